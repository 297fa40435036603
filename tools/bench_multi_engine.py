#!/usr/bin/env python3
"""Single-process multi-device engine (sbv_create with N devices, NCCL gather inside libsbv.so): e2e rate
through sbv_verify_batch with pinned host buffers.  Run under gpurun --gpus N."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from oracle import corpus
import consensus_b200 as sbv

G = torch.cuda.device_count()
n = 65536 * G
b = corpus.make_batch(0, n=65536, K=1024, seed=1)
want1 = oracle.verify_batch(0, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
rep = lambda a: torch.from_numpy(np.ascontiguousarray(np.tile(a, (G, 1)))).pin_memory()
h = {k: rep(b[k]) for k in ("r", "s", "qx", "qy", "digest")}
ok = torch.zeros(n, dtype=torch.uint8).pin_memory()
res = {"devices": G, "signatures_per_call": n}
for devs in sorted({1, G}):
    e = sbv.Engine(n_devices=devs)
    f = lambda: e.verify_batch_ptr(0, n, h["r"].data_ptr(), h["s"].data_ptr(), h["qx"].data_ptr(), h["qy"].data_ptr(), h["digest"].data_ptr(), 32, ok.data_ptr())
    for _ in range(3): f()
    assert (ok.numpy().reshape(G, -1) == want1[None, :]).all()
    t0 = time.perf_counter()
    for _ in range(20): f()
    dt = (time.perf_counter() - t0) / 20
    res[f"engine_{devs}dev_e2e_verifies_per_s"] = n / dt
    e.close()
print(json.dumps(res))
