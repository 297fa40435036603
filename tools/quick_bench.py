#!/usr/bin/env python3
"""First-contact GPU numbers: MAD-rate probe + kernel-only verify throughput for the tunables."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from oracle import corpus, P256, P384
import consensus_b200 as sbv

n = int(os.environ.get("QB_N", 65536))
b = corpus.make_batch(P256, n=n, K=1024, seed=1)
want = oracle.verify_batch(P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
dev = torch.device("cuda:0")
t = {k: torch.from_numpy(b[k]).to(dev) for k in ("r", "s", "qx", "qy", "digest")}
ok = torch.zeros(n, dtype=torch.uint8, device=dev)
res = {}
for w, blk in [(3, 64), (3, 128), (4, 128)]:
    os.environ["SBV_P256_W"] = str(w); os.environ["SBV_P256_BLOCK"] = str(blk)
    e = sbv.Engine(n_devices=1)
    if "mad" not in res:
        res["mad_rate_TMAC_s"] = e.probe_mad_rate() / 1e12
    st = torch.cuda.current_stream().cuda_stream
    run = lambda: e.verify_batch_device(P256, n, t["r"].data_ptr(), t["s"].data_ptr(), t["qx"].data_ptr(), t["qy"].data_ptr(),
                                        t["digest"].data_ptr(), 32, ok.data_ptr(), stream=st)
    for _ in range(3): run()
    torch.cuda.synchronize()
    assert (ok.cpu().numpy() == want).all(), "parity"
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): run()
    c.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(c) / 10
    res[f"W{w}_B{blk}"] = {"ms": ms, "Mverif_s": n / ms / 1e3}
    t0 = time.time(); got = e.verify_batch(P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"]); t1 = time.time()
    res[f"W{w}_B{blk}"]["e2e_pageable_ms"] = (t1 - t0) * 1e3
    e.close()
print(json.dumps(res, indent=1))
