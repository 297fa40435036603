#!/usr/bin/env python3
"""Lone-call latency of sbv_verify_batch (pinned host buffers, one caller) over the batch size, with the key grouping on
and off: decides SBV_GROUP_MIN_BATCH.  Prints a table (gpurun_out/latency_sweep.txt)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from oracle import corpus
import consensus_b200 as sbv

b = corpus.make_batch(oracle.P256, n=65536, K=1024, seed=1)
fields = ("r", "s", "qx", "qy", "digest")
host = {k: torch.from_numpy(np.ascontiguousarray(b[k])).pin_memory() for k in fields}
ok = torch.zeros(65536, dtype=torch.uint8).pin_memory()
rows = []
for thr in (16, 0):
    os.environ["SBV_GROUP_THRESHOLD"] = str(thr)
    eng = sbv.Engine(n_devices=1)
    for n in (256, 1024, 4096, 8192, 16384, 32768, 65536):
        call = lambda: eng.verify_batch_ptr(sbv.P256, n, *(host[k].data_ptr() for k in fields), 32, ok.data_ptr())
        for _ in range(6): call()
        ts = []
        for _ in range(15):
            t0 = time.perf_counter(); call(); ts.append(time.perf_counter() - t0)
        ts.sort()
        rows.append((thr, n, ts[len(ts) // 2] * 1e3))
    eng.close()
out = ["# lone sbv_verify_batch call from pinned host memory (H2D + pipeline + D2H), median of 15, ms; keys: n/64 distinct (C2 corpus prefix)",
       f"{'n':>7s} {'grouping on (T=16)':>20s} {'grouping off':>14s}"]
for n in sorted({r[1] for r in rows}):
    a = [r[2] for r in rows if r[1] == n and r[0] == 16][0]; c = [r[2] for r in rows if r[1] == n and r[0] == 0][0]
    out.append(f"{n:7d} {a:20.3f} {c:14.3f}")
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/latency_sweep.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out))
