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
for w, blk in [(3, 64)]:
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
e = sbv.Engine(n_devices=1)
t0 = time.time(); e.set_keys(np.zeros(1024, np.uint8), b["keys"].reshape(1024, 2, 32)); res["set_keys_1024_s"] = time.time() - t0
slot = torch.from_numpy(b["key_idx"].astype(np.int32)).to(dev)
runk = lambda: e.verify_registered_device(P256, n, slot.data_ptr(), t["r"].data_ptr(), t["s"].data_ptr(), t["digest"].data_ptr(), 32, ok.data_ptr(), stream=st)
for _ in range(3): runk()
torch.cuda.synchronize()
want_k = oracle.verify_batch(P256, b["r"], b["s"], np.ascontiguousarray(b["keys"][b["key_idx"], :32]), np.ascontiguousarray(b["keys"][b["key_idx"], 32:]), b["digest"])
assert (ok.cpu().numpy() == want_k).all(), "parity keyed"
a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): runk()
c.record(); torch.cuda.synchronize()
ms = a.elapsed_time(c) / 10
res["registered_1024keys"] = {"ms": ms, "Mverif_s": n / ms / 1e3}
t0 = time.time(); got = e.verify_registered(P256, b["key_idx"], b["r"], b["s"], b["digest"]); res["registered_1024keys"]["e2e_pageable_ms"] = (time.time() - t0) * 1e3
e.set_keys(np.zeros(16, np.uint8), b["keys"].reshape(1024, 2, 32)[:16])
slot16 = torch.from_numpy((b["key_idx"] % 16).astype(np.int32)).to(dev)
runk = lambda: e.verify_registered_device(P256, n, slot16.data_ptr(), t["r"].data_ptr(), t["s"].data_ptr(), t["digest"].data_ptr(), 32, ok.data_ptr(), stream=st)
for _ in range(3): runk()
a.record()
for _ in range(10): runk()
c.record(); torch.cuda.synchronize()
res["registered_16keys"] = {"ms": a.elapsed_time(c) / 10, "Mverif_s": n / (a.elapsed_time(c) / 10) / 1e3}
e.close()
print(json.dumps(res, indent=1))
# ---- A/B at full batch: one signature per warp vs one per thread (registered-key path, 64K) ----
for limit in (100000, 0):
    os.environ["SBV_KEYED_WARP_LIMIT"] = str(limit)
    e = sbv.Engine(n_devices=1)
    e.set_keys(np.zeros(16, np.uint8), b["keys"].reshape(1024, 2, 32)[:16])
    runk = lambda: e.verify_registered_device(P256, n, slot16.data_ptr(), t["r"].data_ptr(), t["s"].data_ptr(), t["digest"].data_ptr(), 32, ok.data_ptr(), stream=st)
    for _ in range(3): runk()
    a.record()
    for _ in range(5): runk()
    c.record(); torch.cuda.synchronize()
    print(json.dumps({("warp_per_signature" if limit else "thread_per_signature") + "_64k_16keys_Mverif_s": n / (a.elapsed_time(c) / 5) / 1e3}))
    e.close()
# ---- small-batch latency of the registered-key path: warp-per-signature vs thread-per-signature ----
lat = {}
for limit in (2048, 0):
    os.environ["SBV_KEYED_WARP_LIMIT"] = str(limit)
    e = sbv.Engine(n_devices=1)
    e.set_keys(np.zeros(16, np.uint8), b["keys"].reshape(1024, 2, 32)[:16])
    for m in (1, 16, 128, 1024, 2048):
        sl = (b["key_idx"][:m] % 16).astype(np.uint32)
        f = lambda: e.verify_registered(P256, sl, b["r"][:m], b["s"][:m], b["digest"][:m])
        for _ in range(5): f()
        t0 = time.perf_counter()
        for _ in range(50): f()
        lat[f"limit{limit}_n{m}_us"] = (time.perf_counter() - t0) / 50 * 1e6
    e.close()
print(json.dumps(lat, indent=1))
