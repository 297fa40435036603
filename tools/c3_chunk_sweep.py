"""C3 (sbv_hash_verify_batch, 1,048,576 requests of 256 B from pinned host memory) against the chunk size of the chunked
upload (SBV_CHUNK_ITEMS; 0 = the whole shard is uploaded before anything runs).  Verdicts checked against the oracle.
Run on a GPU box: python tools/c3_chunk_sweep.py > gpurun_out/c3_chunk_sweep.txt"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from oracle import corpus

BATCH, T16 = 65536, 16
msgs1, off1 = corpus.make_requests(BATCH, seed=5, fixed_len=256)
dig1 = oracle.sha256_batch(msgs1, off1)
d, kxy = corpus.make_keys(oracle.P256, 4096, seed=71)
kidx = (np.arange(BATCH) % 4096).astype(np.uint32)
r1, s1 = oracle.sign_batch(oracle.P256, d, kidx, dig1, corpus._blocks(73, BATCH, 32, b"k"))
msgs1 = msgs1.copy()
msgs1[np.nonzero((np.arange(BATCH) % 16) == 5)[0] * 256 + 17] ^= 0x40
want = np.tile(oracle.verify_batch(oracle.P256, r1, s1, kxy[kidx, :32].copy(), kxy[kidx, 32:].copy(), oracle.sha256_batch(msgs1, off1)), T16)
n3 = BATCH * T16
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
rep = lambda a: np.ascontiguousarray(np.tile(a, (T16, 1)))
M, OFF = pin(np.tile(msgs1, T16)), pin(np.arange(n3 + 1, dtype=np.uint64) * 256)
R, S, QX, QY = pin(rep(r1)), pin(rep(s1)), pin(rep(kxy[kidx, :32])), pin(rep(kxy[kidx, 32:]))
ok3 = pin(np.zeros(n3, np.uint8))
vp = ctypes.c_void_p

import consensus_b200 as sbv

print("# chunk_items  ms/call  M requests/s  bit_exact")
for items in [0, 32768, 65536, 131072, 262144, 524288]:
    os.environ["SBV_CHUNK_ITEMS"] = str(items)
    eng = sbv.Engine(n_devices=1)

    def c3():
        eng._check(eng._lib.sbv_hash_verify_batch(eng._h, ctypes.c_uint8(0), ctypes.c_size_t(n3), vp(M.data_ptr()), vp(OFF.data_ptr()), vp(R.data_ptr()),
                                                  vp(S.data_ptr()), vp(QX.data_ptr()), vp(QY.data_ptr()), None, vp(ok3.data_ptr())), "sbv_hash_verify_batch")
    for _ in range(3):
        c3()
    best = 1e9
    for _ in range(4):
        ok3.zero_()
        t0 = time.perf_counter()
        c3()
        best = min(best, time.perf_counter() - t0)
    print(f"{items:10d}  {best * 1e3:8.2f}  {n3 / best / 1e6:8.1f}  {bool(np.array_equal(ok3.numpy(), want))}", flush=True)
    eng.close()
