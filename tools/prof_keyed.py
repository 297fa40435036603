#!/usr/bin/env python3
"""Workload for ncu captures of the registered-key kernels and k_prep (see tools/gpu_prof_keyed.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from oracle import corpus
import consensus_b200 as sbv
b = corpus.make_batch(0, n=65536, K=1024, seed=1)
e = sbv.Engine(n_devices=1)
e.set_keys(np.zeros(1024, np.uint8), b["keys"].reshape(1024, 2, 32))
for _ in range(4):
    ok = e.verify_registered(0, b["key_idx"], b["r"], b["s"], b["digest"])
ok16 = e.verify_registered(0, b["key_idx"][:16], b["r"][:16], b["s"][:16], b["digest"][:16])
print(int(ok.sum()), int(ok16.sum()))
e.close()
