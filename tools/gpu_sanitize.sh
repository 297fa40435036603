#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over a small run of every kernel
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0, '/root/repo')
import numpy as np, oracle
from oracle import corpus
import consensus_b200 as sbv
e = sbv.Engine(n_devices=1)
for curve, n in [(0, 300), (1, 130)]:
    b = corpus.make_batch(curve, n=n, K=4, seed=3, corrupt_rate=3)
    want = oracle.verify_batch(curve, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
    assert (e.verify_batch(curve, b["r"], b["s"], b["qx"], b["qy"], b["digest"]) == want).all()
    L = 32 if curve == 0 else 48
    e.set_keys(np.full(4, curve, np.uint8), b["keys"].reshape(4, 2, L))
    wk = oracle.verify_batch(curve, b["r"], b["s"], np.ascontiguousarray(b["keys"][b["key_idx"], :L]), np.ascontiguousarray(b["keys"][b["key_idx"], L:]), b["digest"])
    assert (e.verify_registered(curve, b["key_idx"], b["r"], b["s"], b["digest"]) == wk).all()
msgs, off = corpus.make_requests(200, seed=5, fixed_len=None, lo=1, hi=900)
assert (e.sha256_batch(msgs, off) == oracle.sha256_batch(msgs, off)).all()
cnt, reached = e.quorum([0,0,0,1,1], [1,2,3,1,1], [1,2,3,1,1], [1,1,1,1,1], [1,1,0,1,1], 2, 2)
assert cnt.tolist() == [2, 1]
e.close(); print("sanitizer workload ok")
PY
cd /root/repo
for tool in memcheck racecheck; do
  echo "== $tool"; timeout 1500 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py 2>&1 | tail -6 | tee gpurun_out/sanitizer_$tool.txt
done
