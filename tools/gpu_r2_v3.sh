#!/bin/bash
# round 2, visit 3: parity suite, both bench arms with the final defaults, compute-sanitizer over every kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 900 python bench.py > gpurun_out/v3_bench.json 2> gpurun_out/v3_bench.err; tail -3 gpurun_out/v3_bench.err; cut -c1-250 gpurun_out/v3_bench.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/v3_bench_ref.json 2>> gpurun_out/v3_bench.err; cut -c1-200 gpurun_out/v3_bench_ref.json
cat > /tmp/san.py <<'PY'
import sys, os; sys.path.insert(0, os.getcwd())
import numpy as np, oracle
from oracle import corpus
import consensus_b200 as sbv
e = sbv.Engine(n_devices=1)
for curve, n in [(0, 600), (1, 260)]:
    b = corpus.make_batch(curve, n=n, K=4, seed=3, corrupt_rate=3)      # 4 keys x many signatures: grouped path + generic path for the corrupted keys
    want = oracle.verify_batch(curve, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
    assert (e.verify_batch(curve, b["r"], b["s"], b["qx"], b["qy"], b["digest"]) == want).all()
    L = 32 if curve == 0 else 48
    e.set_keys(np.full(4, curve, np.uint8), b["keys"].reshape(4, 2, L))
    wk = oracle.verify_batch(curve, b["r"], b["s"], np.ascontiguousarray(b["keys"][b["key_idx"], :L]), np.ascontiguousarray(b["keys"][b["key_idx"], L:]), b["digest"])
    assert (e.verify_registered(curve, b["key_idx"], b["r"], b["s"], b["digest"]) == wk).all()
    assert (e.verify_registered(curve, b["key_idx"][:40], b["r"][:40], b["s"][:40], b["digest"][:40]) == wk[:40]).all()   # warp-per-signature kernel
msgs, off = corpus.make_requests(200, seed=5, fixed_len=None, lo=1, hi=900)
assert (e.sha256_batch(msgs, off) == oracle.sha256_batch(msgs, off)).all()
cnt, reached = e.quorum([0,0,0,1,1], [1,2,3,1,1], [1,2,3,1,1], [1,1,1,1,1], [1,1,0,1,1], 2, 2)
assert cnt.tolist() == [2, 1]
b = corpus.make_batch(0, n=300, K=3, seed=9, corrupt_rate=4)
inst = np.repeat(np.arange(20, dtype=np.uint32), 15); snd = np.tile(np.arange(1, 16, dtype=np.uint16), 20)
ok, cnt, reached = e.verify_quorum(0, b["r"], b["s"], b["qx"], b["qy"], b["digest"], inst, snd, snd, np.ones(300, np.uint8), 20, 10)
assert (ok == oracle.verify_batch(0, b["r"], b["s"], b["qx"], b["qy"], b["digest"])).all()
e.close(); print("sanitizer workload ok")
PY
for tool in memcheck racecheck; do
  echo "== $tool"; timeout 1500 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py 2>&1 | tail -6 | tee gpurun_out/v3_sanitizer_$tool.txt
done
