#!/bin/bash
# e2e at N ranks: number of caller threads and NCCL channel cap (A/B; each run prints value / e2e / single-caller e2e)
N=${1:-4}
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-extras > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/ab_$tag.json") if l.startswith("{")][-1])
    print("$tag N=$N value %.1fM e2e %.1fM (callers %d) single %.1fM"%(j["value"]/1e6,j["e2e"]["value"]/1e6,j["e2e"]["callers"],j["e2e"]["single_caller_value"]/1e6))
except Exception as ex: print("$tag failed", ex)
PY
}
run t4_ch1 SBV_BENCH_E2E_THREADS=4
run t3_ch1 SBV_BENCH_E2E_THREADS=3
run t3_chdef SBV_BENCH_E2E_THREADS=3 NCCL_MAX_NCHANNELS=32
run t6_ch1 SBV_BENCH_E2E_THREADS=6
