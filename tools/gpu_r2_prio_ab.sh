#!/bin/bash
# N ranks: the per-step pack + all-gather on a high-priority stream (SBV_GATHER_PRIORITY=1) against the step's own stream
N=${1:-4}
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/prio_$tag.json 2> gpurun_out/prio_$tag.err
  python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/prio_$tag.json") if l.startswith("{")][-1])
    print("$tag N=$N value %.1fM e2e %.1fM (callers %d) single %.1fM"%(j["value"]/1e6,j["e2e"]["value"]/1e6,j["e2e"]["callers"],j["e2e"]["single_caller_value"]/1e6))
except Exception as ex: print("$tag failed", ex); print(open("gpurun_out/prio_$tag.err").read()[-1500:])
PY
}
run plain SBV_GATHER_PRIORITY=0
run hi SBV_GATHER_PRIORITY=1
run hi_t4 SBV_GATHER_PRIORITY=1 SBV_BENCH_E2E_THREADS=4
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/prio_n1.json 2>/dev/null
python - <<PY
import json
j=json.loads([l for l in open("gpurun_out/prio_n1.json") if l.startswith("{")][-1])
print("N=1 same box: value %.1fM e2e %.1fM"%(j["value"]/1e6,j["e2e"]["value"]/1e6))
PY
