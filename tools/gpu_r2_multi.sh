#!/bin/bash
# N-GPU visit: the driver's own launch of bench.py at N ranks, the same without the high-priority exchange stream (A/B), and
# N=1 on the same box.   usage: gpu_r2_multi.sh N [steps] [tests 0/1] [A/B 0/1]
N=${1:-2}
STEPS=${2:-20}
mkdir -p gpurun_out
nvidia-smi -L | head -8
if [ "${3:-0}" = "1" ]; then timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -k "multi_device" 2>&1 | tail -4; fi
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $STEPS --warmup 5 > gpurun_out/m_bench_n$N.json 2> gpurun_out/m_bench_n$N.err; tail -5 gpurun_out/m_bench_n$N.err; cut -c1-300 gpurun_out/m_bench_n$N.json
if [ "${4:-0}" = "1" ]; then SBV_GATHER_PRIORITY=0 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps $STEPS --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/m_bench_n${N}_plain.json 2> gpurun_out/m_bench_n${N}_plain.err; fi
timeout 240 python bench.py --gpus 1 --steps $STEPS --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/m_bench_n1_of$N.json 2>/dev/null
python - <<PY
import json
last=lambda f: json.loads([l for l in open(f) if l.startswith("{")][-1])
a=last("gpurun_out/m_bench_n$N.json"); b=last("gpurun_out/m_bench_n1_of$N.json")
print("N=$N value %.1fM e2e %.1fM (callers %d, single %.1fM) | N=1 value %.1fM e2e %.1fM | efficiency value %.3f e2e %.3f" % (a["value"]/1e6, a["e2e"]["value"]/1e6, a["e2e"]["callers"], a["e2e"]["single_caller_value"]/1e6, b["value"]/1e6, b["e2e"]["value"]/1e6, a["value"]/b["value"]/$N, a["e2e"]["value"]/b["e2e"]["value"]/$N))
print("c4:", {k:v for k,v in a.get("extras",{}).get("c4_quorum_stream",{}).items() if k in ("value","e2e_s","bit_exact_vs_oracle","n_gpus")})
try:
    p=last("gpurun_out/m_bench_n${N}_plain.json")
    print("exchange on the step's own stream (SBV_GATHER_PRIORITY=0): value %.1fM e2e %.1fM" % (p["value"]/1e6, p["e2e"]["value"]/1e6))
except Exception as ex: print("plain run failed", ex)
PY
