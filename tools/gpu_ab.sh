#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 200 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
b=json.loads(sys.stdin.read())
print('value',b['value'],'e2e',b['e2e']['value'],b['e2e']['single_caller_value'],'lat',b['step_latency_ms'],'roof',b['roofline']['frac'],'reg',b['registered_keys']['value'],b['registered_keys']['e2e'])"
bash tools/gpu_qb.sh 2>&1 | tail -22
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_extras.csv python tools/bench_extras.py > /dev/null 2>&1
grep -E "k_sha256|k_quorum" gpurun_out/launches_extras.csv | awk -F'","' '{print $5, $NF}' | head -12
