#!/bin/bash
# round 2, visit 14: spread of the driver's command (20 timed steps) on one box
mkdir -p gpurun_out
for rep in a b c; do
  timeout 60 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/v14_k20$rep.json 2>/dev/null
  python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/v14_k20$rep.json") if l.startswith("{")][-1])
    print("k20$rep value %.1fM ms/step %.3f e2e %.1fM"%(j["value"]/1e6,j["ms_per_step"],j["e2e"]["value"]/1e6))
except Exception as ex: print("failed", ex)
PY
done
