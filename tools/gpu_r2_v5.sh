#!/bin/bash
# round 2, visit 5: G-part split A/B, four streams, short (driver-like) and default step counts
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
for tag in default gsplit0; do
  [ $tag = gsplit0 ] && export SBV_GSPLIT=0
  timeout 400 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/v5_${tag}_s20.json 2> gpurun_out/v5_${tag}_s20.err
  timeout 400 python bench.py --no-extras --no-cpu-baseline > gpurun_out/v5_${tag}_s200.json 2> gpurun_out/v5_${tag}_s200.err
done
unset SBV_GSPLIT
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/v5_*.json")):
    try:
        j=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value %.1fM ms/step %.3f e2e %.1fM single %.1fM lat %.3f kernel %.3f"%(j["value"]/1e6,j["ms_per_step"],j["e2e"]["value"]/1e6,j["e2e"]["single_caller_value"]/1e6,j["step_latency_ms"],j["roofline"]["kernel_ms"]))
    except Exception as ex: print(f, ex)
PY
