#!/bin/bash
# round 2, visit 10: the driver times 20 steps: how many streams should the device-timed steps rotate over? (fill / drain of the pipeline)
mkdir -p gpurun_out
show() {
python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/$1.json") if l.startswith("{")][-1])
    d=j["timing_diag"]
    print("$1 steps %d value %.1fM ms/step %.3f e2e %.1fM | gaps med %.3f max %.3f first %.3f"%(j["steps"],j["value"]/1e6,j["ms_per_step"],j["e2e"]["value"]/1e6,d["step_completion_gap_ms"]["median"],d["step_completion_gap_ms"]["max"],d["first_step_done_ms"]))
    if j["steps"]<=20: print("    done:", d["step_done_ms"])
except Exception as ex: print("$1 failed", ex)
PY
}
for L in 4 6 8; do
  for K in 20 200; do
    SBV_BENCH_LANES=$L timeout 300 python bench.py --steps $K --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/v10_l${L}_k$K.json 2> gpurun_out/v10_l${L}_k$K.err; show v10_l${L}_k$K
  done
done
SBV_BENCH_LANES=4 timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/v10_l4_k20b.json 2>/dev/null; show v10_l4_k20b
