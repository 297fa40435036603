#!/bin/bash
# round 2, visit 11: after creating every scratch set's streams up front: head of the timed region at 20 / 200 steps, 4 / 6 / 8 streams
mkdir -p gpurun_out
show() {
python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/$1.json") if l.startswith("{")][-1])
    d=j["timing_diag"]
    print("$1 steps %d value %.1fM ms/step %.3f e2e %.1fM | gaps med %.3f max %.3f first %.3f"%(j["steps"],j["value"]/1e6,j["ms_per_step"],j["e2e"]["value"]/1e6,d["step_completion_gap_ms"]["median"],d["step_completion_gap_ms"]["max"],d["first_step_done_ms"]))
except Exception as ex: print("$1 failed", ex)
PY
}
for rep in a b c; do
  SBV_BENCH_LANES=4 timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/v11_l4_k20$rep.json 2> gpurun_out/v11.err; show v11_l4_k20$rep
done
for L in 6 8; do
  SBV_BENCH_LANES=$L timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/v11_l${L}_k20.json 2> gpurun_out/v11.err; show v11_l${L}_k20
done
for L in 4 6; do
  SBV_BENCH_LANES=$L timeout 300 python bench.py --steps 200 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/v11_l${L}_k200.json 2> gpurun_out/v11.err; show v11_l${L}_k200
done
