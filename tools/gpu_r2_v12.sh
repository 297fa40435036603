#!/bin/bash
# round 2, visit 12: do the per-step completion events (timing_diag) cost throughput?  same box, alternating
mkdir -p gpurun_out
show() {
python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/$1.json") if l.startswith("{")][-1])
    print("$1 steps %d value %.1fM ms/step %.3f e2e %.1fM"%(j["steps"],j["value"]/1e6,j["ms_per_step"],j["e2e"]["value"]/1e6))
except Exception as ex: print("$1 failed", ex)
PY
}
for rep in a b; do
  for D in 0 1; do
    SBV_BENCH_DIAG=$D timeout 300 python bench.py --steps 200 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/v12_d${D}_$rep.json 2> gpurun_out/v12.err; show v12_d${D}_$rep
  done
done
SBV_BENCH_DIAG=0 SBV_TAB_PRIORITY=0 timeout 300 python bench.py --steps 200 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/v12_d0_notab.json 2> gpurun_out/v12.err; show v12_d0_notab
