#!/bin/bash
# One GPU visit: smoke, bench (both arms), ncu launch list + full capture of the dominant kernel.
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_verify -s 3 -c 1 -f -o gpurun_out/prof_verify python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
