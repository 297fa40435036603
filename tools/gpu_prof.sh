#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_verify -s 3 -c 1 -f -o gpurun_out/prof_verify python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
