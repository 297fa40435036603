#!/bin/bash
# round 2, visit 1: parity first, then the bench and A/B variants of the new pipeline
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | head -1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 50 > gpurun_out/v1_bench.json 2> gpurun_out/v1_bench.err; tail -3 gpurun_out/v1_bench.err; cut -c1-600 gpurun_out/v1_bench.json
SBV_KT_RELAXED=1 timeout 300 python bench.py --steps 50 --no-extras --no-cpu-baseline > gpurun_out/v1_bench_relaxed.json 2> gpurun_out/v1_bench_relaxed.err; cut -c1-300 gpurun_out/v1_bench_relaxed.json
SBV_GROUP_THRESHOLD=0 timeout 300 python bench.py --steps 30 --no-extras --no-cpu-baseline > gpurun_out/v1_bench_generic.json 2> gpurun_out/v1_bench_generic.err; cut -c1-300 gpurun_out/v1_bench_generic.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/v1_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/v1_ncu_bench.log 2>&1
tail -30 gpurun_out/v1_launches.csv | cut -d, -f5,12- | cut -c1-160
