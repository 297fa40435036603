#!/bin/bash
# round 2, visit 4: keys-first upload + latency sweep (decides SBV_GROUP_MIN_BATCH), bench, round-2 parity tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/latency_sweep.py 2>&1 | tail -12
timeout 900 python bench.py > gpurun_out/v4_bench.json 2> gpurun_out/v4_bench.err; tail -3 gpurun_out/v4_bench.err; cut -c1-200 gpurun_out/v4_bench.json
