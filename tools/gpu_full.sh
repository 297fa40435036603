#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -18
timeout 600 python tools/bench_extras.py 2>&1 | tail -50
./consensus_b200/host/sim 1000 100 1
./consensus_b200/host/sim 1000 1 1
