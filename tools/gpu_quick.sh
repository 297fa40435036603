#!/bin/bash
# quick GPU visit: parity tests + tunable sweep
mkdir -p gpurun_out
cat /sys/fs/cgroup/cpu.max 2>/dev/null; nproc
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/quick_bench.py 2>&1 | tail -30
