#!/bin/bash
# round 2, visit 6 (final single-GPU evidence): parity suite, bench (both arms), launch list, ncu --set full of the two
# verification kernels, sanitizer
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/v6_bench.json 2> gpurun_out/v6_bench.err; tail -3 gpurun_out/v6_bench.err; cut -c1-200 gpurun_out/v6_bench.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/v6_bench_ref.json 2>> gpurun_out/v6_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 170 --csv --log-file gpurun_out/v6_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/v6_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_verify_kt|k_gpart" -s 8 -c 2 -f -o gpurun_out/v6_prof_kt python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/v6_ncu_full.log 2>&1
ls -la gpurun_out | grep v6_
