#!/bin/bash
# round 2, visit 2: full parity suite, bench, kernel A/Bs, microbenchmarks, ncu capture of the dominant kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py --steps 50 > gpurun_out/v2_bench.json 2> gpurun_out/v2_bench.err; tail -3 gpurun_out/v2_bench.err; cut -c1-250 gpurun_out/v2_bench.json
for v in 1 2; do SBV_KT_VARIANT=$v timeout 300 python bench.py --steps 40 --no-extras --no-cpu-baseline > gpurun_out/v2_bench_kt$v.json 2> gpurun_out/v2_bench_kt$v.err; cut -c1-200 gpurun_out/v2_bench_kt$v.json; done
SBV_KT_BASES_CALL=1 timeout 300 python bench.py --steps 40 --no-extras --no-cpu-baseline > gpurun_out/v2_bench_basescall.json 2> gpurun_out/v2_bench_basescall.err; cut -c1-200 gpurun_out/v2_bench_basescall.json
./tools/ubench > gpurun_out/v2_ubench.txt 2>&1; tail -8 gpurun_out/v2_ubench.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/v2_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/v2_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_verify_kt -s 4 -c 1 -f -o gpurun_out/v2_prof_kt python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/v2_ncu_full.log 2>&1
ls -la gpurun_out | grep v2_
