#!/bin/bash
# round 2, visit 7: four-lane doubling chain (k_kt_bases4) and the 16-bit P-384 comb: parity suite first, then A/B + bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
SBV_KT_BASES=1 timeout 300 python bench.py --steps 50 --no-extras --no-cpu-baseline > gpurun_out/v7_bases1.json 2> gpurun_out/v7_bases1.err
timeout 900 python bench.py > gpurun_out/v7_bench.json 2> gpurun_out/v7_bench.err; tail -3 gpurun_out/v7_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/v7_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/v7_ncu_bench.log 2>&1
python - <<'PY'
import json
for f in ("v7_bases1","v7_bench"):
    try:
        j=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1])
        print(f, "value %.1fM ms/step %.3f e2e %.1fM single %.1fM lat %.3f"%(j["value"]/1e6,j["ms_per_step"],j["e2e"]["value"]/1e6,j["e2e"]["single_caller_value"]/1e6,j["step_latency_ms"]))
        if "extras" in j: print("   c5", {k:v for k,v in j["extras"]["c5_mixed_curve_64k"].items() if k in ("value","bit_exact_vs_oracle","roofline_frac_canonical")}, "c4 %.1fM"%(j["extras"]["c4_quorum_stream"]["value"]/1e6))
    except Exception as ex: print(f, "failed", ex)
PY
grep -E "k_kt_bases|k_gtable" gpurun_out/v7_launches.csv | cut -d, -f5,12- | sort | uniq -c | sort -rn | head -6 | cut -c1-150
