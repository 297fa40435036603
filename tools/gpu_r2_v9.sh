#!/bin/bash
# round 2, visit 9: parity suite on the final chunk rule, table-construction streams at high priority (A/B), bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
show() {
python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/$1.json") if l.startswith("{")][-1])
    print("$1 value %.1fM ms/step %.3f e2e %.1fM single %.1fM lat %.3f"%(j["value"]/1e6,j["ms_per_step"],j["e2e"]["value"]/1e6,j["e2e"]["single_caller_value"]/1e6,j["step_latency_ms"]))
    ex=j.get("extras")
    if ex:
        print("   c3 %.1fM c4 %.1fM c5 %.1fM"%(ex["c3_sha256_verify_1m"]["value"]/1e6, ex["c4_quorum_stream"]["value"]/1e6, ex["c5_mixed_curve_64k"]["value"]/1e6), [ex[k]["bit_exact_vs_oracle"] for k in ("c3_sha256_verify_1m","c4_quorum_stream","c5_mixed_curve_64k")])
except Exception as ex: print("$1 failed", ex)
PY
}
SBV_TAB_PRIORITY=1 timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/v9_tabprio.json 2> gpurun_out/v9_tabprio.err; show v9_tabprio
timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/v9_plain.json 2> gpurun_out/v9_plain.err; show v9_plain
SBV_TAB_PRIORITY=1 timeout 300 python bench.py --no-extras --no-cpu-baseline > gpurun_out/v9_tabprio2.json 2> gpurun_out/v9_tabprio2.err; show v9_tabprio2
timeout 900 python bench.py > gpurun_out/v9_bench.json 2> gpurun_out/v9_bench.err; tail -2 gpurun_out/v9_bench.err; show v9_bench
