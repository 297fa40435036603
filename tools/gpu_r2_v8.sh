#!/bin/bash
# round 2, visit 8: chunked upload of large host-buffer batches (stage_and_verify): parity suite, C3 against the chunk size, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python tools/c3_chunk_sweep.py > gpurun_out/v8_c3_chunk_sweep.txt 2> gpurun_out/v8_c3.err; cat gpurun_out/v8_c3_chunk_sweep.txt; tail -2 gpurun_out/v8_c3.err
timeout 900 python bench.py > gpurun_out/v8_bench.json 2> gpurun_out/v8_bench.err; tail -3 gpurun_out/v8_bench.err
python - <<'PY'
import json
for f in ("v8_bench",):
    try:
        j=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1])
        print(f, "value %.1fM ms/step %.3f e2e %.1fM single %.1fM lat %.3f"%(j["value"]/1e6,j["ms_per_step"],j["e2e"]["value"]/1e6,j["e2e"]["single_caller_value"]/1e6,j["step_latency_ms"]))
        ex=j["extras"]
        print("   c3", {k:v for k,v in ex["c3_sha256_verify_1m"].items() if k in ("value","e2e_s","bit_exact_vs_oracle","h2d_gbs")})
        print("   c4", {k:v for k,v in ex["c4_quorum_stream"].items() if k in ("value","e2e_s","bit_exact_vs_oracle")})
        print("   c5", {k:v for k,v in ex["c5_mixed_curve_64k"].items() if k in ("value","bit_exact_vs_oracle")})
    except Exception as ex: print(f, "failed", ex)
PY
