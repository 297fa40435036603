#!/bin/bash
# round 2, visit 13 (final single-GPU pass): parity suite, smoke, the driver's command (20 steps), the default bench (200 steps),
# the reference arm, the ncu launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/v13_bench_k20.json 2> gpurun_out/v13_bench_k20.err; tail -2 gpurun_out/v13_bench_k20.err
timeout 900 python bench.py > gpurun_out/v13_bench.json 2> gpurun_out/v13_bench.err; tail -2 gpurun_out/v13_bench.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/v13_bench_ref.json 2> gpurun_out/v13_bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv --log-file gpurun_out/v13_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/v13_ncu_bench.log 2>&1
python - <<'PY'
import json
for f in ("v13_bench_k20","v13_bench","v13_bench_ref"):
    try:
        j=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1])
        if j.get("impl")=="reference": print(f, "value %.3fM"%(j["value"]/1e6), j.get("cpu_baseline")); continue
        print(f, "steps %d value %.1fM ms/step %.3f e2e %.1fM single %.1fM lat %.3f launches %d"%(j["steps"],j["value"]/1e6,j["ms_per_step"],j["e2e"]["value"]/1e6,j["e2e"]["single_caller_value"]/1e6,j["step_latency_ms"],j["gpu_launches"]))
        ex=j.get("extras")
        if ex: print("   c3 %.1fM c4 %.1fM c5 %.1fM"%(ex["c3_sha256_verify_1m"]["value"]/1e6, ex["c4_quorum_stream"]["value"]/1e6, ex["c5_mixed_curve_64k"]["value"]/1e6), [ex[k]["bit_exact_vs_oracle"] for k in ("c3_sha256_verify_1m","c4_quorum_stream","c5_mixed_curve_64k")], "reg %.1fM sim %s"%(j["registered_keys"]["value"]/1e6, str(j.get("sim"))[:120]))
    except Exception as ex: print(f, "failed", ex)
PY
