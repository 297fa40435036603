#!/usr/bin/env python3
"""Turns gpurun_out/{launches.csv,prof_verify.ncu-rep,bench.json} into tracked summaries under profiles/."""
import csv, json, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go = os.path.join(root, "gpurun_out"); out = os.path.join(root, "profiles")
lines = []
rows = [r for r in csv.reader(open(os.path.join(go, "launches.csv"))) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = {}
for r in rows[1:]:
    agg.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")))
step = {k: v for k, v in agg.items() if "k_verify" in k or "k_prep" in k}
tot = sum(sum(v) for v in step.values())
lines.append(f"# ncu launch list ({tag}): ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 3 --warmup 3")
lines.append("# per-launch times are cold-cache and serialised: compare SHARES of the step, not absolutes\n")
lines.append(f"{'kernel':62s} {'launches':>8s} {'avg_us':>10s} {'share_of_step':>14s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    share = f"{sum(v)/tot:13.1%}" if k in step else "   (not in step)"
    lines.append(f"{k[:62]:62s} {len(v):8d} {sum(v)/len(v)/1e3:10.1f} {share}")
open(os.path.join(out, f"{tag}_launches.txt"), "w").write("\n".join(lines) + "\n")
raw = subprocess.run(["ncu", "-i", os.path.join(go, "prof_verify.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines())); h, u, v = rr[0], rr[1], rr[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled", "sm__cycles_elapsed.avg"]
sel = [f"# ncu --set full --clock-control none -k regex:k_verify ({tag}); kernel: {v[h.index('Kernel Name')] if 'Kernel Name' in h else 'k_verify'}", ""]
for a, b, c in zip(h, u, v):
    if any(w in a for w in want) and ".max" not in a and ".min" not in a and "pcsamp" not in a and "per_second" not in a:
        sel.append(f"{a:92s} {c:>18s} {b}")
open(os.path.join(out, f"{tag}_k_verify_ncu.txt"), "w").write("\n".join(sel) + "\n")
for f in ("bench.json", "bench_ref.json"):
    p = os.path.join(go, f)
    if os.path.exists(p):
        open(os.path.join(out, f"{tag}_{f}"), "w").write(open(p).read())
print("wrote profiles/", tag)
