#!/usr/bin/env python3
"""Turns the gpurun_out/ artefacts of the round-2 GPU visits (tools/gpu_r2_*.sh) into the tracked summaries under
profiles/: bench lines, launch shares, ncu counters of the dominant kernel, SASS opcode histograms, A/B table."""
import collections, csv, glob, json, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, out = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
visit = sys.argv[2] if len(sys.argv) > 2 else "v2"


def last_json(path):
    return json.loads([l for l in open(path) if l.startswith("{")][-1])


# ---- bench lines ----
for src, dst in [(f"{visit}_bench.json", f"{tag}_bench.json"), (f"{visit}_bench_ref.json", f"{tag}_bench_ref.json")]:
    p = os.path.join(go, src)
    if os.path.exists(p) and os.path.getsize(p):
        open(os.path.join(out, dst), "w").write(json.dumps(last_json(p)) + "\n")

# ---- launch list: every kernel's share of a step ----
p = os.path.join(go, f"{visit}_launches.csv")
if os.path.exists(p):
    rows = list(csv.reader([l for l in open(p) if l.startswith('"')]))
    hdr = rows[0]; ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki]).replace("void sbv::", "").replace("sbv::", "")
        agg.setdefault(name, []).append(float(r[vi].replace(",", "")))
    step = {k: v for k, v in agg.items() if k.startswith("k_") and "gtable" not in k and "mad_probe" not in k}
    tot = sum(sum(v) / len(v) for v in step.values())
    lines = [f"# ncu launch list ({tag}): ncu --metrics gpu__time_duration.sum --clock-control none -c 170  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras",
             "# per-launch times are cold-cache and serialised: compare SHARES of a step (one launch of each kernel), not absolutes", "",
             f"{'kernel':44s} {'launches':>8s} {'avg_us':>10s} {'share_of_step':>14s}"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]) / len(kv[1])):
        share = f"{sum(v)/len(v)/tot:13.1%}" if k in step else "   (not in step)"
        lines.append(f"{k[:44]:44s} {len(v):8d} {sum(v)/len(v)/1e3:10.1f} {share}")
    lines += ["", "# critical path of an ISOLATED step: grouping -> bases -> fill -> inv -> final -> k_verify_kt (k_prep, k_gpart and k_verify_coz run beside the table kernels);",
              "# pipelined steps overlap the latency-bound table kernels of step i+1 with the k_verify_kt of step i"]
    open(os.path.join(out, f"{tag}_launches.txt"), "w").write("\n".join(lines) + "\n")

# ---- ncu --set full of the verification kernels (k_gpart + k_verify_kt) ----
rep = os.path.join(go, f"{visit}_prof_kt.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines())); h, u = rr[0], rr[1]
    want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
            "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__average_warps_issue_stalled", "sm__cycles_elapsed.avg", "sass__inst_executed_local"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    sel = [f"# ncu --set full --clock-control none --import-source on -k 'regex:k_verify_kt|k_gpart' -s 8 -c 2 ({tag}): the two halves of the fixed-base",
           "# verification of one 65,536-signature step (python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras)", ""]
    tr, names = 0.0, []
    for v in rr[2:]:
        if len(v) != len(h):
            continue
        kn = v[h.index("Kernel Name")]
        names.append(kn.split("(")[0])
        sel += [f"## {kn[:100]}", ""]
        vals = {}
        for a, b, c in zip(h, u, v):
            if any(w in a for w in want) and ".max" not in a and ".min" not in a and "pcsamp" not in a and "per_second" not in a and "Triage" not in a:
                sel.append(f"{a:92s} {c:>18s} {b}")
                vals[a] = (c, b)
        sel.append("")
        tr += sum(float(vals[k][0]) * scale[vals[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in vals)
    open(os.path.join(out, f"{tag}_k_verify_kt_ncu.txt"), "w").write("\n".join(sel) + "\n")
    json.dump({"dram_bytes_per_launch": tr, "kernels": names, "source": f"profiles/{tag}_k_verify_kt_ncu.txt (ncu --set full, k_gpart + k_verify_kt of one 65,536-signature step)"},
              open(os.path.join(out, f"{tag}_traffic.json"), "w"))

# ---- SASS opcode histograms (from the in-tree build: nothing GPU-side) ----
obj = os.path.join(root, "consensus_b200", "build", "inst_p256_kt5.o")
if os.path.exists(obj):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    elf = subprocess.run(["cuobjdump", "-elf", obj], capture_output=True, text=True).stdout
    funcs = {x.split("\n", 1)[0]: x for x in re.split(r"\n\s*Function : ", sass)[1:]}
    kname = [k for k in funcs if "k_verify_ktINS_4P256ELi5ELi64ELi7ELb0ELb0" in k][0]
    ins = re.findall(r"/\*([0-9a-f]{4,5})\*/\s+((?:@!?U?P\d+\s+)?)([A-Z0-9_.]+)([^;]*);", funcs[kname])
    syms = {}
    for m in re.finditer(r"^\s*0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+\S+\s+\S+\s+\S+\s+\$" + re.escape(kname) + r"\$\S*?(p256_f(?:mul|sqr)_call)", elf, re.M):
        syms[m.group(3)] = (int(m.group(1), 16), int(m.group(2), 16))

    def hist(lo, hi):
        c = collections.Counter()
        for off, pred, op, rest in ins:
            if lo <= int(off, 16) < hi:
                if op.startswith("IMAD"):
                    k = "IMAD.WIDE" if "WIDE" in op else ("IMAD.HI" if ".HI" in op else ("IMAD.MOV" if "MOV" in op else ("IMAD.X" if ".X" in op else "IMAD")))
                else:
                    k = op.split(".")[0]
                c[k] += 1
        return c
    lines = [f"# SASS opcode histogram of the P-256 field multiplication / squaring (out-of-line units) and of the fixed-base kernel body ({tag})",
             f"# cuobjdump -sass consensus_b200/build/inst_p256_kt5.o ; kernel {kname[:70]}", ""]
    body_end = min(v[0] for v in syms.values()) if syms else 1 << 30
    for label, (lo, hi) in [("p256_fmul_call", (syms["p256_fmul_call"][0], sum(syms["p256_fmul_call"]))), ("p256_fsqr_call", (syms["p256_fsqr_call"][0], sum(syms["p256_fsqr_call"]))),
                            ("k_verify_kt body (prologue + one addition site + final check)", (0, body_end))]:
        c = hist(lo, hi)
        lines.append(f"{label}: {sum(c.values())} instructions")
        lines.append("   " + ", ".join(f"{k} {n}" for k, n in c.most_common()))
    lines += ["", "round 1 (VERDICT.md): p256_fmul_call = 193 instructions (63 IMAD.WIDE + 90 IADD3 + 11 SEL + 9 IMAD.MOV + ...)",
              "round 2: the term-wise reduction (curve.cuh P256::redc) and the add-delta final subtraction removed 20 instructions; the wide MADs are the product's 64 (one is an IMAD.HI)"]
    open(os.path.join(out, f"{tag}_sass_fmul.txt"), "w").write("\n".join(lines) + "\n")
    # ptxas resource usage of every kernel
    res = []
    for f in sorted(glob.glob(os.path.join(root, "consensus_b200", "build", "inst_*.log"))):
        s = open(f).read()
        for m in re.finditer(r"Compiling entry function '(\S+)'.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s*: Used (\d+) registers", s):
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void sbv::", "").replace("sbv::", "")
            if name.startswith("k_"):
                res.append(f"{name[:58]:58s} regs {m.group(5):>3s}  stack {m.group(2):>4s}  spill st/ld {m.group(3):>4s}/{m.group(4):>4s}")
    open(os.path.join(out, f"{tag}_ptxas.txt"), "w").write("# ptxas -v per kernel (nvcc -gencode arch=compute_100a,code=sm_100a -O3)\n" + "\n".join(sorted(set(res))) + "\n")

# ---- A/B table ----
ab = []
for f, label in [(f"{visit}_bench.json", "default: multiplications out of line, 7 blocks/SM"), (f"{visit}_bench_kt1.json", "SBV_KT_VARIANT=1: multiplications inlined, 7 blocks/SM"),
                 (f"{visit}_bench_kt2.json", "SBV_KT_VARIANT=2: inlined, 6 blocks/SM (no spills)"), (f"{visit}_bench_basescall.json", "SBV_KT_BASES_CALL=1: doubling chain with out-of-line multiplications"),
                 ("v1_bench_generic.json", "SBV_GROUP_THRESHOLD=0: key grouping off (every item on k_verify_coz; visit 1, before the add-delta subtraction)")]:
    p = os.path.join(go, f)
    if os.path.exists(p) and os.path.getsize(p):
        j = last_json(p)
        ab.append(f"| {label} | {j['value']/1e6:.1f} | {j['ms_per_step']:.3f} | {j['step_latency_ms']:.3f} | {j['roofline']['kernel_ms']:.3f} | {j['e2e']['value']/1e6:.1f} |")
if len(ab) >= 4:   # only a visit that ran the variant A/Bs rewrites the table
    open(os.path.join(out, f"{tag}_variants.md"), "w").write(
        f"# A/B of kernel variants ({tag}; bench.py --steps 40..50, 1xB200, same visit)\n\n| variant | value M/s | ms/step (pipelined) | isolated step ms | dominant kernel ms | e2e M/s |\n|---|---|---|---|---|---|\n" + "\n".join(ab) + "\n")
p = os.path.join(go, f"{visit}_ubench.txt")
if os.path.exists(p):
    open(os.path.join(out, f"{tag}_ubench.txt"), "w").write("# tools/ubench.cu on the B200 (issue cycles per warp instruction per SM sub-partition at 1965 MHz)\n" + open(p).read())
print("wrote profiles/", tag)
