#!/usr/bin/env python3
"""bench.py — ECDSA-P256 verifies/s at batch = 64K (BASELINE.json configs[1]) on N B200s.

A "step" is one pass of the hot path over one 65,536-signature batch PER GPU (batches shard
embarrassingly, so per-GPU work is fixed as N grows: weak scaling); with N > 1 every step ends with
the NCCL all-gather of the packed verdict bitmask (the only exchange the path has).

  value      device-timed, inputs already resident in HBM (16 rotating copies = 168 MB > L2)
  e2e        the same metric through the C ABI (sbv_verify_batch) with pinned HOST buffers:
             H2D of the 160 B/item batch and D2H of the verdicts inside the timed region
  roofline   dominant kernel k_verify: achieved wide-MAC/s (W = 272,256 MAC32 per verify)
             over CUDA-event kernel time vs the IMAD.WIDE peak probed in the same run; HBM fraction
             (161 B/verify vs MEASURED_PEAKS.json) reported beside it for completeness
  cpu_baseline  OpenSSL ECDSA_do_verify (oracle/, the stand-in for Go crypto/ecdsa — no Go toolchain
             exists here) on all host cores, same batch, rank 0 / N=1 only

`--impl reference` times that CPU implementation alone (the reference arm).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 65536
KEYS = 1024
MAC32_PER_VERIFY = 272_256       # SURVEY.md §8d canonical count (P-256)
BYTES_PER_VERIFY = 161           # 160 B in + 1 B out
N_COPIES = 16                    # rotating input copies: 16 x 10.5 MB > 126 MB L2
METRIC = "ECDSA-P256 verifies/sec at batch=64K"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for row in self.rows:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the newest
    committed `ncu --set full` summary under profiles/ (None if there is none)."""
    import glob
    import re
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_k_verify_ncu.txt"))):
        tot = 0.0
        found = 0
        for line in open(path):
            m = re.match(r"dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)", line)
            if m:
                scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(m.group(3), None)
                if scale:
                    tot += float(m.group(2)) * scale
                    found += 1
        if found == 2:
            best = (tot, os.path.basename(path))
    return best


def make_workload(rank: int):
    import oracle  # corpus generator + CPU baseline live in the oracle package (test/bench infrastructure)
    from oracle import corpus
    return corpus.make_batch(oracle.P256, n=BATCH, K=KEYS, seed=1 + 1000 * rank)


def run_reference(args, rank, world):
    """Reference arm: the CPU implementation of the path on the box's host cores."""
    if rank != 0:
        return
    import oracle
    b = make_workload(0)
    cores = oracle.ncores()
    keys = b["keys"]
    for _ in range(max(args.warmup, 1)):
        oracle.bench_verify(oracle.P256, b["r"][:8192], b["s"][:8192], keys, b["key_idx"][:8192], b["digest"][:8192], nthreads=cores)
    total = 0.0
    for _ in range(args.steps):
        t, ok = oracle.bench_verify(oracle.P256, b["r"], b["s"], keys, b["key_idx"], b["digest"], nthreads=cores)
        total += t
    value = BATCH * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "verifies/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 limbs (integer)", "data": "synthetic",
        "config": {"workload": "C2: ECDSA-P256 batch verify, 65,536 synthetic sigs, 1,024 keys, 1/16 corrupted", "batch": BATCH},
        "cpu_baseline": {"value": value, "unit": "verifies/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} x the full 65,536-signature batch, OpenSSL 3.0 ECDSA_do_verify (stand-in for Go crypto/ecdsa)"},
        "e2e": {"value": value, "unit": "verifies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="sbv", choices=["sbv", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import consensus_b200 as sbv

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    b = make_workload(rank)
    eng = sbv.Engine(devices=[local_rank])

    fields = ("r", "s", "qx", "qy", "digest")
    host = {k: torch.from_numpy(np.ascontiguousarray(b[k])).pin_memory() for k in fields}
    host_ok = torch.zeros(BATCH, dtype=torch.uint8).pin_memory()
    copies = [{k: host[k].to(dev, non_blocking=True) for k in fields} for _ in range(N_COPIES)]
    d_ok = torch.zeros(BATCH, dtype=torch.uint8, device=dev)
    n_words = BATCH // 32
    pow2 = (2 ** torch.arange(8, device=dev, dtype=torch.int32)).to(torch.uint8)
    gathered = torch.zeros(world * BATCH // 8, dtype=torch.uint8, device=dev) if world > 1 else None
    stream = torch.cuda.current_stream().cuda_stream
    # Consecutive steps are independent batches, so they are enqueued round-robin on three streams: the
    # scalar-preparation kernel of step i+1 (latency-bound: one inversion chain) and the head of its
    # verify kernel overlap the draining tail of step i.  Every step still does all of its work; the
    # timed region is bracketed by events on the main stream that wait for both.
    N_LANES = 3
    lanes = [torch.cuda.Stream(device=dev) for _ in range(N_LANES)]
    d_oks = [torch.zeros(BATCH, dtype=torch.uint8, device=dev) for _ in range(N_LANES)]

    # The verdict gather (pack to a bitmask + NCCL all_gather) of every step runs on ONE dedicated stream,
    # chained by events: collectives of one communicator execute in issue order, so issuing them on the
    # lane streams would re-serialise the lanes at every gather.
    gather_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    verify_done = [torch.cuda.Event() for _ in range(N_LANES)]
    gather_done = [None] * N_LANES

    def device_step(i, pipelined=True):
        c = copies[i % N_COPIES]
        if not pipelined:
            eng.verify_batch_device(sbv.P256, BATCH, c["r"].data_ptr(), c["s"].data_ptr(), c["qx"].data_ptr(), c["qy"].data_ptr(),
                                    c["digest"].data_ptr(), 32, d_ok.data_ptr(), stream=stream)
            return
        k = i % N_LANES
        lane, out = lanes[k], d_oks[k]
        with torch.cuda.stream(lane):
            if gather_done[k] is not None:
                lane.wait_event(gather_done[k])      # the previous user's verdicts have been packed
            eng.verify_batch_device(sbv.P256, BATCH, c["r"].data_ptr(), c["s"].data_ptr(), c["qx"].data_ptr(), c["qy"].data_ptr(),
                                    c["digest"].data_ptr(), 32, out.data_ptr(), stream=lane.cuda_stream)
            if world > 1:
                verify_done[k].record(lane)
        if world > 1:
            with torch.cuda.stream(gather_stream):
                gather_stream.wait_event(verify_done[k])
                packed = (out.view(-1, 8) * pow2).sum(dim=1, dtype=torch.uint8)  # 8 KiB verdict bitmask
                ev = torch.cuda.Event()
                ev.record(gather_stream)
                gather_done[k] = ev                  # `out` may be overwritten once it has been packed
                dist.all_gather_into_tensor(gathered, packed)

    def join_lanes():
        for lane in lanes:
            torch.cuda.current_stream().wait_stream(lane)
        if gather_stream is not None:
            torch.cuda.current_stream().wait_stream(gather_stream)

    def fork_lanes():
        for lane in lanes:
            lane.wait_stream(torch.cuda.current_stream())
        if gather_stream is not None:
            gather_stream.wait_stream(torch.cuda.current_stream())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- correctness gate: the verdicts of this run must equal the oracle's ----
    import oracle
    want = oracle.verify_batch(oracle.P256, b["r"], b["s"], b["qx"], b["qy"], b["digest"])
    device_step(0, pipelined=False)
    torch.cuda.synchronize()
    if not np.array_equal(d_ok.cpu().numpy(), want):
        raise SystemExit("bench: GPU verdicts differ from the oracle — refusing to report a number")

    # ---- device-timed value ----
    for i in range(args.warmup):
        device_step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fork_lanes()
    for i in range(args.steps):
        device_step(args.warmup + i)
    join_lanes()
    e1.record()
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    for k in range(N_LANES):
        if not np.array_equal(d_oks[k].cpu().numpy(), want):
            raise SystemExit("bench: pipelined verdicts differ from the oracle")
    # single-stream steps (no overlap): step latency, and the per-kernel CUDA-event durations the
    # roofline uses (kernel durations are only meaningful when kernels do not share the SMs)
    eng.profile_enable(True)
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record()
    for i in range(20):
        device_step(i, pipelined=False)
    l1.record()
    torch.cuda.synchronize()
    step_latency_ms = l0.elapsed_time(l1) / 20
    prep_ms, verify_ms, pairs = eng.profile_read()
    eng.profile_enable(False)
    launches = eng.kernel_launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    value = world * BATCH * args.steps / (dev_ms * 1e-3)

    # ---- end-to-end through the C ABI with pinned host buffers ----
    # Two host threads each keep one synchronous sbv_verify_batch call in flight (the reference calls
    # its Verifier from concurrent goroutines, view.go:537-541 / consensus.go:302-306); every call does
    # H2D of its 160 B/item batch, both kernels and the D2H of its verdicts.
    ptr = {k: host[k].data_ptr() for k in fields}
    host_oks = [torch.zeros(BATCH, dtype=torch.uint8).pin_memory() for _ in range(2)]
    def e2e_calls(tid, count):
        for _ in range(count):
            eng.verify_batch_ptr(sbv.P256, BATCH, ptr["r"], ptr["s"], ptr["qx"], ptr["qy"], ptr["digest"], 32, host_oks[tid].data_ptr())
    def e2e_run(total):
        ths = [threading.Thread(target=e2e_calls, args=(t, total // 2 + (t < total % 2))) for t in range(2)]
        for t in ths: t.start()
        for t in ths: t.join()
    e2e_run(2 * args.warmup)
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    for hk in host_oks:
        if not np.array_equal(hk.numpy(), want):
            raise SystemExit("bench: e2e verdicts differ from the oracle")
    e2e_value = world * BATCH * args.steps / e2e_s
    # one caller, one call at a time: the latency-bound form of the same number
    host_ok = host_oks[0]
    t0 = time.perf_counter()
    e2e_calls(0, 20)
    e2e_single = world * BATCH * 20 / max_over_ranks(time.perf_counter() - t0)

    # ---- registered-key path (extra, NOT the headline): keys registered once with sbv_set_keys, both
    # scalar multiplications fixed-base.  Same signatures; expected verdicts recomputed against the
    # registered key of each item (corruption classes that swap the key do not apply to this API).
    reg = None
    try:
        keys = b["keys"]
        t0 = time.perf_counter()
        eng.set_keys(np.zeros(KEYS, np.uint8), keys.reshape(KEYS, 2, 32))
        set_keys_s = time.perf_counter() - t0
        want_reg = oracle.verify_batch(oracle.P256, b["r"], b["s"], np.ascontiguousarray(keys[b["key_idx"], :32]),
                                       np.ascontiguousarray(keys[b["key_idx"], 32:]), b["digest"])
        d_slot = torch.from_numpy(b["key_idx"].astype(np.int32)).to(dev)
        def reg_step(i, pipelined=True):
            c = copies[i % N_COPIES]
            if not pipelined:
                eng.verify_registered_device(sbv.P256, BATCH, d_slot.data_ptr(), c["r"].data_ptr(), c["s"].data_ptr(), c["digest"].data_ptr(), 32,
                                             d_ok.data_ptr(), stream=stream)
                return
            lane, out = lanes[i % N_LANES], d_oks[i % N_LANES]
            with torch.cuda.stream(lane):
                eng.verify_registered_device(sbv.P256, BATCH, d_slot.data_ptr(), c["r"].data_ptr(), c["s"].data_ptr(), c["digest"].data_ptr(), 32,
                                             out.data_ptr(), stream=lane.cuda_stream)
        for i in range(args.warmup):
            reg_step(i, pipelined=False)
        torch.cuda.synchronize()
        if not np.array_equal(d_ok.cpu().numpy(), want_reg):
            raise RuntimeError("registered-key verdicts differ from the oracle")
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        fork_lanes()
        for i in range(args.steps):
            reg_step(i)
        join_lanes()
        r1.record()
        barrier()
        reg_ms = max_over_ranks(r0.elapsed_time(r1))
        for k in range(N_LANES):
            if not np.array_equal(d_oks[k].cpu().numpy(), want_reg):
                raise RuntimeError("pipelined registered-key verdicts differ from the oracle")
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for i in range(20):
            reg_step(i, pipelined=False)
        q1.record()
        torch.cuda.synchronize()
        reg_latency_ms = q0.elapsed_time(q1) / 20
        slot_host = torch.from_numpy(b["key_idx"].astype(np.int32)).pin_memory()
        def reg_e2e():
            vp = __import__("ctypes").c_void_p
            eng._check(eng._lib.sbv_verify_registered(eng._h, 0, BATCH, vp(slot_host.data_ptr()), vp(ptr["r"]), vp(ptr["s"]), vp(ptr["digest"]), 32,
                                                      vp(host_ok.data_ptr())), "sbv_verify_registered")
        for _ in range(args.warmup):
            reg_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            reg_e2e()
        reg_e2e_s = max_over_ranks(time.perf_counter() - t0)
        reg = {"value": world * BATCH * args.steps / (reg_ms * 1e-3), "unit": "verifies/s", "ms_per_step": reg_ms / args.steps,
               "step_latency_ms": reg_latency_ms, "e2e": world * BATCH * args.steps / reg_e2e_s, "keys": KEYS, "set_keys_seconds": set_keys_s,
               "note": "sbv_set_keys + sbv_verify_registered: per-key comb tables (512 KiB/key) built once per verification sequence; "
                       "not comparable to the keys-per-item headline"}
    except Exception as ex:  # the extra must never take the headline down
        reg = {"error": str(ex)}

    # ---- roofline of the dominant kernel (k_verify) ----
    mad_peak = eng.probe_mad_rate()                      # wide MAC32/s, measured in this run
    hbm_gbs, hbm_src = load_peaks()
    k_ms = verify_ms / max(pairs, 1)                     # average k_verify launch duration (CUDA events)
    mac_rate = BATCH * MAC32_PER_VERIFY / (k_ms * 1e-3)
    roofline = {
        "bound": "int32-mad (IMAD.WIDE issue rate; neither hbm nor tensor binds this path)",
        "kernel": "k_verify<P256>", "achieved": mac_rate / 1e12, "peak": mad_peak / 1e12, "unit": "TMAC32/s",
        "frac": mac_rate / mad_peak if mad_peak else None, "peak_source": "sbv_probe_mad_rate, same run",
        "kernel_ms": k_ms, "prep_kernel_ms": prep_ms / max(pairs, 1),
        "traffic": (ncu_traffic_bytes() or (None, None))[0], "traffic_source": (ncu_traffic_bytes() or (None, None))[1],
        "algorithmic_bytes_per_launch": BATCH * BYTES_PER_VERIFY,
        "hbm": {"achieved": BATCH * BYTES_PER_VERIFY / (k_ms * 1e-3) / 1e9, "peak": hbm_gbs, "unit": "GB/s",
                "frac": BATCH * BYTES_PER_VERIFY / (k_ms * 1e-3) / 1e9 / hbm_gbs, "peak_source": hbm_src},
    }

    line = {
        "metric": METRIC, "value": value, "unit": "verifies/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 limbs (integer)", "data": "synthetic",
        "config": {"workload": "C2: ECDSA-P256 batch verify, 65,536 synthetic sigs per GPU, 1,024 keys, 1/16 corrupted",
                   "batch_per_gpu": BATCH, "l2": f"{N_COPIES} rotating input copies (168 MB > 126 MB L2)",
                   "pipelining": "consecutive steps rotate over 3 CUDA streams; unpipelined step latency in step_latency_ms",
                   "exchange": "NCCL all_gather of the packed verdict bitmask per step, on a dedicated stream chained by events" if world > 1 else "none (1 GPU)",
                   "sharding": f"batch-parallel x{world}"},
        "e2e": {"value": e2e_value, "unit": "verifies/s", "h2d_bytes_per_step": 160 * BATCH * world, "d2h_bytes_per_step": BATCH * world,
                "callers": 2, "single_caller_value": e2e_single},
        "step_latency_ms": step_latency_ms, "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks, "registered_keys": reg,
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = oracle.ncores()
        reps = 4
        tot = 0.0
        for _ in range(reps):
            t, okc = oracle.bench_verify(oracle.P256, b["r"], b["s"], b["keys"], b["key_idx"], b["digest"], nthreads=cores)
            tot += t
        line["cpu_baseline"] = {"value": BATCH * reps / tot, "unit": "verifies/s", "cores": cores, "kind": "port",
                                "sample": f"{reps} x the full 65,536-signature batch; OpenSSL 3.0 ECDSA_do_verify on pre-built keys "
                                          "(stand-in for Go crypto/ecdsa: no Go toolchain)"}
    eng.close()
    if rank == 0 and world == 1:
        # consensus tx/s at n=4 (BASELINE configs[0]): in-process normal-path simulator, 1,000 txs,
        # RequestBatchMaxCount = 100; accept-all (= stock naive_chain) vs per-call CPU verifier vs GPU verifier
        sim = os.path.join(ROOT, "consensus_b200", "host", "sim")
        try:
            out = subprocess.run([sim, "1000", "100", "1"], capture_output=True, text=True, timeout=300)
            line["consensus_n4"] = json.loads(out.stdout.strip().splitlines()[-1])
        except Exception as ex:
            line["consensus_n4"] = {"error": str(ex)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
