#!/usr/bin/env python3
"""bench.py — read→haplotype realignment throughput (alignments/s) on B200.

One *step* = one pass of the hot path (index haplotypes, sketch reads, map every (read,
haplotype) pair, extend, score, assign) over one batch of synthetic Genotype() payloads.
Workload at N=1 = BASELINE.json configs[1] ("synthetic 30x/30x tumor-normal 2x150bp reads over
a 1 Mb synthetic reference with spiked SNVs/InDels"); every rank gets its own 1 Mb region
(weak scaling, no collective on the data path — windows are independent).

  value : pairs/s with the batch resident in HBM, device time from CUDA events on the
          launching stream, L2 flushed between steps (untimed), max over ranks
  e2e   : pairs/s through lgr_genotype_batch (the C-ABI call the Genotyper adapter makes) from
          pinned host buffers, H2D + kernels + D2H of the assignments inside the timed region
  --impl reference : the CPU oracle (restatement of the reference's minimap2 + Lancet2 scoring
          path — the reference itself cannot be built offline) on all host cores
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

from lancet2_b200 import abi, synth  # noqa: E402

METRIC = "read-haplotype realignments/s (mm_map-equivalent pairs incl. allele assignment)"
UNIT = "alignments/s"


def build_workload(name: str, seed: int):
    if name == "cfg2":
        groups = synth.make_region_groups(seed, ref_len=1_000_000, cov_normal=30.0, cov_tumor=30.0)
        desc = "cfg2: synthetic 30x/30x tumor-normal 2x150bp, 1 Mb reference, spiked SNV/InDels, 1000bp windows step 800"
    elif name == "micro":
        groups = synth.make_groups(seed, 1024, read_len=150, hap_len=1000, n_haps=8, n_reads=256)
        desc = "cfg5 point: L=150, H=1000, P=8, R=256, 1024 groups"
    elif name == "tiny":
        groups = synth.make_region_groups(seed, ref_len=60_000)
        desc = "tiny: 60 kb region (debug)"
    else:
        raise SystemExit(f"unknown workload {name}")
    return groups, desc


def algorithmic_bytes(batch: abi.Batch) -> int:
    """per pair 2-bit read + qualities + 96 B of results; per haplotype its
    2-bit bases once (amortised over its reads)."""
    rl = np.diff(batch.read_off).astype(np.int64)
    per_read_pairs = np.diff(batch.pair_off).astype(np.int64)
    b = int(((rl // 4 + rl + 96) * per_read_pairs).sum())
    b += int((np.diff(batch.hap_off) // 4).sum())
    return b


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                       "100", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f:
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pin_batch(batch: abi.Batch, torch):
    """move every host array of the batch into pinned memory (torch is plumbing only)"""
    keep = []
    for name in ("grp_hap_begin", "grp_read_begin", "grp_var_begin", "hap_off", "hap_bases", "read_off", "read_bases",
                 "read_quals", "read_name_hash", "var_hap_off", "var_start", "var_len", "var_allele", "grp_mid_occ"):
        t = torch.from_numpy(getattr(batch, name)).pin_memory()
        keep.append(t)
        setattr(batch, name, t.numpy())
    batch._pinned = keep


def pin_result(res: abi.Result, torch):
    keep = []
    for name in ("assign",):
        arr = getattr(res, name)
        t = torch.from_numpy(arr.view(np.uint8).reshape(-1)).pin_memory()
        keep.append(t)
        setattr(res, name, t.numpy().view(arr.dtype))
    res._pinned = keep


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, torchrun notices)
    also write to fd 1, so keep a private handle on the real stdout and point fd 1 at stderr for
    everything else."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit_line(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def run_reference(args, rank, world):
    """CPU arm: the oracle on all host cores, bounded sample per step."""
    if rank != 0:
        return
    import oracle_lib as O
    groups, desc = build_workload(args.workload, 42)
    cores = os.cpu_count() or 1
    prm = O.default_params()
    probe = abi.Batch(groups[:8])
    t0 = time.perf_counter()
    O.oracle_genotype(probe, prm, n_threads=cores)
    rate = probe.n_pairs / max(time.perf_counter() - t0, 1e-6)
    want_pairs = rate * 8.0  # ~8 s per step
    sel, acc = [], 0
    for g in groups:
        sel.append(g)
        acc += len(g.reads) * len(g.haps)
        if acc >= want_pairs:
            break
    batch = abi.Batch(sel)
    for _ in range(max(args.warmup, 1) if args.warmup > 0 else 0):
        O.oracle_genotype(abi.Batch(sel[:max(1, len(sel) // 8)]), prm, n_threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.oracle_genotype(batch, prm, n_threads=cores)
    dt = time.perf_counter() - t0
    v = batch.n_pairs * args.steps / dt
    sample = f"first {len(sel)} of {len(groups)} groups ({batch.n_pairs} pairs) per step"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": desc, "note": "CPU restatement of the reference path (oracle port, not Lancet2/minimap2 binaries)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_line(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-inflight", type=int, default=3)
    args = ap.parse_args()
    claim_stdout()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the realignment path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from lancet2_b200.realign import GpuRealigner
    groups, desc = build_workload(args.workload, 42 + rank)
    batch = abi.Batch(groups)
    pin_batch(batch, torch)
    gpu = GpuRealigner(local_rank)
    res = abi.Result(batch, 1 << 20)
    pin_result(res, torch)
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm ----
    gpu.upload(batch)
    for _ in range(max(args.warmup, 3)):
        gpu.run_resident()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    ms_steps, ms_map, launches = [], [], 0
    last = None
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush_buf.fill_(1)
        torch.cuda.synchronize()
        st = gpu.run_resident()
        ms_steps.append(st.ms_kernels)
        ms_map.append(st.ms_k_map)
        launches += st.kernel_launches
        last = st
    barrier()
    wall_resident = time.perf_counter() - wall0
    dev_ms = sum(ms_steps)

    # ---- end-to-end arm (C-ABI call, pinned host buffers, H2D + D2H inside every step) ----
    # single context: one synchronous lgr_genotype_batch per step
    for _ in range(2):
        gpu.genotype_batch(batch, result=res, want_aln=False)
    barrier()
    t0 = time.perf_counter()
    e2e_st = None
    for _ in range(args.steps):
        _, e2e_st = gpu.genotype_batch(batch, result=res, want_aln=False)
    barrier()
    e2e_single_s = time.perf_counter() - t0
    # pipelined: the same per-step call through lgr_submit/lgr_wait with a few batches in flight on
    # ONE context and ONE host thread: every step still carries its own H2D and D2H, but the
    # copies of one step overlap the kernels of the previous one
    depth = max(1, min(args.e2e_inflight, abi.LGR_MAX_INFLIGHT))
    e2e_s = e2e_single_s
    if depth > 1:
        ress = [res]
        for _ in range(depth - 1):
            r2 = abi.Result(batch, 1 << 20)
            pin_result(r2, torch)
            ress.append(r2)

        def pipeline(n_steps):
            open_t = []
            stl = None
            for i in range(n_steps):
                if len(open_t) == depth:
                    stl = gpu.wait(open_t.pop(0))
                t, _ = gpu.submit(batch, result=ress[i % depth], want_aln=False)
                open_t.append(t)
            for t in open_t:
                stl = gpu.wait(t)
            return stl

        pipeline(2 * depth)
        barrier()
        t0 = time.perf_counter()
        e2e_st = pipeline(args.steps)
        barrier()
        e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()

    # max over ranks
    if world > 1:
        t = torch.tensor([dev_ms, e2e_s, e2e_single_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, e2e_single_s = float(t[0]), float(t[1]), float(t[2])
        cnt = torch.tensor([batch.n_pairs], dtype=torch.float64, device="cuda")
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        total_pairs = float(cnt[0])
    else:
        total_pairs = float(batch.n_pairs)

    if rank == 0:
        value = total_pairs * args.steps / (dev_ms * 1e-3)
        e2e_v = total_pairs * args.steps / e2e_s
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        abytes = algorithmic_bytes(batch)
        map_ms = statistics.mean(ms_map)
        achieved = abytes / (map_ms * 1e-3) / 1e9
        # per-launch DRAM traffic and warp-instruction count of the dominant kernel: static for a
        # given build and workload, taken from the committed ncu capture (never measured under ncu here)
        static = {}
        try:
            with open(os.path.join(ROOT, "profiles", "r1_ncu_static.json")) as fh:
                static = json.load(fh)
        except OSError:
            pass
        wl = static.get(args.workload, {}) if world == 1 or args.workload in static else {}
        traffic = wl.get("dram_bytes_per_launch")
        issue = None
        if wl.get("warp_inst_per_launch") and static.get("int_issue_peak_warp_inst_per_s"):
            ach = wl["warp_inst_per_launch"] / (map_ms * 1e-3)
            issue = {"kernel": "k_chain_warp", "achieved": ach, "peak": static["int_issue_peak_warp_inst_per_s"], "unit": "warp-instr/s",
                     "frac": ach / static["int_issue_peak_warp_inst_per_s"],
                     "source": "instruction count: " + wl.get("source", "ncu") + "; peak: " + static.get("int_issue_peak_source", "")}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": desc, "pairs_per_step_per_gpu": batch.n_pairs, "groups": batch.n_groups,
                       "reads": batch.n_reads, "haplotypes": batch.n_haps, "variants": batch.n_vars,
                       "l2": "flushed between timed steps (512 MiB write, untimed)", "timing": "CUDA events on the library stream"},
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": int(e2e_st.h2d_bytes), "d2h_bytes_per_step": int(e2e_st.d2h_bytes),
                    "batches_in_flight": depth,
                    "genotype_calls_per_s": e2e_v * batch.n_groups / max(1, batch.n_pairs),  # groups (= Genotype() payloads ~ windows) per second
                    "synchronous_value": total_pairs * args.steps / e2e_single_s,
                    "ms_h2d": e2e_st.ms_h2d, "ms_kernels": e2e_st.ms_kernels, "ms_d2h": e2e_st.ms_d2h,
                    "note": "every step is one lgr_submit+lgr_wait of the whole batch from pinned host buffers (H2D + kernels + D2H per step, one host thread); batches_in_flight steps are outstanding so copies overlap kernels; synchronous_value is the same through lgr_genotype_batch, one call at a time"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_chain_warp", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                         "algorithmic_bytes_per_launch": abytes, "kernel_ms": map_ms,
                         "note": "path is integer-issue / latency bound, not HBM bound (DESIGN.md §roofline); `issue` is the same kernel against the measured INT issue peak"},
            "issue": issue,
            "work": {"aligned_frac": last.n_aligned / max(1, last.n_pairs), "chain_evals_per_pair": last.chain_evals / max(1, last.n_pairs),
                     "anchors_per_pair": last.n_anchors / max(1, last.n_pairs), "dp_cells_per_pair": last.dp_cells / max(1, last.n_pairs),
                     "dp_cells_full_per_pair": last.dp_cells_full / max(1, last.n_pairs),
                     "gcups_computed": last.dp_cells / (last.ms_kernels * 1e-3) / 1e9,
                     "gcups_reference_rectangles": last.dp_cells_full / (last.ms_kernels * 1e-3) / 1e9,
                     "chain_gevals_per_s": last.chain_evals / (last.ms_kernels * 1e-3) / 1e9,
                     "ms_index": last.ms_k_index, "ms_sketch": last.ms_k_sketch, "ms_map": last.ms_k_map, "ms_ext": last.ms_k_ext,
                     "ms_assign": last.ms_k_assign, "wall_resident_s": wall_resident},
        }
        if not args.no_cpu_baseline and world == 1:
            import oracle_lib as O
            cores = os.cpu_count() or 1
            prm = O.default_params()
            probe = abi.Batch(groups[:8])
            t0 = time.perf_counter()
            O.oracle_genotype(probe, prm, n_threads=cores)
            rate = probe.n_pairs / max(time.perf_counter() - t0, 1e-6)
            sel, acc = [], 0
            for g in groups:
                sel.append(g)
                acc += len(g.reads) * len(g.haps)
                if acc >= rate * 12.0:
                    break
            sb = abi.Batch(sel)
            t0 = time.perf_counter()
            passes = 0
            while passes < 1 or time.perf_counter() - t0 < 10.0:
                O.oracle_genotype(sb, prm, n_threads=cores)
                passes += 1
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": sb.n_pairs * passes / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"first {len(sel)} of {len(groups)} groups ({sb.n_pairs} pairs) x {passes} passes, {dt:.1f} s"}
        emit_line(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
