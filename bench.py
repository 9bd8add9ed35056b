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


def build_workload(name: str, seed: int, rank: int = 0, world: int = 1, tiles: int = 0, procs: int = 1):
    """→ (groups of THIS rank, description, scaling, sample names).  cfg2/micro/tiny: every rank its own
    copy of the shape (weak scaling, seed + rank).  cfg3/cfg4 (BASELINE configs[2]/[3]): ONE workload of
    1 Mb tiles, sharded over the ranks with dispatch.partition_units (strong scaling)."""
    if name in synth.TILED:
        from lancet2_b200.dispatch import partition_units
        spec = synth.TILED[name]
        n_tiles = tiles or spec["tiles"]
        costs = [synth.tile_cost(name, 42, t) for t in range(n_tiles)]
        mine = partition_units(costs, world)[rank]
        groups = synth.make_tiled_groups(name, 42, mine, procs=procs)
        return groups, tiled_desc(name, n_tiles, world), "strong", spec["sample_names"]
    groups, desc = _build_replica_workload(name, seed + rank)
    return groups, desc, "weak", ["normal", "tumor"]


def _cpulist(text: str):
    out = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        out.extend(range(int(lo), int(hi or lo) + 1))
    return out


def _numa_share(avail, local_rank: int, world: int):
    """Cores of the NUMA node each GPU hangs off (/sys/bus/pci/devices/<bdf>/numa_node), split between the
    ranks whose GPUs share that node.  None when the box does not say (single node, virtualised PCI)."""
    try:
        import torch
        nodes = []
        for r in range(world):
            p = torch.cuda.get_device_properties(r)
            bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
            nodes.append(int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read()))
        node = nodes[local_rank]
        if node < 0:
            return None
        ok = set(avail)
        cpus = [c for c in _cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read()) if c in ok]
        peers = [r for r in range(world) if nodes[r] == node]
        share = len(cpus) // len(peers)
        if share < 1:
            return None
        at = peers.index(local_rank)
        return cpus[at * share:(at + 1) * share]
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None


def tiled_desc(name: str, n_tiles: int, world: int) -> str:
    spec = synth.TILED[name]
    cov = "/".join(f"{int(c)}x" for _, c, _ in spec["samples"])
    return (f"{name}: synthetic {cov} ({len(spec['samples'])} samples) 2x150bp, {n_tiles} x 1 Mb reference tiles, spiked SNV/InDels, "
            f"1000bp windows step 800; tiles sharded over {world} rank(s) by dispatch.partition_units")


def _build_replica_workload(name: str, seed: int):
    if name == "cfg2":
        groups = synth.make_region_groups(seed, ref_len=1_000_000, cov_normal=30.0, cov_tumor=30.0)
        desc = "cfg2: synthetic 30x/30x tumor-normal 2x150bp, 1 Mb reference, spiked SNV/InDels, 1000bp windows step 800"
    elif name == "micro":
        groups = synth.make_groups(seed, 1024, read_len=150, hap_len=1000, n_haps=8, n_reads=256)
        desc = "cfg5 point: L=150, H=1000, P=8, R=256, 1024 groups"
    elif name == "l250":
        groups = synth.make_groups(1000 * 250 + 1000 + 8, 64, read_len=250, hap_len=1000, n_haps=8, n_reads=512)
        desc = "cfg5 point: L=250, H=1000, P=8, R=512, 64 groups"
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


_T0 = time.perf_counter()


def log(msg):
    """progress on stderr (stdout carries only the JSON line)"""
    sys.stderr.write(f"[bench {time.perf_counter() - _T0:7.1f}s] {msg}\n")
    sys.stderr.flush()


def emit_line(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def run_reference(args, rank, world):
    """CPU arm: the oracle on all host cores, bounded sample per step."""
    if rank != 0:
        return
    import oracle_lib as O
    olib, oflags = O.load_native_oracle()
    workload = args.workload if args.workload != "auto" else ("cfg2" if world == 1 else "cfg3")
    cores = os.cpu_count() or 1
    # the reference arm times a bounded sample: for a tiled workload the first tile is representative
    groups, desc, scaling, _ = build_workload(workload, 42, 0, 1, tiles=1 if workload in synth.TILED else 0)
    if workload in synth.TILED:  # same configuration string as the GPU arm of this launch
        desc = tiled_desc(workload, args.tiles or synth.TILED[workload]["tiles"], world)
    prm = O.default_params()
    probe = abi.Batch(groups[:8])
    t0 = time.perf_counter()
    O.oracle_genotype(probe, prm, n_threads=cores, lib=olib)
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
        O.oracle_genotype(abi.Batch(sel[:max(1, len(sel) // 8)]), prm, n_threads=cores, lib=olib)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.oracle_genotype(batch, prm, n_threads=cores, lib=olib)
    dt = time.perf_counter() - t0
    v = batch.n_pairs * args.steps / dt
    sample = f"first {len(sel)} of {len(groups)} groups ({batch.n_pairs} pairs) per step"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": desc, "note": "CPU restatement of the reference path (oracle port, not Lancet2/minimap2 binaries)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "flags": oflags, "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_line(line)


def adapter_meta(groups, batch, sample_names):
    """per-read metadata AddToTable needs (names, sample, start, insert size, flag, mapq), synthetic like the reads"""
    rng = np.random.default_rng(1)
    nr = batch.n_reads
    names = [nm for g in groups for nm in g.names]
    if all(g.sample is not None for g in groups):
        sample_id = np.asarray([s for g in groups for s in g.sample], dtype=np.int32)
    else:
        sample_id = np.asarray([0 if nm.startswith("n") else 1 for nm in names], dtype=np.int32)
    meta = {
        "blob": b"\0".join(x.encode() for x in names) + b"\0",
        "samples": b"\0".join(x.encode() for x in sample_names) + b"\0",
        "sample_id": sample_id,
        "start0": rng.integers(10_000, 20_000, nr).astype(np.int64),
        "isize": rng.integers(-500, 500, nr).astype(np.int64),
        "flag": (rng.integers(0, 2, nr) * 0x10 + 0x2).astype(np.uint16),
        "mapq": np.full(nr, 60, dtype=np.uint8),
        "softclip": np.zeros(nr, dtype=np.uint8),
    }
    return meta


def run_adapter(lib, device, batch, meta, threads, window, rounds):
    """The reference-shaped call: every group is one Genotype() payload handed to the C++
    lancet_gpu::GenotypeBatcher by `threads` worker threads (Enqueue/Collect with `window` payloads in
    flight per worker, the split ProcessWindow of SURVEY.md §8f #1); packing into the pinned slab, the
    one host->device copy per device batch, all kernels, the device->host copy of the assignments and
    AddToTable on the workers are all inside the timed region.  Returns (seconds, counters)."""
    fn = lib.lgr_adapter_batcher_bench
    fn.argtypes = [C.c_int, C.POINTER(abi.LgrBatchIn), C.c_char_p, C.c_char_p] + [C.c_void_p] * 6 + \
        [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_char_p, C.c_longlong]
    fn.restype = C.c_int
    bi = batch.c_struct()
    ctr = np.zeros(21, dtype=np.uint64)
    err = C.create_string_buffer(4096)
    rc = fn(device, C.byref(bi), meta["blob"], meta["samples"], meta["sample_id"].ctypes.data, meta["start0"].ctypes.data,
            meta["isize"].ctypes.data, meta["flag"].ctypes.data, meta["mapq"].ctypes.data, meta["softclip"].ctypes.data,
            threads, rounds, window, 1, ctr.ctypes.data, err, len(err))
    if rc != 0:
        raise SystemExit(f"adapter bench failed: {err.value.decode()}")
    return float(ctr[4]) * 1e-9, ctr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", help="auto = cfg2 on one GPU (BASELINE configs[1]), cfg3 sharded over the ranks otherwise (configs[2]); cfg2 | cfg3 | cfg4 | micro | tiny")
    ap.add_argument("--tiles", type=int, default=0, help="cfg3/cfg4: number of 1 Mb tiles (0 = the configuration's own: 50 / 10)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-inflight", type=int, default=3)
    ap.add_argument("--adapter-threads", type=int, default=0, help="worker threads of the adapter arm (0 = min(16, host cores / ranks))")
    ap.add_argument("--adapter-window", type=int, default=32, help="Genotype() payloads a worker keeps enqueued")
    ap.add_argument("--min-step-ms", type=float, default=50.0, help="repeat the batch inside a step until a step carries this much device time")
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
    cores = os.cpu_count() or 1
    # every rank keeps to its own share of the host cores (generation, packing workers, batcher thread)
    per_rank = max(1, cores // max(1, world))
    affinity = "inherited"
    try:
        avail = sorted(os.sched_getaffinity(0))
        if world > 1 and len(avail) >= world:
            mine = _numa_share(avail, local_rank, world)
            if mine:
                affinity = f"numa-local share of {len(mine)} cores"
            else:
                share = len(avail) // world
                mine = avail[local_rank * share:(local_rank + 1) * share]
                affinity = f"equal share of {len(mine)} cores (no NUMA information)"
            os.sched_setaffinity(0, set(mine))
            per_rank = len(mine)
    except (AttributeError, OSError):
        pass
    workload = args.workload if args.workload != "auto" else ("cfg2" if world == 1 else "cfg3")
    groups, desc, scaling, sample_names = build_workload(workload, 42, rank, world, tiles=args.tiles, procs=per_rank)
    if not groups:
        raise SystemExit(f"rank {rank}: no work (fewer tiles than ranks)")
    log(f"rank {rank}: workload {workload}: {len(groups)} groups generated")
    batch = abi.Batch(groups)
    gpu = GpuRealigner(local_rank)
    packed = abi.PackedBatch(groups, gpu.lib)  # the wire format the adapter ships: 2-bit planes, one slab
    packed.pin(torch)
    res = abi.Result(batch, 1 << 20)
    pin_result(res, torch)
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    warm = max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm: the slab is in HBM, a pass = unpack + every kernel of the path ----
    log(f"packed {packed.slab_bytes} bytes; resident arm")
    gpu.upload_packed(packed)
    probe = [gpu.run_resident().ms_kernels for _ in range(warm)]
    reps = max(1, int(np.ceil(args.min_step_ms / max(min(probe), 1e-3))))
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    ms_steps, ms_map, ms_ext, launches = [], [], [], 0
    last = None
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        acc = 0.0
        for _ in range(reps):  # one step = `reps` passes over the batch, L2 flushed (untimed) before each
            flush_buf.fill_(1)
            torch.cuda.synchronize()
            st = gpu.run_resident()
            acc += st.ms_kernels
            ms_map.append(st.ms_k_map)
            ms_ext.append(st.ms_k_ext)
            launches += st.kernel_launches
            last = st
        ms_steps.append(acc)
    barrier()
    wall_resident = time.perf_counter() - wall0
    dev_ms = sum(ms_steps)

    log(f"resident arm done ({reps} passes per step); C-ABI arms")
    # ---- C-ABI arms on the pre-packed slab (pinned): H2D + unpack + kernels + D2H per call ----
    for _ in range(2):
        gpu.genotype_packed(packed, batch, result=res, want_aln=False)
    barrier()
    t0 = time.perf_counter()
    e2e_st = None
    for _ in range(args.steps):
        _, e2e_st = gpu.genotype_packed(packed, batch, result=res, want_aln=False)
    barrier()
    e2e_single_s = time.perf_counter() - t0
    depth = max(1, min(args.e2e_inflight, abi.LGR_MAX_INFLIGHT))
    # every batch in flight owns a device arena the size of this context's: a whole cfg3 shard in one batch is tens of GB
    arena = int(gpu.lib.lgr_arena_bytes(gpu._ctx))
    free_b, _ = torch.cuda.mem_get_info()
    if arena > 0:
        depth = max(1, min(depth, int(free_b * 0.8) // int(arena * 1.6)))
    capi_s = e2e_single_s
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
                t, _ = gpu.submit_packed(packed, batch, result=ress[i % depth], want_aln=False)
                open_t.append(t)
            for t in open_t:
                stl = gpu.wait(t)
            return stl

        pipeline(2 * depth)
        barrier()
        t0 = time.perf_counter()
        e2e_st = pipeline(args.steps)
        barrier()
        capi_s = time.perf_counter() - t0

    # ---- adapter arm: the reference-shaped call (C++ GenotypeBatcher, worker threads, AddToTable) ----
    threads = args.adapter_threads or max(1, min(16, per_rank))
    log(f"C-ABI arms done; adapter arm with {threads} workers")
    meta = adapter_meta(groups, batch, sample_names)
    run_adapter(gpu.lib, local_rank, batch, meta, threads, args.adapter_window, max(2, min(warm, 4)))  # grows pinned staging
    barrier()
    adapter_s, actr = run_adapter(gpu.lib, local_rank, batch, meta, threads, args.adapter_window, args.steps)
    barrier()
    log("adapter arm done")
    clocks = sampler.stop()

    # max over ranks
    if world > 1:
        t = torch.tensor([dev_ms, capi_s, e2e_single_s, adapter_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, capi_s, e2e_single_s, adapter_s = (float(x) for x in t)
        cnt = torch.tensor([batch.n_pairs], dtype=torch.float64, device="cuda")
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        total_pairs = float(cnt[0])
    else:
        total_pairs = float(batch.n_pairs)

    if rank == 0:
        value = total_pairs * args.steps * reps / (dev_ms * 1e-3)
        capi_v = total_pairs * args.steps / capi_s
        adapter_v = total_pairs * args.steps / adapter_s
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        abytes = algorithmic_bytes(batch)
        map_ms = statistics.mean(ms_map)
        ext_ms = statistics.mean(ms_ext)
        achieved = abytes / (map_ms * 1e-3) / 1e9
        # per-launch DRAM traffic / warp-instruction count of the dominant kernel come from a committed ncu
        # capture of THIS build (profiles/r2_ncu_static.json, written by tools/ncu_summary.py); they are
        # constants of the build and workload, never measured under ncu inside a timed run, and are
        # reported as null when no capture of the current kernels has been committed
        static = {}
        try:
            with open(os.path.join(ROOT, "profiles", "r2_ncu_static.json")) as fh:
                static = json.load(fh)
        except OSError:
            pass
        wl = static.get(workload, {})
        traffic = wl.get("dram_bytes_per_launch")
        issue = None
        if wl.get("warp_inst_per_launch") and static.get("int_issue_peak_warp_inst_per_s"):
            ach = wl["warp_inst_per_launch"] / (map_ms * 1e-3)
            issue = {"kernel": wl.get("kernel", "k_chain_warp"), "achieved": ach, "peak": static["int_issue_peak_warp_inst_per_s"], "unit": "warp-instr/s",
                     "frac": ach / static["int_issue_peak_warp_inst_per_s"],
                     "source": "static: instruction count from " + wl.get("source", "ncu") + "; peak: " + static.get("int_issue_peak_source", "")}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": desc, "pairs_per_step_rank0": batch.n_pairs * reps, "pairs_per_step_all_ranks": total_pairs * reps, "passes_per_step": reps,
                       "pairs_per_pass": batch.n_pairs, "groups": batch.n_groups, "reads": batch.n_reads,
                       "haplotypes": batch.n_haps, "variants": batch.n_vars, "input": "packed wire format (2-bit base planes, quality dictionary planes), one slab",
                       "l2": "flushed before every timed pass (512 MiB write, untimed)", "timing": "CUDA events on the library stream, summed over the passes of a step"},
            "e2e": {"value": adapter_v, "unit": UNIT, "h2d_bytes_per_step": int(actr[17]) // max(1, args.steps),
                    "d2h_bytes_per_step": int(actr[18]) // max(1, args.steps),
                    "path": "lancet_gpu::GenotypeBatcher (C++ adapter with the reference's Genotype() call shape): one payload per group, worker threads Enqueue/Collect; packing from the caller's strings into the pinned slab, ONE H2D per device batch, all kernels, D2H of the assignments and AddToTable on the workers inside the timed region",
                    "worker_threads": threads, "payloads_in_flight_per_worker": args.adapter_window, "host_cores": cores, "cpu_affinity": affinity,
                    "genotype_calls_per_s": adapter_v * batch.n_groups / max(1, batch.n_pairs),
                    "device_batches_per_step": float(actr[0]) / max(1, args.steps), "max_payloads_in_one_batch": int(actr[3]),
                    "payloads_rerun_alone": int(actr[19]),
                    "thread_ms_per_step": {"pack_all_workers": float(actr[5]) * 1e-6 / args.steps, "submit_batcher": float(actr[6]) * 1e-6 / args.steps, "of_which_waiting_for_packers": float(actr[20]) * 1e-6 / args.steps,
                                           "wait_batcher": float(actr[7]) * 1e-6 / args.steps, "add_to_table_all_workers": float(actr[8]) * 1e-6 / args.steps,
                                           "wall": adapter_s * 1e3 / args.steps},
                    "l2_flushed": False,
                    "l2_note": "no flush kernel between steps: every device batch runs in one of four submission slots whose arenas (0.75 GiB of working set each on cfg2-sized batches) rotate, and its inputs arrive by a fresh H2D copy, so a batch finds none of its data in the 126 MB L2",
                    "capi": {"value": capi_v, "batches_in_flight": depth, "synchronous_value": total_pairs * args.steps / e2e_single_s,
                             "h2d_bytes_per_step": int(e2e_st.h2d_bytes), "d2h_bytes_per_step": int(e2e_st.d2h_bytes),
                             "ms_h2d_and_unpack": e2e_st.ms_h2d, "ms_kernels": e2e_st.ms_kernels, "ms_d2h": e2e_st.ms_d2h,
                             "note": "the same slab, already packed and pinned, through lgr_submit_packed/lgr_wait (one step = one call: H2D + unpack + kernels + D2H); with batches in flight neighbouring steps fill each other's low-parallelism phases and L2 is not flushed, so this figure can exceed `value`"}},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_chain_warp + k_chain_cold (phase A: seeds, anchors, chaining)", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": ("static: " + wl.get("source", "")) if traffic else None,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                         "algorithmic_bytes_per_launch": abytes, "kernel_ms": map_ms,
                         "note": "path is integer-issue / latency bound, not HBM bound (DESIGN.md §5); `issue` is the same kernel against the measured INT issue peak"},
            "issue": issue,
            "work": {"aligned_frac": last.n_aligned / max(1, last.n_pairs), "chain_evals_per_pair": last.chain_evals / max(1, last.n_pairs),
                     "anchors_per_pair": last.n_anchors / max(1, last.n_pairs), "dp_cells_per_pair": last.dp_cells / max(1, last.n_pairs),
                     "dp_cells_full_per_pair": last.dp_cells_full / max(1, last.n_pairs),
                     "gcups_computed": last.dp_cells / (last.ms_kernels * 1e-3) / 1e9,
                     "gcups_reference_rectangles": last.dp_cells_full / (last.ms_kernels * 1e-3) / 1e9,
                     "gcups_ext_stage": last.dp_cells / (ext_ms * 1e-3) / 1e9,
                     "chain_gevals_per_s": last.chain_evals / (last.ms_kernels * 1e-3) / 1e9,
                     "ms_index": last.ms_k_index, "ms_sketch": last.ms_k_sketch, "ms_map": last.ms_k_map, "ms_ext": last.ms_k_ext,
                     "ms_assign": last.ms_k_assign, "wall_resident_s": wall_resident},
        }
        if not args.no_cpu_baseline and world == 1:
            import oracle_lib as O
            olib, oflags = O.load_native_oracle()
            prm = O.default_params()
            probe_b = abi.Batch(groups[:8])
            t0 = time.perf_counter()
            O.oracle_genotype(probe_b, prm, n_threads=cores, lib=olib)
            rate = probe_b.n_pairs / max(time.perf_counter() - t0, 1e-6)
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
                O.oracle_genotype(sb, prm, n_threads=cores, lib=olib)
                passes += 1
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": sb.n_pairs * passes / dt, "unit": UNIT, "cores": cores, "kind": "port", "flags": oflags,
                                    "sample": f"first {len(sel)} of {len(groups)} groups ({sb.n_pairs} pairs) x {passes} passes, {dt:.1f} s"}
        emit_line(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
