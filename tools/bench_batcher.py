#!/usr/bin/env python3
"""Throughput of the cross-thread batcher (lancet_gpu::GenotypeBatcher): every cfg2 group is one
blocking Genotype() call, issued from T worker threads the way Lancet2's workers call
Genotyper::Genotype (core/variant_builder.cpp:258-259).  Prints one JSON line per thread count:
Genotype() calls/s (= graph components/s ~ windows/s for the hot path alone) and pairs/s, with
AddToTable on the workers inside the timed region.
usage: python tools/bench_batcher.py [threads ...]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from lancet2_b200 import abi, synth  # noqa: E402


def main():
    # each argument is threads or threads:window (window = groups a worker keeps enqueued; 1 = blocking calls)
    cfgs = [tuple(int(v) for v in (x.split(":") + ["1"])[:2]) for x in sys.argv[1:]] or [(1, 1), (16, 1), (4, 16), (16, 16)]
    lib = abi.load_library()
    n_dev = int(os.environ.get("LGR_BENCH_DEVICES", "1"))
    lib.lgr_adapter_batcher_bench.argtypes = [C.c_int, C.POINTER(abi.LgrBatchIn), C.c_char_p, C.c_char_p] + \
        [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_char_p, C.c_longlong]
    lib.lgr_adapter_batcher_bench.restype = C.c_int
    groups = synth.make_region_groups(42, ref_len=1_000_000)
    batch = abi.Batch(groups)
    nr = batch.n_reads
    rng = np.random.default_rng(1)
    names = [nm for g in groups for nm in g.names]
    sample_id = np.asarray([0 if nm.startswith("n") else 1 for nm in names], dtype=np.int32)
    start0 = rng.integers(10_000, 20_000, nr).astype(np.int64)
    isize = rng.integers(-500, 500, nr).astype(np.int64)
    flag = (rng.integers(0, 2, nr) * 0x10 + 0x2).astype(np.uint16)
    mapq = np.full(nr, 60, dtype=np.uint8)
    softclip = np.zeros(nr, dtype=np.uint8)
    bi = batch.c_struct()
    blob = b"\0".join(x.encode() for x in names) + b"\0"
    args = (sample_id.ctypes.data, start0.ctypes.data, isize.ctypes.data, flag.ctypes.data, mapq.ctypes.data, softclip.ctypes.data)
    err = C.create_string_buffer(4096)
    for t, window in cfgs:
        rounds = 32
        ctr = np.zeros(21, dtype=np.uint64)
        for _ in range(2):  # first pass warms the allocators of the process
            rc = lib.lgr_adapter_batcher_bench(0, C.byref(bi), blob, b"normal\0tumor\0", *args, t, rounds, window, n_dev, ctr.ctypes.data, err, len(err))
            assert rc == 0, err.value.decode()
        sec = float(ctr[4]) * 1e-9
        print(json.dumps({"devices": n_dev, "calls_per_device": [int(x) for x in ctr[9:9 + n_dev]], "worker_threads": t, "groups_in_flight_per_worker": window, "genotype_calls_per_s": float(ctr[1]) / sec, "pairs_per_s": float(ctr[2]) / sec,
                          "device_batches": int(ctr[0]), "thread_ms": {"pack_all_workers": float(ctr[5]) * 1e-6, "submit_batcher": float(ctr[6]) * 1e-6, "of_which_waiting_for_packers": float(ctr[20]) * 1e-6, "wait_batcher": float(ctr[7]) * 1e-6, "add_to_table_all_workers": float(ctr[8]) * 1e-6, "wall": sec * 1e3}, "h2d_bytes": int(ctr[17]), "d2h_bytes": int(ctr[18]), "calls": int(ctr[1]), "max_calls_in_one_batch": int(ctr[3]),
                          "workload": "cfg2 groups, one Genotype() payload per group (blocking call when 1 in flight, Enqueue/Collect otherwise), AddToTable on the workers"}), flush=True)


if __name__ == "__main__":
    main()
