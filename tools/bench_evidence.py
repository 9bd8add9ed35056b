#!/usr/bin/env python3
"""AddToTable + FORMAT math for the supports of one cfg2 step, two ways:
  host:   lgr_assign records copied to the host, lancet_gpu::EvidenceColumns::AppendJob (C++, one thread),
          columns copied back, k_fmt_dedup / k_fmt_metrics      (lgr_download + lgr_adapter_format_metrics)
  device: k_evidence_count / _scan / _scatter on the resident records, then the same two kernels
          (lgr_format_from_assign with dev_assign; 28 B of per-read fields go up, the records come back)
One JSON line; the two record sets must be byte-equal."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import bench  # noqa: E402
from lancet2_b200 import abi  # noqa: E402
from lancet2_b200.format_metrics import GpuFormatMetrics  # noqa: E402
from lancet2_b200.realign import GpuRealigner  # noqa: E402
from test_adapter import _meta  # noqa: E402
from test_gpu_evidence_build import variant_tables  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    groups, desc = bench._build_replica_workload(name, 42)
    batch = abi.Batch(groups)
    names, blob, (sample_id, start0, isize, flag, mapq, softclip) = _meta(None, groups, batch)
    batch.read_name_hash[:batch.n_reads] = [abi.x31_hash(n) for n in names]
    k, vlen = variant_tables(groups, batch)
    gpu, fmt = GpuRealigner(0), GpuFormatMetrics(0)
    lib = gpu.lib
    gpu.upload_packed(abi.PackedBatch(groups, lib))
    gpu.run_resident()
    dev, n = gpu.resident_assign()
    res = abi.Result(batch)
    lib.lgr_adapter_format_metrics.argtypes = [C.c_int, C.POINTER(abi.LgrBatchIn), C.c_char_p, C.c_char_p] + [C.c_void_p] * 7 + \
        [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
    lib.lgr_adapter_format_metrics.restype = C.c_int
    bi = batch.c_struct()
    cap = batch.n_vars * 2
    out_h = np.zeros(cap, dtype=abi.FORMAT_DTYPE)
    err = C.create_string_buffer(512)
    kw = dict(n_samples=2, sample_id=sample_id, start0=start0, isize=isize, sam_flag=flag, mapq=mapq, softclip=softclip,
              var_n_alleles=k, var_len=vlen)
    t_host, t_dev, k_dev = [], [], []
    for it in range(reps + 2):
        t0 = time.perf_counter()
        gpu.download(batch, res)                      # D2H of lgr_aln + lgr_assign (the host path needs the records)
        t1 = time.perf_counter()
        ns = lib.lgr_adapter_format_metrics(0, C.byref(bi), blob, b"normal\0tumor\0", sample_id.ctypes.data, start0.ctypes.data,
                                            isize.ctypes.data, flag.ctypes.data, mapq.ctypes.data, softclip.ctypes.data,
                                            res.assign.ctypes.data, out_h.ctypes.data, cap, err, len(err))
        t2 = time.perf_counter()
        assert ns > 0, err.value.decode()
        got, keys, ms = fmt.from_assign(batch, dev_assign=dev, **kw)
        t3 = time.perf_counter()
        if it >= 2:
            t_host.append(((t1 - t0) + (t2 - t1)) * 1e3), t_dev.append((t3 - t2) * 1e3), k_dev.append(ms)
    same = len(got) == ns and got.tobytes() == out_h[:ns].tobytes()
    print(json.dumps({"workload": desc, "supports": int(ns), "evidence_records": int(fmt.debug_evidence()["sup_begin"][-1]),
                      "assign_records": int(n), "records_byte_equal": bool(same),
                      "host_path_ms": {"median": float(np.median(t_host)), "what": "lgr_download (aln + assign) + BuildJobs/AppendJob on one host thread + lgr_format_metrics (new GpuFormatMetrics per call)"},
                      "device_path_ms": {"median": float(np.median(t_dev)), "kernels_ms": float(np.median(k_dev)),
                                         "what": "lgr_format_from_assign(dev_assign): per-read fields H2D, k_evidence_count/scan/scatter, k_fmt_dedup, k_fmt_metrics, records D2H"},
                      "h2d_bytes_device_path": int(batch.n_reads * 28 + batch.n_vars * 12), "d2h_bytes_host_path": int(n * 48)}))
    fmt.close()
    gpu.close()


if __name__ == "__main__":
    main()
