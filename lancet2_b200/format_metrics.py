"""Host-side handle of the device FORMAT path: the evidence streams of many
`caller::VariantSupport` objects in, every FORMAT accessor's value out
(reference: src/lancet/caller/variant_support.h:60-330).  Plumbing only — the arithmetic is
k_fmt_dedup / k_fmt_metrics in csrc/lgr_format.cu; there is no CPU implementation here and the
constructor raises when the CUDA library or a device is missing."""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence, Tuple

import numpy as np

from . import abi


class GpuFormatMetrics:
    def __init__(self, device: int = 0):
        self._main = abi.load_library()
        self._lib = self._main
        # A/B builds of the same four entry points (e.g. csrc/liblgr_format_scan.so, the O(n^2)-scan kernels; the shipped build is -DLGR_FMT_SORT)
        variant = os.environ.get("LGR_FORMAT_LIBRARY")
        if variant:
            self._lib = C.CDLL(variant)
            self._lib.lgr_format_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
            self._lib.lgr_format_create.restype = C.c_int
            self._lib.lgr_format_destroy.argtypes = [C.c_void_p]
            self._lib.lgr_format_destroy.restype = None
            self._lib.lgr_format_last_error.argtypes = [C.c_void_p]
            self._lib.lgr_format_last_error.restype = C.c_char_p
            self._lib.lgr_format_metrics.argtypes = [C.c_void_p, C.POINTER(abi.LgrEvidenceIn), C.c_void_p, C.POINTER(C.c_float)]
            self._lib.lgr_format_metrics.restype = C.c_int
        self._ctx = C.c_void_p()
        rc = self._lib.lgr_format_create(device, C.byref(self._ctx))
        if rc != 0:
            msg = self._lib.lgr_format_last_error(None).decode()
            self._ctx = C.c_void_p()
            raise RuntimeError(f"lgr_format_create failed ({self._main.lgr_strerror(rc).decode()}): {msg}")

    def close(self) -> None:
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.lgr_format_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def compute(self, batch: "abi.EvidenceBatch | Sequence[dict]") -> Tuple[np.ndarray, float]:
        """(records of abi.FORMAT_DTYPE, one per support; CUDA-event ms of the two kernels)."""
        if not isinstance(batch, abi.EvidenceBatch):
            batch = abi.EvidenceBatch(batch)
        out = np.zeros(batch.n_supports, dtype=abi.FORMAT_DTYPE)
        st = batch.c_struct()
        ms = C.c_float(0.0)
        rc = self._lib.lgr_format_metrics(self._ctx, C.byref(st), out.ctypes.data, C.byref(ms))
        self.last_rc = rc
        if rc == abi.LGR_E_PARTIAL:  # supports with more than LGR_FMT_MAX_ALLELES alleles are flagged, not fatal
            return out, float(ms.value)
        if rc != 0:
            raise RuntimeError(f"lgr_format_metrics failed ({self._main.lgr_strerror(rc).decode()}): "
                               f"{self._lib.lgr_format_last_error(self._ctx).decode()}")
        return out, float(ms.value)

    def from_assign(self, batch: "abi.Batch", *, n_samples: int, sample_id, start0, isize, sam_flag, mapq, softclip,
                    var_n_alleles, var_len, host_assign=None, dev_assign: int = 0):
        """AddToTable + FORMAT math on the device (lgr_format_from_assign): evidence columns are built by
        k_evidence_* from the realignment's lgr_assign records — `dev_assign` (a device address from
        GpuRealigner.resident_assign()) or `host_assign` (numpy records, uploaded) — and the per-read fields.
        → (records, keys[S,3] = group / variant / sample id, CUDA-event ms of all kernels)."""
        if self._lib is not self._main:
            raise RuntimeError("from_assign is not part of the A/B FORMAT libraries")
        keep = [np.ascontiguousarray(a, dtype=dt) for a, dt in (
            (batch.grp_read_begin, np.int32), (batch.grp_var_begin, np.int32),
            (np.diff(batch.grp_hap_begin), np.int32), (var_n_alleles, np.int32), (var_len, np.int32), (isize, np.int64),
            (start0, np.int64), (batch.read_name_hash, np.uint32), (sample_id, np.int32), (sam_flag, np.uint16),
            (mapq, np.uint8), (softclip, np.uint8))]
        st = abi.LgrAssignBatch()
        st.dev_assign = dev_assign or None
        if host_assign is not None:
            host_assign = np.ascontiguousarray(host_assign)
            st.host_assign = host_assign.ctypes.data
        st.n_assign = int(batch.n_assign)
        st.n_groups, st.n_reads, st.n_vars, st.n_samples = batch.n_groups, batch.n_reads, batch.n_vars, n_samples
        for name, arr in zip(("grp_read_begin", "grp_var_begin", "grp_n_haps", "var_n_alleles", "var_len", "read_insert_size",
                              "read_aln_start", "read_name_hash", "read_sample", "read_sam_flag", "read_map_qual",
                              "read_soft_clipped"), keep):
            setattr(st, name, arr.ctypes.data)
        cap = max(1, batch.n_vars * n_samples)
        out = np.zeros(cap, dtype=abi.FORMAT_DTYPE)
        keys = np.zeros((cap, 3), dtype=np.int32)
        n_sup, ms = C.c_int32(0), C.c_float(0.0)
        rc = self._lib.lgr_format_from_assign(self._ctx, C.byref(st), out.ctypes.data, cap, keys.ctypes.data, C.byref(n_sup), C.byref(ms))
        self.last_rc = rc
        if rc not in (0, abi.LGR_E_PARTIAL):
            raise RuntimeError(f"lgr_format_from_assign failed ({self._main.lgr_strerror(rc).decode()}): "
                               f"{self._lib.lgr_format_last_error(self._ctx).decode()}")
        return out[:n_sup.value], keys[:n_sup.value], float(ms.value)

    def debug_evidence(self) -> dict:
        """The evidence columns the last from_assign built on the device (test hook)."""
        probe = abi.LgrEvidenceIn()
        if self._lib.lgr_format_debug_evidence(self._ctx, C.byref(probe)) != 0:
            raise RuntimeError("lgr_format_debug_evidence failed")
        n, s = int(probe.n_evidence), int(probe.n_supports)
        cols = {"sup_begin": np.zeros(s + 1, np.int64), "sup_n_alleles": np.zeros(s, np.int32),
                "sup_variant_len": np.zeros(s, np.int32), "sup_total_haps": np.zeros(s, np.int32)}
        for name, dt in abi.EVIDENCE_FIELDS:
            cols[name] = np.zeros(n, dtype=dt)
        st = abi.LgrEvidenceIn()
        for name, arr in cols.items():
            setattr(st, name, arr.ctypes.data if arr.size else None)
        if self._lib.lgr_format_debug_evidence(self._ctx, C.byref(st)) != 0:
            raise RuntimeError("lgr_format_debug_evidence failed")
        return cols
