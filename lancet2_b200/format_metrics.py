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
