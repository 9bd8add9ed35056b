"""Host-side handle on the repeat-detection kernel (lgr_repeat_*, SURVEY.md §8f #3 first half):
`HasRepeat(SlidingView(seq, k), max_mismatches)` of the reference (src/lancet/base/repeat.cpp:348-371,
called from cbdg/graph.h:127-131 and core/variant_builder.cpp:116-117) for many (window, k) jobs in
one device call.  No CPU path exists: construction raises when the library or a GPU is missing."""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

import numpy as np

from . import abi


_JOB_DTYPE = np.dtype([("seq_off", "<i8"), ("seq_len", "<i4"), ("k", "<i4"), ("max_mismatches", "<i4"), ("reserved", "<i4")])
assert _JOB_DTYPE.itemsize == C.sizeof(abi.LgrRepeatJob)


class GpuRepeatScan:
    def __init__(self, device: int = 0):
        self._lib = abi.load_library()
        self._ctx = C.c_void_p()
        rc = self._lib.lgr_repeat_create(device, C.byref(self._ctx))
        if rc != 0:
            msg = self._lib.lgr_repeat_last_error(None).decode()
            self._ctx = C.c_void_p()
            raise RuntimeError(f"lgr_repeat_create failed ({self._lib.lgr_strerror(rc).decode()}): {msg}")
        self.last_rc = 0

    def close(self) -> None:
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.lgr_repeat_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def scan(self, jobs: Sequence[Tuple[bytes, int, int]]) -> Tuple[np.ndarray, float]:
        """jobs: (sequence, k, max_mismatches).  → (uint8 answers: 0 / 1 / LGR_REPEAT_TOO_LONG, kernel ms).
        Equal sequences are uploaded once (a window is asked once per k of the graph's k-loop)."""
        offs, blobs, pos = {}, [], 0
        arr = np.zeros(max(len(jobs), 1), dtype=_JOB_DTYPE)
        seq_off = np.empty(len(jobs), dtype=np.int64)
        for i, (seq, _, _) in enumerate(jobs):
            key = id(seq) if len(seq) > 64 else seq
            at = offs.get(key)
            if at is None:
                at = offs[key] = pos
                blobs.append(seq)
                pos += len(seq)
            seq_off[i] = at
        if jobs:
            arr["seq_off"][:len(jobs)] = seq_off
            arr["seq_len"][:len(jobs)] = [len(j[0]) for j in jobs]
            arr["k"][:len(jobs)] = [j[1] for j in jobs]
            arr["max_mismatches"][:len(jobs)] = [j[2] for j in jobs]
        buf = np.frombuffer(b"".join(blobs) + b"\0", dtype=np.uint8)
        out = np.zeros(max(len(jobs), 1), dtype=np.uint8)
        ms = C.c_float(0.0)
        rc = self._lib.lgr_repeat_scan(self._ctx, buf.ctypes.data, pos, arr.ctypes.data, len(jobs), out.ctypes.data, C.byref(ms))
        self.last_rc = rc
        if rc not in (0, abi.LGR_E_PARTIAL):
            raise RuntimeError(f"lgr_repeat_scan failed ({self._lib.lgr_strerror(rc).decode()}): "
                               f"{self._lib.lgr_repeat_last_error(self._ctx).decode()}")
        return out[:len(jobs)], float(ms.value)
