"""TEST INFRASTRUCTURE: builds/loads tests/hostemu/libhostemu.so — the kernels' per-lane
scalar core (lgr_core.cuh) compiled by g++ — so CPU tests can diff it against the oracle."""
import ctypes as C
import os
import subprocess

from lancet2_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostemu", "hostemu.cpp")
SO = os.path.join(HERE, "hostemu", "libhostemu.so")
CORE = os.path.join(os.path.dirname(HERE), "lancet2_b200", "csrc", "lgr_core.cuh")
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    stale = (not os.path.exists(SO)) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(CORE))
    if stale:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", "-o",
                               SO, SRC])
    lib = C.CDLL(SO)
    lib.emu_genotype_batch.argtypes = [C.POINTER(abi.LgrParams), C.POINTER(abi.LgrBatchIn), C.POINTER(abi.LgrBatchOut),
                                       C.POINTER(abi.LgrStats), C.c_int]
    lib.emu_genotype_batch.restype = C.c_int
    _lib = lib
    return lib


def emu_genotype(batch, params, arena=1 << 20):
    lib = load()
    res = abi.Result(batch, arena)
    bi, bo = batch.c_struct(), res.c_struct()
    st = abi.LgrStats()
    rc = lib.emu_genotype_batch(C.byref(params), C.byref(bi), C.byref(bo), C.byref(st), 0)
    return rc, res, st


def selfcheck():
    """(failures, co-linear chains checked, closed-form extensions checked, closed-form chain tails checked,
    skipped radix passes checked) so far: the host
    emulation verifies the warp kernels' closed forms against the scalar paths whenever their
    preconditions hold (LGR_CORE_SELFCHECK in lgr_core.cuh)."""
    lib = load()
    out = (C.c_longlong * 5)()
    lib.emu_selfcheck(out)
    return tuple(int(x) for x in out)
