"""CPU checks of the C-ABI library: it builds, loads, exports every symbol the header
declares, and refuses to run (loudly) without a GPU — there is no CPU fallback."""
import ctypes as C
import os
import re

import pytest

from lancet2_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(abi.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    return abi.load_library()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "lancet_gpu_realign.h")).read()
    declared = set(re.findall(r"\b(lgr_[a-z0-9_]+)\s*\(", hdr)) - {"lgr_ctx"}
    assert declared == set(abi.declared_symbols())
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.lgr_abi_version() == 2


def test_struct_layouts_match_header(lib):
    assert C.sizeof(abi.LgrParams) == 4 * 40
    assert abi.ALN_DTYPE.itemsize == 64 and abi.ASSIGN_DTYPE.itemsize == 48
    p = abi.LgrParams()
    lib.lgr_default_params(C.byref(p))
    assert (p.k, p.w, p.a, p.b, p.q, p.e, p.bw, p.end_bonus, p.max_gap, p.max_gap_ref, p.min_dp_max, p.best_n) == \
        (11, 5, 1, 4, 12, 3, 10000, 10000, 200, 5000, 80, 1)
    import oracle_lib as O
    q = O.default_params()
    assert bytes(p) == bytes(q)


def test_x31_hash_matches_python_and_oracle(lib):
    import oracle_lib as O
    for name in ["r0", "g12r511", "A00123:45:HXXXX:1:1101:1000:2000", ""]:
        assert lib.lgr_x31_hash(name.encode()) == abi.x31_hash(name) == O.load_oracle().orc_x31_hash(name.encode())


def test_no_gpu_means_loud_failure(lib):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    ctx = C.c_void_p()
    p = abi.LgrParams()
    lib.lgr_default_params(C.byref(p))
    rc = lib.lgr_create(0, C.byref(p), C.byref(ctx))
    assert rc < 0 and not ctx.value
    assert b"CUDA" in lib.lgr_strerror(rc) or b"device" in lib.lgr_strerror(rc)


def test_bad_params_rejected(lib):
    ctx = C.c_void_p()
    p = abi.LgrParams()
    lib.lgr_default_params(C.byref(p))
    p.end_bonus = 100
    assert lib.lgr_create(0, C.byref(p), C.byref(ctx)) == -1
    assert b"end_bonus" in lib.lgr_last_error(None)


def test_missing_library_raises():
    with pytest.raises(RuntimeError):
        abi.load_library("/nonexistent/liblancet_gpu_realign.so")
