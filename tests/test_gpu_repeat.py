"""SURVEY.md §8f #3 (first half) — k_repeat_scan through the C-ABI against the CPU oracle (bit-exact:
one byte per job)."""
import numpy as np
import pytest

import repeat_lib
from lancet2_b200 import abi
from lancet2_b200.repeat_scan import GpuRepeatScan

pytestmark = pytest.mark.gpu

ORC = repeat_lib.oracle()


def _expect(jobs):
    return np.array([ORC.orc_has_repeat(s, len(s), k, mm) for s, k, mm in jobs], dtype=np.uint8)


def test_reference_known_answers_as_windows():
    scan = GpuRepeatScan()
    jobs = [(b"ACGT", 4, 0), (b"ACG", 4, 2), (b"", 4, 2), (b"AAAAA", 4, 0), (b"ACGTACGT", 4, 0), (b"ACGTACGA", 4, 0),
            (b"ACGTACGA", 4, 1), (b"ACGTacgt", 4, 0), (b"A" * 33, 32, 0), (b"A" * 32 + b"C", 32, 0), (b"A" * 32 + b"C", 32, 1)]
    got, _ = scan.scan(jobs)
    assert got.tolist() == _expect(jobs).tolist() == [0, 0, 0, 1, 1, 0, 1, 0, 1, 0, 1]


def test_windows_match_oracle():
    scan = GpuRepeatScan()
    jobs = repeat_lib.window_jobs(seed=11, n_windows=60)
    got, ms = scan.scan(jobs)
    want = _expect(jobs)
    assert np.array_equal(got, want)
    assert 0 < want.sum() < len(jobs) and ms > 0


def test_ragged_lengths_and_thresholds():
    rng = np.random.default_rng(5)
    scan = GpuRepeatScan()
    jobs = []
    for n in (1, 2, 31, 32, 33, 63, 64, 65, 95, 96, 97, 255, 256, 257, 1023, 1024, 1025):
        for k in (1, 2, 7, 31, 32, 33, 64):
            seq = repeat_lib.random_window(rng, n, b"ACGT" if (n + k) % 3 else b"AC")
            for mm in (0, 1, 2, 5):
                jobs.append((seq, k, mm))
    got, _ = scan.scan(jobs)
    assert np.array_equal(got, _expect(jobs))


def test_planted_repeat_at_every_mismatch_count():
    rng = np.random.default_rng(9)
    scan = GpuRepeatScan()
    jobs = []
    for k in (25, 61, 101):
        for planted in range(0, 5):
            seq = repeat_lib.plant_repeat(rng, repeat_lib.random_window(rng, 900), k, planted)
            for mm in range(0, 5):
                jobs.append((seq, k, mm))
    got, _ = scan.scan(jobs)
    want = _expect(jobs)
    assert np.array_equal(got, want)
    # a copy with `planted` substitutions is a repeat exactly from max_mismatches = planted upward
    for i, (_, k, mm) in enumerate(jobs):
        planted = (i // 5) % 5
        if mm >= planted:
            assert got[i] == 1


def test_max_length_and_too_long():
    rng = np.random.default_rng(3)
    scan = GpuRepeatScan()
    full = repeat_lib.random_window(rng, abi.LGR_REPEAT_MAX_LEN)
    longer = repeat_lib.random_window(rng, abi.LGR_REPEAT_MAX_LEN + 1)
    tail = repeat_lib.plant_repeat(rng, full, 40, 2, gap=abi.LGR_REPEAT_MAX_LEN - 40)  # the last diagonal alone answers
    jobs = [(full, 40, 2), (tail, 40, 2), (longer, 40, 2), (tail, 40, 1)]
    got, _ = scan.scan(jobs)
    assert scan.last_rc == abi.LGR_E_PARTIAL
    assert got.tolist() == [ORC.orc_has_repeat(full, len(full), 40, 2), 1, abi.LGR_REPEAT_TOO_LONG,
                            ORC.orc_has_repeat(tail, len(tail), 40, 1)]


def test_empty_batch_and_bad_jobs():
    scan = GpuRepeatScan()
    got, _ = scan.scan([])
    assert len(got) == 0 and scan.last_rc == 0
    with pytest.raises(RuntimeError):
        scan.scan([(b"ACGTACGT", 0, 0)])
    with pytest.raises(RuntimeError):
        scan.scan([(b"ACGTACGT", 4, -1)])
