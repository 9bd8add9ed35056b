"""CPU checks of the packed wire format (include/lancet_gpu_realign.h "Packed wire format"):
the library's host packer (lgr_pack_group, csrc/lgr_pack.h) against an independent numpy decoder
written from the header's description.  The device decoder (k_unpack_*) is checked on the GPU by
tests/test_gpu_parity.py::test_packed_path_matches_plain_path."""
import ctypes as C

import numpy as np
import pytest

from lancet2_b200 import abi, synth


@pytest.fixture(scope="module")
def lib():
    return abi.load_library()


def code_of(c: int) -> int:
    """code byte of one base: low nibble = minimap2 seq_nt4_table (A,C,G,T/U = 0..3 in either case,
    everything else 4), high nibble = Lancet2 ENCODE_TABLE (reference scoring_constants.h:48-74:
    A/a C/c G/g T/t = 0..3, everything else — U included — 4)."""
    ch = chr(c).upper() if c < 128 else "?"
    mm = {"A": 0, "C": 1, "G": 2, "T": 3, "U": 3}.get(ch, 4)
    ln = {"A": 0, "C": 1, "G": 2, "T": 3}.get(ch, 4)
    return mm | ln << 4


CODE = np.asarray([code_of(c) for c in range(256)], dtype=np.uint8)


def decode_planes(words: np.ndarray, n_bits: int, length: int) -> np.ndarray:
    """words: [chunks * n_bits] u32, chunk-major; symbol i of the sequence = bits (i % 32) of the
    chunk's n_bits plane words"""
    out = np.zeros(length, dtype=np.uint8)
    for i in range(length):
        c, b = divmod(i, 32)
        v = 0
        for p in range(n_bits):
            v |= ((int(words[c * n_bits + p]) >> b) & 1) << p
        out[i] = v
    return out


def decode_slab(pb: abi.PackedBatch):
    """→ list of dicts (one per group) with hap_codes, read_codes, read_quals (lists of arrays),
    name_hash, var_start/len/allele, mid_occ"""
    slab = pb.slab
    out = []
    for g in range(pb.n_groups):
        d = abi.LgrGroupDir.from_buffer_copy(slab[pb.rec_bytes + 40 * g: pb.rec_bytes + 40 * (g + 1)].tobytes())
        assert d.rec_off % 16 == 0
        rec = slab[d.rec_off:]
        hdr = abi.LgrGroupRecHdr.from_buffer_copy(rec[:64].tobytes())
        assert hdr.magic == abi.LGR_PACK_MAGIC and hdr.rec_bytes % 16 == 0
        P, R, V = d.n_haps, d.n_reads, d.n_vars
        hap_len = rec[hdr.off_hap_len: hdr.off_hap_len + 4 * P].view(np.int32)
        read_len = rec[hdr.off_read_len: hdr.off_read_len + 2 * R].view(np.uint16)
        name_hash = rec[hdr.off_name_hash: hdr.off_name_hash + 4 * R].view(np.uint32).copy()
        vh = V * P
        vs = rec[hdr.off_var: hdr.off_var + 4 * vh].view(np.int32).copy()
        vl = rec[hdr.off_var + 4 * vh: hdr.off_var + 8 * vh].view(np.int32).copy()
        va = rec[hdr.off_var + 8 * vh: hdr.off_var + 9 * vh].view(np.int8).copy()
        assert int(hap_len.sum()) == d.hap_bases and int(read_len.astype(np.int64).sum()) == d.read_bases
        assert d.max_hap_len == (int(hap_len.max()) if P else 0) and d.max_read_len == (int(read_len.max()) if R else 0)
        n_exc = hdr.n_exc
        epos = rec[hdr.off_exc: hdr.off_exc + 4 * n_exc].view(np.uint32)
        ecode = rec[hdr.off_exc + 4 * n_exc: hdr.off_exc + 5 * n_exc]
        lut = np.asarray(list(hdr.qual_lut), dtype=np.uint8)

        def seqs(lens, off_planes):
            res, chunk = [], 0
            for ln in lens.tolist():
                nc = (ln + 31) // 32
                w = rec[off_planes + 8 * chunk: off_planes + 8 * (chunk + nc)].view(np.uint32)
                nt = decode_planes(w, 2, ln)
                res.append((nt * 0x11).astype(np.uint8))
                chunk += nc
            return res
        haps = seqs(hap_len, hdr.off_hap_planes)
        reads = seqs(read_len, hdr.off_read_planes)
        hap_start = np.concatenate([[0], np.cumsum(hap_len)]).astype(np.int64)
        read_start = np.concatenate([[0], np.cumsum(read_len.astype(np.int64))])
        for p, cd in zip(epos.tolist(), ecode.tolist()):
            is_read, pos = p >> 31, p & 0x7FFFFFFF
            starts, arrs = (read_start, reads) if is_read else (hap_start, haps)
            i = int(np.searchsorted(starts, pos, side="right") - 1)
            arrs[i][pos - int(starts[i])] = cd
        quals, chunk, base = [], 0, 0
        for ln in read_len.tolist():
            nc = (ln + 31) // 32
            if hdr.qual_bits == 8:
                quals.append(rec[hdr.off_qual + base: hdr.off_qual + base + ln].copy())
            else:
                nb = hdr.qual_bits
                w = rec[hdr.off_qual + 4 * nb * chunk: hdr.off_qual + 4 * nb * (chunk + nc)].view(np.uint32)
                quals.append(lut[decode_planes(w, nb, ln)])
            chunk += nc
            base += ln
        out.append(dict(haps=haps, reads=reads, quals=quals, name_hash=name_hash, vs=vs, vl=vl, va=va, mid_occ=d.mid_occ,
                        qual_bits=hdr.qual_bits, n_exc=n_exc))
    return out


def check_round_trip(groups, lib):
    pb = abi.PackedBatch(groups, lib)
    dec = decode_slab(pb)
    assert len(dec) == len(groups)
    for g, d in zip(groups, dec):
        assert len(d["haps"]) == len(g.haps) and len(d["reads"]) == len(g.reads)
        for h, got in zip(g.haps, d["haps"]):
            assert np.array_equal(CODE[np.frombuffer(h, dtype=np.uint8)], got)
        for r, q, got, gq in zip(g.reads, g.quals, d["reads"], d["quals"]):
            assert np.array_equal(CODE[np.frombuffer(r, dtype=np.uint8)], got)
            assert np.array_equal(np.frombuffer(q, dtype=np.uint8), gq)
        assert d["name_hash"].tolist() == [abi.x31_hash(n) for n in g.names]
        flat = [x for var in g.variants for x in var]
        assert d["vs"].tolist() == [x[0] for x in flat] and d["vl"].tolist() == [x[1] for x in flat]
        assert d["va"].tolist() == [x[2] for x in flat]
        assert d["mid_occ"] == g.mid_occ
    return pb, dec


def test_synthetic_groups_round_trip(lib):
    groups = synth.make_groups(7, 3, read_len=150, hap_len=700, n_haps=4, n_reads=40)
    pb, dec = check_round_trip(groups, lib)
    assert all(d["qual_bits"] == 2 and d["n_exc"] == 0 for d in dec)  # three Phred bins, plain ACGT
    plain = sum(len(h) for g in groups for h in g.haps) + 2 * sum(len(r) for g in groups for r in g.reads)
    assert pb.slab_bytes < 0.45 * plain  # 2-bit bases + 2-bit quality indices + lengths/hashes


def test_every_byte_value_and_ragged_lengths(lib):
    rng = np.random.default_rng(3)
    haps = [bytes(range(256)) * 2, b"ACGTNacgtnUuRYKM", b""]
    reads, quals = [], []
    for ln in [0, 1, 31, 32, 33, 63, 64, 65, 150, 1024]:
        r = rng.choice(np.frombuffer(b"ACGTNacgtURYn", dtype=np.uint8), size=ln)
        reads.append(r.tobytes())
        quals.append(rng.integers(0, 94, size=ln).astype(np.uint8).tobytes())
    g = abi.Group(haps=haps, reads=reads, quals=quals, names=[f"q{i}" for i in range(len(reads))],
                  variants=[[(5, 1, 0), (5, 2, 1), (-1, 0, -1)]], mid_occ=37)
    _, dec = check_round_trip([g], lib)
    assert dec[0]["qual_bits"] == 8 and dec[0]["n_exc"] > 0


@pytest.mark.parametrize("n_values,bits", [(1, 2), (4, 2), (5, 4), (16, 4), (17, 8)])
def test_quality_dictionary_widths(lib, n_values, bits):
    rng = np.random.default_rng(n_values)
    vals = rng.choice(94, size=n_values, replace=False).astype(np.uint8)
    reads = [rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=ln).tobytes() for ln in (97, 150, 32, 250)]
    quals = []
    for r in reads:
        q = vals[rng.integers(0, n_values, size=len(r))]
        q[:n_values] = vals[:len(q)][:n_values] if len(q) >= n_values else q[:n_values]
        quals.append(q.tobytes())
    quals[1] = (np.resize(vals, 150)).tobytes()  # every value present
    g = abi.Group(haps=[b"ACGT" * 50], reads=reads, quals=quals, names=["a", "b", "c", "d"], variants=[])
    _, dec = check_round_trip([g], lib)
    assert dec[0]["qual_bits"] == bits


def test_limits_and_bad_descriptors(lib):
    g = abi.Group(haps=[b"A" * 100], reads=[b"C" * (abi.LGR_MAX_READ_LEN + 1)], quals=[b"#" * (abi.LGR_MAX_READ_LEN + 1)],
                  names=["x"], variants=[])
    with pytest.raises(ValueError):
        abi.PackedBatch([g], lib)
    assert lib.lgr_packed_group_bytes(None) == 0
    d = abi.LgrGroupDesc()
    d.n_haps, d.n_reads = 0, 1  # reads without a REF haplotype
    assert lib.lgr_packed_group_bytes(C.byref(d)) == 0


def test_empty_batch_and_empty_group(lib):
    pb = abi.PackedBatch([], lib)
    assert pb.n_groups == 0
    g = abi.Group(haps=[b"ACGTACGT"], reads=[], quals=[], names=[], variants=[])
    check_round_trip([g], lib)


def test_every_tail_length_with_short_dictionary(lib):
    """Partial last chunks are read as the last 32 bytes of the string and shifted down (no padded copy): every
    length 1..100, three Phred values (2-bit planes), a non-ACGT base in the very last position of every other read."""
    rng = np.random.default_rng(11)
    vals = np.asarray([2, 23, 37], dtype=np.uint8)
    reads, quals = [], []
    for ln in range(1, 101):
        r = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=ln)
        if ln % 2:
            r[-1] = ord("N")
        reads.append(r.tobytes())
        quals.append(vals[rng.integers(0, 3, size=ln)].tobytes())
    g = abi.Group(haps=[b"ACGTTGCA" * 9 + b"n"], reads=reads, quals=quals, names=[f"t{i}" for i in range(100)], variants=[])
    _, dec = check_round_trip([g, g], lib)
    assert dec[0]["qual_bits"] == 2 and dec[0]["n_exc"] == 51
