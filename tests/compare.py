"""Field-by-field comparison of two Result objects (oracle vs device / host-emu)."""
import numpy as np

from lancet2_b200 import abi

ALN_FIELDS = ["valid", "score", "rs", "re", "qs", "qe", "rev", "dp_score", "dp_max", "mlen", "blen", "n_ambi", "nm",
              "n_cigar", "n_regs"]
ASG_FIELDS = ["local_score", "local_identity", "folded_read_pos", "global_score", "ref_nm", "own_hap_nm", "hap_id",
              "allele", "base_qual", "assigned"]


def compare_results(batch: abi.Batch, want: abi.Result, got: abi.Result, max_report: int = 10, check_aln=True):
    """returns a list of human-readable mismatch strings (empty = bit-exact)"""
    errs = []
    n = batch.n_pairs
    if check_aln:
        for f in ALN_FIELDS:
            a, b = want.aln[f][:n], got.aln[f][:n]
            bad = np.nonzero(a != b)[0]
            for i in bad[:max_report]:
                errs.append(f"aln[{i}].{f}: want {a[i]} got {b[i]}  (want cigar {want.cigar_string(int(i))} got {got.cigar_string(int(i))})")
        both = (want.aln["valid"][:n] == 1) & (got.aln["valid"][:n] == 1) & (want.aln["n_cigar"][:n] == got.aln["n_cigar"][:n])
        # inline cigars (the common case) in one vectorised sweep; arena cigars one by one
        w_in = both & (want.aln["cigar_off"][:n] < 0) & (got.aln["cigar_off"][:n] < 0)
        wi = want.cigar_inline[:n * abi.LGR_CIGAR_INLINE].reshape(n, abi.LGR_CIGAR_INLINE)
        gi = got.cigar_inline[:n * abi.LGR_CIGAR_INLINE].reshape(n, abi.LGR_CIGAR_INLINE)
        live = np.arange(abi.LGR_CIGAR_INLINE)[None, :] < want.aln["n_cigar"][:n, None]
        bad_inline = np.nonzero(w_in & ((wi != gi) & live).any(axis=1))[0]
        rest = np.nonzero(both & ~w_in)[0]
        nbad = 0
        for i in list(bad_inline) + [int(i) for i in rest if want.cigar(int(i)) != got.cigar(int(i))]:
            nbad += 1
            if nbad <= max_report:
                errs.append(f"cigar[{i}]: want {want.cigar_string(int(i))} got {got.cigar_string(int(i))}")
    m = batch.n_assign
    for f in ASG_FIELDS:
        a, b = want.assign[f][:m], got.assign[f][:m]
        if a.dtype.kind == "f":
            bad = np.nonzero(a.view(np.uint64) != b.view(np.uint64))[0]
        else:
            bad = np.nonzero(a != b)[0]
        for i in bad[:max_report]:
            errs.append(f"assign[{i}].{f}: want {a[i]!r} got {b[i]!r}")
    return errs
