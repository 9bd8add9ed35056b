// TEST INFRASTRUCTURE: lancet2_b200/csrc/lgr_format.cuh — the code k_fmt_dedup / k_fmt_metrics run —
// compiled by g++ with the 128 threads of a CTA emulated (lgr_fmt::CtaHost), behind the same
// signature as the C-ABI's lgr_format_metrics, so CPU tests can diff the device arithmetic
// against the reference (oracle/_ref) and the golden vectors without a GPU.
#include <vector>

#include "../../lancet2_b200/csrc/lgr_format.cuh"

static const double kPhred[256] = {
#include "../../lancet2_b200/csrc/phred_lut.inc"
};

// split_tasks != 0: one call per (support, task), as k_fmt_metrics launches them; 0: all tasks in one call
extern "C" int emu_format_metrics_ex(const lgr_evidence_in* in, lgr_format* out, int split_tasks) {
  using namespace lgr_fmt;
  const int S = in->n_supports;
  const int64_t N = in->n_evidence;
  for (int s = 0; s < S; ++s) {
    if (in->sup_n_alleles[s] < 1) return LGR_E_ARG;
    if (in->sup_n_alleles[s] > LGR_FMT_MAX_ALLELES) return LGR_E_LIMIT;
    for (int64_t i = in->sup_begin[s]; i < in->sup_begin[s + 1]; ++i)
      if (in->allele[i] >= in->sup_n_alleles[s]) return LGR_E_ARG;
  }
  std::vector<uint8_t> keep((size_t)N + 1);
  for (int s = 0; s < S; ++s)
    for (int64_t i = in->sup_begin[s]; i < in->sup_begin[s + 1]; ++i)
      keep[(size_t)i] = dedup_keep(in->allele, in->rname_hash, in->sup_begin[s], i);
  Ev e{in->insert_size, in->aln_start, in->aln_score, in->folded_pos, in->rname_hash, in->ref_nm, in->own_hap_nm,
       in->hap_id,      in->allele,    in->flags,     in->base_qual,  in->map_qual,   keep.data()};
  CtaHost w;
  for (int t = 0; t < (split_tasks ? kNumTasks : 1); ++t)
    for (int s = 0; s < S; ++s)
      support_metrics(w, e, in->sup_begin[s], in->sup_begin[s + 1], in->sup_n_alleles[s], in->sup_variant_len[s],
                      in->sup_total_haps[s], kPhred, &out[s], split_tasks ? 1u << t : (unsigned)kTaskAll);
  return LGR_OK;
}

extern "C" int emu_format_metrics(const lgr_evidence_in* in, lgr_format* out) { return emu_format_metrics_ex(in, out, 1); }
