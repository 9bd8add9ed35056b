// TEST INFRASTRUCTURE: C wrappers around the reference's OWN scoring functions,
// compiled unmodified from /root/reference/src (see ../Makefile target `ref`).
// Used only by tests/ to pin oracle/genotype_oracle.cpp and to generate the
// golden vectors under tests/golden/ (tests/golden/gen_scoring_golden.py).
#include <cstdint>
#include <cstring>
#include <string_view>
#include <vector>

#include "lancet/caller/allele_scoring_types.h"
#include "lancet/caller/combined_scorer.h"
#include "lancet/caller/genotyper.h"
#include "lancet/caller/local_scorer.h"
#include "lancet/caller/scoring_constants.h"
#include "lancet/caller/variant_support.h"
#include "lancet/hts/cigar_unit.h"
#include "lancet/hts/cigar_utils.h"
#include "lancet/hts/phred_quality.h"

using lancet::hts::CigarUnit;

static std::vector<CigarUnit> FromBam(const uint32_t* c, int n) {
  std::vector<CigarUnit> v;
  v.reserve(n);
  for (int i = 0; i < n; ++i) v.emplace_back(c[i]);
  return v;
}

extern "C" {

double ref_phred_err(uint32_t q) { return lancet::hts::PhredToErrorProb(q); }

void ref_encode(const char* s, int n, uint8_t* out) {
  auto v = lancet::caller::EncodeSequence(std::string_view(s, (size_t)n));
  std::memcpy(out, v.data(), v.size());
}

uint32_t ref_edit_distance(const uint32_t* cigar, int n, const uint8_t* q, int qn, const uint8_t* t, int tn) {
  return lancet::hts::ComputeEditDistance(FromBam(cigar, n), absl::Span<uint8_t const>(q, qn),
                                          absl::Span<uint8_t const>(t, tn));
}

uint64_t ref_refpos_to_qpos(const uint32_t* cigar, int n, uint64_t ref_pos) {
  return lancet::hts::CigarRefPosToQueryPos(FromBam(cigar, n), (size_t)ref_pos);
}

double ref_softclip_penalty(const uint32_t* cigar, int n) {
  return lancet::caller::ComputeSoftClipPenalty(FromBam(cigar, n));
}

void ref_local_score(const uint32_t* cigar, int n, const uint8_t* q, int qn, const uint8_t* t, int tn,
                     const uint8_t* quals, int qualn, int32_t aln_start, int32_t var_start, int32_t var_len,
                     double* out3, uint8_t* bq) {
  auto const r = lancet::caller::ComputeLocalScore(
      FromBam(cigar, n), absl::Span<uint8_t const>(q, qn), absl::Span<uint8_t const>(t, tn),
      absl::Span<uint8_t const>(quals, qualn), aln_start, var_start, var_len, lancet::caller::SCORING_MATRIX);
  out3[0] = r.mPbqScore, out3[1] = r.mRawScore, out3[2] = r.mIdentity;
  *bq = r.mBaseQual;
}

// ScoreReadAtVariant (combined_scorer.cpp:60-108).  hap = full encoded haplotype.
// out_f[0..3] = local_score, local_identity, folded_pos, CombinedScore();
// out_i[0..5] = global_score, own_hap_nm, hap_id, allele, base_qual
void ref_score_read_at_variant(const uint32_t* cigar, int n, int32_t score, int32_t rs, int32_t re, int hap_idx,
                               const uint8_t* hap, int hap_n, const uint8_t* q, int qn, const uint8_t* quals,
                               int32_t var_start, int32_t var_len, int allele, double* out_f, int64_t* out_i) {
  lancet::caller::Mm2AlnResult aln;
  aln.mCigar = FromBam(cigar, n);
  aln.mScore = score, aln.mRefStart = rs, aln.mRefEnd = re, aln.mHapIdx = (size_t)hap_idx;
  lancet::caller::ReadAlnContext ctx{absl::Span<uint8_t const>(q, qn), absl::Span<uint8_t const>(quals, qn),
                                     (size_t)qn};
  lancet::caller::HapVariantBounds b{var_start, var_len, (lancet::caller::AlleleIndex)allele};
  auto const r = lancet::caller::ScoreReadAtVariant(aln, absl::Span<uint8_t const>(hap, hap_n), ctx, b);
  out_f[0] = r.mLocalScore, out_f[1] = r.mLocalIdentity, out_f[2] = r.mFoldedReadPos, out_f[3] = r.CombinedScore();
  out_i[0] = r.mGlobalScore, out_i[1] = r.mOwnHapNm, out_i[2] = r.mAssignedHaplotypeId, out_i[3] = r.mAllele;
  out_i[4] = r.mBaseQualAtVar;
}

// ComputeHaplotypeEditDistance (combined_scorer.cpp:24-38) for a single alignment on hap_idx
uint32_t ref_hap_edit_distance(const uint32_t* cigar, int n, int32_t rs, int32_t re, int aln_hap, int hap_idx,
                               const uint8_t* hap, int hap_n, const uint8_t* q, int qn) {
  std::vector<lancet::caller::Mm2AlnResult> alns(1);
  alns[0].mCigar = FromBam(cigar, n);
  alns[0].mRefStart = rs, alns[0].mRefEnd = re, alns[0].mHapIdx = (size_t)aln_hap;
  return lancet::caller::ComputeHaplotypeEditDistance(alns, absl::Span<uint8_t const>(hap, hap_n),
                                                      absl::Span<uint8_t const>(q, qn), (size_t)qn,
                                                      (size_t)hap_idx);
}

}  // extern "C"

// The reference's own VariantSupport::AddEvidence (variant_support.cpp:23-67, compiled
// unmodified) over one evidence stream; dumps the per-allele vectors as text.  Pins the
// adapter's AddEvidence restatement (lancet2_b200/host/gpu_genotyper.cpp).
// (plain snprintf/std::string: no iostreams, the host python may carry another libstdc++)
#include <algorithm>
#include <cstdio>
#include <string>
namespace {
void AppI(std::string& s, long long v) { char b[32]; std::snprintf(b, sizeof(b), "%lld,", v); s += b; }
void AppD(std::string& s, double v) { char b[48]; std::snprintf(b, sizeof(b), "%a,", v); s += b; }
}  // namespace
extern "C" int ref_add_evidence_dump(int n, const long long* isize, const long long* start, const double* aln,
                                     const double* fold, const unsigned* hash, const unsigned* ref_nm,
                                     const unsigned* own_nm, const unsigned* hap_id, const unsigned char* allele,
                                     const unsigned char* rev, const unsigned char* bq, const unsigned char* mapq,
                                     const unsigned char* softclip, const unsigned char* proper, char* out, long long cap) {
  using namespace lancet::caller;
  VariantSupport vs;
  for (int i = 0; i < n; ++i) {
    VariantSupport::ReadEvidence ev{};
    ev.mInsertSize = isize[i], ev.mAlignmentStart = start[i], ev.mAlnScore = aln[i], ev.mFoldedReadPos = fold[i];
    ev.mRnameHash = hash[i], ev.mRefNm = ref_nm[i], ev.mOwnHapNm = own_nm[i], ev.mAssignedHaplotypeId = hap_id[i];
    ev.mAllele = allele[i], ev.mStrand = rev[i] ? Strand::REV : Strand::FWD;
    ev.mBaseQual = bq[i], ev.mMapQual = mapq[i], ev.mIsSoftClipped = softclip[i] != 0, ev.mIsProperPair = proper[i] != 0;
    vs.AddEvidence(ev);
  }
  std::string s;
  auto const ad = vs.AlleleData();
  for (std::size_t a = 0; a < ad.size(); ++a) {
    auto const& d = ad[a];
    s += "A" + std::to_string(a) + "|fwdbq:";
    for (auto v : d.mFwdBaseQuals) AppI(s, v);
    s += "|revbq:";
    for (auto v : d.mRevBaseQuals) AppI(s, v);
    s += "|mapq:";
    for (auto v : d.mMapQuals) AppI(s, v);
    s += "|aln:";
    for (auto v : d.mAlnScores) AppD(s, v);
    s += "|isz:";
    for (auto v : d.mProperPairIsizes) AppD(s, v);
    s += "|fold:";
    for (auto v : d.mFoldedReadPositions) AppD(s, v);
    s += "|refnm:";
    for (auto v : d.mRefNmValues) AppD(s, v);
    s += "|ownnm:";
    for (auto v : d.mOwnHapNmValues) AppD(s, v);
    s += "|starts:";
    for (auto v : d.mAlignmentStarts) AppI(s, v);
    s += "|hapids:";
    for (auto v : d.mHaplotypeIds) AppI(s, v);
    s += "|sc:" + std::to_string(d.mSoftClipCount) + "|hashes:";
    std::vector<std::pair<unsigned, int>> hs;
    for (auto const& kv : d.mNameHashes) hs.emplace_back(kv.first, (int)(kv.second == Strand::REV));
    std::sort(hs.begin(), hs.end());
    for (auto& kv : hs) s += std::to_string(kv.first) + ":" + std::to_string(kv.second) + ",";
    s += "\n";
  }
  if ((long long)s.size() + 1 > cap) return -1;
  std::memcpy(out, s.c_str(), s.size() + 1);
  return (int)s.size();
}

// ---- every FORMAT accessor of the reference's VariantSupport on one evidence stream,
// written into the C-ABI's lgr_format so that tests compare field by field.
#include "../../include/lancet_gpu_realign.h"
extern "C" int ref_support_metrics(int n, const long long* isize, const long long* start, const double* aln,
                                   const double* fold, const unsigned* hash, const unsigned* ref_nm,
                                   const unsigned* own_nm, const unsigned* hap_id, const unsigned char* allele,
                                   const unsigned char* flags, const unsigned char* bq, const unsigned char* mapq,
                                   int n_alleles, int variant_len, int total_haps, lgr_format* out) {
  using namespace lancet::caller;
  if (n_alleles < 1 || n_alleles > LGR_FMT_MAX_ALLELES) return -1;
  VariantSupport vs;
  for (int i = 0; i < n; ++i) {
    VariantSupport::ReadEvidence ev{};
    ev.mInsertSize = isize[i], ev.mAlignmentStart = start[i], ev.mAlnScore = aln[i], ev.mFoldedReadPos = fold[i];
    ev.mRnameHash = hash[i], ev.mRefNm = ref_nm[i], ev.mOwnHapNm = own_nm[i], ev.mAssignedHaplotypeId = hap_id[i];
    ev.mAllele = allele[i], ev.mStrand = (flags[i] & LGR_EV_REV) ? Strand::REV : Strand::FWD;
    ev.mBaseQual = bq[i], ev.mMapQual = mapq[i];
    ev.mIsSoftClipped = (flags[i] & LGR_EV_SOFTCLIP) != 0, ev.mIsProperPair = (flags[i] & LGR_EV_PROPER_PAIR) != 0;
    vs.AddEvidence(ev);
  }
  std::memset(out, 0, sizeof(*out));
  out->n_alleles = (uint32_t)n_alleles;
  auto const ad = vs.AlleleData();
  for (int a = 0; a < n_alleles; ++a) {
    auto const idx = static_cast<AlleleIndex>(a);
    out->raw_pbq[a] = vs.RawPosteriorBaseQual(idx), out->rms_mq[a] = vs.RmsMappingQual(idx);
    out->mean_aln[a] = vs.MeanAlnScore(idx);
    out->fwd[a] = (uint32_t)vs.FwdCount(idx), out->rev[a] = (uint32_t)vs.RevCount(idx);
    out->soft_clip[a] = (std::size_t)a < ad.size() ? (uint32_t)ad[a].mSoftClipCount : 0u;
  }
  out->n_kept = (uint32_t)vs.TotalSampleCov();
  out->sb = vs.StrandBiasLogOR(), out->sca = vs.SoftClipAsymmetry();
  auto opt = [&](std::optional<double> v, uint32_t bit, double* dst) {
    if (v.has_value()) out->valid |= bit, *dst = *v;
  };
  opt(vs.FragLengthDelta(), LGR_FMT_HAS_FLD, &out->fld);
  opt(vs.MappingQualCohenD(), LGR_FMT_HAS_MQCD, &out->mqcd);
  opt(vs.ReadPosCohenD(), LGR_FMT_HAS_RPCD, &out->rpcd);
  opt(vs.BaseQualCohenD(), LGR_FMT_HAS_BQCD, &out->bqcd);
  opt(vs.AlleleMismatchDelta((std::size_t)variant_len), LGR_FMT_HAS_ASMD, &out->asmd);
  opt(vs.ComputeFSSE(), LGR_FMT_HAS_FSSE, &out->fsse);
  opt(vs.ComputeAHDD(), LGR_FMT_HAS_AHDD, &out->ahdd);
  opt(vs.ComputeHSE((std::size_t)total_haps), LGR_FMT_HAS_HSE, &out->hse);
  auto const pls = vs.ComputePLs((std::size_t)n_alleles);
  for (std::size_t g = 0; g < pls.size() && g < LGR_FMT_MAX_GENOTYPES; ++g) out->pl[g] = pls[g];
  out->gq = VariantSupport::ComputeGQ(absl::MakeConstSpan(pls.data(), pls.size()));
  auto const lods = vs.ComputeContinuousMixtureLods((std::size_t)n_alleles);
  for (std::size_t a = 0; a < lods.size() && a < LGR_FMT_MAX_ALLELES; ++a) out->cmlod[a] = lods[a];
  return 0;
}

// the same over a whole lgr_evidence_in (S supports) in one call — the CPU baseline of
// tools/bench_format.py times this loop (the reference's own code, one thread).
extern "C" int ref_support_metrics_batch(const lgr_evidence_in* in, lgr_format* out) {
  for (int s = 0; s < in->n_supports; ++s) {
    const long long b = in->sup_begin[s];
    const int n = (int)(in->sup_begin[s + 1] - b);
    const int rc = ref_support_metrics(n, (const long long*)in->insert_size + b, (const long long*)in->aln_start + b,
                                       in->aln_score + b, in->folded_pos + b, in->rname_hash + b, in->ref_nm + b,
                                       in->own_hap_nm + b, in->hap_id + b, in->allele + b, in->flags + b, in->base_qual + b,
                                       in->map_qual + b, in->sup_n_alleles[s], in->sup_variant_len[s], in->sup_total_haps[s],
                                       &out[s]);
    if (rc != 0) return rc;
  }
  return 0;
}
