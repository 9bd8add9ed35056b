// lgr_format.cuh — VariantSupport aggregation + FORMAT math for one support (one variant x one
// sample), written ONCE for the device warp (lgr_format.cu) and for the g++ host emulation
// (tests/hostemu/format_emu.cpp).  SURVEY.md §8f #2.
//
// Mapping: one 128-thread CTA per (support, task).  Every reduction over the support's evidence
// records is a thread-strided partial (thread t takes records t, t+128, ...) followed by an
// xor-butterfly inside each warp and a fixed-order sum of the four warp totals, so all threads
// end up with the same bits; the scalar math after a reduction is executed uniformly by all
// threads and only the leader stores.  The `W` policy supplies `reduce` (shuffles + shared
// memory on the device, a loop over 128 emulated threads on the host), so host and device run
// the same arithmetic in the same order and differ only in libm (log10/log2/log/lgamma/pow).
//
// Reference being restated (file:line in the reference tree, src/lancet/...):
//   caller/variant_support.cpp:23-67   AddEvidence (first-seen dedup by read-name hash per allele)
//   caller/variant_support.cpp:140-335 the accessors; caller/variant_support.h:362-412 the three
//                                      templates (effect size, mean delta, pooled entropy)
//   base/mann_whitney.h:13-65          MannWhitneyEffectSize
//   caller/posterior_base_qual.cpp:13-40, caller/genotype_likelihood.cpp:29-205
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/lancet_gpu_realign.h"

#if defined(__CUDACC__)
#define FMT_HD __host__ __device__ __forceinline__
#else
#define FMT_HD inline
#endif

namespace lgr_fmt {

constexpr int kLanes = 32;
constexpr int kWarps = 4;
constexpr int kThreads = kLanes * kWarps;  // threads cooperating on one (support, task)

// SoA views of one batch of evidence (device pointers on the device, host pointers in the emulation)
struct Ev {
  const int64_t* insert_size;
  const int64_t* aln_start;
  const double* aln_score;
  const double* folded_pos;
  const uint32_t* rname_hash;
  const uint32_t* ref_nm;
  const uint32_t* own_hap_nm;
  const uint32_t* hap_id;
  const uint8_t* allele;
  const uint8_t* flags;
  const uint8_t* base_qual;
  const uint8_t* map_qual;
  const uint8_t* keep;  // result of dedup_keep for every record
};

template <int ND, int NI>
struct Acc {
  double d[ND > 0 ? ND : 1];
  long long i[NI > 0 ? NI : 1];
  FMT_HD Acc() {
    for (int k = 0; k < (ND > 0 ? ND : 1); ++k) d[k] = 0.0;
    for (int k = 0; k < (NI > 0 ? NI : 1); ++k) i[k] = 0;
  }
};

// AddEvidence's try_emplace (variant_support.cpp:28-29): record i survives iff no earlier record
// of the same support carries the same (allele, read-name hash).
FMT_HD uint8_t dedup_keep(const uint8_t* allele, const uint32_t* hash, int64_t begin, int64_t i) {
  const uint8_t a = allele[i];
  const uint32_t h = hash[i];
  for (int64_t j = begin; j < i; ++j)
    if (hash[j] == h && allele[j] == a) return 0;
  return 1;
}

// base/mann_whitney.h:47-64 from the exact integer rank statistics:
// r2 = 2 x (sum of ALT mid-ranks), tie = sum over tie groups of t^3 - t.
FMT_HD double mw_effect(long long r2, long long tie, long long n_ref_i, long long n_alt_i) {
  const double n_ref = (double)n_ref_i, n_alt = (double)n_alt_i;
  const double alt_rank_sum = (double)r2 / 2.0;
  const double tie_correction = (double)tie;
  const double u_stat = alt_rank_sum - ((n_alt * (n_alt + 1.0)) / 2.0);
  const double mean_u = (n_ref * n_alt) / 2.0;
  const double n_total = (double)(n_ref_i + n_alt_i);
  const double var_u = (n_ref * n_alt / 12.0) * ((n_total + 1.0) - (tie_correction / (n_total * (n_total - 1.0))));
  if (var_u <= 0.0) return 0.0;
  const double z_score = (u_stat - mean_u) / sqrt(var_u);
  return z_score / sqrt(n_total);
}

// Mann-Whitney over a byte-valued field: each thread owns kValsPerThread of the 256 values and counts,
// in one pass over the records, how many REF / ALT records equal each of them and how many are smaller.
constexpr int kValsPerThread = 256 / kThreads;
template <class W, class Get>
FMT_HD double mw_bytes(W& w, const Ev& e, int64_t b, int64_t n_end, long long n_ref, long long n_alt, const Get& get) {
  Acc<0, 2> r = w.template reduce<Acc<0, 2>>([&](int lane) {
    int eq_ref[kValsPerThread], eq_alt[kValsPerThread], less[kValsPerThread];
#pragma unroll
    for (int t = 0; t < kValsPerThread; ++t) eq_ref[t] = eq_alt[t] = less[t] = 0;
    const int v0 = lane * kValsPerThread;
    for (int64_t j = b; j < n_end; ++j) {
      if (!e.keep[j]) continue;
      const int v = get(j);
      const int alt = e.allele[j] != 0;
#pragma unroll
      for (int t = 0; t < kValsPerThread; ++t) {
        const int eq = v == v0 + t;
        eq_ref[t] += eq & (alt ^ 1);
        eq_alt[t] += eq & alt;
        less[t] += v < v0 + t;
      }
    }
    Acc<0, 2> a;
#pragma unroll
    for (int t = 0; t < kValsPerThread; ++t) {
      const long long ts = (long long)eq_ref[t] + eq_alt[t];
      if (ts == 0) continue;
      a.i[0] += (long long)eq_alt[t] * (2LL * less[t] + ts + 1);  // 2 x mid-rank = i + 1 + jdx
      a.i[1] += ts * ts * ts - ts;
    }
    return a;
  });
  return mw_effect(r.i[0], r.i[1], n_ref, n_alt);
}

// Mann-Whitney over the f64 folded read positions: rank by counting (no sort, no scratch).
template <class W>
FMT_HD double mw_folded(W& w, const Ev& e, int64_t b, int64_t n_end, long long n_ref, long long n_alt) {
  Acc<0, 2> r = w.template reduce<Acc<0, 2>>([&](int lane) {
    Acc<0, 2> a;
    for (int64_t i = b + lane; i < n_end; i += kThreads) {
      if (!e.keep[i]) continue;
      const double x = e.folded_pos[i];
      long long less = 0, eq = 0;
      for (int64_t j = b; j < n_end; ++j) {
        if (!e.keep[j]) continue;
        const double y = e.folded_pos[j];
        less += y < x;
        eq += y == x;
      }
      if (e.allele[i] != 0) a.i[0] += 2 * less + eq + 1;
      a.i[1] += eq * eq - 1;  // summed over the eq members of a tie group: t^3 - t
    }
    return a;
  });
  return mw_effect(r.i[0], r.i[1], n_ref, n_alt);
}

// variant_support.h:386-412 AltPooledEntropy: normalised Shannon entropy of the pooled ALT
// records' bins, bin = val / width with C truncation (width 3: FSSE's start bins, 1: HSE's
// haplotype ids).  One lane per ALT record counts the record's bin mates; a bin contributes its
// term at its first record.  The bin test inside the scan is a range test on the raw value (one
// division per record, none per pair).  Measured and rejected (profiles/README.md): a
// warp-uniform outer loop with the lanes sharing each scan — one dependent butterfly per record
// costs more than the idle REF lanes do.
template <class W, class Val>
FMT_HD double alt_entropy(W& w, const Ev& e, int64_t b, int64_t n_end, long long n_alt, double max_bins, long long width,
                          const Val& val) {
  const double total = (double)n_alt;
  Acc<1, 0> r = w.template reduce<Acc<1, 0>>([&](int lane) {
    Acc<1, 0> a;
    for (int64_t i = b + lane; i < n_end; i += kThreads) {
      if (!e.keep[i] || e.allele[i] == 0) continue;
      const long long key = val(i) / width;
      const long long lo = key > 0 ? key * width : key * width - (width - 1);
      const long long hi = key < 0 ? key * width : key * width + (width - 1);
      int count = 0, earlier = 0;
      for (int64_t j = b; j < n_end; ++j) {
        const long long v = val(j);
        const int hit = (e.keep[j] != 0) & (e.allele[j] != 0) & (v >= lo) & (v <= hi);
        count += hit;
        earlier += hit & (j < i);
      }
      if (earlier) continue;  // not the bin's first record
      const double prob = (double)count / total;
      a.d[0] -= prob * log2(prob);
    }
    return a;
  });
  const double max_entropy = log2(total < max_bins ? total : max_bins);
  return max_entropy > 0.0 ? (r.d[0] / max_entropy) : 0.0;
}

#ifdef LGR_FMT_SORT
// ---- sort instead of scan — compiled only with -DLGR_FMT_SORT: checked on the
// CPU against the scan build and the reference, run on a GPU for one shape only, therefore not
// yet the default device build.  Supports of up to kSortCap records sort their keys in shared memory
// with a bitonic network (one `each` phase per stage: the compare-exchanges of a stage are
// disjoint, so the host emulation may run them thread by thread); larger supports keep the scan.
constexpr int kSortCap = 2048;
constexpr uint64_t kSortSentinel = ~0ull;  // dropped records sort behind every real key

FMT_HD uint64_t ordered_f64(double x) {  // u64 image with the order of the doubles; -0.0 and +0.0 tie
  if (x == 0.0) x = 0.0;
  uint64_t u;
  memcpy(&u, &x, sizeof u);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
FMT_HD uint64_t ordered_i64(long long v) { return (uint64_t)v ^ 0x8000000000000000ull; }

template <class W, class Key>
FMT_HD int fill_and_sort(W& w, int n, const Key& key) {  // key(idx) → sentinel for records that do not take part
  uint64_t* keys = w.sort_keys();
  uint8_t* tags = w.sort_tags();
  int m = 2;
  while (m < n) m <<= 1;
  w.each([&](int t) {
    for (int idx = t; idx < m; idx += kThreads) {
      uint8_t tag = 0;
      keys[idx] = idx < n ? key(idx, &tag) : kSortSentinel;
      tags[idx] = tag;
    }
  });
  for (int k = 2; k <= m; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1)
      w.each([&](int t) {
        for (int idx = t; idx < m; idx += kThreads) {
          const int partner = idx ^ j;
          if (partner <= idx) continue;
          const bool up = (idx & k) == 0;
          const uint64_t a = keys[idx], c = keys[partner];
          if ((a > c) == up && a != c) {
            keys[idx] = c, keys[partner] = a;
            const uint8_t ta = tags[idx];
            tags[idx] = tags[partner], tags[partner] = ta;
          }
        }
      });
  return m;
}

FMT_HD int lower_bound_u64(const uint64_t* keys, int n, uint64_t x) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (keys[mid] < x) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}
FMT_HD int upper_bound_u64(const uint64_t* keys, int n, uint64_t x) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (keys[mid] <= x) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

template <class W>
FMT_HD double mw_folded_sorted(W& w, const Ev& e, int64_t b, int64_t n_end, long long n_ref, long long n_alt) {
  const int n = (int)(n_end - b), n_kept = (int)(n_ref + n_alt);
  fill_and_sort(w, n, [&](int idx, uint8_t* tag) {
    const int64_t i = b + idx;
    if (!e.keep[i]) return kSortSentinel;
    *tag = e.allele[i] != 0;
    return ordered_f64(e.folded_pos[i]);
  });
  const uint64_t* keys = w.sort_keys();
  const uint8_t* tags = w.sort_tags();
  Acc<0, 2> r = w.template reduce<Acc<0, 2>>([&](int lane) {
    Acc<0, 2> a;
    for (int p = lane; p < n_kept; p += kThreads) {
      const uint64_t x = keys[p];
      const long long lo = lower_bound_u64(keys, n_kept, x), hi = upper_bound_u64(keys, n_kept, x);
      if (tags[p]) a.i[0] += lo + hi + 1;  // 2 x mid-rank = i + 1 + jdx
      if (p == lo) {
        const long long ts = hi - lo;
        a.i[1] += ts * ts * ts - ts;
      }
    }
    return a;
  });
  return mw_effect(r.i[0], r.i[1], n_ref, n_alt);
}

template <class W, class Val>
FMT_HD double alt_entropy_sorted(W& w, const Ev& e, int64_t b, int64_t n_end, long long n_alt, double max_bins, long long width,
                                 const Val& val) {
  const int n = (int)(n_end - b), na = (int)n_alt;
  fill_and_sort(w, n, [&](int idx, uint8_t*) {
    const int64_t i = b + idx;
    if (!e.keep[i] || e.allele[i] == 0) return kSortSentinel;
    return ordered_i64(val(i) / width);
  });
  const uint64_t* keys = w.sort_keys();
  const double total = (double)n_alt;
  Acc<1, 0> r = w.template reduce<Acc<1, 0>>([&](int lane) {
    Acc<1, 0> a;
    for (int p = lane; p < na; p += kThreads) {
      const uint64_t x = keys[p];
      if (p > 0 && keys[p - 1] == x) continue;  // not the bin's first key
      const double prob = (double)(upper_bound_u64(keys, na, x) - p) / total;
      a.d[0] -= prob * log2(prob);
    }
    return a;
  });
  const double max_entropy = log2(total < max_bins ? total : max_bins);
  return max_entropy > 0.0 ? (r.d[0] / max_entropy) : 0.0;
}
#endif  // LGR_FMT_SORT

// genotype_likelihood.cpp:29-46 LogDirichletMultinomial
FMT_HD double log_dm(const int* counts, const double* alphas, int K) {
  double log_prob = 0.0, alpha_sum = 0.0, count_alpha_sum = 0.0;
  for (int k = 0; k < K; ++k) {
    const double c = (double)counts[k], al = alphas[k];
    log_prob += lgamma(c + al) - lgamma(al);
    alpha_sum += al;
    count_alpha_sum += c + al;
  }
  log_prob += lgamma(alpha_sum) - lgamma(count_alpha_sum);
  return log_prob;
}

// genotype_likelihood.cpp:109-163 ComputeGenotypePLs + NormalizeToPLs + ComputeGenotypeQuality
FMT_HD void genotype_pls(const int* counts, int K, uint32_t* pl, uint32_t* gq) {
  const double kBackground = 0.005, kOverdispersion = 0.01, kAlphaFloor = 1e-6;
  const double precision = (1.0 - kOverdispersion) / kOverdispersion;
  double ll[LGR_FMT_MAX_GENOTYPES];
  int g = 0;
  for (int allele_b = 0; allele_b < K; ++allele_b) {
    for (int allele_a = 0; allele_a <= allele_b; ++allele_a) {
      double alphas[LGR_FMT_MAX_ALLELES];
      const double main_mass = 1.0 - kBackground;
      for (int k = 0; k < K; ++k) {
        double mu = kBackground / K;
        if (allele_a == allele_b) {
          if (k == allele_a) mu += main_mass;
        } else if (k == allele_a || k == allele_b) {
          mu += main_mass / 2.0;
        }
        const double al = mu * precision;
        alphas[k] = al > kAlphaFloor ? al : kAlphaFloor;
      }
      ll[g++] = log_dm(counts, alphas, K);
    }
  }
  const int G = g;
  double best = ll[0];
  for (int i = 1; i < G; ++i) best = ll[i] > best ? ll[i] : best;
  const double kPlCap = 4294967295.0 / 2.0, kLn10 = 2.302585092994045684;
  uint32_t min1 = 0xffffffffu, min2 = 0xffffffffu;
  for (int i = 0; i < G; ++i) {
    const double raw = -10.0 * (ll[i] - best) / kLn10;
    const uint32_t v = (uint32_t)round(raw < kPlCap ? raw : kPlCap);
    pl[i] = v;
    if (v < min1) {
      min2 = min1;
      min1 = v;
    } else if (v < min2) {
      min2 = v;
    }
  }
  for (int i = G; i < LGR_FMT_MAX_GENOTYPES; ++i) pl[i] = 0;
  const uint32_t d = min2 - min1;
  *gq = G < 2 ? 0u : (d < 99u ? d : 99u);
}

// genotype_likelihood.cpp:75-84 PileupLogLikelihood under one allele-fraction vector
template <class W>
FMT_HD double pileup_loglk(W& w, const Ev& e, int64_t b, int64_t n_end, int K, const double* frac, const double* phred) {
  double log_lk = 0.0;
  const double denom = (double)(K - 1 > 1 ? K - 1 : 1);
  for (int called_as = 0; called_as < K; ++called_as) {
    const double f = frac[called_as];
    Acc<2, 0> r = w.template reduce<Acc<2, 0>>([&](int lane) {
      Acc<2, 0> a;
      for (int64_t i = b + lane; i < n_end; i += kThreads) {
        if (!e.keep[i] || e.allele[i] != called_as) continue;
        const double error_prob = phred[e.base_qual[i]];
        const double mismatch_prob = error_prob / denom;
        const double match_bonus = (1.0 - error_prob) - mismatch_prob;
        const double prob_read = mismatch_prob + (f * match_bonus);
        const double term = log10(prob_read > 1e-15 ? prob_read : 1e-15);
        if (e.flags[i] & LGR_EV_REV) a.d[1] += term;
        else a.d[0] += term;
      }
      return a;
    });
    log_lk += r.d[0];
    log_lk += r.d[1];
  }
  return log_lk;
}

// The FORMAT quantities of support [b, n_end) fall into independent tasks; the device runs one
// warp per (support, task) so that a batch of a few hundred supports still fills the SMs, the
// host emulation runs them all in one call.  Every task writes its own fields of `out`.
enum : unsigned {
  kTaskStats = 1u,   // per-allele statistics, SB, SCA, FLD, ASMD, AHDD, PL, GQ, valid bits
  kTaskMqcd = 2u,
  kTaskBqcd = 4u,
  kTaskRpcd = 8u,
  kTaskFsse = 16u,
  kTaskHse = 32u,
  kTaskCmlod = 64u,
  kTaskAll = 127u,
};
constexpr int kNumTasks = 7;

// `phred`: the 256-entry PhredToErrorProb table.
template <class W>
FMT_HD void support_metrics(W& w, const Ev& e, int64_t b, int64_t n_end, int K, int variant_len, int total_haps,
                            const double* phred, lgr_format* out, unsigned tasks = kTaskAll) {
  const bool lead = w.leader();
  // allele depths after the dedup (TotalAlleleCov) — every task needs them
  int cov[LGR_FMT_MAX_ALLELES];
  {
    Acc<0, LGR_FMT_MAX_ALLELES> c = w.template reduce<Acc<0, LGR_FMT_MAX_ALLELES>>([&](int lane) {
      Acc<0, LGR_FMT_MAX_ALLELES> acc;
      for (int64_t i = b + lane; i < n_end; i += kThreads) {
        if (!e.keep[i]) continue;
        const int al = e.allele[i];
#pragma unroll
        for (int a = 0; a < LGR_FMT_MAX_ALLELES; ++a) acc.i[a] += al == a;
      }
      return acc;
    });
    for (int a = 0; a < LGR_FMT_MAX_ALLELES; ++a) cov[a] = (int)c.i[a];
  }
  const long long n_ref = cov[0];
  long long n_alt = 0;
  for (int a = 1; a < K; ++a) n_alt += cov[a];

  if (tasks & kTaskStats) {
    long long ref_fwd = 0, ref_rev = 0, alt_fwd = 0, alt_rev = 0, ref_sc = 0, alt_sc = 0;
    long long isz_n[2] = {0, 0}, isz_sum[2] = {0, 0}, refnm_sum[2] = {0, 0}, ownnm_sum[2] = {0, 0};
    if (lead) {
      out->n_alleles = (uint32_t)K;
      for (int a = K; a < LGR_FMT_MAX_ALLELES; ++a) {
        out->raw_pbq[a] = out->rms_mq[a] = out->mean_aln[a] = out->cmlod[a] = 0.0;
        out->fwd[a] = out->rev[a] = out->soft_clip[a] = 0;
      }
    }
    // ---- per-allele statistics (PerAlleleData vectors, variant_support.cpp:140-209) ----
    for (int a = 0; a < K; ++a) {
      Acc<3, 8> r = w.template reduce<Acc<3, 8>>([&](int lane) {
        Acc<3, 8> acc;
        for (int64_t i = b + lane; i < n_end; i += kThreads) {
          if (!e.keep[i] || e.allele[i] != a) continue;
          const unsigned fl = e.flags[i];
          if (fl & LGR_EV_REV) acc.i[1] += 1;
          else acc.i[0] += 1;
          acc.i[2] += (fl & LGR_EV_SOFTCLIP) ? 1 : 0;
          const long long mq = e.map_qual[i];
          acc.i[3] += mq * mq;
          if ((fl & LGR_EV_PROPER_PAIR) && e.insert_size[i] != 0) acc.i[4] += 1, acc.i[5] += e.insert_size[i];
          acc.i[6] += e.ref_nm[i];
          acc.i[7] += e.own_hap_nm[i];
          acc.d[0] += e.aln_score[i];
          const double eps = phred[e.base_qual[i]];
          const double ok = 1.0 - eps;
          acc.d[1] += log10(eps > 1e-300 ? eps : 1e-300);
          acc.d[2] += log10(ok > 1e-300 ? ok : 1e-300);
        }
        return acc;
      });
      const long long cnt = r.i[0] + r.i[1];
      const int grp = a == 0 ? 0 : 1;
      if (a == 0) ref_fwd = r.i[0], ref_rev = r.i[1], ref_sc = r.i[2];
      else alt_fwd += r.i[0], alt_rev += r.i[1], alt_sc += r.i[2];
      isz_n[grp] += r.i[4], isz_sum[grp] += r.i[5], refnm_sum[grp] += r.i[6], ownnm_sum[grp] += r.i[7];
      double pbq = 0.0;
      if (cnt > 0) {  // posterior_base_qual.cpp:13-40
        const double log_err = r.d[1], log_ok = r.d[2];
        const double max_log = log_err > log_ok ? log_err : log_ok;
        const double min_log = log_err < log_ok ? log_err : log_ok;
        const double log_sum = max_log + log10(1.0 + pow(10.0, min_log - max_log));
        pbq = -10.0 * (log_err - log_sum);
      }
      if (lead) {
        out->fwd[a] = (uint32_t)r.i[0], out->rev[a] = (uint32_t)r.i[1], out->soft_clip[a] = (uint32_t)r.i[2];
        out->rms_mq[a] = cnt > 0 ? sqrt((double)r.i[3] / (double)cnt) : 0.0;
        out->mean_aln[a] = cnt > 0 ? r.d[0] / (double)cnt : 0.0;
        out->raw_pbq[a] = pbq;
      }
    }
    uint32_t valid = 0;
    // ---- StrandBiasLogOR / SoftClipAsymmetry (variant_support.cpp:182-225) ----
    const double sb = log(((double)(ref_fwd + 1) * (double)(alt_rev + 1)) / ((double)(ref_rev + 1) * (double)(alt_fwd + 1)));
    const double alt_frac = n_alt > 0 ? (double)alt_sc / (double)n_alt : 0.0;
    const double ref_frac = n_ref > 0 ? (double)ref_sc / (double)n_ref : 0.0;
    const double sca = alt_frac - ref_frac;
    // ---- MeanAltMinusRef users: FLD, ASMD, AHDD (variant_support.h:362-384) ----
    double fld = 0.0, asmd = 0.0, ahdd = 0.0;
    if (isz_n[0] > 0 && isz_n[1] > 0) {
      valid |= LGR_FMT_HAS_FLD;
      fld = ((double)isz_sum[1] / (double)isz_n[1] - 0.0) - (double)isz_sum[0] / (double)isz_n[0];
    }
    if (n_ref > 0 && n_alt > 0) {
      valid |= LGR_FMT_HAS_ASMD | LGR_FMT_HAS_AHDD | LGR_FMT_HAS_MQCD | LGR_FMT_HAS_RPCD | LGR_FMT_HAS_BQCD;
      asmd = ((double)refnm_sum[1] / (double)n_alt - (double)variant_len) - (double)refnm_sum[0] / (double)n_ref;
      ahdd = ((double)ownnm_sum[1] / (double)n_alt - 0.0) - (double)ownnm_sum[0] / (double)n_ref;
    }
    if (n_alt >= 3) valid |= LGR_FMT_HAS_FSSE | (total_haps >= 2 ? LGR_FMT_HAS_HSE : 0u);
    // ---- PL / GQ from the allele depths (variant_support.cpp:294-310) ----
    uint32_t pl[LGR_FMT_MAX_GENOTYPES], gq = 0;
    genotype_pls(cov, K, pl, &gq);
    if (lead) {
      out->sb = sb, out->sca = sca, out->fld = fld, out->asmd = asmd, out->ahdd = ahdd;
      for (int g = 0; g < LGR_FMT_MAX_GENOTYPES; ++g) out->pl[g] = pl[g];
      out->gq = gq, out->valid = valid, out->n_kept = (uint32_t)(n_ref + n_alt);
    }
  }
  // ---- Mann-Whitney effect sizes: MQCD, BQCD, RPCD (variant_support.cpp:234-263) ----
  const bool both = n_ref > 0 && n_alt > 0;
  if (tasks & kTaskMqcd) {
    const double v = both ? mw_bytes(w, e, b, n_end, n_ref, n_alt, [&](int64_t j) { return (int)e.map_qual[j]; }) : 0.0;
    if (lead) out->mqcd = v;
  }
  if (tasks & kTaskBqcd) {
    const double v = both ? mw_bytes(w, e, b, n_end, n_ref, n_alt, [&](int64_t j) { return (int)e.base_qual[j]; }) : 0.0;
    if (lead) out->bqcd = v;
  }
  if (tasks & kTaskRpcd) {
#ifdef LGR_FMT_SORT
    const double v = !both ? 0.0 : (n_end - b <= kSortCap ? mw_folded_sorted(w, e, b, n_end, n_ref, n_alt) : mw_folded(w, e, b, n_end, n_ref, n_alt));
#else
    const double v = both ? mw_folded(w, e, b, n_end, n_ref, n_alt) : 0.0;
#endif
    if (lead) out->rpcd = v;
  }
  // ---- pooled-ALT entropies: FSSE (3 bp start bins, <= 20 bins) and HSE (variant_support.cpp:270-291) ----
  if (tasks & kTaskFsse) {
    auto start_of = [&](int64_t j) { return (long long)e.aln_start[j]; };
#ifdef LGR_FMT_SORT
    const double v = n_alt < 3 ? 0.0 : (n_end - b <= kSortCap ? alt_entropy_sorted(w, e, b, n_end, n_alt, 20.0, 3, start_of)
                                                             : alt_entropy(w, e, b, n_end, n_alt, 20.0, 3, start_of));
#else
    const double v = n_alt >= 3 ? alt_entropy(w, e, b, n_end, n_alt, 20.0, 3, start_of) : 0.0;
#endif
    if (lead) out->fsse = v;
  }
  if (tasks & kTaskHse) {
    auto hap_of = [&](int64_t j) { return (long long)e.hap_id[j]; };
#ifdef LGR_FMT_SORT
    const double v = !(n_alt >= 3 && total_haps >= 2)
                         ? 0.0
                         : (n_end - b <= kSortCap ? alt_entropy_sorted(w, e, b, n_end, n_alt, (double)total_haps, 1, hap_of)
                                                  : alt_entropy(w, e, b, n_end, n_alt, (double)total_haps, 1, hap_of));
#else
    const double v = n_alt >= 3 && total_haps >= 2 ? alt_entropy(w, e, b, n_end, n_alt, (double)total_haps, 1, hap_of) : 0.0;
#endif
    if (lead) out->hse = v;
  }
  // ---- CMLOD (genotype_likelihood.cpp:165-205) ----
  if (tasks & kTaskCmlod) {
    double lod[LGR_FMT_MAX_ALLELES];
    for (int a = 0; a < LGR_FMT_MAX_ALLELES; ++a) lod[a] = 0.0;
    long long total_depth = 0;
    for (int a = 0; a < K; ++a) total_depth += cov[a];
    if (K >= 2 && total_depth > 0) {
      double frac_mle[LGR_FMT_MAX_ALLELES], frac_null[LGR_FMT_MAX_ALLELES];
      for (int a = 0; a < K; ++a) frac_mle[a] = (double)cov[a] / (double)total_depth;
      const double ll_mle = pileup_loglk(w, e, b, n_end, K, frac_mle, phred);
      for (int t = 1; t < K; ++t) {
        if (cov[t] == 0) continue;
        for (int a = 0; a < K; ++a) frac_null[a] = frac_mle[a];
        const double null_mass = frac_null[t];
        frac_null[t] = 0.0;
        const double remaining = 1.0 - null_mass;
        if (remaining <= 0.0) {
          frac_null[0] = 1.0;
        } else {
          for (int a = 0; a < K; ++a) frac_null[a] /= remaining;
        }
        const double ll_null = pileup_loglk(w, e, b, n_end, K, frac_null, phred);
        const double dlt = ll_mle - ll_null;
        lod[t] = dlt > 0.0 ? dlt : 0.0;
      }
    }
    if (lead)
      for (int a = 0; a < K; ++a) out->cmlod[a] = lod[a];
  }
}

// the reduction on 128 emulated threads (host) — bit-identical to the device version: the
// xor-butterfly inside each group of 32 (IEEE addition is commutative, so at every level both
// partners add the same two numbers), then (w0 + w1) + (w2 + w3) over the four warp totals.
struct CtaHost {
  FMT_HD bool leader() const { return true; }
#ifdef LGR_FMT_SORT
  uint64_t keys_[kSortCap];
  uint8_t tags_[kSortCap];
  uint64_t* sort_keys() { return keys_; }
  uint8_t* sort_tags() { return tags_; }
  template <class F>
  inline void each(const F& f) {  // one phase: every thread once, then a barrier
    for (int t = 0; t < kThreads; ++t) f(t);
  }
#endif
  template <class A, class F>
  inline A reduce(const F& f) {
    A p[kThreads], q[kThreads];
    for (int l = 0; l < kThreads; ++l) p[l] = f(l);
    constexpr int nd = (int)(sizeof(p[0].d) / sizeof(double)), ni = (int)(sizeof(p[0].i) / sizeof(long long));
    for (int off = kLanes / 2; off > 0; off >>= 1) {
      for (int l = 0; l < kThreads; ++l) {
        for (int k = 0; k < nd; ++k) q[l].d[k] = p[l].d[k] + p[l ^ off].d[k];
        for (int k = 0; k < ni; ++k) q[l].i[k] = p[l].i[k] + p[l ^ off].i[k];
      }
      for (int l = 0; l < kThreads; ++l) p[l] = q[l];
    }
    A r;
    for (int k = 0; k < nd; ++k) r.d[k] = (p[0].d[k] + p[kLanes].d[k]) + (p[2 * kLanes].d[k] + p[3 * kLanes].d[k]);
    for (int k = 0; k < ni; ++k) r.i[k] = (p[0].i[k] + p[kLanes].i[k]) + (p[2 * kLanes].i[k] + p[3 * kLanes].i[k]);
    return r;
  }
};
static_assert(kWarps == 4, "the fixed-order sum of the warp totals is written for four warps");

}  // namespace lgr_fmt
