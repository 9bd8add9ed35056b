// TEST INFRASTRUCTURE — CPU oracle.  Not part of the product path: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library, and only as the checker / the timed CPU baseline.
//
// Restates, batch-in → batch-out with the SAME structs as the GPU C-ABI
// (include/lancet_gpu_realign.h), what the reference does per Genotype() call:
//   Genotyper::ResetData / AlignToAllHaplotypes / AssignReadToAlleles
//     (reference: src/lancet/caller/genotyper.cpp:243-267, 376-411, 269-321)
//   ScoreReadAtVariant / ComputeHaplotypeEditDistance (combined_scorer.cpp:60-108, 24-38)
//   ComputeLocalScore / ComputeSoftClipPenalty        (local_scorer.cpp:166-279, 290-305)
//   hts::ComputeEditDistance / CigarRefPosToQueryPos  (hts/cigar_utils.h:48-94, 104-139)
// The minimap2 half is oracle/mm2_restate.cpp (PARITY UNPINNED, see its header);
// the Lancet-owned half is pinned against the reference's own sources compiled
// unmodified into oracle/_ref (oracle/Makefile; golden vectors tests/golden/, tests/test_oracle_scoring.py) and
// against the 11 known-answer cases of tests/hts/cigar_utils_test.cpp:58-172.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/lancet_gpu_realign.h"
#include "mm2_restate.hpp"

namespace {

const double kPhredErr[256] = {
#include "../lancet2_b200/csrc/phred_lut.inc"
};

// scoring_constants.h:35-41 — 5x5 matrix, target row × query column, N scores 0
inline int LocalMat(uint8_t t, uint8_t q) {
  if (t == 4 || q == 4) return 0;
  return t == q ? 1 : -4;
}

// scoring_constants.h:48-74 — ENCODE_TABLE (note: unlike minimap2, U/u → 4)
inline uint8_t LancetEncode(uint8_t c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
  }
}

struct CigOp {
  char op;  // 'M','I','D','N','S','H','P','=','X'
  uint32_t len;
};

struct LocalScore {
  double pbq = 0.0, raw = 0.0, identity = 0.0;
  uint8_t base_qual = 0;
};

// local_scorer.cpp:166-279 (+ RegionAccumulator :32-113)
LocalScore ComputeLocalScore(const std::vector<CigOp>& cigar, const uint8_t* qry, size_t qry_n,
                             const uint8_t* tgt, size_t tgt_n, const uint8_t* quals, size_t qual_n,
                             int32_t aln_start, int32_t var_start, int32_t var_len) {
  LocalScore out;
  if (cigar.empty() || var_len == 0) return out;
  const int32_t var_end = var_start + var_len;
  double pbq = 0.0, raw = 0.0;
  size_t matches = 0, aligned = 0;
  uint8_t min_bq = 255;
  int32_t tpos = 0;
  size_t qpos = 0;
  auto in_region = [&](int32_t t) {
    const int32_t abs_pos = aln_start + t;
    return abs_pos >= var_start && abs_pos < var_end;
  };
  auto track_bq = [&](size_t qp) {
    if (qp < qual_n) min_bq = std::min(min_bq, quals[qp]);
  };
  for (const CigOp& u : cigar) {
    const bool consumes_ref = u.op == 'M' || u.op == 'D' || u.op == 'N' || u.op == '=' || u.op == 'X';
    if (aln_start + tpos >= var_end && consumes_ref) break;
    switch (u.op) {
      case 'M': case '=': case 'X':
        for (uint32_t i = 0; i < u.len; ++i, ++tpos, ++qpos) {
          if (!in_region(tpos)) continue;
          ++aligned;
          if (!(qpos >= qry_n || (size_t)tpos >= tgt_n)) {
            const int r = LocalMat(tgt[tpos], qry[qpos]);
            raw += (double)r;
            const double weight = qpos < qual_n ? 1.0 - kPhredErr[quals[qpos]] : 1.0;
            pbq += (double)r * weight;
            matches += qry[qpos] == tgt[tpos] ? 1 : 0;
          }
          track_bq(qpos);
        }
        break;
      case 'I': {
        const bool in = in_region(tpos);
        for (uint32_t i = 0; i < u.len; ++i, ++qpos) {
          if (!in) continue;
          ++aligned;
          track_bq(qpos);
          pbq += 3.0;  // +SCORING_GAP_EXTEND, sign as in the reference (local_scorer.cpp:235)
        }
        break;
      }
      case 'D':
        for (uint32_t i = 0; i < u.len; ++i, ++tpos) {
          if (in_region(tpos)) {
            ++aligned;
            pbq += 3.0;  // local_scorer.cpp:253
          }
        }
        if (qpos > 0 && qpos - 1 < qual_n) min_bq = std::min(min_bq, quals[qpos - 1]);
        if (qpos < qual_n) min_bq = std::min(min_bq, quals[qpos]);
        break;
      case 'S':
        qpos += u.len;
        break;
      case 'N':
        tpos += (int32_t)u.len;
        break;
      default:
        break;
    }
  }
  out.pbq = pbq;
  out.raw = raw;
  out.identity = aligned > 0 ? (double)matches / (double)aligned : 0.0;
  out.base_qual = min_bq == 255 ? 0 : min_bq;
  return out;
}

// local_scorer.cpp:290-305
double SoftClipPenalty(const std::vector<CigOp>& cigar) {
  if (cigar.empty()) return 0.0;
  const int32_t c5 = cigar.front().op == 'S' ? (int32_t)cigar.front().len : 0;
  const int32_t c3 = cigar.size() > 1 && cigar.back().op == 'S' ? (int32_t)cigar.back().len : 0;
  return (double)(c5 + c3) * 4;
}

// hts/cigar_utils.h:48-94
uint32_t EditDistance(const std::vector<CigOp>& cigar, const uint8_t* qry, size_t qry_n,
                      const uint8_t* tgt, size_t tgt_n) {
  uint32_t nm = 0;
  size_t qpos = 0, tpos = 0;
  for (const CigOp& u : cigar) {
    switch (u.op) {
      case 'M':
        for (uint32_t i = 0; i < u.len; ++i, ++qpos, ++tpos)
          if (qpos < qry_n && tpos < tgt_n && qry[qpos] != tgt[tpos]) ++nm;
        break;
      case '=': qpos += u.len, tpos += u.len; break;
      case 'X': nm += u.len, qpos += u.len, tpos += u.len; break;
      case 'I': nm += u.len, qpos += u.len; break;
      case 'D': nm += u.len, tpos += u.len; break;
      case 'S': qpos += u.len; break;
      case 'N': tpos += u.len; break;
      default: break;
    }
  }
  return nm;
}

// hts/cigar_utils.h:104-139
size_t RefPosToQueryPos(const std::vector<CigOp>& cigar, size_t ref_pos) {
  size_t qpos = 0, tpos = 0;
  for (const CigOp& u : cigar) {
    switch (u.op) {
      case 'M': case '=': case 'X':
        for (uint32_t i = 0; i < u.len; ++i, ++qpos, ++tpos)
          if (tpos == ref_pos) return qpos;
        break;
      case 'I': case 'S': qpos += u.len; break;
      case 'D': case 'N':
        for (uint32_t i = 0; i < u.len; ++i, ++tpos)
          if (tpos == ref_pos) return qpos;
        break;
      default: break;
    }
  }
  return qpos;
}

const char kBamOps[] = "MIDNSHP=XB";

// genotyper.cpp:45-69 BuildCigar
std::vector<CigOp> BuildCigar(const mm2r::Reg& hit, int qlen) {
  std::vector<CigOp> c;
  if (!hit.has_p || hit.cigar.empty()) return c;
  if (hit.qs > 0) c.push_back({'S', (uint32_t)hit.qs});
  for (uint32_t v : hit.cigar) c.push_back({kBamOps[v & 0xf], v >> 4});
  if (hit.qe < qlen) c.push_back({'S', (uint32_t)(qlen - hit.qe)});
  return c;
}

struct Aln {  // Mm2AlnResult (genotyper.h:34-43)
  std::vector<CigOp> cigar;
  int32_t score, rs, re;
  int hap;
};

struct GroupView {
  int g;
  int hap0, P, read0, R, var0, V;
};

void GenotypeGroup(const lgr_params& prm, const lgr_batch_in& in, const GroupView& gv,
                   const int64_t* pair_off, const int64_t* asg_off, lgr_batch_out* out,
                   std::atomic<int64_t>* arena_used, int64_t* dp_cells_full, int64_t* chain_evals,
                   int64_t* n_anchors, int read_begin, int read_end,
                   const std::vector<mm2r::HapIndex>& idx,
                   const std::vector<std::vector<uint8_t>>& hap_enc, int32_t mid_occ) {
  for (int rl = read_begin; rl < read_end; ++rl) {
    const int r = gv.read0 + rl;
    const uint8_t* rseq = in.read_bases + in.read_off[r];
    const uint8_t* rqual = in.read_quals + in.read_off[r];
    const int qlen = (int)(in.read_off[r + 1] - in.read_off[r]);
    std::vector<Aln> alns;
    // AlignToAllHaplotypes (genotyper.cpp:376-411)
    for (int h = 0; h < gv.P; ++h) {
      mm2r::MapDebug dbg;
      std::vector<mm2r::Reg> regs = mm2r::Map(idx[h], rseq, qlen, in.read_name_hash[r], prm, mid_occ, &dbg);
      *dp_cells_full += dbg.dp_cells_full;
      *chain_evals += dbg.chain_evals;
      *n_anchors += (int64_t)dbg.anchors.size();
      lgr_aln& o = out->aln[pair_off[r] + h];
      std::memset(&o, 0, sizeof(o));
      o.cigar_off = -1;
      if (regs.empty()) continue;
      const mm2r::Reg& top = regs[0];
      o.valid = 1;
      o.score = top.score, o.rs = top.rs, o.re = top.re, o.qs = top.qs, o.qe = top.qe;
      o.rev = top.rev, o.dp_score = top.dp_score, o.dp_max = top.dp_max;
      o.mlen = top.mlen, o.blen = top.blen, o.n_ambi = top.n_ambi;
      o.n_cigar = (int32_t)top.cigar.size();
      o.n_regs = (int32_t)regs.size();
      uint32_t* dst;
      if (o.n_cigar <= LGR_CIGAR_INLINE) {
        dst = out->cigar_inline + (pair_off[r] + h) * LGR_CIGAR_INLINE;
      } else {
        const int64_t off = arena_used->fetch_add(o.n_cigar);
        if (off + o.n_cigar > out->cigar_arena_cap) { o.n_cigar = 0; dst = nullptr; o.cigar_off = -2; }
        else { o.cigar_off = (int32_t)off; dst = out->cigar_arena + off; }
      }
      if (dst) std::memcpy(dst, top.cigar.data(), sizeof(uint32_t) * top.cigar.size());
      Aln a;
      a.cigar = BuildCigar(top, qlen);
      a.score = top.score, a.rs = top.rs, a.re = top.re, a.hap = h;
      alns.push_back(std::move(a));
    }
    // AssignReadToAlleles (genotyper.cpp:269-321)
    for (int v = 0; v < gv.V; ++v) {
      lgr_assign& s = out->assign[asg_off[r] + v];
      std::memset(&s, 0, sizeof(s));
    }
    if (alns.empty()) continue;
    std::vector<uint8_t> qenc(qlen);
    for (int i = 0; i < qlen; ++i) qenc[i] = LancetEncode(rseq[i]);
    // ComputeHaplotypeEditDistance against REF hap (combined_scorer.cpp:24-38)
    uint32_t ref_nm = (uint32_t)qlen;
    for (const Aln& a : alns) {
      if (a.hap != 0 || a.rs >= a.re) continue;
      ref_nm = EditDistance(a.cigar, qenc.data(), qenc.size(), hap_enc[0].data() + a.rs,
                            (size_t)(a.re - a.rs));
      break;
    }
    for (const Aln& a : alns) {
      const std::vector<uint8_t>& henc = hap_enc[a.hap];
      const uint8_t* tgt = henc.data() + a.rs;
      const size_t tgt_n = (size_t)(a.re - a.rs);
      const uint32_t own_nm = EditDistance(a.cigar, qenc.data(), qenc.size(), tgt, tgt_n);
      out->aln[pair_off[r] + a.hap].nm = (int32_t)own_nm;
      for (int v = 0; v < gv.V; ++v) {
        const int64_t vh = in.var_hap_off[gv.var0 + v] + a.hap;
        const int allele = in.var_allele[vh];
        if (allele < 0) continue;  // ExtractHapBounds → nullopt
        const int32_t vstart = in.var_start[vh], vlen = in.var_len[vh];
        if (!(vstart + vlen > a.rs && vstart < a.re)) continue;  // OverlapsAlignment
        // ScoreReadAtVariant (combined_scorer.cpp:60-108)
        const LocalScore loc = ComputeLocalScore(a.cigar, qenc.data(), qenc.size(), tgt, tgt_n, rqual,
                                                 (size_t)qlen, a.rs, vstart, vlen);
        const double global_adjusted = (double)a.score - SoftClipPenalty(a.cigar);
        lgr_assign cand;
        std::memset(&cand, 0, sizeof(cand));
        cand.allele = (int8_t)allele;
        cand.global_score = (int32_t)(global_adjusted - loc.raw);
        cand.local_score = loc.pbq;
        cand.local_identity = loc.identity;
        cand.base_qual = loc.base_qual;
        cand.hap_id = (uint32_t)a.hap;
        cand.own_hap_nm = own_nm;
        size_t var_start_in_aln = 0;
        if (vstart > a.rs) var_start_in_aln = (size_t)(vstart - a.rs);
        const size_t qpos_at_var = RefPosToQueryPos(a.cigar, var_start_in_aln);
        const double rel = qlen > 0 ? (double)qpos_at_var / (double)qlen : 0.5;
        cand.folded_read_pos = std::min(rel, 1.0 - rel);
        cand.ref_nm = ref_nm;
        cand.assigned = 1;
        lgr_assign& cur = out->assign[asg_off[r] + v];
        const double cs_new = (double)cand.global_score + cand.local_score * cand.local_identity;
        const double cs_cur = (double)cur.global_score + cur.local_score * cur.local_identity;
        if (cur.assigned && cs_new <= cs_cur) continue;  // ties keep the earlier haplotype
        cur = cand;
      }
    }
  }
}

}  // namespace

extern "C" {

// Same contract as lgr_genotype_batch, on the CPU, with `n_threads` host threads
// (reads of a group are split across threads; results are order-independent).
int orc_genotype_batch(const lgr_params* prm, const lgr_batch_in* in, lgr_batch_out* out,
                       int n_threads, lgr_stats* stats) {
  if (!prm || !in || !out) return LGR_E_ARG;
  std::vector<int64_t> pair_off(in->n_reads + 1), asg_off(in->n_reads + 1);
  {
    int64_t po = 0, ao = 0;
    for (int g = 0; g < in->n_groups; ++g) {
      const int P = in->grp_hap_begin[g + 1] - in->grp_hap_begin[g];
      const int V = in->grp_var_begin[g + 1] - in->grp_var_begin[g];
      for (int r = in->grp_read_begin[g]; r < in->grp_read_begin[g + 1]; ++r) {
        pair_off[r] = po, asg_off[r] = ao;
        po += P, ao += V;
      }
    }
    pair_off[in->n_reads] = po, asg_off[in->n_reads] = ao;
    if (po > out->n_pairs || ao > out->n_assign) return LGR_E_ARG;
  }
  std::atomic<int64_t> arena_used{0};
  if (n_threads < 1) n_threads = 1;
  std::vector<int64_t> cells(n_threads, 0), evals(n_threads, 0), anchors(n_threads, 0);
  std::atomic<int> next_group{0};
  auto worker = [&](int tid) {
    for (;;) {
      const int g = next_group.fetch_add(1);
      if (g >= in->n_groups) break;
      GroupView gv;
      gv.g = g;
      gv.hap0 = in->grp_hap_begin[g], gv.P = in->grp_hap_begin[g + 1] - gv.hap0;
      gv.read0 = in->grp_read_begin[g], gv.R = in->grp_read_begin[g + 1] - gv.read0;
      gv.var0 = in->grp_var_begin[g], gv.V = in->grp_var_begin[g + 1] - gv.var0;
      // ResetData (genotyper.cpp:243-267)
      std::vector<mm2r::HapIndex> idx(gv.P);
      std::vector<std::vector<uint8_t>> henc(gv.P);
      for (int h = 0; h < gv.P; ++h) {
        const uint8_t* hs = in->hap_bases + in->hap_off[gv.hap0 + h];
        const int hl = (int)(in->hap_off[gv.hap0 + h + 1] - in->hap_off[gv.hap0 + h]);
        mm2r::BuildHapIndex(hs, hl, prm->w, prm->k, idx[h]);
        henc[h].resize(hl);
        for (int i = 0; i < hl; ++i) henc[h][i] = LancetEncode(hs[i]);
      }
      int32_t mid_occ = prm->mid_occ;
      if (in->grp_mid_occ && in->grp_mid_occ[g] > 0) mid_occ = in->grp_mid_occ[g];
      if (mid_occ <= 0) mid_occ = gv.P > 0 ? mm2r::MidOccFromIndex(idx[0], *prm) : prm->min_mid_occ;
      GenotypeGroup(*prm, *in, gv, pair_off.data(), asg_off.data(), out, &arena_used, &cells[tid],
                    &evals[tid], &anchors[tid], 0, gv.R, idx, henc, mid_occ);
    }
  };
  if (n_threads == 1) {
    worker(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t) th.emplace_back(worker, t);
    for (auto& t : th) t.join();
  }
  out->cigar_arena_used = arena_used.load();
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    stats->n_pairs = pair_off[in->n_reads];
    for (int t = 0; t < n_threads; ++t)
      stats->dp_cells_full += cells[t], stats->chain_evals += evals[t], stats->n_anchors += anchors[t];
    for (int64_t i = 0; i < stats->n_pairs; ++i) stats->n_aligned += out->aln[i].valid;
  }
  if (out->cigar_arena_used > out->cigar_arena_cap) return LGR_E_CIGAR_OVERFLOW;
  return LGR_OK;
}

// ---- fine-grained entry points for stage-level parity tests -----------------

int orc_sketch(const uint8_t* seq, int len, int w, int k, uint64_t* x, uint64_t* y, int cap) {
  std::vector<mm2r::U128> v;
  if (len > 0) mm2r::Sketch(seq, len, w, k, 0, v);
  const int n = (int)v.size();
  for (int i = 0; i < n && i < cap; ++i) x[i] = v[i].x, y[i] = v[i].y;
  return n;
}

int orc_hap_mid_occ(const lgr_params* prm, const uint8_t* hap, int len) {
  mm2r::HapIndex idx;
  mm2r::BuildHapIndex(hap, len, prm->w, prm->k, idx);
  return mm2r::MidOccFromIndex(idx, *prm);
}

// One mm_map with every intermediate exposed.  Arrays may be NULL.
int orc_map_debug(const lgr_params* prm, const uint8_t* hap, int hlen, const uint8_t* read, int qlen,
                  uint32_t qname_hash, int32_t mid_occ, lgr_aln* aln, uint32_t* cigar, int cigar_cap,
                  uint64_t* anchor_x, uint64_t* anchor_y, int32_t* f, int32_t* p, int anchor_cap,
                  int32_t* n_anchor, uint64_t* u, int u_cap, int32_t* n_u) {
  mm2r::HapIndex idx;
  mm2r::BuildHapIndex(hap, hlen, prm->w, prm->k, idx);
  if (mid_occ <= 0) mid_occ = mm2r::MidOccFromIndex(idx, *prm);
  mm2r::MapDebug dbg;
  std::vector<mm2r::Reg> regs = mm2r::Map(idx, read, qlen, qname_hash, *prm, mid_occ, &dbg);
  if (n_anchor) *n_anchor = (int32_t)dbg.anchors.size();
  for (int i = 0; i < (int)dbg.anchors.size() && i < anchor_cap; ++i) {
    if (anchor_x) anchor_x[i] = dbg.anchors[i].x;
    if (anchor_y) anchor_y[i] = dbg.anchors[i].y;
    if (f) f[i] = dbg.f[i];
    if (p) p[i] = (int32_t)dbg.p[i];
  }
  if (n_u) *n_u = (int32_t)dbg.u.size();
  for (int i = 0; i < (int)dbg.u.size() && i < u_cap; ++i)
    if (u) u[i] = dbg.u[i];
  if (aln) {
    std::memset(aln, 0, sizeof(*aln));
    aln->cigar_off = -1;
    if (!regs.empty()) {
      const mm2r::Reg& t = regs[0];
      aln->valid = 1, aln->score = t.score, aln->rs = t.rs, aln->re = t.re, aln->qs = t.qs, aln->qe = t.qe;
      aln->rev = t.rev, aln->dp_score = t.dp_score, aln->dp_max = t.dp_max, aln->mlen = t.mlen;
      aln->blen = t.blen, aln->n_ambi = t.n_ambi, aln->n_cigar = (int32_t)t.cigar.size();
      aln->n_regs = (int32_t)regs.size();
      for (int i = 0; i < aln->n_cigar && i < cigar_cap; ++i)
        if (cigar) cigar[i] = t.cigar[i];
    }
  }
  return (int)regs.size();
}

// ksw2 extension in isolation (for DP-kernel parity): returns n_cigar
int orc_extz(const uint8_t* q, int qlen, const uint8_t* t, int tlen, int a, int b, int sc_ambi, int gapo,
             int gape, int end_bonus, int right, int rev_cigar, int32_t* out5, uint32_t* cigar,
             int cigar_cap) {
  int8_t mat[25];
  const int8_t aa = (int8_t)(a < 0 ? -a : a), bb = (int8_t)(b > 0 ? -b : b),
               amb = (int8_t)(sc_ambi > 0 ? -sc_ambi : sc_ambi);
  for (int i = 0; i < 4; ++i) {
    for (int j = 0; j < 4; ++j) mat[i * 5 + j] = i == j ? aa : bb;
    mat[i * 5 + 4] = amb;
  }
  for (int j = 0; j < 5; ++j) mat[20 + j] = amb;
  mm2r::ExtzResult ez;
  mm2r::ExtzOnly(qlen, q, tlen, t, mat, gapo, gape, end_bonus,
                 (right ? mm2r::kEzRight : 0) | (rev_cigar ? mm2r::kEzRevCigar : 0), ez);
  out5[0] = ez.max, out5[1] = ez.mqe, out5[2] = ez.mqe_t, out5[3] = ez.reach_end, out5[4] = ez.max_t;
  const int n = (int)ez.cigar.size();
  for (int i = 0; i < n && i < cigar_cap; ++i) cigar[i] = ez.cigar[i];
  return n;
}

// Lancet-owned scoring pieces in isolation (pinned against oracle/_ref).
// cigar as BAM u32 (len<<4|op) INCLUDING S ops.
static std::vector<CigOp> FromBam(const uint32_t* c, int n) {
  std::vector<CigOp> v;
  for (int i = 0; i < n; ++i) v.push_back({kBamOps[c[i] & 0xf], c[i] >> 4});
  return v;
}
uint32_t orc_edit_distance(const uint32_t* cigar, int n, const uint8_t* q, int qn, const uint8_t* t, int tn) {
  return EditDistance(FromBam(cigar, n), q, (size_t)qn, t, (size_t)tn);
}
uint64_t orc_refpos_to_qpos(const uint32_t* cigar, int n, uint64_t ref_pos) {
  return RefPosToQueryPos(FromBam(cigar, n), (size_t)ref_pos);
}
double orc_softclip_penalty(const uint32_t* cigar, int n) { return SoftClipPenalty(FromBam(cigar, n)); }
void orc_local_score(const uint32_t* cigar, int n, const uint8_t* q, int qn, const uint8_t* t, int tn,
                     const uint8_t* quals, int qualn, int32_t aln_start, int32_t var_start, int32_t var_len,
                     double* out3, uint8_t* bq) {
  const LocalScore s = ComputeLocalScore(FromBam(cigar, n), q, (size_t)qn, t, (size_t)tn, quals,
                                         (size_t)qualn, aln_start, var_start, var_len);
  out3[0] = s.pbq, out3[1] = s.raw, out3[2] = s.identity;
  *bq = s.base_qual;
}
double orc_phred_err(uint32_t q) { return kPhredErr[q > 255 ? 255 : q]; }
uint8_t orc_lancet_encode(uint8_t c) { return LancetEncode(c); }
uint32_t orc_x31_hash(const char* s) { return mm2r::X31HashString(s); }

void orc_default_params(lgr_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->k = 11, p->w = 5;
  p->a = 1, p->b = 4, p->q = 12, p->e = 3, p->sc_ambi = 1;
  p->bw = 10000, p->zdrop = 100000, p->end_bonus = 10000;
  p->max_gap = 200, p->max_gap_ref = 5000;
  p->max_chain_skip = 25, p->max_chain_iter = 5000, p->min_cnt = 3, p->min_chain_score = 40;
  p->min_dp_max = 80;
  p->mid_occ = 0, p->min_mid_occ = 10, p->max_mid_occ = 1000000, p->max_max_occ = 4095;
  p->occ_dist = 500, p->best_n = 1, p->seed = 11;
  p->mid_occ_frac = 2e-4f, p->q_occ_frac = 0.01f, p->chain_gap_scale = 0.8f, p->chain_skip_scale = 0.0f;
  p->mask_level = 0.5f, p->pri_ratio = 0.8f, p->max_clip_ratio = 1.0f;
  p->mask_len = 0x7fffffff;
  p->cigar_arena_ops = 1 << 20;
}

}  // extern "C"
