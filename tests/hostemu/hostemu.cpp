// TEST INFRASTRUCTURE — compiles the kernels' per-lane scalar core
// (lancet2_b200/csrc/lgr_core.cuh) with g++ and drives it pair by pair on the CPU,
// so that the control flow the GPU lanes execute can be diffed against the oracle
// without a GPU.  It is never linked into or loaded by the product library.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/lancet_gpu_realign.h"
#define LGR_CORE_SELFCHECK 1  // closed forms of the warp kernels verified against the scalar paths as they run
#include "../../lancet2_b200/csrc/lgr_core.cuh"

static const double kPhredErr[256] = {
#include "../../lancet2_b200/csrc/phred_lut.inc"
};

static lgr::DevParams MakeDev(const lgr_params& p) {
  lgr::DevParams d{};
  d.k = p.k, d.w = p.w, d.a = p.a, d.b = p.b, d.q = p.q, d.e = p.e, d.sc_ambi = p.sc_ambi, d.bw = p.bw;
  d.end_bonus = p.end_bonus, d.max_gap = p.max_gap, d.max_gap_ref = p.max_gap_ref, d.max_skip = p.max_chain_skip;
  d.max_iter = p.max_chain_iter, d.min_cnt = p.min_cnt, d.min_sc = p.min_chain_score, d.min_dp_max = p.min_dp_max;
  d.max_max_occ = p.max_max_occ, d.occ_dist = p.occ_dist, d.best_n = p.best_n, d.seed = p.seed, d.mask_len = p.mask_len;
  d.pen_gap = (float)(p.chain_gap_scale * 0.01 * p.k);
  d.pen_skip = (float)(p.chain_skip_scale * 0.01 * p.k);
  d.mask_level = p.mask_level, d.pri_ratio = p.pri_ratio, d.max_clip_ratio = p.max_clip_ratio, d.q_occ_frac = p.q_occ_frac;
  d.min_strand_sc = (int32_t)(p.max_gap * 0.8);
  return d;
}

extern "C" int emu_genotype_batch(const lgr_params* prm, const lgr_batch_in* in, lgr_batch_out* out, lgr_stats* stats,
                                  int small_only_inline) {
  using namespace lgr;
  const DevParams P = MakeDev(*prm);
  const int cap = 65535;  // the 16-bit positions of Ws / RadixScratch
  std::vector<int32_t> wsbuf((size_t)A_COUNT * cap);
  Ws<1> ws{wsbuf.data(), cap};
  RadixScratch rsx;
  ChainCounters ctr{};
  std::vector<uint8_t> hapc(in->hap_off[in->n_haps]), readc(in->read_off[in->n_reads]);
  for (int64_t i = 0; i < (int64_t)hapc.size(); ++i) hapc[i] = encode_base(in->hap_bases[i]);
  for (int64_t i = 0; i < (int64_t)readc.size(); ++i) readc[i] = encode_base(in->read_bases[i]);
  // index
  std::vector<std::vector<uint64_t>> idx(in->n_haps);
  std::vector<int32_t> hmid(in->n_haps);
  for (int h = 0; h < in->n_haps; ++h) {
    const int len = (int)(in->hap_off[h + 1] - in->hap_off[h]);
    std::vector<uint64_t> x(len + 1);
    std::vector<uint32_t> y(len + 1);
    const int n = len > 0 ? sketch_any(hapc.data() + in->hap_off[h], len, P.w, P.k, x.data(), y.data(), len + 1) : 0;
    idx[h].resize(n);
    for (int i = 0; i < n; ++i) idx[h][i] = (x[i] >> 8) << kIdxShift | y[i];
    std::sort(idx[h].begin(), idx[h].end());
    hmid[h] = hap_mid_occ(idx[h].data(), n, prm->mid_occ_frac, prm->min_mid_occ, prm->max_mid_occ);
  }
  int64_t po = 0, ao = 0, arena_used = 0;
  std::vector<uint32_t> ext_arena(1 << 22);
  int64_t ext_used = 0;
  auto alloc = [&](int n) -> int64_t { const int64_t o = ext_used; ext_used += n; return ext_used <= (int64_t)ext_arena.size() ? o : -1; };
  int rc = LGR_OK;
  for (int g = 0; g < in->n_groups; ++g) {
    const int h0 = in->grp_hap_begin[g], Pn = in->grp_hap_begin[g + 1] - h0;
    const int v0 = in->grp_var_begin[g], V = in->grp_var_begin[g + 1] - v0;
    int32_t mid_occ = prm->mid_occ;
    if (in->grp_mid_occ && in->grp_mid_occ[g] > 0) mid_occ = in->grp_mid_occ[g];
    if (mid_occ <= 0) mid_occ = Pn > 0 ? hmid[h0] : prm->min_mid_occ;
    for (int r = in->grp_read_begin[g]; r < in->grp_read_begin[g + 1]; ++r) {
      const int qlen = (int)(in->read_off[r + 1] - in->read_off[r]);
      const uint8_t* rc_ = readc.data() + in->read_off[r];
      const uint8_t* rq = in->read_quals + in->read_off[r];
      std::vector<uint64_t> mx(qlen + 1);
      std::vector<uint32_t> my(qlen + 1);
      int mz_n = qlen > 0 ? sketch_any(rc_, qlen, P.w, P.k, mx.data(), my.data(), qlen + 1) : 0;
      if (P.q_occ_frac > 0.0f) mz_n = seed_mz_flt(mx.data(), my.data(), mz_n, mid_occ, P.q_occ_frac);
      ReadView rv{rc_, qlen};
      std::vector<AlnOut> alns(Pn);
      std::vector<std::vector<uint32_t>> cigs(Pn);
      for (int h = 0; h < Pn; ++h) {
        const int hl = (int)(in->hap_off[h0 + h + 1] - in->hap_off[h0 + h]);
        const uint8_t* hc = hapc.data() + in->hap_off[h0 + h];
        PairIn pin{rv, hc, hl, idx[h0 + h].data(), (int)idx[h0 + h].size(), mx.data(), my.data(), mz_n,
                   in->read_name_hash[r], mid_occ};
        int n_regs = 0;
        AlnOut ao_{};
        ao_.cigar_off = -1;
        const int st = qlen > 0 ? map_chain_phase<1>(P, pin, ws, &rsx, &n_regs, &ctr) : kMapNoHit;
        if (st == kMapOverflow) rc = LGR_E_LIMIT;
        if (st == kMapOk) {
          std::vector<RegRec> regs(n_regs);
          for (int i = 0; i < n_regs; ++i) {
            export_reg<1>(ws, i, qlen, &regs[i]);
            for (int side = 0; side < 2; ++side) {
              ExtRec& E = regs[i].ext[side];
              if (E.m <= 0) continue;
              const int T = prune_cols(P, E.m, E.n);
              std::vector<uint8_t> dir((size_t)E.m * T);
              std::vector<int32_t> hcol(std::max(E.m, T) + 1), ecol(std::max(E.m, T) + 1);
              std::vector<uint32_t> tmp(2 * E.m + 4);
              if (!run_ext_scalar(P, rv, hc, &regs[i], side, dir.data(), hcol.data(), ecol.data(), tmp.data(),
                                  (int)tmp.size(), ext_arena.data(), alloc, &ctr))
                rc = LGR_E_LIMIT;
            }
          }
          std::vector<uint32_t> c1(3 * qlen + 16), c2(3 * qlen + 16);
          FinishScratch fs{c1.data(), c2.data(), (int)c1.size()};
          const int nc = finish_pair(P, rv, hc, regs.data(), n_regs, ext_arena.data(), fs, &ao_);
          if (nc < 0) rc = LGR_E_LIMIT;
          else cigs[h].assign(fs.best, fs.best + nc);
        }
        alns[h] = ao_;
        lgr_aln* o = &out->aln[po + h];
        std::memcpy(o, &ao_, sizeof(lgr_aln));
        if (ao_.valid) {
          if (ao_.n_cigar <= LGR_CIGAR_INLINE) {
            std::memcpy(out->cigar_inline + (po + h) * LGR_CIGAR_INLINE, cigs[h].data(), 4 * cigs[h].size());
          } else {
            o->cigar_off = (int32_t)arena_used;
            std::memcpy(out->cigar_arena + arena_used, cigs[h].data(), 4 * cigs[h].size());
            arena_used += ao_.n_cigar;
          }
        }
      }
      // assign
      uint32_t ref_nm = (uint32_t)qlen;
      if (Pn > 0 && alns[0].valid && alns[0].rs < alns[0].re) ref_nm = (uint32_t)alns[0].nm;
      for (int v = 0; v < V; ++v) {
        lgr_assign* dst = &out->assign[ao + v];
        std::memset(dst, 0, sizeof(*dst));
        bool have = false;
        double best = 0;
        for (int h = 0; h < Pn; ++h) {
          if (!alns[h].valid) continue;
          const int64_t vh = in->var_hap_off[v0 + v] + h;
          const int allele = in->var_allele[vh];
          if (allele < 0) continue;
          const int32_t vs = in->var_start[vh], vl = in->var_len[vh];
          if (!(vs + vl > alns[h].rs && vs < alns[h].re)) continue;
          AssignOut cand;
          score_read_variant(alns[h], cigs[h].data(), rc_, rq, qlen, hapc.data() + in->hap_off[h0 + h], vs, vl, allele, h,
                             ref_nm, kPhredErr, &cand);
          const double cs = (double)cand.global_score + cand.local_score * cand.local_identity;
          if (have && cs <= best) continue;
          have = true, best = cs;
          std::memcpy(dst, &cand, sizeof(*dst));
        }
      }
      po += Pn, ao += V;
    }
  }
  out->cigar_arena_used = arena_used;
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    stats->n_pairs = po, stats->dp_cells = ctr.dp_cells, stats->dp_cells_full = ctr.dp_cells_full;
    stats->chain_evals = ctr.chain_evals, stats->n_anchors = ctr.n_anchors;
    for (int64_t i = 0; i < po; ++i) stats->n_aligned += out->aln[i].valid;
  }
  return rc;
}

static_assert(sizeof(lgr::AlnOut) == sizeof(lgr_aln), "AlnOut must mirror lgr_aln");
static_assert(sizeof(lgr::AssignOut) == sizeof(lgr_assign), "AssignOut must mirror lgr_assign");

// closed-form self checks accumulated by the runs so far: out[0] = failures, out[1] = co-linear
// chains checked, out[2] = closed-form extensions checked, out[3] = closed-form chain tails checked,
// out[4] = skipped radix passes (> 64 strictly sorted anchors) checked
extern "C" void emu_selfcheck(long long* out) {
  out[0] = lgr::lgr_selfcheck_failures, out[1] = lgr::lgr_selfcheck_colinear_seen, out[2] = lgr::lgr_selfcheck_ext_seen;
  out[3] = lgr::lgr_selfcheck_tail_seen, out[4] = lgr::lgr_selfcheck_sorted_seen;
}
