// lgr_kernels_finish.cuh — finish stage (cigar assembly, mm_update_extra, filters), k_finish_warp, k_assign
// Part of the single translation unit lgr_gpu.cu (included there, in order); see that file's header.
#ifndef LANCET2_B200_LGR_KERNELS_FINISH_CUH_
#define LANCET2_B200_LGR_KERNELS_FINISH_CUH_

namespace {

// ---- warp-parallel pieces of the finish phase (all 32 lanes call; results are uniform) ----
__device__ __forceinline__ int32_t warp_core_score(const DevParams& P, const ReadView& rv, int rev, const uint8_t* hap,
                                                   int c_qs, int c_rs, int len) {
  const int lane = threadIdx.x & 31;
  int32_t sc = 0;
  for (int j = lane; j < len; j += 32) {
    const int qc = rv.at(rev, c_qs + j), tc = hap[c_rs + j] & 0xf;
    sc += (qc >= 4 || tc >= 4) ? P.e : (qc == tc ? P.a : -P.b);
  }
  return __reduce_add_sync(0xffffffffu, sc);
}

// mm_update_extra: the running score s = max(s + m, 0) with its maximum is a (max,+) recurrence;
// a chunk of 32 columns is folded with an ordered tree reduction of (A,B,C,D):
//   s_out = max(s + A, B), best = max(s + C, D).  All quantities are integers (upstream keeps them
// in doubles that only ever hold integers, dp_max = (int)(max + .499)).
__device__ __noinline__ void warp_update_extra(const DevParams& P, const ReadView& rv, int rev, const uint8_t* hap, int qb,
                                               int tb, const uint32_t* c, int n, RegFinal* out) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  constexpr int NEG = -(1 << 28);
  int32_t toff = 0, qoff = 0, blen = 0, mlen = 0, n_ambi_tot = 0, s = 0, mx = 0;
  for (int k = 0; k < n; ++k) {
    const uint32_t op = c[k] & 0xf;
    const int len = (int)(c[k] >> 4);
    if (op == 0) {
      for (int base = 0; base < len; base += 32) {
        const int l = base + lane;
        const bool valid = l < len;
        int m = 0;
        bool ambi = false, diff = false;
        if (valid) {
          const int cq = rv.at(rev, qb + qoff + l), ct = hap[tb + toff + l] & 0xf;
          ambi = ct > 3 || cq > 3;
          diff = !ambi && ct != cq;
          m = sub_score(P, ct, cq);
        }
        const int na = __popc(__ballot_sync(full, ambi)), nd = __popc(__ballot_sync(full, diff));
        const int cnt = len - base < 32 ? len - base : 32;
        blen += cnt - na, mlen += cnt - (na + nd), n_ambi_tot += na;
        if (__ballot_sync(full, valid && m < 0) == 0) {
          s += __reduce_add_sync(full, m);
          if (s > mx) mx = s;
        } else {
          int A = valid ? m : 0, B = valid ? 0 : NEG, C = valid ? m : NEG, Dd = valid ? 0 : NEG;
          for (int o = 1; o < 32; o <<= 1) {
            const int Ay = __shfl_down_sync(full, A, o), By = __shfl_down_sync(full, B, o);
            const int Cy = __shfl_down_sync(full, C, o), Dy = __shfl_down_sync(full, Dd, o);
            if (lane + o < 32) {
              int d2 = B + Cy;
              if (Dd > d2) d2 = Dd;
              if (Dy > d2) d2 = Dy;
              const int c2 = A + Cy > C ? A + Cy : C;
              const int b2 = B + Ay > By ? B + Ay : By;
              A = A + Ay, B = b2, C = c2, Dd = d2;
              if (B < NEG) B = NEG;
              if (C < NEG) C = NEG;
              if (Dd < NEG) Dd = NEG;
            }
          }
          A = __shfl_sync(full, A, 0), B = __shfl_sync(full, B, 0), C = __shfl_sync(full, C, 0), Dd = __shfl_sync(full, Dd, 0);
          int best = s + C > Dd ? s + C : Dd;
          if (best > mx) mx = best;
          s = s + A > B ? s + A : B;
        }
      }
      toff += len, qoff += len;
    } else if (op == 1 || op == 2) {
      int na = 0;
      for (int base = 0; base < len; base += 32) {
        const int l = base + lane;
        bool ambi = false;
        if (l < len) ambi = op == 1 ? rv.at(rev, qb + qoff + l) > 3 : (hap[tb + toff + l] & 0xf) > 3;
        na += __popc(__ballot_sync(full, ambi));
      }
      blen += len - na, n_ambi_tot += na;
      s -= P.q + P.e;
      if (s < 0) s = 0;
      if (op == 1) qoff += len;
      else toff += len;
    } else if (op == 3) {
      toff += len;
    }
  }
  out->blen = blen, out->mlen = mlen, out->n_ambi = n_ambi_tot, out->dp_max = mx;
}

__device__ __forceinline__ int32_t warp_edit_distance(const uint8_t* read_codes, int qlen, const uint8_t* hap, int rs, int re,
                                                      int qs, const uint32_t* c, int n) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  int32_t nm = 0;
  int qpos = qs, tpos = 0;
  const int tn = re - rs;
  for (int k = 0; k < n; ++k) {
    const uint32_t op = c[k] & 0xf;
    const int len = (int)(c[k] >> 4);
    if (op == 0) {
      for (int base = 0; base < len; base += 32) {
        const int l = base + lane;
        bool mis = false;
        if (l < len) {
          const int qp = qpos + l, tp = tpos + l;
          mis = qp < qlen && tp < tn && (read_codes[qp] >> 4) != (hap[rs + tp] >> 4);
        }
        nm += __popc(__ballot_sync(full, mis));
      }
      qpos += len, tpos += len;
    } else if (op == 1) {
      nm += len, qpos += len;
    } else if (op == 2) {
      nm += len, tpos += len;
    } else if (op == 3) {
      tpos += len;
    }
  }
  return nm;
}

// One pass over a gap-free forward-strand alignment (cigar = one M op, the normal case):
// mm_update_extra's mlen / blen / n_ambi / dp_max, the ungapped core score and NM together,
// from the same two code bytes per column.
__device__ __forceinline__ void warp_finish_pure_m(const DevParams& P, const uint8_t* read_codes, const uint8_t* hap, int qb, int tb,
                                                int len, int c_qs, int c_qe, RegFinal* out, int32_t* core_out, int32_t* nm_out) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  constexpr int NEG = -(1 << 28);
  int32_t blen = 0, mlen = 0, n_ambi_tot = 0, s = 0, mx = 0, core = 0, nm = 0;
  for (int base = 0; base < len; base += 32) {
    const int l = base + lane;
    const bool valid = l < len;
    int m = 0, cm = 0;
    bool ambi = false, diff = false, mis = false;
    if (valid) {
      const int qb_ = read_codes[qb + l], tb_ = hap[tb + l];
      const int cq = qb_ & 0xf, ct = tb_ & 0xf;
      ambi = ct > 3 || cq > 3;
      diff = !ambi && ct != cq;
      m = ambi ? -P.sc_ambi : (diff ? -P.b : P.a);
      mis = (qb_ >> 4) != (tb_ >> 4);
      if (qb + l >= c_qs && qb + l < c_qe) cm = ambi ? P.e : (diff ? -P.b : P.a);
    }
    const int na = __popc(__ballot_sync(full, ambi)), nd = __popc(__ballot_sync(full, diff));
    nm += __popc(__ballot_sync(full, mis));
    core += __reduce_add_sync(full, cm);
    const int cnt = len - base < 32 ? len - base : 32;
    blen += cnt - na, mlen += cnt - (na + nd), n_ambi_tot += na;
    if (__ballot_sync(full, valid && m < 0) == 0) {
      s += __reduce_add_sync(full, m);
      if (s > mx) mx = s;
    } else {
      int A = valid ? m : 0, B = valid ? 0 : NEG, C = valid ? m : NEG, Dd = valid ? 0 : NEG;
      for (int o = 1; o < 32; o <<= 1) {
        const int Ay = __shfl_down_sync(full, A, o), By = __shfl_down_sync(full, B, o);
        const int Cy = __shfl_down_sync(full, C, o), Dy = __shfl_down_sync(full, Dd, o);
        if (lane + o < 32) {
          int d2 = B + Cy;
          if (Dd > d2) d2 = Dd;
          if (Dy > d2) d2 = Dy;
          const int c2 = A + Cy > C ? A + Cy : C;
          const int b2 = B + Ay > By ? B + Ay : By;
          A = A + Ay, B = b2, C = c2, Dd = d2;
          if (B < NEG) B = NEG;
          if (C < NEG) C = NEG;
          if (Dd < NEG) Dd = NEG;
        }
      }
      A = __shfl_sync(full, A, 0), B = __shfl_sync(full, B, 0), C = __shfl_sync(full, C, 0), Dd = __shfl_sync(full, Dd, 0);
      const int best = s + C > Dd ? s + C : Dd;
      if (best > mx) mx = best;
      s = s + A > B ? s + A : B;
    }
  }
  out->blen = blen, out->mlen = mlen, out->n_ambi = n_ambi_tot, out->dp_max = mx;
  *core_out = core, *nm_out = nm;
}

// finish_pair (lgr_core.cuh) with the per-base loops spread over the warp.  Uniform control flow;
// cigar assembly / mm_fix_cigar stay scalar on lane 0.  Returns the op count of the winning cigar
// (in fs.best), or -1 on scratch overflow; *out is valid on every lane.
struct TrackBlock {  // surviving regs of one pair (mm_set_parent / mm_select_sub inputs), one per warp in shared memory
  uint64_t key[kTrack];
  int32_t qs[kTrack], qe[kTrack], rs[kTrack], re[kTrack], score[kTrack];
};
#ifndef LGR_FIN_CHUNK
#define LGR_FIN_CHUNK 8
#endif
constexpr int kFinChunk = LGR_FIN_CHUNK;  // pairs a warp takes from the finish queue per atomic
constexpr int kFinSmemCig = 64;  // cigar ops of a reg kept in shared memory; longer ones use the HBM scratch

__device__ __forceinline__ int finish_pair_warp(const Dev& D, const ReadView& rv, const uint8_t* hap, const RegRec* regs, int n_regs,
                                             FinishScratch& fs, TrackBlock* trk, RegRec* s_reg, RegFinal* s_bf, AlnOut* out) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const DevParams& P = D.P;
  const int qlen = rv.qlen;
  int best = -1, n_surv = 0;
  int32_t best_nm = -1;
  uint64_t best_key = 0;
  // the best reg so far and the final record live in the warp's shared-memory slots (s_bf, out): as locals they
  // were spilled to local memory, whose write-through traffic to L2 bounded this kernel
  int32_t *s_qs = trk->qs, *s_qe = trk->qe, *s_rs = trk->rs, *s_re = trk->re, *s_score = trk->score;
  uint64_t* s_key = trk->key;
  for (int r = 0; r < n_regs; ++r) {
    // the reg comes out of HBM once, in one coalesced read, and is read from shared memory from here on
    __syncwarp();
    {
      constexpr int kWords = (int)(sizeof(RegRec) / 4);
      const uint32_t* src = reinterpret_cast<const uint32_t*>(&regs[r]);
      uint32_t* dst = reinterpret_cast<uint32_t*>(s_reg);
      for (int w = lane; w < kWords; w += 32) dst[w] = src[w];
    }
    __syncwarp();
    const RegRec& rg = *s_reg;
    RegAsm ra;
    const ExtRec& L = rg.ext[0];
    const ExtRec& R = rg.ext[1];
    const bool pure = rg.rev == 0 && (L.m <= 0 || (L.n_cig == 1 && (L.inl[0] & 0xf) == 0)) &&
                      (R.m <= 0 || (R.n_cig == 1 && (R.inl[0] & 0xf) == 0));
    if (pure) {
      // both tails are pure match runs (or absent) on the forward strand: mm_append_cigar x3 merges them
      // with the core into ONE M op and mm_fix_cigar has nothing to do (assemble_fix_reg with n == 1) —
      // plain arithmetic on every lane instead of the scalar assembly on lane 0 and eight broadcasts
      int32_t len = rg.c_qe - rg.c_qs;
      ra.dp_ext = 0, ra.rs = rg.c_rs, ra.qs = rg.c_qs, ra.re = rg.c_re, ra.qe = rg.c_qe;
      if (L.m > 0) len += (int32_t)(L.inl[0] >> 4), ra.dp_ext += L.max, ra.rs = rg.c_rs - (L.mqe_t + 1), ra.qs = 0;
      if (R.m > 0) len += (int32_t)(R.inl[0] >> 4), ra.dp_ext += R.max, ra.re = rg.c_re + (R.mqe_t + 1), ra.qe = qlen;
      ra.n = 1, ra.qb = ra.qs, ra.tb = ra.rs;
      if (lane == 0) fs.cig[0] = (uint32_t)len << 4;
    } else {
      int okf = 1;
      if (lane == 0) okf = assemble_fix_reg(rv, hap, rg, D.ext_arena, fs.cig, fs.cap, &ra) ? 1 : 0;
      okf = __shfl_sync(full, okf, 0);
      if (!okf) return -1;
      ra.n = __shfl_sync(full, ra.n, 0), ra.rs = __shfl_sync(full, ra.rs, 0), ra.re = __shfl_sync(full, ra.re, 0);
      ra.qs = __shfl_sync(full, ra.qs, 0), ra.qe = __shfl_sync(full, ra.qe, 0), ra.qb = __shfl_sync(full, ra.qb, 0);
      ra.tb = __shfl_sync(full, ra.tb, 0), ra.dp_ext = __shfl_sync(full, ra.dp_ext, 0);
    }
    __syncwarp();
    const int rev = rg.rev, c_qs = rg.c_qs, c_qe = rg.c_qe, c_rs = rg.c_rs;
    const int32_t score = rg.score, cnt = rg.cnt;
    const uint32_t hash = rg.hash;
    RegFinal rf;
    int32_t nm_reg = -1;
    if (ra.n == 1 && rev == 0 && (fs.cig[0] & 0xf) == 0) {
      int32_t core = 0;
      warp_finish_pure_m(P, rv.codes, hap, ra.qb, ra.tb, (int)(fs.cig[0] >> 4), c_qs, c_qe, &rf, &core, &nm_reg);
      rf.dp_score = ra.dp_ext + core;
    } else {
      RegFinal far;  // the out-of-line call's result lives in local memory; `rf` itself stays in registers
      warp_update_extra(P, rv, rev, hap, ra.qb, ra.tb, fs.cig, ra.n, &far);
      rf.dp_max = far.dp_max, rf.mlen = far.mlen, rf.blen = far.blen, rf.n_ambi = far.n_ambi;
      rf.dp_score = ra.dp_ext + warp_core_score(P, rv, rev, hap, c_qs, c_rs, c_qe - c_qs);
    }
    rf.rs = ra.rs, rf.re = ra.re, rf.qs = ra.qs, rf.qe = ra.qe, rf.n_cig = ra.n;
    bool flt = false;
    if (cnt < P.min_cnt) flt = true;
    if (rf.mlen < P.min_sc) flt = true;
    else if (rf.dp_max < P.min_dp_max) flt = true;
    else if ((float)rf.qs > (float)qlen * P.max_clip_ratio && (float)(qlen - rf.qe) > (float)qlen * P.max_clip_ratio) flt = true;
    if (flt) continue;
    const uint64_t key = (uint64_t)(uint32_t)rf.dp_max << 32 | hash;
    if (n_surv < kTrack && lane == 0) {
      s_qs[n_surv] = rf.qs, s_qe[n_surv] = rf.qe, s_rs[n_surv] = rf.rs, s_re[n_surv] = rf.re;
      s_score[n_surv] = score, s_key[n_surv] = key;
    }
    ++n_surv;
    if (best < 0 || key >= best_key) {
      best = r, best_key = key, best_nm = nm_reg;
      if (lane == 0) *s_bf = rf;
      uint32_t* tmp = fs.best;
      fs.best = fs.cig;
      fs.cig = tmp;
    }
  }
  if (best < 0) {
    if (lane == 0) *out = AlnOut{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, -1, 0};
    __syncwarp();
    return 0;
  }
  __syncwarp();
  const int n_ret = n_surv > 1 ? select_returned(P, n_surv, s_qs, s_qe, s_rs, s_re, s_score, s_key) : n_surv;
  const RegFinal bf = *s_bf;
  const int32_t nm = best_nm >= 0 ? best_nm : warp_edit_distance(rv.codes, qlen, hap, bf.rs, bf.re, bf.qs, fs.best, bf.n_cig);
  if (lane == 0)
    *out = AlnOut{1, regs[best].score, bf.rs, bf.re, bf.qs, bf.qe, regs[best].rev, bf.dp_score, bf.dp_max, bf.mlen, bf.blen,
                  bf.n_ambi, nm, bf.n_cig, -1, n_ret};
  __syncwarp();
  return bf.n_cig;
}

// Phase B2 kernel: one warp per parked pair, every extension already done: the warp-parallel
// finish (assemble, fix, extra, filter, sort) and the final record.
__global__ void __launch_bounds__(128, LGR_FIN_MINB) k_finish_warp(const __grid_constant__ Dev D) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  __shared__ TrackBlock s_trk[4];
  __shared__ uint32_t s_cig[4][2 * kFinSmemCig];
  __shared__ RegRec s_regs[4];
  __shared__ RegFinal s_bf[4];
  __shared__ __align__(16) AlnOut s_out[4];
  uint32_t* fin0 = D.fin_scratch + (size_t)gwarp * 2 * D.fin_cap;
  TrackBlock* trk = &s_trk[threadIdx.x >> 5];
  uint32_t* scig = s_cig[threadIdx.x >> 5];
  long long n_aligned = 0;
  // pairs are taken kFinChunk at a time: the work per pair is uniform enough here, and one queue
  // atomic per pair (hundreds of thousands on one address) is a measurable share of the kernel
  long long chunk_next = 0, chunk_end = 0;
  for (;;) {
    if (chunk_next >= chunk_end) {
      if (lane == 0) chunk_next = atomicAdd((unsigned long long*)&D.ctr[C_FINPOS], (unsigned long long)kFinChunk);
      chunk_next = __shfl_sync(full, chunk_next, 0);
      if (chunk_next >= D.n_pairs) break;
      chunk_end = chunk_next + kFinChunk < D.n_pairs ? chunk_next + kFinChunk : D.n_pairs;
    }
    const long long pair = chunk_next++;
    const PairReg d = D.pair_reg[pair];
    if (d.n <= 0) continue;
    const int read = d.read;
    const uint8_t* hapc = D.hap_codes + D.hap_off[d.hap];
    const int64_t roff = D.read_off[read];
    ReadView rv{D.read_codes + roff, (int)(D.read_off[read + 1] - roff)};
    RegRec* regs = D.regs + d.first;
    // cigars live in shared memory; the rare reg with more ops than fit reruns on the HBM scratch
    // (one inlined body, run a second time on the HBM scratch in the rare overflow case: the record and the
    // per-reg structs then live in registers — as an out-of-line call they went through local memory, and the
    // spills around the call were 2.7 GB of L1→L2 write traffic per cfg2 step)
    FinishScratch fs{scig, scig + kFinSmemCig, kFinSmemCig};
    AlnOut* ao = &s_out[threadIdx.x >> 5];
    int nc = -1;
    for (int attempt = 0; attempt < 2 && nc < 0; ++attempt) {
      if (attempt == 1) {
        __syncwarp();
        fs = FinishScratch{fin0, fin0 + D.fin_cap, D.fin_cap};
      }
      nc = finish_pair_warp(D, rv, hapc, regs, d.n, fs, trk, &s_regs[threadIdx.x >> 5], &s_bf[threadIdx.x >> 5], ao);
    }
    if (lane == 0) {
      if (nc < 0) {
        flag_err(D, D.read_grp[read], E_CIG_SCRATCH);
        write_invalid(&D.aln[pair]);
      } else {
        store_final(D, D.read_grp[read], pair, *ao, fs.best, nc);
        n_aligned += ao->valid;
      }
    }
    __syncwarp();
  }
  if (lane == 0) atomicAdd((unsigned long long*)&D.ctr[C_ALIGNED], (unsigned long long)n_aligned);
}

// one lane per (read, variant): AssignReadToAlleles' inner loops (genotyper.cpp:294-318)
__global__ void __launch_bounds__(128) k_assign(const __grid_constant__ Dev D) {
  const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= D.n_assign) return;
  // read r with asg_off[r] <= slot < asg_off[r+1]
  int lo = 0, hi = D.n_reads;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (D.asg_off[mid] <= slot) lo = mid;
    else hi = mid;
  }
  const int r = lo;  // asg_off[r] <= slot < asg_off[r+1]
  const int v_local = (int)(slot - D.asg_off[r]);
  const int g = D.read_grp[r];
  const int h0 = D.grp_hap_begin[g], Pn = D.grp_hap_begin[g + 1] - h0;
  const int v = D.grp_var_begin[g] + v_local;
  const int64_t roff = D.read_off[r];
  const int qlen = (int)(D.read_off[r + 1] - roff);
  const int64_t pair0 = D.pair_off[r];
  AssignOut best;
  best.local_score = best.local_identity = best.folded_read_pos = 0.0;
  best.global_score = 0, best.ref_nm = best.own_hap_nm = best.hap_id = 0, best.allele = 0, best.base_qual = 0, best.assigned = 0;
  for (int i = 0; i < 5; ++i) best.pad[i] = 0;
  double best_cs = 0.0;
  uint32_t ref_nm = (uint32_t)qlen;
  {
    const AlnOut& a0 = D.aln[pair0];
    if (Pn > 0 && a0.valid && a0.rs < a0.re) ref_nm = (uint32_t)a0.nm;
  }
  for (int h = 0; h < Pn; ++h) {
    const AlnOut a = D.aln[pair0 + h];
    if (!a.valid) continue;
    const int64_t vh = D.var_hap_off[v] + h;
    const int allele = D.var_allele[vh];
    if (allele < 0) continue;
    const int32_t vs = D.var_start[vh], vl = D.var_len[vh];
    if (!(vs + vl > a.rs && vs < a.re)) continue;
    const uint32_t* cig = a.cigar_off < 0 ? D.cigar_inline + (pair0 + h) * LGR_CIGAR_INLINE : D.cigar_arena + a.cigar_off;
    AssignOut cand;
    score_read_variant(a, cig, D.read_codes + roff, D.read_quals + roff, qlen, D.hap_codes + D.hap_off[h0 + h], vs, vl, allele, h,
                       ref_nm, c_phred_err, &cand);
    const double cs = (double)cand.global_score + cand.local_score * cand.local_identity;
    if (best.assigned && cs <= best_cs) continue;
    best = cand, best_cs = cs;
  }
  D.assign[slot] = best;
}

}  // namespace

#endif  // LANCET2_B200_LGR_KERNELS_FINISH_CUH_
