// lgr_kernels_ext.cuh — ksw2-style extension: anti-diagonal wavefront DP, warp traceback, k_ext_warp
// Part of the single translation unit lgr_gpu.cu (included there, in order); see that file's header.
#ifndef LANCET2_B200_LGR_KERNELS_EXT_CUH_
#define LANCET2_B200_LGR_KERNELS_EXT_CUH_

namespace {

// ---------------------------------------------------------------------------------------
// ext_dp_warp: one warp computes one extension tail.  Lane l owns query row j = 32*blk + l and
// sweeps the target columns; on step s it computes cell (i = s - l, j).  H and the F flowing
// down a column travel to the lane below with two shuffles per step; E stays in the lane.
// Rows beyond 32 are processed in further passes with the boundary row (H, F) kept in scratch.
// Direction bytes are stored diagonal-major ([blk][s][lane]) so that every step is one
// coalesced 32-byte store.  Same recurrences, tie rules and column pruning as ext_dp_scalar.
// All 32 lanes must call it; results are written by lane 0 into reg->ext[side].
// ---------------------------------------------------------------------------------------
constexpr int kDynPruneMinRows = 12;  // shorter tails: the static bound is already small, skip the extra pass
constexpr int32_t kBandLow = -30000;   // stand-in for H/E/F of cells outside the band: below every true value (checked per task)
constexpr int kBandLowPacked = (int)(((uint32_t)kBandLow & 0xffffu) | ((uint32_t)kBandLow << 16));

__device__ __noinline__ void ext_dp_warp(const Dev& D, int grp, RegRec* reg, int side, const ReadView& rv, const uint8_t* hapc,
                                         uint8_t* dir_g, uint8_t* dir_s, int dir_s_cap, int32_t* Hb, int32_t* Fb, uint32_t* wcig,
                                         long long* cells, long long* cells_full) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const DevParams& P = D.P;
  const int q = P.q, e = P.e;
  const int32_t sc_match = P.a, sc_mis = -P.b, sc_amb = -P.sc_ambi;  // in registers: the loop's generic stores could alias P
  const int m = reg->ext[side].m, n = reg->ext[side].n;
  int T = prune_cols(P, m, n);
  int band_ins = 1 << 20, band_del = 1 << 20;  // |i - j| kept on the insertion / deletion side (no band by default)
  const bool right = side == 0;
  ExtQuery qf{rv, reg->rev, side, reg->c_qs, reg->c_qe};
  ExtTarget tf{hapc, side, reg->c_rs, reg->c_re};
  // Data-dependent column bound (exact).  Any path ending in the last query row at target column
  // i = m-1+d (d > 0) deletes at least d target bases: its score is <= a*m - q - e*d, and every
  // cell at or beyond that column is bounded the same way.  If LB is the score of SOME path that
  // ends in the last row at an earlier column, columns with a*m - q - e*d <= LB can hold neither
  // the first maximum of the last row nor the global maximum, and the traceback never enters
  // them (cells only depend on smaller columns).  prune_cols uses the worst case LB = -b*m;
  // here every lane scores one concrete family of paths — the main diagonal for p bases, one
  // gap of delta = lane-15 (deletion > 0, insertion < 0), then the shifted diagonal — and the
  // warp keeps the best, which for a tail that crosses an indel is close to the optimum.
  if (n >= m && P.e > 0 && m >= kDynPruneMinRows && 2 * m + 32 <= dir_s_cap) {
    // stage the m query codes and the first m+16 target codes in the warp's shared-memory slice
    // (free until the direction bytes are written) so that the 32 lanes read bytes, not functors
    uint8_t* sq = dir_s;
    uint8_t* st = dir_s + m;
    const int nt = m + 16 < n ? m + 16 : n;
    for (int x = lane; x < m; x += 32) sq[x] = (uint8_t)qf(x);
    for (int x = lane; x < nt; x += 32) st[x] = (uint8_t)tf(x);
    __syncwarp();
    const int delta = lane - 15;
    int32_t lb = kNegInf;
    if (delta >= 0 ? m + delta <= n : -delta < m) {
      const int k = delta < 0 ? -delta : 0;  // inserted query bases
      const uint8_t* t0 = st;
      const uint8_t* t1 = st + (delta > 0 ? delta : 0);
      const uint8_t* q1 = sq + k;
      int32_t p0 = 0, ps = 0, best = 0;      // P0[p], shifted prefix, max(P0 - shifted)
      const int steps = m - k;
      for (int p = 0; p < steps; ++p) {
        const int tcp = t0[p], tcs = t1[p], qcp = sq[p], qcs = q1[p];
        p0 += (tcp > 3 || qcp > 3) ? sc_amb : (tcp == qcp ? sc_match : sc_mis);
        ps += (tcs > 3 || qcs > 3) ? sc_amb : (tcs == qcs ? sc_match : sc_mis);
        const int32_t dlt = p0 - ps;
        if (dlt > best) best = dlt;
      }
      lb = best + ps - (delta != 0 ? q + e * (delta < 0 ? -delta : delta) : 0);
    }
    lb = __reduce_max_sync(full, lb);
    const int X = P.a * m - q - lb;
    const int Dd = X <= 0 ? 0 : X / e;
    if (m + Dd < T) T = m + Dd;
    // The same LB bounds the band.  A path through a cell with i - j = d > 0 has deleted >= d target bases
    // (score <= a*m - q - e*d), one through a cell with j - i = d > 0 has inserted >= d query bases and can
    // match at most m - d of them (score <= a*(m-d) - q - e*d), whether it ends in the last row or right
    // there.  Cells whose bound is < LB can hold neither the last-row maximum, nor the global maximum, nor a
    // traceback cell, and any cell they feed either loses to an in-band input or is itself unreachable by the
    // traceback (a tie would be a path through the cell scoring >= LB).  So they may be skipped, or computed
    // from inputs that are merely <= the true values (kBandLow), without changing any result.
    // kBandLow must undercut every true H/E/F and survive ~100 steps of drift inside int16.
    const int64_t lowest = -(3 * (int64_t)q + (int64_t)e * (T + m + 4) + (int64_t)(P.b > P.sc_ambi ? P.b : P.sc_ambi) * m);
    if (lowest >= kBandLow + 64 && e <= 16) {
      band_del = Dd;
      band_ins = X <= 0 ? 0 : X / (P.a + e);
    }
    __syncwarp();  // the staging bytes are dead from here on; the slice becomes direction storage
  }
  const int nblk = (m + 31) >> 5;
  // direction bytes: block b holds rows [32b, 32b+rows_b) as [step][row]; shared memory when the
  // whole matrix fits the warp's slice, else the HBM scratch
  const int rows_last = m - (nblk - 1) * 32;
  const int dir_bytes = (nblk - 1) * 32 * (T + 31) + rows_last * (T + rows_last - 1);
  uint8_t* dir = dir_bytes <= dir_s_cap ? dir_s : dir_g;
  int32_t ezmax = 0, mqe = kNegInf, mqe_t = -1;
  for (int blk = 0; blk < nblk; ++blk) {
    const int j = blk * 32 + lane;
    const int rows = m - blk * 32 < 32 ? m - blk * 32 : 32;
    const bool row_ok = lane < rows;
    const int qc = row_ok ? qf(j) : 4;
    int32_t e_cur = -(q + e * (j + 1)) - q - e;  // E(0, j)
    int32_t diag = j == 0 ? 0 : -(q + e * j);    // H(-1, j-1)
    int32_t hf = 0;                               // packed (H low16, F-out high16) of my last cell
    uint8_t* dblk = dir + (size_t)blk * 32 * (T + 31);
    const int nsteps = T + rows - 1;
    const bool save_bnd = blk + 1 < nblk;
    // Band (see above): the block only runs the anti-diagonals [s_lo, s_hi) that hold in-band cells of its
    // rows.  Everything on anti-diagonal s_lo - 1 and before is out of band (i - j <= -band_ins - 2 there), so
    // a lane that starts mid-row takes kBandLow for the three inputs it never saw; a lane that starts at
    // column 0 keeps the true boundary values.  Past the band's right edge the boundary row of the previous
    // block holds kBandLow (filled below).
    const int s_lo = blk * 32 - band_ins - 1 > 0 ? blk * 32 - band_ins - 1 : 0;
    const int s_hi = band_del < nsteps && blk * 32 + 63 + band_del < nsteps ? blk * 32 + 63 + band_del : nsteps;
    if (s_lo > 0) {
      hf = kBandLowPacked;
      if (s_lo - lane > 0) e_cur = kBandLow, diag = kBandLow;
    }
    // Per 32 steps every lane fetches one target base (and one packed boundary cell for row
    // blocks > 0): coalesced, off the per-step dependency chain, and the step body stays
    // branch-free — lane l takes its base t[s-l] with one indexed shuffle out of the current or
    // previous 32-base register window, lane 0 takes its boundary input with a broadcast.
    const int s_first = s_lo & ~31;
    int tprev = 4, tcur = 4, bcur = 0;
    {
      const int ti = s_first - 32 + lane;  // the window before the first one (lanes look back up to 31 columns)
      tcur = ti >= 0 && ti < T ? tf(ti) : 4;
    }
    for (int s0 = s_first; s0 < s_hi; s0 += 32) {
      const int ti = s0 + lane;
      tprev = tcur;
      tcur = ti < T ? tf(ti) : 4;
      if (blk > 0) bcur = ti < T ? Hb[ti] : 0;
      const int kmax = s_hi - s0 < 32 ? s_hi - s0 : 32;
      for (int k = s0 < s_lo ? s_lo - s0 : 0; k < kmax; ++k) {
        const int s = s0 + k;
        const int i = s - lane;
        const int tc = __shfl_sync(full, k >= lane ? tcur : tprev, (k - lane) & 31);
        int up_hf = __shfl_up_sync(full, hf, 1);
        int feed;
        if (blk == 0) {
          const int32_t h0 = -(q + e * (s + 1));
          feed = (int)(((uint32_t)h0 & 0xffffu) | ((uint32_t)(h0 - q - e) << 16));
        } else {
          feed = __shfl_sync(full, bcur, k);
        }
        if (lane == 0) up_hf = feed;
        const int32_t up_h = (int32_t)(int16_t)(up_hf & 0xffff);
        const int32_t up_f = up_hf >> 16;
        if (row_ok && i >= 0 && i < T) {
          const int32_t sc = (tc > 3 || qc > 3) ? sc_amb : (tc == qc ? sc_match : sc_mis);
          uint8_t d;
          int32_t en, fn;
          const int32_t h = ext_cell(diag + sc, e_cur, up_f, q, e, right, &d, &en, &fn);
          dblk[s * rows + lane] = d;
          diag = up_h;
          e_cur = en;
          hf = (int)(((uint32_t)h & 0xffffu) | ((uint32_t)fn << 16));
          if (h > ezmax) ezmax = h;
          if (j == m - 1 && h > mqe) mqe = h, mqe_t = i;
          if (save_bnd && lane == 31) Hb[i] = hf;
        }
      }
    }
    if (save_bnd && s_hi < nsteps) {
      // columns of the boundary row this block did not reach but the next one will ask for: all beyond the
      // band's right edge of row 32*blk + 31
      const int from = s_hi - 31 > 0 ? s_hi - 31 : 0;
      const int to = s_hi + 32 < T ? s_hi + 32 : T;
      for (int x = from + lane; x < to; x += 32) Hb[x] = kBandLowPacked;
    }
    __syncwarp();
  }
  ezmax = __reduce_max_sync(full, ezmax);
  mqe_t = __shfl_sync(full, mqe_t, (m - 1) & 31);
  __syncwarp();
  // ksw_backtrack, warp-cooperative: the path mostly runs down the diagonal, so the 32 lanes
  // fetch the direction bytes of the next 32 diagonal cells in one go and the (warp-uniform)
  // state machine walks them by shuffle; a gap step leaves the diagonal and refetches.  One
  // memory round trip per <= 32 steps instead of one per step (the bytes of a long tail sit in L2).
  // Runs of equal ops are counted in registers and pushed once (same result as ksw_push_cigar).
  CigBuf cb{wcig, 0, D.wcig_cap};
  {
    auto dirf = [&](int i, int j) -> uint32_t {
      const int b = j >> 5, l = j & 31;
      const int rows = m - b * 32 < 32 ? m - b * 32 : 32;
      return dir[(size_t)b * 32 * (T + 31) + (size_t)(i + l) * rows + l];
    };
    int i = mqe_t, j = m - 1, state = 0;
    uint32_t run_op = 0;
    int run_len = 0;
    auto emit = [&](uint32_t op) {
      if (run_len > 0 && op == run_op) {
        ++run_len;
      } else {
        if (run_len > 0 && lane == 0) cb.push(run_op, run_len);
        run_op = op, run_len = 1;
      }
    };
    while (i >= 0 && j >= 0) {
      const int wi = i - lane, wj = j - lane;
      const uint32_t dv = (wi >= 0 && wj >= 0) ? dirf(wi, wj) : 0u;
      for (int k = 0; k < 32; ++k) {
        const uint32_t tmp = __shfl_sync(full, dv, k);
        if (state == 0) state = tmp & 7;
        else if (!(tmp >> (state + 2) & 1)) state = 0;
        if (state == 0) state = tmp & 7;
        if (state == 0) {
          emit(0), --i, --j;
          if (i < 0 || j < 0) break;
        } else {
          if (state == 1) emit(2), --i;
          else emit(1), --j;
          break;  // off this diagonal
        }
      }
    }
    if (lane == 0) {
      if (run_len > 0) cb.push(run_op, run_len);
      if (i >= 0) cb.push(2, i + 1);
      if (j >= 0) cb.push(1, j + 1);
      if (side != 0 && cb.n <= cb.cap) {  // right extension: ksw2 reverses the backtrack order
        for (int a = 0; a < cb.n >> 1; ++a) {
          const uint32_t t = cb.ops[a];
          cb.ops[a] = cb.ops[cb.n - 1 - a];
          cb.ops[cb.n - 1 - a] = t;
        }
      }
    }
  }
  if (lane == 0) {
    ExtRec& E = reg->ext[side];
    E.max = ezmax;
    E.mqe_t = mqe_t;
    E.n_cig = cb.n;
    if (cb.n > D.wcig_cap) {
      flag_err(D, grp, E_CIG_SCRATCH);
      E.n_cig = 0;
    } else if (cb.n <= kInlineCig) {
      E.cig_off = -1;
      for (int c = 0; c < cb.n; ++c) E.inl[c] = wcig[c];
    } else {
      const long long o = atomicAdd((unsigned long long*)&D.ctr[C_EXTARENA], (unsigned long long)cb.n);
      if (o + cb.n > D.ext_arena_cap) {
        flag_err(D, grp, E_EXT_ARENA);
        E.n_cig = 0;
      } else {
        E.cig_off = (int32_t)o;
        for (int c = 0; c < cb.n; ++c) D.ext_arena[o + c] = wcig[c];
      }
    }
    *cells += (long long)m * T;
    *cells_full += (long long)m * n;
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------
// ext_dp_multi<W>: 32 / W extension tails on one warp, W lanes each (W = 8: m <= 8, W = 16: m <= 16).
// Most tails that reach the wavefront are a dozen rows (a mismatch or indel near a read end): one
// tail per warp left half the lanes idle on every DP instruction (round-1 ncu: 14-15 of 32 active).
// Same recurrence, tie rules, column bounds and backtrack as ext_dp_warp, with
//   * one row block (m <= W), so no boundary row;
//   * the tail's query and target codes staged in the segment's shared-memory slice (T is a few
//     dozen columns at these sizes), so a step needs ONE shuffle (H|F from the row above, width W);
//   * direction bytes [step][row] in the same slice (ext_class guarantees they fit);
//   * the segment-wide backtrack: W diagonal cells per fetch.
// Lane sl == 0 of a segment writes its tail's ExtRec.  Control flow is warp-uniform: loop bounds
// are the maximum over the segments, the per-segment work is predicated.
// ---------------------------------------------------------------------------------------
template <int W>
__device__ __noinline__ void ext_dp_multi(const Dev& D, const TaskRec* tasks, int n_here, uint8_t* dir_warp, uint32_t* cig_warp,
                                          long long* cells, long long* cells_full) {
  constexpr int NSEG = 32 / W;
  constexpr int kSegBytes = kDirSmemPerWarp / NSEG;
  constexpr int kSegCig = 2 * W + 8;  // ops of one tail's cigar (<= 2m + 2)
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, seg = lane / W, sl = lane % W;
  const DevParams& P = D.P;
  const int q = P.q, e = P.e;
  const int32_t sc_match = P.a, sc_mis = -P.b, sc_amb = -P.sc_ambi;
  const bool seg_ok = seg < n_here;
  TaskRec tk = TaskRec{0, 0, 0, 0};
  if (seg_ok) tk = tasks[seg];
  RegRec* reg = &D.regs[tk.reg];
  const int side = tk.side;
  const bool right = side == 0;
  const uint8_t* hapc = D.hap_codes + D.hap_off[tk.hap];
  const int64_t roff = D.read_off[tk.read];
  ReadView rv{D.read_codes + roff, (int)(D.read_off[tk.read + 1] - roff)};
  const int m = seg_ok ? reg->ext[side].m : 0, n = seg_ok ? reg->ext[side].n : 0;
  int T = seg_ok ? prune_cols(P, m, n) : 0;
  ExtQuery qf{rv, seg_ok ? reg->rev : 0, side, seg_ok ? reg->c_qs : 0, seg_ok ? reg->c_qe : 0};
  ExtTarget tf{hapc, side, seg_ok ? reg->c_rs : 0, seg_ok ? reg->c_re : 0};
  uint8_t* sq = dir_warp + (size_t)seg * kSegBytes;  // [W] query codes
  uint8_t* st = sq + W;                              // [kSegHead - W] target codes of the first T columns
  uint8_t* dir = sq + kSegHead;                      // [T + m - 1][m] direction bytes
  uint32_t* cig = cig_warp + (size_t)seg * kSegCig;
  // stage the codes (T is the static bound here; the data-dependent bound below can only shrink it)
  if (sl < m) sq[sl] = (uint8_t)qf(sl);
  for (int x = sl; x < T; x += W) st[x] = (uint8_t)tf(x);
  __syncwarp();
  // ---- data-dependent column bound, as in ext_dp_warp: 32 gap families delta = -15 .. 16 ----
  {
    const bool want = seg_ok && n >= m && P.e > 0 && m >= kDynPruneMinRows;
    int32_t lb = kNegInf;
    if (W < 32 && __any_sync(full, want)) {
      for (int round = 0; round < NSEG; ++round) {
        const int delta = round * W + sl - 15;
        if (want && (delta >= 0 ? m + delta <= n : -delta < m)) {
          const int k = delta < 0 ? -delta : 0;
          const int sh = delta > 0 ? delta : 0;
          int32_t p0 = 0, ps = 0, best = 0;
          const int steps = m - k;
          bool in_stage = steps + sh <= T;  // the shifted diagonal must stay inside the staged columns
          if (in_stage) {
            for (int p = 0; p < steps; ++p) {
              const int tcp = st[p], tcs = st[sh + p], qcp = sq[p], qcs = sq[k + p];
              p0 += (tcp > 3 || qcp > 3) ? sc_amb : (tcp == qcp ? sc_match : sc_mis);
              ps += (tcs > 3 || qcs > 3) ? sc_amb : (tcs == qcs ? sc_match : sc_mis);
              const int32_t dlt = p0 - ps;
              if (dlt > best) best = dlt;
            }
            const int32_t v = best + ps - (delta != 0 ? q + e * (delta < 0 ? -delta : delta) : 0);
            if (v > lb) lb = v;
          }
        }
      }
      for (int o = W / 2; o > 0; o >>= 1) {
        const int32_t v = __shfl_xor_sync(full, lb, o, W);
        if (v > lb) lb = v;
      }
      if (want && lb > kNegInf) {
        const int X = P.a * m - q - lb;
        const int Dd = X <= 0 ? 0 : X / e;
        if (m + Dd < T) T = m + Dd;
      }
    }
  }
  // ---- wavefront: lane sl owns query row sl, step s computes cell (i = s - sl, sl) ----
  const int nsteps = seg_ok ? T + m - 1 : 0;
  int nsteps_max = nsteps;
  for (int o = 16; o >= W; o >>= 1) {
    const int v = __shfl_xor_sync(full, nsteps_max, o);
    if (v > nsteps_max) nsteps_max = v;
  }
  const bool row_ok = seg_ok && sl < m;
  const int qc = row_ok ? sq[sl] : 4;
  int32_t e_cur = -(q + e * (sl + 1)) - q - e;  // E(0, j)
  int32_t diag = sl == 0 ? 0 : -(q + e * sl);   // H(-1, j-1)
  int32_t hf = 0;                                // packed (H low16, F-out high16) of my last cell
  int32_t ezmax = 0, mqe = kNegInf, mqe_t = -1;
  for (int s = 0; s < nsteps_max; ++s) {
    int up_hf = __shfl_up_sync(full, hf, 1, W);
    if (sl == 0) {
      const int32_t h0 = -(q + e * (s + 1));
      up_hf = (int)(((uint32_t)h0 & 0xffffu) | ((uint32_t)(h0 - q - e) << 16));
    }
    const int i = s - sl;
    if (row_ok && i >= 0 && i < T) {
      const int tc = st[i];
      const int32_t up_h = (int32_t)(int16_t)(up_hf & 0xffff);
      const int32_t up_f = up_hf >> 16;
      const int32_t sc = (tc > 3 || qc > 3) ? sc_amb : (tc == qc ? sc_match : sc_mis);
      uint8_t d;
      int32_t en, fn;
      const int32_t h = ext_cell(diag + sc, e_cur, up_f, q, e, right, &d, &en, &fn);
      dir[s * m + sl] = d;
      diag = up_h;
      e_cur = en;
      hf = (int)(((uint32_t)h & 0xffffu) | ((uint32_t)fn << 16));
      if (h > ezmax) ezmax = h;
      if (sl == m - 1 && h > mqe) mqe = h, mqe_t = i;
    }
  }
  for (int o = W / 2; o > 0; o >>= 1) {
    const int32_t v = __shfl_xor_sync(full, ezmax, o, W);
    if (v > ezmax) ezmax = v;
  }
  mqe_t = __shfl_sync(full, mqe_t, m > 0 ? m - 1 : 0, W);
  __syncwarp();
  // ---- ksw_backtrack, one segment each: W diagonal cells per fetch ----
  CigBuf cb{cig, 0, kSegCig};
  {
    int i = mqe_t, j = m - 1, state = 0;
    uint32_t run_op = 0;
    int run_len = 0;
    auto emit = [&](uint32_t op) {
      if (run_len > 0 && op == run_op) {
        ++run_len;
      } else {
        if (run_len > 0 && sl == 0) cb.push(run_op, run_len);
        run_op = op, run_len = 1;
      }
    };
    bool active = seg_ok && i >= 0 && j >= 0;
    while (__any_sync(full, active)) {
      const int wi = i - sl, wj = j - sl;
      const uint32_t dv = (active && wi >= 0 && wj >= 0) ? dir[(wi + wj) * m + wj] : 0u;
      bool off_diag = false;
      for (int k = 0; k < W; ++k) {
        const uint32_t tmp = __shfl_sync(full, dv, k, W);
        if (active && !off_diag) {
          if (state == 0) state = tmp & 7;
          else if (!(tmp >> (state + 2) & 1)) state = 0;
          if (state == 0) state = tmp & 7;
          if (state == 0) {
            emit(0), --i, --j;
            if (i < 0 || j < 0) active = false;
          } else {
            if (state == 1) emit(2), --i;
            else emit(1), --j;
            off_diag = true;  // left this diagonal: refetch
            if (i < 0 || j < 0) active = false;
          }
        }
      }
    }
    if (seg_ok && sl == 0) {
      if (run_len > 0) cb.push(run_op, run_len);
      if (i >= 0) cb.push(2, i + 1);
      if (j >= 0) cb.push(1, j + 1);
      if (side != 0 && cb.n <= cb.cap) {  // right extension: ksw2 reverses the backtrack order
        for (int a = 0; a < cb.n >> 1; ++a) {
          const uint32_t t = cb.ops[a];
          cb.ops[a] = cb.ops[cb.n - 1 - a];
          cb.ops[cb.n - 1 - a] = t;
        }
      }
    }
  }
  if (seg_ok && sl == 0) {
    ExtRec& E = reg->ext[side];
    E.max = ezmax;
    E.mqe_t = mqe_t;
    E.n_cig = cb.n;
    if (cb.n > kSegCig) {
      flag_err(D, D.read_grp[tk.read], E_CIG_SCRATCH);
      E.n_cig = 0;
    } else if (cb.n <= kInlineCig) {
      E.cig_off = -1;
      for (int c = 0; c < cb.n; ++c) E.inl[c] = cig[c];
    } else {
      const long long o = atomicAdd((unsigned long long*)&D.ctr[C_EXTARENA], (unsigned long long)cb.n);
      if (o + cb.n > D.ext_arena_cap) {
        flag_err(D, D.read_grp[tk.read], E_EXT_ARENA);
        E.n_cig = 0;
      } else {
        E.cig_off = (int32_t)o;
        for (int c = 0; c < cb.n; ++c) D.ext_arena[o + c] = cig[c];
      }
    }
    *cells += (long long)m * T;
    *cells_full += (long long)m * n;
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------
// ext_dp_multi_t<W>: the TRANSPOSED sub-warp wavefront for tails with few target columns (n <= W,
// n < m): lane sl owns target column i = sl and sweeps the query rows, step s computes cell
// (i = sl, j = s - sl).  A read that overhangs a haplotype end leaves ~100 query bases against the
// handful of haplotype bases outside the first anchor; with lanes on rows that is four 32-row blocks
// of a few columns each (every block pays 31 steps of pipeline fill for ~5 useful ones).  Here the
// whole tail is one pass of m + n - 1 steps with n lanes busy, and 32 / W tails share the warp.
// Roles swap with respect to ext_dp_multi: F (the gap that runs along the query) stays in the lane,
// H and E travel to the lane on the right.  Same cell function, tie rules, maxima and backtrack.
// ---------------------------------------------------------------------------------------
template <int W>
__device__ __noinline__ void ext_dp_multi_t(const Dev& D, const TaskRec* tasks, int n_here, uint8_t* dir_warp, uint32_t* cig_warp,
                                            long long* cells, long long* cells_full) {
  constexpr int NSEG = 32 / W;
  constexpr int kSegBytes = kDirSmemPerWarp / NSEG;
  constexpr int kSegCig = 2 * W + 8;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, seg = lane / W, sl = lane % W;
  const DevParams& P = D.P;
  const int q = P.q, e = P.e;
  const int32_t sc_match = P.a, sc_mis = -P.b, sc_amb = -P.sc_ambi;
  const bool seg_ok = seg < n_here;
  TaskRec tk = TaskRec{0, 0, 0, 0};
  if (seg_ok) tk = tasks[seg];
  RegRec* reg = &D.regs[tk.reg];
  const int side = tk.side;
  const bool right = side == 0;
  const uint8_t* hapc = D.hap_codes + D.hap_off[tk.hap];
  const int64_t roff = D.read_off[tk.read];
  ReadView rv{D.read_codes + roff, (int)(D.read_off[tk.read + 1] - roff)};
  const int m = seg_ok ? reg->ext[side].m : 0, n = seg_ok ? reg->ext[side].n : 0;  // n < m, n <= W: every column is computed
  ExtQuery qf{rv, seg_ok ? reg->rev : 0, side, seg_ok ? reg->c_qs : 0, seg_ok ? reg->c_qe : 0};
  ExtTarget tf{hapc, side, seg_ok ? reg->c_rs : 0, seg_ok ? reg->c_re : 0};
  uint8_t* sq = dir_warp + (size_t)seg * kSegBytes;  // [m] query codes
  uint8_t* dir = sq + m + n;                         // [m + n - 1][n] direction bytes (the target codes stay in registers)
  uint32_t* cig = cig_warp + (size_t)seg * kSegCig;
  for (int x = sl; x < m; x += W) sq[x] = (uint8_t)qf(x);
  const bool col_ok = seg_ok && sl < n;
  const int tc = col_ok ? tf(sl) : 4;
  __syncwarp();
  const int nsteps = seg_ok ? m + n - 1 : 0;
  int nsteps_max = nsteps;
  for (int o = 16; o >= W; o >>= 1) {
    const int v = __shfl_xor_sync(full, nsteps_max, o);
    if (v > nsteps_max) nsteps_max = v;
  }
  int32_t f_cur = -(q + e * (sl + 1)) - q - e;  // F(i, 0)
  int32_t diag = sl == 0 ? 0 : -(q + e * sl);   // H(i-1, -1)
  int32_t he = 0;                                // packed (H low16, E-out high16) of my last cell
  int32_t ezmax = 0, mqe = kNegInf;
  for (int s = 0; s < nsteps_max; ++s) {
    int left = __shfl_up_sync(full, he, 1, W);
    if (sl == 0) {
      const int32_t h0 = -(q + e * (s + 1));    // H(-1, j), j = s
      left = (int)(((uint32_t)h0 & 0xffffu) | ((uint32_t)(h0 - q - e) << 16));
    }
    const int j = s - sl;
    if (col_ok && j >= 0 && j < m) {
      const int qc = sq[j];
      const int32_t left_h = (int32_t)(int16_t)(left & 0xffff);
      const int32_t left_e = left >> 16;
      const int32_t sc = (tc > 3 || qc > 3) ? sc_amb : (tc == qc ? sc_match : sc_mis);
      uint8_t d;
      int32_t en, fn;
      const int32_t h = ext_cell(diag + sc, left_e, f_cur, q, e, right, &d, &en, &fn);
      dir[s * n + sl] = d;
      diag = left_h;
      f_cur = fn;
      he = (int)(((uint32_t)h & 0xffffu) | ((uint32_t)en << 16));
      if (h > ezmax) ezmax = h;
      if (j == m - 1) mqe = h;  // this column's cell of the last query row
    }
  }
  // first maximum of the last row in column order: (max score, then smallest column)
  int32_t mqe_t = col_ok ? sl : -1;
  for (int o = W / 2; o > 0; o >>= 1) {
    const int32_t v = __shfl_xor_sync(full, ezmax, o, W);
    if (v > ezmax) ezmax = v;
    const int32_t oq = __shfl_xor_sync(full, mqe, o, W), ot = __shfl_xor_sync(full, mqe_t, o, W);
    if (oq > mqe || (oq == mqe && ot >= 0 && (mqe_t < 0 || ot < mqe_t))) mqe = oq, mqe_t = ot;
  }
  __syncwarp();
  CigBuf cb{cig, 0, kSegCig};
  {
    int i = mqe_t, j = m - 1, state = 0;
    uint32_t run_op = 0;
    int run_len = 0;
    auto emit = [&](uint32_t op) {
      if (run_len > 0 && op == run_op) {
        ++run_len;
      } else {
        if (run_len > 0 && sl == 0) cb.push(run_op, run_len);
        run_op = op, run_len = 1;
      }
    };
    bool active = seg_ok && i >= 0 && j >= 0;
    while (__any_sync(full, active)) {
      const int wi = i - sl, wj = j - sl;
      const uint32_t dv = (active && wi >= 0 && wj >= 0) ? dir[(wi + wj) * n + wi] : 0u;
      bool off_diag = false;
      for (int k = 0; k < W; ++k) {
        const uint32_t tmp = __shfl_sync(full, dv, k, W);
        if (active && !off_diag) {
          if (state == 0) state = tmp & 7;
          else if (!(tmp >> (state + 2) & 1)) state = 0;
          if (state == 0) state = tmp & 7;
          if (state == 0) {
            emit(0), --i, --j;
            if (i < 0 || j < 0) active = false;
          } else {
            if (state == 1) emit(2), --i;
            else emit(1), --j;
            off_diag = true;
            if (i < 0 || j < 0) active = false;
          }
        }
      }
    }
    if (seg_ok && sl == 0) {
      if (run_len > 0) cb.push(run_op, run_len);
      if (i >= 0) cb.push(2, i + 1);
      if (j >= 0) cb.push(1, j + 1);
      if (side != 0 && cb.n <= cb.cap) {
        for (int a = 0; a < cb.n >> 1; ++a) {
          const uint32_t t = cb.ops[a];
          cb.ops[a] = cb.ops[cb.n - 1 - a];
          cb.ops[cb.n - 1 - a] = t;
        }
      }
    }
  }
  if (seg_ok && sl == 0) {
    ExtRec& E = reg->ext[side];
    E.max = ezmax;
    E.mqe_t = mqe_t;
    E.n_cig = cb.n;
    if (cb.n > kSegCig) {
      flag_err(D, D.read_grp[tk.read], E_CIG_SCRATCH);
      E.n_cig = 0;
    } else if (cb.n <= kInlineCig) {
      E.cig_off = -1;
      for (int c = 0; c < cb.n; ++c) E.inl[c] = cig[c];
    } else {
      const long long o = atomicAdd((unsigned long long*)&D.ctr[C_EXTARENA], (unsigned long long)cb.n);
      if (o + cb.n > D.ext_arena_cap) {
        flag_err(D, D.read_grp[tk.read], E_EXT_ARENA);
        E.n_cig = 0;
      } else {
        E.cig_off = (int32_t)o;
        for (int c = 0; c < cb.n; ++c) D.ext_arena[o + c] = cig[c];
      }
    }
    *cells += (long long)m * n;
    *cells_full += (long long)m * n;
  }
  __syncwarp();
}

#ifdef LGR_EXT_HIST
__device__ unsigned long long g_ext_hist[256];  // [m] task count, [128 + m] warp cycles (debug builds only)
#endif

// Phase B1 kernel: the extensions no closed form covered, through the anti-diagonal wavefront.
// Persistent warps drain the three size classes, longest first: one tail per warp (class 2), then
// two (class 1) and four (class 0) tails per warp.  Nothing but DP code lives here.
__global__ void __launch_bounds__(128, LGR_EXT_MINB) k_ext_warp(const __grid_constant__ Dev D) {
  __shared__ __align__(16) uint8_t s_dir[4 * kDirSmemPerWarp];
  __shared__ uint32_t s_cig[4][4 * 24];  // per-segment cigar staging of the sub-warp classes (4 x (2*8+8) = 2 x (2*16+8))
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  uint8_t* dir = D.dir_scratch + (size_t)gwarp * D.dir_per_warp;
  int32_t* Hb = D.bnd_scratch + (size_t)gwarp * D.bnd_per_warp;
  int32_t* Fb = Hb + D.bnd_per_warp / 2;
  uint32_t* wcig = D.wcig_scratch + (size_t)gwarp * D.wcig_cap;
  long long cells = 0, cells_full = 0;
  {
    long long n_task = D.ctr[C_NTASK + 2];
    if (n_task > D.tasks_cap) n_task = D.tasks_cap;
    const TaskRec* tasks = D.tasks + 2 * (size_t)D.tasks_cap;
    for (;;) {
      long long t = 0;
      if (lane == 0) t = atomicAdd((unsigned long long*)&D.ctr[C_TASKPOS + 2], 1ULL);
      t = __shfl_sync(full, t, 0);
      if (t >= n_task) break;
      const TaskRec tk = tasks[t];
      const uint8_t* hapc = D.hap_codes + D.hap_off[tk.hap];
      const int64_t roff = D.read_off[tk.read];
      ReadView rv{D.read_codes + roff, (int)(D.read_off[tk.read + 1] - roff)};
      long long c1 = 0, c2 = 0;
#ifdef LGR_EXT_HIST
      const long long t_begin = clock64();
#endif
      ext_dp_warp(D, D.read_grp[tk.read], &D.regs[tk.reg], tk.side, rv, hapc, dir, s_dir + warp * kDirSmemPerWarp, kDirSmemPerWarp, Hb,
                  Fb, wcig, &c1, &c2);
#ifdef LGR_EXT_HIST
      if (lane == 0) {
        int mb = D.regs[tk.reg].ext[tk.side].m;
        mb = mb > 127 ? 127 : mb;
        atomicAdd(&g_ext_hist[mb], 1ULL);
        atomicAdd(&g_ext_hist[128 + mb], (unsigned long long)(clock64() - t_begin));
      }
#endif
      cells += c1, cells_full += c2;
      __syncwarp();
    }
  }
  {
    long long n_task = D.ctr[C_NTASK + 1];
    if (n_task > D.tasks_cap) n_task = D.tasks_cap;
    const TaskRec* tasks = D.tasks + 1 * (size_t)D.tasks_cap;
    for (;;) {
      long long t = 0;
      if (lane == 0) t = atomicAdd((unsigned long long*)&D.ctr[C_TASKPOS + 1], 2ULL);
      t = __shfl_sync(full, t, 0);
      if (t >= n_task) break;
      ext_dp_multi<16>(D, tasks + t, (int)(n_task - t < 2 ? n_task - t : 2), s_dir + warp * kDirSmemPerWarp, s_cig[warp], &cells, &cells_full);
    }
  }
  {
    long long n_task = D.ctr[C_NTASK + 4];
    if (n_task > D.tasks_cap) n_task = D.tasks_cap;
    const TaskRec* tasks = D.tasks + 4 * (size_t)D.tasks_cap;
    for (;;) {
      long long t = 0;
      if (lane == 0) t = atomicAdd((unsigned long long*)&D.ctr[C_TASKPOS + 4], 2ULL);
      t = __shfl_sync(full, t, 0);
      if (t >= n_task) break;
      ext_dp_multi_t<16>(D, tasks + t, (int)(n_task - t < 2 ? n_task - t : 2), s_dir + warp * kDirSmemPerWarp, s_cig[warp], &cells, &cells_full);
    }
  }
  {
    long long n_task = D.ctr[C_NTASK + 3];
    if (n_task > D.tasks_cap) n_task = D.tasks_cap;
    const TaskRec* tasks = D.tasks + 3 * (size_t)D.tasks_cap;
    for (;;) {
      long long t = 0;
      if (lane == 0) t = atomicAdd((unsigned long long*)&D.ctr[C_TASKPOS + 3], 4ULL);
      t = __shfl_sync(full, t, 0);
      if (t >= n_task) break;
      ext_dp_multi_t<8>(D, tasks + t, (int)(n_task - t < 4 ? n_task - t : 4), s_dir + warp * kDirSmemPerWarp, s_cig[warp], &cells, &cells_full);
    }
  }
  {
    long long n_task = D.ctr[C_NTASK];
    if (n_task > D.tasks_cap) n_task = D.tasks_cap;
    const TaskRec* tasks = D.tasks;
    for (;;) {
      long long t = 0;
      if (lane == 0) t = atomicAdd((unsigned long long*)&D.ctr[C_TASKPOS], 4ULL);
      t = __shfl_sync(full, t, 0);
      if (t >= n_task) break;
      ext_dp_multi<8>(D, tasks + t, (int)(n_task - t < 4 ? n_task - t : 4), s_dir + warp * kDirSmemPerWarp, s_cig[warp], &cells, &cells_full);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {  // per-lane partial sums (lane 0 / the segment leaders) → total
    cells += __shfl_xor_sync(full, cells, o);
    cells_full += __shfl_xor_sync(full, cells_full, o);
  }
  if (lane == 0) {
    atomicAdd((unsigned long long*)&D.ctr[C_CELLS], (unsigned long long)cells);
    atomicAdd((unsigned long long*)&D.ctr[C_CELLSFULL], (unsigned long long)cells_full);
  }
}

}  // namespace

#endif  // LANCET2_B200_LGR_KERNELS_EXT_CUH_
