// lgr_gpu.cu — sm_100a kernels + the C-ABI of the B200 read→haplotype realignment path.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
//
// Pipeline per batch (all on one stream, no host sync inside; DESIGN.md has layout and rooflines):
//   k_encode            ASCII → code bytes (nt4 | Lancet code) for haplotypes and reads
//   k_hap_sketch_warp   one warp per haplotype: minimizer sketch → unsorted table
//   k_hap_sort          one CTA per haplotype: bitonic sort of the table (the "index"), hash-bucket
//                       directory, mid_occ the haplotype would latch
//   k_group_mid         effective mid_occ per group
//   k_read_sketch       one lane per read: sketch + mm_seed_mz_flt
//   k_chain_warp        ONE WARP PER (read, haplotype) PAIR: seeds → anchors → sort → chain DP →
//                       backtrack → regs → SR stretch; regs parked in HBM (RegRec)
//   k_chain_overflow    the same for pairs whose anchors exceed the shared-memory cap (lane per pair)
//   k_finish_warp       one warp per parked pair: extensions (closed forms / anti-diagonal
//                       wavefront DP with shuffle neighbour exchange + traceback), cigar assembly,
//                       mm_fix_cigar, mm_update_extra, filter/sort, NM → lgr_aln
//   k_assign            one lane per (read, variant): local scoring + best-allele selection
// There is no host fallback: every entry point fails with an error code when CUDA fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/lancet_gpu_realign.h"
#include "lgr_core.cuh"

namespace {

using namespace lgr;

__constant__ double c_phred_err[256] = {
#include "phred_lut.inc"
};

// counters (int64 slots in device memory)
enum Ctr {
  C_ITEM = 0, C_NTASK, C_REGS, C_EXTARENA, C_CIGARENA, C_NOVF, C_ERR, C_EVALS, C_ANCH, C_CELLS, C_CELLSFULL,
  C_ALIGNED, C_TASKPOS, C_FINPOS, C_OVFPOS, C_COUNT
};
enum ErrBits { E_REG_ARENA = 1, E_EXT_ARENA = 2, E_CIG_ARENA = 4, E_ANCHOR_CAP = 8, E_CIG_SCRATCH = 16, E_MZ_CAP = 32 };

constexpr int kBucketBits = 9;
constexpr int kBuckets = 1 << kBucketBits;

struct TaskRec {  // one extension that needs the wavefront DP
  int32_t reg, side, read, hap;
};

struct PairReg {  // per pair: its parked RegRecs in the arena (n == 0: nothing to finish)
  int32_t first, n, read, hap;
};

struct Dev {      // everything the kernels need, passed by value
  DevParams P;
  int n_groups, n_haps, n_reads, n_vars;
  int64_t n_pairs, n_assign;
  // inputs
  const int32_t *grp_hap_begin, *grp_read_begin, *grp_var_begin;
  const int64_t *hap_off, *read_off, *var_hap_off;
  const uint8_t *hap_bases, *read_bases, *read_quals;
  const uint32_t* name_hash;
  const int32_t *var_start, *var_len;
  const int8_t* var_allele;
  const int32_t* read_grp;      // [NR]
  const int32_t* hap_grp;       // [NH]
  const int64_t *pair_off, *asg_off;  // [NR+1]
  const int32_t *item_hap, *item_r0, *item_n;  // phase-A work items: (hap, first read, #reads<=32)
  int n_items;
  // derived
  uint8_t *hap_codes, *read_codes;
  uint64_t* idx;                // [hap_off-indexed] sorted minimizer tables
  int32_t *idx_n, *hap_mid;     // [NH]
  uint16_t* bkt;                // [NH][kBuckets+1] start of every hash bucket (top hash bits) in the sorted table
  int bkt_shift;                // hash >> bkt_shift = bucket
  int32_t* grp_mid;             // [G] in: >0 fixed, <=0 latch from first hap; out: effective
  uint64_t* mz_x;               // [read_off-indexed]
  uint32_t* mz_y;
  int32_t* mz_n;                // [NR]
  uint64_t* mz_cnt;             // [NR][2] bucket counters of the read's minimizer hashes
  // phase A workspace
  int32_t* ws;                  // k_chain_overflow: [warps][A_COUNT][cap][32 lanes] workspace
  int ws_cap;
  uint32_t* fin_scratch;        // [warps][2][fin_cap] cigar staging of k_finish_warp
  int fin_cap;
  // parked pairs / tails
  RegRec* regs;  int64_t regs_cap;
  PairReg* pair_reg;             // [n_pairs]
  TaskRec* tasks; int64_t tasks_cap;
  uint32_t* ext_arena; int64_t ext_arena_cap;
  int32_t* ovf_read; int32_t* ovf_hap; int64_t ovf_cap;
  // k_ext_big scratch
  uint8_t* dir_scratch; int64_t dir_per_warp;
  int32_t* bnd_scratch; int64_t bnd_per_warp;   // Hb/Fb boundary rows
  uint32_t* wcig_scratch; int wcig_cap;
  RegRec* wreg_scratch;          // [warps][CAP] regs of the pair a warp is working on
  RadixScratch* rsx_scratch;     // [warps]
  // outputs
  AlnOut* aln; uint32_t* cigar_inline; uint32_t* cigar_arena; int64_t cigar_arena_cap;
  AssignOut* assign;
  long long* ctr;
};

// ---------------------------------------------------------------------------------------
// both buffers come from cudaMalloc (256-byte aligned): 16 bases per lane and iteration
__global__ void k_encode(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int64_t n) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n16 = n >> 4;
  for (int64_t v = tid; v < n16; v += stride) {
    uint4 w = reinterpret_cast<const uint4*>(src)[v];
    uint32_t* p = &w.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t x = p[j];
      p[j] = (uint32_t)encode_base((uint8_t)x) | (uint32_t)encode_base((uint8_t)(x >> 8)) << 8 |
             (uint32_t)encode_base((uint8_t)(x >> 16)) << 16 | (uint32_t)encode_base((uint8_t)(x >> 24)) << 24;
    }
    reinterpret_cast<uint4*>(dst)[v] = w;
  }
  for (int64_t i = (n16 << 4) + tid; i < n; i += stride) dst[i] = encode_base(src[i]);
}

// one lane per haplotype: sketch → table entries (hash<<17 | pos<<1|strand), unsorted
__global__ void k_hap_sketch(const __grid_constant__ Dev D) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= D.n_haps) return;
  const int64_t off = D.hap_off[h];
  const int len = (int)(D.hap_off[h + 1] - off);
  uint64_t* tab = D.idx + off;
  struct XW {  // x arrives first and is staged in the slot, y finalises the packed entry
    uint64_t* t;
    __device__ uint64_t& operator[](int i) const { return t[i]; }
  };
  struct YW {
    uint64_t* t;
    struct Ref {
      uint64_t* p;
      __device__ void operator=(uint32_t y) const { *p = (*p >> 8) << kIdxShift | (uint64_t)y; }
    };
    __device__ Ref operator[](int i) const { return Ref{t + i}; }
  };
  int n = 0;
  if (len > 0) {
    if (D.P.w == 5) {
      int m = 0;
      n = sketch_sr<5>(D.hap_codes + off, len, D.P.k, [&](uint64_t x, uint32_t y) {
        if (m < len) tab[m] = (x >> 8) << kIdxShift | (uint64_t)y;
        ++m;
      });
    } else {
      n = sketch(D.hap_codes + off, len, D.P.w, D.P.k, XW{tab}, YW{tab}, len);
    }
  }
  if (n > len) { n = len; atomicOr((unsigned long long*)&D.ctr[C_ERR], (unsigned long long)E_MZ_CAP); }
  D.idx_n[h] = n;
}

// one WARP per haplotype (odd k, w == 5), one LANE per position.  With odd k no k-mer is its own
// reverse complement, so mm_sketch never skips an iteration and its window state before
// position i is a pure function of the W records before i: the ring holds exactly those, and
// `min` is their right-most minimum (a new record takes over on <=, the rescan keeps the last
// of equals, otherwise nothing to the right of `min` can be <= it).  Every lane rebuilds that
// state from its W predecessors, runs the one `MinimizerWindow::step` of its own position
// (same code as the sequential sketch) and the warp concatenates the emissions in order.
template <typename XT>
__global__ void __launch_bounds__(128) k_hap_sketch_warp(const __grid_constant__ Dev D) {
  constexpr XT kNone = MinimizerWindow<5, XT>::kMax;
  constexpr int W = 5;
  __shared__ XT s_x[4][32 + W];
  __shared__ uint32_t s_y[4][32 + W];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int h = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (h >= D.n_haps) return;
  const int64_t off = D.hap_off[h];
  const int len = (int)(D.hap_off[h + 1] - off);
  const uint8_t* codes = D.hap_codes + off;
  uint64_t* tab = D.idx + off;
  const int k = D.P.k;
  const XT mask = (XT)((1ULL << 2 * k) - 1);
  XT* sx = s_x[warp];
  uint32_t* sy = s_y[warp];
  if (lane < W) sx[lane] = kNone, sy[lane] = UINT32_MAX;  // records "before" position 0
  int n = 0;
  int run_in = 0;  // unambiguous run length ending just before this chunk
  for (int base = 0; base < len; base += 32) {
    const int i = base + lane;
    const int c = i < len ? (codes[i] & 0xf) : 4;
    // run length: distance to the last ambiguous base at or before i (inclusive scan of "last N")
    int lastn = c > 3 ? i : -1;
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(full, lastn, o);
      if (lane >= o && v > lastn) lastn = v;
    }
    const int l = lastn >= 0 ? i - lastn : run_in + lane + 1;
    XT ix = kNone;
    uint32_t iy = UINT32_MAX;
    if (i < len && l >= k) {
      XT k0 = 0, k1 = 0;
      for (int t = 0; t < k; ++t) {
        const XT b = (XT)(codes[i - k + 1 + t] & 0xf);
        k0 = k0 << 2 | b;
        k1 = k1 >> 2 | ((XT)3 ^ b) << 2 * (k - 1);
      }
      const int z = k0 < k1 ? 0 : 1;
      const XT key = z ? k1 : k0;
      const XT hv = sizeof(XT) == 4 ? (XT)hash64_mask_narrow((uint32_t)key, (uint32_t)mask) : (XT)hash64_mask((uint64_t)key, (uint64_t)mask);
      ix = hv << 8 | (XT)k;
      iy = (uint32_t)i << 1 | (uint32_t)z;
    }
    __syncwarp();
    sx[W + lane] = ix, sy[W + lane] = iy;
    run_in = __shfl_sync(full, l, 31);
    __syncwarp();
    // the state before position i, from records i-W .. i-1
    MinimizerWindow<W, XT> win;
    win.k = k;
    win.min_x = kNone, win.min_y = UINT32_MAX, win.min_idx = W - 1;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      win.wx[j] = sx[lane + j], win.wy[j] = sy[lane + j];
      if (win.min_x >= win.wx[j]) win.min_x = win.wx[j], win.min_y = win.wy[j], win.min_idx = j;
    }
    const MinimizerWindow<W, XT> before = win;
    int cnt = 0;
    if (i < len) {
      auto count = [&](XT, uint32_t) { ++cnt; };
      win.step(ix, iy, l, count);
      if (i == len - 1) win.finish(count);
    }
    int pos = cnt;
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(full, pos, o);
      if (lane >= o) pos += v;
    }
    const int total = __shfl_sync(full, pos, 31);
    if (cnt > 0) {
      int m = n + pos - cnt;
      auto put = [&](XT x, uint32_t y) {
        if (m < len) tab[m] = (uint64_t)(x >> 8) << kIdxShift | (uint64_t)y;
        ++m;
      };
      win = before;
      win.step(ix, iy, l, put);
      if (i == len - 1) win.finish(put);
    }
    n += total;
    __syncwarp();
    if (lane < W) sx[lane] = sx[32 + lane], sy[lane] = sy[32 + lane];  // carry the last W records over
  }
  if (lane == 0) {
    if (n > len) { n = len; atomicOr((unsigned long long*)&D.ctr[C_ERR], (unsigned long long)E_MZ_CAP); }
    D.idx_n[h] = n;
  }
}

// one CTA per haplotype: in-place bitonic sort of its table (keys are unique)
__global__ void k_hap_sort(const __grid_constant__ Dev D, float mid_occ_frac, int min_mid, int max_mid) {
  const int h = blockIdx.x;
  uint64_t* tab = D.idx + D.hap_off[h];
  const int n = D.idx_n[h];
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  extern __shared__ uint64_t s_tab[];
  const bool use_smem = np2 <= 2048;
  uint64_t* a = use_smem ? s_tab : tab;
  if (use_smem) {
    for (int i = threadIdx.x; i < np2; i += blockDim.x) s_tab[i] = i < n ? tab[i] : UINT64_MAX;
    __syncthreads();
  }
  // all-ascending bitonic network (first step of every merge pairs i with its mirror
  // i ^ (k-1)); with ascending comparators only, slots >= n act as +inf padding and are
  // simply skipped.
  const int lim = use_smem ? np2 : n;
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1, first = 1; j > 0; j >>= 1, first = 0) {
      for (int i = threadIdx.x; i < lim; i += blockDim.x) {
        const int l = first ? (i ^ (k - 1)) : (i ^ j);
        if (l > i && l < lim) {
          const uint64_t vi = a[i], vl = a[l];
          if (vi > vl) a[i] = vl, a[l] = vi;
        }
      }
      __syncthreads();
    }
  }
  if (use_smem)
    for (int i = threadIdx.x; i < n; i += blockDim.x) tab[i] = s_tab[i];
  // bucket directory: entries are sorted by hash, so the entries whose top hash bits equal b are
  // the contiguous range [bkt[b], bkt[b+1]); k_chain_warp starts its lookups there
  {
    uint16_t* bk = D.bkt + (size_t)h * (kBuckets + 1);
    for (int b = threadIdx.x; b <= kBuckets; b += blockDim.x)
      bk[b] = (uint16_t)idx_lower_bound(a, n, ((uint64_t)b << D.bkt_shift) << kIdxShift);
  }
  // mid_occ this haplotype would latch (mm_idx_cal_max_occ + clamp): histogram of the run
  // lengths of equal hashes; the kk-th smallest run length is read off the cumulative counts.
  __shared__ int s_hist[64];
  __shared__ int s_keys;
  if (threadIdx.x < 64) s_hist[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_keys = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const uint64_t key = a[i] >> kIdxShift;
    if (i == 0 || (a[i - 1] >> kIdxShift) != key) {
      int len = 1;
      while (i + len < n && (a[i + len] >> kIdxShift) == key) ++len;
      atomicAdd(&s_hist[len < 63 ? len : 63], 1);
      atomicAdd(&s_keys, 1);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t mid = INT32_MAX;
    if (mid_occ_frac > 0.f && s_keys > 0) {
      const uint32_t kk = (uint32_t)((1. - (double)mid_occ_frac) * (double)s_keys);
      uint32_t cum = 0;
      int v = 1;
      for (; v < 63; ++v) {
        cum += (uint32_t)s_hist[v];
        if (cum > kk) break;
      }
      if (v < 63) {
        mid = v + 1;
        if (mid < min_mid) mid = min_mid;
        if (max_mid > min_mid && mid > max_mid) mid = max_mid;
      } else {
        mid = hap_mid_occ(a, n, mid_occ_frac, min_mid, max_mid);  // very long runs: exact slow path
      }
    } else {
      if (mid < min_mid) mid = min_mid;
      if (max_mid > min_mid && mid > max_mid) mid = max_mid;
    }
    D.hap_mid[h] = mid;
  }
}

__global__ void k_group_mid(const __grid_constant__ Dev D, int min_mid) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= D.n_groups) return;
  if (D.grp_mid[g] > 0) return;
  const int h0 = D.grp_hap_begin[g];
  D.grp_mid[g] = D.grp_hap_begin[g + 1] > h0 ? D.hap_mid[h0] : min_mid;
}

// one lane per read: sketch.  Independent of the haplotype index, so it runs on a second stream
// next to the haplotype kernels.  Alongside the minimizers it leaves 32 saturating 4-bit
// counters of their hashes (bucket = low hash bits) for k_read_filter.
__global__ void k_read_sketch(const __grid_constant__ Dev D) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= D.n_reads) return;
  const int64_t off = D.read_off[r];
  const int len = (int)(D.read_off[r + 1] - off);
  int n = 0;
  uint64_t cnt_lo = 0, cnt_hi = 0;
  if (len > 0) {
    uint64_t* mzx = D.mz_x + off;
    uint32_t* mzy = D.mz_y + off;
    auto emit = [&](uint64_t x, uint32_t y) {
      if (n < len) mzx[n] = x, mzy[n] = y;
      ++n;
      const int b = (int)(x >> 8) & 31, sh = (b & 15) * 4;
      uint64_t& w = b < 16 ? cnt_lo : cnt_hi;
      if (((w >> sh) & 15) < 15) w += 1ULL << sh;
    };
    if (D.P.w == 5) sketch_sr<5>(D.read_codes + off, len, D.P.k, emit);
    else n = sketch(D.read_codes + off, len, D.P.w, D.P.k, mzx, mzy, len), cnt_lo = cnt_hi = ~0ULL;
    if (n > len) { n = len; atomicOr((unsigned long long*)&D.ctr[C_ERR], (unsigned long long)E_MZ_CAP); }
  }
  D.mz_n[r] = n;
  D.mz_cnt[2 * (size_t)r] = cnt_lo, D.mz_cnt[2 * (size_t)r + 1] = cnt_hi;
}

// one lane per read: mm_seed_mz_flt (q_occ_max = the group's mid_occ, known once the haplotype
// tables exist).  A minimizer can only repeat more than q_occ_max times if its bucket counter
// does, so the O(n^2) filter only runs for the few reads where some bucket got that full.
__global__ void k_read_filter(const __grid_constant__ Dev D) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= D.n_reads || D.P.q_occ_frac <= 0.0f) return;
  const int n = D.mz_n[r];
  const int q_occ_max = D.grp_mid[D.read_grp[r]];
  if (n <= q_occ_max) return;
  const uint64_t cnt_lo = D.mz_cnt[2 * (size_t)r], cnt_hi = D.mz_cnt[2 * (size_t)r + 1];
  bool may_repeat = true;  // counters saturate at 15: above that nothing can be ruled out
  if (q_occ_max < 15) {
    may_repeat = false;
    for (int b = 0; b < 16; ++b)
      may_repeat |= (int)((cnt_lo >> (4 * b)) & 15) > q_occ_max || (int)((cnt_hi >> (4 * b)) & 15) > q_occ_max;
  }
  if (!may_repeat) return;
  const int64_t off = D.read_off[r];
  D.mz_n[r] = seed_mz_flt(D.mz_x + off, D.mz_y + off, n, q_occ_max, D.P.q_occ_frac);
}

__device__ __forceinline__ void write_invalid(AlnOut* o) {
  o->valid = 0, o->score = 0, o->rs = 0, o->re = 0, o->qs = 0, o->qe = 0, o->rev = 0, o->dp_score = 0, o->dp_max = 0;
  o->mlen = 0, o->blen = 0, o->n_ambi = 0, o->nm = 0, o->n_cigar = 0, o->cigar_off = -1, o->n_regs = 0;
}

__device__ __forceinline__ void store_final(const Dev& D, int64_t pair, const AlnOut& a, const uint32_t* cig, int nc) {
  AlnOut o = a;
  if (nc <= LGR_CIGAR_INLINE) {
    o.cigar_off = -1;
    uint32_t* dst = D.cigar_inline + pair * LGR_CIGAR_INLINE;
    for (int i = 0; i < nc; ++i) dst[i] = cig[i];
  } else {
    const long long off = atomicAdd((unsigned long long*)&D.ctr[C_CIGARENA], (unsigned long long)nc);
    if (off + nc > D.cigar_arena_cap) {
      atomicOr((unsigned long long*)&D.ctr[C_ERR], (unsigned long long)E_CIG_ARENA);
      o.cigar_off = -2, o.n_cigar = 0;
    } else {
      o.cigar_off = (int32_t)off;
      for (int i = 0; i < nc; ++i) D.cigar_arena[off + i] = cig[i];
    }
  }
  D.aln[pair] = o;
}

// ---------------------------------------------------------------------------------------
// k_chain_overflow: phase A for the pairs whose seeds/anchors exceeded the shared-memory cap of
// k_chain_warp (tandem repeats: hundreds to thousands of anchors).  One LANE per pair running the
// scalar core (map_chain_phase) over a 16384-anchor HBM workspace interleaved per warp; regs are
// parked exactly like k_chain_warp does.  Exits immediately when the overflow list is empty.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_chain_overflow(const __grid_constant__ Dev D) {
  const int lane = threadIdx.x & 31;
  const int gthread = blockIdx.x * blockDim.x + threadIdx.x;
  const int gwarp = gthread >> 5;
  const long long n_work = D.ctr[C_NOVF] < D.ovf_cap ? D.ctr[C_NOVF] : D.ovf_cap;
  if (n_work == 0) return;
  Ws<32> ws;
  ws.caps = D.ws_cap;  // 16384: fits the low half, chain arrays the same size
  ws.base = D.ws + (size_t)gwarp * A_COUNT * D.ws_cap * 32 + lane;
  RadixScratch rsx;
  ChainCounters ctr{0, 0, 0, 0};
  for (;;) {
    long long item = 0;
    if (lane == 0) item = atomicAdd((unsigned long long*)&D.ctr[C_OVFPOS], 32ULL);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= n_work) break;
    if (item + lane < n_work) {
      const int r = D.ovf_read[item + lane], h = D.ovf_hap[item + lane];
      const int g = D.read_grp[r];
      const int64_t pair = D.pair_off[r] + (h - D.grp_hap_begin[g]);
      const int64_t roff = D.read_off[r], hoff = D.hap_off[h];
      const int qlen = (int)(D.read_off[r + 1] - roff);
      const int hlen = (int)(D.hap_off[h + 1] - hoff);
      ReadView rv{D.read_codes + roff, qlen};
      PairIn pin{rv, D.hap_codes + hoff, hlen, D.idx + hoff, D.idx_n[h], D.mz_x + roff, D.mz_y + roff, D.mz_n[r], D.name_hash[r], D.grp_mid[g]};
      int n_regs = 0;
      const int st = qlen > 0 ? map_chain_phase<32>(D.P, pin, ws, &rsx, &n_regs, &ctr) : kMapNoHit;
      PairReg pr{0, 0, r, h};
      if (st == kMapOverflow) {
        atomicOr((unsigned long long*)&D.ctr[C_ERR], (unsigned long long)E_ANCHOR_CAP);
        write_invalid(&D.aln[pair]);
      } else if (st == kMapNoHit) {
        write_invalid(&D.aln[pair]);
      } else {
        const long long first = atomicAdd((unsigned long long*)&D.ctr[C_REGS], (unsigned long long)n_regs);
        if (first + n_regs > D.regs_cap) {
          atomicOr((unsigned long long*)&D.ctr[C_ERR], (unsigned long long)E_REG_ARENA);
          write_invalid(&D.aln[pair]);
        } else {
          pr = PairReg{(int32_t)first, n_regs, r, h};
          for (int i = 0; i < n_regs; ++i) {
            RegRec* rg = &D.regs[first + i];
            export_reg<32>(ws, i, qlen, rg);
            for (int side = 0; side < 2; ++side) {
              if (rg->ext[side].m <= 0) continue;
              const long long ti = atomicAdd((unsigned long long*)&D.ctr[C_NTASK], 1ULL);
              if (ti < D.tasks_cap) D.tasks[ti] = TaskRec{(int32_t)(first + i), side, r, h};
              else atomicOr((unsigned long long*)&D.ctr[C_ERR], (unsigned long long)E_REG_ARENA);
            }
          }
        }
      }
      D.pair_reg[pair] = pr;
    }
    __syncwarp();
  }
  for (int o = 16; o > 0; o >>= 1) {
    ctr.chain_evals += __shfl_down_sync(0xffffffffu, ctr.chain_evals, o);
    ctr.n_anchors += __shfl_down_sync(0xffffffffu, ctr.n_anchors, o);
  }
  if (lane == 0) {
    atomicAdd((unsigned long long*)&D.ctr[C_EVALS], (unsigned long long)ctr.chain_evals);
    atomicAdd((unsigned long long*)&D.ctr[C_ANCH], (unsigned long long)ctr.n_anchors);
  }
}

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

__device__ __noinline__ void ext_dp_warp(const Dev& D, RegRec* reg, int side, const ReadView& rv, const uint8_t* hapc,
                                         uint8_t* dir_g, uint8_t* dir_s, int dir_s_cap, int32_t* Hb, int32_t* Fb, uint32_t* wcig,
                                         long long* cells, long long* cells_full) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const DevParams& P = D.P;
  const int q = P.q, e = P.e;
  const int32_t sc_match = P.a, sc_mis = -P.b, sc_amb = -P.sc_ambi;  // in registers: the loop's generic stores could alias P
  const int m = reg->ext[side].m, n = reg->ext[side].n;
  int T = prune_cols(P, m, n);
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
    // Per 32 steps every lane fetches one target base (and one packed boundary cell for row
    // blocks > 0): coalesced, off the per-step dependency chain, and the step body stays
    // branch-free — lane l takes its base t[s-l] with one indexed shuffle out of the current or
    // previous 32-base register window, lane 0 takes its boundary input with a broadcast.
    int tprev = 4, tcur = 4, bcur = 0;
    for (int s0 = 0; s0 < nsteps; s0 += 32) {
      const int ti = s0 + lane;
      tprev = tcur;
      tcur = ti < T ? tf(ti) : 4;
      if (blk > 0) bcur = ti < T ? Hb[ti] : 0;
      const int kmax = nsteps - s0 < 32 ? nsteps - s0 : 32;
      for (int k = 0; k < kmax; ++k) {
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
      atomicOr((unsigned long long*)&D.ctr[C_ERR], (unsigned long long)E_CIG_SCRATCH);
      E.n_cig = 0;
    } else if (cb.n <= kInlineCig) {
      E.cig_off = -1;
      for (int c = 0; c < cb.n; ++c) E.inl[c] = wcig[c];
    } else {
      const long long o = atomicAdd((unsigned long long*)&D.ctr[C_EXTARENA], (unsigned long long)cb.n);
      if (o + cb.n > D.ext_arena_cap) {
        atomicOr((unsigned long long*)&D.ctr[C_ERR], (unsigned long long)E_EXT_ARENA);
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
// k_map_warp: ONE WARP PER (read, haplotype) PAIR.  A CTA (8 warps) takes a work item =
// (haplotype, up to 64 consecutive reads of its group), stages the haplotype's code bytes and
// minimizer table in shared memory, and its warps walk the reads.  Per pair, all chain state
// (seeds, anchors, f/p/t, chains, regs: A_COUNT arrays of CAP int32) lives in the warp's slice
// of shared memory:
//   seeds     32 minimizers at a time: binary search in the staged table, ballot-compacted
//   anchors   warp prefix sum over occurrence counts
//   chain DP  for anchor i, 32 predecessors j at a time: comput_sc in parallel, then minimap2's
//             sequential max / max_skip / break automaton reproduced exactly with a prefix-max
//             scan, a (max,+) scan for the saturating skip counter and ballots
//   tail      backtrack → regs → stretch: the scalar core (map_chain_tail) on lane 0
//   extension short tails scalar on lane 0, long tails on the whole warp (ext_dp_warp)
//   finish    scalar core on lane 0 (cigar assembly, mm_fix_cigar, mm_update_extra, NM)
// Pairs whose seeds/anchors exceed CAP are appended to the overflow list (k_chain_overflow).
// ---------------------------------------------------------------------------------------
constexpr int kWarpsPerCta = 4;
constexpr int kMapOkColinear = 2;  // warp_seed_chain: chain DP done by the co-linear closed form

template <int CAP>
__device__ __forceinline__ int warp_seed_chain(const Dev& D, const PairIn& in, const uint16_t* bkt, const Ws<1>& ws,
                                               RadixScratch* rsx, ChainCounters* ctr, int* n_a_out) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const DevParams& P = D.P;
  auto ax = ws.arr(A_AX), ay = ws.arr(A_AY), sx = ws.arr(A_SX), sy = ws.arr(A_SY);
  auto f = ws.arr(A_F), p = ws.arr(A_P), t = ws.arr(A_T), perm = ws.arr(A_PERM);
  auto seedq = ws.arr(A_SEEDQ), seedn = ws.arr(A_SEEDN), seeds = ws.arr(A_SEEDS);
  const int qlen = in.read.qlen;
  // ---- seeds (mm_seed_collect_all) ----
  int n_m = 0, n_high = 0;
  for (int base = 0; base < in.mz_n; base += 32) {
    const int i = base + lane;
    int occ = 0, s0 = 0;
    uint32_t sq = 0;
    if (i < in.mz_n) {
      const uint64_t mx = in.mz_x[i];
      const uint64_t hx = mx >> 8;
      {  // bisection restricted to the minimizer's hash bucket (usually 0-2 entries)
        const int b = (int)(hx >> D.bkt_shift);
        int lo = bkt[b];
        int hi = bkt[b + 1];
        const uint64_t key = hx << kIdxShift;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (in.idx[mid] < key) lo = mid + 1;
          else hi = mid;
        }
        s0 = lo;
      }
      int s1 = s0;
      while (s1 < in.idx_n && (in.idx[s1] >> kIdxShift) == hx && s1 - s0 < 8) ++s1;
      if (s1 - s0 == 8) s1 = idx_lower_bound(in.idx, in.idx_n, (hx + 1) << kIdxShift);  // long run: finish by bisection
      occ = s1 - s0;
      if (occ > 0) {
        uint32_t tandem = 0;
        if (i > 0 && hx == in.mz_x[i - 1] >> 8) tandem = 1;
        if (i < in.mz_n - 1 && hx == in.mz_x[i + 1] >> 8) tandem = 1;
        sq = in.mz_y[i] | (uint32_t)(mx & 0xff) << 20 | tandem << 28;
      }
    }
    const unsigned hit = __ballot_sync(full, occ > 0);
    const int pos = n_m + __popc(hit & ((1u << lane) - 1));
    if (occ > 0 && pos < CAP) seedq[pos] = (int32_t)sq, seedn[pos] = occ, seeds[pos] = s0;
    n_high += __popc(__ballot_sync(full, occ > in.mid_occ));
    n_m += __popc(hit);
  }
  if (n_m > CAP) return kMapOverflow;
  __syncwarp();
  if (n_high > 0) {
    if (lane == 0) seed_select(P, seedq, seedn, n_m, qlen, in.mid_occ);
    __syncwarp();
  }
  // ---- anchors (collect_seed_hits) ----
  int n_a = 0;
  for (int base = 0; base < n_m; base += 32) {
    const int i = base + lane;
    int occ = 0;
    uint32_t sq = 0;
    if (i < n_m) {
      sq = (uint32_t)seedq[i];
      if (!(sq >> 29 & 1)) occ = seedn[i];
    }
    int inc = occ;
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(full, inc, o);
      if (lane >= o) inc += v;
    }
    const int total = __shfl_sync(full, inc, 31);
    if (n_a + total > CAP) return kMapOverflow;
    if (occ > 0) {
      const uint32_t q_pos = sq & 0xfffff, q_span = sq >> 20 & 0xff, tandem = sq >> 28 & 1;
      const int s0 = seeds[i];
      int off = n_a + inc - occ;
      for (int k = 0; k < occ; ++k, ++off) {
        const uint32_t rk = (uint32_t)(in.idx[s0 + k] & ((1u << kIdxShift) - 1));
        const uint32_t rpos = rk >> 1;
        uint32_t x32, qp;
        if ((rk & 1) == (q_pos & 1)) {
          x32 = rpos;
          qp = q_pos >> 1;
        } else {
          x32 = 1u << 31 | rpos;
          qp = (uint32_t)(qlen - ((int32_t)(q_pos >> 1) + 1 - (int32_t)q_span) - 1);
        }
        ax[off] = (int32_t)x32;
        ay[off] = (int32_t)(tandem << 24 | q_span << 16 | (qp & 0xffff));
      }
    }
    n_a += total;
  }
  if (lane == 0 && ctr) ctr->n_anchors += n_a;
  *n_a_out = n_a;
  if (n_a == 0) return kMapNoHit;
  __syncwarp();
  // ---- radix_sort_128x(a): already-sorted fast path, else the exact emulation on lane 0 ----
  {
    // sorted input is left alone by upstream's insertion sort (n <= 64, stable); its in-place
    // radix passes (n > 64) may permute elements whose keys tie, so beyond 64 anchors only a
    // STRICTLY increasing key sequence — which has exactly one sorted order — can skip the emulation
    bool ok = true, strict = true;
    for (int i = lane + 1; i < n_a; i += 32) {
      const uint32_t cur = (uint32_t)ax[i], prev = (uint32_t)ax[i - 1];
      ok &= cur >= prev, strict &= cur > prev;
    }
    const bool sorted = __all_sync(full, ok);
    const bool strictly = __all_sync(full, strict);
    if (sorted && (n_a <= 64 || strictly)) {
      for (int i = lane; i < n_a; i += 32) sx[i] = ax[i], sy[i] = ay[i];
    } else {
      if (lane == 0) {
        for (int i = 0; i < n_a; ++i) perm[i] = i;
        radix_sort_perm(perm, n_a, [&](int32_t id) { return anchor_x64((uint32_t)ax[id]); }, rsx);
        for (int i = 0; i < n_a; ++i) sx[i] = ax[perm[i]], sy[i] = ay[perm[i]];
      }
    }
  }
  for (int i = lane; i < n_a; i += 32) t[i] = 0;
  __syncwarp();
  // ---- mg_lchain_dp ----
  int32_t max_dist_x = P.max_gap_ref > 0 ? P.max_gap_ref : P.max_gap;
  int32_t max_dist_y = qlen > P.max_gap ? qlen : P.max_gap;
  if (max_dist_x < P.bw) max_dist_x = P.bw;
  if (max_dist_y < P.bw) max_dist_y = P.bw;
  int st = 0, max_ii = -1;
  long long n_iter = 0;
  // ---- co-linear fast path ----------------------------------------------------------------
  // All anchors on one strand and one diagonal, strictly increasing, equal spans, no skip
  // penalty: then for every i the best predecessor is i-1 (sc_{i-1} = f[i-1] + min(span, dq) >=
  // f[j] + min(span, dq_ij) for all j < i-1 because sum(min(span, g)) >= min(span, sum g); ties
  // go to the first j scanned, i-1), there is no gap penalty (dd = 0), so f is a prefix sum and
  // p[i] = i-1.  The predecessor scan upstream visits min(i, max_skip + 2) anchors for anchor i
  // (j = i-1 raises the maximum, every further j is stamped by its successor's predecessor link
  // and bumps the skip counter until it exceeds max_skip), which gives its iteration count.
  {
    const uint32_t x0 = (uint32_t)sx[0], y0 = (uint32_t)sy[0];
    const int diag0 = anchor_rpos(x0) - anchor_qpos(y0), span0 = anchor_span(y0);
    bool ok = true;
    for (int i = lane; i < n_a; i += 32) {
      const uint32_t x = (uint32_t)sx[i], y = (uint32_t)sy[i];
      ok &= (x >> 31) == (x0 >> 31) && anchor_rpos(x) - anchor_qpos(y) == diag0 && anchor_span(y) == span0;
      if (i > 0) ok &= anchor_rpos(x) > anchor_rpos((uint32_t)sx[i - 1]);
    }
    const int tot_span = anchor_rpos((uint32_t)sx[n_a - 1]) - anchor_rpos(x0);
    const bool colinear = __all_sync(full, ok) && P.pen_skip == 0.0f && P.max_skip >= 0 && n_a <= P.max_iter &&
                          tot_span <= max_dist_x && tot_span <= max_dist_y && span0 > 0;
    if (colinear) {
      int32_t carry = span0;  // f[0]
      for (int base = 0; base < n_a; base += 32) {
        const int i = base + lane;
        int32_t c = 0;
        if (i > 0 && i < n_a) {
          const int32_t dq = anchor_rpos((uint32_t)sx[i]) - anchor_rpos((uint32_t)sx[i - 1]);
          c = dq < span0 ? dq : span0;
        }
        for (int o = 1; o < 32; o <<= 1) {
          const int32_t v = __shfl_up_sync(full, c, o);
          if (lane >= o) c += v;
        }
        if (i < n_a) f[i] = carry + c, p[i] = i - 1;
        carry += __shfl_sync(full, c, 31);
      }
      const long long cap_it = P.max_skip + 2, nm1 = n_a - 1;  // sum_{i=1}^{n_a-1} min(i, cap_it)
      n_iter = nm1 <= cap_it ? nm1 * (nm1 + 1) / 2 : cap_it * (cap_it + 1) / 2 + (nm1 - cap_it) * cap_it;
      if (lane == 0 && ctr) ctr->chain_evals += n_iter;
      __syncwarp();
      return kMapOkColinear;
    }
  }
  for (int i = 0; i < n_a; ++i) {
    const uint32_t xi = (uint32_t)sx[i], yi = (uint32_t)sy[i];
    while (st < i && ((xi >> 31) != ((uint32_t)sx[st] >> 31) || anchor_rpos(xi) > anchor_rpos((uint32_t)sx[st]) + max_dist_x)) ++st;
    if (i - st > P.max_iter) st = i - P.max_iter;
    int32_t max_f = anchor_span(yi);
    int max_j = -1, n_skip = 0, end_j = st - 1;
    for (int jb = i - 1; jb >= st; jb -= 32) {
      const int j = jb - lane;
      int32_t sc = INT32_MIN;
      int pj = -1;
      if (j >= st) {
        sc = comput_sc(xi, yi, (uint32_t)sx[j], (uint32_t)sy[j], max_dist_x, max_dist_y, P.bw, P.pen_gap, P.pen_skip);
        if (sc != INT32_MIN) sc += f[j], pj = p[j];
      }
      const bool valid = sc != INT32_MIN;
      // stamps t[p[j]] = i of this chunk: predecessors processed earlier (higher j, lower lane) are
      // visible to later lanes after the barrier; stamps written by lanes at/after a break only
      // touch entries that are never read again for this i.
      if (valid && pj >= 0) t[pj] = i;
      __syncwarp();
      const bool stamped = valid && t[j] == i;
      // record setters of the sequential "sc > max_f" test: lanes whose score exceeds max_f and
      // every earlier lane of the chunk.  Usually there is at most one, so they are peeled off
      // with ballots instead of a prefix-max scan.
      unsigned nmask_all = 0;
      {
        int32_t cur = max_f;
        unsigned cand = __ballot_sync(full, valid && sc > cur);
        while (cand) {
          const int l = __ffs(cand) - 1;
          nmask_all |= 1u << l;
          cur = __shfl_sync(full, sc, l);
          cand = __ballot_sync(full, valid && sc > cur) & ~((2u << l) - 1);
        }
      }
      const bool newmax = (nmask_all >> lane) & 1;
      const int ev = newmax ? -1 : (stamped ? 1 : 0);
      const unsigned smask = __ballot_sync(full, ev == 1);
      int last = 31;
      unsigned brk = 0;
      // the skip counter only ever decrements on a record setter; when every record setter comes
      // before the first stamped lane and the counter enters at 0 (or there is none), the
      // decrements are no-ops and the counter is a running popcount of the stamped lanes
      const bool simple = nmask_all == 0 || (n_skip == 0 && (smask == 0 || (31 - __clz(nmask_all)) < (__ffs(smask) - 1)));
      if (simple) {
        const int need = P.max_skip + 1 - n_skip;  // stamped lanes until the break
        const int tot = __popc(smask);
        if (need <= tot) {
          last = (int)__fns(smask, 0, need);
          brk = 1u << last;
          n_skip = P.max_skip + 1;
        } else {
          n_skip += tot;
        }
      } else {
        // general case: n -> max(n + a, b) per lane, composed left to right ((max,+) scan)
        int a = ev, b = ev == 1 ? 1 : 0;
        for (int o = 1; o < 32; o <<= 1) {
          const int au = __shfl_up_sync(full, a, o), bu = __shfl_up_sync(full, b, o);
          if (lane >= o) {
            const int nb = bu + a;
            b = nb > b ? nb : b;
            a = au + a;
          }
        }
        int n_after = n_skip + a;
        if (b > n_after) n_after = b;
        brk = __ballot_sync(full, ev == 1 && n_after > P.max_skip);
        last = brk ? __ffs(brk) - 1 : 31;
        n_skip = __shfl_sync(full, n_after, last);
      }
      const unsigned nmask = nmask_all & (last == 31 ? 0xffffffffu : ((2u << last) - 1));
      if (nmask) {
        const int src = 31 - __clz(nmask);
        max_f = __shfl_sync(full, sc, src);
        max_j = jb - src;
      }
      const int n_in = jb - st + 1 < 32 ? jb - st + 1 : 32;
      if (brk) {
        n_iter += last + 1;
        end_j = jb - last;
        break;
      }
      n_iter += n_in;
      __syncwarp();
    }
    bool far = true;
    if (max_ii >= 0) {
      const uint32_t xm = (uint32_t)sx[max_ii];
      far = (xi >> 31) != (xm >> 31) || anchor_rpos(xi) - anchor_rpos(xm) > max_dist_x;
    }
    if (max_ii < 0 || far) {
      int32_t bf = INT32_MIN;
      int bj = -1;
      for (int j = i - 1 - lane; j >= st; j -= 32)
        if (f[j] > bf) bf = f[j], bj = j;
      for (int o = 16; o > 0; o >>= 1) {
        const int32_t of = __shfl_xor_sync(full, bf, o);
        const int oj = __shfl_xor_sync(full, bj, o);
        if (of > bf || (of == bf && oj > bj)) bf = of, bj = oj;
      }
      max_ii = bj;
    }
    if (max_ii >= 0 && max_ii < end_j) {
      const int32_t tmp = comput_sc(xi, yi, (uint32_t)sx[max_ii], (uint32_t)sy[max_ii], max_dist_x, max_dist_y, P.bw, P.pen_gap, P.pen_skip);
      if (tmp != INT32_MIN && max_f < tmp + f[max_ii]) max_f = tmp + f[max_ii], max_j = max_ii;
    }
    __syncwarp();
    if (lane == 0) f[i] = max_f, p[i] = max_j;
    if (max_ii < 0) {
      max_ii = i;
    } else {
      const uint32_t xm = (uint32_t)sx[max_ii];
      const bool near = (xi >> 31) == (xm >> 31) && anchor_rpos(xi) - anchor_rpos(xm) <= max_dist_x;
      if (near && f[max_ii] < max_f) max_ii = i;
    }
    __syncwarp();
  }
  if (lane == 0 && ctr) ctr->chain_evals += n_iter;
  return kMapOk;
}

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
__device__ __noinline__ void warp_finish_pure_m(const DevParams& P, const uint8_t* read_codes, const uint8_t* hap, int qb, int tb,
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
constexpr int kFinSmemCig = 64;  // cigar ops of a reg kept in shared memory; longer ones use the HBM scratch

__device__ __noinline__ int finish_pair_warp(const Dev& D, const ReadView& rv, const uint8_t* hap, const RegRec* regs, int n_regs,
                                             FinishScratch& fs, TrackBlock* trk, AlnOut* out) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const DevParams& P = D.P;
  const int qlen = rv.qlen;
  int best = -1, n_surv = 0;
  int32_t best_nm = -1;
  uint64_t best_key = 0;
  RegFinal bf;
  bf.n_cig = 0;
  int32_t *s_qs = trk->qs, *s_qe = trk->qe, *s_rs = trk->rs, *s_re = trk->re, *s_score = trk->score;
  uint64_t* s_key = trk->key;
  for (int r = 0; r < n_regs; ++r) {
    RegAsm ra;
    int okf = 1;
    if (lane == 0) okf = assemble_fix_reg(rv, hap, regs[r], D.ext_arena, fs.cig, fs.cap, &ra) ? 1 : 0;
    okf = __shfl_sync(full, okf, 0);
    if (!okf) return -1;
    ra.n = __shfl_sync(full, ra.n, 0), ra.rs = __shfl_sync(full, ra.rs, 0), ra.re = __shfl_sync(full, ra.re, 0);
    ra.qs = __shfl_sync(full, ra.qs, 0), ra.qe = __shfl_sync(full, ra.qe, 0), ra.qb = __shfl_sync(full, ra.qb, 0);
    ra.tb = __shfl_sync(full, ra.tb, 0), ra.dp_ext = __shfl_sync(full, ra.dp_ext, 0);
    __syncwarp();
    const int rev = regs[r].rev, c_qs = regs[r].c_qs, c_qe = regs[r].c_qe, c_rs = regs[r].c_rs;
    const int32_t score = regs[r].score, cnt = regs[r].cnt;
    const uint32_t hash = regs[r].hash;
    RegFinal rf;
    int32_t nm_reg = -1;
    if (ra.n == 1 && rev == 0 && (fs.cig[0] & 0xf) == 0) {
      int32_t core = 0;
      warp_finish_pure_m(P, rv.codes, hap, ra.qb, ra.tb, (int)(fs.cig[0] >> 4), c_qs, c_qe, &rf, &core, &nm_reg);
      rf.dp_score = ra.dp_ext + core;
    } else {
      warp_update_extra(P, rv, rev, hap, ra.qb, ra.tb, fs.cig, ra.n, &rf);
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
      best = r, best_key = key, bf = rf, best_nm = nm_reg;
      uint32_t* tmp = fs.best;
      fs.best = fs.cig;
      fs.cig = tmp;
    }
  }
  out->valid = 0, out->score = 0, out->rs = out->re = out->qs = out->qe = 0, out->rev = 0, out->dp_score = 0;
  out->dp_max = 0, out->mlen = out->blen = out->n_ambi = 0, out->nm = 0, out->n_cigar = 0, out->cigar_off = -1;
  out->n_regs = 0;
  if (best < 0) return 0;
  __syncwarp();
  const int n_ret = n_surv > 1 ? select_returned(P, n_surv, s_qs, s_qe, s_rs, s_re, s_score, s_key) : n_surv;
  out->valid = 1;
  out->score = regs[best].score;
  out->rs = bf.rs, out->re = bf.re, out->qs = bf.qs, out->qe = bf.qe;
  out->rev = regs[best].rev;
  out->dp_score = bf.dp_score, out->dp_max = bf.dp_max, out->mlen = bf.mlen, out->blen = bf.blen;
  out->n_ambi = bf.n_ambi;
  out->n_cigar = bf.n_cig;
  out->n_regs = n_ret;
  out->nm = best_nm >= 0 ? best_nm : warp_edit_distance(rv.codes, qlen, hap, bf.rs, bf.re, bf.qs, fs.best, bf.n_cig);
  return bf.n_cig;
}

// ---------------------------------------------------------------------------------------
// warp_chain_tail_fast: map_chain_tail for the overwhelmingly common shape — every anchor with
// f >= min_sc lies on ONE chain that is accepted.  All lanes execute it uniformly on the shared
// arrays.  Returns kMapOk (one reg, R_* / stretch filled exactly as map_chain_tail would),
// kMapNoHit, or -2 when the shape is different (then lane 0 runs the exact scalar map_chain_tail
// from scratch).
// ---------------------------------------------------------------------------------------
__device__ __noinline__ int warp_chain_tail_fast(const DevParams& P, int qlen, int hap_len, uint32_t name_hash, const Ws<1>& ws, int n_a) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  auto sx = ws.arr(A_SX), sy = ws.arr(A_SY), f = ws.arr(A_F), p = ws.arr(A_P), t = ws.arr(A_T), v = ws.arr(A_V);
  auto cx = ws.arr(A_CX), cy = ws.arr(A_CY);
  // z = anchors with f >= min_sc; its top (processed first upstream) is the max (f, index)
  int n_z = 0;
  int32_t bf = INT32_MIN;
  int bi = -1;
  for (int base = 0; base < n_a; base += 32) {
    const int i = base + lane;
    const bool in = i < n_a && f[i] >= P.min_sc;
    n_z += __popc(__ballot_sync(full, in));
    if (in && (f[i] > bf || (f[i] == bf && i > bi))) bf = f[i], bi = i;
    if (i < n_a) t[i] = 0;
  }
  if (n_z == 0) return kMapNoHit;
  for (int o = 16; o > 0; o >>= 1) {
    const int32_t of = __shfl_xor_sync(full, bf, o);
    const int oi = __shfl_xor_sync(full, bi, o);
    if (of > bf || (of == bf && oi > bi)) bf = of, bi = oi;
  }
  if (n_z > 64) {
    // beyond 64 candidates upstream's in-place radix pass orders ties arbitrarily: the chain it
    // starts with is only certain when the maximum is unique (every other candidate is then
    // swallowed by that chain or sends us to the general path below)
    int ties = 0;
    for (int base = 0; base < n_a; base += 32) {
      const int i = base + lane;
      ties += __popc(__ballot_sync(full, i < n_a && f[i] == bf));
    }
    if (ties > 1) return -2;  // exact scalar path
  }
  __syncwarp();
  // mg_chain_bk_end + collection for the top anchor (t[] is all zero: first chain)
  const int zi = bi;
  const int32_t zx = bf;
  int end_i;
  {
    int i = zi, max_i = zi;
    int32_t max_s = 0;
    do {
      i = p[i];
      const int32_t s = i < 0 ? zx : zx - f[i];
      if (s > max_s) max_s = s, max_i = i;
      else if (max_s - s > P.bw) break;
    } while (i >= 0);
    end_i = max_i;
  }
  int cnt = 0;
  int i;
  for (i = zi; i != end_i; i = p[i]) {
    if (lane == 0) v[cnt] = i, t[i] = 1;
    ++cnt;
  }
  const int32_t sc = i < 0 ? zx : zx - f[i];
  __syncwarp();
  if (!(sc >= P.min_sc && cnt >= P.min_cnt)) return -2;  // rejected top chain: general path
  // any other candidate left?  (they would start further chains upstream)
  {
    bool other = false;
    for (int j = lane; j < n_a; j += 32) other |= f[j] >= P.min_sc && t[j] == 0;
    if (__any_sync(full, other)) return -2;
  }
  // compact_a: ascending anchors of the chain
  for (int j = lane; j < cnt; j += 32) {
    const int id = v[cnt - 1 - j];
    cx[j] = sx[id], cy[j] = sy[id];
  }
  __syncwarp();
  // mm_gen_regs for the single chain
  uint32_t hash = name_hash;
  hash ^= wang_hash((uint32_t)qlen) + wang_hash((uint32_t)P.seed);
  hash = wang_hash(hash);
  const uint32_t x0 = (uint32_t)cx[0], y0 = (uint32_t)cy[0];
  const uint32_t h = (uint32_t)hash64_full((hash64_full(anchor_x64(x0)) + hash64_full(anchor_y64(y0))) ^ hash);
  const int rev = (int)(x0 >> 31);
  // mm_max_stretch (uniform sequential scan over the chain)
  int as1 = 0, cnt1 = cnt;
  if (cnt >= 2) {
    int32_t max_score = -1, max_i = -1, max_len = 0;
    int32_t score = anchor_span(y0), len = 1;
    int k;
    uint32_t px = x0, py = y0;
    for (k = 0; k < cnt - 1; ++k) {
      const uint32_t nx = (uint32_t)cx[k + 1], ny = (uint32_t)cy[k + 1];
      const int32_t q_span = anchor_span(ny);
      const int32_t lr = anchor_rpos(nx) - anchor_rpos(px);
      const int32_t lq = anchor_qpos(ny) - anchor_qpos(py);
      if (lq == lr) {
        score += lq < q_span ? lq : q_span;
        ++len;
      } else {
        if (score > max_score) max_score = score, max_len = len, max_i = k - len + 1;
        score = q_span;
        len = 1;
      }
      px = nx, py = ny;
    }
    if (score > max_score) max_score = score, max_len = len, max_i = k - len + 1;
    as1 = max_i, cnt1 = max_len;
  }
  const uint32_t ys = (uint32_t)cy[as1];
  const int32_t rs = anchor_rpos((uint32_t)cx[as1]) + 1 - anchor_span(ys);
  const int32_t qs = anchor_qpos(ys) + 1 - anchor_span(ys);
  const int32_t re = anchor_rpos((uint32_t)cx[as1 + cnt1 - 1]) + 1;
  const int32_t qe = anchor_qpos((uint32_t)cy[as1 + cnt1 - 1]) + 1;
  int32_t l = qs;
  l += l * P.a + P.end_bonus > P.q ? (l * P.a + P.end_bonus - P.q) / P.e : 0;
  const int32_t rs0 = rs - l > 0 ? rs - l : 0;
  l = qlen - qe;
  l += l * P.a + P.end_bonus > P.q ? (l * P.a + P.end_bonus - P.q) / P.e : 0;
  const int32_t re0 = re + l < hap_len ? re + l : hap_len;
  __syncwarp();
  if (lane == 0) {
    ws.arr(R_SCORE)[0] = sc, ws.arr(R_CNT)[0] = cnt, ws.arr(R_AS)[0] = 0, ws.arr(R_HASH)[0] = (int32_t)((uint32_t)cnt ^ h);
    ws.arr(R_REV)[0] = rev, ws.arr(R_PARENT)[0] = 0, ws.arr(R_ID)[0] = 0;
    ws.arr(R_QS)[0] = qs, ws.arr(R_QE)[0] = qe, ws.arr(R_RS)[0] = rs, ws.arr(R_RE)[0] = re;
    f[0] = rs0, p[0] = re0;
  }
  __syncwarp();
  return kMapOk;
}

// Tail of a co-linear pair in closed form: f is strictly increasing and p[i] = i-1, so the top
// of z is the last anchor, mg_chain_bk_end walks to the start (s = zx - f[i] grows all the way,
// f > 0), the chain is ALL anchors with score f[n_a-1]; if it fails min_sc / min_cnt every other
// candidate is already marked used, so there is no hit.  All anchors share the diagonal, hence
// mm_max_stretch returns the whole chain.  Fills the same R_* slots as map_chain_tail.
__device__ __forceinline__ int warp_chain_tail_colinear(const DevParams& P, int qlen, int hap_len, uint32_t name_hash,
                                                        const Ws<1>& ws, int n_a) {
  const int lane = threadIdx.x & 31;
  auto sx = ws.arr(A_SX), sy = ws.arr(A_SY), f = ws.arr(A_F), p = ws.arr(A_P);
  const int32_t sc = f[n_a - 1];
  if (!(sc >= P.min_sc && n_a >= P.min_cnt)) return kMapNoHit;
  const uint32_t x0 = (uint32_t)sx[0], y0 = (uint32_t)sy[0], x1 = (uint32_t)sx[n_a - 1], y1 = (uint32_t)sy[n_a - 1];
  uint32_t hash = name_hash;
  hash ^= wang_hash((uint32_t)qlen) + wang_hash((uint32_t)P.seed);
  hash = wang_hash(hash);
  const uint32_t h = (uint32_t)hash64_full((hash64_full(anchor_x64(x0)) + hash64_full(anchor_y64(y0))) ^ hash);
  const int32_t span = anchor_span(y0);
  const int32_t rs = anchor_rpos(x0) + 1 - span, qs = anchor_qpos(y0) + 1 - span;
  const int32_t re = anchor_rpos(x1) + 1, qe = anchor_qpos(y1) + 1;
  int32_t l = qs;
  l += l * P.a + P.end_bonus > P.q ? (l * P.a + P.end_bonus - P.q) / P.e : 0;
  const int32_t rs0 = rs - l > 0 ? rs - l : 0;
  l = qlen - qe;
  l += l * P.a + P.end_bonus > P.q ? (l * P.a + P.end_bonus - P.q) / P.e : 0;
  const int32_t re0 = re + l < hap_len ? re + l : hap_len;
  __syncwarp();
  if (lane == 0) {
    ws.arr(R_SCORE)[0] = sc, ws.arr(R_CNT)[0] = n_a, ws.arr(R_AS)[0] = 0, ws.arr(R_HASH)[0] = (int32_t)((uint32_t)n_a ^ h);
    ws.arr(R_REV)[0] = (int32_t)(x0 >> 31), ws.arr(R_PARENT)[0] = 0, ws.arr(R_ID)[0] = 0;
    ws.arr(R_QS)[0] = qs, ws.arr(R_QE)[0] = qe, ws.arr(R_RS)[0] = rs, ws.arr(R_RE)[0] = re;
    f[0] = rs0, p[0] = re0;
  }
  __syncwarp();
  return kMapOk;
}

// Closed forms of an extension tail (both proven in DESIGN.md §4, both checked against the DP by
// the parity tests):
//  * exact match, n >= m: the m query bases equal the first m target bases (no ambiguity codes).
//    The only path reaching m*a is the gap-free diagonal ⇒ max = mqe = m*a at target offset m-1,
//    cigar mM.
//  * overhang, n < m: the first n query bases equal the n target bases and the LAST query base
//    differs from the last target base.  Every path ends in column <= n-1, has at most n matches
//    and at least m-n inserted bases; n*a - (q + e(m-n)) is reached only by "n matches, then one
//    insertion of m-n" (an insertion anywhere earlier would have to match t[n-1] with q[m-1]).
//    ⇒ max = n*a, mqe_t = n-1, cigar nM (m-n)I in alignment order (the left extension reports it
//    outward-in as (m-n)I nM).  Needs a, q, e > 0.
__device__ __forceinline__ bool warp_ext_exact(const DevParams& P, const ReadView& rv, const uint8_t* hapc, RegRec* reg, int side,
                                               int64_t* cells_full) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int m = reg->ext[side].m, n = reg->ext[side].n;
  if (P.a <= 0) return false;
  const int nn = n < m ? n : m;  // bases that must match
  if (n < m && (P.q <= 0 || P.e <= 0)) return false;
  ExtQuery qf{rv, reg->rev, side, reg->c_qs, reg->c_qe};
  ExtTarget tf{hapc, side, reg->c_rs, reg->c_re};
  bool same = true;
  for (int j = lane; j < nn; j += 32) {
    const int qc = qf(j), tc = tf(j);
    same &= qc == tc && qc < 4;
  }
  if (n < m && lane == 0) same &= qf(m - 1) != tf(n - 1);
  if (!__all_sync(full, same)) return false;
  if (lane == 0) {
    ExtRec& E = reg->ext[side];
    E.max = nn * P.a, E.mqe_t = nn - 1, E.cig_off = -1;
    if (n >= m) {
      E.n_cig = 1, E.inl[0] = (uint32_t)m << 4;
    } else {
      E.n_cig = 2;
      const uint32_t mop = (uint32_t)n << 4, iop = (uint32_t)(m - n) << 4 | 1u;
      if (side == 0) E.inl[0] = iop, E.inl[1] = mop;
      else E.inl[0] = mop, E.inl[1] = iop;
    }
    *cells_full += (int64_t)m * n;
  }
  __syncwarp();
  return true;
}

constexpr int kWarpItemReads = 4;  // reads per warp work item (all against one haplotype) in machine-filling batches

// Phase A kernel: seeds → anchors → chain DP → regs, one warp per pair.  Every pair with at least
// one reg is parked: its RegRecs go to the arena and its PairReg slot tells k_finish_warp where.
#ifndef LGR_CHAIN_CARVEOUT
#define LGR_CHAIN_CARVEOUT -1  // cudaSharedmemCarveoutDefault: the chain kernel gains from every KB left to L1 (measured)
#endif
#ifndef LGR_CHAIN_MINB
#define LGR_CHAIN_MINB 9
#endif
#ifndef LGR_EXT_MINB
#define LGR_EXT_MINB 8
#endif
#ifndef LGR_FIN_MINB
#define LGR_FIN_MINB 8
#endif
constexpr int kRegCap = 16;  // chains per pair held in shared memory (more → overflow pass)

template <int CAP>
__global__ void __launch_bounds__(kWarpsPerCta * 32, LGR_CHAIN_MINB) k_chain_warp(const __grid_constant__ Dev D) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int32_t* s_ws = reinterpret_cast<int32_t*>(smem_raw);
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gwarp = blockIdx.x * kWarpsPerCta + warp;
  Ws<1> ws{s_ws + (size_t)warp * Ws<1>::elems(CAP, kRegCap), Ws<1>::pack(CAP, kRegCap)};
  RadixScratch* rsx = D.rsx_scratch + gwarp;
  ChainCounters ctr{0, 0, 0, 0};
  for (;;) {
    long long item = 0;
    if (lane == 0) item = atomicAdd((unsigned long long*)&D.ctr[C_ITEM], 1ULL);
    item = __shfl_sync(full, item, 0);
    if (item >= D.n_items) break;
    const int h = D.item_hap[item], r0 = D.item_r0[item], nr = D.item_n[item];
    const int64_t hoff = D.hap_off[h];
    const int hlen = (int)(D.hap_off[h + 1] - hoff);
    const int idx_n = D.idx_n[h];
    const uint8_t* hapc = D.hap_codes + hoff;
    const uint64_t* idx = D.idx + hoff;
    const int g = D.hap_grp[h];
    const int h_local = h - D.grp_hap_begin[g];
    const int mid_occ = D.grp_mid[g];
    for (int rr = 0; rr < nr; ++rr) {
      const int r = r0 + rr;
      const int64_t pair = D.pair_off[r] + h_local;
      const int64_t roff = D.read_off[r];
      const int qlen = (int)(D.read_off[r + 1] - roff);
      ReadView rv{D.read_codes + roff, qlen};
      PairIn pin{rv, hapc, hlen, idx, idx_n, D.mz_x + roff, D.mz_y + roff, D.mz_n[r], D.name_hash[r], mid_occ};
      int n_a = 0, n_regs = 0;
      int st = qlen > 0 ? warp_seed_chain<CAP>(D, pin, D.bkt + (size_t)h * (kBuckets + 1), ws, rsx, &ctr, &n_a) : kMapNoHit;
      if (st == kMapOkColinear) {
        st = warp_chain_tail_colinear(D.P, qlen, hlen, pin.name_hash, ws, n_a);
        n_regs = 1;
      } else if (st == kMapOk) {
        st = warp_chain_tail_fast(D.P, qlen, hlen, pin.name_hash, ws, n_a);
        n_regs = 1;
        if (st == -2) {
          if (lane == 0) st = map_chain_tail<1>(D.P, qlen, hlen, pin.name_hash, ws, rsx, n_a, &n_regs);
          st = __shfl_sync(full, st, 0);
          n_regs = __shfl_sync(full, n_regs, 0);
        }
      }
      long long first = -1;
      if (lane == 0) {
        if (st == kMapOverflow) {
          const long long o = atomicAdd((unsigned long long*)&D.ctr[C_NOVF], 1ULL);
          if (o < D.ovf_cap) D.ovf_read[o] = r, D.ovf_hap[o] = h;
          else atomicOr((unsigned long long*)&D.ctr[C_ERR], (unsigned long long)E_ANCHOR_CAP);
        } else if (st == kMapNoHit) {
          write_invalid(&D.aln[pair]);
          D.pair_reg[pair] = PairReg{0, 0, r, h};
        } else {
          first = atomicAdd((unsigned long long*)&D.ctr[C_REGS], (unsigned long long)n_regs);
          if (first + n_regs > D.regs_cap) {
            atomicOr((unsigned long long*)&D.ctr[C_ERR], (unsigned long long)E_REG_ARENA);
            write_invalid(&D.aln[pair]);
            D.pair_reg[pair] = PairReg{0, 0, r, h};
            first = -1;
          } else {
            D.pair_reg[pair] = PairReg{(int32_t)first, n_regs, r, h};
            for (int i = 0; i < n_regs; ++i) export_reg<1>(ws, i, qlen, &D.regs[first + i]);
          }
        }
      }
      first = __shfl_sync(full, first, 0);
      __syncwarp();
      if (first >= 0) {
        // extensions: closed forms here (warp-parallel compare), everything else → wavefront queue
        for (int i = 0; i < n_regs; ++i) {
          RegRec* rg = &D.regs[first + i];
          for (int side = 0; side < 2; ++side) {
            if (rg->ext[side].m <= 0) continue;
            if (warp_ext_exact(D.P, rv, hapc, rg, side, &ctr.dp_cells_full)) continue;
            if (lane == 0) {
              const long long ti = atomicAdd((unsigned long long*)&D.ctr[C_NTASK], 1ULL);
              if (ti < D.tasks_cap) D.tasks[ti] = TaskRec{(int32_t)(first + i), side, r, h};
              else atomicOr((unsigned long long*)&D.ctr[C_ERR], (unsigned long long)E_REG_ARENA);
            }
          }
        }
      }
      __syncwarp();
    }
  }
  if (lane == 0) {
    atomicAdd((unsigned long long*)&D.ctr[C_EVALS], (unsigned long long)ctr.chain_evals);
    atomicAdd((unsigned long long*)&D.ctr[C_ANCH], (unsigned long long)ctr.n_anchors);
    atomicAdd((unsigned long long*)&D.ctr[C_CELLSFULL], (unsigned long long)ctr.dp_cells_full);
  }
}

#ifdef LGR_EXT_HIST
__device__ unsigned long long g_ext_hist[256];  // [m] task count, [128 + m] warp cycles (debug builds only)
#endif

// Phase B1 kernel: the extensions no closed form covered, one warp per queued extension, through
// the anti-diagonal wavefront.  Nothing but DP code lives here, so resident warps share one hot loop.
constexpr int kDirSmemPerWarp = 4096;  // direction bytes of one extension kept in shared memory when they fit

__global__ void __launch_bounds__(128, LGR_EXT_MINB) k_ext_warp(const __grid_constant__ Dev D) {
  __shared__ uint8_t s_dir[4 * kDirSmemPerWarp];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  uint8_t* dir = D.dir_scratch + (size_t)gwarp * D.dir_per_warp;
  int32_t* Hb = D.bnd_scratch + (size_t)gwarp * D.bnd_per_warp;
  int32_t* Fb = Hb + D.bnd_per_warp / 2;
  uint32_t* wcig = D.wcig_scratch + (size_t)gwarp * D.wcig_cap;
  long long cells = 0, cells_full = 0;
  long long n_task = D.ctr[C_NTASK];
  if (n_task > D.tasks_cap) n_task = D.tasks_cap;
  for (;;) {
    long long t = 0;
    if (lane == 0) t = atomicAdd((unsigned long long*)&D.ctr[C_TASKPOS], 1ULL);
    t = __shfl_sync(full, t, 0);
    if (t >= n_task) break;
    const TaskRec tk = D.tasks[t];
    const uint8_t* hapc = D.hap_codes + D.hap_off[tk.hap];
    const int64_t roff = D.read_off[tk.read];
    ReadView rv{D.read_codes + roff, (int)(D.read_off[tk.read + 1] - roff)};
    long long c1 = 0, c2 = 0;
#ifdef LGR_EXT_HIST
    const long long t_begin = clock64();
#endif
    ext_dp_warp(D, &D.regs[tk.reg], tk.side, rv, hapc, dir, s_dir + (threadIdx.x >> 5) * kDirSmemPerWarp, kDirSmemPerWarp, Hb, Fb, wcig,
                &c1, &c2);
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
  if (lane == 0) {
    atomicAdd((unsigned long long*)&D.ctr[C_CELLS], (unsigned long long)cells);
    atomicAdd((unsigned long long*)&D.ctr[C_CELLSFULL], (unsigned long long)cells_full);
  }
}

// Phase B2 kernel: one warp per parked pair, every extension already done: the warp-parallel
// finish (assemble, fix, extra, filter, sort) and the final record.
__global__ void __launch_bounds__(128, LGR_FIN_MINB) k_finish_warp(const __grid_constant__ Dev D) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  __shared__ TrackBlock s_trk[4];
  __shared__ uint32_t s_cig[4][2 * kFinSmemCig];
  uint32_t* fin0 = D.fin_scratch + (size_t)gwarp * 2 * D.fin_cap;
  TrackBlock* trk = &s_trk[threadIdx.x >> 5];
  uint32_t* scig = s_cig[threadIdx.x >> 5];
  long long n_aligned = 0;
  for (;;) {
    long long pair = 0;
    if (lane == 0) pair = atomicAdd((unsigned long long*)&D.ctr[C_FINPOS], 1ULL);
    pair = __shfl_sync(full, pair, 0);
    if (pair >= D.n_pairs) break;
    const PairReg d = D.pair_reg[pair];
    if (d.n <= 0) continue;
    const int read = d.read;
    const uint8_t* hapc = D.hap_codes + D.hap_off[d.hap];
    const int64_t roff = D.read_off[read];
    ReadView rv{D.read_codes + roff, (int)(D.read_off[read + 1] - roff)};
    RegRec* regs = D.regs + d.first;
    // cigars live in shared memory; the rare reg with more ops than fit reruns on the HBM scratch
    FinishScratch fs{scig, scig + kFinSmemCig, kFinSmemCig};
    AlnOut ao;
    int nc = finish_pair_warp(D, rv, hapc, regs, d.n, fs, trk, &ao);
    if (nc < 0) {
      __syncwarp();
      fs = FinishScratch{fin0, fin0 + D.fin_cap, D.fin_cap};
      nc = finish_pair_warp(D, rv, hapc, regs, d.n, fs, trk, &ao);
    }
    if (lane == 0) {
      if (nc < 0) {
        atomicOr((unsigned long long*)&D.ctr[C_ERR], (unsigned long long)E_CIG_SCRATCH);
        write_invalid(&D.aln[pair]);
      } else {
        store_final(D, pair, ao, fs.best, nc);
        n_aligned += ao.valid;
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

// =========================================================================================
// host side
// =========================================================================================
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

}  // namespace

struct lgr_ctx {
  int device = 0;
  cudaStream_t stream = nullptr, stream2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  lgr_params prm;
  DevParams P;
  std::string err;
  int sm_count = 0;
  // grow-only device buffers
  std::vector<DevBuf*> all;
  DevBuf b_grp_hap, b_grp_read, b_grp_var, b_hap_off, b_read_off, b_var_hap_off, b_hap_bases, b_read_bases, b_read_quals,
      b_name_hash, b_var_start, b_var_len, b_var_allele, b_read_grp, b_hap_grp, b_pair_off, b_asg_off, b_item_hap, b_item_r0,
      b_item_n, b_hap_codes, b_read_codes, b_idx, b_idx_n, b_hap_mid, b_grp_mid, b_mz_x, b_mz_y, b_mz_n, b_fin, b_regs,
      b_pair_reg, b_ext_arena, b_ovf_read, b_ovf_hap, b_dir, b_bnd, b_wcig, b_aln, b_cig_inline, b_cig_arena, b_assign,
      b_ctr, b_ws_big, b_wreg, b_rsx, b_bkt, b_mz_cnt, b_tasks;
  Dev D;
  bool resident = false;
  int max_read_len = 0, max_hap_len = 0;
  int64_t hap_bytes = 0, read_bytes = 0;
  int ext_blocks = 0, fin_blocks = 0, warp_blocks = 0, warp_cap = 64;
  size_t warp_smem = 0;
  cudaEvent_t ev[12];
  long long* h_ctr = nullptr;  // pinned copy of the device counters
  int launches = 0;
  // asynchronous submissions (lgr_submit/lgr_wait): child contexts, one per slot in flight
  lgr_ctx* slot[LGR_MAX_INFLIGHT] = {};
  bool slot_busy[LGR_MAX_INFLIGHT] = {};
  lgr_batch_out* slot_out[LGR_MAX_INFLIGHT] = {};
  int64_t slot_h2d[LGR_MAX_INFLIGHT] = {}, slot_d2h[LGR_MAX_INFLIGHT] = {};
  // host staging of helper arrays
  std::vector<int32_t> h_read_grp, h_hap_grp, h_item_hap, h_item_r0, h_item_n, h_grp_mid;
  std::vector<int64_t> h_pair_off, h_asg_off;
};

static thread_local std::string g_create_err;

#define LGR_CUDA(ctx, call)                                                                  \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                       \
      return LGR_E_CUDA;                                                                     \
    }                                                                                        \
  } while (0)

static int ensure(lgr_ctx* c, DevBuf& b, size_t bytes) {
  if (bytes < 256) bytes = 256;
  if (b.cap >= bytes) return LGR_OK;
  if (b.p) cudaFree(b.p);
  b.p = nullptr, b.cap = 0;
  size_t want = bytes + bytes / 2;  // cudaFree/cudaMalloc synchronise the device: regrow rarely
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    c->err = std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e);
    return LGR_E_NOMEM;
  }
  b.cap = want;
  return LGR_OK;
}

extern "C" {

int lgr_abi_version(void) { return LGR_ABI_VERSION; }

void lgr_default_params(lgr_params* p) {
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
  p->mask_len = INT_MAX;
  p->cigar_arena_ops = 1 << 20;
}

const char* lgr_strerror(int code) {
  if (code == LGR_E_BUSY) return "all submission slots in flight";
  switch (code) {
    case LGR_OK: return "ok";
    case LGR_E_ARG: return "bad argument or inconsistent batch";
    case LGR_E_CUDA: return "CUDA runtime failure";
    case LGR_E_NO_DEVICE: return "no usable CUDA device (this path has no CPU fallback)";
    case LGR_E_LIMIT: return "a sequence or intermediate exceeds a device-path cap";
    case LGR_E_CIGAR_OVERFLOW: return "cigar overflow arena exhausted";
    case LGR_E_NOMEM: return "out of device memory";
    default: return "unknown error";
  }
}

const char* lgr_last_error(const lgr_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

uint32_t lgr_x31_hash(const char* s) {
  uint32_t h = (uint32_t)(int32_t)(signed char)*s;
  if (h)
    for (++s; *s; ++s) h = (h << 5) - h + (uint32_t)(int32_t)(signed char)*s;
  return h;
}

int lgr_pair_offsets(const lgr_batch_in* in, int64_t* pair_off, int64_t* asg_off) {
  if (!in || !pair_off || !asg_off) return LGR_E_ARG;
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
  return LGR_OK;
}

static int validate_params(const lgr_params* p, std::string& err) {
  auto bad = [&](const char* m) { err = m; return LGR_E_ARG; };
  if (p->k < 1 || 2 * p->k + kIdxShift > 64) return bad("k must satisfy 2k+17 <= 64 (k <= 23)");
  if (p->w < 1 || p->w > kMaxWindow) return bad("w must be in [1,32]");
  if (p->a < 0 || p->b < 0 || p->q < 0 || p->e < 1 || p->sc_ambi < 0) return bad("scores must be non-negative, e >= 1");
  if (p->b > 2 * (p->q + p->e)) return bad("mismatch penalty exceeds 2(q+e): ksw2 returns early, unsupported");
  // regime of the reference: the extension always reaches the query end and never z-drops
  if (p->end_bonus < (p->a + std::max(p->b, p->sc_ambi)) * LGR_MAX_READ_LEN + p->q + p->e * LGR_MAX_READ_LEN)
    return bad("end_bonus too small: device path requires reach_end for every extension (reference uses 10000)");
  if (p->zdrop < 16 * LGR_MAX_READ_LEN) return bad("zdrop too small: device path requires that z-drop never fires (reference uses 100000)");
  if (p->occ_dist != 0 && p->occ_dist < 256) return bad("occ_dist must be 0 or >= 256");
  if (p->min_cnt < 1 || p->best_n < 0) return bad("min_cnt >= 1, best_n >= 0");
  return LGR_OK;
}

int lgr_create(int device_ordinal, const lgr_params* params, lgr_ctx** out) {
  if (!out) return LGR_E_ARG;
  *out = nullptr;
  lgr_params p;
  if (params) p = *params;
  else lgr_default_params(&p);
  int rc = validate_params(&p, g_create_err);
  if (rc != LGR_OK) return rc;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0 || device_ordinal < 0 || device_ordinal >= n_dev) {
    g_create_err = "no usable CUDA device";
    (void)cudaGetLastError();
    return LGR_E_NO_DEVICE;
  }
  lgr_ctx* c = new lgr_ctx();
  c->device = device_ordinal;
  c->prm = p;
  if (cudaSetDevice(device_ordinal) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    g_create_err = "cudaSetDevice/cudaStreamCreate failed";
    delete c;
    return LGR_E_CUDA;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device_ordinal);
  c->sm_count = prop.multiProcessorCount;
  for (auto& e : c->ev) cudaEventCreate(&e);
  if (cudaMallocHost((void**)&c->h_ctr, sizeof(long long) * (C_COUNT + 1)) != cudaSuccess) {
    g_create_err = "cudaMallocHost failed";
    delete c;
    return LGR_E_CUDA;
  }
  cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
  DevParams& d = c->P;
  std::memset(&d, 0, sizeof(d));
  d.k = p.k, d.w = p.w, d.a = p.a, d.b = p.b, d.q = p.q, d.e = p.e, d.sc_ambi = p.sc_ambi, d.bw = p.bw;
  d.end_bonus = p.end_bonus, d.max_gap = p.max_gap, d.max_gap_ref = p.max_gap_ref, d.max_skip = p.max_chain_skip;
  d.max_iter = p.max_chain_iter, d.min_cnt = p.min_cnt, d.min_sc = p.min_chain_score, d.min_dp_max = p.min_dp_max;
  d.max_max_occ = p.max_max_occ, d.occ_dist = p.occ_dist, d.best_n = p.best_n, d.seed = p.seed, d.mask_len = p.mask_len;
  d.pen_gap = (float)(p.chain_gap_scale * 0.01 * p.k);
  d.pen_skip = (float)(p.chain_skip_scale * 0.01 * p.k);
  d.mask_level = p.mask_level, d.pri_ratio = p.pri_ratio, d.max_clip_ratio = p.max_clip_ratio, d.q_occ_frac = p.q_occ_frac;
  d.min_strand_sc = (int32_t)(p.max_gap * 0.8);
  *out = c;
  return LGR_OK;
}

void lgr_destroy(lgr_ctx* c) {
  if (!c) return;
  for (lgr_ctx*& ch : c->slot) {
    lgr_destroy(ch);
    ch = nullptr;
  }
  cudaSetDevice(c->device);
  if (c->h_ctr) cudaFreeHost(c->h_ctr);
  DevBuf* bufs[] = {&c->b_grp_hap, &c->b_grp_read, &c->b_grp_var, &c->b_hap_off, &c->b_read_off, &c->b_var_hap_off, &c->b_hap_bases,
                    &c->b_read_bases, &c->b_read_quals, &c->b_name_hash, &c->b_var_start, &c->b_var_len, &c->b_var_allele,
                    &c->b_read_grp, &c->b_hap_grp, &c->b_pair_off, &c->b_asg_off, &c->b_item_hap, &c->b_item_r0, &c->b_item_n,
                    &c->b_hap_codes, &c->b_read_codes, &c->b_idx, &c->b_idx_n, &c->b_hap_mid, &c->b_grp_mid, &c->b_mz_x, &c->b_mz_y,
                    &c->b_mz_n, &c->b_fin, &c->b_regs, &c->b_pair_reg, &c->b_tasks, &c->b_ext_arena, &c->b_ovf_read,
                    &c->b_ovf_hap, &c->b_dir, &c->b_bnd, &c->b_wcig, &c->b_aln, &c->b_cig_inline, &c->b_cig_arena, &c->b_assign,
                    &c->b_ctr, &c->b_ws_big, &c->b_wreg, &c->b_rsx, &c->b_bkt, &c->b_mz_cnt};
  for (DevBuf* b : bufs)
    if (b->p) cudaFree(b->p);
  for (auto& e : c->ev) cudaEventDestroy(e);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

#ifdef LGR_EXT_HIST
int lgr_debug_ext_hist(unsigned long long* out256, int reset) {
  if (cudaMemcpyFromSymbol(out256, g_ext_hist, sizeof(unsigned long long) * 256) != cudaSuccess) return LGR_E_CUDA;
  if (reset) {
    static const unsigned long long zero[256] = {};
    cudaMemcpyToSymbol(g_ext_hist, zero, sizeof(zero));
  }
  return LGR_OK;
}
#endif

void* lgr_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
    (void)cudaGetLastError();
    return nullptr;
  }
  return p;
}

void lgr_free_pinned(void* p) {
  if (p) cudaFreeHost(p);
}

void* lgr_stream(lgr_ctx* c) { return c ? (void*)c->stream : nullptr; }

static constexpr int kCapBig = 16384;    // anchors per lane in the overflow pass
static constexpr int kBigWarps = 4 * 19; // warps of the overflow pass (workspace = 28 arrays * cap * 4 B per lane, pre-allocated)

static int validate_batch(lgr_ctx* c, const lgr_batch_in* in) {
  auto bad = [&](const std::string& m, int code = LGR_E_ARG) { c->err = m; return code; };
  if (!in || in->n_groups < 0 || in->n_haps < 0 || in->n_reads < 0 || in->n_vars < 0) return bad("null or negative counts");
  if (in->n_groups > 0 && (!in->grp_hap_begin || !in->grp_read_begin || !in->grp_var_begin)) return bad("null group arrays");
  if (in->n_groups == 0) return LGR_OK;
  if (in->grp_hap_begin[0] != 0 || in->grp_read_begin[0] != 0 || in->grp_var_begin[0] != 0) return bad("group prefix arrays must start at 0");
  if (in->grp_hap_begin[in->n_groups] != in->n_haps || in->grp_read_begin[in->n_groups] != in->n_reads ||
      in->grp_var_begin[in->n_groups] != in->n_vars)
    return bad("group prefix arrays do not end at the totals");
  c->max_read_len = 0, c->max_hap_len = 0;
  for (int g = 0; g < in->n_groups; ++g) {
    if (in->grp_hap_begin[g + 1] < in->grp_hap_begin[g] || in->grp_read_begin[g + 1] < in->grp_read_begin[g] ||
        in->grp_var_begin[g + 1] < in->grp_var_begin[g])
      return bad("group prefix arrays must be non-decreasing");
    if (in->grp_read_begin[g + 1] > in->grp_read_begin[g] && in->grp_hap_begin[g + 1] == in->grp_hap_begin[g])
      return bad("a group with reads needs at least the REF haplotype");
  }
  for (int h = 0; h < in->n_haps; ++h) {
    const int64_t l = in->hap_off[h + 1] - in->hap_off[h];
    if (l < 0) return bad("hap_off must be non-decreasing");
    if (l > LGR_MAX_HAP_LEN) return bad("haplotype longer than LGR_MAX_HAP_LEN", LGR_E_LIMIT);
    c->max_hap_len = std::max<int>(c->max_hap_len, (int)l);
  }
  for (int r = 0; r < in->n_reads; ++r) {
    const int64_t l = in->read_off[r + 1] - in->read_off[r];
    if (l < 0) return bad("read_off must be non-decreasing");
    if (l > LGR_MAX_READ_LEN) return bad("read longer than LGR_MAX_READ_LEN", LGR_E_LIMIT);
    c->max_read_len = std::max<int>(c->max_read_len, (int)l);
  }
  {  // the warp wavefront exchanges H/F as packed int16: bound |H| for the longest read
    const int Lm = c->max_read_len, mm = std::max(c->prm.b, c->prm.sc_ambi);
    const int64_t Tm = Lm + ((int64_t)(c->prm.a + mm) * Lm) / c->prm.e + 2;
    if (c->prm.q + (int64_t)c->prm.e * (Tm + Lm) + (int64_t)(mm + c->prm.a) * Lm > 32000)
      return bad("scores could leave the int16 range of the extension kernel for this read length / scoring", LGR_E_LIMIT);
  }
  // ksw2 band (w = 1.5*bw + 1) must never bind
  if ((int64_t)c->max_hap_len + c->max_read_len >= (int64_t)(c->prm.bw * 1.5))
    return bad("haplotype+read length reaches the ksw2 band; unsupported by the device path", LGR_E_LIMIT);
  for (int v = 0; v < in->n_vars; ++v)
    if (in->var_hap_off[v + 1] < in->var_hap_off[v]) return bad("var_hap_off must be non-decreasing");
  return LGR_OK;
}

#define UP(buf, src, bytes)                                                                              \
  do {                                                                                                   \
    if ((rc = ensure(c, c->buf, (bytes))) != LGR_OK) return rc;                                          \
    if ((bytes) > 0) LGR_CUDA(c, cudaMemcpyAsync(c->buf.p, (src), (bytes), cudaMemcpyHostToDevice, c->stream)); \
    h2d += (bytes);                                                                                      \
  } while (0)

static int upload_impl(lgr_ctx* c, const lgr_batch_in* in, int64_t* h2d_bytes) {
  int rc = validate_batch(c, in);
  if (rc != LGR_OK) return rc;
  LGR_CUDA(c, cudaSetDevice(c->device));
  int64_t h2d = 0;
  const int G = in->n_groups, NH = in->n_haps, NR = in->n_reads, NV = in->n_vars;
  // host helper arrays
  c->h_read_grp.resize(NR + 1), c->h_hap_grp.resize(NH + 1), c->h_pair_off.resize(NR + 1), c->h_asg_off.resize(NR + 1);
  c->h_grp_mid.resize(G + 1);
  c->h_item_hap.clear(), c->h_item_r0.clear(), c->h_item_n.clear();
  int64_t po = 0, ao = 0;
  // reads per warp work item: several reads of one haplotype amortise the item's fixed cost when
  // the batch fills the machine many times over; a small batch (one Genotype() call) is latency
  // bound instead and wants every pair on its own warp
  int64_t pairs_total = 0;
  for (int g = 0; g < G; ++g)
    pairs_total += (int64_t)(in->grp_read_begin[g + 1] - in->grp_read_begin[g]) * (in->grp_hap_begin[g + 1] - in->grp_hap_begin[g]);
  const int64_t warps_resident = (int64_t)c->sm_count * 36;
  const int item_reads = pairs_total >= 8 * warps_resident ? kWarpItemReads : (pairs_total >= 3 * warps_resident ? 2 : 1);
  for (int g = 0; g < G; ++g) {
    const int h0 = in->grp_hap_begin[g], h1 = in->grp_hap_begin[g + 1];
    const int r0 = in->grp_read_begin[g], r1 = in->grp_read_begin[g + 1];
    const int V = in->grp_var_begin[g + 1] - in->grp_var_begin[g];
    for (int h = h0; h < h1; ++h) c->h_hap_grp[h] = g;
    for (int r = r0; r < r1; ++r) {
      c->h_read_grp[r] = g, c->h_pair_off[r] = po, c->h_asg_off[r] = ao;
      po += h1 - h0, ao += V;
    }
    for (int h = h0; h < h1; ++h)
      for (int r = r0; r < r1; r += item_reads) {
        c->h_item_hap.push_back(h), c->h_item_r0.push_back(r), c->h_item_n.push_back(std::min(item_reads, r1 - r));
      }
    int32_t mid = c->prm.mid_occ;
    if (in->grp_mid_occ && in->grp_mid_occ[g] > 0) mid = in->grp_mid_occ[g];
    c->h_grp_mid[g] = mid;
  }
  c->h_pair_off[NR] = po, c->h_asg_off[NR] = ao;
  if (po > (int64_t)1 << 30) { c->err = "more than 2^30 (read, haplotype) pairs in one batch"; return LGR_E_LIMIT; }
  const int64_t hap_bytes = NH ? in->hap_off[NH] : 0, read_bytes = NR ? in->read_off[NR] : 0;
  const int64_t nvh = NV ? in->var_hap_off[NV] : 0;
  c->hap_bytes = hap_bytes, c->read_bytes = read_bytes;
  UP(b_grp_hap, in->grp_hap_begin, sizeof(int32_t) * (G + 1));
  UP(b_grp_read, in->grp_read_begin, sizeof(int32_t) * (G + 1));
  UP(b_grp_var, in->grp_var_begin, sizeof(int32_t) * (G + 1));
  UP(b_hap_off, in->hap_off, sizeof(int64_t) * (NH + 1));
  UP(b_read_off, in->read_off, sizeof(int64_t) * (NR + 1));
  UP(b_var_hap_off, in->var_hap_off, sizeof(int64_t) * (NV + 1));
  UP(b_hap_bases, in->hap_bases, (size_t)hap_bytes);
  UP(b_read_bases, in->read_bases, (size_t)read_bytes);
  UP(b_read_quals, in->read_quals, (size_t)read_bytes);
  UP(b_name_hash, in->read_name_hash, sizeof(uint32_t) * NR);
  UP(b_var_start, in->var_start, sizeof(int32_t) * nvh);
  UP(b_var_len, in->var_len, sizeof(int32_t) * nvh);
  UP(b_var_allele, in->var_allele, (size_t)nvh);
  UP(b_read_grp, c->h_read_grp.data(), sizeof(int32_t) * NR);
  UP(b_hap_grp, c->h_hap_grp.data(), sizeof(int32_t) * NH);
  UP(b_pair_off, c->h_pair_off.data(), sizeof(int64_t) * (NR + 1));
  UP(b_asg_off, c->h_asg_off.data(), sizeof(int64_t) * (NR + 1));
  UP(b_item_hap, c->h_item_hap.data(), sizeof(int32_t) * c->h_item_hap.size());
  UP(b_item_r0, c->h_item_r0.data(), sizeof(int32_t) * c->h_item_r0.size());
  UP(b_item_n, c->h_item_n.data(), sizeof(int32_t) * c->h_item_n.size());
  UP(b_grp_mid, c->h_grp_mid.data(), sizeof(int32_t) * G);
  // derived / scratch / outputs
  const int64_t n_pairs = po, n_assign = ao;
  if ((rc = ensure(c, c->b_hap_codes, hap_bytes)) || (rc = ensure(c, c->b_read_codes, read_bytes)) ||
      (rc = ensure(c, c->b_idx, sizeof(uint64_t) * hap_bytes)) || (rc = ensure(c, c->b_idx_n, sizeof(int32_t) * NH)) ||
      (rc = ensure(c, c->b_hap_mid, sizeof(int32_t) * NH)) || (rc = ensure(c, c->b_bkt, sizeof(uint16_t) * (size_t)NH * (kBuckets + 1))) || (rc = ensure(c, c->b_mz_x, sizeof(uint64_t) * read_bytes)) ||
      (rc = ensure(c, c->b_mz_y, sizeof(uint32_t) * read_bytes)) || (rc = ensure(c, c->b_mz_n, sizeof(int32_t) * NR)) || (rc = ensure(c, c->b_mz_cnt, sizeof(uint64_t) * 2 * (size_t)NR)))
    return rc;
  const int fin_cap = 2 * c->max_read_len + 16;
  const int Lm = std::max(c->max_read_len, 1);
  const int Tmax = Lm + ((c->prm.a + std::max(c->prm.b, c->prm.sc_ambi)) * Lm) / c->prm.e + 2;
  // warp-per-pair kernel: CAP anchors per pair in shared memory
  c->warp_cap = c->max_read_len <= 160 ? 64 : 128;
  c->warp_smem = (size_t)kWarpsPerCta * Ws<1>::elems(c->warp_cap, kRegCap) * sizeof(int32_t);
  {
    int per_sm = 0;
    cudaError_t e1, e2;
    if (c->warp_cap == 64) {
      e1 = cudaFuncSetAttribute(k_chain_warp<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->warp_smem);
      cudaFuncSetAttribute(k_chain_warp<64>, cudaFuncAttributePreferredSharedMemoryCarveout, LGR_CHAIN_CARVEOUT);
      e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_chain_warp<64>, kWarpsPerCta * 32, c->warp_smem);
    } else {
      e1 = cudaFuncSetAttribute(k_chain_warp<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->warp_smem);
      cudaFuncSetAttribute(k_chain_warp<128>, cudaFuncAttributePreferredSharedMemoryCarveout, LGR_CHAIN_CARVEOUT);
      e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_chain_warp<128>, kWarpsPerCta * 32, c->warp_smem);
    }
    if (e1 != cudaSuccess || e2 != cudaSuccess || per_sm < 1) {
      c->err = "k_chain_warp does not fit on this device (shared memory / registers)";
      return LGR_E_CUDA;
    }
    c->warp_blocks = c->sm_count * per_sm;
  }
  {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ext_warp, 128, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
    c->ext_blocks = c->sm_count * per_sm;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_finish_warp, 128, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
    c->fin_blocks = c->sm_count * per_sm;
  }
  const int64_t ext_warps = std::max<int64_t>((int64_t)std::max(c->ext_blocks, c->fin_blocks) * 4, (int64_t)c->warp_blocks * kWarpsPerCta);
  const int64_t dir_per_warp = (int64_t)((Lm + 31) / 32) * (Tmax + 32) * 32;
  const int64_t bnd_per_warp = 2 * (int64_t)(Tmax + 32);
  const int wcig_cap = 2 * Lm + 8;
  const int64_t regs_cap = n_pairs + n_pairs / 4 + 1024;
  const int64_t ext_arena_cap = 4 * n_pairs + (1 << 20);
  const int64_t cig_arena_cap = std::max<int64_t>(c->prm.cigar_arena_ops, 1024);
  if ((rc = ensure(c, c->b_wreg, sizeof(RegRec) * (size_t)ext_warps * c->warp_cap)) ||
      (rc = ensure(c, c->b_rsx, sizeof(RadixScratch) * (size_t)ext_warps)) ||
      (rc = ensure(c, c->b_fin, sizeof(uint32_t) * (size_t)ext_warps * 2 * fin_cap)) ||
      (rc = ensure(c, c->b_regs, sizeof(RegRec) * (size_t)regs_cap)) || (rc = ensure(c, c->b_pair_reg, sizeof(PairReg) * (size_t)n_pairs)) ||
      (rc = ensure(c, c->b_tasks, sizeof(TaskRec) * (size_t)regs_cap * 2)) ||
      (rc = ensure(c, c->b_ext_arena, sizeof(uint32_t) * (size_t)ext_arena_cap)) ||
      (rc = ensure(c, c->b_ovf_read, sizeof(int32_t) * (size_t)(n_pairs + 32))) ||
      (rc = ensure(c, c->b_ovf_hap, sizeof(int32_t) * (size_t)(n_pairs + 32))) ||
      (rc = ensure(c, c->b_dir, (size_t)ext_warps * dir_per_warp)) ||
      (rc = ensure(c, c->b_bnd, sizeof(int32_t) * (size_t)ext_warps * bnd_per_warp)) ||
      (rc = ensure(c, c->b_wcig, sizeof(uint32_t) * (size_t)ext_warps * wcig_cap)) ||
      (rc = ensure(c, c->b_aln, sizeof(AlnOut) * (size_t)n_pairs)) ||
      (rc = ensure(c, c->b_cig_inline, sizeof(uint32_t) * (size_t)n_pairs * LGR_CIGAR_INLINE)) ||
      (rc = ensure(c, c->b_cig_arena, sizeof(uint32_t) * (size_t)cig_arena_cap)) ||
      (rc = ensure(c, c->b_assign, sizeof(AssignOut) * (size_t)n_assign)) || (rc = ensure(c, c->b_ctr, sizeof(long long) * C_COUNT)) ||
      (rc = ensure(c, c->b_ws_big, sizeof(int32_t) * (size_t)(kBigWarps * 32) * A_COUNT * kCapBig)))
    return rc;
  Dev& D = c->D;
  std::memset(&D, 0, sizeof(D));
  D.P = c->P;
  D.n_groups = G, D.n_haps = NH, D.n_reads = NR, D.n_vars = NV, D.n_pairs = n_pairs, D.n_assign = n_assign;
  D.grp_hap_begin = (int32_t*)c->b_grp_hap.p, D.grp_read_begin = (int32_t*)c->b_grp_read.p, D.grp_var_begin = (int32_t*)c->b_grp_var.p;
  D.hap_off = (int64_t*)c->b_hap_off.p, D.read_off = (int64_t*)c->b_read_off.p, D.var_hap_off = (int64_t*)c->b_var_hap_off.p;
  D.hap_bases = (uint8_t*)c->b_hap_bases.p, D.read_bases = (uint8_t*)c->b_read_bases.p, D.read_quals = (uint8_t*)c->b_read_quals.p;
  D.name_hash = (uint32_t*)c->b_name_hash.p;
  D.var_start = (int32_t*)c->b_var_start.p, D.var_len = (int32_t*)c->b_var_len.p, D.var_allele = (int8_t*)c->b_var_allele.p;
  D.read_grp = (int32_t*)c->b_read_grp.p, D.hap_grp = (int32_t*)c->b_hap_grp.p;
  D.pair_off = (int64_t*)c->b_pair_off.p, D.asg_off = (int64_t*)c->b_asg_off.p;
  D.item_hap = (int32_t*)c->b_item_hap.p, D.item_r0 = (int32_t*)c->b_item_r0.p, D.item_n = (int32_t*)c->b_item_n.p;
  D.n_items = (int)c->h_item_hap.size();
  D.hap_codes = (uint8_t*)c->b_hap_codes.p, D.read_codes = (uint8_t*)c->b_read_codes.p;
  D.idx = (uint64_t*)c->b_idx.p, D.idx_n = (int32_t*)c->b_idx_n.p, D.hap_mid = (int32_t*)c->b_hap_mid.p;
  D.grp_mid = (int32_t*)c->b_grp_mid.p;
  D.bkt = (uint16_t*)c->b_bkt.p;
  D.bkt_shift = 2 * c->prm.k > kBucketBits ? 2 * c->prm.k - kBucketBits : 0;
  D.mz_x = (uint64_t*)c->b_mz_x.p, D.mz_y = (uint32_t*)c->b_mz_y.p, D.mz_n = (int32_t*)c->b_mz_n.p;
  D.mz_cnt = (uint64_t*)c->b_mz_cnt.p;
  D.ws = nullptr, D.ws_cap = 0;
  D.wreg_scratch = (RegRec*)c->b_wreg.p, D.rsx_scratch = (RadixScratch*)c->b_rsx.p;
  D.fin_scratch = (uint32_t*)c->b_fin.p, D.fin_cap = fin_cap;
  D.regs = (RegRec*)c->b_regs.p, D.regs_cap = regs_cap;
  D.pair_reg = (PairReg*)c->b_pair_reg.p;
  D.tasks = (TaskRec*)c->b_tasks.p, D.tasks_cap = regs_cap * 2;
  D.ext_arena = (uint32_t*)c->b_ext_arena.p, D.ext_arena_cap = ext_arena_cap;
  D.ovf_read = (int32_t*)c->b_ovf_read.p, D.ovf_hap = (int32_t*)c->b_ovf_hap.p, D.ovf_cap = n_pairs;
  D.dir_scratch = (uint8_t*)c->b_dir.p, D.dir_per_warp = dir_per_warp;
  D.bnd_scratch = (int32_t*)c->b_bnd.p, D.bnd_per_warp = bnd_per_warp;
  D.wcig_scratch = (uint32_t*)c->b_wcig.p, D.wcig_cap = wcig_cap;
  D.aln = (AlnOut*)c->b_aln.p, D.cigar_inline = (uint32_t*)c->b_cig_inline.p, D.cigar_arena = (uint32_t*)c->b_cig_arena.p;
  D.cigar_arena_cap = cig_arena_cap;
  D.assign = (AssignOut*)c->b_assign.p;
  D.ctr = (long long*)c->b_ctr.p;
  c->resident = true;
  if (h2d_bytes) *h2d_bytes = h2d;
  return LGR_OK;
}

// enqueue one pass of the whole path on the context's streams; no host synchronisation
static int run_launch(lgr_ctx* c) {
  if (!c->resident) { c->err = "no batch uploaded"; return LGR_E_ARG; }
  LGR_CUDA(c, cudaSetDevice(c->device));
  Dev& D = c->D;
  cudaStream_t s = c->stream;
  int launches = 0;
  // the group mid_occ array is an in/out: restore the requested values before every run
  if (D.n_groups > 0)
    LGR_CUDA(c, cudaMemcpyAsync(D.grp_mid, c->h_grp_mid.data(), sizeof(int32_t) * D.n_groups, cudaMemcpyHostToDevice, s));
  LGR_CUDA(c, cudaMemsetAsync(D.ctr, 0, sizeof(long long) * C_COUNT, s));
  cudaEventRecord(c->ev[0], s);
  const int64_t hb = c->hap_bytes, rb = c->read_bytes;
  if (D.n_pairs > 0) {
    const int enc_blocks = c->sm_count * 8;
    // read side (encode + sketch) on the second stream, haplotype side (encode, sketch, sort,
    // mid_occ) on the main one; they join before the minimizer filter
    cudaStream_t s2 = c->stream2;
    cudaEventRecord(c->ev_fork, s);
    cudaStreamWaitEvent(s2, c->ev_fork, 0);
    k_encode<<<enc_blocks, 256, 0, s2>>>(D.read_bases, D.read_codes, rb);
    k_read_sketch<<<(D.n_reads + 127) / 128, 128, 0, s2>>>(D);
    cudaEventRecord(c->ev_join, s2);
    k_encode<<<enc_blocks, 256, 0, s>>>(D.hap_bases, D.hap_codes, hb);
    if (D.P.w == 5 && (D.P.k & 1) && 2 * D.P.k + 8 <= 32) k_hap_sketch_warp<uint32_t><<<(D.n_haps + 3) / 4, 128, 0, s>>>(D);
    else if (D.P.w == 5 && (D.P.k & 1)) k_hap_sketch_warp<uint64_t><<<(D.n_haps + 3) / 4, 128, 0, s>>>(D);
    else k_hap_sketch<<<(D.n_haps + 63) / 64, 64, 0, s>>>(D);
    k_hap_sort<<<D.n_haps, 128, 2048 * sizeof(uint64_t), s>>>(D, c->prm.mid_occ_frac, c->prm.min_mid_occ, c->prm.max_mid_occ);
    k_group_mid<<<(D.n_groups + 127) / 128, 128, 0, s>>>(D, c->prm.min_mid_occ);
    cudaStreamWaitEvent(s, c->ev_join, 0);
    launches += 6;
    cudaEventRecord(c->ev[1], s);
    k_read_filter<<<(D.n_reads + 127) / 128, 128, 0, s>>>(D);
    launches += 1;
    cudaEventRecord(c->ev[2], s);
    if (c->warp_cap == 64) k_chain_warp<64><<<c->warp_blocks, kWarpsPerCta * 32, c->warp_smem, s>>>(D);
    else k_chain_warp<128><<<c->warp_blocks, kWarpsPerCta * 32, c->warp_smem, s>>>(D);
    launches += 1;
    cudaEventRecord(c->ev[9], s);
    {
      // overflow pass (pairs whose seeds/anchors exceed the shared-memory cap): lane-per-pair over
      // a large HBM workspace; exits immediately when the list is empty (no host round trip)
      Dev D2 = D;
      D2.ws = (int32_t*)c->b_ws_big.p, D2.ws_cap = kCapBig;
      k_chain_overflow<<<kBigWarps / 4, 128, 0, s>>>(D2);
      launches += 1;
    }
    k_ext_warp<<<c->ext_blocks, 128, 0, s>>>(D);
    k_finish_warp<<<c->fin_blocks, 128, 0, s>>>(D);
    launches += 2;
    cudaEventRecord(c->ev[3], s);
    if (D.n_assign > 0) {
      k_assign<<<(unsigned)((D.n_assign + 127) / 128), 128, 0, s>>>(D);
      launches += 1;
    }
  } else {
    cudaEventRecord(c->ev[1], s), cudaEventRecord(c->ev[2], s), cudaEventRecord(c->ev[9], s), cudaEventRecord(c->ev[3], s);
  }
  cudaEventRecord(c->ev[4], s);
  LGR_CUDA(c, cudaMemcpyAsync(c->h_ctr, D.ctr, sizeof(long long) * C_COUNT, cudaMemcpyDeviceToHost, s));
  c->launches = launches;
  return LGR_OK;
}

// after the stream has drained: statistics and the device-side limit flags
static int run_finish(lgr_ctx* c, lgr_stats* st) {
  LGR_CUDA(c, cudaGetLastError());
  Dev& D = c->D;
  const long long* hctr = c->h_ctr;
  const int launches = c->launches;
  if (st) {
    float ms;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[4]); st->ms_kernels = ms;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]); st->ms_k_index = ms;
    cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]); st->ms_k_sketch = ms;
    cudaEventElapsedTime(&ms, c->ev[2], c->ev[9]); st->ms_k_map = ms;
    cudaEventElapsedTime(&ms, c->ev[9], c->ev[3]); st->ms_k_ext = ms;
    cudaEventElapsedTime(&ms, c->ev[3], c->ev[4]); st->ms_k_assign = ms;
    st->n_pairs = D.n_pairs, st->n_aligned = hctr[C_ALIGNED];
    st->dp_cells = hctr[C_CELLS], st->dp_cells_full = hctr[C_CELLSFULL];
    st->chain_evals = hctr[C_EVALS], st->n_anchors = hctr[C_ANCH];
    st->kernel_launches = launches;
  }
  if (hctr[C_ERR]) {
    char buf[160];
    snprintf(buf, sizeof(buf), "device path limit hit (flags 0x%llx: 1 reg arena, 2 ext arena, 4 cigar arena, 8 anchor cap, 16 cigar scratch, 32 minimizer cap)",
             hctr[C_ERR]);
    c->err = buf;
    return (hctr[C_ERR] == E_CIG_ARENA) ? LGR_E_CIGAR_OVERFLOW : LGR_E_LIMIT;
  }
  return LGR_OK;
}

static int run_impl(lgr_ctx* c, lgr_stats* st) {
  int rc = run_launch(c);
  if (rc != LGR_OK) return rc;
  LGR_CUDA(c, cudaStreamSynchronize(c->stream));
  return run_finish(c, st);
}

// enqueue the device→host copies whose sizes are known up front (records, inline cigars,
// assignments); the overflow cigar arena follows in download_finish once its fill is known
static int download_launch(lgr_ctx* c, lgr_batch_out* out, int64_t* d2h_bytes) {
  if (!c->resident || !out) { c->err = "nothing to download"; return LGR_E_ARG; }
  LGR_CUDA(c, cudaSetDevice(c->device));
  Dev& D = c->D;
  int64_t d2h = 0;
  if (out->aln) {
    if (out->n_pairs < D.n_pairs) { c->err = "out->n_pairs too small"; return LGR_E_ARG; }
    LGR_CUDA(c, cudaMemcpyAsync(out->aln, D.aln, sizeof(AlnOut) * D.n_pairs, cudaMemcpyDeviceToHost, c->stream));
    d2h += sizeof(AlnOut) * D.n_pairs;
    if (out->cigar_inline) {
      LGR_CUDA(c, cudaMemcpyAsync(out->cigar_inline, D.cigar_inline, sizeof(uint32_t) * D.n_pairs * LGR_CIGAR_INLINE, cudaMemcpyDeviceToHost, c->stream));
      d2h += sizeof(uint32_t) * D.n_pairs * LGR_CIGAR_INLINE;
    }
    LGR_CUDA(c, cudaMemcpyAsync(c->h_ctr + C_COUNT, D.ctr + C_CIGARENA, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
  }
  if (out->assign && D.n_assign > 0) {
    if (out->n_assign < D.n_assign) { c->err = "out->n_assign too small"; return LGR_E_ARG; }
    LGR_CUDA(c, cudaMemcpyAsync(out->assign, D.assign, sizeof(AssignOut) * D.n_assign, cudaMemcpyDeviceToHost, c->stream));
    d2h += sizeof(AssignOut) * D.n_assign;
  }
  if (d2h_bytes) *d2h_bytes = d2h;
  return LGR_OK;
}

// stream already drained by the caller
static int download_finish(lgr_ctx* c, lgr_batch_out* out, int64_t* d2h_bytes) {
  Dev& D = c->D;
  if (out->aln) {
    const long long used = c->h_ctr[C_COUNT];
    out->cigar_arena_used = used;
    if (used > 0) {
      if (!out->cigar_arena || out->cigar_arena_cap < used) { c->err = "host cigar arena too small"; return LGR_E_CIGAR_OVERFLOW; }
      LGR_CUDA(c, cudaMemcpyAsync(out->cigar_arena, D.cigar_arena, sizeof(uint32_t) * used, cudaMemcpyDeviceToHost, c->stream));
      LGR_CUDA(c, cudaStreamSynchronize(c->stream));
      if (d2h_bytes) *d2h_bytes += sizeof(uint32_t) * used;
    }
  }
  return LGR_OK;
}

static int download_impl(lgr_ctx* c, lgr_batch_out* out, int64_t* d2h_bytes) {
  int rc = download_launch(c, out, d2h_bytes);
  if (rc != LGR_OK) return rc;
  LGR_CUDA(c, cudaStreamSynchronize(c->stream));
  return download_finish(c, out, d2h_bytes);
}

int lgr_upload(lgr_ctx* c, const lgr_batch_in* in) {
  if (!c) return LGR_E_ARG;
  int rc = upload_impl(c, in, nullptr);
  if (rc == LGR_OK) LGR_CUDA(c, cudaStreamSynchronize(c->stream));
  return rc;
}

int lgr_run_resident(lgr_ctx* c, lgr_stats* stats) {
  if (!c) return LGR_E_ARG;
  if (stats) std::memset(stats, 0, sizeof(*stats));
  return run_impl(c, stats);
}

int lgr_download(lgr_ctx* c, lgr_batch_out* out) {
  if (!c) return LGR_E_ARG;
  return download_impl(c, out, nullptr);
}

int lgr_genotype_batch(lgr_ctx* c, const lgr_batch_in* in, lgr_batch_out* out, lgr_stats* stats) {
  if (!c || !in || !out) return LGR_E_ARG;
  lgr_stats st;
  std::memset(&st, 0, sizeof(st));
  int64_t h2d = 0, d2h = 0;
  cudaEventRecord(c->ev[5], c->stream);
  int rc = upload_impl(c, in, &h2d);
  if (rc != LGR_OK) return rc;
  cudaEventRecord(c->ev[6], c->stream);
  rc = run_impl(c, &st);
  if (rc != LGR_OK && rc != LGR_E_LIMIT && rc != LGR_E_CIGAR_OVERFLOW) return rc;
  const int run_rc = rc;
  cudaEventRecord(c->ev[7], c->stream);
  rc = download_impl(c, out, &d2h);
  cudaEventRecord(c->ev[8], c->stream);
  cudaStreamSynchronize(c->stream);
  float ms;
  cudaEventElapsedTime(&ms, c->ev[5], c->ev[6]); st.ms_h2d = ms;
  cudaEventElapsedTime(&ms, c->ev[7], c->ev[8]); st.ms_d2h = ms;
  st.h2d_bytes = h2d, st.d2h_bytes = d2h;
  if (stats) *stats = st;
  return run_rc != LGR_OK ? run_rc : rc;
}

int lgr_submit(lgr_ctx* c, const lgr_batch_in* in, lgr_batch_out* out, lgr_ticket* ticket) {
  if (!c || !in || !out || !ticket) return LGR_E_ARG;
  int t = -1;
  for (int i = 0; i < LGR_MAX_INFLIGHT && t < 0; ++i)
    if (!c->slot_busy[i]) t = i;
  if (t < 0) { c->err = "all submission slots are in flight; lgr_wait one first"; return LGR_E_BUSY; }
  if (!c->slot[t]) {
    int rc = lgr_create(c->device, &c->prm, &c->slot[t]);
    if (rc != LGR_OK) { c->err = g_create_err; return rc; }
  }
  lgr_ctx* ch = c->slot[t];
  int64_t h2d = 0, d2h = 0;
  cudaEventRecord(ch->ev[5], ch->stream);
  int rc = upload_impl(ch, in, &h2d);
  if (rc == LGR_OK) {
    cudaEventRecord(ch->ev[6], ch->stream);
    rc = run_launch(ch);
  }
  if (rc == LGR_OK) {
    cudaEventRecord(ch->ev[7], ch->stream);
    rc = download_launch(ch, out, &d2h);
    cudaEventRecord(ch->ev[8], ch->stream);
  }
  if (rc != LGR_OK) {
    cudaStreamSynchronize(ch->stream);  // leave the slot idle
    c->err = ch->err;
    return rc;
  }
  c->slot_busy[t] = true, c->slot_out[t] = out, c->slot_h2d[t] = h2d, c->slot_d2h[t] = d2h;
  *ticket = t;
  return LGR_OK;
}

int lgr_wait(lgr_ctx* c, lgr_ticket t, lgr_stats* stats) {
  if (!c || t < 0 || t >= LGR_MAX_INFLIGHT || !c->slot_busy[t]) {
    if (c) c->err = "lgr_wait: unknown ticket";
    return LGR_E_ARG;
  }
  lgr_ctx* ch = c->slot[t];
  c->slot_busy[t] = false;
  if (cudaStreamSynchronize(ch->stream) != cudaSuccess) {
    c->err = std::string("lgr_wait: ") + cudaGetErrorString(cudaGetLastError());
    return LGR_E_CUDA;
  }
  lgr_stats st;
  std::memset(&st, 0, sizeof(st));
  const int run_rc = run_finish(ch, &st);
  int rc = LGR_OK;
  int64_t d2h = c->slot_d2h[t];
  if (run_rc == LGR_OK || run_rc == LGR_E_LIMIT || run_rc == LGR_E_CIGAR_OVERFLOW) rc = download_finish(ch, c->slot_out[t], &d2h);
  float ms;
  cudaEventElapsedTime(&ms, ch->ev[5], ch->ev[6]); st.ms_h2d = ms;
  cudaEventElapsedTime(&ms, ch->ev[7], ch->ev[8]); st.ms_d2h = ms;
  st.h2d_bytes = c->slot_h2d[t], st.d2h_bytes = d2h;
  if (stats) *stats = st;
  if (run_rc != LGR_OK || rc != LGR_OK) c->err = ch->err;
  return run_rc != LGR_OK ? run_rc : rc;
}

int lgr_hap_mid_occ(lgr_ctx* c, const uint8_t* hap, int32_t hap_len, int32_t* mid_occ) {
  if (!c || !hap || hap_len < 0 || !mid_occ) return LGR_E_ARG;
  if (hap_len > LGR_MAX_HAP_LEN) { c->err = "haplotype longer than LGR_MAX_HAP_LEN"; return LGR_E_LIMIT; }
  // a one-haplotype, zero-read batch through the index kernels
  const int32_t gb[2] = {0, 1}, zero2[2] = {0, 0};
  const int64_t ho[2] = {0, hap_len}, ro[1] = {0}, vo[1] = {0};
  lgr_batch_in in;
  std::memset(&in, 0, sizeof(in));
  in.n_groups = 1, in.n_haps = 1, in.n_reads = 0, in.n_vars = 0;
  in.grp_hap_begin = gb, in.grp_read_begin = zero2, in.grp_var_begin = zero2;
  in.hap_off = ho, in.hap_bases = hap, in.read_off = ro, in.var_hap_off = vo;
  int rc = upload_impl(c, &in, nullptr);
  if (rc != LGR_OK) return rc;
  Dev& D = c->D;
  cudaStream_t s = c->stream;
  LGR_CUDA(c, cudaMemsetAsync(D.ctr, 0, sizeof(long long) * C_COUNT, s));
  k_encode<<<8, 256, 0, s>>>(D.hap_bases, D.hap_codes, hap_len);
  k_hap_sketch<<<1, 64, 0, s>>>(D);
  k_hap_sort<<<1, 128, 2048 * sizeof(uint64_t), s>>>(D, c->prm.mid_occ_frac, c->prm.min_mid_occ, c->prm.max_mid_occ);
  LGR_CUDA(c, cudaMemcpyAsync(mid_occ, D.hap_mid, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  LGR_CUDA(c, cudaStreamSynchronize(s));
  LGR_CUDA(c, cudaGetLastError());
  c->resident = false;
  return LGR_OK;
}

}  // extern "C"
