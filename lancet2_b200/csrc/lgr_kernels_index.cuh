// lgr_kernels_index.cuh — encode, haplotype sketch / sort / mid_occ, read sketch / filter kernels
// Part of the single translation unit lgr_gpu.cu (included there, in order); see that file's header.
#ifndef LANCET2_B200_LGR_KERNELS_INDEX_CUH_
#define LANCET2_B200_LGR_KERNELS_INDEX_CUH_

namespace {

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
  if (n > len) { n = len; flag_err(D, D.hap_grp[h], E_MZ_CAP); }
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
    if (n > len) { n = len; flag_err(D, D.hap_grp[h], E_MZ_CAP); }
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
  const int req = D.grp_mid_req[g];
  const int h0 = D.grp_hap_begin[g];
  D.grp_mid[g] = req > 0 ? req : (D.grp_hap_begin[g + 1] > h0 ? D.hap_mid[h0] : min_mid);
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
    if (n > len) { n = len; flag_err(D, D.read_grp[r], E_MZ_CAP); }
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

}  // namespace

#endif  // LANCET2_B200_LGR_KERNELS_INDEX_CUH_
