// TEST INFRASTRUCTURE — CPU oracle (see mm2_restate.hpp header: PARITY UNPINNED).
// Restates minimap2 v2.30 mm_map() for Lancet2's Genotyper option set.
// Each function names the upstream minimap2 file/function it restates and the
// reference call site that reaches it.
#include "mm2_restate.hpp"

#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <cstring>

namespace mm2r {

// ---------------------------------------------------------------------------
// sketch.c
// ---------------------------------------------------------------------------
uint8_t Nt4(uint8_t c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3;
    default: return 4;
  }
}

// sketch.c: hash64 (invertible integer hash restricted to 2k bits)
static inline uint64_t Hash64Mask(uint64_t key, uint64_t mask) {
  key = (~key + (key << 21)) & mask;
  key = key ^ key >> 24;
  key = ((key + (key << 3)) + (key << 8)) & mask;
  key = key ^ key >> 14;
  key = ((key + (key << 2)) + (key << 4)) & mask;
  key = key ^ key >> 28;
  key = (key + (key << 31)) & mask;
  return key;
}

// sketch.c: mm_sketch, non-HPC branch.  Reached from mm_idx_str
// (genotyper.cpp:250) and collect_minimizers inside mm_map (genotyper.cpp:387).
void Sketch(const uint8_t* str, int len, int w, int k, uint32_t rid, std::vector<U128>& out) {
  assert(len > 0 && w > 0 && w < 256 && k > 0 && k <= 28);
  const uint64_t shift1 = 2 * (k - 1), mask = (1ULL << 2 * k) - 1;
  uint64_t kmer[2] = {0, 0};
  U128 buf[256];
  U128 mn = {UINT64_MAX, UINT64_MAX};
  for (int j = 0; j < w; ++j) buf[j] = {UINT64_MAX, UINT64_MAX};
  int l = 0, buf_pos = 0, min_pos = 0, kmer_span = 0;
  for (int i = 0; i < len; ++i) {
    const int c = Nt4(str[i]);
    U128 info = {UINT64_MAX, UINT64_MAX};
    if (c < 4) {
      kmer_span = l + 1 < k ? l + 1 : k;
      kmer[0] = (kmer[0] << 2 | (uint64_t)c) & mask;
      kmer[1] = (kmer[1] >> 2) | (3ULL ^ (uint64_t)c) << shift1;
      if (kmer[0] == kmer[1]) continue;  // strand-ambiguous k-mer: skipped entirely
      const int z = kmer[0] < kmer[1] ? 0 : 1;
      ++l;
      if (l >= k && kmer_span < 256) {
        info.x = Hash64Mask(kmer[z], mask) << 8 | (uint64_t)kmer_span;
        info.y = (uint64_t)rid << 32 | (uint32_t)i << 1 | (uint64_t)z;
      }
    } else {
      l = 0;
      kmer_span = 0;
    }
    buf[buf_pos] = info;
    if (l == w + k - 1 && mn.x != UINT64_MAX) {  // first full window: flush ties of the minimum
      for (int j = buf_pos + 1; j < w; ++j)
        if (mn.x == buf[j].x && buf[j].y != mn.y) out.push_back(buf[j]);
      for (int j = 0; j < buf_pos; ++j)
        if (mn.x == buf[j].x && buf[j].y != mn.y) out.push_back(buf[j]);
    }
    if (info.x <= mn.x) {  // new minimum (ties replace: rightmost wins)
      if (l >= w + k && mn.x != UINT64_MAX) out.push_back(mn);
      mn = info;
      min_pos = buf_pos;
    } else if (buf_pos == min_pos) {  // the minimum left the window
      if (l >= w + k - 1 && mn.x != UINT64_MAX) out.push_back(mn);
      mn.x = UINT64_MAX;
      for (int j = buf_pos + 1; j < w; ++j)
        if (mn.x >= buf[j].x) mn = buf[j], min_pos = j;
      for (int j = 0; j <= buf_pos; ++j)
        if (mn.x >= buf[j].x) mn = buf[j], min_pos = j;
      if (l >= w + k - 1 && mn.x != UINT64_MAX) {
        for (int j = buf_pos + 1; j < w; ++j)
          if (mn.x == buf[j].x && mn.y != buf[j].y) out.push_back(buf[j]);
        for (int j = 0; j <= buf_pos; ++j)
          if (mn.x == buf[j].x && mn.y != buf[j].y) out.push_back(buf[j]);
      }
    }
    if (++buf_pos == w) buf_pos = 0;
  }
  if (mn.x != UINT64_MAX) out.push_back(mn);
}

// ---------------------------------------------------------------------------
// ksort.h: KRADIX_SORT_INIT — in-place MSD radix sort, 8 bits per pass,
// insertion sort for <= 64 elements.  NOT stable for > 64 elements; minimap2's
// results depend on the exact permutation, so it is restated exactly.
// ---------------------------------------------------------------------------
namespace {
constexpr int kRsMinSize = 64;
constexpr int kRsMaxBits = 8;

template <typename T, typename KeyFn>
void RsInsertSort(T* beg, T* end, KeyFn key) {
  for (T* i = beg + 1; i < end; ++i) {
    if (key(*i) < key(*(i - 1))) {
      T tmp = *i;
      T* j;
      for (j = i; j > beg && key(tmp) < key(*(j - 1)); --j) *j = *(j - 1);
      *j = tmp;
    }
  }
}

template <typename T, typename KeyFn>
void RsSort(T* beg, T* end, int n_bits, int s, KeyFn key) {
  struct Bucket {
    T *b, *e;
  };
  const int size = 1 << n_bits, m = size - 1;
  Bucket b[1 << kRsMaxBits];
  Bucket* be = b + size;
  for (Bucket* k = b; k != be; ++k) k->b = k->e = beg;
  for (T* i = beg; i != end; ++i) ++b[key(*i) >> s & m].e;
  for (Bucket* k = b + 1; k != be; ++k) {
    k->e += (k - 1)->e - beg;
    k->b = (k - 1)->e;
  }
  for (Bucket* k = b; k != be;) {
    if (k->b != k->e) {
      Bucket* l = b + (key(*k->b) >> s & m);
      if (l != k) {
        T tmp = *k->b, swap;
        do {
          swap = tmp;
          tmp = *l->b;
          *l->b++ = swap;
          l = b + (key(tmp) >> s & m);
        } while (l != k);
        *k->b++ = tmp;
      } else {
        ++k->b;
      }
    } else {
      ++k;
    }
  }
  b->b = beg;
  for (Bucket* k = b + 1; k != be; ++k) k->b = (k - 1)->e;
  if (s) {
    s = s > n_bits ? s - n_bits : 0;
    for (Bucket* k = b; k != be; ++k) {
      if (k->e - k->b > kRsMinSize) RsSort(k->b, k->e, n_bits, s, key);
      else if (k->e - k->b > 1) RsInsertSort(k->b, k->e, key);
    }
  }
}

template <typename T, typename KeyFn>
void RadixSort(T* beg, T* end, KeyFn key) {
  if (end - beg <= kRsMinSize) RsInsertSort(beg, end, key);
  else RsSort(beg, end, kRsMaxBits, (8 - 1) * kRsMaxBits, key);
}
}  // namespace

void RadixSort128x(U128* beg, U128* end) {
  RadixSort(beg, end, [](const U128& a) { return a.x; });
}
void RadixSort64(uint64_t* beg, uint64_t* end) {
  RadixSort(beg, end, [](const uint64_t& a) { return a; });
}

// khash.h
uint32_t X31HashString(const char* s) {
  uint32_t h = (uint32_t)(int32_t)(signed char)*s;
  if (h)
    for (++s; *s; ++s) h = (h << 5) - h + (uint32_t)(int32_t)(signed char)*s;
  return h;
}
uint32_t WangHash(uint32_t key) {
  key += ~(key << 15);
  key ^= (key >> 10);
  key += (key << 3);
  key ^= (key >> 6);
  key += ~(key << 11);
  key ^= (key >> 16);
  return key;
}

// ---------------------------------------------------------------------------
// index.c: mm_idx_str / mm_idx_add / worker_post / mm_idx_get / mm_idx_cal_max_occ
// The upstream bucketed khash is replaced by one sorted (key, y) table: mm_idx_get
// returns, for n > 1, the occurrences sorted by y (worker_post radix_sort_64), which
// a (key, y) sort reproduces; for n == 1 the single y.
// ---------------------------------------------------------------------------
void BuildHapIndex(const uint8_t* hap, int len, int w, int k, HapIndex& idx) {
  idx.k = k;
  idx.w = w;
  idx.seq4.resize(len);
  for (int i = 0; i < len; ++i) idx.seq4[i] = Nt4(hap[i]);
  idx.keys.clear();
  idx.vals.clear();
  if (len <= 0) return;
  std::vector<U128> a;
  Sketch(hap, len, w, k, 0, a);
  std::vector<std::pair<uint64_t, uint64_t>> kv(a.size());
  for (size_t i = 0; i < a.size(); ++i) kv[i] = {a[i].x >> 8, a[i].y};
  std::sort(kv.begin(), kv.end());
  idx.keys.resize(kv.size());
  idx.vals.resize(kv.size());
  for (size_t i = 0; i < kv.size(); ++i) idx.keys[i] = kv[i].first, idx.vals[i] = kv[i].second;
}

const uint64_t* HapIndex::Get(uint64_t minier, int* n) const {
  auto lo = std::lower_bound(keys.begin(), keys.end(), minier);
  auto hi = std::upper_bound(lo, keys.end(), minier);
  *n = (int)(hi - lo);
  if (*n == 0) return nullptr;
  return vals.data() + (lo - keys.begin());
}

int32_t HapIndex::CalMaxOcc(float f) const {
  if (f <= 0.) return INT32_MAX;
  std::vector<uint32_t> occ;
  for (size_t i = 0; i < keys.size();) {
    size_t j = i;
    while (j < keys.size() && keys[j] == keys[i]) ++j;
    occ.push_back((uint32_t)(j - i));
    i = j;
  }
  const size_t n = occ.size();
  if (n == 0) return INT32_MAX;
  const uint32_t kk = (uint32_t)((1. - f) * n);  // ks_ksmall index (0-based k-th smallest)
  std::nth_element(occ.begin(), occ.begin() + kk, occ.end());
  return (int32_t)(occ[kk] + 1);
}

// options.c: mm_mapopt_update — only the mid_occ rule matters here
// (genotyper.cpp:263-266; SURVEY Appendix A.2: latched from the first index).
int32_t MidOccFromIndex(const HapIndex& idx, const lgr_params& p) {
  int32_t mid = idx.CalMaxOcc(p.mid_occ_frac);
  if (mid < p.min_mid_occ) mid = p.min_mid_occ;
  if (p.max_mid_occ > p.min_mid_occ && mid > p.max_mid_occ) mid = p.max_mid_occ;
  return mid;
}

// ---------------------------------------------------------------------------
// seed.c
// ---------------------------------------------------------------------------
namespace {

constexpr uint64_t kSeedTandem = 1ULL << 42;

struct Seed {
  uint32_t n;
  uint32_t q_pos;
  uint32_t q_span : 31, flt : 1;
  uint32_t seg_id : 31, is_tandem : 1;
  const uint64_t* cr;
};

// seed.c: mm_seed_mz_flt
void SeedMzFlt(std::vector<U128>& mv, int32_t q_occ_max, float q_occ_frac) {
  if ((int64_t)mv.size() <= q_occ_max || q_occ_frac <= 0.0f || q_occ_max <= 0) return;
  std::vector<U128> a(mv.size());
  for (size_t i = 0; i < mv.size(); ++i) a[i].x = mv[i].x, a[i].y = i;
  RadixSort128x(a.data(), a.data() + a.size());
  const size_t n = mv.size();
  for (size_t st = 0, i = 1; i <= n; ++i) {
    if (i == n || a[i].x != a[st].x) {
      const int32_t cnt = (int32_t)(i - st);
      if (cnt > q_occ_max && cnt > n * q_occ_frac)
        for (size_t j = st; j < i; ++j) mv[a[j].y].x = 0;
      st = i;
    }
  }
  size_t j = 0;
  for (size_t i = 0; i < n; ++i)
    if (mv[i].x != 0) mv[j++] = mv[i];
  mv.resize(j);
}

// seed.c: mm_seed_select — per streak of high-occurrence seeds keep the
// max_high_occ least frequent ones.  The upstream binary max-heap is a device
// for "replace the current maximum (n, then index) when a strictly less
// frequent seed arrives"; restated with an explicit max search.
void SeedSelect(int32_t n, Seed* a, int len, int max_occ, int max_max_occ, int dist) {
  constexpr int kMaxMaxHighOcc = 128;
  if (n == 0 || n == 1) return;
  int32_t m = 0;
  for (int32_t i = 0; i < n; ++i)
    if ((int32_t)a[i].n > max_occ) ++m;
  if (m == 0) return;
  int32_t last0 = -1;
  for (int32_t i = 0; i <= n; ++i) {
    if (i == n || (int32_t)a[i].n <= max_occ) {
      if (i - last0 > 1) {
        const int32_t ps = last0 < 0 ? 0 : (int32_t)(a[last0].q_pos >> 1);
        const int32_t pe = i == n ? len : (int32_t)(a[i].q_pos >> 1);
        const int32_t st = last0 + 1, en = i;
        int32_t max_high_occ = (int32_t)((double)(pe - ps) / dist + .499);
        if (max_high_occ > 0) {
          if (max_high_occ > kMaxMaxHighOcc) max_high_occ = kMaxMaxHighOcc;
          std::vector<uint64_t> b;
          int32_t j = st;
          for (; j < en && (int32_t)b.size() < max_high_occ; ++j)
            b.push_back((uint64_t)a[j].n << 32 | (uint32_t)j);
          for (; j < en; ++j) {
            auto top = std::max_element(b.begin(), b.end());
            if ((int32_t)a[j].n < (int32_t)(*top >> 32)) *top = (uint64_t)a[j].n << 32 | (uint32_t)j;
          }
          for (uint64_t v : b) a[(uint32_t)v].flt = 1;
        }
        for (int32_t j = st; j < en; ++j) a[j].flt ^= 1;
        for (int32_t j = st; j < en; ++j)
          if ((int32_t)a[j].n > max_max_occ) a[j].flt = 1;
      }
      last0 = i;
    }
  }
}

// ---------------------------------------------------------------------------
// lchain.c
// ---------------------------------------------------------------------------
// mmpriv.h: mg_log2 — bit-trick log2 approximation.  The reference builds
// minimap2 with "-O3 -march=x86-64-v3" (cmake/build_minimap2.sh:22,
// cmake/compiler_flags.cmake:151), where both clang (-ffp-contract=on) and gcc
// (-ffp-contract=fast) contract a*b+c into FMA; the two multiply-adds below are
// therefore written as explicit fmaf so that every build of the oracle and the
// CUDA kernel (__fmaf_rn) agree bit for bit.
inline float MgLog2(float x) {
  union {
    float f;
    uint32_t i;
  } z = {x};
  float log_2 = (float)(int)(((z.i >> 23) & 255) - 128);
  z.i &= ~(255u << 23);
  z.i += 127u << 23;
  const float t = std::fmaf(-0.34484843f, z.f, 2.02466578f);
  log_2 += std::fmaf(t, z.f, -0.67487759f);
  return log_2;
}

// lchain.c: comput_sc (n_seg == 1, !is_cdna)
inline int32_t ComputSc(const U128& ai, const U128& aj, int32_t max_dist_x, int32_t max_dist_y,
                        int32_t bw, float chn_pen_gap, float chn_pen_skip) {
  const int32_t dq = (int32_t)ai.y - (int32_t)aj.y;
  if (dq <= 0 || dq > max_dist_x) return INT32_MIN;
  const int32_t dr = (int32_t)(ai.x - aj.x);
  if (dr == 0 || dq > max_dist_y) return INT32_MIN;
  const int32_t dd = dr > dq ? dr - dq : dq - dr;
  if (dd > bw) return INT32_MIN;
  const int32_t dg = dr < dq ? dr : dq;
  const int32_t q_span = (int32_t)(aj.y >> 32 & 0xff);
  int32_t sc = q_span < dg ? q_span : dg;
  if (dd || dg > q_span) {
    // lin_pen = pen_gap*dd + pen_skip*dg : with pen_skip*dg rounded separately or
    // fused the result is identical whenever pen_skip == 0 (the reference's
    // setting); written unfused.
    const float lin_pen = chn_pen_gap * (float)dd + chn_pen_skip * (float)dg;
    const float log_pen = dd >= 1 ? MgLog2((float)(dd + 1)) : 0.0f;
    sc -= (int)(lin_pen + .5f * log_pen);
  }
  return sc;
}

// lchain.c: mg_chain_bk_end
int64_t ChainBkEnd(int32_t max_drop, const U128* z, const int32_t* f, const int64_t* p, int32_t* t,
                   int64_t k) {
  int64_t i = (int64_t)z[k].y, end_i = -1, max_i = i;
  int32_t max_s = 0;
  if (i < 0 || t[i] != 0) return i;
  do {
    t[i] = 2;
    end_i = i = p[i];
    const int32_t s = i < 0 ? (int32_t)z[k].x : (int32_t)z[k].x - f[i];
    if (s > max_s) max_s = s, max_i = i;
    else if (max_s - s > max_drop) break;
  } while (i >= 0 && t[i] == 0);
  for (i = (int64_t)z[k].y; i >= 0 && i != end_i; i = p[i]) t[i] = 0;
  return max_i;
}

// lchain.c: mg_chain_backtrack
void ChainBacktrack(int64_t n, const int32_t* f, const int64_t* p, std::vector<int32_t>& v,
                    std::vector<int32_t>& t, int32_t min_cnt, int32_t min_sc, int32_t max_drop,
                    std::vector<uint64_t>& u, int32_t* n_v_) {
  u.clear();
  *n_v_ = 0;
  std::vector<U128> z;
  for (int64_t i = 0; i < n; ++i)
    if (f[i] >= min_sc) z.push_back({(uint64_t)f[i], (uint64_t)i});
  const int64_t n_z = (int64_t)z.size();
  if (n_z == 0) return;
  RadixSort128x(z.data(), z.data() + n_z);
  std::fill(t.begin(), t.end(), 0);
  int64_t n_v = 0;
  for (int64_t k = n_z - 1; k >= 0; --k) {
    if (t[z[k].y] == 0) {
      const int64_t n_v0 = n_v;
      const int64_t end_i = ChainBkEnd(max_drop, z.data(), f, p, t.data(), k);
      int64_t i;
      for (i = (int64_t)z[k].y; i != end_i; i = p[i]) v[n_v++] = (int32_t)i, t[i] = 1;
      const int32_t sc = i < 0 ? (int32_t)z[k].x : (int32_t)z[k].x - f[i];
      if (sc >= min_sc && n_v > n_v0 && n_v - n_v0 >= min_cnt)
        u.push_back((uint64_t)sc << 32 | (uint64_t)(n_v - n_v0));
      else
        n_v = n_v0;
    }
  }
  *n_v_ = (int32_t)n_v;
}

// lchain.c: compact_a
std::vector<U128> CompactA(std::vector<uint64_t>& u, int32_t n_v, const std::vector<int32_t>& v,
                           const std::vector<U128>& a) {
  const int32_t n_u = (int32_t)u.size();
  std::vector<U128> b(n_v);
  int64_t k = 0;
  for (int32_t i = 0; i < n_u; ++i) {
    const int32_t k0 = (int32_t)k, ni = (int32_t)u[i];
    for (int32_t j = 0; j < ni; ++j) b[k++] = a[v[k0 + (ni - j - 1)]];
  }
  std::vector<U128> w(n_u);
  k = 0;
  for (int32_t i = 0; i < n_u; ++i) {
    w[i].x = b[k].x;
    w[i].y = (uint64_t)k << 32 | (uint32_t)i;
    k += (int32_t)u[i];
  }
  RadixSort128x(w.data(), w.data() + n_u);
  std::vector<uint64_t> u2(n_u);
  std::vector<U128> out(n_v);
  k = 0;
  for (int32_t i = 0; i < n_u; ++i) {
    const int32_t j = (int32_t)w[i].y, n = (int32_t)u[j];
    u2[i] = u[j];
    std::memcpy(&out[k], &b[w[i].y >> 32], (size_t)n * sizeof(U128));
    k += n;
  }
  u = u2;
  return out;
}

// lchain.c: mg_lchain_dp (is_cdna = 0, n_seg = 1).  Returns the compacted anchors.
std::vector<U128> LchainDp(int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter,
                           int min_cnt, int min_sc, float chn_pen_gap, float chn_pen_skip,
                           const std::vector<U128>& a, std::vector<uint64_t>& u, MapDebug* dbg) {
  u.clear();
  const int64_t n = (int64_t)a.size();
  if (n == 0) return {};
  const int32_t max_drop = bw;
  if (max_dist_x < bw) max_dist_x = bw;
  if (max_dist_y < bw) max_dist_y = bw;
  std::vector<int64_t> p(n);
  std::vector<int32_t> f(n), v(n), t(n, 0);
  int64_t st = 0, max_ii = -1, n_iter = 0;
  for (int64_t i = 0; i < n; ++i) {
    int64_t max_j = -1, end_j;
    int32_t max_f = (int32_t)(a[i].y >> 32 & 0xff), n_skip = 0;
    while (st < i && (a[i].x >> 32 != a[st].x >> 32 || a[i].x > a[st].x + (uint64_t)max_dist_x)) ++st;
    if (i - st > max_iter) st = i - max_iter;
    int64_t j;
    for (j = i - 1; j >= st; --j) {
      int32_t sc = ComputSc(a[i], a[j], max_dist_x, max_dist_y, bw, chn_pen_gap, chn_pen_skip);
      ++n_iter;
      if (sc == INT32_MIN) continue;
      sc += f[j];
      if (sc > max_f) {
        max_f = sc, max_j = j;
        if (n_skip > 0) --n_skip;
      } else if (t[j] == (int32_t)i) {
        if (++n_skip > max_skip) break;
      }
      if (p[j] >= 0) t[p[j]] = (int32_t)i;
    }
    end_j = j;
    if (max_ii < 0 || a[i].x - a[max_ii].x > (uint64_t)(int64_t)max_dist_x) {  // unsigned compare, as upstream
      int32_t mx = INT32_MIN;
      max_ii = -1;
      for (j = i - 1; j >= st; --j)
        if (mx < f[j]) mx = f[j], max_ii = j;
    }
    if (max_ii >= 0 && max_ii < end_j) {
      const int32_t tmp =
          ComputSc(a[i], a[max_ii], max_dist_x, max_dist_y, bw, chn_pen_gap, chn_pen_skip);
      if (tmp != INT32_MIN && max_f < tmp + f[max_ii]) max_f = tmp + f[max_ii], max_j = max_ii;
    }
    f[i] = max_f, p[i] = max_j;
    v[i] = max_j >= 0 && v[max_j] > max_f ? v[max_j] : max_f;
    if (max_ii < 0 ||
        (a[i].x - a[max_ii].x <= (uint64_t)(int64_t)max_dist_x && f[max_ii] < f[i]))
      max_ii = i;
  }
  if (dbg) dbg->f = f, dbg->p = p, dbg->chain_evals = n_iter;
  int32_t n_v = 0;
  ChainBacktrack(n, f.data(), p.data(), v, t, min_cnt, min_sc, max_drop, u, &n_v);
  if (u.empty()) return {};
  return CompactA(u, n_v, v, a);
}

// ---------------------------------------------------------------------------
// hit.c
// ---------------------------------------------------------------------------
inline uint64_t Hash64(uint64_t key) {
  key = ~key + (key << 21);
  key = key ^ key >> 24;
  key = ((key + (key << 3)) + (key << 8));
  key = key ^ key >> 14;
  key = ((key + (key << 2)) + (key << 4));
  key = key ^ key >> 28;
  key = key + (key << 31);
  return key;
}

// hit.c: mm_reg_set_coor
void RegSetCoor(Reg& r, int32_t qlen, const std::vector<U128>& a) {
  const int32_t k = r.as, q_span = (int32_t)(a[k].y >> 32 & 0xff);
  r.rev = a[k].x >> 63;
  r.rs = (int32_t)a[k].x + 1 > q_span ? (int32_t)a[k].x + 1 - q_span : 0;
  r.re = (int32_t)a[k + r.cnt - 1].x + 1;
  if (!r.rev) {
    r.qs = (int32_t)a[k].y + 1 - q_span;
    r.qe = (int32_t)a[k + r.cnt - 1].y + 1;
  } else {
    r.qs = qlen - ((int32_t)a[k + r.cnt - 1].y + 1);
    r.qe = qlen - ((int32_t)a[k].y + 1 - q_span);
  }
}

// hit.c: mm_gen_regs
std::vector<Reg> GenRegs(uint32_t hash, int qlen, const std::vector<uint64_t>& u,
                         const std::vector<U128>& a) {
  const int n_u = (int)u.size();
  if (n_u == 0) return {};
  std::vector<U128> z(n_u);
  int k = 0;
  for (int i = 0; i < n_u; ++i) {
    const uint32_t h = (uint32_t)Hash64((Hash64(a[k].x) + Hash64(a[k].y)) ^ hash);
    z[i].x = u[i] ^ h;
    z[i].y = (uint64_t)k << 32 | (uint32_t)(int32_t)u[i];
    k += (int32_t)u[i];
  }
  RadixSort128x(z.data(), z.data() + n_u);
  std::reverse(z.begin(), z.end());
  std::vector<Reg> r(n_u);
  for (int i = 0; i < n_u; ++i) {
    Reg& ri = r[i];
    ri.id = i;
    ri.parent = -1;  // MM_PARENT_UNSET
    ri.score = ri.score0 = (int32_t)(z[i].x >> 32);
    ri.hash = (uint32_t)z[i].x;
    ri.cnt = (int32_t)z[i].y;
    ri.as = (int32_t)(z[i].y >> 32);
    RegSetCoor(ri, qlen, a);
  }
  return r;
}

// hit.c: mm_set_parent (hard_mask_level = 0, no ALT contigs)
void SetParent(float mask_level, int mask_len, std::vector<Reg>& r, int sub_diff) {
  const int n = (int)r.size();
  if (n <= 0) return;
  for (int i = 0; i < n; ++i) r[i].id = i;
  std::vector<uint64_t> cov(n);
  std::vector<int> w(n);
  w[0] = 0, r[0].parent = 0;
  int k = 1;
  for (int i = 1; i < n; ++i) {
    Reg& ri = r[i];
    const int si = ri.qs, ei = ri.qe;
    int n_cov = 0, uncov_len = 0;
    for (int j = 0; j < k; ++j) {
      const Reg& rp = r[w[j]];
      int sj = rp.qs, ej = rp.qe;
      if (ej <= si || sj >= ei) continue;
      if (sj < si) sj = si;
      if (ej > ei) ej = ei;
      cov[n_cov++] = (uint64_t)sj << 32 | (uint32_t)ej;
    }
    int j = k;
    if (n_cov > 0) {
      int x = si;
      std::sort(cov.begin(), cov.begin() + n_cov);
      for (int c = 0; c < n_cov; ++c) {
        if ((int)(cov[c] >> 32) > x) uncov_len += (int)(cov[c] >> 32) - x;
        x = (int32_t)cov[c] > x ? (int32_t)cov[c] : x;
      }
      if (ei > x) uncov_len += ei - x;
      for (j = 0; j < k; ++j) {
        Reg& rp = r[w[j]];
        const int sj = rp.qs, ej = rp.qe;
        if (ej <= si || sj >= ei) continue;
        const int mn = ej - sj < ei - si ? ej - sj : ei - si;
        const int mx = ej - sj > ei - si ? ej - sj : ei - si;
        const int ol = si < sj ? (ei < sj ? 0 : ei < ej ? ei - sj : ej - sj)
                               : (ej < si ? 0 : ej < ei ? ej - si : ei - si);
        if ((float)ol / mn - (float)uncov_len / mx > mask_level && uncov_len <= mask_len) {
          int cnt_sub = 0;
          int sci = ri.score;
          ri.parent = rp.parent;
          rp.subsc = rp.subsc > sci ? rp.subsc : sci;
          if (ri.cnt >= rp.cnt) cnt_sub = 1;
          if (rp.has_p && ri.has_p && (rp.rs != ri.rs || rp.re != ri.re || ol != mn)) {
            sci = ri.dp_max;
            rp.dp_max2 = rp.dp_max2 > sci ? rp.dp_max2 : sci;
            if (rp.dp_max - ri.dp_max <= sub_diff) cnt_sub = 1;
          }
          if (cnt_sub) ++rp.n_sub;
          break;
        }
      }
    }
    if (j == k) w[k++] = i, ri.parent = i, ri.n_sub = 0;
  }
}

// hit.c: mm_sync_regs (id/parent bookkeeping after hits were removed)
void SyncRegs(std::vector<Reg>& regs) {
  const int n = (int)regs.size();
  if (n <= 0) return;
  int max_id = -1;
  for (int i = 0; i < n; ++i) max_id = std::max(max_id, regs[i].id);
  std::vector<int> tmp(max_id + 1, -1);
  for (int i = 0; i < n; ++i)
    if (regs[i].id >= 0) tmp[regs[i].id] = i;
  for (int i = 0; i < n; ++i) {
    Reg& r = regs[i];
    r.id = i;
    if (r.parent >= 0 && tmp[r.parent] >= 0) r.parent = tmp[r.parent];
    else r.parent = -1;
  }
}

// hit.c: mm_select_sub
void SelectSub(float pri_ratio, int min_diff, int best_n, int check_strand, int min_strand_sc,
               std::vector<Reg>& r) {
  if (!(pri_ratio > 0.0f) || r.empty()) return;
  const int n = (int)r.size();
  int k = 0, n_2nd = 0;
  for (int i = 0; i < n; ++i) {
    const int p = r[i].parent;
    if (p == i) {
      r[k++] = r[i];
    } else if ((r[i].score >= r[p].score * pri_ratio || r[i].score + min_diff >= r[p].score) &&
               n_2nd < best_n) {
      if (!(r[i].qs == r[p].qs && r[i].qe == r[p].qe && r[i].rs == r[p].rs && r[i].re == r[p].re))
        r[k++] = r[i], ++n_2nd;
    } else if (check_strand && n_2nd < best_n && r[i].score > min_strand_sc &&
               r[p].rev != r[i].rev) {
      r[i].strand_retained = true;
      r[k++] = r[i], ++n_2nd;
    }
  }
  if (k != n) {
    r.resize(k);
    SyncRegs(r);
  }
}

// ---------------------------------------------------------------------------
// align.c
// ---------------------------------------------------------------------------
// ksw2.h: ksw_push_cigar
inline void PushCigar(std::vector<uint32_t>& c, uint32_t op, int len) {
  if (c.empty() || op != (c.back() & 0xf)) c.push_back((uint32_t)len << 4 | op);
  else c.back() += (uint32_t)len << 4;
}

// align.c: mm_append_cigar
void AppendCigar(Reg& r, const std::vector<uint32_t>& c) {
  if (c.empty()) return;
  r.has_p = true;
  size_t st = 0;
  if (!r.cigar.empty() && (r.cigar.back() & 0xf) == (c[0] & 0xf)) {
    r.cigar.back() += (c[0] >> 4) << 4;
    st = 1;
  }
  r.cigar.insert(r.cigar.end(), c.begin() + st, c.end());
}

// align.c: mm_max_stretch
void MaxStretch(const Reg& r, const std::vector<U128>& a, int32_t* as, int32_t* cnt) {
  *as = r.as, *cnt = r.cnt;
  if (r.cnt < 2) return;
  int32_t max_score = -1, max_i = -1, max_len = 0;
  int32_t score = (int32_t)(a[r.as].y >> 32 & 0xff), len = 1;
  int32_t i;
  for (i = r.as; i < r.as + r.cnt - 1; ++i) {
    const int32_t q_span = (int32_t)(a[i + 1].y >> 32 & 0xff);
    const int32_t lr = (int32_t)a[i + 1].x - (int32_t)a[i].x;
    const int32_t lq = (int32_t)a[i + 1].y - (int32_t)a[i].y;
    if (lq == lr) {
      score += lq < q_span ? lq : q_span;
      ++len;
    } else {
      if (score > max_score) max_score = score, max_len = len, max_i = i - len + 1;
      score = q_span;
      len = 1;
    }
  }
  if (score > max_score) max_score = score, max_len = len, max_i = i - len + 1;
  *as = max_i, *cnt = max_len;
}

// align.c: mm_fix_cigar
void FixCigar(Reg& r, const uint8_t* qseq, const uint8_t* tseq, int* qshift, int* tshift) {
  std::vector<uint32_t>& c = r.cigar;
  int32_t toff = 0, qoff = 0;
  bool to_shrink = false;
  *qshift = *tshift = 0;
  if (c.size() <= 1) return;
  const uint32_t n = (uint32_t)c.size();
  for (uint32_t k = 0; k < n; ++k) {  // indel left alignment
    const uint32_t op = c[k] & 0xf, len = c[k] >> 4;
    if (len == 0) to_shrink = true;
    if (op == 0) {
      toff += len, qoff += len;
    } else if (op == 1 || op == 2) {
      if (k > 0 && k < n - 1 && (c[k - 1] & 0xf) == 0 && (c[k + 1] & 0xf) == 0) {
        int l;
        const int prev_len = (int)(c[k - 1] >> 4);
        if (op == 1) {
          for (l = 0; l < prev_len; ++l)
            if (qseq[qoff - 1 - l] != qseq[qoff + (int)len - 1 - l]) break;
        } else {
          for (l = 0; l < prev_len; ++l)
            if (tseq[toff - 1 - l] != tseq[toff + (int)len - 1 - l]) break;
        }
        if (l > 0) c[k - 1] -= (uint32_t)l << 4, c[k + 1] += (uint32_t)l << 4, qoff -= l, toff -= l;
        if (l == prev_len) to_shrink = true;
      }
      if (op == 1) qoff += len;
      else toff += len;
    } else if (op == 3) {
      toff += len;
    }
  }
  for (uint32_t k = 0; k + 2 < n; ++k) {  // runs like 5I6D7I → one I and one D
    if ((c[k] & 0xf) > 0 && (c[k] & 0xf) + (c[k + 1] & 0xf) == 3) {
      uint32_t l, s[3] = {0, 0, 0};
      for (l = k; l < n; ++l) {
        const uint32_t op = c[l] & 0xf;
        if (op == 1 || op == 2 || c[l] >> 4 == 0) s[op] += c[l] >> 4;
        else break;
      }
      if (s[1] > 0 && s[2] > 0 && l - k > 2) {
        c[k] = s[1] << 4 | 1;
        c[k + 1] = s[2] << 4 | 2;
        for (k += 2; k < l; ++k) c[k] &= 0xf;
        to_shrink = true;
      }
      k = l;
    }
  }
  if (to_shrink) {
    size_t l = 0;
    for (size_t k = 0; k < c.size(); ++k)
      if (c[k] >> 4 != 0) c[l++] = c[k];
    c.resize(l);
    l = 0;
    for (size_t k = 0; k < c.size(); ++k) {
      if (k == c.size() - 1 || (c[k] & 0xf) != (c[k + 1] & 0xf)) c[l++] = c[k];
      else c[k + 1] += c[k] >> 4 << 4;
    }
    c.resize(l);
  }
  if (!c.empty() && ((c[0] & 0xf) == 1 || (c[0] & 0xf) == 2)) {  // strip a leading I or D
    const int32_t l = (int32_t)(c[0] >> 4);
    if ((c[0] & 0xf) == 1) {
      if (r.rev) r.qe -= l;
      else r.qs += l;
      *qshift = l;
    } else {
      r.rs += l, *tshift = l;
    }
    c.erase(c.begin());
  }
}

// align.c: mm_update_extra (is_eqx = 0, log_gap = 0 because MM_F_SR is set)
void UpdateExtra(Reg& r, const uint8_t* qseq, const uint8_t* tseq, const int8_t* mat, int8_t q,
                 int8_t e) {
  if (!r.has_p) return;
  int qshift, tshift;
  int32_t toff = 0, qoff = 0;
  double s = 0.0, mx = 0.0;
  FixCigar(r, qseq, tseq, &qshift, &tshift);
  qseq += qshift, tseq += tshift;
  r.blen = r.mlen = 0;
  for (uint32_t c : r.cigar) {
    const uint32_t op = c & 0xf, len = c >> 4;
    if (op == 0) {
      int n_ambi = 0, n_diff = 0;
      for (uint32_t l = 0; l < len; ++l) {
        const int cq = qseq[qoff + l], ct = tseq[toff + l];
        if (ct > 3 || cq > 3) ++n_ambi;
        else if (ct != cq) ++n_diff;
        s += mat[ct * 5 + cq];
        if (s < 0) s = 0;
        else mx = mx > s ? mx : s;
      }
      r.blen += len - n_ambi, r.mlen += len - (n_ambi + n_diff), r.n_ambi += n_ambi;
      toff += len, qoff += len;
    } else if (op == 1) {
      int n_ambi = 0;
      for (uint32_t l = 0; l < len; ++l)
        if (qseq[qoff + l] > 3) ++n_ambi;
      r.blen += len - n_ambi, r.n_ambi += n_ambi;
      s -= q + e;  // non-log gap cost branch (SR)
      if (s < 0) s = 0;
      qoff += len;
    } else if (op == 2) {
      int n_ambi = 0;
      for (uint32_t l = 0; l < len; ++l)
        if (tseq[toff + l] > 3) ++n_ambi;
      r.blen += len - n_ambi, r.n_ambi += n_ambi;
      s -= q + e;
      if (s < 0) s = 0;
      toff += len;
    } else if (op == 3) {
      toff += len;
    }
  }
  r.dp_max = (int32_t)(mx + .499);
}

// ksw2.h: ksw_gen_simple_mat
void GenSimpleMat(int m, int8_t* mat, int8_t a, int8_t b, int8_t sc_ambi) {
  a = a < 0 ? -a : a;
  b = b > 0 ? -b : b;
  sc_ambi = sc_ambi > 0 ? -sc_ambi : sc_ambi;
  for (int i = 0; i < m - 1; ++i) {
    for (int j = 0; j < m - 1; ++j) mat[i * m + j] = i == j ? a : b;
    mat[i * m + m - 1] = sc_ambi;
  }
  for (int j = 0; j < m; ++j) mat[(m - 1) * m + j] = sc_ambi;
}

// align.c: mm_align1, MM_F_SR branch only (MM_F_SR is always set, genotyper.cpp:109)
void Align1(const lgr_params& opt, const HapIndex& mi, int qlen, const uint8_t* const qseq0[2],
            Reg& r, const std::vector<U128>& a, MapDebug* dbg) {
  if (r.cnt == 0) return;
  const int32_t rev = (int32_t)(a[r.as].x >> 63);
  int8_t mat[25];
  GenSimpleMat(5, mat, (int8_t)opt.a, (int8_t)opt.b, (int8_t)opt.sc_ambi);
  const int32_t hap_len = (int32_t)mi.seq4.size();

  int32_t as1, cnt1;
  MaxStretch(r, a, &as1, &cnt1);
  int32_t rs = (int32_t)a[as1].x + 1 - (int32_t)(a[as1].y >> 32 & 0xff);
  int32_t qs = (int32_t)a[as1].y + 1 - (int32_t)(a[as1].y >> 32 & 0xff);
  int32_t re = (int32_t)a[as1 + cnt1 - 1].x + 1;
  int32_t qe = (int32_t)a[as1 + cnt1 - 1].y + 1;

  const int32_t qs0 = 0, qe0 = qlen;
  int32_t l = qs;
  l += l * opt.a + opt.end_bonus > opt.q ? (l * opt.a + opt.end_bonus - opt.q) / opt.e : 0;
  const int32_t rs0 = rs - l > 0 ? rs - l : 0;
  l = qlen - qe;
  l += l * opt.a + opt.end_bonus > opt.q ? (l * opt.a + opt.end_bonus - opt.q) / opt.e : 0;
  const int32_t re0 = re + l < hap_len ? re + l : hap_len;

  int32_t rs1, qs1, re1, qe1;
  const uint8_t* qseq_all = qseq0[rev];
  r.cigar.clear();
  r.has_p = false;
  r.dp_score = 0;
  r.n_ambi = 0;

  if (qs > 0 && rs > 0) {  // left extension on reversed sequences
    std::vector<uint8_t> qrev(qseq_all + qs0, qseq_all + qs), trev(mi.seq4.begin() + rs0,
                                                                  mi.seq4.begin() + rs);
    std::reverse(qrev.begin(), qrev.end());
    std::reverse(trev.begin(), trev.end());
    ExtzResult ez;
    ExtzOnly(qs - qs0, qrev.data(), rs - rs0, trev.data(), mat, opt.q, opt.e, opt.end_bonus,
             kEzRight | kEzRevCigar, ez);
    if (dbg) dbg->dp_cells_full += (int64_t)(qs - qs0) * (rs - rs0);
    if (!ez.cigar.empty()) {
      AppendCigar(r, ez.cigar);
      r.dp_score += ez.max;
    }
    rs1 = rs - (ez.reach_end ? ez.mqe_t + 1 : ez.max_t + 1);
    qs1 = qs - (ez.reach_end ? qs - qs0 : ez.max_q + 1);
  } else {
    rs1 = rs, qs1 = qs;
  }
  re1 = rs, qe1 = qs;

  {  // SR "gap filling": one ungapped block over the longest co-linear stretch
    re = (int32_t)a[as1 + cnt1 - 1].x + 1;
    qe = (int32_t)a[as1 + cnt1 - 1].y + 1;
    re1 = re, qe1 = qe;
    assert(qe - qs == re - rs);
    const uint8_t* qseq = &qseq_all[qs];
    const uint8_t* tseq = &mi.seq4[rs];
    int32_t score = 0;
    for (int32_t j = 0; j < qe - qs; ++j) {
      if (qseq[j] >= 4 || tseq[j] >= 4) score += opt.e;  // upstream adds e2 (== e here)
      else score += qseq[j] == tseq[j] ? opt.a : -opt.b;
    }
    std::vector<uint32_t> c;
    PushCigar(c, 0, qe - qs);
    AppendCigar(r, c);
    r.dp_score += score;
    rs = re, qs = qe;
  }

  if (qe < qe0 && re < re0) {  // right extension
    ExtzResult ez;
    ExtzOnly(qe0 - qe, &qseq_all[qe], re0 - re, &mi.seq4[re], mat, opt.q, opt.e, opt.end_bonus, 0,
             ez);
    if (dbg) dbg->dp_cells_full += (int64_t)(qe0 - qe) * (re0 - re);
    if (!ez.cigar.empty()) {
      AppendCigar(r, ez.cigar);
      r.dp_score += ez.max;
    }
    re1 = re + (ez.reach_end ? ez.mqe_t + 1 : ez.max_t + 1);
    qe1 = qe + (ez.reach_end ? qe0 - qe : ez.max_q + 1);
  }

  r.rs = rs1, r.re = re1;
  if (rev) r.qs = qlen - qe1, r.qe = qlen - qs1;
  else r.qs = qs1, r.qe = qe1;
  if (r.has_p) UpdateExtra(r, &qseq_all[qs1], &mi.seq4[rs1], mat, (int8_t)opt.q, (int8_t)opt.e);
}

// hit.c: mm_filter_regs
void FilterRegs(const lgr_params& opt, int qlen, std::vector<Reg>& regs) {
  size_t k = 0;
  for (size_t i = 0; i < regs.size(); ++i) {
    const Reg& r = regs[i];
    bool flt = false;
    if (r.cnt < opt.min_cnt) flt = true;
    if (r.has_p) {
      if (r.mlen < opt.min_chain_score) flt = true;
      else if (r.dp_max < opt.min_dp_max) flt = true;
      else if (r.qs > qlen * opt.max_clip_ratio && qlen - r.qe > qlen * opt.max_clip_ratio) flt = true;
    }
    if (!flt) {
      if (k < i) regs[k] = regs[i];
      ++k;
    }
  }
  regs.resize(k);
}

// hit.c: mm_hit_sort
void HitSort(std::vector<Reg>& r) {
  const int n = (int)r.size();
  if (n <= 1) return;
  std::vector<U128> aux;
  for (int i = 0; i < n; ++i) {
    if (r[i].cnt > 0) {
      const int score = r[i].has_p ? r[i].dp_max : r[i].score;
      aux.push_back({(uint64_t)score << 32 | r[i].hash, (uint64_t)i});
    }
  }
  RadixSort128x(aux.data(), aux.data() + aux.size());
  std::vector<Reg> t(aux.size());
  for (int i = (int)aux.size() - 1; i >= 0; --i) t[aux.size() - 1 - i] = r[aux[i].y];
  r = t;
}

}  // namespace

// ---------------------------------------------------------------------------
// map.c: mm_map_frag for one segment (mm_map), MM_F_CIGAR|MM_F_SR, no re-chain
// (max_occ = 0), no long join, no mm_est_err.
// ---------------------------------------------------------------------------
std::vector<Reg> Map(const HapIndex& mi, const uint8_t* read, int qlen, uint32_t qname_hash,
                     const lgr_params& opt, int32_t mid_occ, MapDebug* dbg) {
  if (qlen <= 0) return {};
  uint32_t hash = qname_hash;
  hash ^= WangHash((uint32_t)qlen) + WangHash((uint32_t)opt.seed);
  hash = WangHash(hash);

  // collect_minimizers + mm_seed_mz_flt
  std::vector<U128> mv;
  Sketch(read, qlen, mi.w, mi.k, 0, mv);
  if (opt.q_occ_frac > 0.0f) SeedMzFlt(mv, mid_occ, opt.q_occ_frac);
  if (dbg) dbg->mv = mv;

  // seed.c: mm_seed_collect_all + mm_collect_matches
  std::vector<Seed> m;
  m.reserve(mv.size());
  for (size_t i = 0; i < mv.size(); ++i) {
    int t;
    const uint64_t* cr = mi.Get(mv[i].x >> 8, &t);
    if (t == 0) continue;
    Seed q{};
    q.q_pos = (uint32_t)mv[i].y, q.q_span = mv[i].x & 0xff, q.cr = cr, q.n = (uint32_t)t;
    q.seg_id = (uint32_t)(mv[i].y >> 32);
    q.is_tandem = q.flt = 0;
    if (i > 0 && mv[i].x >> 8 == mv[i - 1].x >> 8) q.is_tandem = 1;
    if (i + 1 < mv.size() && mv[i].x >> 8 == mv[i + 1].x >> 8) q.is_tandem = 1;
    m.push_back(q);
  }
  const int max_occ = mid_occ;
  if (opt.occ_dist > 0 && opt.max_max_occ > max_occ) {
    SeedSelect((int32_t)m.size(), m.data(), qlen, max_occ, opt.max_max_occ, opt.occ_dist);
  } else {
    for (auto& s : m)
      if ((int)s.n > max_occ) s.flt = 1;
  }
  int rep_st = 0, rep_en = 0, rep_len = 0;
  size_t n_m = 0;
  for (size_t i = 0; i < m.size(); ++i) {
    const Seed& q = m[i];
    if (q.flt) {
      const int en = (int)(q.q_pos >> 1) + 1, st = en - (int)q.q_span;
      if (st > rep_en) {
        rep_len += rep_en - rep_st;
        rep_st = st, rep_en = en;
      } else {
        rep_en = en;
      }
    } else {
      m[n_m++] = q;
    }
  }
  rep_len += rep_en - rep_st;
  m.resize(n_m);
  if (dbg) dbg->rep_len = rep_len;

  // map.c: collect_seed_hits
  std::vector<U128> a;
  for (const Seed& q : m) {
    for (uint32_t k = 0; k < q.n; ++k) {
      const uint64_t rk = q.cr[k];
      const int32_t rpos = (int32_t)((uint32_t)rk >> 1);
      U128 p;
      if ((rk & 1) == (q.q_pos & 1)) {
        p.x = (rk & 0xffffffff00000000ULL) | (uint64_t)(uint32_t)rpos;
        p.y = (uint64_t)q.q_span << 32 | q.q_pos >> 1;
      } else {
        p.x = 1ULL << 63 | (rk & 0xffffffff00000000ULL) | (uint64_t)(uint32_t)rpos;
        p.y = (uint64_t)q.q_span << 32 |
              (uint32_t)(qlen - ((int32_t)(q.q_pos >> 1) + 1 - (int32_t)q.q_span) - 1);
      }
      if (q.is_tandem) p.y |= kSeedTandem;
      a.push_back(p);
    }
  }
  RadixSort128x(a.data(), a.data() + a.size());
  if (dbg) dbg->anchors = a;

  // chaining limits (map.c: mm_map_frag)
  const int max_chain_gap_qry = qlen > opt.max_gap ? qlen : opt.max_gap;  // is_sr
  const int max_chain_gap_ref = opt.max_gap_ref > 0 ? opt.max_gap_ref : opt.max_gap;
  const float chn_pen_gap = (float)(opt.chain_gap_scale * 0.01 * mi.k);
  const float chn_pen_skip = (float)(opt.chain_skip_scale * 0.01 * mi.k);
  std::vector<uint64_t> u;
  std::vector<U128> ca =
      LchainDp(max_chain_gap_ref, max_chain_gap_qry, opt.bw, opt.max_chain_skip,
               opt.max_chain_iter, opt.min_cnt, opt.min_chain_score, chn_pen_gap, chn_pen_skip, a, u,
               dbg);
  if (dbg) dbg->u = u, dbg->chained = ca;

  std::vector<Reg> regs = GenRegs(hash, qlen, u, ca);

  // map.c: chain_post
  SetParent(opt.mask_level, opt.mask_len, regs, opt.a * 2 + opt.b);
  SelectSub(opt.pri_ratio, mi.k * 2, opt.best_n, 1, (int)(opt.max_gap * 0.8), regs);
  if (dbg) dbg->n_regs_chain = (int32_t)regs.size();

  // map.c: align_regs → align.c: mm_align_skeleton
  std::vector<uint8_t> qf(qlen), qr(qlen);
  for (int i = 0; i < qlen; ++i) {
    qf[i] = Nt4(read[i]);
    qr[qlen - 1 - i] = qf[i] < 4 ? 3 - qf[i] : 4;
  }
  const uint8_t* qseq0[2] = {qf.data(), qr.data()};
  for (Reg& r : regs) Align1(opt, mi, qlen, qseq0, r, ca, dbg);
  FilterRegs(opt, qlen, regs);
  HitSort(regs);
  SetParent(opt.mask_level, opt.mask_len, regs, opt.a * 2 + opt.b);
  SelectSub(opt.pri_ratio, mi.k * 2, opt.best_n, 0, (int)(opt.max_gap * 0.8), regs);
  return regs;
}

// ---------------------------------------------------------------------------
// ksw2_extz2_sse.c (KSW_EZ_EXTZ_ONLY, with CIGAR, exact max) restated in absolute
// scores.  ksw2 keeps Suzuki–Kasahara differences in int8 lanes; with q+e = 15
// and |mismatch| <= 2(q+e) no lane saturates, so the differences encode exactly
//   H(i,j) = max{ H(i-1,j-1)+s(i,j), E(i,j), F(i,j) }
//   E(i+1,j) = max{H(i,j)-q, E(i,j)} - e        F(i,j+1) = max{H(i,j)-q, F(i,j)} - e
// with H(-1,-1)=0, H(i,-1) = -(q+e(i+1)), H(-1,j) = -(q+e(j+1)), and the direction
// byte rules of the two code paths (default = left-aligned gaps; KSW_EZ_RIGHT).
// The band (w = 1.5*bw+1 = 15001) and z-drop (100000) never bind for the sizes
// this path sees; the batch validator enforces that.
// ---------------------------------------------------------------------------
void ExtzOnly(int qlen, const uint8_t* query, int tlen, const uint8_t* target, const int8_t* mat,
              int gapo, int gape, int end_bonus, int flag, ExtzResult& ez) {
  ez = ExtzResult();
  if (qlen <= 0 || tlen <= 0) return;
  const bool right = (flag & kEzRight) != 0;
  const int q = gapo, e = gape;
  std::vector<uint8_t> dir((size_t)tlen * qlen);
  std::vector<int32_t> hcol(qlen), ecol(qlen);
  for (int j = 0; j < qlen; ++j) {
    hcol[j] = -(q + e * (j + 1));
    ecol[j] = hcol[j] - q - e;
  }
  for (int i = 0; i < tlen; ++i) {
    int32_t hdiag = i == 0 ? 0 : -(q + e * i);
    const int32_t hleft = -(q + e * (i + 1));
    int32_t f = hleft - q - e;
    uint8_t* drow = &dir[(size_t)i * qlen];
    for (int j = 0; j < qlen; ++j) {
      const int32_t hd = hdiag + mat[target[i] * 5 + query[j]];
      const int32_t ee = ecol[j];
      int32_t h;
      uint8_t d;
      if (!right) {
        d = ee > hd ? 1 : 0;
        h = ee > hd ? ee : hd;
        if (f > h) d = 2, h = f;
      } else {
        d = hd > ee ? 0 : 1;
        h = hd > ee ? hd : ee;
        if (!(h > f)) d = 2, h = f;
      }
      const int32_t ho = h - q;
      if (!right) {
        if (ee > ho) d |= 0x08;
        if (f > ho) d |= 0x10;
      } else {
        if (ee >= ho) d |= 0x08;
        if (f >= ho) d |= 0x10;
      }
      drow[j] = d;
      hdiag = hcol[j];
      hcol[j] = h;
      ecol[j] = (ee > ho ? ee : ho) - e;
      f = (f > ho ? f : ho) - e;
      if (h > ez.max) ez.max = h, ez.max_t = i, ez.max_q = j;  // location unused when reach_end
    }
    if (hcol[qlen - 1] > ez.mqe) ez.mqe = hcol[qlen - 1], ez.mqe_t = i;
  }
  int i0, j0;
  if (ez.mqe + end_bonus > ez.max) {
    ez.reach_end = 1;
    i0 = ez.mqe_t, j0 = qlen - 1;
  } else if (ez.max_t >= 0 && ez.max_q >= 0) {
    i0 = ez.max_t, j0 = ez.max_q;
  } else {
    return;
  }
  // ksw2.h: ksw_backtrack
  int i = i0, j = j0, state = 0;
  std::vector<uint32_t>& c = ez.cigar;
  while (i >= 0 && j >= 0) {
    const uint8_t tmp = dir[(size_t)i * qlen + j];
    if (state == 0) state = tmp & 7;
    else if (!(tmp >> (state + 2) & 1)) state = 0;
    if (state == 0) state = tmp & 7;
    if (state == 0) PushCigar(c, 0, 1), --i, --j;
    else if (state == 1) PushCigar(c, 2, 1), --i;
    else PushCigar(c, 1, 1), --j;
  }
  if (i >= 0) PushCigar(c, 2, i + 1);
  if (j >= 0) PushCigar(c, 1, j + 1);
  if (!(flag & kEzRevCigar)) std::reverse(c.begin(), c.end());
}

}  // namespace mm2r
