// ext_dpx_bench.cu — settles the "DPX cells" question of the north_star with a measurement
// (the round-1 claim "s16x2 DPX cells are slower" had no committed numbers).
//
// The ksw2-style extension cell of this path (lgr_core.cuh: ext_cell) produces, besides H/E/F, a
// direction byte (source + two continuation bits) per cell.  Two implementations of the same
// anti-diagonal wavefront over one warp are timed on identical synthetic tails:
//   A  one query row per lane, 32-bit arithmetic on values that travel as packed int16 pairs
//      (what k_ext_warp does): IMNMX / compare / select
//   B  two query rows per lane as s16x2, the recurrences through the DPX instructions
//      (__vibmax_s16x2 = VIBMNMX with predicates for the direction bits, __vimax3_s16x2)
// Both write the direction bytes to shared memory (the traceback needs them) and report the same
// checksums (sum of direction bytes, maximum H), so a speed difference is not a work difference.
// B' is B without the direction bytes: the ceiling DPX would give a score-only aligner.
// Output: one JSON line with GCUPS of A, B, B' over full-occupancy grids.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/ext_dpx_bench tools/ext_dpx_bench.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                   \
  do {                                                                          \
    cudaError_t e_ = (x);                                                       \
    if (e_ != cudaSuccess) {                                                    \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                  \
      exit(1);                                                                  \
    }                                                                           \
  } while (0)

constexpr int kT = 96;        // target columns of every synthetic tail
constexpr int kWarps = 4;     // warps per CTA
constexpr int kQ = 12, kE = 3, kA = 1, kB = 4;

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16, x *= 0x7feb352du, x ^= x >> 15, x *= 0x846ca68bu, x ^= x >> 16;
  return x;
}
// base c of tail `w`: the query copies the target with 6 % substitutions
__device__ __forceinline__ int tbase(uint32_t w, int i) { return (int)(mix(w * 977u + (uint32_t)i) & 3u); }
__device__ __forceinline__ int qbase(uint32_t w, int j) {
  const uint32_t h = mix(w * 131u + 7919u * (uint32_t)j);
  return (h & 15u) == 0 ? (int)((h >> 8) & 3u) : tbase(w, j);
}

// ---- A: one row per lane, M = 32 rows ------------------------------------------------------
__global__ void __launch_bounds__(kWarps * 32) k_scalar(int reps, unsigned long long* out) {
  __shared__ uint8_t s_dir[kWarps][(kT + 31) * 32];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gw = blockIdx.x * kWarps + warp;
  uint8_t* dir = s_dir[warp];
  unsigned long long dsum = 0;
  int hmax = 0;
  for (int rep = 0; rep < reps; ++rep) {
    const uint32_t w = gw * 1315423911u + (uint32_t)rep;
    const int j = lane;
    const int qc = qbase(w, j);
    int e_cur = -(kQ + kE * (j + 1)) - kQ - kE;
    int diag = j == 0 ? 0 : -(kQ + kE * j);
    int hf = 0;
    const int nsteps = kT + 31;
    for (int s = 0; s < nsteps; ++s) {
      int up_hf = __shfl_up_sync(full, hf, 1);
      if (lane == 0) {
        const int h0 = -(kQ + kE * (s + 1));
        up_hf = (int)(((uint32_t)h0 & 0xffffu) | ((uint32_t)(h0 - kQ - kE) << 16));
      }
      const int i = s - lane;
      if (i >= 0 && i < kT) {
        const int tc = tbase(w, i);
        const int up_h = (int)(int16_t)(up_hf & 0xffff), f = up_hf >> 16;
        const int hd = diag + (tc == qc ? kA : -kB), ee = e_cur;
        int d = ee > hd ? 1 : 0;
        int h = ee > hd ? ee : hd;
        if (f > h) d = 2, h = f;
        const int ho = h - kQ;
        if (ee > ho) d |= 0x08;
        if (f > ho) d |= 0x10;
        e_cur = (ee > ho ? ee : ho) - kE;
        const int fn = (f > ho ? f : ho) - kE;
        dir[s * 32 + lane] = (uint8_t)d;
        diag = up_h;
        hf = (int)(((uint32_t)h & 0xffffu) | ((uint32_t)fn << 16));
        hmax = h > hmax ? h : hmax;
      }
    }
    __syncwarp();
    for (int x = lane; x < nsteps * 32; x += 32) {
      const int s = x >> 5, l = x & 31, i = s - l;
      if (i >= 0 && i < kT) dsum += dir[x];
    }
    __syncwarp();
  }
  for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(full, dsum, o), hmax = max(hmax, __shfl_xor_sync(full, hmax, o));
  if (lane == 0) atomicAdd(&out[0], dsum), atomicMax(&out[1], (unsigned long long)hmax);
}

// ---- B: two rows per lane as s16x2, M = 64 rows ---------------------------------------------
__device__ __forceinline__ unsigned pack2(int lo, int hi) { return ((unsigned)lo & 0xffffu) | ((unsigned)hi << 16); }

template <bool DIR>
__global__ void __launch_bounds__(kWarps * 32) k_dpx(int reps, unsigned long long* out) {
  __shared__ uint16_t s_dir[kWarps][(kT + 63) * 32];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gw = blockIdx.x * kWarps + warp;
  uint16_t* dir = s_dir[warp];
  unsigned long long dsum = 0;
  int hmax = 0;
  const unsigned q2 = pack2(kQ, kQ), e2 = pack2(kE, kE);
  for (int rep = 0; rep < reps; ++rep) {
    const uint32_t w = gw * 1315423911u + (uint32_t)rep;  // the same tails as k_scalar: rows 0..31 must give the same bytes
    const int j0 = 2 * lane, j1 = j0 + 1;
    const int qc0 = qbase(w, j0), qc1 = qbase(w, j1);
    unsigned E2 = pack2(-(kQ + kE * (j0 + 1)) - kQ - kE, -(kQ + kE * (j1 + 1)) - kQ - kE);
    int diag_lo = j0 == 0 ? 0 : -(kQ + kE * j0);  // H(-1, j0-1)
    int hlo1 = -(kQ + kE * (j0 + 1));             // low row's H one step ago; before its first column: H(-1, j0)
    int hlo2 = 0;                                  // ... two steps ago (the high row's diagonal)
    int flo1 = 0;                                  // low row's F-out one step ago
    unsigned hf_hi = 0;                            // (H, F-out) of my high row's last cell
    int tc_prev = 0;
    const int nsteps = kT + 63;
    for (int s = 0; s < nsteps; ++s) {
      unsigned up = __shfl_up_sync(full, hf_hi, 1);
      if (lane == 0) {
        const int h0 = -(kQ + kE * (s + 1));
        up = pack2(h0, h0 - kQ - kE);
      }
      const int i0 = s - 2 * lane, i1 = i0 - 1;
      const bool ok0 = i0 >= 0 && i0 < kT, ok1 = i1 >= 0 && i1 < kT;
      if (ok0 || ok1) {
        const int tc0 = ok0 ? tbase(w, i0) : 0;
        const int up_h = (int)(int16_t)(up & 0xffff), up_f = (int)up >> 16;
        // high row (i1, j1): up = the low row's cell of the previous step, diagonal = two steps ago
        const int dg_hi = i1 == 0 ? -(kQ + kE * (j0 + 1)) : hlo2;  // H(i1-1, j0); H(-1, j0) on the first column
        const unsigned hd2 = pack2(diag_lo + (tc0 == qc0 ? kA : -kB), dg_hi + (tc_prev == qc1 ? kA : -kB));
        const unsigned F2 = pack2(up_f, flo1);
        bool p1h, p1l, p2h, p2l, p3h, p3l, p4h, p4l;
        const unsigned h2a = __vibmax_s16x2(hd2, E2, &p1h, &p1l);  // p1: hd >= ee
        const unsigned h2 = __vibmax_s16x2(h2a, F2, &p2h, &p2l);   // p2: max(hd, ee) >= f
        const unsigned ho2 = __vsub2(h2, q2);
        const unsigned en2 = __vsub2(__vibmax_s16x2(ho2, E2, &p3h, &p3l), e2);  // p3: ho >= ee
        const unsigned fn2 = __vsub2(__vibmax_s16x2(ho2, F2, &p4h, &p4l), e2);  // p4: ho >= f
        if (DIR) {
          const unsigned dlo = (!p2l ? 2u : (!p1l ? 1u : 0u)) | (!p3l ? 8u : 0u) | (!p4l ? 16u : 0u);
          const unsigned dhi = (!p2h ? 2u : (!p1h ? 1u : 0u)) | (!p3h ? 8u : 0u) | (!p4h ? 16u : 0u);
          dir[s * 32 + lane] = (uint16_t)((ok0 ? dlo : 0u) | (ok1 ? dhi : 0u) << 8);
        }
        const int h_lo = (int)(int16_t)(h2 & 0xffff), h_hi = (int)h2 >> 16;
        // commit per row (a row outside its column range keeps its state)
        unsigned keep = 0;
        if (ok0) {
          diag_lo = up_h;
          hlo2 = hlo1, hlo1 = h_lo, flo1 = (int)(int16_t)(fn2 & 0xffff);
          keep |= 0x0000ffffu;
          if (lane < 16) hmax = h_lo > hmax ? h_lo : hmax;
        } else {
          hlo2 = hlo1;
        }
        if (ok1) {
          hf_hi = pack2(h_hi, (int)fn2 >> 16);
          keep |= 0xffff0000u;
          if (lane < 16) hmax = h_hi > hmax ? h_hi : hmax;
        }
        E2 = (en2 & keep) | (E2 & ~keep);
        tc_prev = tc0;
      }
    }
    if (DIR) {
      __syncwarp();
      for (int x = lane; x < nsteps * 32; x += 32) {
        const int s = x >> 5, l = x & 31, i0 = s - 2 * l, i1 = i0 - 1;
        if (l >= 16) continue;  // rows 0..31 only: comparable with k_scalar
        if (i0 >= 0 && i0 < kT) dsum += dir[x] & 0xff;  // slots of steps outside a row's columns are never written
        if (i1 >= 0 && i1 < kT) dsum += dir[x] >> 8;
      }
      __syncwarp();
    }
  }
  for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(full, dsum, o), hmax = max(hmax, __shfl_xor_sync(full, hmax, o));
  if (lane == 0) atomicAdd(&out[0], dsum), atomicMax(&out[1], (unsigned long long)hmax);
}

template <typename F>
static double time_ms(F launch) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int it = 0; it < 5; ++it) {
    CK(cudaEventRecord(a));
    launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    best = ms < best ? ms : best;
  }
  return best;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  unsigned long long* d_out;
  CK(cudaMalloc(&d_out, 6 * sizeof(unsigned long long)));
  CK(cudaMemset(d_out, 0, 6 * sizeof(unsigned long long)));
  const int blocks = prop.multiProcessorCount * 8, reps = 64;
  const double ms_a = time_ms([&] { k_scalar<<<blocks, kWarps * 32>>>(reps, d_out); });
  const double ms_b = time_ms([&] { k_dpx<true><<<blocks, kWarps * 32>>>(reps, d_out + 2); });
  const double ms_c = time_ms([&] { k_dpx<false><<<blocks, kWarps * 32>>>(reps, d_out + 4); });
  unsigned long long h[6];
  CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
  const double tails = (double)blocks * kWarps * reps;
  const double cells_a = tails * 32 * kT, cells_b = tails * 64 * kT;
  printf("{\"device\": \"%s\", \"tail\": \"%d target columns; A: 32 query rows per warp, B: 64 (two per lane)\", "
         "\"scalar_one_row_per_lane\": {\"ms\": %.3f, \"gcups\": %.1f, \"hmax\": %llu}, "
         "\"dpx_s16x2_two_rows_per_lane\": {\"ms\": %.3f, \"gcups\": %.1f, \"hmax\": %llu}, "
         "\"dpx_s16x2_no_direction_bytes\": {\"ms\": %.3f, \"gcups\": %.1f, \"hmax\": %llu}, "
         "\"rows_0_31_identical\": %s, \"direction_byte_sum\": [%llu, %llu], "
         "\"note\": \"best of 5 launches (6 launches accumulate into the checksums), %d CTAs x %d warps, %d tails per warp; A and B run the same "
         "tails, B with 64 rows: its rows 0..31 are the DP of A, so their direction-byte sums and maxima must agree\"}\n",
         prop.name, kT, ms_a, cells_a / ms_a / 1e6, h[1], ms_b, cells_b / ms_b / 1e6, h[3], ms_c, cells_b / ms_c / 1e6, h[5],
         (h[0] == h[2] && h[1] == h[3]) ? "true" : "false", h[0], h[2], blocks, kWarps, reps);
  return 0;
}
