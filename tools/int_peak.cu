// int_peak.cu — measures the integer / DPX / shuffle issue peaks the realignment kernels are bound
// by ("the INT/DPX peak is not in MEASURED_PEAKS.json; the builder must measure it").
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/int_peak tools/int_peak.cu
// Run on the GPU box; prints one JSON object.  Every kernel is a long unrolled dependent-chain ×
// ILP loop so that the measured rate is the pipe's issue rate, not memory or launch overhead.
#include <cuda_runtime.h>

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

constexpr int kIlp = 8;
constexpr int kInner = 64;

// op codes
enum { OP_IADD = 0, OP_IMAD, OP_LOP3, OP_VIADDMAX, OP_VIMAX3, OP_VIMAX3_S32, OP_SHFL, OP_IMNMX, OP_COUNT };

template <int OP>
__device__ __forceinline__ unsigned step(unsigned a, unsigned b, unsigned c) {
  if (OP == OP_IADD) return a + b + c;                                   // IADD3
  if (OP == OP_IMAD) return a * b + c;                                   // IMAD
  if (OP == OP_LOP3) return (a & b) ^ c;                                 // LOP3
  if (OP == OP_VIADDMAX) return __viaddmax_s16x2(a, b, c);               // VIADDMNMX.S16x2 (DPX)
  if (OP == OP_VIMAX3) return __vimax3_s16x2(a, b, c);                   // VIMNMX3.S16x2 (DPX)
  if (OP == OP_VIMAX3_S32) return (unsigned)__vimax3_s32((int)a, (int)b, (int)c);
  if (OP == OP_SHFL) return __shfl_xor_sync(0xffffffffu, a, 1) + c;      // SHFL + IADD
  if (OP == OP_IMNMX) return (unsigned)max((int)a, (int)b) + c;          // IMNMX + IADD
  return a;
}

template <int OP>
__global__ void __launch_bounds__(256) k_peak(unsigned* out, int iters, unsigned seed) {
  unsigned v[kIlp];
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int i = 0; i < kIlp; ++i) v[i] = seed * (t + 1) + i * 2654435761u;
  const unsigned b = seed | 1u, c = seed ^ 0x9e3779b9u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < kInner; ++j) {
#pragma unroll
      for (int i = 0; i < kIlp; ++i) v[i] = step<OP>(v[i], b, c + j);
    }
  }
  unsigned acc = 0;
#pragma unroll
  for (int i = 0; i < kIlp; ++i) acc ^= v[i];
  if (acc == 0x12345678u) out[t] = acc;  // keeps the chain alive
}

template <int OP>
double run(int sms, unsigned* out, int ops_per_step) {
  const int blocks = sms * 8, threads = 256, iters = 2000;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  k_peak<OP><<<blocks, threads>>>(out, 50, 1234567u);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0));
    k_peak<OP><<<blocks, threads>>>(out, iters, 1234567u + rep);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double thread_ops = (double)blocks * threads * iters * kInner * kIlp * ops_per_step;
  return thread_ops / (best * 1e-3) / 1e12;  // tera thread-ops / s
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  unsigned* out;
  CK(cudaMalloc(&out, sizeof(unsigned) * p.multiProcessorCount * 8 * 256));
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("{\"device\": \"%s\", \"sms\": %d, \"sm_clock_mhz_attr\": %d, \"unit\": \"tera thread-ops/s\"", p.name,
         p.multiProcessorCount, clk_khz / 1000);
  printf(", \"iadd3\": %.2f", run<OP_IADD>(p.multiProcessorCount, out, 1));
  printf(", \"imad\": %.2f", run<OP_IMAD>(p.multiProcessorCount, out, 1));
  printf(", \"lop3\": %.2f", run<OP_LOP3>(p.multiProcessorCount, out, 1));
  printf(", \"imnmx_plus_iadd\": %.2f", run<OP_IMNMX>(p.multiProcessorCount, out, 2));
  printf(", \"viaddmax_s16x2\": %.2f", run<OP_VIADDMAX>(p.multiProcessorCount, out, 1));
  printf(", \"vimax3_s16x2\": %.2f", run<OP_VIMAX3>(p.multiProcessorCount, out, 1));
  printf(", \"vimax3_s32\": %.2f", run<OP_VIMAX3_S32>(p.multiProcessorCount, out, 1));
  printf(", \"shfl_plus_iadd\": %.2f", run<OP_SHFL>(p.multiProcessorCount, out, 2));
  // warp-instruction issue ceiling implied by the best single-op rate: thread-ops / 32
  printf(", \"note\": \"s16x2 ops process 2 cells per thread-op; warp-instr/s = value/32\"}\n");
  return 0;
}
