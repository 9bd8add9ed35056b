// lgr_repeat.cu — SURVEY.md §8f #3 (first half): repeat detection over the sliding k-mers of
// reference windows, batched.  C-ABI in include/lancet_gpu_realign.h (lgr_repeat_*).
//
// Replaces, per (window sequence, k, max_mismatches) job:
//   cbdg::Graph::HasExactOrApproxRepeat    graph.h:127-131  (HasRepeat(SlidingView(seq, k), 2), once per k of the k-loop)
//   VariantBuilder::ShouldSkipWindow       variant_builder.cpp:116-117 (HasExactRepeat(SlidingView(seq, max_k)))
//   base::HasRepeat / IsWithinHammingDist  base/repeat.cpp:55-217, 348-371
// Answer: do two k-mers at different offsets differ in at most max_mismatches BYTE positions?
//
// The reference compares every k-mer pair (O(n^2) pairs, SIMD compare with early exit).  Here the
// pairs are organised by diagonal: for an offset d, x_d[p] = (seq[p] != seq[p + d]) and the Hamming
// distance of the k-mers at i and i + d is the number of set bits of x_d in [i, i + k) — a sliding
// window sum.  One CTA takes one job and stages the sequence in shared memory; each of its warps
// takes diagonals d = warp, warp + 8, ...: one ballot per 32 positions turns x_d into bit masks with
// a running popcount, then every lane evaluates one window start with two prefix look-ups.  A
// diagonal costs 2 * n / 32 warp steps whatever k is.  All warps stop at the first hit.
// Byte compares on the raw ASCII, like the reference (case and IUPAC letters count as written).
// There is no CPU fallback: lgr_repeat_create returns LGR_E_NO_DEVICE without a GPU.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <string>

#include "../../include/lancet_gpu_realign.h"

namespace {

constexpr int kRepThreads = 256;
constexpr int kRepWarps = kRepThreads / 32;
constexpr int kRepMaxLen = LGR_REPEAT_MAX_LEN;        // bases of one job's sequence (staged in shared memory)
constexpr int kRepWords = kRepMaxLen / 32 + 2;        // mask words of one diagonal

struct RepDev {
  const uint8_t* seqs;
  const lgr_repeat_job* jobs;
  uint8_t* out;
  int n_jobs;
};

__global__ void __launch_bounds__(kRepThreads) k_repeat_scan(const RepDev D) {
  __shared__ uint8_t s_seq[kRepMaxLen];
  __shared__ uint32_t s_mask[kRepWarps][kRepWords];
  __shared__ uint16_t s_pre[kRepWarps][kRepWords];
  __shared__ int s_found;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int job = blockIdx.x; job < D.n_jobs; job += gridDim.x) {
    const lgr_repeat_job j = D.jobs[job];
    const int len = j.seq_len, k = j.k, mm = j.max_mismatches;
    __syncthreads();  // the previous job is finished with the shared arrays
    if (threadIdx.x == 0) s_found = 0;
    if (len > kRepMaxLen || k <= 0 || mm < 0) {  // not computed: flagged for the host
      if (threadIdx.x == 0) D.out[job] = len > kRepMaxLen ? LGR_REPEAT_TOO_LONG : 0;
      continue;
    }
    const uint8_t* seq = D.seqs + j.seq_off;
    for (int p = threadIdx.x; p < len; p += kRepThreads) s_seq[p] = seq[p];
    __syncthreads();
    const int n_kmers = len - k + 1;  // base::SlidingView: none when the sequence is shorter than k
    uint32_t* mask = s_mask[warp];
    uint16_t* pre = s_pre[warp];
    for (int d = 1 + warp; d < n_kmers; d += kRepWarps) {
      // another warp found a repeat: one lane reads the flag (atomics on both sides, no data race) and the
      // whole warp follows its answer, so the lanes can never disagree about leaving the loop
      int stop = 0;
      if (lane == 0) stop = atomicOr(&s_found, 0);
      if (__shfl_sync(full, stop, 0)) break;
      const int span = len - d;            // positions p with a partner p + d
      const int starts = n_kmers - d;      // k-mer pairs (i, i + d) on this diagonal
      const int words = (span + 31) >> 5;
      int run = 0;
      for (int c = 0; c < words; ++c) {
        const int p = (c << 5) + lane;
        const unsigned m = __ballot_sync(full, p < span && s_seq[p] != s_seq[p + d]);
        if (lane == 0) mask[c] = m, pre[c] = (uint16_t)run;
        run += __popc(m);
      }
      if (lane == 0) mask[words] = 0, pre[words] = (uint16_t)run;  // P(span) when span is a multiple of 32
      __syncwarp();
      bool hit = false;
      for (int base = 0; base < starts; base += 32) {
        const int i = base + lane;
        if (i < starts) {
          const int e = i + k;
          const int pb = pre[i >> 5] + __popc(mask[i >> 5] & ((1u << (i & 31)) - 1u));
          const int pe = pre[e >> 5] + __popc(mask[e >> 5] & ((1u << (e & 31)) - 1u));
          hit |= pe - pb <= mm;
        }
      }
      if (__any_sync(full, hit)) {
        if (lane == 0) atomicExch(&s_found, 1);
        break;
      }
      __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x == 0) D.out[job] = s_found ? 1 : 0;
  }
}

}  // namespace

struct lgr_rep_ctx {
  int device = 0, sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  void *d_seq = nullptr, *d_jobs = nullptr, *d_out = nullptr;
  size_t cap_seq = 0, cap_jobs = 0, cap_out = 0;
  std::string err;
};

static thread_local std::string g_rep_create_err;

#define REP_CUDA(c, call)                                                      \
  do {                                                                         \
    cudaError_t e_ = (call);                                                   \
    if (e_ != cudaSuccess) {                                                   \
      (c)->err = std::string(#call) + ": " + cudaGetErrorString(e_);           \
      return LGR_E_CUDA;                                                       \
    }                                                                          \
  } while (0)

static int rep_ensure(lgr_rep_ctx* c, void** p, size_t* cap, size_t bytes) {
  if (bytes < 256) bytes = 256;
  if (*cap >= bytes) return LGR_OK;
  if (*p) cudaFree(*p);
  *p = nullptr, *cap = 0;
  if (cudaMalloc(p, bytes + bytes / 2) != cudaSuccess) {
    c->err = "cudaMalloc failed";
    (void)cudaGetLastError();
    return LGR_E_NOMEM;
  }
  *cap = bytes + bytes / 2;
  return LGR_OK;
}

extern "C" {

int lgr_repeat_create(int device_ordinal, lgr_rep_ctx** out) {
  if (!out) return LGR_E_ARG;
  *out = nullptr;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0 || device_ordinal < 0 || device_ordinal >= n_dev) {
    g_rep_create_err = "no usable CUDA device (this path has no CPU fallback)";
    (void)cudaGetLastError();
    return LGR_E_NO_DEVICE;
  }
  lgr_rep_ctx* c = new lgr_rep_ctx();
  c->device = device_ordinal;
  cudaDeviceProp prop;
  if (cudaSetDevice(device_ordinal) != cudaSuccess || cudaGetDeviceProperties(&prop, device_ordinal) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    g_rep_create_err = "cudaSetDevice / cudaStreamCreate failed";
    delete c;
    return LGR_E_CUDA;
  }
  c->sm_count = prop.multiProcessorCount;
  cudaEventCreate(&c->ev0), cudaEventCreate(&c->ev1);
  *out = c;
  return LGR_OK;
}

void lgr_repeat_destroy(lgr_rep_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->d_seq) cudaFree(c->d_seq);
  if (c->d_jobs) cudaFree(c->d_jobs);
  if (c->d_out) cudaFree(c->d_out);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char* lgr_repeat_last_error(const lgr_rep_ctx* c) { return c ? c->err.c_str() : g_rep_create_err.c_str(); }

int lgr_repeat_scan(lgr_rep_ctx* c, const uint8_t* seqs, int64_t seq_bytes, const lgr_repeat_job* jobs, int32_t n_jobs,
                    uint8_t* has_repeat, float* ms_kernels) {
  if (!c || n_jobs < 0 || seq_bytes < 0 || (n_jobs > 0 && (!jobs || !has_repeat)) || (seq_bytes > 0 && !seqs)) return LGR_E_ARG;
  c->err.clear();
  if (ms_kernels) *ms_kernels = 0.0f;
  if (n_jobs == 0) return LGR_OK;
  bool too_long = false;
  for (int i = 0; i < n_jobs; ++i) {
    const lgr_repeat_job& j = jobs[i];
    if (j.seq_len < 0 || j.seq_off < 0 || j.seq_off + j.seq_len > seq_bytes || j.k <= 0 || j.max_mismatches < 0)
      return c->err = "bad repeat job (offsets outside the sequence buffer, k <= 0 or max_mismatches < 0)", LGR_E_ARG;
    too_long |= j.seq_len > LGR_REPEAT_MAX_LEN;
  }
  REP_CUDA(c, cudaSetDevice(c->device));
  int rc;
  if ((rc = rep_ensure(c, &c->d_seq, &c->cap_seq, (size_t)seq_bytes)) != LGR_OK) return rc;
  if ((rc = rep_ensure(c, &c->d_jobs, &c->cap_jobs, sizeof(lgr_repeat_job) * (size_t)n_jobs)) != LGR_OK) return rc;
  if ((rc = rep_ensure(c, &c->d_out, &c->cap_out, (size_t)n_jobs)) != LGR_OK) return rc;
  if (seq_bytes > 0) REP_CUDA(c, cudaMemcpyAsync(c->d_seq, seqs, (size_t)seq_bytes, cudaMemcpyHostToDevice, c->stream));
  REP_CUDA(c, cudaMemcpyAsync(c->d_jobs, jobs, sizeof(lgr_repeat_job) * (size_t)n_jobs, cudaMemcpyHostToDevice, c->stream));
  RepDev D{(const uint8_t*)c->d_seq, (const lgr_repeat_job*)c->d_jobs, (uint8_t*)c->d_out, n_jobs};
  REP_CUDA(c, cudaEventRecord(c->ev0, c->stream));
  const int grid = n_jobs < c->sm_count * 8 ? n_jobs : c->sm_count * 8;
  k_repeat_scan<<<grid, kRepThreads, 0, c->stream>>>(D);
  REP_CUDA(c, cudaGetLastError());
  REP_CUDA(c, cudaEventRecord(c->ev1, c->stream));
  REP_CUDA(c, cudaMemcpyAsync(has_repeat, c->d_out, (size_t)n_jobs, cudaMemcpyDeviceToHost, c->stream));
  REP_CUDA(c, cudaStreamSynchronize(c->stream));
  if (ms_kernels) REP_CUDA(c, cudaEventElapsedTime(ms_kernels, c->ev0, c->ev1));
  if (too_long) {
    c->err = "some sequences are longer than LGR_REPEAT_MAX_LEN (their answers are LGR_REPEAT_TOO_LONG)";
    return LGR_E_PARTIAL;
  }
  return LGR_OK;
}

}  // extern "C"
