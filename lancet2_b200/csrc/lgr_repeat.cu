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
// diagonal costs 2 * n / 32 warp steps whatever k is.  The mask does not depend on k at all, so consecutive jobs over the
// same sequence with the same threshold (a window's whole k-loop) share one CTA: the mask of a diagonal is built once and
// every still-open k is answered from it.  A job is closed at its first hit; the CTA stops when all its jobs are closed.
// Byte compares on the raw ASCII, like the reference (case and IUPAC letters count as written).
// There is no CPU fallback: lgr_repeat_create returns LGR_E_NO_DEVICE without a GPU.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/lancet_gpu_realign.h"

namespace {

constexpr int kRepThreads = 256;
constexpr int kRepWarps = kRepThreads / 32;
constexpr int kRepMaxLen = LGR_REPEAT_MAX_LEN;        // bases of one job's sequence (staged in shared memory)
constexpr int kRepWords = kRepMaxLen / 32 + 2;        // mask words of one diagonal

struct RepGroup {  // consecutive jobs over the same sequence with the same max_mismatches: they share the mismatch masks
  int32_t first, n;
};

struct RepDev {
  const uint8_t* seqs;
  const lgr_repeat_job* jobs;
  const RepGroup* groups;
  uint8_t* out;
  int* next_group;  // work counter (groups differ by 20x in cost: a static stride would pair the heavy ones up)
  int n_groups;
};

constexpr int kRepGroupJobs = 32;  // k values handled by one CTA (a window's k-loop has 20)

// One CTA per job group.  The mismatch mask of a diagonal depends on the sequence and the offset only, not on k, so
// the warp builds it once and then answers every still-open k of the group from it: (1 + K) * n / 32 warp steps per
// diagonal instead of 2 K * n / 32 (K = 20 for a window's k-loop: 1.9x fewer steps).
__global__ void __launch_bounds__(kRepThreads) k_repeat_scan(const RepDev D) {
  __shared__ uint8_t s_seq[kRepMaxLen];
  __shared__ uint32_t s_mask[kRepWarps][kRepWords];
  __shared__ uint16_t s_pre[kRepWarps][kRepWords];
  __shared__ int s_k[kRepGroupJobs];
  __shared__ unsigned s_found;  // bit j: job j of the group has its repeat
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ int s_grp;
  for (;;) {
    __syncthreads();  // the previous group is finished with the shared arrays (and with s_grp)
    if (threadIdx.x == 0) s_grp = atomicAdd(D.next_group, 1);
    __syncthreads();
    const int grp = s_grp;
    if (grp >= D.n_groups) break;
    const RepGroup G = D.groups[grp];
    const lgr_repeat_job j0 = D.jobs[G.first];
    const int len = j0.seq_len, mm = j0.max_mismatches;
    if (threadIdx.x == 0) s_found = 0;
    if (threadIdx.x < G.n) s_k[threadIdx.x] = D.jobs[G.first + threadIdx.x].k;
    if (len > kRepMaxLen) {  // not computed: flagged for the host
      if (threadIdx.x < G.n) D.out[G.first + threadIdx.x] = LGR_REPEAT_TOO_LONG;
      continue;
    }
    const uint8_t* seq = D.seqs + j0.seq_off;
    for (int p = threadIdx.x; p < len; p += kRepThreads) s_seq[p] = seq[p];
    __syncthreads();
    int k_min = s_k[0];
    for (int q = 1; q < G.n; ++q) k_min = s_k[q] < k_min ? s_k[q] : k_min;
    const unsigned all = G.n == 32 ? full : (1u << G.n) - 1u;
    const int d_end = len - k_min + 1;  // diagonals of the smallest k (base::SlidingView: none when the sequence is shorter than k)
    uint32_t* mask = s_mask[warp];
    uint16_t* pre = s_pre[warp];
    for (int d = 1 + warp; d < d_end; d += kRepWarps) {
      // which jobs are still open: one lane reads the flags (atomics on both sides, no data race) and the whole
      // warp follows its answer, so the lanes can never disagree about leaving the loop
      unsigned found = 0;
      if (lane == 0) found = atomicOr(&s_found, 0u);
      found = __shfl_sync(full, found, 0);
      if (found == all) break;
      const int span = len - d;  // positions p with a partner p + d
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
      unsigned newly = 0;
      for (int q = 0; q < G.n; ++q) {
        if (found >> q & 1u) continue;
        const int k = s_k[q];
        const int starts = len - k + 1 - d;  // k-mer pairs (i, i + d) of this k on the diagonal
        if (starts <= 0) continue;
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
        if (__any_sync(full, hit)) newly |= 1u << q;
      }
      if (newly && lane == 0) atomicOr(&s_found, newly);
      __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x < G.n) D.out[G.first + threadIdx.x] = (uint8_t)(s_found >> threadIdx.x & 1u);
  }
}

}  // namespace

struct lgr_rep_ctx {
  int device = 0, sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  void *d_seq = nullptr, *d_jobs = nullptr, *d_out = nullptr, *d_groups = nullptr, *d_ctr = nullptr;
  size_t cap_seq = 0, cap_jobs = 0, cap_out = 0, cap_groups = 0, cap_ctr = 0;
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
  if (c->d_groups) cudaFree(c->d_groups);
  if (c->d_ctr) cudaFree(c->d_ctr);
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
  // consecutive jobs over the same sequence with the same threshold share one CTA (and its mismatch masks)
  std::vector<RepGroup> groups;
  for (int i = 0; i < n_jobs;) {
    int n = 1;
    while (i + n < n_jobs && n < kRepGroupJobs && jobs[i + n].seq_off == jobs[i].seq_off && jobs[i + n].seq_len == jobs[i].seq_len &&
           jobs[i + n].max_mismatches == jobs[i].max_mismatches)
      ++n;
    groups.push_back(RepGroup{i, n});
    i += n;
  }
  // the queue hands out the expensive groups first (cost ~ jobs x length^2): a short tail
  std::stable_sort(groups.begin(), groups.end(), [&](const RepGroup& a, const RepGroup& b) {
    const int64_t la = jobs[a.first].seq_len, lb = jobs[b.first].seq_len;
    return (int64_t)(a.n + 1) * la * la > (int64_t)(b.n + 1) * lb * lb;
  });
  REP_CUDA(c, cudaSetDevice(c->device));
  int rc;
  if ((rc = rep_ensure(c, &c->d_seq, &c->cap_seq, (size_t)seq_bytes)) != LGR_OK) return rc;
  if ((rc = rep_ensure(c, &c->d_jobs, &c->cap_jobs, sizeof(lgr_repeat_job) * (size_t)n_jobs)) != LGR_OK) return rc;
  if ((rc = rep_ensure(c, &c->d_out, &c->cap_out, (size_t)n_jobs)) != LGR_OK) return rc;
  if ((rc = rep_ensure(c, &c->d_groups, &c->cap_groups, sizeof(RepGroup) * groups.size())) != LGR_OK) return rc;
  if ((rc = rep_ensure(c, &c->d_ctr, &c->cap_ctr, sizeof(int))) != LGR_OK) return rc;
  REP_CUDA(c, cudaMemsetAsync(c->d_ctr, 0, sizeof(int), c->stream));
  if (seq_bytes > 0) REP_CUDA(c, cudaMemcpyAsync(c->d_seq, seqs, (size_t)seq_bytes, cudaMemcpyHostToDevice, c->stream));
  REP_CUDA(c, cudaMemcpyAsync(c->d_jobs, jobs, sizeof(lgr_repeat_job) * (size_t)n_jobs, cudaMemcpyHostToDevice, c->stream));
  REP_CUDA(c, cudaMemcpyAsync(c->d_groups, groups.data(), sizeof(RepGroup) * groups.size(), cudaMemcpyHostToDevice, c->stream));
  RepDev D{(const uint8_t*)c->d_seq, (const lgr_repeat_job*)c->d_jobs, (const RepGroup*)c->d_groups, (uint8_t*)c->d_out, (int*)c->d_ctr, (int)groups.size()};
  REP_CUDA(c, cudaEventRecord(c->ev0, c->stream));
  const int n_groups = (int)groups.size();
  const int grid = n_groups < c->sm_count * 8 ? n_groups : c->sm_count * 8;
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
