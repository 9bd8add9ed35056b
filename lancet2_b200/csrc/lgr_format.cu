// lgr_format.cu — VariantSupport aggregation + FORMAT math on the device:
// kernels k_fmt_dedup / k_fmt_metrics and the C-ABI entry points lgr_format_* declared in
// include/lancet_gpu_realign.h.  Own translation unit, compiled with -fmad=false so that the
// f64 expressions of lgr_format.cuh are evaluated as written (the host emulation is compiled
// with -ffp-contract=off).  No CPU fallback: a missing device is LGR_E_NO_DEVICE.
#include <cuda_runtime.h>

#include <cstring>

#include <string>
#include <vector>

#include "lgr_format.cuh"

namespace {

using lgr_fmt::Ev;

// global (L1-cached) rather than __constant__: lanes index it with different qualities
__device__ const double g_phred[256] = {
#include "phred_lut.inc"
};

struct FmtDev {
  Ev e;
  const int64_t* sup_begin;
  const int32_t* sup_n_alleles;
  const int32_t* sup_variant_len;
  const int32_t* sup_total_haps;
  uint8_t* keep;
  lgr_format* out;
  int32_t n_supports;
  int64_t n_evidence;
};

// one thread per evidence record: AddEvidence's first-seen rule against the earlier records of
// the same support (variant_support.cpp:28-29).  Supports are a few hundred records, so the scan
// is short and its loads are warp-uniform or coalesced.
__global__ void __launch_bounds__(256) k_fmt_dedup(const __grid_constant__ FmtDev D) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D.n_evidence) return;
  int lo = 0, hi = D.n_supports;  // last support with sup_begin <= i
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (D.sup_begin[mid] <= i) lo = mid;
    else hi = mid;
  }
  D.keep[i] = lgr_fmt::dedup_keep(D.e.allele, D.e.rname_hash, D.sup_begin[lo], i);
}

struct CtaDev {
  int tid;
  __device__ __forceinline__ bool leader() const { return tid == 0; }
#ifdef LGR_FMT_SORT  // not in the default build until it has run on a GPU
  uint64_t* keys_;
  uint8_t* tags_;
  __device__ __forceinline__ uint64_t* sort_keys() const { return keys_; }
  __device__ __forceinline__ uint8_t* sort_tags() const { return tags_; }
  template <class F>
  __device__ __forceinline__ void each(const F& f) {
    f(tid);
    __syncthreads();
  }
#endif
  template <class A, class F>
  __device__ __forceinline__ A reduce(const F& f) {
    constexpr int nd = (int)(sizeof(A::d) / sizeof(double)), ni = (int)(sizeof(A::i) / sizeof(long long));
    __shared__ double s_d[lgr_fmt::kWarps][nd];
    __shared__ long long s_i[lgr_fmt::kWarps][ni];
    A a = f(tid);
#pragma unroll
    for (int off = lgr_fmt::kLanes / 2; off > 0; off >>= 1) {
#pragma unroll
      for (int k = 0; k < nd; ++k) a.d[k] = a.d[k] + __shfl_xor_sync(0xffffffffu, a.d[k], off);
#pragma unroll
      for (int k = 0; k < ni; ++k) a.i[k] = a.i[k] + __shfl_xor_sync(0xffffffffu, a.i[k], off);
    }
    if ((tid & 31) == 0) {
#pragma unroll
      for (int k = 0; k < nd; ++k) s_d[tid >> 5][k] = a.d[k];
#pragma unroll
      for (int k = 0; k < ni; ++k) s_i[tid >> 5][k] = a.i[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < nd; ++k) a.d[k] = (s_d[0][k] + s_d[1][k]) + (s_d[2][k] + s_d[3][k]);
#pragma unroll
    for (int k = 0; k < ni; ++k) a.i[k] = (s_i[0][k] + s_i[1][k]) + (s_i[2][k] + s_i[3][k]);
    __syncthreads();  // the partials may be overwritten by the next reduction of this instantiation
    return a;
  }
};

// one 128-thread CTA per (support, task): blockIdx.y is the task (lgr_fmt::kTask*), supports are
// taken in a grid-stride loop along x.  A batch of a few hundred supports is a few thousand CTAs.
__global__ void __launch_bounds__(lgr_fmt::kThreads) k_fmt_metrics(const __grid_constant__ FmtDev D) {
  const unsigned task = 1u << blockIdx.y;
#ifdef LGR_FMT_SORT
  __shared__ uint64_t s_keys[lgr_fmt::kSortCap];
  __shared__ uint8_t s_tags[lgr_fmt::kSortCap];
  CtaDev w{(int)threadIdx.x, s_keys, s_tags};
#else
  CtaDev w{(int)threadIdx.x};
#endif
  for (int s = (int)blockIdx.x; s < D.n_supports; s += (int)gridDim.x) {
    // a support with more alleles than the record's fixed arrays hold is not computed here (the host marks its
    // record LGR_FMT_WIDE): it runs as an empty one-allele support so that nothing indexes past the arrays
    const int K = D.sup_n_alleles[s];
    const bool wide = K > LGR_FMT_MAX_ALLELES;
    const int64_t b = D.sup_begin[s];
    lgr_fmt::support_metrics(w, D.e, b, wide ? b : D.sup_begin[s + 1], wide ? 1 : K, D.sup_variant_len[s], D.sup_total_haps[s], g_phred,
                             &D.out[s], task);
  }
}

struct Buf {
  void* p = nullptr;
  size_t cap = 0;
};

}  // namespace

struct lgr_fmt_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  Buf b_in, b_keep, b_out;  // one packed input arena, the dedup flags, the results
};

static thread_local std::string g_fmt_create_err;

#define FMT_CUDA(ctx, call)                                            \
  do {                                                                 \
    cudaError_t e_ = (call);                                           \
    if (e_ != cudaSuccess) {                                           \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_); \
      return LGR_E_CUDA;                                               \
    }                                                                  \
  } while (0)

static int fmt_ensure(lgr_fmt_ctx* c, Buf& b, size_t bytes) {
  if (bytes < 256) bytes = 256;
  if (b.cap >= bytes) return LGR_OK;
  if (b.p) cudaFree(b.p);
  b.p = nullptr, b.cap = 0;
  const size_t want = bytes + bytes / 2;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    c->err = std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e);
    return LGR_E_NOMEM;
  }
  b.cap = want;
  return LGR_OK;
}

extern "C" {

int lgr_format_create(int device_ordinal, lgr_fmt_ctx** out) {
  if (!out) return LGR_E_ARG;
  *out = nullptr;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0) {
    g_fmt_create_err = "no CUDA device (this path has no CPU fallback)";
    return LGR_E_NO_DEVICE;
  }
  if (device_ordinal < 0 || device_ordinal >= n_dev) {
    g_fmt_create_err = "device ordinal out of range";
    return LGR_E_ARG;
  }
  lgr_fmt_ctx* c = new lgr_fmt_ctx();
  c->device = device_ordinal;
  cudaError_t e = cudaSetDevice(device_ordinal);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device_ordinal);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
  if (e != cudaSuccess) {
    g_fmt_create_err = std::string("lgr_format_create: ") + cudaGetErrorString(e);
    delete c;
    return LGR_E_CUDA;
  }
  *out = c;
  return LGR_OK;
}

void lgr_format_destroy(lgr_fmt_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  for (Buf* b : {&c->b_in, &c->b_keep, &c->b_out})
    if (b->p) cudaFree(b->p);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char* lgr_format_last_error(const lgr_fmt_ctx* c) { return c ? c->err.c_str() : g_fmt_create_err.c_str(); }

int lgr_format_metrics(lgr_fmt_ctx* c, const lgr_evidence_in* in, lgr_format* out, float* ms_kernels) {
  if (!c || !in || (!out && in->n_supports > 0)) return LGR_E_ARG;
  c->err.clear();
  if (ms_kernels) *ms_kernels = 0.0f;
  const int S = in->n_supports;
  const int64_t N = in->n_evidence;
  if (S < 0 || N < 0) return c->err = "negative sizes", LGR_E_ARG;
  if (S == 0) return LGR_OK;
  int n_wide = 0;
  if (!in->sup_begin || !in->sup_n_alleles || !in->sup_variant_len || !in->sup_total_haps)
    return c->err = "missing support arrays", LGR_E_ARG;
  if (in->sup_begin[0] != 0 || in->sup_begin[S] != N) return c->err = "sup_begin must span [0, n_evidence]", LGR_E_ARG;
  for (int s = 0; s < S; ++s) {
    if (in->sup_begin[s + 1] < in->sup_begin[s]) return c->err = "sup_begin not monotone", LGR_E_ARG;
    if (in->sup_n_alleles[s] < 1) return c->err = "support with fewer than one allele", LGR_E_ARG;
    n_wide += in->sup_n_alleles[s] > LGR_FMT_MAX_ALLELES;  // not an error of the batch: see LGR_FMT_WIDE
  }
  if (N > 0 && !(in->insert_size && in->aln_start && in->aln_score && in->folded_pos && in->rname_hash && in->ref_nm &&
                 in->own_hap_nm && in->hap_id && in->allele && in->flags && in->base_qual && in->map_qual))
    return c->err = "missing evidence arrays", LGR_E_ARG;
  for (int s = 0; s < S; ++s)
    for (int64_t i = in->sup_begin[s]; i < in->sup_begin[s + 1]; ++i)
      if (in->allele[i] >= in->sup_n_alleles[s]) return c->err = "evidence allele index >= the support's allele count", LGR_E_ARG;

  FMT_CUDA(c, cudaSetDevice(c->device));
  // one packed arena: 8-byte arrays first, then 4-byte, then bytes (every section 16-byte aligned)
  auto up16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
  const size_t n = (size_t)N;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += up16(bytes);
    return o;
  };
  const size_t o_begin = take((size_t)(S + 1) * 8), o_isz = take(n * 8), o_start = take(n * 8), o_aln = take(n * 8),
               o_fold = take(n * 8), o_k = take((size_t)S * 4), o_vl = take((size_t)S * 4), o_th = take((size_t)S * 4),
               o_hash = take(n * 4), o_rnm = take(n * 4), o_onm = take(n * 4), o_hid = take(n * 4), o_al = take(n),
               o_fl = take(n), o_bq = take(n), o_mq = take(n);
  int rc;
  if ((rc = fmt_ensure(c, c->b_in, off)) != LGR_OK) return rc;
  if ((rc = fmt_ensure(c, c->b_keep, n)) != LGR_OK) return rc;
  if ((rc = fmt_ensure(c, c->b_out, (size_t)S * sizeof(lgr_format))) != LGR_OK) return rc;
  char* base = (char*)c->b_in.p;
  auto h2d = [&](size_t o, const void* src, size_t bytes) -> cudaError_t {
    return bytes ? cudaMemcpyAsync(base + o, src, bytes, cudaMemcpyHostToDevice, c->stream) : cudaSuccess;
  };
  FMT_CUDA(c, h2d(o_begin, in->sup_begin, (size_t)(S + 1) * 8));
  FMT_CUDA(c, h2d(o_k, in->sup_n_alleles, (size_t)S * 4));
  FMT_CUDA(c, h2d(o_vl, in->sup_variant_len, (size_t)S * 4));
  FMT_CUDA(c, h2d(o_th, in->sup_total_haps, (size_t)S * 4));
  FMT_CUDA(c, h2d(o_isz, in->insert_size, n * 8));
  FMT_CUDA(c, h2d(o_start, in->aln_start, n * 8));
  FMT_CUDA(c, h2d(o_aln, in->aln_score, n * 8));
  FMT_CUDA(c, h2d(o_fold, in->folded_pos, n * 8));
  FMT_CUDA(c, h2d(o_hash, in->rname_hash, n * 4));
  FMT_CUDA(c, h2d(o_rnm, in->ref_nm, n * 4));
  FMT_CUDA(c, h2d(o_onm, in->own_hap_nm, n * 4));
  FMT_CUDA(c, h2d(o_hid, in->hap_id, n * 4));
  FMT_CUDA(c, h2d(o_al, in->allele, n));
  FMT_CUDA(c, h2d(o_fl, in->flags, n));
  FMT_CUDA(c, h2d(o_bq, in->base_qual, n));
  FMT_CUDA(c, h2d(o_mq, in->map_qual, n));

  FmtDev D;
  D.e.insert_size = (const int64_t*)(base + o_isz), D.e.aln_start = (const int64_t*)(base + o_start);
  D.e.aln_score = (const double*)(base + o_aln), D.e.folded_pos = (const double*)(base + o_fold);
  D.e.rname_hash = (const uint32_t*)(base + o_hash), D.e.ref_nm = (const uint32_t*)(base + o_rnm);
  D.e.own_hap_nm = (const uint32_t*)(base + o_onm), D.e.hap_id = (const uint32_t*)(base + o_hid);
  D.e.allele = (const uint8_t*)(base + o_al), D.e.flags = (const uint8_t*)(base + o_fl);
  D.e.base_qual = (const uint8_t*)(base + o_bq), D.e.map_qual = (const uint8_t*)(base + o_mq);
  D.e.keep = (const uint8_t*)c->b_keep.p;
  D.sup_begin = (const int64_t*)(base + o_begin), D.sup_n_alleles = (const int32_t*)(base + o_k);
  D.sup_variant_len = (const int32_t*)(base + o_vl), D.sup_total_haps = (const int32_t*)(base + o_th);
  D.keep = (uint8_t*)c->b_keep.p, D.out = (lgr_format*)c->b_out.p;
  D.n_supports = S, D.n_evidence = N;

  FMT_CUDA(c, cudaEventRecord(c->ev0, c->stream));
  if (N > 0) {
    const unsigned blocks = (unsigned)((N + 255) / 256);
    k_fmt_dedup<<<blocks, 256, 0, c->stream>>>(D);
    FMT_CUDA(c, cudaGetLastError());
  }
  {
    // one CTA per (support, task); beyond 32 CTAs per SM and task the supports are grid-strided
    const int cap = c->sm_count * 32;
    const dim3 grid((unsigned)(S < cap ? S : cap), (unsigned)lgr_fmt::kNumTasks);
    k_fmt_metrics<<<grid, lgr_fmt::kThreads, 0, c->stream>>>(D);
    FMT_CUDA(c, cudaGetLastError());
  }
  FMT_CUDA(c, cudaEventRecord(c->ev1, c->stream));
  FMT_CUDA(c, cudaMemcpyAsync(out, c->b_out.p, (size_t)S * sizeof(lgr_format), cudaMemcpyDeviceToHost, c->stream));
  FMT_CUDA(c, cudaStreamSynchronize(c->stream));
  if (ms_kernels) FMT_CUDA(c, cudaEventElapsedTime(ms_kernels, c->ev0, c->ev1));
  if (n_wide > 0) {
    for (int s = 0; s < S; ++s)
      if (in->sup_n_alleles[s] > LGR_FMT_MAX_ALLELES) {
        std::memset(&out[s], 0, sizeof(lgr_format));
        out[s].n_alleles = (uint32_t)in->sup_n_alleles[s], out[s].valid = LGR_FMT_WIDE;
      }
    c->err = "some supports have more than LGR_FMT_MAX_ALLELES alleles (records flagged LGR_FMT_WIDE)";
    return LGR_E_PARTIAL;
  }
  return LGR_OK;
}

}  // extern "C"
