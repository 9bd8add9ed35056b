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


// ---------------------------------------------------------------------------------------------
// k_evidence_*: AddToTable on the device.  From the resident
// lgr_assign records of a realignment batch and a few per-read columns, build the SoA evidence of
// every (variant, sample) support exactly as lancet_gpu::EvidenceColumns::AppendJob does on the host
// (genotyper.cpp:423-456): supports in (group, variant, sample-first-seen) order, records in read
// order.  The realignment's X31 name hash stands in for absl::HashOf(qname): the dedup only asks
// whether two records of a support carry the same name.
// ---------------------------------------------------------------------------------------------
constexpr int kEvMaxSamples = 32;

struct EvDev {
  const lgr_assign* assign;       // resident records of the realignment batch
  const int32_t* grp_read_begin;  // [G+1]
  const int32_t* grp_var_begin;   // [G+1]
  const int64_t* grp_asg_begin;   // [G+1] first assign record of the group
  const int32_t* grp_n_haps;      // [G]
  const int32_t* var_grp;         // [NV]
  const int32_t* var_n_alleles;   // [NV]
  const int32_t* var_len;         // [NV]
  const int64_t* r_insert;        // per read
  const int64_t* r_start;
  const uint32_t* r_hash;
  const int32_t* r_sample;
  const uint16_t* r_flag;
  const uint8_t* r_mapq;
  const uint8_t* r_soft;
  int32_t n_vars, n_samples;
  // slots: variant v owns [v * n_samples, (v+1) * n_samples), filled in sample-first-seen order
  int32_t* slot_cnt;
  int32_t* slot_sample;
  int64_t* slot_begin;
  long long* totals;  // [0] supports, [1] evidence records, [2] error flags
  // outputs: the columns of lgr_evidence_in + the support tables
  int64_t* sup_begin;
  int32_t* sup_n_alleles;
  int32_t* sup_variant_len;
  int32_t* sup_total_haps;
  int32_t* sup_key;  // [3 S]: group, variant, sample id
  int64_t* o_insert;
  int64_t* o_start;
  double* o_aln;
  double* o_fold;
  uint32_t* o_hash;
  uint32_t* o_rnm;
  uint32_t* o_onm;
  uint32_t* o_hid;
  uint8_t* o_allele;
  uint8_t* o_flags;
  uint8_t* o_bq;
  uint8_t* o_mq;
};

// one CTA per variant: per-sample record counts and the first read that names each sample
__global__ void __launch_bounds__(128) k_evidence_count(const __grid_constant__ EvDev D) {
  __shared__ int s_cnt[kEvMaxSamples];
  __shared__ int s_first[kEvMaxSamples];
  for (int v = blockIdx.x; v < D.n_vars; v += gridDim.x) {
    const int g = D.var_grp[v];
    const int r0 = D.grp_read_begin[g], R = D.grp_read_begin[g + 1] - r0;
    const int V = D.grp_var_begin[g + 1] - D.grp_var_begin[g], vi = v - D.grp_var_begin[g];
    const lgr_assign* a = D.assign + D.grp_asg_begin[g] + vi;
    __syncthreads();
    if (threadIdx.x < kEvMaxSamples) s_cnt[threadIdx.x] = 0, s_first[threadIdx.x] = 0x7fffffff;
    __syncthreads();
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
      if (!a[(int64_t)r * V].assigned) continue;
      const int sm = D.r_sample[r0 + r];
      atomicAdd(&s_cnt[sm], 1);
      atomicMin(&s_first[sm], r);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      // FindOrCreate's creation order: samples by the first read that contributes to this variant
      int order[kEvMaxSamples], n = 0;
      for (int sm = 0; sm < D.n_samples; ++sm)
        if (s_cnt[sm] > 0) {
          int at = n++;
          while (at > 0 && s_first[order[at - 1]] > s_first[sm]) order[at] = order[at - 1], --at;
          order[at] = sm;
        }
      for (int k = 0; k < D.n_samples; ++k) {
        D.slot_cnt[(int64_t)v * D.n_samples + k] = k < n ? s_cnt[order[k]] : 0;
        D.slot_sample[(int64_t)v * D.n_samples + k] = k < n ? order[k] : -1;
      }
    }
  }
}

// one CTA: exclusive scans over the slots → support numbers and evidence ranges, support tables
__global__ void __launch_bounds__(1024) k_evidence_scan(const __grid_constant__ EvDev D) {
  __shared__ long long s_ev[32];
  __shared__ int s_sup[32];
  __shared__ long long s_carry_ev;
  __shared__ int s_carry_sup;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n_slots = (int64_t)D.n_vars * D.n_samples;
  if (threadIdx.x == 0) s_carry_ev = 0, s_carry_sup = 0;
  __syncthreads();
  for (int64_t base = 0; base < n_slots; base += blockDim.x) {
    const int64_t slot = base + threadIdx.x;
    const int cnt = slot < n_slots ? D.slot_cnt[slot] : 0;
    long long ev = cnt;
    int sup = cnt > 0;
    for (int o = 1; o < 32; o <<= 1) {
      const long long e2 = __shfl_up_sync(0xffffffffu, ev, o);
      const int s2 = __shfl_up_sync(0xffffffffu, sup, o);
      if (lane >= o) ev += e2, sup += s2;
    }
    if (lane == 31) s_ev[warp] = ev, s_sup[warp] = sup;
    __syncthreads();
    if (warp == 0) {
      long long e = s_ev[lane];
      int u = s_sup[lane];
      for (int o = 1; o < 32; o <<= 1) {
        const long long e2 = __shfl_up_sync(0xffffffffu, e, o);
        const int u2 = __shfl_up_sync(0xffffffffu, u, o);
        if (lane >= o) e += e2, u += u2;
      }
      s_ev[lane] = e, s_sup[lane] = u;
    }
    __syncthreads();
    const long long ev_excl = s_carry_ev + (warp > 0 ? s_ev[warp - 1] : 0) + ev - cnt;
    const int sup_excl = s_carry_sup + (warp > 0 ? s_sup[warp - 1] : 0) + sup - (cnt > 0);
    if (slot < n_slots) {
      D.slot_begin[slot] = ev_excl;
      if (cnt > 0) {
        const int v = (int)(slot / D.n_samples), g = D.var_grp[v];
        D.sup_begin[sup_excl] = ev_excl;
        D.sup_n_alleles[sup_excl] = D.var_n_alleles[v];
        D.sup_variant_len[sup_excl] = D.var_len[v];
        D.sup_total_haps[sup_excl] = D.grp_n_haps[g];
        D.sup_key[3 * sup_excl] = g, D.sup_key[3 * sup_excl + 1] = v, D.sup_key[3 * sup_excl + 2] = D.slot_sample[slot];
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) s_carry_ev += s_ev[31], s_carry_sup += s_sup[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    D.sup_begin[s_carry_sup] = s_carry_ev;
    D.totals[0] = s_carry_sup, D.totals[1] = s_carry_ev;
  }
}

// one warp per slot: the support's records in read order (ballot compaction), every column of ReadEvidence
__global__ void __launch_bounds__(128) k_evidence_scatter(const __grid_constant__ EvDev D) {
  const int lane = threadIdx.x & 31;
  const int64_t n_slots = (int64_t)D.n_vars * D.n_samples;
  for (int64_t slot = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5); slot < n_slots; slot += (int64_t)gridDim.x * 4) {
    if (D.slot_cnt[slot] == 0) continue;
    const int v = (int)(slot / D.n_samples), sm = D.slot_sample[slot], g = D.var_grp[v];
    const int r0 = D.grp_read_begin[g], R = D.grp_read_begin[g + 1] - r0;
    const int V = D.grp_var_begin[g + 1] - D.grp_var_begin[g], vi = v - D.grp_var_begin[g];
    const int K = D.var_n_alleles[v];
    const lgr_assign* ab = D.assign + D.grp_asg_begin[g] + vi;
    int64_t at = D.slot_begin[slot];
    for (int base = 0; base < R; base += 32) {
      const int r = base + lane;
      lgr_assign a;
      bool take = false;
      if (r < R && D.r_sample[r0 + r] == sm) {
        a = ab[(int64_t)r * V];
        take = a.assigned != 0;
      }
      const unsigned m = __ballot_sync(0xffffffffu, take);
      if (take) {
        const int64_t o = at + __popc(m & ((1u << lane) - 1u));
        const int rr = r0 + r;
        D.o_insert[o] = D.r_insert[rr], D.o_start[o] = D.r_start[rr];
        D.o_aln[o] = (double)a.global_score + (a.local_score * a.local_identity);  // CombinedScore(); -fmad=false: as written
        D.o_fold[o] = a.folded_read_pos;
        D.o_hash[o] = D.r_hash[rr], D.o_rnm[o] = a.ref_nm, D.o_onm[o] = a.own_hap_nm, D.o_hid[o] = a.hap_id;
        D.o_allele[o] = (uint8_t)a.allele;
        const unsigned fl = D.r_flag[rr];
        D.o_flags[o] = (uint8_t)(((fl & 0x10u) ? LGR_EV_REV : 0u) | (D.r_soft[rr] ? LGR_EV_SOFTCLIP : 0u) | ((fl & 0x2u) ? LGR_EV_PROPER_PAIR : 0u));
        D.o_bq[o] = a.base_qual, D.o_mq[o] = D.r_mapq[rr];
        if (a.allele < 0 || (int)a.allele >= K) atomicOr((unsigned long long*)&D.totals[2], 1ull);
      }
      at += __popc(m);
    }
  }
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
  Buf b_ev, b_key;           // lgr_format_from_assign: per-read / per-variant inputs + slots, support keys
  lgr_evidence_in last_ev{};  // device pointers of the columns the last call worked on (lgr_format_debug_evidence)
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

// AddToTable + FORMAT math without the host round trip: evidence columns are built on the device from the
// realignment's resident lgr_assign records (k_evidence_*), then k_fmt_dedup / k_fmt_metrics run on them.
int lgr_format_from_assign(lgr_fmt_ctx* c, const lgr_assign_batch* in, lgr_format* out, int32_t out_cap, int32_t* sup_key,
                           int32_t* n_supports, float* ms_kernels) {
  if (!c || !in || !n_supports) return LGR_E_ARG;
  c->err.clear();
  *n_supports = 0;
  if (ms_kernels) *ms_kernels = 0.0f;
  const int G = in->n_groups, NR = in->n_reads, NV = in->n_vars, NS = in->n_samples;
  if (G < 0 || NR < 0 || NV < 0 || in->n_assign < 0) return c->err = "negative sizes", LGR_E_ARG;
  if (NS < 1 || NS > kEvMaxSamples) return c->err = "n_samples must be in [1, 32]", LGR_E_LIMIT;
  if (G == 0 || NV == 0 || in->n_assign == 0) return LGR_OK;
  if (!in->dev_assign && !in->host_assign) return c->err = "neither dev_assign nor host_assign", LGR_E_ARG;
  if (!in->grp_read_begin || !in->grp_var_begin || !in->grp_n_haps || !in->var_n_alleles || !in->var_len || !in->read_insert_size ||
      !in->read_aln_start || !in->read_name_hash || !in->read_sample || !in->read_sam_flag || !in->read_map_qual ||
      !in->read_soft_clipped)
    return c->err = "missing input arrays", LGR_E_ARG;
  if (in->grp_read_begin[0] != 0 || in->grp_var_begin[0] != 0 || in->grp_read_begin[G] != NR || in->grp_var_begin[G] != NV)
    return c->err = "group prefix arrays do not span the reads / variants", LGR_E_ARG;
  std::vector<int64_t> asg_begin((size_t)G + 1, 0);
  std::vector<int32_t> var_grp((size_t)NV);
  for (int g = 0; g < G; ++g) {
    const int R = in->grp_read_begin[g + 1] - in->grp_read_begin[g], V = in->grp_var_begin[g + 1] - in->grp_var_begin[g];
    if (R < 0 || V < 0) return c->err = "group prefix arrays must be non-decreasing", LGR_E_ARG;
    asg_begin[(size_t)g + 1] = asg_begin[(size_t)g] + (int64_t)R * V;
    for (int v = in->grp_var_begin[g]; v < in->grp_var_begin[g + 1]; ++v) var_grp[(size_t)v] = g;
  }
  if (asg_begin[(size_t)G] != in->n_assign) return c->err = "n_assign is not the sum of reads x variants over the groups", LGR_E_ARG;
  for (int r = 0; r < NR; ++r)
    if (in->read_sample[r] < 0 || in->read_sample[r] >= NS) return c->err = "read_sample out of range", LGR_E_ARG;
  for (int v = 0; v < NV; ++v)
    if (in->var_n_alleles[v] < 1) return c->err = "variant with fewer than one allele", LGR_E_ARG;
  const int64_t n_slots = (int64_t)NV * NS;
  if (n_slots > out_cap && (!out || !sup_key)) return c->err = "out / sup_key missing", LGR_E_ARG;

  FMT_CUDA(c, cudaSetDevice(c->device));
  auto up16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += up16(bytes);
    return o;
  };
  // inputs + slots
  const size_t nr = (size_t)NR, nv = (size_t)NV, ng = (size_t)G, na = (size_t)in->n_assign;
  const size_t i_asg = take(in->dev_assign ? 0 : na * sizeof(lgr_assign)), i_grb = take((ng + 1) * 4), i_gvb = take((ng + 1) * 4),
               i_gab = take((ng + 1) * 8), i_gnh = take(ng * 4), i_vg = take(nv * 4), i_vk = take(nv * 4), i_vl = take(nv * 4),
               i_ins = take(nr * 8), i_st = take(nr * 8), i_hash = take(nr * 4), i_smp = take(nr * 4), i_flag = take(nr * 2),
               i_mq = take(nr), i_sc = take(nr), i_scnt = take((size_t)n_slots * 4), i_ssmp = take((size_t)n_slots * 4),
               i_sbeg = take((size_t)n_slots * 8), i_tot = take(3 * 8);
  int rc;
  if ((rc = fmt_ensure(c, c->b_ev, off)) != LGR_OK) return rc;
  char* eb = (char*)c->b_ev.p;
  // columns (capacity: one record per assign record, one support per slot)
  off = 0;
  const size_t S_cap = (size_t)n_slots;
  const size_t o_begin = take((S_cap + 1) * 8), o_isz = take(na * 8), o_start = take(na * 8), o_aln = take(na * 8), o_fold = take(na * 8),
               o_k = take(S_cap * 4), o_vl = take(S_cap * 4), o_th = take(S_cap * 4), o_hash = take(na * 4), o_rnm = take(na * 4),
               o_onm = take(na * 4), o_hid = take(na * 4), o_al = take(na), o_fl = take(na), o_bq = take(na), o_mq = take(na);
  if ((rc = fmt_ensure(c, c->b_in, off)) != LGR_OK) return rc;
  if ((rc = fmt_ensure(c, c->b_keep, na)) != LGR_OK) return rc;
  if ((rc = fmt_ensure(c, c->b_out, S_cap * sizeof(lgr_format))) != LGR_OK) return rc;
  if ((rc = fmt_ensure(c, c->b_key, S_cap * 3 * sizeof(int32_t))) != LGR_OK) return rc;
  char* base = (char*)c->b_in.p;
  auto h2d = [&](size_t o, const void* src, size_t bytes) -> cudaError_t {
    return bytes ? cudaMemcpyAsync(eb + o, src, bytes, cudaMemcpyHostToDevice, c->stream) : cudaSuccess;
  };
  if (!in->dev_assign) FMT_CUDA(c, h2d(i_asg, in->host_assign, na * sizeof(lgr_assign)));
  FMT_CUDA(c, h2d(i_grb, in->grp_read_begin, (ng + 1) * 4));
  FMT_CUDA(c, h2d(i_gvb, in->grp_var_begin, (ng + 1) * 4));
  FMT_CUDA(c, h2d(i_gab, asg_begin.data(), (ng + 1) * 8));
  FMT_CUDA(c, h2d(i_gnh, in->grp_n_haps, ng * 4));
  FMT_CUDA(c, h2d(i_vg, var_grp.data(), nv * 4));
  FMT_CUDA(c, h2d(i_vk, in->var_n_alleles, nv * 4));
  FMT_CUDA(c, h2d(i_vl, in->var_len, nv * 4));
  FMT_CUDA(c, h2d(i_ins, in->read_insert_size, nr * 8));
  FMT_CUDA(c, h2d(i_st, in->read_aln_start, nr * 8));
  FMT_CUDA(c, h2d(i_hash, in->read_name_hash, nr * 4));
  FMT_CUDA(c, h2d(i_smp, in->read_sample, nr * 4));
  FMT_CUDA(c, h2d(i_flag, in->read_sam_flag, nr * 2));
  FMT_CUDA(c, h2d(i_mq, in->read_map_qual, nr));
  FMT_CUDA(c, h2d(i_sc, in->read_soft_clipped, nr));
  FMT_CUDA(c, cudaMemsetAsync(eb + i_tot, 0, 3 * 8, c->stream));

  EvDev E;
  E.assign = in->dev_assign ? in->dev_assign : (const lgr_assign*)(eb + i_asg);
  E.grp_read_begin = (const int32_t*)(eb + i_grb), E.grp_var_begin = (const int32_t*)(eb + i_gvb);
  E.grp_asg_begin = (const int64_t*)(eb + i_gab), E.grp_n_haps = (const int32_t*)(eb + i_gnh);
  E.var_grp = (const int32_t*)(eb + i_vg), E.var_n_alleles = (const int32_t*)(eb + i_vk), E.var_len = (const int32_t*)(eb + i_vl);
  E.r_insert = (const int64_t*)(eb + i_ins), E.r_start = (const int64_t*)(eb + i_st), E.r_hash = (const uint32_t*)(eb + i_hash);
  E.r_sample = (const int32_t*)(eb + i_smp), E.r_flag = (const uint16_t*)(eb + i_flag), E.r_mapq = (const uint8_t*)(eb + i_mq);
  E.r_soft = (const uint8_t*)(eb + i_sc);
  E.n_vars = NV, E.n_samples = NS;
  E.slot_cnt = (int32_t*)(eb + i_scnt), E.slot_sample = (int32_t*)(eb + i_ssmp), E.slot_begin = (int64_t*)(eb + i_sbeg);
  E.totals = (long long*)(eb + i_tot);
  E.sup_begin = (int64_t*)(base + o_begin), E.sup_n_alleles = (int32_t*)(base + o_k), E.sup_variant_len = (int32_t*)(base + o_vl);
  E.sup_total_haps = (int32_t*)(base + o_th), E.sup_key = (int32_t*)c->b_key.p;
  E.o_insert = (int64_t*)(base + o_isz), E.o_start = (int64_t*)(base + o_start), E.o_aln = (double*)(base + o_aln);
  E.o_fold = (double*)(base + o_fold), E.o_hash = (uint32_t*)(base + o_hash), E.o_rnm = (uint32_t*)(base + o_rnm);
  E.o_onm = (uint32_t*)(base + o_onm), E.o_hid = (uint32_t*)(base + o_hid), E.o_allele = (uint8_t*)(base + o_al);
  E.o_flags = (uint8_t*)(base + o_fl), E.o_bq = (uint8_t*)(base + o_bq), E.o_mq = (uint8_t*)(base + o_mq);

  FMT_CUDA(c, cudaEventRecord(c->ev0, c->stream));
  k_evidence_count<<<NV < c->sm_count * 16 ? NV : c->sm_count * 16, 128, 0, c->stream>>>(E);
  k_evidence_scan<<<1, 1024, 0, c->stream>>>(E);
  {
    const int64_t want = (n_slots + 3) / 4;
    k_evidence_scatter<<<(unsigned)(want < c->sm_count * 16 ? want : c->sm_count * 16), 128, 0, c->stream>>>(E);
  }
  FMT_CUDA(c, cudaGetLastError());
  long long totals[3] = {0, 0, 0};
  FMT_CUDA(c, cudaMemcpyAsync(totals, eb + i_tot, sizeof(totals), cudaMemcpyDeviceToHost, c->stream));
  FMT_CUDA(c, cudaStreamSynchronize(c->stream));
  const int S = (int)totals[0];
  const int64_t N = totals[1];
  if (totals[2] != 0) return c->err = "an assignment names an allele outside its variant's allele count", LGR_E_ARG;
  *n_supports = S;
  if (S > out_cap) return c->err = "out / sup_key too small for the supports of this batch", LGR_E_ARG;

  FmtDev D;
  D.e.insert_size = E.o_insert, D.e.aln_start = E.o_start, D.e.aln_score = E.o_aln, D.e.folded_pos = E.o_fold;
  D.e.rname_hash = E.o_hash, D.e.ref_nm = E.o_rnm, D.e.own_hap_nm = E.o_onm, D.e.hap_id = E.o_hid;
  D.e.allele = E.o_allele, D.e.flags = E.o_flags, D.e.base_qual = E.o_bq, D.e.map_qual = E.o_mq;
  D.e.keep = (const uint8_t*)c->b_keep.p;
  D.sup_begin = E.sup_begin, D.sup_n_alleles = E.sup_n_alleles, D.sup_variant_len = E.sup_variant_len, D.sup_total_haps = E.sup_total_haps;
  D.keep = (uint8_t*)c->b_keep.p, D.out = (lgr_format*)c->b_out.p;
  D.n_supports = S, D.n_evidence = N;
  c->last_ev = lgr_evidence_in{};
  c->last_ev.n_supports = S, c->last_ev.n_evidence = N;
  c->last_ev.sup_begin = E.sup_begin, c->last_ev.sup_n_alleles = E.sup_n_alleles, c->last_ev.sup_variant_len = E.sup_variant_len;
  c->last_ev.sup_total_haps = E.sup_total_haps, c->last_ev.insert_size = E.o_insert, c->last_ev.aln_start = E.o_start;
  c->last_ev.aln_score = E.o_aln, c->last_ev.folded_pos = E.o_fold, c->last_ev.rname_hash = E.o_hash, c->last_ev.ref_nm = E.o_rnm;
  c->last_ev.own_hap_nm = E.o_onm, c->last_ev.hap_id = E.o_hid, c->last_ev.allele = E.o_allele, c->last_ev.flags = E.o_flags;
  c->last_ev.base_qual = E.o_bq, c->last_ev.map_qual = E.o_mq;
  if (S > 0) {
    if (N > 0) {
      k_fmt_dedup<<<(unsigned)((N + 255) / 256), 256, 0, c->stream>>>(D);
      FMT_CUDA(c, cudaGetLastError());
    }
    const int cap = c->sm_count * 32;
    const dim3 grid((unsigned)(S < cap ? S : cap), (unsigned)lgr_fmt::kNumTasks);
    k_fmt_metrics<<<grid, lgr_fmt::kThreads, 0, c->stream>>>(D);
    FMT_CUDA(c, cudaGetLastError());
  }
  FMT_CUDA(c, cudaEventRecord(c->ev1, c->stream));
  if (S > 0) {
    FMT_CUDA(c, cudaMemcpyAsync(out, c->b_out.p, (size_t)S * sizeof(lgr_format), cudaMemcpyDeviceToHost, c->stream));
    FMT_CUDA(c, cudaMemcpyAsync(sup_key, c->b_key.p, (size_t)S * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  }
  FMT_CUDA(c, cudaStreamSynchronize(c->stream));
  if (ms_kernels) FMT_CUDA(c, cudaEventElapsedTime(ms_kernels, c->ev0, c->ev1));
  int n_wide = 0;
  for (int s = 0; s < S; ++s) {
    const int K = in->var_n_alleles[sup_key[3 * s + 1]];
    if (K > LGR_FMT_MAX_ALLELES) {
      std::memset(&out[s], 0, sizeof(lgr_format));
      out[s].n_alleles = (uint32_t)K, out[s].valid = LGR_FMT_WIDE;
      ++n_wide;
    }
  }
  if (n_wide > 0) {
    c->err = "some supports have more than LGR_FMT_MAX_ALLELES alleles (records flagged LGR_FMT_WIDE)";
    return LGR_E_PARTIAL;
  }
  return LGR_OK;
}

// Test hook: copy the evidence columns the last lgr_format_from_assign built on the device into caller buffers
// (dst's pointers are DESTINATIONS with room for n_evidence / n_supports entries; NULL ones are skipped).
int lgr_format_debug_evidence(lgr_fmt_ctx* c, lgr_evidence_in* dst) {
  if (!c || !dst) return LGR_E_ARG;
  const lgr_evidence_in& L = c->last_ev;
  const size_t n = (size_t)L.n_evidence, S = (size_t)L.n_supports;
  dst->n_supports = L.n_supports, dst->n_evidence = L.n_evidence;
  FMT_CUDA(c, cudaSetDevice(c->device));
  auto cp = [&](const void* d, const void* s, size_t bytes) -> cudaError_t {
    return d && s && bytes ? cudaMemcpy(const_cast<void*>(d), s, bytes, cudaMemcpyDeviceToHost) : cudaSuccess;
  };
  if (S > 0) {
    FMT_CUDA(c, cp(dst->sup_begin, L.sup_begin, (S + 1) * 8));
    FMT_CUDA(c, cp(dst->sup_n_alleles, L.sup_n_alleles, S * 4));
    FMT_CUDA(c, cp(dst->sup_variant_len, L.sup_variant_len, S * 4));
    FMT_CUDA(c, cp(dst->sup_total_haps, L.sup_total_haps, S * 4));
  }
  FMT_CUDA(c, cp(dst->insert_size, L.insert_size, n * 8));
  FMT_CUDA(c, cp(dst->aln_start, L.aln_start, n * 8));
  FMT_CUDA(c, cp(dst->aln_score, L.aln_score, n * 8));
  FMT_CUDA(c, cp(dst->folded_pos, L.folded_pos, n * 8));
  FMT_CUDA(c, cp(dst->rname_hash, L.rname_hash, n * 4));
  FMT_CUDA(c, cp(dst->ref_nm, L.ref_nm, n * 4));
  FMT_CUDA(c, cp(dst->own_hap_nm, L.own_hap_nm, n * 4));
  FMT_CUDA(c, cp(dst->hap_id, L.hap_id, n * 4));
  FMT_CUDA(c, cp(dst->allele, L.allele, n));
  FMT_CUDA(c, cp(dst->flags, L.flags, n));
  FMT_CUDA(c, cp(dst->base_qual, L.base_qual, n));
  FMT_CUDA(c, cp(dst->map_qual, L.map_qual, n));
  return LGR_OK;
}

}  // extern "C"
