// lgr_kernels_unpack.cuh — device side of the packed wire format (include/lancet_gpu_realign.h,
// host side lgr_pack.h): the slab arrives in ONE copy; these two kernels derive from it every array
// the realignment kernels read.  Part of the single translation unit lgr_gpu.cu.
//   k_unpack_scan   one CTA: exclusive prefix sums over the group directory (haplotypes, reads,
//                   variants, bases, bounds-table entries, pairs, assignments, work items)
//   k_unpack_group  one CTA per group: per-sequence offsets (block scans of the lengths), name hashes,
//                   bounds table, read→group / pair / assignment offsets, the phase-A work items
//   k_unpack_decode one warp per sequence of the batch: bit planes → code bytes / Phred bytes (32
//                   coalesced bytes per step) and the sequence's share of the exception list
#ifndef LANCET2_B200_LGR_KERNELS_UNPACK_CUH_
#define LANCET2_B200_LGR_KERNELS_UNPACK_CUH_

namespace {

// block-wide exclusive scan of one int64 per thread (blockDim.x = 32 * NW); returns the exclusive
// prefix, *total = block sum.  s_w: NW + 1 int64 slots of shared memory, reusable after the call.
template <int NW>
__device__ __forceinline__ long long block_excl_scan(long long v, long long* s_w, long long* total) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long inc = v;
  for (int o = 1; o < 32; o <<= 1) {
    const long long u = __shfl_up_sync(full, inc, o);
    if (lane >= o) inc += u;
  }
  __syncthreads();  // s_w may still be read from the previous call
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    long long w = lane < NW ? s_w[lane] : 0;
    long long wi = w;
    for (int o = 1; o < 32; o <<= 1) {
      const long long u = __shfl_up_sync(full, wi, o);
      if (lane >= o) wi += u;
    }
    if (lane < NW) s_w[lane] = wi - w;
    if (lane == 31) s_w[NW] = wi;
  }
  __syncthreads();
  *total = s_w[NW];
  return s_w[warp] + inc - v;
}

__global__ void __launch_bounds__(1024) k_unpack_scan(const __grid_constant__ Dev D) {
  __shared__ long long s_w[33];
  const int G = D.n_groups;
  long long c_h = 0, c_r = 0, c_v = 0, c_hb = 0, c_rb = 0, c_vh = 0, c_p = 0, c_a = 0, c_it = 0;
  int32_t* ghb = const_cast<int32_t*>(D.grp_hap_begin);
  int32_t* grb = const_cast<int32_t*>(D.grp_read_begin);
  int32_t* gvb = const_cast<int32_t*>(D.grp_var_begin);
  for (int base = 0; base < G; base += 1024) {
    const int g = base + (int)threadIdx.x;
    long long nh = 0, nr = 0, nv = 0, hb = 0, rb = 0;
    int mid = 0;
    if (g < G) {
      const lgr_group_dir d = D.dir[g];
      nh = d.n_haps, nr = d.n_reads, nv = d.n_vars, hb = d.hap_bases, rb = d.read_bases, mid = d.mid_occ;
    }
    long long t;
    const long long e_h = block_excl_scan<32>(nh, s_w, &t); const long long t_h = t;
    const long long e_r = block_excl_scan<32>(nr, s_w, &t); const long long t_r = t;
    const long long e_v = block_excl_scan<32>(nv, s_w, &t); const long long t_v = t;
    const long long e_hb = block_excl_scan<32>(hb, s_w, &t); const long long t_hb = t;
    const long long e_rb = block_excl_scan<32>(rb, s_w, &t); const long long t_rb = t;
    const long long e_vh = block_excl_scan<32>(nv * nh, s_w, &t); const long long t_vh = t;
    const long long e_p = block_excl_scan<32>(nr * nh, s_w, &t); const long long t_p = t;
    const long long e_a = block_excl_scan<32>(nr * nv, s_w, &t); const long long t_a = t;
    const long long n_it = nh * ((nr + D.item_reads - 1) / D.item_reads);
    const long long e_it = block_excl_scan<32>(n_it, s_w, &t); const long long t_it = t;
    if (g < G) {
      ghb[g] = (int32_t)(c_h + e_h), grb[g] = (int32_t)(c_r + e_r), gvb[g] = (int32_t)(c_v + e_v);
      D.grp_hapbase[g] = c_hb + e_hb, D.grp_readbase[g] = c_rb + e_rb, D.grp_vh[g] = c_vh + e_vh;
      D.grp_pair[g] = c_p + e_p, D.grp_asg[g] = c_a + e_a, D.grp_item[g] = (int32_t)(c_it + e_it);
      const_cast<int32_t*>(D.grp_mid_req)[g] = mid > 0 ? mid : D.mid_occ_param;
    }
    c_h += t_h, c_r += t_r, c_v += t_v, c_hb += t_hb, c_rb += t_rb, c_vh += t_vh, c_p += t_p, c_a += t_a, c_it += t_it;
  }
  if (threadIdx.x == 0) {
    ghb[G] = (int32_t)c_h, grb[G] = (int32_t)c_r, gvb[G] = (int32_t)c_v;
    D.grp_hapbase[G] = c_hb, D.grp_readbase[G] = c_rb, D.grp_vh[G] = c_vh, D.grp_pair[G] = c_p, D.grp_asg[G] = c_a;
    D.grp_item[G] = (int32_t)c_it;
    // sentinels of the per-sequence offset arrays
    const_cast<int64_t*>(D.hap_off)[c_h] = c_hb;
    const_cast<int64_t*>(D.read_off)[c_r] = c_rb;
    const_cast<int64_t*>(D.var_hap_off)[c_v] = c_vh;
    const_cast<int64_t*>(D.pair_off)[c_r] = c_p;
    const_cast<int64_t*>(D.asg_off)[c_r] = c_a;
  }
}

constexpr int kUnpackThreads = 256;

// offsets of the sequences of one kind (haplotypes or reads) of one group: block scans of the lengths
constexpr uint64_t kSeqHasExc = 1ull << 55;  // descriptor flag: the sequence's group carries an exception list

// offsets of the sequences of one kind (haplotypes or reads) of one group: block scans of the lengths.
// Besides the offset arrays every sequence gets a descriptor for k_unpack_decode: the absolute slab
// offset of its first plane word (and, for reads, of its quality data, with the plane count in the top
// byte), so that the decode kernel reaches the payload bytes without walking directory → record header.
template <bool READS>
__device__ __forceinline__ void unpack_offsets(const Dev& D, const uint8_t* rec, const lgr_group_rec_hdr& hdr, uint64_t rec_off, int n_seq,
                                               int first, long long gbase, int g, long long pair0, long long asg0, int P, int V,
                                               long long* s_w) {
  const int tid = threadIdx.x;
  const int32_t* hap_len = reinterpret_cast<const int32_t*>(rec + hdr.off_hap_len);
  const uint16_t* read_len = reinterpret_cast<const uint16_t*>(rec + hdr.off_read_len);
  int64_t* out_off = const_cast<int64_t*>(READS ? D.read_off : D.hap_off) + first;
  uint64_t* out_plane = (READS ? D.read_plane : D.hap_plane) + first;
  const uint64_t exc = hdr.n_exc ? kSeqHasExc : 0;
  const uint64_t plane0 = rec_off + (READS ? hdr.off_read_planes : hdr.off_hap_planes);
  const uint64_t qual0 = rec_off + hdr.off_qual;
  const uint64_t qbits = hdr.qual_bits;
  long long carry_off = 0, carry_chk = 0;
  for (int base = 0; base < n_seq; base += kUnpackThreads) {
    const int i = base + tid;
    const int len = i < n_seq ? (READS ? (int)read_len[i] : hap_len[i]) : 0;
    long long t_len, t_chk;
    const long long ex_len = block_excl_scan<kUnpackThreads / 32>(len, s_w, &t_len);
    const long long ex_chk = block_excl_scan<kUnpackThreads / 32>((len + 31) >> 5, s_w, &t_chk);
    if (i < n_seq) {
      const long long off = carry_off + ex_len, chk = carry_chk + ex_chk;
      out_off[i] = gbase + off;
      out_plane[i] = (plane0 + 8ull * (uint64_t)chk) | exc;
      if (READS) {
        const int r = first + i;
        D.read_qoff[r] = (qbits == 8 ? qual0 + (uint64_t)off : qual0 + 4ull * qbits * (uint64_t)chk) | qbits << 56;
        const_cast<int32_t*>(D.read_grp)[r] = g;
        const_cast<int64_t*>(D.pair_off)[r] = pair0 + (long long)i * P;
        const_cast<int64_t*>(D.asg_off)[r] = asg0 + (long long)i * V;
        const_cast<uint32_t*>(D.name_hash)[r] = reinterpret_cast<const uint32_t*>(rec + hdr.off_name_hash)[i];
      } else {
        const_cast<int32_t*>(D.hap_grp)[first + i] = g;
      }
    }
    carry_off += t_len, carry_chk += t_chk;
  }
}

// one CTA per group: everything but the sequence bytes
__global__ void __launch_bounds__(kUnpackThreads) k_unpack_group(const __grid_constant__ Dev D) {
  __shared__ long long s_w[kUnpackThreads / 32 + 1];
  const int tid = threadIdx.x;
  for (int g = blockIdx.x; g < D.n_groups; g += gridDim.x) {
    const lgr_group_dir d = D.dir[g];
    const uint8_t* rec = D.slab + d.rec_off;
    const lgr_group_rec_hdr hdr = *reinterpret_cast<const lgr_group_rec_hdr*>(rec);
    const int P = d.n_haps, R = d.n_reads, V = d.n_vars;
    const int hb = D.grp_hap_begin[g], rb = D.grp_read_begin[g], vb = D.grp_var_begin[g];
    const long long hbase = D.grp_hapbase[g], rbase = D.grp_readbase[g], vh0 = D.grp_vh[g];
    const long long pair0 = D.grp_pair[g], asg0 = D.grp_asg[g];
    unpack_offsets<false>(D, rec, hdr, d.rec_off, P, hb, hbase, g, 0, 0, P, V, s_w);
    unpack_offsets<true>(D, rec, hdr, d.rec_off, R, rb, rbase, g, pair0, asg0, P, V, s_w);
    if (tid < 4) D.grp_lut[4 * (size_t)g + tid] = reinterpret_cast<const uint32_t*>(hdr.qual_lut)[tid];
    // ExtractHapBounds' dense table
    {
      const long long vh = (long long)V * P;
      const int32_t* vs = reinterpret_cast<const int32_t*>(rec + hdr.off_var);
      const int32_t* vl = vs + vh;
      const int8_t* va = reinterpret_cast<const int8_t*>(vl + vh);
      for (long long x = tid; x < vh; x += kUnpackThreads) {
        const_cast<int32_t*>(D.var_start)[vh0 + x] = vs[x];
        const_cast<int32_t*>(D.var_len)[vh0 + x] = vl[x];
        const_cast<int8_t*>(D.var_allele)[vh0 + x] = va[x];
      }
      for (int v = tid; v < V; v += kUnpackThreads) const_cast<int64_t*>(D.var_hap_off)[vb + v] = vh0 + (long long)v * P;
    }
    // phase-A work items: (haplotype, first read, #reads) in haplotype-major order
    {
      const int per_hap = (R + D.item_reads - 1) / D.item_reads;
      const int it0 = D.grp_item[g];
      for (int x = tid; x < P * per_hap; x += kUnpackThreads) {
        const int h = x / per_hap, k = x - h * per_hap;
        const int r0 = k * D.item_reads;
        const_cast<int32_t*>(D.item_hap)[it0 + x] = hb + h;
        const_cast<int32_t*>(D.item_r0)[it0 + x] = rb + r0;
        const_cast<int32_t*>(D.item_n)[it0 + x] = R - r0 < D.item_reads ? R - r0 : D.item_reads;
      }
    }
    __syncthreads();
  }
}

// one WARP per sequence of the whole batch (haplotypes first, then reads): bit planes → code bytes
// (and Phred bytes for reads), 32 coalesced bytes per step, then the sequence's share of the exception
// list.  Flat over the batch and fed by per-sequence descriptors (two independent loads, then the
// payload), so the copy does not depend on group sizes or on a chain of header lookups.
__global__ void __launch_bounds__(kUnpackThreads) k_unpack_decode(const __grid_constant__ Dev D) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const long long n_seq = (long long)D.n_haps + D.n_reads;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < n_seq; q += warps) {
    const bool is_read = q >= D.n_haps;
    const int idx = is_read ? (int)(q - D.n_haps) : (int)q;
    const int64_t* offs = is_read ? D.read_off : D.hap_off;
    const long long o = offs[idx];
    const int l = (int)(offs[idx + 1] - o);
    const uint64_t pd = is_read ? D.read_plane[idx] : D.hap_plane[idx];
    const uint64_t qd = is_read ? D.read_qoff[idx] : 0;
    const uint32_t* planes = reinterpret_cast<const uint32_t*>(D.slab + (pd & ~kSeqHasExc));
    uint8_t* codes = (is_read ? D.read_codes : D.hap_codes) + o;
    for (int c = 0; (c << 5) < l; ++c) {
      const uint2 w = *reinterpret_cast<const uint2*>(planes + 2 * c);  // (low-bit plane, high-bit plane), 8-byte aligned
      const int b = (c << 5) + lane;
      if (b < l) codes[b] = (uint8_t)(((w.x >> lane & 1u) | (w.y >> lane & 1u) << 1) * 0x11u);
    }
    if (is_read) {
      const int qbits = (int)(qd >> 56);
      const uint8_t* qsrc = D.slab + (qd & ((1ull << 55) - 1));
      uint8_t* quals = const_cast<uint8_t*>(D.read_quals) + o;
      if (qbits == 8) {
        for (int b = lane; b < l; b += 32) quals[b] = qsrc[b];
      } else {
        const int g = D.read_grp[idx];
        const uint32_t lutw = D.grp_lut[4 * (size_t)g + (lane >> 2 & 3)];
        const uint32_t lut_lo = lutw >> (8 * (lane & 3)) & 0xffu;  // dictionary entry `lane` (< 16) lives in lane `lane`
        const uint32_t* qplanes = reinterpret_cast<const uint32_t*>(qsrc);
        // (the two plane counts spelled out: as one loop over a run-time plane count this line was half of the
        // kernel's instructions)
        if (qbits == 2) {
          for (int c = 0; (c << 5) < l; ++c) {
            const uint2 w = *reinterpret_cast<const uint2*>(qplanes + 2 * c);  // plane words are 8-byte aligned in the record
            const uint32_t k = (w.x >> lane & 1u) | (w.y >> lane & 1u) << 1;
            const uint32_t v = __shfl_sync(full, lut_lo, (int)k);
            const int b = (c << 5) + lane;
            if (b < l) quals[b] = (uint8_t)v;
          }
        } else {
          for (int c = 0; (c << 5) < l; ++c) {
            const uint2 w0 = *reinterpret_cast<const uint2*>(qplanes + 4 * c);  // 8-byte, not 16-byte, alignment is guaranteed
            const uint2 w1 = *reinterpret_cast<const uint2*>(qplanes + 4 * c + 2);
            const uint32_t k = (w0.x >> lane & 1u) | (w0.y >> lane & 1u) << 1 | (w1.x >> lane & 1u) << 2 | (w1.y >> lane & 1u) << 3;
            const uint32_t v = __shfl_sync(full, lut_lo, (int)k);
            const int b = (c << 5) + lane;
            if (b < l) quals[b] = (uint8_t)v;
          }
        }
      }
    }
    if (pd & kSeqHasExc) {  // exceptions (N, IUPAC, U): the entries of the group that fall inside this sequence
      __syncwarp();
      const int g = is_read ? D.read_grp[idx] : D.hap_grp[idx];
      const uint8_t* rec = D.slab + D.dir[g].rec_off;
      const lgr_group_rec_hdr* hdr = reinterpret_cast<const lgr_group_rec_hdr*>(rec);
      const uint32_t n_exc = hdr->n_exc;
      const uint32_t* epos = reinterpret_cast<const uint32_t*>(rec + hdr->off_exc);
      const uint8_t* ecode = reinterpret_cast<const uint8_t*>(epos + n_exc);
      const long long start = o - (is_read ? D.grp_readbase[g] : D.grp_hapbase[g]);
      for (uint32_t e = lane; e < n_exc; e += 32) {
        const uint32_t p = epos[e];
        const long long pos = p & 0x7fffffffu;
        if ((p >> 31) == (is_read ? 1u : 0u) && pos >= start && pos < start + l) codes[pos - start] = ecode[e];
      }
    }
  }
}

}  // namespace

#endif  // LANCET2_B200_LGR_KERNELS_UNPACK_CUH_
