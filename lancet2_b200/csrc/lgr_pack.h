// lgr_pack.h — host side of the packed wire format (include/lancet_gpu_realign.h, "Packed wire
// format"): one Genotype() payload → one self-describing group record of bit planes, written by the
// thread that owns the payload straight into the pinned slab the batch is copied from.
// Pure host code (no CUDA call); header-only so that the C++ adapter inlines it.  The device side is
// lgr_kernels_unpack.cuh; tests/test_pack_format.py holds an independent numpy decoder.
//
// What is packed is what Genotyper::ResetData / AlignToAllHaplotypes hand to minimap2 and what
// AssignReadToAlleles reads (reference: src/lancet/caller/genotyper.cpp:243-267, 376-411): haplotype
// strings, read SeqPtr()/QualPtr()/Length(), the X31 hash of QnamePtr(), and ExtractHapBounds' table.
#ifndef LANCET2_B200_LGR_PACK_H_
#define LANCET2_B200_LGR_PACK_H_

#include <cstdint>
#include <cstring>

#include "../../include/lancet_gpu_realign.h"

#if defined(__x86_64__) && (defined(__GNUC__) || defined(__clang__)) && !defined(__CUDA_ARCH__)
#include <immintrin.h>
#define LGR_PACK_X86 1
#else
#define LGR_PACK_X86 0
#endif

namespace lgr_pack {

// code byte of one base: low nibble minimap2 nt4, high nibble Lancet2 ENCODE_TABLE
// (scoring_constants.h:48-74; 'U' is T for minimap2 and N for Lancet2).  Same table as
// lgr::encode_base in lgr_core.cuh (tests/test_pack_format.py checks the two against each other).
inline std::uint8_t code_of(std::uint8_t c) {
  switch (c) {
    case 'A': case 'a': return 0x00;
    case 'C': case 'c': return 0x11;
    case 'G': case 'g': return 0x22;
    case 'T': case 't': return 0x33;
    case 'U': case 'u': return 0x43;
    default: return 0x44;
  }
}

inline bool is_acgt(std::uint8_t c) {
  c &= 0xDF;
  return c == 'A' || c == 'C' || c == 'G' || c == 'T';
}

inline std::size_t align_up(std::size_t x, std::size_t a) { return (x + a - 1) / a * a; }
inline int chunks_of(int len) { return (len + 31) >> 5; }

struct Plan {
  int qual_bits = 8;
  std::uint8_t lut[16] = {};
  std::int64_t n_exc = 0;
  std::int64_t hap_bases = 0, read_bases = 0, hap_chunks = 0, read_chunks = 0;
  int max_hap_len = 0, max_read_len = 0;
  std::size_t off_hap_len = 0, off_read_len = 0, off_name_hash = 0, off_var = 0, off_hap_planes = 0, off_read_planes = 0,
              off_qual = 0, off_exc = 0, bytes = 0;
  int rc = LGR_OK;
};

// ---- 32 bytes → validity mask / bit planes ---------------------------------------------------
#if LGR_PACK_X86
__attribute__((target("avx2"))) inline std::uint32_t valid32_avx2(const std::uint8_t* s) {
  const __m256i x = _mm256_and_si256(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(s)), _mm256_set1_epi8((char)0xDF));
  const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(x, _mm256_set1_epi8('A')), _mm256_cmpeq_epi8(x, _mm256_set1_epi8('C'))),
                                     _mm256_or_si256(_mm256_cmpeq_epi8(x, _mm256_set1_epi8('G')), _mm256_cmpeq_epi8(x, _mm256_set1_epi8('T'))));
  return (std::uint32_t)_mm256_movemask_epi8(ok);
}
__attribute__((target("avx2"))) inline void planes32_avx2(const std::uint8_t* s, std::uint32_t* lo, std::uint32_t* hi) {
  const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s));
  const __m256i t = _mm256_xor_si256(x, _mm256_srli_epi16(x, 1));  // bit1 = c1^c2, bit2 = c2^c3
  *lo = (std::uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(t, 6));
  *hi = (std::uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(t, 5));
}
// quality bytes against a dictionary of at most 4 values: index planes + "all in the dictionary"
__attribute__((target("avx2"))) inline std::uint32_t qual32_avx2(const std::uint8_t* q, const std::uint8_t* lut, int n_lut, std::uint32_t* lo,
                                                                 std::uint32_t* hi) {
  const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(q));
  const __m256i e0 = _mm256_cmpeq_epi8(x, _mm256_set1_epi8((char)lut[0]));
  const __m256i e1 = n_lut > 1 ? _mm256_cmpeq_epi8(x, _mm256_set1_epi8((char)lut[1])) : _mm256_setzero_si256();
  const __m256i e2 = n_lut > 2 ? _mm256_cmpeq_epi8(x, _mm256_set1_epi8((char)lut[2])) : _mm256_setzero_si256();
  const __m256i e3 = n_lut > 3 ? _mm256_cmpeq_epi8(x, _mm256_set1_epi8((char)lut[3])) : _mm256_setzero_si256();
  *lo = (std::uint32_t)_mm256_movemask_epi8(_mm256_or_si256(e1, e3));
  *hi = (std::uint32_t)_mm256_movemask_epi8(_mm256_or_si256(e2, e3));
  return (std::uint32_t)_mm256_movemask_epi8(_mm256_or_si256(_mm256_or_si256(e0, e1), _mm256_or_si256(e2, e3)));
}
inline bool have_avx2() {
  static const bool v = __builtin_cpu_supports("avx2");
  return v;
}
#else
inline bool have_avx2() { return false; }
#endif

inline std::uint32_t valid32_scalar(const std::uint8_t* s, int n) {
  std::uint32_t m = 0;
  for (int i = 0; i < n; ++i) m |= (std::uint32_t)is_acgt(s[i]) << i;
  return m;
}
inline void planes32_scalar(const std::uint8_t* s, int n, std::uint32_t* lo, std::uint32_t* hi) {
  std::uint32_t l = 0, h = 0;
  for (int i = 0; i < n; ++i) {
    const std::uint32_t c = s[i], code = ((c >> 1) ^ (c >> 2)) & 3u;
    l |= (code & 1u) << i, h |= (code >> 1) << i;
  }
  *lo = l, *hi = h;
}

// count the bases of s[0..len) that need an exception entry
#if LGR_PACK_X86
// (whole-sequence routines carry the target attribute themselves, so that the 32-byte helpers inline into their
// loops: as separate calls per chunk the helpers were a third of the packing time)
__attribute__((target("avx2"))) inline std::int64_t count_exceptions_avx2(const std::uint8_t* s, int len) {
  std::int64_t n = 0;
  int i = 0;
  for (; i + 32 <= len; i += 32) n += 32 - __builtin_popcount(valid32_avx2(s + i));
  if (i < len) {
    // the partial last chunk: the LAST 32 bytes of the string, of which the top len - i bits are new
    // (no padded copy, and nothing is read beyond the caller's string); len >= 32 is the caller's check
    const std::uint32_t ok = valid32_avx2(s + len - 32) >> (32 - (len - i));
    n += (len - i) - __builtin_popcount(ok);
  }
  return n;
}
#endif
inline std::int64_t count_exceptions(const std::uint8_t* s, int len) {
#if LGR_PACK_X86
  if (len >= 32 && have_avx2()) return count_exceptions_avx2(s, len);
#endif
  std::int64_t n = 0;
  for (int i = 0; i < len; ++i) n += !is_acgt(s[i]);
  return n;
}

// bases of one sequence → planes (2 words per chunk) + exception entries (pos = base + index)
inline void emit_chunk(const std::uint8_t* s, int off, int n, std::uint32_t lo, std::uint32_t hi, std::uint32_t ok, std::uint32_t* planes,
                       int c, std::uint32_t pos_base, std::uint32_t* exc_pos, std::uint8_t* exc_code, std::int64_t* n_exc) {
  const std::uint32_t full = n == 32 ? 0xffffffffu : ((1u << n) - 1u);
  std::uint32_t bad = ~ok & full;
  lo &= ok, hi &= ok;  // exception positions carry 0 bits: the record is a pure function of the payload
  planes[2 * c] = lo, planes[2 * c + 1] = hi;
  while (bad) {
    const int b = __builtin_ctz(bad);
    bad &= bad - 1;
    exc_pos[*n_exc] = pos_base + (std::uint32_t)(off + b);
    exc_code[*n_exc] = code_of(s[off + b]);
    ++*n_exc;
  }
}
#if LGR_PACK_X86
__attribute__((target("avx2"))) inline void pack_bases_avx2(const std::uint8_t* s, int len, std::uint32_t* planes, std::uint32_t pos_base,
                                                            std::uint32_t* exc_pos, std::uint8_t* exc_code, std::int64_t* n_exc) {
  const int nc = chunks_of(len);
  for (int c = 0; c < nc; ++c) {
    const int off = c << 5, n = len - off < 32 ? len - off : 32;
    // the last, partial chunk is read as the LAST 32 bytes of the string and shifted down: never a read past
    // the caller's string, and no padded copy (len >= 32 is the caller's check)
    const std::uint8_t* src = n == 32 ? s + off : s + len - 32;
    std::uint32_t lo, hi;
    planes32_avx2(src, &lo, &hi);
    std::uint32_t ok = valid32_avx2(src);
    if (n < 32) lo >>= 32 - n, hi >>= 32 - n, ok >>= 32 - n;
    emit_chunk(s, off, n, lo, hi, ok, planes, c, pos_base, exc_pos, exc_code, n_exc);
  }
}
#endif
inline void pack_bases(const std::uint8_t* s, int len, std::uint32_t* planes, std::uint32_t pos_base, std::uint32_t* exc_pos,
                       std::uint8_t* exc_code, std::int64_t* n_exc) {
#if LGR_PACK_X86
  if (len >= 32 && have_avx2()) return pack_bases_avx2(s, len, planes, pos_base, exc_pos, exc_code, n_exc);
#endif
  const int nc = chunks_of(len);
  for (int c = 0; c < nc; ++c) {
    const int off = c << 5, n = len - off < 32 ? len - off : 32;
    std::uint32_t lo, hi;
    planes32_scalar(s + off, n, &lo, &hi);
    emit_chunk(s, off, n, lo, hi, valid32_scalar(s + off, n), planes, c, pos_base, exc_pos, exc_code, n_exc);
  }
}

// the dictionary of the group's quality values: at most 16 distinct → lut (ascending), else 8 bits raw
inline int build_qual_lut(const lgr_group_desc* g, std::uint8_t* lut) {
  bool seen[256] = {};
  for (int r = 0; r < g->n_reads; ++r) {
    const std::uint8_t* q = g->read_qual[r];
    const int n = g->read_len[r];
    for (int i = 0; i < n; ++i) seen[q[i]] = true;
  }
  int k = 0;
  for (int v = 0; v < 256; ++v)
    if (seen[v]) {
      if (k == 16) return 8;
      lut[k++] = (std::uint8_t)v;
    }
  for (int i = k; i < 16; ++i) lut[i] = k ? lut[k - 1] : 0;
  return k <= 4 ? 2 : 4;
}

// does every quality byte of the group lie in lut[0..n_lut)? (n_lut <= 4)
#if LGR_PACK_X86
__attribute__((target("avx2"))) inline bool quals_fit_avx2(const std::uint8_t* q, int n, const std::uint8_t* lut, int n_lut) {
  std::uint32_t lo, hi;
  int i = 0;
  for (; i + 32 <= n; i += 32)
    if (qual32_avx2(q + i, lut, n_lut, &lo, &hi) != 0xffffffffu) return false;
  // the last 32 bytes cover the partial chunk (bytes seen twice do not matter here); n >= 32 is the caller's check
  return i == n || qual32_avx2(q + n - 32, lut, n_lut, &lo, &hi) == 0xffffffffu;
}
// 2-bit quality planes of one read (n >= 32): two words per chunk, index 0 beyond the end
__attribute__((target("avx2"))) inline void pack_quals2_avx2(const std::uint8_t* q, int n, const std::uint8_t* lut, std::uint32_t* w) {
  const int nc = chunks_of(n);
  for (int c = 0; c < nc; ++c) {
    const int off = c << 5, m = n - off < 32 ? n - off : 32;
    (void)qual32_avx2(m == 32 ? q + off : q + n - 32, lut, 4, &w[2 * c], &w[2 * c + 1]);
    if (m < 32) w[2 * c] >>= 32 - m, w[2 * c + 1] >>= 32 - m;
  }
}
#endif
inline bool quals_fit(const lgr_group_desc* g, const std::uint8_t* lut, int n_lut) {
  for (int r = 0; r < g->n_reads; ++r) {
    const std::uint8_t* q = g->read_qual[r];
    const int n = g->read_len[r];
#if LGR_PACK_X86
    if (n >= 32 && have_avx2()) {
      if (!quals_fit_avx2(q, n, lut, n_lut)) return false;
      continue;
    }
#endif
    for (int i = 0; i < n; ++i) {
      bool ok = false;
      for (int k = 0; k < n_lut; ++k) ok |= q[i] == lut[k];
      if (!ok) return false;
    }
  }
  return true;
}

// sizing pass: exception count, quality dictionary, section offsets.  `hint` (may be NULL) is the
// dictionary of the caller's previous group: consecutive windows of one run share it, and checking
// it is one vector compare per 32 bytes while building one is a byte loop.
inline Plan plan_group(const lgr_group_desc* g, const Plan* hint = nullptr) {
  Plan p;
  if (!g || g->n_haps < 0 || g->n_reads < 0 || g->n_vars < 0 || (g->n_reads > 0 && g->n_haps == 0)) {
    p.rc = LGR_E_ARG;
    return p;
  }
  for (int h = 0; h < g->n_haps; ++h) {
    const int l = g->hap_len[h];
    if (l < 0) { p.rc = LGR_E_ARG; return p; }
    if (l > LGR_MAX_HAP_LEN) { p.rc = LGR_E_LIMIT; return p; }
    p.hap_bases += l, p.hap_chunks += chunks_of(l);
    if (l > p.max_hap_len) p.max_hap_len = l;
    p.n_exc += count_exceptions(g->hap_seq[h], l);
  }
  for (int r = 0; r < g->n_reads; ++r) {
    const int l = g->read_len[r];
    if (l < 0) { p.rc = LGR_E_ARG; return p; }
    if (l > LGR_MAX_READ_LEN) { p.rc = LGR_E_LIMIT; return p; }
    p.read_bases += l, p.read_chunks += chunks_of(l);
    if (l > p.max_read_len) p.max_read_len = l;
    p.n_exc += count_exceptions(g->read_seq[r], l);
  }
  if (p.hap_bases > INT32_MAX || p.read_bases > INT32_MAX) { p.rc = LGR_E_LIMIT; return p; }
  if (hint && hint->qual_bits == 2 && quals_fit(g, hint->lut, 4)) {
    p.qual_bits = 2;
    std::memcpy(p.lut, hint->lut, 16);
  } else {
    p.qual_bits = build_qual_lut(g, p.lut);
  }
  const std::size_t vh = (std::size_t)g->n_vars * (std::size_t)g->n_haps;
  std::size_t o = sizeof(lgr_group_rec_hdr);
  p.off_hap_len = o, o += align_up(sizeof(std::int32_t) * (std::size_t)g->n_haps, 8);
  p.off_read_len = o, o += align_up(sizeof(std::uint16_t) * (std::size_t)g->n_reads, 8);
  p.off_name_hash = o, o += align_up(sizeof(std::uint32_t) * (std::size_t)g->n_reads, 8);
  p.off_var = o, o += align_up(9 * vh, 8);  // i32 start[vh], i32 len[vh], i8 allele[vh]
  p.off_hap_planes = o, o += 8 * (std::size_t)p.hap_chunks;
  p.off_read_planes = o, o += 8 * (std::size_t)p.read_chunks;
  p.off_qual = o;
  o += p.qual_bits == 8 ? align_up((std::size_t)p.read_bases, 8) : align_up(4 * (std::size_t)p.qual_bits * (std::size_t)p.read_chunks, 8);
  p.off_exc = o, o += align_up(5 * (std::size_t)p.n_exc, 8);  // u32 pos[n], u8 code[n]
  p.bytes = align_up(o, 16);
  if (p.bytes > 0xffffffffull) p.rc = LGR_E_LIMIT;
  return p;
}

// pack the payload into dst (p.bytes bytes, 8-byte aligned); fills *dir except rec_off
inline int pack_group(const lgr_group_desc* g, const Plan& p, void* dst, lgr_group_dir* dir) {
  if (p.rc != LGR_OK) return p.rc;
  std::uint8_t* base = static_cast<std::uint8_t*>(dst);
  lgr_group_rec_hdr hdr;
  std::memset(&hdr, 0, sizeof(hdr));
  hdr.magic = LGR_PACK_MAGIC, hdr.qual_bits = (std::uint32_t)p.qual_bits, hdr.n_exc = (std::uint32_t)p.n_exc;
  hdr.rec_bytes = (std::uint32_t)p.bytes;
  std::memcpy(hdr.qual_lut, p.lut, 16);
  hdr.off_hap_len = (std::uint32_t)p.off_hap_len, hdr.off_read_len = (std::uint32_t)p.off_read_len;
  hdr.off_name_hash = (std::uint32_t)p.off_name_hash, hdr.off_var = (std::uint32_t)p.off_var;
  hdr.off_hap_planes = (std::uint32_t)p.off_hap_planes, hdr.off_read_planes = (std::uint32_t)p.off_read_planes;
  hdr.off_qual = (std::uint32_t)p.off_qual, hdr.off_exc = (std::uint32_t)p.off_exc;
  std::memcpy(base, &hdr, sizeof(hdr));
  std::int32_t* hap_len = reinterpret_cast<std::int32_t*>(base + p.off_hap_len);
  std::uint16_t* read_len = reinterpret_cast<std::uint16_t*>(base + p.off_read_len);
  std::uint32_t* name_hash = reinterpret_cast<std::uint32_t*>(base + p.off_name_hash);
  const std::size_t vh = (std::size_t)g->n_vars * (std::size_t)g->n_haps;
  std::int32_t* var_start = reinterpret_cast<std::int32_t*>(base + p.off_var);
  std::int32_t* var_len = var_start + vh;
  std::int8_t* var_allele = reinterpret_cast<std::int8_t*>(var_len + vh);
  std::uint32_t* hap_planes = reinterpret_cast<std::uint32_t*>(base + p.off_hap_planes);
  std::uint32_t* read_planes = reinterpret_cast<std::uint32_t*>(base + p.off_read_planes);
  std::uint32_t* exc_pos = reinterpret_cast<std::uint32_t*>(base + p.off_exc);
  std::uint8_t* exc_code = reinterpret_cast<std::uint8_t*>(exc_pos + p.n_exc);
  if (vh) {
    std::memcpy(var_start, g->var_start, sizeof(std::int32_t) * vh);
    std::memcpy(var_len, g->var_len, sizeof(std::int32_t) * vh);
    std::memcpy(var_allele, g->var_allele, vh);
  }
  std::int64_t n_exc = 0;
  std::uint32_t pos = 0;
  std::int64_t chunk = 0;
  for (int h = 0; h < g->n_haps; ++h) {
    const int l = g->hap_len[h];
    hap_len[h] = l;
    pack_bases(g->hap_seq[h], l, hap_planes + 2 * chunk, pos, exc_pos, exc_code, &n_exc);
    pos += (std::uint32_t)l, chunk += chunks_of(l);
  }
  pos = 0, chunk = 0;
  std::uint8_t* qraw = base + p.off_qual;
  std::uint32_t* qplanes = reinterpret_cast<std::uint32_t*>(base + p.off_qual);
  std::uint8_t inv[256];
  if (p.qual_bits == 4) {
    std::memset(inv, 0, sizeof(inv));
    for (int k = 15; k >= 0; --k) inv[p.lut[k]] = (std::uint8_t)k;
  }
  for (int r = 0; r < g->n_reads; ++r) {
    const int l = g->read_len[r];
    read_len[r] = (std::uint16_t)l;
    name_hash[r] = g->read_name_hash[r];
    pack_bases(g->read_seq[r], l, read_planes + 2 * chunk, pos | 0x80000000u, exc_pos, exc_code, &n_exc);
    const std::uint8_t* q = g->read_qual[r];
    if (p.qual_bits == 8) {
      if (l) std::memcpy(qraw + pos, q, (std::size_t)l);
    } else {
      const int nc = chunks_of(l);
#if LGR_PACK_X86
      if (p.qual_bits == 2 && l >= 32 && have_avx2()) {
        pack_quals2_avx2(q, l, p.lut, qplanes + (std::size_t)chunk * 2);
        pos += (std::uint32_t)l, chunk += nc;
        continue;
      }
#endif
      for (int c = 0; c < nc; ++c) {
        const int off = c << 5, n = l - off < 32 ? l - off : 32;
        std::uint32_t* w = qplanes + (std::size_t)(chunk + c) * (std::size_t)p.qual_bits;
        if (p.qual_bits == 2) {
          std::uint32_t lo = 0, hi = 0;
          for (int i = 0; i < n; ++i) {
            const std::uint8_t v = q[off + i];
            // same planes as the vector form: entries a short dictionary repeats OR together
            lo |= (std::uint32_t)(v == p.lut[1] || v == p.lut[3]) << i, hi |= (std::uint32_t)(v == p.lut[2] || v == p.lut[3]) << i;
          }
          w[0] = lo, w[1] = hi;
        } else {
          std::uint32_t b0 = 0, b1 = 0, b2 = 0, b3 = 0;
          for (int i = 0; i < n; ++i) {
            const std::uint32_t k = inv[q[off + i]];
            b0 |= (k & 1u) << i, b1 |= (k >> 1 & 1u) << i, b2 |= (k >> 2 & 1u) << i, b3 |= (k >> 3) << i;
          }
          w[0] = b0, w[1] = b1, w[2] = b2, w[3] = b3;
        }
      }
    }
    pos += (std::uint32_t)l, chunk += chunks_of(l);
  }
  if (n_exc != p.n_exc) return LGR_E_ARG;  // the payload changed between plan and pack
  if (dir) {
    dir->n_haps = g->n_haps, dir->n_reads = g->n_reads, dir->n_vars = g->n_vars;
    dir->hap_bases = (std::int32_t)p.hap_bases, dir->read_bases = (std::int32_t)p.read_bases;
    dir->mid_occ = g->mid_occ;
    dir->max_hap_len = p.max_hap_len, dir->max_read_len = p.max_read_len;
  }
  return LGR_OK;
}

}  // namespace lgr_pack

#endif  // LANCET2_B200_LGR_PACK_H_
