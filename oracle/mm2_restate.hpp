// TEST INFRASTRUCTURE — CPU oracle.  Not part of the product path.
//
// mm2_restate: a from-scratch, single-threaded C++ restatement of the stages of
// lh3/minimap2 **v2.30** that `mm_map()` executes under the option set Lancet2's
// Genotyper installs (reference: src/lancet/caller/genotyper.cpp:89-191 for the
// options, :243-267 for mm_idx_str/mm_mapopt_update, :376-411 for mm_map).
//
// PARITY UNPINNED: minimap2 is a build-time download of the reference
// (cmake/dependencies.cmake:160-171) and is not present in /root/reference nor
// anywhere in the build container; the reference has no test that calls mm_map
// (SURVEY.md §8c).  Everything here restates the published minimap2 algorithm
// (files named per function below) from knowledge of the upstream source, and is
// anchored on the reference's own call sites.  It has not been diffed against a
// real minimap2 2.30 binary.
#ifndef ORACLE_MM2_RESTATE_HPP_
#define ORACLE_MM2_RESTATE_HPP_

#include <cstdint>
#include <string>
#include <vector>

#include "../include/lancet_gpu_realign.h"

namespace mm2r {

struct U128 {
  uint64_t x, y;
};

// minimap2 sketch.c: seq_nt4_table (A/a 0, C/c 1, G/g 2, T/t/U/u 3, else 4)
uint8_t Nt4(uint8_t c);

// minimap2 sketch.c: mm_sketch(km, str, len, w, k, rid, is_hpc=0, p)
void Sketch(const uint8_t* str, int len, int w, int k, uint32_t rid, std::vector<U128>& out);

// ksort.h: radix_sort_128x / radix_sort_64 (in-place MSD radix, insertion sort <= 64)
void RadixSort128x(U128* beg, U128* end);
void RadixSort64(uint64_t* beg, uint64_t* end);

// khash.h hashes used by mm_map_frag
uint32_t X31HashString(const char* s);
uint32_t WangHash(uint32_t key);

// minimap2 index.c: mm_idx_str for ONE sequence (the reference builds one index
// per haplotype, genotyper.cpp:248-252).
struct HapIndex {
  int k = 0, w = 0;
  std::vector<uint8_t> seq4;   // mm_idx_t::S content: nt4 codes 0..4 (4 bits/base upstream)
  std::vector<uint64_t> keys;  // sorted minimizer hashes (x>>8), one per occurrence
  std::vector<uint64_t> vals;  // y = rid<<32 | pos<<1 | strand, sorted within a key
  // mm_idx_get: occurrences of `minier`; returns pointer into vals and count
  const uint64_t* Get(uint64_t minier, int* n) const;
  // mm_idx_cal_max_occ(mi, f)
  int32_t CalMaxOcc(float f) const;
};
void BuildHapIndex(const uint8_t* hap, int len, int w, int k, HapIndex& idx);

// mm_mapopt_update's mid_occ rule (options.c)
int32_t MidOccFromIndex(const HapIndex& idx, const lgr_params& p);

struct Reg {  // the mm_reg1_t / mm_extra_t fields this path touches
  int32_t id = 0, cnt = 0, score = 0, score0 = 0, qs = 0, qe = 0, rs = 0, re = 0;
  int32_t parent = 0, subsc = 0, as = 0, mlen = 0, blen = 0, n_sub = 0;
  uint32_t hash = 0;
  bool rev = false, strand_retained = false, has_p = false;
  int32_t dp_score = 0, dp_max = 0, dp_max2 = 0, n_ambi = 0;
  std::vector<uint32_t> cigar;
};

struct MapDebug {  // every intermediate, for GPU-vs-oracle stage tests
  std::vector<U128> mv;       // query minimizers after mm_seed_mz_flt
  std::vector<U128> anchors;  // sorted seed hits fed to mg_lchain_dp
  std::vector<int32_t> f;     // chain score per anchor
  std::vector<int64_t> p;     // chain predecessor per anchor
  std::vector<uint64_t> u;    // chains: score<<32 | n
  std::vector<U128> chained;  // compact_a output
  int64_t chain_evals = 0;
  int64_t dp_cells_full = 0;  // ksw2 rectangle cells (tlen*qlen per extension)
  int32_t n_regs_chain = 0;   // regs after chain_post
  int32_t rep_len = 0;
};

// mm_map(): returns the final reg list (regs[0] is what the reference consumes).
// qname_hash = __ac_X31_hash_string(qname).  mid_occ must be > 0.
std::vector<Reg> Map(const HapIndex& idx, const uint8_t* read, int qlen, uint32_t qname_hash,
                     const lgr_params& p, int32_t mid_occ, MapDebug* dbg = nullptr);

// ksw2_extz2_sse.c + ksw2.h:ksw_backtrack in absolute-score form.
struct ExtzResult {
  int32_t max = 0, max_q = -1, max_t = -1, mqe = -0x40000000, mqe_t = -1;
  int reach_end = 0, zdropped = 0;
  std::vector<uint32_t> cigar;
};
// flags
enum { kEzRight = 1, kEzRevCigar = 2 };
void ExtzOnly(int qlen, const uint8_t* q, int tlen, const uint8_t* t, const int8_t* mat, int gapo,
              int gape, int end_bonus, int flag, ExtzResult& ez);

}  // namespace mm2r

#endif  // ORACLE_MM2_RESTATE_HPP_
