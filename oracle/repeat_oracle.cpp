// TEST INFRASTRUCTURE — CPU restatement of the reference's repeat detection,
// pinned against the reference's own src/lancet/base/repeat.cpp compiled into oracle/_ref
// (tests/test_oracle_repeat.py) and against its known-answer tests (tests/base/repeat_test.cpp).
// Only tests/, __graft_entry__.smoke() and tools/bench_repeat.py's CPU baseline may call this.
#include <algorithm>
#include <cstdint>
#include <cstring>

extern "C" {

// lancet::base::HammingDist (reference src/lancet/base/repeat.cpp:219-330): byte-level count of
// positions where two equal-length strings differ.
uint64_t orc_hamming_dist(const char* a, const char* b, int64_t n) {
  uint64_t d = 0;
  for (int64_t i = 0; i < n; ++i) d += a[i] != b[i];
  return d;
}

// lancet::base::HasRepeat (repeat.cpp:348-371) over base::SlidingView(seq, k) (sliding.h:17-34):
// true iff two k-mers at different offsets differ in at most max_mismatches byte positions.
// max_mismatches == 0 is the reference's hash-set duplicate check (same predicate).  Fewer than two
// k-mers (len < k + 1) -> false.  Bytes are compared raw: case and IUPAC letters count as written.
// Also reports the smallest Hamming distance seen before the early exit would have fired (for tests).
int orc_has_repeat(const char* seq, int64_t len, int64_t k, int64_t max_mismatches) {
  if (k <= 0 || len < k + 1) return 0;
  const int64_t n = len - k + 1;
  for (int64_t i = 0; i < n; ++i)
    for (int64_t j = i + 1; j < n; ++j) {
      int64_t d = 0;
      for (int64_t p = 0; p < k && d <= max_mismatches; ++p) d += seq[i + p] != seq[j + p];
      if (d <= max_mismatches) return 1;
    }
  return 0;
}

// HasRepeat over an explicit k-mer list (n_kmers strings of k bytes, concatenated) — the form the
// reference's own known-answer tests use (tests/base/repeat_test.cpp).
int orc_has_repeat_kmers(const char* kmers, int64_t n_kmers, int64_t k, int64_t max_mismatches) {
  for (int64_t i = 0; i < n_kmers; ++i)
    for (int64_t j = i + 1; j < n_kmers; ++j) {
      int64_t d = 0;
      for (int64_t p = 0; p < k && d <= max_mismatches; ++p) d += kmers[i * k + p] != kmers[j * k + p];
      if (d <= max_mismatches) return 1;
    }
  return 0;
}

// exhaustive companion for the tests: the minimum Hamming distance over all k-mer pairs (-1 if < 2 k-mers)
int64_t orc_min_kmer_distance(const char* seq, int64_t len, int64_t k) {
  if (k <= 0 || len < k + 1) return -1;
  const int64_t n = len - k + 1;
  int64_t best = k + 1;
  for (int64_t i = 0; i < n; ++i)
    for (int64_t j = i + 1; j < n; ++j) {
      int64_t d = 0;
      for (int64_t p = 0; p < k && d < best; ++p) d += seq[i + p] != seq[j + p];
      best = std::min(best, d);
    }
  return best;
}

}  // extern "C"
