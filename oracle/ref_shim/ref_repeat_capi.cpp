// TEST INFRASTRUCTURE: C wrappers around the reference's OWN repeat detection,
// src/lancet/base/repeat.cpp compiled unmodified (see ../Makefile target `ref`).  Used only by tests/
// and tools/bench_repeat.py's CPU baseline to pin oracle/repeat_oracle.cpp and the CUDA kernel.
#include <cstdint>
#include <string_view>
#include <vector>

#include "absl/types/span.h"
#include "lancet/base/repeat.h"

extern "C" {

// lancet::base::HammingDist (repeat.cpp:219)
uint64_t ref_hamming_dist(const char* a, const char* b, int64_t n) {
  return lancet::base::HammingDist(std::string_view(a, (size_t)n), std::string_view(b, (size_t)n));
}

// lancet::base::HasRepeat (repeat.cpp:348) over the sliding k-mers of seq — what
// cbdg::Graph::HasExactOrApproxRepeat (graph.h:127-131, max_mismatches = 2) and
// VariantBuilder::ShouldSkipWindow (variant_builder.cpp:116-117, max_mismatches = 0) compute.
// The k-mer views are built as base::SlidingView builds them (sliding.h:17-34: none when the
// sequence is shorter than k, else every offset 0 .. len - k).
int ref_has_repeat(const char* seq, int64_t len, int64_t k, int64_t max_mismatches) {
  std::vector<std::string_view> kmers;
  if (len >= k && k > 0) {
    kmers.reserve((size_t)(len - k + 1));
    for (int64_t i = 0; i + k <= len; ++i) kmers.emplace_back(seq + i, (size_t)k);
  }
  return lancet::base::HasRepeat(absl::MakeConstSpan(kmers), (size_t)max_mismatches) ? 1 : 0;
}

// the same over an explicit list of n_kmers k-mers stored back to back (the reference's KAT form)
int ref_has_repeat_kmers(const char* kmers, int64_t n_kmers, int64_t k, int64_t max_mismatches) {
  std::vector<std::string_view> views;
  for (int64_t i = 0; i < n_kmers; ++i) views.emplace_back(kmers + i * k, (size_t)k);
  return lancet::base::HasRepeat(absl::MakeConstSpan(views), (size_t)max_mismatches) ? 1 : 0;
}

}  // extern "C"
