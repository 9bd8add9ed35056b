// TEST INFRASTRUCTURE shim: absl::InlinedVector → std::vector
#ifndef SHIM_ABSL_INLINED_VECTOR_H_
#define SHIM_ABSL_INLINED_VECTOR_H_
#include <cstddef>
#include <vector>
namespace absl {
template <typename T, std::size_t N>
using InlinedVector = std::vector<T>;
}
#endif
