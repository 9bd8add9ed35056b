// TEST INFRASTRUCTURE — shim so that the reference's base/repeat.cpp compiles unmodified without
// abseil: HasRepeat only needs reserve / insert returning {iterator, inserted} on string_view keys.
#ifndef REF_SHIM_ABSL_CONTAINER_FLAT_HASH_SET_H_
#define REF_SHIM_ABSL_CONTAINER_FLAT_HASH_SET_H_
#include <unordered_set>
namespace absl {
template <class T, class H = std::hash<T>, class E = std::equal_to<T>>
using flat_hash_set = std::unordered_set<T, H, E>;
}  // namespace absl
#endif  // REF_SHIM_ABSL_CONTAINER_FLAT_HASH_SET_H_
