// TEST INFRASTRUCTURE shim: absl::flat_hash_map → std::unordered_map
#ifndef SHIM_ABSL_FLAT_HASH_MAP_H_
#define SHIM_ABSL_FLAT_HASH_MAP_H_
#include <unordered_map>
namespace absl {
template <typename K, typename V, typename H = std::hash<K>, typename E = std::equal_to<K>>
using flat_hash_map = std::unordered_map<K, V, H, E>;
}
#endif
