/*
 * TEST INFRASTRUCTURE ONLY (oracle/): stand-in for robin-hood-hashing, which the
 * reference Makefile clones unpinned at install time (reference src/Makefile:16,52-53)
 * and which is absent here (no network).  The reference uses it purely as an
 * associative container (hashtrie.hpp:49-50), so aliasing std::unordered_map cannot
 * change any query result -- only the builder's bucket order and the CPU baseline's
 * speed (std::unordered_map is slower than robin_hood's flat map; stated wherever the
 * CPU number is reported).
 */
#ifndef ORACLE_SHIM_ROBIN_HOOD_H
#define ORACLE_SHIM_ROBIN_HOOD_H
#include <unordered_map>
namespace robin_hood {
template <class K, class V>
using unordered_map = std::unordered_map<K, V>;
}
#endif
