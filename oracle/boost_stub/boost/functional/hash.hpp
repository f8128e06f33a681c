// Minimal stand-in for <boost/functional/hash.hpp> so the reference's cpp/parallel_weighted_astar.cpp
// (line 30, 107) compiles without Boost installed.  Only boost::hash_range is used there, as the bucket
// hash of a std::unordered_set; its value is unobservable outside the container.
#pragma once
#include <cstddef>
namespace boost {
template <class It> inline std::size_t hash_range(It first, It last) {
  std::size_t seed = 0;
  for (; first != last; ++first) seed ^= static_cast<std::size_t>(*first) + 0x9e3779b97f4a7c15ULL + (seed << 6) + (seed >> 2);
  return seed;
}
}  // namespace boost
