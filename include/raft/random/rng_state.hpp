// Minimal stand-in for raft::random::RngState (seed + base_subsequence + generator type), the only
// raft::random type ML::kmeans::KMeansParams embeds (reference kmeans_params.hpp:25).
#pragma once
#include <cstdint>
namespace raft {
namespace random {
enum GeneratorType { GenPhilox = 0, GenPC };
struct RngState {
  explicit RngState(uint64_t _seed) : seed(_seed) {}
  RngState(uint64_t _seed, GeneratorType _type) : seed(_seed), type(_type) {}
  RngState(uint64_t _seed, uint64_t _base_subsequence, GeneratorType _type)
    : seed(_seed), base_subsequence(_base_subsequence), type(_type) {}
  uint64_t seed{0};
  uint64_t base_subsequence{0};
  GeneratorType type{GenPhilox};
};
}  // namespace random
}  // namespace raft
