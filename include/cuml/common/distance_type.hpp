// ML::distance::DistanceType values the k-means path accepts
// (subset of reference cpp/include/cuml/common/distance_type.hpp:13-36; same numeric values).
#pragma once
namespace ML {
namespace distance {
enum class DistanceType : int {
  L2Expanded     = 0,  // squared Euclidean, expanded form
  L2SqrtExpanded = 1   // Euclidean, expanded form
};
}  // namespace distance
}  // namespace ML
