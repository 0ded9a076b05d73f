// Compile-only check of the drop-in boundary: every one of the 14 ML::kmeans free functions the reference declares
// (cpp/include/cuml/cluster/kmeans.hpp:41-79 fit, :110-130 partition-list fit, :154-195 predict, :213-242 transform)
// exists in include/cuml/cluster/kmeans.hpp with exactly the reference's parameter types, and KMeansParams carries
// the reference's fields and defaults (cpp/include/cuml/cluster/kmeans_params.hpp:17-32).  A wrong parameter type
// makes the overload selection below ill-formed, so the translation unit does not compile.
#include <cuml/cluster/kmeans.hpp>

#include <cstdint>
#include <cstdio>
#include <type_traits>

using H = const raft::handle_t&;
using P = const ML::kmeans::KMeansParams&;

template <typename T, typename I>
using fit_t = void (*)(H, P, const T*, I, I, const T*, T*, T&, I&);
template <typename T>
using fit_parts_t = void (*)(H, P, const T* const*, const int64_t*, int64_t, int64_t, const T* const*, T*, T&, int64_t&);
template <typename T, typename I>
using predict_t = void (*)(H, P, const T*, const T*, I, I, const T*, bool, I*, T&);
template <typename T, typename I>
using transform_t = void (*)(H, P, const T*, const T*, I, I, T*);

int main()
{
  const void* fns[] = {
    reinterpret_cast<const void*>(static_cast<fit_t<float, int>>(&ML::kmeans::fit)),
    reinterpret_cast<const void*>(static_cast<fit_t<double, int>>(&ML::kmeans::fit)),
    reinterpret_cast<const void*>(static_cast<fit_t<float, int64_t>>(&ML::kmeans::fit)),
    reinterpret_cast<const void*>(static_cast<fit_t<double, int64_t>>(&ML::kmeans::fit)),
    reinterpret_cast<const void*>(static_cast<fit_parts_t<float>>(&ML::kmeans::fit)),
    reinterpret_cast<const void*>(static_cast<fit_parts_t<double>>(&ML::kmeans::fit)),
    reinterpret_cast<const void*>(static_cast<predict_t<float, int>>(&ML::kmeans::predict)),
    reinterpret_cast<const void*>(static_cast<predict_t<double, int>>(&ML::kmeans::predict)),
    reinterpret_cast<const void*>(static_cast<predict_t<float, int64_t>>(&ML::kmeans::predict)),
    reinterpret_cast<const void*>(static_cast<predict_t<double, int64_t>>(&ML::kmeans::predict)),
    reinterpret_cast<const void*>(static_cast<transform_t<float, int>>(&ML::kmeans::transform)),
    reinterpret_cast<const void*>(static_cast<transform_t<double, int>>(&ML::kmeans::transform)),
    reinterpret_cast<const void*>(static_cast<transform_t<float, int64_t>>(&ML::kmeans::transform)),
    reinterpret_cast<const void*>(static_cast<transform_t<double, int64_t>>(&ML::kmeans::transform)),
  };
  static_assert(sizeof(fns) / sizeof(fns[0]) == 14, "the reference declares 14 entry points");

  // KMeansParams: field types and defaults of the reference
  ML::kmeans::KMeansParams p;
  static_assert(std::is_same<decltype(p.n_clusters), int>::value, "n_clusters");
  static_assert(std::is_same<decltype(p.max_iter), int>::value, "max_iter");
  static_assert(std::is_same<decltype(p.tol), double>::value, "tol");
  static_assert(std::is_same<decltype(p.n_init), int>::value, "n_init");
  static_assert(std::is_same<decltype(p.oversampling_factor), double>::value, "oversampling_factor");
  static_assert(std::is_same<decltype(p.batch_samples), int>::value, "batch_samples");
  static_assert(std::is_same<decltype(p.batch_centroids), int>::value, "batch_centroids");
  static_assert(std::is_same<decltype(p.init_size), int64_t>::value, "init_size");
  static_assert(std::is_same<decltype(p.device_buffer_samples), int64_t>::value, "device_buffer_samples");
  static_assert(std::is_same<decltype(p.rng_state.seed), uint64_t>::value, "rng_state.seed");
  static_assert(static_cast<int>(ML::kmeans::KMeansParams::InitMethod::KMeansPlusPlus) == 0 &&
                  static_cast<int>(ML::kmeans::KMeansParams::InitMethod::Random) == 1 &&
                  static_cast<int>(ML::kmeans::KMeansParams::InitMethod::Array) == 2,
                "InitMethod values");
  static_assert(static_cast<int>(ML::distance::DistanceType::L2Expanded) == 0 &&
                  static_cast<int>(ML::distance::DistanceType::L2SqrtExpanded) == 1,
                "DistanceType values");
  int bad = 0;
  bad += p.metric != ML::distance::DistanceType::L2Expanded;
  bad += p.n_clusters != 8;
  bad += p.init != ML::kmeans::KMeansParams::InitMethod::KMeansPlusPlus;
  bad += p.max_iter != 300;
  bad += p.tol != 1e-4;
  bad += p.rng_state.seed != 0;
  bad += p.n_init != 1;
  bad += p.oversampling_factor != 2.0;
  bad += p.batch_samples != (1 << 15);
  bad += p.batch_centroids != 0;
  bad += p.init_size != 0;
  bad += p.device_buffer_samples != 0;
  for (const void* f : fns) bad += (f == nullptr);
  std::printf("%s\n", bad ? "MISMATCH" : "surface ok");
  return bad;
}
