// ML::kmeans::{fit, predict, transform} -- the C++ operator surface of the reference
// (cpp/include/cuml/cluster/kmeans.hpp:41-79 fit, :110-130 partition-list fit, :154-195 predict,
// :213-242 transform), same signatures and argument meaning, implemented as thin inline
// forwarders to the C-ABI (include/cuml_b200/kmeans_c.h).  Where the reference's shims
// (cpp/src/kmeans/*.cu) forward to cuvs::cluster::kmeans, these forward to libcuml_b200.so.
// Errors surface as C++ exceptions like the reference's RAFT_EXPECTS / RAFT_CUDA_TRY:
// std::invalid_argument for bad arguments, std::runtime_error otherwise.
#pragma once
#include <cuml/cluster/kmeans_params.hpp>
#include <cuml/common/export.hpp>
#include <cuml_b200/kmeans_c.h>
#include <raft/core/handle.hpp>

#include <cstdint>
#include <stdexcept>
#include <string>

namespace CUML_EXPORT ML {
namespace kmeans {

namespace detail {
// field-by-field copy, the to_cuvs() role (reference cpp/src/kmeans/kmeans_params.hpp:15-34)
inline cuml_b200_kmeans_params_t to_c(const KMeansParams& p)
{
  cuml_b200_kmeans_params_t c;
  c.metric                = static_cast<int32_t>(p.metric);
  c.n_clusters            = p.n_clusters;
  c.init                  = static_cast<int32_t>(p.init);
  c.max_iter              = p.max_iter;
  c.tol                   = p.tol;
  c.verbosity             = static_cast<int32_t>(p.verbosity);
  c.rng_seed              = p.rng_state.seed;
  c.rng_base_subsequence  = p.rng_state.base_subsequence;
  c.rng_type              = static_cast<int32_t>(p.rng_state.type);
  c.n_init                = p.n_init;
  c.oversampling_factor   = p.oversampling_factor;
  c.batch_samples         = p.batch_samples;
  c.batch_centroids       = p.batch_centroids;
  c.init_size             = p.init_size;
  c.device_buffer_samples = p.device_buffer_samples;
  return c;
}
inline void raise(int status)
{
  if (status == CUML_B200_SUCCESS) return;
  const std::string msg = cuml_b200_last_error();
  if (status == CUML_B200_INVALID_ARGUMENT) throw std::invalid_argument(msg);
  throw std::runtime_error(msg);
}
}  // namespace detail

#define CUML_B200_FIT(T, IDX, SUFFIX)                                                                    \
  inline void fit(const raft::handle_t& handle, const KMeansParams& params, const T* X, IDX n_samples,    \
                  IDX n_features, const T* sample_weight, T* centroids, T& inertia, IDX& n_iter)          \
  {                                                                                                       \
    const cuml_b200_kmeans_params_t c = detail::to_c(params);                                             \
    detail::raise(cuml_b200_kmeans_fit_##SUFFIX(handle.c_handle(), &c, X, n_samples, n_features,          \
                                                sample_weight, centroids, &inertia, &n_iter));            \
  }
CUML_B200_FIT(float, int, f32_i32)
CUML_B200_FIT(double, int, f64_i32)
CUML_B200_FIT(float, int64_t, f32_i64)
CUML_B200_FIT(double, int64_t, f64_i64)
#undef CUML_B200_FIT

#define CUML_B200_FIT_PARTS(T, SUFFIX)                                                                    \
  inline void fit(const raft::handle_t& handle, const KMeansParams& params, const T* const* X_parts,       \
                  const int64_t* n_samples_parts, int64_t n_parts, int64_t n_features,                     \
                  const T* const* sample_weight_parts, T* centroids, T& inertia, int64_t& n_iter)          \
  {                                                                                                        \
    const cuml_b200_kmeans_params_t c = detail::to_c(params);                                              \
    detail::raise(cuml_b200_kmeans_fit_parts_##SUFFIX(handle.c_handle(), &c, X_parts, n_samples_parts,     \
                                                      n_parts, n_features, sample_weight_parts, centroids, \
                                                      &inertia, &n_iter));                                 \
  }
CUML_B200_FIT_PARTS(float, f32)
CUML_B200_FIT_PARTS(double, f64)
#undef CUML_B200_FIT_PARTS

#define CUML_B200_PREDICT(T, IDX, SUFFIX)                                                                  \
  inline void predict(const raft::handle_t& handle, const KMeansParams& params, const T* centroids,         \
                      const T* X, IDX n_samples, IDX n_features, const T* sample_weight,                    \
                      bool normalize_weights, IDX* labels, T& inertia)                                      \
  {                                                                                                         \
    const cuml_b200_kmeans_params_t c = detail::to_c(params);                                               \
    detail::raise(cuml_b200_kmeans_predict_##SUFFIX(handle.c_handle(), &c, centroids, X, n_samples,         \
                                                    n_features, sample_weight, normalize_weights ? 1 : 0,   \
                                                    labels, &inertia));                                     \
  }
CUML_B200_PREDICT(float, int, f32_i32)
CUML_B200_PREDICT(double, int, f64_i32)
CUML_B200_PREDICT(float, int64_t, f32_i64)
CUML_B200_PREDICT(double, int64_t, f64_i64)
#undef CUML_B200_PREDICT

#define CUML_B200_TRANSFORM(T, IDX, SUFFIX)                                                              \
  inline void transform(const raft::handle_t& handle, const KMeansParams& params, const T* centroids,     \
                        const T* X, IDX n_samples, IDX n_features, T* X_new)                              \
  {                                                                                                       \
    const cuml_b200_kmeans_params_t c = detail::to_c(params);                                             \
    detail::raise(cuml_b200_kmeans_transform_##SUFFIX(handle.c_handle(), &c, centroids, X, n_samples,     \
                                                      n_features, X_new));                                \
  }
CUML_B200_TRANSFORM(float, int, f32_i32)
CUML_B200_TRANSFORM(double, int, f64_i32)
CUML_B200_TRANSFORM(float, int64_t, f32_i64)
CUML_B200_TRANSFORM(double, int64_t, f64_i64)
#undef CUML_B200_TRANSFORM

}  // namespace kmeans
}  // namespace CUML_EXPORT ML
