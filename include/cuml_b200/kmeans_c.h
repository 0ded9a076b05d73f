/*
 * cuml_b200 -- C-ABI of the B200-native k-means engine.
 *
 * This is the drop-in boundary for the k-means hot path of rapidsai/cuml: every entry point
 * below replaces one `ML::kmeans::*` overload of the reference (all of which forward to
 * cuvs::cluster::kmeans, an un-vendored dependency).  Plain pointers and sizes only; no C++,
 * torch or raft types cross this line.  The C++ surface the reference exports
 * (namespace ML::kmeans, include/cuml/cluster/kmeans.hpp in this repo) and the Python
 * estimator (cuml_b200.cluster.KMeans) are thin wrappers over these symbols.
 *
 * Conventions (mirroring the reference, cpp/include/cuml/cluster/kmeans.hpp and
 * wiki/cpp/DEVELOPER_GUIDE.md:11-20,56-72,324-362):
 *   - all work is ordered on the handle's CUDA stream; calls are result-synchronous
 *     (inertia / n_iter are host outputs valid on return);
 *   - the caller owns every buffer (centroids [k,d], labels [n], X_new [n,k]);
 *   - one handle per calling thread; different handles may be used concurrently;
 *   - errors: non-zero status + thread-local message (cuml_b200_last_error); the C++
 *     wrappers re-throw them as exceptions like the reference's RAFT_EXPECTS.
 *   - There is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef CUML_B200_KMEANS_C_H
#define CUML_B200_KMEANS_C_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CUML_B200_API __attribute__((visibility("default")))
#else
#define CUML_B200_API
#endif

/* status codes */
enum {
  CUML_B200_SUCCESS          = 0,
  CUML_B200_INVALID_ARGUMENT = 1, /* reference: RAFT_EXPECTS / logic_error -> ValueError   */
  CUML_B200_CUDA_ERROR       = 2, /* reference: RAFT_CUDA_TRY -> raft::cuda_error          */
  CUML_B200_NCCL_ERROR       = 3, /* reference: comm.sync_stream != SUCCESS                */
  CUML_B200_INTERNAL_ERROR   = 4
};

/* ML::distance::DistanceType values used by the path
 * (reference cpp/include/cuml/common/distance_type.hpp:13-36). */
enum { CUML_B200_L2Expanded = 0, CUML_B200_L2SqrtExpanded = 1 };

/* ML::kmeans::KMeansParams::InitMethod (reference cpp/include/cuml/cluster/kmeans_params.hpp:18) */
enum { CUML_B200_INIT_KMeansPlusPlus = 0, CUML_B200_INIT_Random = 1, CUML_B200_INIT_Array = 2 };

/* C mirror of ML::kmeans::KMeansParams, same field order and defaults
 * (reference cpp/include/cuml/cluster/kmeans_params.hpp:17-32; rng_state flattened). */
typedef struct cuml_b200_kmeans_params {
  int32_t  metric;                /* = CUML_B200_L2Expanded                                  */
  int32_t  n_clusters;            /* = 8                                                     */
  int32_t  init;                  /* = CUML_B200_INIT_KMeansPlusPlus                         */
  int32_t  max_iter;              /* = 300                                                   */
  double   tol;                   /* = 1e-4; raw sum ||dc||^2 < tol; tol <= 0 never stops    */
  int32_t  verbosity;             /* = 2 (rapids_logger::level_enum::info; <= 1 logs every iteration) */
  uint64_t rng_seed;              /* = 0   raft::random::RngState::seed                      */
  uint64_t rng_base_subsequence;  /* = 0   raft::random::RngState::base_subsequence          */
  int32_t  rng_type;              /* = 0   GenPhilox (only Philox is implemented)            */
  int32_t  n_init;                /* = 1                                                     */
  double   oversampling_factor;   /* = 2.0; 0 => sequential k-means++                        */
  int32_t  batch_samples;         /* = 1<<15 (accepted; tiling is chosen by the engine)      */
  int32_t  batch_centroids;       /* = 0                                                     */
  int64_t  init_size;             /* = 0; out-of-core path: rows sampled for seeding (0 => min(3k, n)) */
  int64_t  device_buffer_samples; /* = 0; > 0: host partitions with more rows are streamed in batches this big */
} cuml_b200_kmeans_params_t;

CUML_B200_API void cuml_b200_kmeans_params_default(cuml_b200_kmeans_params_t* p);

/* ---- handle: the sliver of raft::handle_t the path uses (stream + injected NCCL comm) ----
 * reference: raft::handle_t forward-declared at cpp/include/cuml/cluster/kmeans.hpp:11-13;
 * "NCCL communicator that must be initialized on handle" kmeans.hpp:86-90.                  */
typedef struct cuml_b200_handle cuml_b200_handle_t;

/* stream: cudaStream_t (NULL = a stream owned by the handle).  comm: ncclComm_t or NULL.    */
CUML_B200_API int cuml_b200_handle_create(cuml_b200_handle_t** out, void* stream, void* nccl_comm,
                                          int rank, int n_ranks);
CUML_B200_API int cuml_b200_handle_destroy(cuml_b200_handle_t* h);
CUML_B200_API int cuml_b200_handle_sync(cuml_b200_handle_t* h);      /* handle.sync_stream() */
CUML_B200_API void* cuml_b200_handle_stream(cuml_b200_handle_t* h);
CUML_B200_API const char* cuml_b200_last_error(void);
CUML_B200_API const char* cuml_b200_version(void);

/* NCCL bring-up helpers (the raft_dask.common.comms.Comms role, reference
 * python/cuml/cuml/dask/cluster/kmeans.py:189-190): rank 0 makes a unique id, the host side
 * broadcasts its 128 bytes by any means (torch.distributed store), every rank inits.        */
CUML_B200_API int cuml_b200_nccl_unique_id(void* id_out_128_bytes);
CUML_B200_API int cuml_b200_handle_init_comm(cuml_b200_handle_t* h, const void* id_128_bytes,
                                             int rank, int n_ranks);

/* Peer-memory communicator: the library's own collectives over NVLink / NVSwitch peer mappings (or CUDA IPC on one
 * device) instead of NCCL -- the same raft::comms role (allreduce / allgather / bcast as the reference's multi-GPU
 * fit uses them, cpp/include/cuml/cluster/kmeans.hpp:86-90), with the per-iteration all-reduce fused into the centroid
 * update.  Every rank (one process per rank) creates its exchange window and gets a 64-byte CUDA IPC handle; the host
 * side gathers the n_ranks handles in rank order by any means and every rank attaches.  slot_bytes = 0 picks the
 * default (2 MiB per rank and parity).  After attach the handle's collectives no longer use NCCL.                  */
CUML_B200_API int cuml_b200_peer_window_create(cuml_b200_handle_t* h, int n_ranks, size_t slot_bytes,
                                               void* ipc_handle_out_64_bytes);
CUML_B200_API int cuml_b200_peer_window_attach(cuml_b200_handle_t* h, const void* all_ipc_handles_n_ranks_x_64_bytes,
                                               int rank, int n_ranks);

/* ---- fit: single array.  X [n,d] row-major host or device (auto-detected like
 * ML::is_device_or_managed_type, reference cpp/src/ml_cuda_utils.h:21-33), sample_weight [n] or
 * NULL (same residency), centroids [k,d] DEVICE in/out.
 * Replaces ML::kmeans::fit overloads, reference kmeans.hpp:41-79 (impl kmeans_fit.cu:101-231). */
CUML_B200_API int cuml_b200_kmeans_fit_f32_i32(cuml_b200_handle_t*, const cuml_b200_kmeans_params_t*,
    const float* X, int32_t n_samples, int32_t n_features, const float* sample_weight,
    float* centroids, float* inertia, int32_t* n_iter);
CUML_B200_API int cuml_b200_kmeans_fit_f64_i32(cuml_b200_handle_t*, const cuml_b200_kmeans_params_t*,
    const double* X, int32_t n_samples, int32_t n_features, const double* sample_weight,
    double* centroids, double* inertia, int32_t* n_iter);
CUML_B200_API int cuml_b200_kmeans_fit_f32_i64(cuml_b200_handle_t*, const cuml_b200_kmeans_params_t*,
    const float* X, int64_t n_samples, int64_t n_features, const float* sample_weight,
    float* centroids, float* inertia, int64_t* n_iter);
CUML_B200_API int cuml_b200_kmeans_fit_f64_i64(cuml_b200_handle_t*, const cuml_b200_kmeans_params_t*,
    const double* X, int64_t n_samples, int64_t n_features, const double* sample_weight,
    double* centroids, double* inertia, int64_t* n_iter);

/* ---- fit: partition list (multi-GPU row shards / out-of-core).  Cross-rank reduction over the
 * handle's NCCL communicator.  Replaces reference kmeans.hpp:110-130 (impl kmeans_fit.cu:23-99,
 * 237-318).                                                                                  */
CUML_B200_API int cuml_b200_kmeans_fit_parts_f32(cuml_b200_handle_t*, const cuml_b200_kmeans_params_t*,
    const float* const* X_parts, const int64_t* n_samples_parts, int64_t n_parts, int64_t n_features,
    const float* const* sample_weight_parts, float* centroids, float* inertia, int64_t* n_iter);
CUML_B200_API int cuml_b200_kmeans_fit_parts_f64(cuml_b200_handle_t*, const cuml_b200_kmeans_params_t*,
    const double* const* X_parts, const int64_t* n_samples_parts, int64_t n_parts, int64_t n_features,
    const double* const* sample_weight_parts, double* centroids, double* inertia, int64_t* n_iter);

/* fit + the labels of its own final assignment pass: labels_parts[i] (DEVICE int32 [n_samples_parts[i]], entries may be
 * NULL) receives partition i's labels.  The reference's estimator runs a second, redundant E-step after every fit to
 * obtain labels_ (python/cuml/cuml/cluster/kmeans.pyx:803-812); the fit already computed them for the inertia.
 * `inertia` is the fit's inertia with weights normalised to sum(w) = n_samples, as for the calls above.             */
CUML_B200_API int cuml_b200_kmeans_fit_parts_labels_f32(cuml_b200_handle_t*, const cuml_b200_kmeans_params_t*,
    const float* const* X_parts, const int64_t* n_samples_parts, int64_t n_parts, int64_t n_features,
    const float* const* sample_weight_parts, float* centroids, float* inertia, int64_t* n_iter,
    int32_t* const* labels_parts);
CUML_B200_API int cuml_b200_kmeans_fit_parts_labels_f64(cuml_b200_handle_t*, const cuml_b200_kmeans_params_t*,
    const double* const* X_parts, const int64_t* n_samples_parts, int64_t n_parts, int64_t n_features,
    const double* const* sample_weight_parts, double* centroids, double* inertia, int64_t* n_iter,
    int32_t* const* labels_parts);

/* ---- predict: nearest centroid + weighted inertia.  All pointers DEVICE; labels dtype = index
 * type.  Replaces reference kmeans.hpp:154-195 (impl kmeans_predict.cu:19-135).              */
CUML_B200_API int cuml_b200_kmeans_predict_f32_i32(cuml_b200_handle_t*, const cuml_b200_kmeans_params_t*,
    const float* centroids, const float* X, int32_t n_samples, int32_t n_features,
    const float* sample_weight, int normalize_weights, int32_t* labels, float* inertia);
CUML_B200_API int cuml_b200_kmeans_predict_f64_i32(cuml_b200_handle_t*, const cuml_b200_kmeans_params_t*,
    const double* centroids, const double* X, int32_t n_samples, int32_t n_features,
    const double* sample_weight, int normalize_weights, int32_t* labels, double* inertia);
CUML_B200_API int cuml_b200_kmeans_predict_f32_i64(cuml_b200_handle_t*, const cuml_b200_kmeans_params_t*,
    const float* centroids, const float* X, int64_t n_samples, int64_t n_features,
    const float* sample_weight, int normalize_weights, int64_t* labels, float* inertia);
CUML_B200_API int cuml_b200_kmeans_predict_f64_i64(cuml_b200_handle_t*, const cuml_b200_kmeans_params_t*,
    const double* centroids, const double* X, int64_t n_samples, int64_t n_features,
    const double* sample_weight, int normalize_weights, int64_t* labels, double* inertia);

/* ---- transform: X_new [n,k] distances under params->metric (squared for L2Expanded).
 * Replaces reference kmeans.hpp:213-242 (impl kmeans_transform.cu:18-78).                    */
CUML_B200_API int cuml_b200_kmeans_transform_f32_i32(cuml_b200_handle_t*, const cuml_b200_kmeans_params_t*,
    const float* centroids, const float* X, int32_t n_samples, int32_t n_features, float* X_new);
CUML_B200_API int cuml_b200_kmeans_transform_f64_i32(cuml_b200_handle_t*, const cuml_b200_kmeans_params_t*,
    const double* centroids, const double* X, int32_t n_samples, int32_t n_features, double* X_new);
CUML_B200_API int cuml_b200_kmeans_transform_f32_i64(cuml_b200_handle_t*, const cuml_b200_kmeans_params_t*,
    const float* centroids, const float* X, int64_t n_samples, int64_t n_features, float* X_new);
CUML_B200_API int cuml_b200_kmeans_transform_f64_i64(cuml_b200_handle_t*, const cuml_b200_kmeans_params_t*,
    const double* centroids, const double* X, int64_t n_samples, int64_t n_features, double* X_new);

/* ---- measurement / test hooks (not part of the reference surface) -------------------------
 * One Lloyd iteration on device-resident data: E-step (fused distance+argmin) then M-step
 * (sums / weights / inertia) then centroid update; used by bench.py to time the hot path with
 * CUDA events and by the parity tests for single-step checks (SURVEY.md 8c protocol (ii)).
 * labels: int32 [n] device (out).  sums_out: double [k*d + k + 1] device (out; S | W | inertia
 * wrt the input centroids), may be NULL.  centroids updated in place.  shift2_out: device
 * double, may be NULL.  engine: 0 = auto, 1 = force SIMT fp32 path, 2 = force tcgen05 path.  */
CUML_B200_API int cuml_b200_kmeans_lloyd_step_f32(cuml_b200_handle_t*, const float* X, int64_t n_samples,
    int64_t n_features, const float* sample_weight, int32_t n_clusters, float* centroids,
    int32_t* labels, double* sums_out, double* shift2_out, int engine);
/* E-step only (labels, optional per-row min distance surrogate); engine as above.            */
CUML_B200_API int cuml_b200_kmeans_assign_f32(cuml_b200_handle_t*, const float* X, int64_t n_samples,
    int64_t n_features, int32_t n_clusters, const float* centroids, int32_t* labels, int engine);
/* Test hook for the tensor-core engine: labels plus the raw x.c accumulators as a
 * [n, k_pad] matrix (k_pad returned; call once with dots == NULL to size the buffer).        */
CUML_B200_API int cuml_b200_kmeans_debug_dots_f32(cuml_b200_handle_t*, const float* X, int64_t n_samples,
    int64_t n_features, int32_t n_clusters, const float* centroids, int32_t* labels, float* dots,
    int64_t* k_pad_out);
/* Counters: number of kernels this library launched on the calling thread since the last
 * reset (bench.py's gpu_launches), and the average duration in ms of the dominant kernel
 * recorded with CUDA events on the handle's stream when timing is enabled.                   */
CUML_B200_API void    cuml_b200_launch_count_reset(void);
CUML_B200_API int64_t cuml_b200_launch_count(void);
CUML_B200_API int     cuml_b200_kernel_timing_enable(cuml_b200_handle_t*, int enable);
CUML_B200_API int     cuml_b200_kernel_timing_read(cuml_b200_handle_t*, double* fused_ms_total,
                                                   int64_t* fused_launches, double* update_ms_total,
                                                   int64_t* update_launches);
/* 1 if the tcgen05 engine supports (n_features, n_clusters) for fp32.                        */
CUML_B200_API int cuml_b200_kmeans_tc_supported(int64_t n_features, int32_t n_clusters);
/* 1 if an unweighted fp32 fit of this shape runs the E-step and the M-step as ONE kernel (one pass over X per
 * iteration: n_features = 16, n_clusters <= 64).  Used by bench.py to label its roofline. */
CUML_B200_API int cuml_b200_kmeans_fused_update(cuml_b200_handle_t* handle, int64_t n_features, int32_t n_clusters);
/* Which E-step kernel a fit of this shape takes on the handle's device: 0 CUDA-core fp32, 1 tcgen05 one CTA
 * per tile (3xTF32), 2 tcgen05 CTA pair (3xTF32), 3 tcgen05 CTA pair (tf32 main term + two bf16 correction
 * terms), 4 tcgen05 with the X operand in tensor memory (opt-in), 5 tcgen05 one CTA per tile with the bf16
 * correction terms (opt-in).  Used by bench.py for its roofline. */
CUML_B200_API int cuml_b200_kmeans_estep_variant(cuml_b200_handle_t* handle, int64_t n_features,
                                                 int32_t n_clusters);

#ifdef __cplusplus
}
#endif
#endif /* CUML_B200_KMEANS_C_H */
