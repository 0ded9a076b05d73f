// fit / predict / transform drivers and the extern "C" boundary (include/cuml_b200/kmeans_c.h).
//
// Mirrors the reference's forwarding layer: pointer residency detection and host->device staging
// (cpp/src/kmeans/kmeans_fit.cu:101-231, cpp/src/ml_cuda_utils.h:21-33), the partition-list
// overload (kmeans_fit.cu:23-99,237-318), predict (kmeans_predict.cu:19-135) and transform
// (kmeans_transform.cu:18-78).  Everything below those shims -- which in the reference is the
// un-vendored cuVS -- is this library's own CUDA code.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <limits>
#include <random>
#include <unordered_set>

#include "lloyd.cuh"

namespace cb2 {

Handle* make_handle(void* stream, void* comm, int rank, int n_ranks);
void free_handle(Handle* h);

namespace {

thread_local std::string t_last_error;

template <typename F>
int guarded(F&& f)
{
  try {
    f();
    return CUML_B200_SUCCESS;
  } catch (const Error& e) {
    t_last_error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    t_last_error = e.what();
    return CUML_B200_INTERNAL_ERROR;
  } catch (...) {
    t_last_error = "unknown error";
    return CUML_B200_INTERNAL_ERROR;
  }
}

void check_params(const cuml_b200_kmeans_params_t& p)
{
  CB2_EXPECTS(p.n_clusters > 0, "n_clusters=" + std::to_string(p.n_clusters) + " should be a positive integer.");
  CB2_EXPECTS(p.max_iter >= 0, "max_iter must be >= 0");
  CB2_EXPECTS(p.metric == CUML_B200_L2Expanded || p.metric == CUML_B200_L2SqrtExpanded,
              "only L2Expanded / L2SqrtExpanded metrics are supported by k-means");
  CB2_EXPECTS(p.init >= 0 && p.init <= 2, "invalid init method");
  CB2_EXPECTS(p.n_init >= 1, "n_init must be >= 1");
  CB2_EXPECTS(p.oversampling_factor >= 0.0, "oversampling_factor must be >= 0");
  CB2_EXPECTS(p.device_buffer_samples >= 0, "device_buffer_samples must be >= 0");
  // raft::random::GeneratorType: GenPhilox = 0, GenPC = 1.  The seeded inits draw from a Philox4x32-10 keyed by the
  // global row index (what makes a sharded fit draw what the single-GPU fit draws); a PCG request changes results in
  // the reference, so it is refused here rather than silently served by Philox.
  CB2_EXPECTS(p.init == CUML_B200_INIT_Array || p.rng_type == 0,
              "rng_state.type: only the Philox generator (GenPhilox) is implemented for the seeded inits");
}

// verbosity <= debug (rapids_logger::level_enum numbering: trace 0, debug 1, info 2 ...): one line per Lloyd iteration
// with the squared centroid shift, as cuVS logs it at debug level.  Reads one double back (synchronises).
double log_iteration(Handle& h, const cuml_b200_kmeans_params_t& params, int run, int64_t iter, const double* shift2_dev)
{
  CB2_CUDA(cudaMemcpyAsync(h.pinned, shift2_dev, sizeof(double), cudaMemcpyDeviceToHost, h.stream));
  CB2_CUDA(cudaStreamSynchronize(h.stream));
  if (h.rank == 0)
    std::fprintf(stderr, "[cuml_b200] [debug] KMeans run %d iteration %lld: squared centroid shift %.6e (tol %.3e)\n", run,
                 static_cast<long long>(iter), h.pinned[0], params.tol);
  return h.pinned[0];
}

// device-resident view of caller partitions; host partitions are staged into owned buffers
template <typename T>
struct Staged {
  std::vector<Part<T>> parts;
  std::vector<DevBuf<T>> owned;
};

template <typename T>
void stage_parts(Handle& h, const T* const* X_parts, const int64_t* n_parts_rows, int64_t n_parts, int64_t d,
                 const T* const* w_parts, Staged<T>& out)
{
  // residency is decided by the first non-empty partition (reference kmeans_fit.cu:237-246)
  bool on_device = true;
  for (int64_t i = 0; i < n_parts; ++i) {
    if (n_parts_rows[i] > 0) {
      on_device = is_device_pointer(X_parts[i]);
      break;
    }
  }
  for (int64_t i = 0; i < n_parts; ++i) {
    const int64_t n = n_parts_rows[i];
    CB2_EXPECTS(n >= 0, "negative partition size");
    if (n == 0) continue;
    CB2_EXPECTS(X_parts[i] != nullptr, "null partition pointer");
    const T* wp = w_parts ? w_parts[i] : nullptr;
    if (on_device) {
      out.parts.push_back(Part<T>{X_parts[i], n, wp});
    } else {
      // host-resident and within device_buffer_samples (or no buffer size given): staged whole; larger inputs
      // take fit_streamed() above
      out.owned.emplace_back(static_cast<size_t>(n) * d, h.stream);
      T* dx = out.owned.back().get();
      CB2_CUDA(cudaMemcpyAsync(dx, X_parts[i], sizeof(T) * n * d, cudaMemcpyHostToDevice, h.stream));
      T* dw = nullptr;
      if (wp) {
        out.owned.emplace_back(static_cast<size_t>(n), h.stream);
        dw = out.owned.back().get();
        CB2_CUDA(cudaMemcpyAsync(dw, wp, sizeof(T) * n, cudaMemcpyHostToDevice, h.stream));
      }
      out.parts.push_back(Part<T>{dx, n, dw});
    }
  }
}

// sum / max of a few host scalars over the ranks (one tiny device all-reduce + one synchronisation)
void allreduce_host(Handle& h, double* v, int count, bool take_max = false)
{
  if (h.n_ranks <= 1) return;
  DevBuf<double> cell(static_cast<size_t>(count), h.stream);
  CB2_CUDA(cudaMemcpyAsync(cell.get(), v, sizeof(double) * count, cudaMemcpyHostToDevice, h.stream));
  if (take_max) comms::allreduce_max_f64(h, cell.get(), count);
  else comms::allreduce_sum_f64(h, cell.get(), count);
  CB2_CUDA(cudaMemcpyAsync(v, cell.get(), sizeof(double) * count, cudaMemcpyDeviceToHost, h.stream));
  CB2_CUDA(cudaStreamSynchronize(h.stream));
}

int64_t allreduce_i64_host(Handle& h, int64_t v)
{
  double dv = static_cast<double>(v);
  allreduce_host(h, &dv, 1);
  return static_cast<int64_t>(dv + 0.5);
}

// Every rank passes its local verdict ("" = fine).  If any rank failed, ALL ranks throw: a rank-local argument error
// must not leave the other ranks waiting in the next collective (reference: the all-worker preflight of
// dask/cluster/kmeans.py:97-113 plays this role on the client).
void collective_agree(Handle& h, const std::string& local_error)
{
  double failed = local_error.empty() ? 0.0 : 1.0;
  allreduce_host(h, &failed, 1);
  if (failed > 0.0) {
    if (!local_error.empty()) throw Error(CUML_B200_INVALID_ARGUMENT, local_error);
    throw Error(CUML_B200_INVALID_ARGUMENT, "k-means fit: " + std::to_string(static_cast<int>(failed + 0.5)) +
                                              " other rank(s) rejected their arguments; no rank proceeds");
  }
}

// exclusive prefix of n_local over ranks
int64_t rank_row_offset(Handle& h, int64_t n_local)
{
  if (h.n_ranks <= 1) return 0;
  DevBuf<int64_t> sb(1, h.stream), rb(h.n_ranks, h.stream);
  CB2_CUDA(cudaMemcpyAsync(sb.get(), &n_local, sizeof(int64_t), cudaMemcpyHostToDevice, h.stream));
  comms::allgather_bytes(h, sb.get(), rb.get(), sizeof(int64_t));
  std::vector<int64_t> cnt(h.n_ranks);
  CB2_CUDA(cudaMemcpyAsync(cnt.data(), rb.get(), sizeof(int64_t) * h.n_ranks, cudaMemcpyDeviceToHost, h.stream));
  CB2_CUDA(cudaStreamSynchronize(h.stream));
  int64_t off = 0;
  for (int i = 0; i < h.rank; ++i) off += cnt[i];
  return off;
}

// what a rank learns about its own arguments before any collective
struct Preflight {
  int64_t n_local   = 0;
  bool host_rows    = false;   // first non-empty partition is host-resident (reference kmeans_fit.cu:237-246)
  bool device_rows  = false;
  bool want_stream  = false;   // host-resident and more rows than device_buffer_samples
};

template <typename T>
Preflight inspect_parts(Handle& h, const cuml_b200_kmeans_params_t& params, const T* const* X_parts,
                        const int64_t* n_parts_rows, int64_t n_parts, int64_t d, const T* centroids)
{
  check_params(params);
  CB2_EXPECTS(d >= 1 && d <= std::numeric_limits<int>::max(), "n_features out of range");
  CB2_EXPECTS(centroids != nullptr && is_device_pointer(centroids), "centroids must be device accessible");
  Preflight pf;
  bool first = true;
  for (int64_t i = 0; i < n_parts; ++i) {
    CB2_EXPECTS(n_parts_rows[i] >= 0, "negative partition size");
    if (n_parts_rows[i] == 0) continue;
    CB2_EXPECTS(X_parts[i] != nullptr, "null partition pointer");
    if (first) {
      (is_device_pointer(X_parts[i]) ? pf.device_rows : pf.host_rows) = true;
      first = false;
    }
    pf.n_local += n_parts_rows[i];
  }
  pf.want_stream = pf.host_rows && params.device_buffer_samples > 0 && pf.n_local > params.device_buffer_samples;
  if (params.init == CUML_B200_INIT_Random && h.n_ranks > 1) {
    // the reference's preflight rule (kmeans_mg.py:63-81), checked here so that it fails on every rank together
    const int S    = std::min(h.n_ranks, params.n_clusters);
    const int mine = (h.rank < S) ? params.n_clusters / S + (h.rank == 0 ? params.n_clusters % S : 0) : 0;
    CB2_EXPECTS(pf.n_local >= mine, "init='random' requires rank " + std::to_string(h.rank) + " to sample " +
                                      std::to_string(mine) + " initial centroid(s), but this rank only has " +
                                      std::to_string(pf.n_local) + " row(s)");
  }
  if (params.init == CUML_B200_INIT_KMeansPlusPlus && params.oversampling_factor == 0.0)
    CB2_EXPECTS(h.n_ranks == 1, "init='k-means++' or oversampling_factor=0 not supported for multi-GPU KMeans");
  return pf;
}

// `rows` distinct row indices in [0, n), sorted (Floyd's algorithm; the init_size sample)
std::vector<int64_t> sample_rows(int64_t n, int64_t rows, uint64_t seed)
{
  std::vector<int64_t> out;
  if (rows >= n) {
    out.resize(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) out[static_cast<size_t>(i)] = i;
    return out;
  }
  std::mt19937_64 gen(seed * 0xD1B54A32D192ED03ull + 0x9E3779B97F4A7C15ull);
  std::unordered_set<int64_t> chosen;
  chosen.reserve(static_cast<size_t>(rows) * 2);
  for (int64_t j = n - rows; j < n; ++j) {
    const int64_t t = static_cast<int64_t>(gen() % static_cast<uint64_t>(j + 1));
    if (!chosen.insert(t).second) chosen.insert(j);
  }
  out.assign(chosen.begin(), chosen.end());
  std::sort(out.begin(), out.end());
  return out;
}

// ---- out-of-core fit: host-resident partitions streamed through two device buffers ------------------
// Role of the reference's host-data path (kmeans_fit.cu:167-231 with host X -> cuVS batched fit,
// KMeansParams::device_buffer_samples / init_size, kmeans.pyx:546-564): every Lloyd iteration walks the host partitions
// in batches of `device_buffer_samples` rows -- H2D copy, E-step, M-step accumulation into the same packed sums -- so X
// never has to fit device memory.  The copy of batch b + 1 (second stream, second buffer) overlaps the kernels of
// batch b.  Seeding (other than init=Array) runs on a random sample of `init_size` rows (0: min(3 k, n), the
// reference's documented default), drawn with the fit's seed.  labels_parts (optional): per-partition DEVICE int32
// arrays that receive the labels of the final pass.
template <typename T>
void fit_streamed(Handle& h, const cuml_b200_kmeans_params_t& params, const T* const* X_parts,
                  const int64_t* n_parts_rows, int64_t n_parts, int64_t d, const T* const* w_parts, T* centroids,
                  T& inertia_out, int64_t& n_iter_out, int64_t n_local, int32_t* const* labels_parts)
{
  NvtxRange nvtx_fit("cuml_b200::kmeans::fit (out-of-core)");
  const int k  = params.n_clusters;
  const int di = static_cast<int>(d);
  const int64_t batch    = params.device_buffer_samples;
  const int64_t n_global = allreduce_i64_host(h, n_local);
  bool weighted = false;
  for (int64_t i = 0; i < n_parts; ++i) weighted = weighted || (w_parts && w_parts[i] && n_parts_rows[i] > 0);
  double wflags[2] = {weighted ? 1.0 : 0.0, 0.0};
  if (weighted)
    for (int64_t i = 0; i < n_parts; ++i)
      for (int64_t r = 0; r < n_parts_rows[i]; ++r) wflags[1] += (w_parts && w_parts[i]) ? static_cast<double>(w_parts[i][r]) : 1.0;
  else
    wflags[1] = static_cast<double>(n_local);
  allreduce_host(h, wflags, 2);
  weighted = wflags[0] > 0.0;                     // weighted on any rank => weighted everywhere
  std::string err;
  if (n_global < k) err = "n_samples=" + std::to_string(n_global) + " should be >= n_clusters=" + std::to_string(k) + ".";
  else if (weighted && !(wflags[1] > 0.0)) err = "sample weights must have a positive sum";
  if (!err.empty()) throw Error(CUML_B200_INVALID_ARGUMENT, err);   // same verdict on every rank: global quantities
  const double wscale = weighted ? static_cast<double>(n_global) / wflags[1] : 1.0;

  // two staging buffers; copies on the handle's second stream, kernels on its main stream
  if (!h.aux_stream) CB2_CUDA(cudaStreamCreateWithFlags(&h.aux_stream, cudaStreamNonBlocking));
  cudaEvent_t copied[2], consumed[2];
  for (int i = 0; i < 2; ++i) {
    CB2_CUDA(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
    CB2_CUDA(cudaEventCreateWithFlags(&consumed[i], cudaEventDisableTiming));
  }
  struct EventGuard {
    cudaEvent_t* a; cudaEvent_t* b;
    ~EventGuard() { for (int i = 0; i < 2; ++i) { cudaEventDestroy(a[i]); cudaEventDestroy(b[i]); } }
  } ev_guard{copied, consumed};
  DevBuf<T> xb[2], wb[2];
  for (int i = 0; i < 2; ++i) {
    xb[i].alloc(static_cast<size_t>(batch) * d, h.stream);
    if (weighted) wb[i].alloc(static_cast<size_t>(batch), h.stream);
  }
  std::vector<T> ones;   // partitions without weights inside a weighted fit
  int64_t batch_no = 0;
  // enqueue the copy of `rows` rows of partition `pi` from `off` into staging buffer s (after its previous reader)
  auto stage = [&](int s, int64_t pi, int64_t off, int64_t rows, bool wait_consumed) {
    if (wait_consumed) CB2_CUDA(cudaStreamWaitEvent(h.aux_stream, consumed[s], 0));
    CB2_CUDA(cudaMemcpyAsync(xb[s].get(), X_parts[pi] + off * d, sizeof(T) * rows * d, cudaMemcpyHostToDevice, h.aux_stream));
    if (weighted) {
      if (w_parts && w_parts[pi]) {
        CB2_CUDA(cudaMemcpyAsync(wb[s].get(), w_parts[pi] + off, sizeof(T) * rows, cudaMemcpyHostToDevice, h.aux_stream));
      } else {
        if (ones.size() < static_cast<size_t>(rows)) ones.assign(static_cast<size_t>(batch), T(1));
        CB2_CUDA(cudaMemcpyAsync(wb[s].get(), ones.data(), sizeof(T) * rows, cudaMemcpyHostToDevice, h.aux_stream));
      }
    }
    CB2_CUDA(cudaEventRecord(copied[s], h.aux_stream));
  };
  // calls f(buffer, partition, offset, rows) for every batch of this rank, in order, with the batch already staged
  auto for_each_batch = [&](auto&& f) {
    for (int64_t pi = 0; pi < n_parts; ++pi)
      for (int64_t off = 0; off < n_parts_rows[pi]; off += batch, ++batch_no) {
        const int s        = static_cast<int>(batch_no & 1);
        const int64_t rows = std::min(batch, n_parts_rows[pi] - off);
        stage(s, pi, off, rows, batch_no >= 2);
        CB2_CUDA(cudaStreamWaitEvent(h.stream, copied[s], 0));
        f(s, pi, off, rows);
        CB2_CUDA(cudaEventRecord(consumed[s], h.stream));
      }
  };

  std::vector<Part<T>> buf_part{Part<T>{xb[0].get(), batch, weighted ? wb[0].get() : nullptr}};
  LloydSolver<T> solver(h, buf_part, di, k, ENGINE_AUTO);

  // seeding sample (init_size role): random rows of this rank, proportional to its share of the data
  DevBuf<T> seed_x, seed_w;
  int64_t n_seed = 0;
  if (params.init != CUML_B200_INIT_Array) {
    const int64_t want_global = params.init_size > 0 ? params.init_size : std::min<int64_t>(3 * static_cast<int64_t>(k), n_global);
    n_seed = std::min<int64_t>(n_local, ceil_div(want_global * n_local, std::max<int64_t>(n_global, 1)));
    const std::vector<int64_t> pick = sample_rows(n_local, n_seed, params.rng_seed + static_cast<uint64_t>(h.rank));
    std::vector<T> hx(static_cast<size_t>(n_seed) * d), hw(weighted ? static_cast<size_t>(n_seed) : 0);
    int64_t pi = 0, base = 0;
    for (int64_t s_i = 0; s_i < n_seed; ++s_i) {
      const int64_t g = pick[static_cast<size_t>(s_i)];
      while (g >= base + n_parts_rows[pi]) base += n_parts_rows[pi++];
      std::memcpy(hx.data() + static_cast<size_t>(s_i) * d, X_parts[pi] + (g - base) * d, sizeof(T) * d);
      if (weighted) hw[s_i] = (w_parts && w_parts[pi]) ? w_parts[pi][g - base] : T(1);
    }
    seed_x.alloc(std::max<size_t>(hx.size(), 1), h.stream);
    CB2_CUDA(cudaMemcpyAsync(seed_x.get(), hx.data(), sizeof(T) * hx.size(), cudaMemcpyHostToDevice, h.stream));
    if (weighted) {
      seed_w.alloc(std::max<size_t>(hw.size(), 1), h.stream);
      CB2_CUDA(cudaMemcpyAsync(seed_w.get(), hw.data(), sizeof(T) * hw.size(), cudaMemcpyHostToDevice, h.stream));
    }
    CB2_CUDA(cudaStreamSynchronize(h.stream));   // hx / hw go out of scope
  }
  const int64_t seed_global = (params.init != CUML_B200_INIT_Array) ? allreduce_i64_host(h, n_seed) : 0;
  if (params.init != CUML_B200_INIT_Array && seed_global < k)
    throw Error(CUML_B200_INVALID_ARGUMENT, "init_size=" + std::to_string(seed_global) + " should be >= n_clusters=" + std::to_string(k) + ".");
  std::vector<Part<T>> seed_parts;
  if (n_seed > 0) seed_parts.push_back(Part<T>{seed_x.get(), n_seed, weighted ? seed_w.get() : nullptr});
  SeedContext<T> sctx{h, seed_parts, di, n_seed, seed_global,
                      (params.init != CUML_B200_INIT_Array) ? rank_row_offset(h, n_seed) : 0, params.rng_seed, ENGINE_AUTO};

  double* packed     = solver.packed();
  const size_t count = solver.packed_count();
  const int n_init   = (params.init == CUML_B200_INIT_Array) ? 1 : params.n_init;
  DevBuf<T> trial(static_cast<size_t>(k) * di, h.stream);
  double best_inertia = std::numeric_limits<double>::infinity();
  int64_t best_iter   = 0;
  // final pass of a run: labels (kept when this run is the best so far) and cost with the final centroids
  auto final_pass = [&](T* C, bool keep_labels) {
    solver.prepare(C);
    bool first = true;
    for_each_batch([&](int s, int64_t pi, int64_t off, int64_t rows) {
      solver.set_part(xb[s].get(), weighted ? wb[s].get() : nullptr, rows);
      solver.assign_one(C, xb[s].get(), rows, solver.labels(0));
      solver.inertia_only(C, !first);
      if (keep_labels && labels_parts && labels_parts[pi])
        CB2_CUDA(cudaMemcpyAsync(labels_parts[pi] + off, solver.labels(0), sizeof(int32_t) * rows, cudaMemcpyDeviceToDevice, h.stream));
      first = false;
    });
    if (first) CB2_CUDA(cudaMemsetAsync(packed + count - 1, 0, sizeof(double), h.stream));   // a rank without rows
    double* cell = packed + count - 1;
    comms::allreduce_sum_f64(h, cell, 1);
    CB2_CUDA(cudaMemcpyAsync(h.pinned, cell, sizeof(double), cudaMemcpyDeviceToHost, h.stream));
    CB2_CUDA(cudaStreamSynchronize(h.stream));
    return h.pinned[0] * wscale;
  };
  for (int run = 0; run < n_init; ++run) {
    T* C = (n_init == 1) ? centroids : trial.get();
    sctx.seed = params.rng_seed + 0x9E3779B97F4A7C15ull * static_cast<uint64_t>(run);
    {
      NvtxRange nvtx_seed("cuml_b200::kmeans::init");
      if (params.init == CUML_B200_INIT_Random) init_random<T>(sctx, k, C);
      else if (params.init != CUML_B200_INIT_Array && params.oversampling_factor == 0.0) init_kmeans_plus_plus<T>(sctx, k, C);
      else if (params.init != CUML_B200_INIT_Array) init_scalable<T>(sctx, params, C);
    }
    int64_t iters = 0;
    {
      NvtxRange nvtx_lloyd("cuml_b200::kmeans::lloyd (streamed batches)");
      while (iters < params.max_iter) {
        solver.prepare(C);
        bool first = true;
        for_each_batch([&](int s, int64_t, int64_t, int64_t rows) {
          solver.set_part(xb[s].get(), weighted ? wb[s].get() : nullptr, rows);
          solver.assign_one(C, xb[s].get(), rows, solver.labels(0));
          solver.accumulate(C, false, !first);
          first = false;
        });
        if (first) CB2_CUDA(cudaMemsetAsync(packed, 0, count * sizeof(double), h.stream));   // a rank without rows
        exchange_and_finalize<T>(h, packed, count, C, k, di);
        ++iters;
        if (params.verbosity <= 1)
          log_iteration(h, params, run, iters, packed + count);
        if (params.tol > 0.0) {
          CB2_CUDA(cudaMemcpyAsync(h.pinned, packed + count, sizeof(double), cudaMemcpyDeviceToHost, h.stream));
          CB2_CUDA(cudaStreamSynchronize(h.stream));
          if (h.pinned[0] < params.tol) break;
        }
      }
    }
    NvtxRange nvtx_final("cuml_b200::kmeans::final assign + inertia");
    const bool last_run = run == n_init - 1;
    double inertia      = final_pass(C, /*keep_labels=*/n_init == 1);
    if (inertia < best_inertia || run == 0) {
      best_inertia = inertia;
      best_iter    = iters;
      if (C != centroids)
        CB2_CUDA(cudaMemcpyAsync(centroids, C, sizeof(T) * k * di, cudaMemcpyDeviceToDevice, h.stream));
    }
    if (last_run && n_init > 1 && labels_parts) final_pass(centroids, true);   // labels of the winning run
  }
  CB2_CUDA(cudaStreamSynchronize(h.stream));
  CB2_CUDA(cudaStreamSynchronize(h.aux_stream));
  inertia_out = static_cast<T>(best_inertia);
  n_iter_out  = best_iter;
}

template <typename T>
void fit_parts_impl(Handle& h, const cuml_b200_kmeans_params_t& params, const T* const* X_parts,
                    const int64_t* n_parts_rows, int64_t n_parts, int64_t d, const T* const* w_parts, T* centroids,
                    T& inertia_out, int64_t& n_iter_out, int32_t* const* labels_parts = nullptr)
{
  CB2_CUDA(cudaSetDevice(h.device));
  // ---- rank-local validation, then one collective verdict (a multi-rank fit fails on every rank or on none) ----
  Preflight pf;
  std::string local_error;
  try {
    pf = inspect_parts<T>(h, params, X_parts, n_parts_rows, n_parts, d, centroids);
  } catch (const Error& e) {
    local_error = e.what();
  }
  collective_agree(h, local_error);
  const int k  = params.n_clusters;
  const int di = static_cast<int>(d);

  // streamed or staged is a GLOBAL decision: the two paths issue different collective sequences
  double sflags[2] = {pf.want_stream ? 1.0 : 0.0, pf.device_rows ? 1.0 : 0.0};
  allreduce_host(h, sflags, 2);
  if (sflags[0] > 0.0) {
    if (sflags[1] > 0.0)
      throw Error(CUML_B200_INVALID_ARGUMENT, "device_buffer_samples: host partitions are streamed on some ranks while others hold "
                                              "device-resident partitions; use one residency on all ranks");
    fit_streamed<T>(h, params, X_parts, n_parts_rows, n_parts, d, w_parts, centroids, inertia_out, n_iter_out, pf.n_local,
                    labels_parts);
    return;
  }
  NvtxRange nvtx_fit("cuml_b200::kmeans::fit");
  // CUML_B200_TRACE=1: host-timed phases (adds stream synchronisations; measurement aid only)
  static const bool trace = std::getenv("CUML_B200_TRACE") != nullptr;
  auto tnow = [&] {
    if (trace) cudaStreamSynchronize(h.stream);
    return std::chrono::steady_clock::now();
  };
  auto tms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  const auto t_0 = tnow();
  Staged<T> st;
  {
    NvtxRange nvtx_stage("cuml_b200::kmeans::stage host partitions");
    stage_parts<T>(h, X_parts, n_parts_rows, n_parts, d, w_parts, st);
  }
  const auto t_1 = tnow();
  if (trace) std::printf("[cuml_b200 trace] stage_parts %.1f ms\n", tms(t_0, t_1));
  int64_t n_local = 0;
  for (auto& p : st.parts) n_local += p.n;

  // global row count, weights normalised so that sum(w) == n_samples (cuVS checkWeight role; centroids are invariant to
  // the scale, so only the inertia is rescaled)
  bool weighted = false;
  for (auto& p : st.parts) weighted = weighted || (p.w != nullptr);
  double g[3] = {static_cast<double>(n_local), weighted ? 1.0 : 0.0, 0.0};
  if (weighted)
    for (auto& p : st.parts) g[2] += p.w ? sum_weights<T>(h, p.w, p.n) : static_cast<double>(p.n);
  else
    g[2] = static_cast<double>(n_local);
  allreduce_host(h, g, 3);
  const int64_t n_global = static_cast<int64_t>(g[0] + 0.5);
  weighted               = g[1] > 0.0;
  CB2_EXPECTS(n_global >= k, "n_samples=" + std::to_string(n_global) + " should be >= n_clusters=" + std::to_string(k) + ".");
  CB2_EXPECTS(!weighted || g[2] > 0.0, "sample weights must have a positive sum");
  const double wscale = weighted ? static_cast<double>(n_global) / g[2] : 1.0;

  LloydSolver<T> solver(h, st.parts, di, k, ENGINE_AUTO);
  SeedContext<T> sctx{h, st.parts, di, n_local, n_global, rank_row_offset(h, n_local), params.rng_seed, ENGINE_AUTO};

  const int n_init = (params.init == CUML_B200_INIT_Array) ? 1 : params.n_init;
  DevBuf<T> trial(static_cast<size_t>(k) * di, h.stream);
  double best_inertia = std::numeric_limits<double>::infinity();
  int64_t best_iter   = 0;
  // the labels of the final pass, written per caller partition (empty partitions were dropped by stage_parts)
  auto export_labels = [&] {
    if (!labels_parts) return;
    size_t sp = 0;
    for (int64_t i = 0; i < n_parts; ++i) {
      if (n_parts_rows[i] == 0) continue;
      if (labels_parts[i])
        CB2_CUDA(cudaMemcpyAsync(labels_parts[i], solver.labels(sp), sizeof(int32_t) * n_parts_rows[i], cudaMemcpyDeviceToDevice,
                                 h.stream));
      ++sp;
    }
  };
  for (int run = 0; run < n_init; ++run) {
    T* C = (n_init == 1) ? centroids : trial.get();
    sctx.seed = params.rng_seed + 0x9E3779B97F4A7C15ull * static_cast<uint64_t>(run);
    {
      NvtxRange nvtx_seed("cuml_b200::kmeans::init");
      if (params.init == CUML_B200_INIT_Array) {
        // centroids already hold the caller's initial centres
      } else if (params.init == CUML_B200_INIT_Random) {
        init_random<T>(sctx, k, C);
      } else if (params.oversampling_factor == 0.0) {
        init_kmeans_plus_plus<T>(sctx, k, C);
      } else {
        init_scalable<T>(sctx, params, C);
      }
    }
    const auto t_2 = tnow();
    int64_t iters = 0;
    {
      NvtxRange nvtx_lloyd("cuml_b200::kmeans::lloyd");
      if (params.verbosity <= 1) {   // debug / trace: one line per iteration (costs a synchronisation each)
        for (; iters < params.max_iter;) {
          solver.step(C);
          ++iters;
          const double shift2 = log_iteration(h, params, run, iters, solver.packed() + solver.packed_count());
          if (params.tol > 0.0 && shift2 < params.tol) break;
        }
      } else {
        iters = solver.run(C, params.max_iter, params.tol);
      }
    }
    const auto t_3 = tnow();
    // final E-step + cost with the final centroids
    NvtxRange nvtx_final("cuml_b200::kmeans::final assign + inertia");
    solver.assign(C);
    const double inertia = solver.inertia(C) * wscale;
    const auto t_4 = tnow();
    if (trace)
      std::printf("[cuml_b200 trace] setup+init %.1f ms, %lld Lloyd iterations %.1f ms, final assign+inertia %.1f ms\n",
                  tms(t_1, t_2), static_cast<long long>(iters), tms(t_2, t_3), tms(t_3, t_4));
    if (inertia < best_inertia || run == 0) {
      best_inertia = inertia;
      best_iter    = iters;
      if (C != centroids)
        CB2_CUDA(cudaMemcpyAsync(centroids, C, sizeof(T) * k * di, cudaMemcpyDeviceToDevice, h.stream));
      export_labels();
    }
  }
  CB2_CUDA(cudaStreamSynchronize(h.stream));
  inertia_out = static_cast<T>(best_inertia);
  n_iter_out  = best_iter;
}

template <typename T, typename LabelT>
void predict_impl(Handle& h, const cuml_b200_kmeans_params_t& params, const T* centroids, const T* X, int64_t n,
                  int64_t d, const T* w, bool normalize_weights, LabelT* labels, T& inertia_out)
{
  check_params(params);
  CB2_EXPECTS(n >= 0 && d >= 1, "invalid shape");
  CB2_EXPECTS(n == 0 || (X && is_device_pointer(X)), "X must be device accessible for predict");
  CB2_EXPECTS(centroids && is_device_pointer(centroids), "centroids must be device accessible");
  CB2_EXPECTS(n == 0 || labels != nullptr, "labels must not be null");
  CB2_CUDA(cudaSetDevice(h.device));
  const int k = params.n_clusters, di = static_cast<int>(d);
  if (n == 0) {
    inertia_out = T(0);
    return;
  }
  NvtxRange nvtx_predict("cuml_b200::kmeans::predict");
  SoloGuard solo(h);  // predict is rank-local (no collectives), reference dask/cluster/kmeans.py:237-243
  std::vector<Part<T>> parts{Part<T>{X, n, w}};
  LloydSolver<T> solver(h, parts, di, k, ENGINE_AUTO);
  solver.assign(centroids);
  double inertia = solver.inertia(centroids);
  if (w && normalize_weights) {
    const double ws = sum_weights<T>(h, w, n);
    CB2_EXPECTS(ws > 0.0, "sample weights must have a positive sum");
    inertia *= static_cast<double>(n) / ws;
  }
  if (std::is_same<LabelT, int32_t>::value) {
    CB2_CUDA(cudaMemcpyAsync(labels, solver.labels(), sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, h.stream));
  } else {
    labels_to_i64(h, solver.labels(), n, reinterpret_cast<int64_t*>(labels));
  }
  CB2_CUDA(cudaStreamSynchronize(h.stream));
  inertia_out = static_cast<T>(inertia);
}

template <typename T>
void transform_impl(Handle& h, const cuml_b200_kmeans_params_t& params, const T* centroids, const T* X, int64_t n,
                    int64_t d, T* X_new)
{
  check_params(params);
  CB2_EXPECTS(n >= 0 && d >= 1, "invalid shape");
  if (n == 0) return;
  CB2_EXPECTS(X && is_device_pointer(X), "X must be device accessible for transform");
  CB2_EXPECTS(centroids && is_device_pointer(centroids), "centroids must be device accessible");
  CB2_EXPECTS(X_new && is_device_pointer(X_new), "X_new must be device accessible");
  CB2_CUDA(cudaSetDevice(h.device));
  NvtxRange nvtx_transform("cuml_b200::kmeans::transform");
  const int k = params.n_clusters, di = static_cast<int>(d);
  const bool want_sqrt = params.metric == CUML_B200_L2SqrtExpanded;
  if constexpr (std::is_same<T, float>::value) {
    // tensor-core distance matrix (3xTF32, same kernels as the E-step with a distance-writing epilogue)
    static const bool tc_off = std::getenv("CUML_B200_TRANSFORM_TC") && std::atoi(std::getenv("CUML_B200_TRANSFORM_TC")) == 0;
    // the tensor-core epilogue stores 16 bytes at a time into X_new when n_clusters % 4 == 0: X_new must be aligned too
    if (!tc_off && tc_transform_supported(h, d, k) && reinterpret_cast<uintptr_t>(X) % 16 == 0 &&
        reinterpret_cast<uintptr_t>(X_new) % 16 == 0 && engine_from_env(ENGINE_AUTO) != ENGINE_SIMT) {
      TcCentroids cen;
      tc_prepare(h, centroids, k, di, cen, /*allow_bf16=*/false);
      DevBuf<float> xn(static_cast<size_t>(n), h.stream);
      row_norms<float>(h, X, n, di, xn.get());
      TcDistOut dist;
      dist.out = X_new; dist.xnorm = xn.get(); dist.sqrt = want_sqrt ? 1 : 0;
      tc_assign(h, X, n, di, k, cen, nullptr, nullptr, &dist);
      CB2_CUDA(cudaStreamSynchronize(h.stream));
      return;
    }
  }
  DevBuf<T> cn(k, h.stream);
  row_norms<T>(h, centroids, k, di, cn.get());
  simt_transform<T>(h, X, n, di, centroids, k, cn.get(), X_new, want_sqrt);
  CB2_CUDA(cudaStreamSynchronize(h.stream));
}

}  // namespace
}  // namespace cb2

using namespace cb2;

extern "C" {

void cuml_b200_kmeans_params_default(cuml_b200_kmeans_params_t* p)
{
  if (!p) return;
  p->metric                = CUML_B200_L2Expanded;
  p->n_clusters            = 8;
  p->init                  = CUML_B200_INIT_KMeansPlusPlus;
  p->max_iter              = 300;
  p->tol                   = 1e-4;
  p->verbosity             = 2;   // rapids_logger::level_enum::info
  p->rng_seed              = 0;
  p->rng_base_subsequence  = 0;
  p->rng_type              = 0;
  p->n_init                = 1;
  p->oversampling_factor   = 2.0;
  p->batch_samples         = 1 << 15;
  p->batch_centroids       = 0;
  p->init_size             = 0;
  p->device_buffer_samples = 0;
}

int cuml_b200_handle_create(cuml_b200_handle_t** out, void* stream, void* nccl_comm, int rank, int n_ranks)
{
  return guarded([&] {
    CB2_EXPECTS(out != nullptr, "null output handle pointer");
    *out = reinterpret_cast<cuml_b200_handle_t*>(make_handle(stream, nccl_comm, rank, n_ranks));
  });
}
int cuml_b200_handle_destroy(cuml_b200_handle_t* h)
{
  return guarded([&] { free_handle(reinterpret_cast<Handle*>(h)); });
}
int cuml_b200_handle_sync(cuml_b200_handle_t* h)
{
  return guarded([&] {
    CB2_EXPECTS(h != nullptr, "null handle");
    CB2_CUDA(cudaStreamSynchronize(reinterpret_cast<Handle*>(h)->stream));
  });
}
void* cuml_b200_handle_stream(cuml_b200_handle_t* h) { return h ? reinterpret_cast<Handle*>(h)->stream : nullptr; }
const char* cuml_b200_last_error(void) { return t_last_error.c_str(); }
const char* cuml_b200_version(void) { return "cuml_b200 0.1.0 (sm_100a)"; }

int cuml_b200_nccl_unique_id(void* id_out)
{
  return guarded([&] {
    CB2_EXPECTS(id_out != nullptr, "null id buffer");
    nccl::unique_id(id_out);
  });
}
int cuml_b200_handle_init_comm(cuml_b200_handle_t* h, const void* id, int rank, int n_ranks)
{
  return guarded([&] {
    CB2_EXPECTS(h && id, "null handle or id");
    CB2_EXPECTS(n_ranks >= 1 && rank >= 0 && rank < n_ranks, "invalid rank / n_ranks");
    nccl::init_rank(*reinterpret_cast<Handle*>(h), id, rank, n_ranks);
  });
}

int cuml_b200_peer_window_create(cuml_b200_handle_t* h, int n_ranks, size_t slot_bytes, void* ipc_handle_out_64_bytes)
{
  return guarded([&] {
    CB2_EXPECTS(h && ipc_handle_out_64_bytes, "null handle or output buffer");
    peer::window_create(*reinterpret_cast<Handle*>(h), slot_bytes, n_ranks, ipc_handle_out_64_bytes);
  });
}
int cuml_b200_peer_window_attach(cuml_b200_handle_t* h, const void* all_ipc_handles, int rank, int n_ranks)
{
  return guarded([&] {
    CB2_EXPECTS(h && all_ipc_handles, "null handle or handle list");
    peer::window_attach(*reinterpret_cast<Handle*>(h), all_ipc_handles, rank, n_ranks);
  });
}

#define HANDLE(h) (*reinterpret_cast<Handle*>(h))
#define REQUIRE_HANDLE(h) CB2_EXPECTS((h) != nullptr, "null handle")

#define DEFINE_FIT(SUFFIX, T, IDX)                                                                              \
  int cuml_b200_kmeans_fit_##SUFFIX(cuml_b200_handle_t* h, const cuml_b200_kmeans_params_t* params, const T* X, \
                                    IDX n_samples, IDX n_features, const T* sample_weight, T* centroids,        \
                                    T* inertia, IDX* n_iter)                                                    \
  {                                                                                                             \
    return guarded([&] {                                                                                        \
      REQUIRE_HANDLE(h);                                                                                        \
      CB2_EXPECTS(params && inertia && n_iter, "null argument");                                                \
      CB2_EXPECTS(n_samples >= 0 && (n_samples == 0 || X != nullptr), "invalid X");                             \
      const T* xp[1]     = {X};                                                                                 \
      const T* wp[1]     = {sample_weight};                                                                     \
      int64_t np[1]      = {static_cast<int64_t>(n_samples)};                                                   \
      int64_t it         = 0;                                                                                   \
      fit_parts_impl<T>(HANDLE(h), *params, xp, np, 1, static_cast<int64_t>(n_features),                        \
                        sample_weight ? wp : nullptr, centroids, *inertia, it);                                 \
      *n_iter = static_cast<IDX>(it);                                                                           \
    });                                                                                                         \
  }
DEFINE_FIT(f32_i32, float, int32_t)
DEFINE_FIT(f64_i32, double, int32_t)
DEFINE_FIT(f32_i64, float, int64_t)
DEFINE_FIT(f64_i64, double, int64_t)

#define DEFINE_FIT_PARTS(SUFFIX, T)                                                                                \
  int cuml_b200_kmeans_fit_parts_##SUFFIX(cuml_b200_handle_t* h, const cuml_b200_kmeans_params_t* params,          \
                                          const T* const* X_parts, const int64_t* n_samples_parts, int64_t n_parts, \
                                          int64_t n_features, const T* const* sample_weight_parts, T* centroids,   \
                                          T* inertia, int64_t* n_iter)                                             \
  {                                                                                                                \
    return guarded([&] {                                                                                           \
      REQUIRE_HANDLE(h);                                                                                           \
      CB2_EXPECTS(params && inertia && n_iter, "null argument");                                                   \
      CB2_EXPECTS(n_parts >= 0 && (n_parts == 0 || (X_parts && n_samples_parts)), "invalid partition list");       \
      fit_parts_impl<T>(HANDLE(h), *params, X_parts, n_samples_parts, n_parts, n_features, sample_weight_parts,    \
                        centroids, *inertia, *n_iter);                                                             \
    });                                                                                                            \
  }
DEFINE_FIT_PARTS(f32, float)
DEFINE_FIT_PARTS(f64, double)

#define DEFINE_FIT_PARTS_LABELS(SUFFIX, T)                                                                          \
  int cuml_b200_kmeans_fit_parts_labels_##SUFFIX(cuml_b200_handle_t* h, const cuml_b200_kmeans_params_t* params,    \
                                                 const T* const* X_parts, const int64_t* n_samples_parts,           \
                                                 int64_t n_parts, int64_t n_features,                               \
                                                 const T* const* sample_weight_parts, T* centroids, T* inertia,     \
                                                 int64_t* n_iter, int32_t* const* labels_parts)                     \
  {                                                                                                                 \
    return guarded([&] {                                                                                            \
      REQUIRE_HANDLE(h);                                                                                            \
      CB2_EXPECTS(params && inertia && n_iter, "null argument");                                                    \
      CB2_EXPECTS(n_parts >= 0 && (n_parts == 0 || (X_parts && n_samples_parts)), "invalid partition list");        \
      fit_parts_impl<T>(HANDLE(h), *params, X_parts, n_samples_parts, n_parts, n_features, sample_weight_parts,     \
                        centroids, *inertia, *n_iter, labels_parts);                                                \
    });                                                                                                             \
  }
DEFINE_FIT_PARTS_LABELS(f32, float)
DEFINE_FIT_PARTS_LABELS(f64, double)

#define DEFINE_PREDICT(SUFFIX, T, IDX)                                                                             \
  int cuml_b200_kmeans_predict_##SUFFIX(cuml_b200_handle_t* h, const cuml_b200_kmeans_params_t* params,            \
                                        const T* centroids, const T* X, IDX n_samples, IDX n_features,             \
                                        const T* sample_weight, int normalize_weights, IDX* labels, T* inertia)    \
  {                                                                                                                \
    return guarded([&] {                                                                                           \
      REQUIRE_HANDLE(h);                                                                                           \
      CB2_EXPECTS(params && inertia, "null argument");                                                             \
      predict_impl<T, IDX>(HANDLE(h), *params, centroids, X, static_cast<int64_t>(n_samples),                      \
                           static_cast<int64_t>(n_features), sample_weight, normalize_weights != 0, labels,        \
                           *inertia);                                                                              \
    });                                                                                                            \
  }
DEFINE_PREDICT(f32_i32, float, int32_t)
DEFINE_PREDICT(f64_i32, double, int32_t)
DEFINE_PREDICT(f32_i64, float, int64_t)
DEFINE_PREDICT(f64_i64, double, int64_t)

#define DEFINE_TRANSFORM(SUFFIX, T, IDX)                                                                       \
  int cuml_b200_kmeans_transform_##SUFFIX(cuml_b200_handle_t* h, const cuml_b200_kmeans_params_t* params,      \
                                          const T* centroids, const T* X, IDX n_samples, IDX n_features,       \
                                          T* X_new)                                                            \
  {                                                                                                            \
    return guarded([&] {                                                                                       \
      REQUIRE_HANDLE(h);                                                                                       \
      CB2_EXPECTS(params != nullptr, "null argument");                                                         \
      transform_impl<T>(HANDLE(h), *params, centroids, X, static_cast<int64_t>(n_samples),                     \
                        static_cast<int64_t>(n_features), X_new);                                              \
    });                                                                                                        \
  }
DEFINE_TRANSFORM(f32_i32, float, int32_t)
DEFINE_TRANSFORM(f64_i32, double, int32_t)
DEFINE_TRANSFORM(f32_i64, float, int64_t)
DEFINE_TRANSFORM(f64_i64, double, int64_t)

// ---- measurement / test hooks ----------------------------------------------------------------
int cuml_b200_kmeans_lloyd_step_f32(cuml_b200_handle_t* h, const float* X, int64_t n, int64_t d, const float* w,
                                    int32_t k, float* centroids, int32_t* labels, double* sums_out,
                                    double* shift2_out, int engine)
{
  return guarded([&] {
    REQUIRE_HANDLE(h);
    CB2_EXPECTS(X && centroids && n > 0 && d > 0 && k > 0, "invalid argument");
    Handle& hh = HANDLE(h);
    // the solver (operand buffers, partial tables) is cached on the handle between steps
    struct Cache {
      const float* X = nullptr;
      int64_t n = 0, d = 0;
      int k = 0, engine = -1;
      const float* w = nullptr;
      std::unique_ptr<LloydSolver<float>> solver;
    };
    auto cp = std::static_pointer_cast<Cache>(hh.step_cache);
    if (!cp) {
      cp            = std::make_shared<Cache>();
      hh.step_cache = cp;
    }
    Cache& cache = *cp;
    if (!cache.solver || cache.X != X || cache.n != n || cache.d != d || cache.k != k || cache.engine != engine ||
        cache.w != w) {
      cache.solver.reset();
      std::vector<Part<float>> parts{Part<float>{X, n, w}};
      cache.solver = std::make_unique<LloydSolver<float>>(hh, parts, static_cast<int>(d), k, engine);
      cache.X = X; cache.n = n; cache.d = d; cache.k = k; cache.engine = engine; cache.w = w;
    }
    LloydSolver<float>& s = *cache.solver;
    s.step(centroids, sums_out != nullptr);
    const size_t cnt = s.packed_count();
    if (sums_out) CB2_CUDA(cudaMemcpyAsync(sums_out, s.packed(), sizeof(double) * cnt, cudaMemcpyDeviceToDevice, hh.stream));
    if (shift2_out) CB2_CUDA(cudaMemcpyAsync(shift2_out, s.packed() + cnt, sizeof(double), cudaMemcpyDeviceToDevice, hh.stream));
    if (labels) CB2_CUDA(cudaMemcpyAsync(labels, s.labels(), sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, hh.stream));
  });
}

int cuml_b200_kmeans_assign_f32(cuml_b200_handle_t* h, const float* X, int64_t n, int64_t d, int32_t k,
                                const float* centroids, int32_t* labels, int engine)
{
  return guarded([&] {
    REQUIRE_HANDLE(h);
    CB2_EXPECTS(X && centroids && labels && n > 0 && d > 0 && k > 0, "invalid argument");
    Handle& hh = HANDLE(h);
    std::vector<Part<float>> parts;  // no owned partitions: labels are written straight to the caller
    LloydSolver<float> s(hh, parts, static_cast<int>(d), k, engine);
    if (engine == ENGINE_TC || (engine == ENGINE_AUTO && reinterpret_cast<uintptr_t>(X) % 16 != 0)) {
      CB2_EXPECTS(reinterpret_cast<uintptr_t>(X) % 16 == 0 || engine != ENGINE_TC, "X must be 16-byte aligned");
    }
    s.prepare(centroids);
    s.assign_one(centroids, X, n, labels);
  });
}

int cuml_b200_kmeans_debug_dots_f32(cuml_b200_handle_t* h, const float* X, int64_t n, int64_t d, int32_t k,
                                    const float* centroids, int32_t* labels, float* dots, int64_t* k_pad_out)
{
  return guarded([&] {
    REQUIRE_HANDLE(h);
    CB2_EXPECTS(X && centroids && labels && n > 0 && d > 0 && k > 0, "invalid argument");
    Handle& hh = HANDLE(h);
    TcCentroids cen;
    tc_prepare(hh, centroids, k, static_cast<int>(d), cen);
    if (k_pad_out) *k_pad_out = cen.k_pad;
    if (dots || labels) tc_assign(hh, X, n, static_cast<int>(d), k, cen, labels, dots);
    CB2_CUDA(cudaStreamSynchronize(hh.stream));
  });
}

void cuml_b200_launch_count_reset(void) { reset_launches(); }
int64_t cuml_b200_launch_count(void) { return launches(); }

int cuml_b200_kernel_timing_enable(cuml_b200_handle_t* h, int enable)
{
  return guarded([&] {
    REQUIRE_HANDLE(h);
    Handle& hh = HANDLE(h);
    hh.timing  = enable != 0;
    for (auto* v : {&hh.fused_events, &hh.update_events}) {
      for (auto& e : *v) hh.event_pool.push_back(e);
      v->clear();
    }
  });
}

int cuml_b200_kernel_timing_read(cuml_b200_handle_t* h, double* fused_ms, int64_t* fused_n, double* update_ms,
                                 int64_t* update_n)
{
  return guarded([&] {
    REQUIRE_HANDLE(h);
    Handle& hh = HANDLE(h);
    CB2_CUDA(cudaStreamSynchronize(hh.stream));
    auto total = [&](std::vector<EventPair>& v, double* ms, int64_t* cnt) {
      double s = 0.0;
      for (auto& e : v) {
        float t = 0.f;
        CB2_CUDA(cudaEventElapsedTime(&t, e.a, e.b));
        s += t;
      }
      if (ms) *ms = s;
      if (cnt) *cnt = static_cast<int64_t>(v.size());
      for (auto& e : v) hh.event_pool.push_back(e);
      v.clear();
    };
    total(hh.fused_events, fused_ms, fused_n);
    total(hh.update_events, update_ms, update_n);
  });
}

int cuml_b200_kmeans_tc_supported(int64_t d, int32_t k) { return tc_supported(d, k) ? 1 : 0; }

int cuml_b200_kmeans_fused_update(cuml_b200_handle_t* h, int64_t d, int32_t k)
{
  if (!h || d <= 0 || d > std::numeric_limits<int>::max() || k <= 0) return 0;
  Handle& hh = HANDLE(h);
  return (hh.cc_major == 10 && tc_supported(d, k) && tc_fused_update_supported(hh, static_cast<int>(d), k)) ? 1 : 0;
}

int cuml_b200_kmeans_estep_variant(cuml_b200_handle_t* h, int64_t d, int32_t k)
{
  if (!h || d <= 0 || d > std::numeric_limits<int>::max() || k <= 0) return 0;
  Handle& hh = HANDLE(h);
  if (hh.cc_major != 10 || !tc_supported(d, k)) return 0;
  return tc_variant(hh, static_cast<int>(d), k);
}

}  // extern "C"
