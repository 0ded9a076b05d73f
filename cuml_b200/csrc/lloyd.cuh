// cuml_b200 internal: the Lloyd solver over a list of device-resident row partitions.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "kernels.cuh"

namespace cb2 {

template <typename T>
struct Part {
  const T* X;  // [n, d] device
  int64_t n;
  const T* w;  // [n] device or null
};

enum Engine { ENGINE_AUTO = 0, ENGINE_SIMT = 1, ENGINE_TC = 2 };

inline int engine_from_env(int engine)
{
  if (engine != ENGINE_AUTO) return engine;
  const char* e = std::getenv("CUML_B200_ENGINE");
  if (!e) return ENGINE_AUTO;
  if (!std::strcmp(e, "simt")) return ENGINE_SIMT;
  if (!std::strcmp(e, "tc")) return ENGINE_TC;
  return ENGINE_AUTO;
}

// Cross-rank sum of the packed M-step result followed by the centroid update.  With the peer-memory communicator the
// two are one exchange: push kernel + a finalize kernel that waits for the peers' slots (peer_comm.cu); otherwise
// ncclAllReduce + finalize.
template <typename T>
inline void exchange_and_finalize(Handle& h, double* packed, size_t count, T* C, int k, int d)
{
  if (h.n_ranks > 1 && h.use_peer && peer::allreduce_finalize<T>(h, packed, count, C, k, d, packed + count)) return;
  comms::allreduce_sum_f64(h, packed, count);
  finalize_centroids<T>(h, packed, C, k, d, packed + count);
}

// One object per (dataset, k): owns labels, operand buffers and the M-step workspace.
template <typename T>
class LloydSolver {
 public:
  LloydSolver(Handle& h, std::vector<Part<T>> parts, int d, int k, int engine = ENGINE_AUTO)
    : h_(h), parts_(std::move(parts)), d_(d), k_(k)
  {
    engine      = engine_from_env(engine);
    n_local_    = 0;
    int64_t nmx = 0;
    capacity_   = parts_.empty() ? 0 : parts_[0].n;
    bool aligned = true;
    for (auto& p : parts_) {
      n_local_ += p.n;
      nmx = std::max(nmx, p.n);
      if (p.n > 0 && reinterpret_cast<uintptr_t>(p.X) % 16 != 0) aligned = false;
    }
    bool tc_ok = std::is_same<T, float>::value && tc_supported(d, k) && h.cc_major == 10 && aligned;
    if (engine == ENGINE_TC) {
      CB2_EXPECTS(tc_ok, "tcgen05 engine requested but unsupported for this problem (needs fp32, sm_100, "
                         "n_features % 4 == 0, n_features <= 1024, 16-byte aligned X)");
    }
    use_tc_ = (engine == ENGINE_SIMT) ? false : tc_ok;
    // label storage: every partition starts on a 16-byte boundary and the tail is padded so the
    // M-step's bulk copies may read one tile past the end
    int64_t off = 0;
    for (auto& p : parts_) {
      label_off_.push_back(off);
      off += (p.n + 3) & ~int64_t(3);
    }
    labels_.alloc(static_cast<size_t>(off + 512), h.stream);
    CB2_CUDA(cudaMemsetAsync(labels_.get(), 0, (off + 512) * sizeof(int32_t), h.stream));
    cnorm_.alloc(k, h.stream);
    packed_.alloc(static_cast<size_t>(k) * d + k + 2, h.stream);
    use_tma_update_ = false;
    if constexpr (std::is_same<T, float>::value) {
      use_tma_update_ = aligned && tma_update_supported(h, d, k) && engine != ENGINE_SIMT;
    }
    if (!use_tma_update_) update_plan<T>(h, std::max<int64_t>(nmx, 1), d, k, ws_);
  }

  bool uses_tensor_cores() const { return use_tc_; }
  int64_t n_local() const { return n_local_; }
  int32_t* labels(size_t part = 0) { return labels_.get() + label_off_[part]; }
  // copy the labels of all partitions, densely concatenated, to out (device)
  void export_labels(int32_t* out)
  {
    int64_t o = 0;
    for (size_t i = 0; i < parts_.size(); ++i) {
      CB2_CUDA(cudaMemcpyAsync(out + o, labels(i), sizeof(int32_t) * parts_[i].n, cudaMemcpyDeviceToDevice, h_.stream));
      o += parts_[i].n;
    }
  }
  double* packed() { return packed_.get(); }  // S | W | inertia | shift2
  size_t packed_count() const { return static_cast<size_t>(k_) * d_ + k_ + 1; }

  // E-step for every partition -> labels
  void assign(const T* C)
  {
    prepare(C);
    for (size_t i = 0; i < parts_.size(); ++i) assign_one(C, parts_[i].X, parts_[i].n, labels(i));
  }

  // E-step on arbitrary rows with the operand buffers of the last prepare()/assign()
  void prepare(const T* C)
  {
    if (use_tc_) {
      if constexpr (std::is_same<T, float>::value) tc_prepare(h_, C, k_, d_, tc_);
    } else {
      row_norms<T>(h_, C, k_, d_, cnorm_.get());
    }
  }
  void assign_one(const T* C, const T* X, int64_t n, int32_t* labels)
  {
    if (use_tc_) {
      if constexpr (std::is_same<T, float>::value) tc_assign(h_, X, n, d_, k_, tc_, labels);
    } else {
      simt_assign<T>(h_, X, n, d_, C, k_, cnorm_.get(), labels, nullptr);
    }
  }

  // M-step accumulation over all partitions using the current labels: packed = S | W | inertia.
  // The inertia cell (wrt C, exact difference form) is only filled when with_inertia is set; the
  // Lloyd loop itself does not need it (the stopping rule is the centroid shift).
  // `into`: add to what packed already holds (out-of-core batches of one iteration)
  void accumulate(const T* C, bool with_inertia, bool into = false)
  {
    EventPair ev{};
    if (h_.timing) ev = h_.begin_event();
    double* inertia_cell = packed_.get() + packed_count() - 1;
    if (parts_.empty() && !into) CB2_CUDA(cudaMemsetAsync(packed_.get(), 0, packed_count() * sizeof(double), h_.stream));
    bool need_separate_inertia = with_inertia;
    if (use_tma_update_) {
      if constexpr (std::is_same<T, float>::value) {
        // consumer classes balanced by the cluster weights the previous accumulate left in packed
        const uint8_t* cls_map = nullptr;
        if (have_weights_ && !parts_.empty())
          cls_map = tma_update_balance(h_, packed_.get() + static_cast<size_t>(k_) * d_, d_, k_, cls_map_);
        for (size_t i = 0; i < parts_.size(); ++i)
          tma_update_accumulate(h_, parts_[i].X, parts_[i].n, d_, labels(i), parts_[i].w, k_, tma_S_, tma_W_,
                                packed_.get(), i != 0 || into, cls_map);
        have_weights_ = true;
      }
      if (!with_inertia && !into) CB2_CUDA(cudaMemsetAsync(inertia_cell, 0, sizeof(double), h_.stream));
    } else {
      for (size_t i = 0; i < parts_.size(); ++i)
        update_accumulate<T>(h_, ws_, parts_[i].X, parts_[i].n, d_, labels(i), parts_[i].w, C, k_, packed_.get(),
                             i != 0 || into, true);
      need_separate_inertia = false;  // the generic kernel produces it in the same pass
    }
    if (need_separate_inertia) inertia_only(C, into);
    if (h_.timing) h_.end_event(ev, false);
  }

  void inertia_only(const T* C, bool into = false)
  {
    double* inertia_cell = packed_.get() + packed_count() - 1;
    if (parts_.empty() && !into) CB2_CUDA(cudaMemsetAsync(inertia_cell, 0, sizeof(double), h_.stream));
    for (size_t i = 0; i < parts_.size(); ++i)
      compute_inertia<T>(h_, parts_[i].X, parts_[i].n, d_, labels(i), parts_[i].w, C, inertia_cell, i != 0 || into);
  }

  // out-of-core use: the single partition is a device staging buffer whose fill level changes per batch
  // and which alternates between two buffers (n must not exceed the row count the solver was built with)
  void set_part(const T* X, const T* w, int64_t n)
  {
    CB2_EXPECTS(parts_.size() == 1 && n >= 0 && n <= capacity_, "set_part: single-partition solver, n within capacity");
    parts_[0] = Part<T>{X, n, w};
    n_local_  = n;
  }

  // E-step and M-step in one pass over X where the fused kernel applies (fp32, n_features = 16, k <= 64, no weights,
  // at least two rows per partition): labels + packed sums / weights.  false = not applicable, nothing was launched.
  bool assign_accumulate_fused(const T* C)
  {
    if constexpr (!std::is_same<T, float>::value) {
      return false;
    } else {
      if (!use_tc_ || parts_.empty() || !tc_fused_update_supported(h_, d_, k_)) return false;
      for (auto& p : parts_)
        if (p.w != nullptr || p.n < 2) return false;
      prepare(C);
      for (size_t i = 0; i < parts_.size(); ++i) {
        TcMstepOut ms;
        ms.partial_S = &tma_S_;
        ms.partial_W = &tma_W_;
        tc_assign(h_, parts_[i].X, parts_[i].n, d_, k_, tc_, labels(i), nullptr, nullptr, nullptr, &ms);
        CB2_EXPECTS(ms.row_blocks > 0, "fused E+M kernel was planned but not launched");
        EventPair ev{};
        if (h_.timing) ev = h_.begin_event();
        tma_update_reduce(h_, tma_S_.get(), tma_W_.get(), ms.row_blocks, k_, d_, packed_.get(), i != 0);
        if (h_.timing) h_.end_event(ev, false);
      }
      CB2_CUDA(cudaMemsetAsync(packed_.get() + packed_count() - 1, 0, sizeof(double), h_.stream));
      have_weights_ = true;
      return true;
    }
  }

  // One full Lloyd iteration, centroids updated in place; squared shift left at packed[count]
  void step(T* C, bool with_inertia = false)
  {
    if (!with_inertia && assign_accumulate_fused(C)) {
      exchange_and_finalize<T>(h_, packed_.get(), packed_count(), C, k_, d_);
      return;
    }
    assign(C);
    accumulate(C, with_inertia);
    exchange_and_finalize<T>(h_, packed_.get(), packed_count(), C, k_, d_);
  }

  // inertia of the current labelling wrt C (exact difference form), all ranks; host result
  double inertia(const T* C)
  {
    inertia_only(C);
    double* cell = packed_.get() + packed_count() - 1;
    comms::allreduce_sum_f64(h_, cell, 1);
    CB2_CUDA(cudaMemcpyAsync(h_.pinned, cell, sizeof(double), cudaMemcpyDeviceToHost, h_.stream));
    CB2_CUDA(cudaStreamSynchronize(h_.stream));
    return h_.pinned[0];
  }

  // Lloyd iterations from the centroids in C (in place).  Returns executed iterations.
  // Stopping rule of the reference's GPU path: stop after the iteration whose raw squared
  // centroid shift is < tol (tol <= 0 never stops early and needs no host read-back at all).
  int64_t run(T* C, int max_iter, double tol)
  {
    int64_t it = 0;
    for (; it < max_iter;) {
      step(C);
      ++it;
      if (tol > 0.0) {
        CB2_CUDA(cudaMemcpyAsync(h_.pinned, packed_.get() + packed_count(), sizeof(double), cudaMemcpyDeviceToHost,
                                 h_.stream));
        CB2_CUDA(cudaStreamSynchronize(h_.stream));
        if (h_.pinned[0] < tol) break;
      }
    }
    return it;
  }

 private:
  Handle& h_;
  std::vector<Part<T>> parts_;
  int d_, k_;
  int64_t n_local_ = 0;
  int64_t capacity_ = 0;
  bool use_tc_     = false;
  bool use_tma_update_ = false;
  std::vector<int64_t> label_off_;
  DevBuf<float> tma_S_, tma_W_;
  DevBuf<uint8_t> cls_map_;
  bool have_weights_ = false;
  DevBuf<int32_t> labels_;
  DevBuf<T> cnorm_;
  DevBuf<double> packed_;
  TcCentroids tc_;
  UpdateWorkspace<T> ws_;
};

// ---- seeding (seeding.cu) -------------------------------------------------------------------
template <typename T>
struct SeedContext {
  Handle& h;
  std::vector<Part<T>> parts;
  int d;
  int64_t n_local;
  int64_t n_global;
  int64_t row_offset;  // global index of this rank's first row
  uint64_t seed;
  int engine;
};

template <typename T>
void init_random(SeedContext<T>& ctx, int k, T* C);
template <typename T>
void init_kmeans_plus_plus(SeedContext<T>& ctx, int k, T* C);  // sequential D^2 sampling (single rank)
template <typename T>
void init_scalable(SeedContext<T>& ctx, const cuml_b200_kmeans_params_t& params, T* C);  // k-means||

}  // namespace cb2
