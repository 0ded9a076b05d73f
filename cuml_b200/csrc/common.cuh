// cuml_b200 internal: error handling, handle, device buffers, launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/cuml_b200/kmeans_c.h"

namespace cb2 {

struct Error : public std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define CB2_STR2(x) #x
#define CB2_STR(x) CB2_STR2(x)
#define CB2_CUDA(call)                                                                        \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      throw ::cb2::Error(CUML_B200_CUDA_ERROR, std::string("CUDA error: ") +                  \
                                                 cudaGetErrorString(e__) + " at " __FILE__   \
                                                 ":" CB2_STR(__LINE__) " (" #call ")");       \
    }                                                                                         \
  } while (0)
#define CB2_EXPECTS(cond, msg)                                                                \
  do {                                                                                        \
    if (!(cond)) throw ::cb2::Error(CUML_B200_INVALID_ARGUMENT, std::string(msg));            \
  } while (0)
#define CB2_CHECK_LAUNCH()                                                                    \
  do {                                                                                        \
    ::cb2::count_launch();                                                                    \
    CB2_CUDA(cudaGetLastError());                                                             \
  } while (0)

void count_launch();
int64_t launches();
void reset_launches();

// cudaFuncSetAttribute applies to the current device only, and the boundary may be entered from several host
// threads (one handle per thread): one flag per device behind a mutex, not a function-local `static bool`.
class PerDeviceOnce {
  std::mutex m_;
  std::vector<char> done_;

 public:
  template <typename F>
  void run(int device, F&& f)
  {
    std::lock_guard<std::mutex> lock(m_);
    if (device >= static_cast<int>(done_.size())) done_.resize(device + 1, 0);
    if (!done_[device]) {
      f();
      done_[device] = 1;
    }
  }
};

struct EventPair {
  cudaEvent_t a, b;
};

struct PeerComm;   // peer_comm.cu

// the sliver of raft::handle_t the k-means path uses
struct Handle {
  cudaStream_t stream = nullptr;
  bool own_stream      = false;
  void* comm           = nullptr;  // ncclComm_t
  bool own_comm        = false;
  PeerComm* peer       = nullptr;  // peer-memory communicator (peer_comm.cu); used instead of NCCL once attached
  bool use_peer        = false;
  int rank             = 0;
  int n_ranks          = 1;
  int device           = 0;
  int sm_count         = 148;
  size_t smem_optin    = 0;
  int cc_major = 0, cc_minor = 0;
  // kernel timing (bench roofline): CUDA events on `stream` around the dominant kernels
  bool timing = false;
  std::vector<EventPair> fused_events, update_events;
  std::vector<EventPair> event_pool;
  // pinned scalar mailbox for per-iteration convergence read-back
  double* pinned = nullptr;
  // device scratch of the multi-block centroid update: FIN_BLOCKS shift partials + the block arrival counter
  static constexpr int FIN_BLOCKS = 64;
  double* fin_scratch = nullptr;
  // solver cached by the lloyd_step measurement hook (freed with the handle)
  std::shared_ptr<void> step_cache;
  // copy stream of the double-buffered host -> device pipelines (out-of-core fit, chunked staging), created on first use
  cudaStream_t aux_stream = nullptr;

  EventPair begin_event();
  void end_event(EventPair ev, bool fused);
};

// run a section rank-locally (no collectives) on a handle that carries a communicator
struct SoloGuard {
  Handle& h;
  int n_ranks, rank;
  void* comm;
  explicit SoloGuard(Handle& h_) : h(h_), n_ranks(h_.n_ranks), rank(h_.rank), comm(h_.comm)
  {
    h.n_ranks = 1;
    h.rank    = 0;
    h.comm    = nullptr;
  }
  ~SoloGuard()
  {
    h.n_ranks = n_ranks;
    h.rank    = rank;
    h.comm    = comm;
  }
};

template <typename T>
struct DevBuf {
  T* p           = nullptr;
  size_t n       = 0;
  cudaStream_t s = nullptr;
  bool direct    = false;   // cudaMalloc / cudaFree instead of the stream-ordered pool
  DevBuf() = default;
  DevBuf(size_t n_, cudaStream_t s_) { alloc(n_, s_); }
  void alloc(size_t n_, cudaStream_t s_)
  {
    release();
    n = n_;
    s = s_;
    // Staging-sized buffers bypass the stream-ordered pool: growing and trimming the pool by tens of GB costs
    // hundreds of milliseconds per fit (measured: tools/e2e_probe.py), a plain cudaMalloc a few.
    direct = n * sizeof(T) >= (size_t(1) << 30);
    if (n) {
      if (direct) CB2_CUDA(cudaMalloc(reinterpret_cast<void**>(&p), n * sizeof(T)));
      else CB2_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&p), n * sizeof(T), s));
    }
  }
  void release()
  {
    if (p) {
      if (direct) {
        cudaStreamSynchronize(s);   // work queued on the owning stream may still use the buffer
        cudaFree(p);
      } else {
        cudaFreeAsync(p, s);
      }
    }
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
  DevBuf(const DevBuf&)            = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept { *this = std::move(o); }
  DevBuf& operator=(DevBuf&& o) noexcept
  {
    if (this != &o) {
      release();
      p = o.p; n = o.n; s = o.s; direct = o.direct;
      o.p = nullptr; o.n = 0;
    }
    return *this;
  }
  T* get() const { return p; }
};

inline bool is_device_pointer(const void* ptr)
{
  // reference: ML::is_device_or_managed_type, cpp/src/ml_cuda_utils.h:21-33
  cudaPointerAttributes att{};
  cudaError_t e = cudaPointerGetAttributes(&att, ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return att.type == cudaMemoryTypeDevice || att.type == cudaMemoryTypeManaged;
}

// NVTX range per phase of a fit / predict / transform (the reference wraps every estimator method in an NVTX range,
// python/cuml/cuml/internals/base.py:126-128,267-285).  Header-only NVTX3: free unless a profiler is attached.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&)            = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// boolean environment switch with a default (A/B measurement aid; the defaults are the measured best)
inline bool env_flag(const char* name, bool dflt)
{
  const char* e = std::getenv(name);
  return e ? std::atoi(e) != 0 : dflt;
}

// ---- NCCL (loaded lazily with dlopen; only the MG path needs it) -------------------------
namespace nccl {
void allreduce_sum_f64(Handle& h, double* buf, size_t count);
void allreduce_max_f64(Handle& h, double* buf, size_t count);
void broadcast_bytes(Handle& h, void* buf, size_t bytes, int root);
void allgather_bytes(Handle& h, const void* send, void* recv, size_t bytes_per_rank);
void unique_id(void* out128);
void init_rank(Handle& h, const void* id128, int rank, int n_ranks);
void destroy(Handle& h);
}  // namespace nccl

// ---- peer-memory collectives (peer_comm.cu): the library's own kernels over NVLink / same-device IPC ----
namespace peer {
void window_create(Handle& h, size_t slot_bytes, int n_ranks, void* ipc_handle_out64);
void window_attach(Handle& h, const void* all_handles, int rank, int n_ranks);
void destroy(Handle& h);
void allreduce_sum_f64(Handle& h, double* buf, size_t count);
void allreduce_max_f64(Handle& h, double* buf, size_t count);
void broadcast_bytes(Handle& h, void* buf, size_t bytes, int root);
void allgather_bytes(Handle& h, const void* send, void* recv, size_t bytes_per_rank);
// all-reduce of the packed M-step sums fused with the centroid update; false = not applicable (caller falls back)
template <typename T>
bool allreduce_finalize(Handle& h, double* packed, size_t count, T* C, int k, int d, double* shift2_out);
}  // namespace peer

// ---- what the path calls: dispatch on the communicator the handle carries (raft::comms::comms_t role) ----
namespace comms {
inline void allreduce_sum_f64(Handle& h, double* buf, size_t count)
{
  if (h.n_ranks <= 1) return;
  h.use_peer ? peer::allreduce_sum_f64(h, buf, count) : nccl::allreduce_sum_f64(h, buf, count);
}
inline void allreduce_max_f64(Handle& h, double* buf, size_t count)
{
  if (h.n_ranks <= 1) return;
  h.use_peer ? peer::allreduce_max_f64(h, buf, count) : nccl::allreduce_max_f64(h, buf, count);
}
inline void broadcast_bytes(Handle& h, void* buf, size_t bytes, int root)
{
  if (h.n_ranks <= 1) return;
  h.use_peer ? peer::broadcast_bytes(h, buf, bytes, root) : nccl::broadcast_bytes(h, buf, bytes, root);
}
inline void allgather_bytes(Handle& h, const void* send, void* recv, size_t bytes_per_rank)
{
  if (h.n_ranks > 1 && h.use_peer) return peer::allgather_bytes(h, send, recv, bytes_per_rank);
  nccl::allgather_bytes(h, send, recv, bytes_per_rank);   // also the 1-rank copy
}
}  // namespace comms

}  // namespace cb2
