// Peer-memory collectives over NVLink / NVSwitch (or the same device): the library's own replacement for the
// small, latency-bound NCCL calls of the k-means path (one all-reduce of k*d+k+1 doubles per Lloyd iteration, plus
// the all-gathers / broadcasts of seeding).  NCCL stays available as the checked baseline (handle_init_comm).
//
// Every rank owns a *window* in device memory (cudaMalloc, exported with cudaIpcGetMemHandle and mapped by the
// other ranks):
//
//     data  [2 parities][n_ranks slots][slot_bytes]     slot r of parity p holds rank r's contribution
//     flags [2 parities][n_ranks] u64                    epoch number of the contribution in the slot
//     counter u32                                        block arrival counter of the local push kernel
//
// One exchange (epoch e, parity e & 1):
//   push     every rank copies its contribution into slot [me] of EVERY window (its own included; remote windows are
//            written with plain stores through NVLink -- posted writes, no round trip); the last block to finish
//            issues a system-scope fence and stores e into flags[me] of every window (st.release.sys).
//   consume  a kernel on the rank's own stream polls its LOCAL flags (ld.acquire.sys) until all n_ranks slots carry
//            epoch e, then reads the slots from local HBM in rank order.  Every rank therefore reduces the same
//            values in the same order: all-reduce results are bitwise identical on all ranks and independent of
//            arrival order.
// The push never waits, the consume only waits for pushes, and a rank pushes epoch e+2 into a parity only after it
// consumed e+1, i.e. after every peer pushed e+1, i.e. after every peer consumed e: two parities suffice and the
// protocol cannot deadlock as long as every rank's stream makes progress.  Two ranks that share one device (the
// single-GPU test topology; NCCL refuses it) work too: the waiting kernel is time-sliced against the other process.
//
// reference role: raft::comms::comms_t allreduce / allgather / bcast as used by cuVS's multi-GPU k-means.
#include <cstring>

#include "common.cuh"

namespace cb2 {

namespace {

constexpr int PEER_MAX_RANKS = 16;

struct PeerView {
  char* win[PEER_MAX_RANKS];   // mapped windows, win[rank] is the local one
  int rank, n_ranks;
  uint64_t slot_bytes;
};

__host__ __device__ inline size_t data_off(const PeerView& v, int parity, int slot)
{
  return (static_cast<size_t>(parity) * v.n_ranks + slot) * v.slot_bytes;
}
__host__ __device__ inline size_t flags_off(const PeerView& v) { return static_cast<size_t>(2) * v.n_ranks * v.slot_bytes; }
__host__ __device__ inline size_t counter_off(const PeerView& v) { return flags_off(v) + sizeof(uint64_t) * 2 * PEER_MAX_RANKS; }
inline size_t window_bytes(int n_ranks, size_t slot_bytes)
{
  return static_cast<size_t>(2) * n_ranks * slot_bytes + sizeof(uint64_t) * 2 * PEER_MAX_RANKS + 64;
}

__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p)
{
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// copy `bytes` (multiple of 16; src 16-byte aligned) into slot [rank] of every window, then publish the epoch
__global__ void __launch_bounds__(256) peer_push_kernel(PeerView v, const char* __restrict__ src, size_t bytes,
                                                        int parity, uint64_t epoch)
{
  const size_t n16   = bytes / 16;
  const size_t off   = data_off(v, parity, v.rank);
  const uint4* s     = reinterpret_cast<const uint4*>(src);
  const size_t step  = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += step) {
    const uint4 val = s[i];
    for (int r = 0; r < v.n_ranks; ++r) reinterpret_cast<uint4*>(v.win[r] + off)[i] = val;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned* counter = reinterpret_cast<unsigned*>(v.win[v.rank] + counter_off(v));
    const unsigned t  = atomicAdd(counter, 1u);
    if (t == gridDim.x - 1) {
      *counter = 0;   // the next push on this stream starts after this kernel
      __threadfence_system();
      for (int r = 0; r < v.n_ranks; ++r)
        st_release_sys(reinterpret_cast<uint64_t*>(v.win[r] + flags_off(v)) + parity * PEER_MAX_RANKS + v.rank, epoch);
    }
  }
}

// block-wide wait until every slot of the parity carries `epoch`
__device__ __forceinline__ void peer_wait(const PeerView& v, int parity, uint64_t epoch)
{
  if (threadIdx.x < v.n_ranks) {
    const uint64_t* f = reinterpret_cast<const uint64_t*>(v.win[v.rank] + flags_off(v)) + parity * PEER_MAX_RANKS + threadIdx.x;
    while (ld_acquire_sys(f) < epoch) __nanosleep(64);
  }
  __syncthreads();
}

// out[i] = op over ranks (rank order) of slot_r[i], doubles
template <int OP /*0 sum, 1 max*/>
__global__ void __launch_bounds__(256) peer_reduce_f64_kernel(PeerView v, double* __restrict__ out, size_t count,
                                                              int parity, uint64_t epoch)
{
  peer_wait(v, parity, epoch);
  const char* base  = v.win[v.rank];
  const size_t step = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += step) {
    double acc = __ldcg(reinterpret_cast<const double*>(base + data_off(v, parity, 0)) + i);
    for (int r = 1; r < v.n_ranks; ++r) {
      const double x = __ldcg(reinterpret_cast<const double*>(base + data_off(v, parity, r)) + i);
      acc            = OP == 0 ? acc + x : (x > acc ? x : acc);
    }
    out[i] = acc;
  }
}

// out[r * bytes_per_rank + i] = slot_r[i]  (first_rank .. first_rank + n_out ranks)
__global__ void __launch_bounds__(256) peer_gather_kernel(PeerView v, char* __restrict__ out, size_t bytes_per_rank,
                                                          int first_rank, int n_out, int parity, uint64_t epoch)
{
  peer_wait(v, parity, epoch);
  const char* base  = v.win[v.rank];
  const size_t step = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t tot  = bytes_per_rank * n_out;
  if (bytes_per_rank % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0) {
    const size_t per16 = bytes_per_rank / 16;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < tot / 16; i += step) {
      const int r = static_cast<int>(i / per16);
      reinterpret_cast<uint4*>(out)[i] =
        __ldcg(reinterpret_cast<const uint4*>(base + data_off(v, parity, first_rank + r)) + (i - r * per16));
    }
  } else {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < tot; i += step) {
      const int r = static_cast<int>(i / bytes_per_rank);
      out[i]      = __ldcg(base + data_off(v, parity, first_rank + r) + (i - r * bytes_per_rank));
    }
  }
}

// The per-iteration exchange of the Lloyd loop fused with the centroid update: wait for all ranks' packed sums,
// add them in rank order (written back to `packed`, which the next iteration's class balancing reads), divide,
// keep the old centroid of an empty cluster, squared shift.  Multi-block; the last block adds the per-block shift
// partials in block order (deterministic).
template <typename T>
__global__ void __launch_bounds__(256) peer_finalize_kernel(PeerView v, double* __restrict__ packed, T* __restrict__ C,
                                                            int k, int d, double* __restrict__ shift2_out,
                                                            double* __restrict__ block_shift, unsigned* __restrict__ done,
                                                            int parity, uint64_t epoch)
{
  __shared__ double red[8];
  __shared__ bool last;
  peer_wait(v, parity, epoch);
  const char* base   = v.win[v.rank];
  const int64_t kd   = static_cast<int64_t>(k) * d;
  const int64_t tot  = kd + k + 1;
  const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
  double acc = 0.0;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < tot; e += step) {
    double s = 0.0;
    for (int r = 0; r < v.n_ranks; ++r) s += __ldcg(reinterpret_cast<const double*>(base + data_off(v, parity, r)) + e);
    packed[e] = s;
    if (e < kd) {
      const int j = static_cast<int>(e / d);
      double wj   = 0.0;
      for (int r = 0; r < v.n_ranks; ++r) wj += __ldcg(reinterpret_cast<const double*>(base + data_off(v, parity, r)) + kd + j);
      const T old = C[e];
      T nw        = old;
      if (wj > 0.0) nw = static_cast<T>(s / wj);
      const double df = static_cast<double>(nw) - static_cast<double>(old);
      acc += df * df;
      C[e] = nw;
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < static_cast<int>(blockDim.x) / 32; ++i) s += red[i];
    block_shift[blockIdx.x] = s;
    __threadfence();
    last = atomicAdd(done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double s = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(block_shift + b);
    if (shift2_out) *shift2_out = s;
    *done = 0;
  }
}

}  // namespace

struct PeerComm {
  PeerView view{};
  void* window = nullptr;
  size_t bytes = 0;
  bool attached = false;
  uint64_t epoch = 0;
  DevBuf<char> stage;          // 16-byte padded copy of odd-sized contributions
};

namespace peer {

static PeerComm& pc(Handle& h)
{
  if (!h.peer || !h.peer->attached) throw Error(CUML_B200_INVALID_ARGUMENT, "handle has no attached peer-memory communicator");
  return *h.peer;
}

void window_create(Handle& h, size_t slot_bytes, int n_ranks, void* ipc_handle_out64)
{
  CB2_EXPECTS(n_ranks >= 1 && n_ranks <= PEER_MAX_RANKS, "peer communicator: 1 <= n_ranks <= 16");
  CB2_EXPECTS(!h.peer, "peer window already created on this handle");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (slot_bytes == 0) slot_bytes = size_t(2) << 20;
  slot_bytes = (slot_bytes + 255) & ~size_t(255);
  CB2_CUDA(cudaSetDevice(h.device));
  auto p   = std::make_unique<PeerComm>();
  p->bytes = window_bytes(n_ranks, slot_bytes);
  CB2_CUDA(cudaMalloc(&p->window, p->bytes));
  CB2_CUDA(cudaMemset(p->window, 0, p->bytes));
  CB2_CUDA(cudaDeviceSynchronize());   // zeroed before any peer can learn the handle
  p->view.slot_bytes = slot_bytes;
  p->view.n_ranks    = n_ranks;
  cudaIpcMemHandle_t ipc;
  CB2_CUDA(cudaIpcGetMemHandle(&ipc, p->window));
  std::memcpy(ipc_handle_out64, &ipc, 64);
  h.peer = p.release();
}

void window_attach(Handle& h, const void* all_handles, int rank, int n_ranks)
{
  CB2_EXPECTS(h.peer && !h.peer->attached, "peer_window_create must precede peer_window_attach (once)");
  PeerComm& p = *h.peer;
  CB2_EXPECTS(n_ranks == p.view.n_ranks && rank >= 0 && rank < n_ranks, "peer communicator: rank / n_ranks mismatch");
  CB2_CUDA(cudaSetDevice(h.device));
  p.view.rank = rank;
  for (int r = 0; r < n_ranks; ++r) {
    if (r == rank) {
      p.view.win[r] = static_cast<char*>(p.window);
      continue;
    }
    cudaIpcMemHandle_t ipc;
    std::memcpy(&ipc, static_cast<const char*>(all_handles) + static_cast<size_t>(r) * 64, 64);
    void* q = nullptr;
    CB2_CUDA(cudaIpcOpenMemHandle(&q, ipc, cudaIpcMemLazyEnablePeerAccess));
    p.view.win[r] = static_cast<char*>(q);
  }
  p.attached    = true;
  h.use_peer    = true;
  h.rank        = rank;
  h.n_ranks     = n_ranks;
}

void destroy(Handle& h)
{
  if (!h.peer) return;
  PeerComm& p = *h.peer;
  cudaStreamSynchronize(h.stream);
  if (p.attached)
    for (int r = 0; r < p.view.n_ranks; ++r)
      if (r != p.view.rank && p.view.win[r]) cudaIpcCloseMemHandle(p.view.win[r]);
  p.stage.release();
  if (p.window) cudaFree(p.window);
  delete h.peer;
  h.peer = nullptr;
}

// contribute `bytes` (<= slot_bytes) and return (parity, epoch) of the exchange
static void push(Handle& h, PeerComm& p, const void* src, size_t bytes, int& parity, uint64_t& epoch)
{
  const size_t padded = (bytes + 15) & ~size_t(15);
  const char* s       = static_cast<const char*>(src);
  if (padded != bytes || reinterpret_cast<uintptr_t>(src) % 16 != 0) {
    if (p.stage.n < padded) p.stage.alloc(std::max<size_t>(padded, 4096), h.stream);
    CB2_CUDA(cudaMemsetAsync(p.stage.get() + (padded - 16), 0, 16, h.stream));
    CB2_CUDA(cudaMemcpyAsync(p.stage.get(), src, bytes, cudaMemcpyDeviceToDevice, h.stream));
    s = p.stage.get();
  }
  epoch  = ++p.epoch;
  parity = static_cast<int>(epoch & 1);
  const unsigned blocks = static_cast<unsigned>(std::min<size_t>(32, std::max<size_t>(1, padded / 16 / 256)));
  peer_push_kernel<<<blocks, 256, 0, h.stream>>>(p.view, s, padded, parity, epoch);
  CB2_CHECK_LAUNCH();
}

template <int OP>
static void allreduce_f64(Handle& h, double* buf, size_t count)
{
  PeerComm& p      = pc(h);
  const size_t cap = p.view.slot_bytes / sizeof(double);
  for (size_t off = 0; off < count; off += cap) {
    const size_t c = std::min(cap, count - off);
    int parity;
    uint64_t epoch;
    push(h, p, buf + off, c * sizeof(double), parity, epoch);
    const unsigned blocks = static_cast<unsigned>(std::min<size_t>(32, std::max<size_t>(1, c / 256)));
    peer_reduce_f64_kernel<OP><<<blocks, 256, 0, h.stream>>>(p.view, buf + off, c, parity, epoch);
    CB2_CHECK_LAUNCH();
  }
}
void allreduce_sum_f64(Handle& h, double* buf, size_t count) { allreduce_f64<0>(h, buf, count); }
void allreduce_max_f64(Handle& h, double* buf, size_t count) { allreduce_f64<1>(h, buf, count); }

void allgather_bytes(Handle& h, const void* send, void* recv, size_t bytes_per_rank)
{
  PeerComm& p      = pc(h);
  const size_t cap = p.view.slot_bytes;
  for (size_t off = 0; off < bytes_per_rank; off += cap) {
    const size_t c = std::min(cap, bytes_per_rank - off);
    int parity;
    uint64_t epoch;
    push(h, p, static_cast<const char*>(send) + off, c, parity, epoch);
    if (off == 0 && c == bytes_per_rank) {
      const unsigned blocks = static_cast<unsigned>(std::min<size_t>(32, std::max<size_t>(1, c * p.view.n_ranks / 16 / 256)));
      peer_gather_kernel<<<blocks, 256, 0, h.stream>>>(p.view, static_cast<char*>(recv), c, 0, p.view.n_ranks, parity, epoch);
      CB2_CHECK_LAUNCH();
    } else {
      // chunked: rank r's piece lands at recv + r * bytes_per_rank + off
      for (int r = 0; r < p.view.n_ranks; ++r) {
        peer_gather_kernel<<<16, 256, 0, h.stream>>>(p.view, static_cast<char*>(recv) + r * bytes_per_rank + off, c, r, 1,
                                                     parity, epoch);
        CB2_CHECK_LAUNCH();
      }
    }
  }
}

void broadcast_bytes(Handle& h, void* buf, size_t bytes, int root)
{
  // every rank contributes (the exchange is symmetric and these messages are small); only the root's slot is read
  PeerComm& p      = pc(h);
  const size_t cap = p.view.slot_bytes;
  for (size_t off = 0; off < bytes; off += cap) {
    const size_t c = std::min(cap, bytes - off);
    int parity;
    uint64_t epoch;
    push(h, p, static_cast<char*>(buf) + off, c, parity, epoch);
    peer_gather_kernel<<<16, 256, 0, h.stream>>>(p.view, static_cast<char*>(buf) + off, c, root, 1, parity, epoch);
    CB2_CHECK_LAUNCH();
  }
}

template <typename T>
bool allreduce_finalize(Handle& h, double* packed, size_t count, T* C, int k, int d, double* shift2_out)
{
  if (!h.peer || !h.peer->attached || h.n_ranks <= 1) return false;
  PeerComm& p = *h.peer;
  if (count * sizeof(double) > p.view.slot_bytes) return false;
  int parity;
  uint64_t epoch;
  push(h, p, packed, count * sizeof(double), parity, epoch);
  const unsigned blocks = static_cast<unsigned>(std::min<size_t>(Handle::FIN_BLOCKS, std::max<size_t>(1, count / 512)));
  peer_finalize_kernel<T><<<blocks, 256, 0, h.stream>>>(p.view, packed, C, k, d, shift2_out, h.fin_scratch,
                                                        reinterpret_cast<unsigned*>(h.fin_scratch + Handle::FIN_BLOCKS),
                                                        parity, epoch);
  CB2_CHECK_LAUNCH();
  return true;
}
template bool allreduce_finalize<float>(Handle&, double*, size_t, float*, int, int, double*);
template bool allreduce_finalize<double>(Handle&, double*, size_t, double*, int, int, double*);

}  // namespace peer
}  // namespace cb2
