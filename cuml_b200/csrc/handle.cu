// Handle, error mailbox, launch accounting and the lazily-loaded NCCL shim.
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "common.cuh"

namespace cb2 {

static thread_local int64_t t_launches = 0;
void count_launch() { ++t_launches; }
int64_t launches() { return t_launches; }
void reset_launches() { t_launches = 0; }

EventPair Handle::begin_event()
{
  EventPair ev;
  if (!event_pool.empty()) {
    ev = event_pool.back();
    event_pool.pop_back();
  } else {
    CB2_CUDA(cudaEventCreate(&ev.a));
    CB2_CUDA(cudaEventCreate(&ev.b));
  }
  CB2_CUDA(cudaEventRecord(ev.a, stream));
  return ev;
}
void Handle::end_event(EventPair ev, bool fused)
{
  CB2_CUDA(cudaEventRecord(ev.b, stream));
  (fused ? fused_events : update_events).push_back(ev);
}

// ------------------------------------------------------------------------------------------
// NCCL through dlopen: the single-GPU library has no link-time NCCL dependency; the MG path
// resolves libnccl.so.2 (torch's bundled copy if already loaded in-process, else the system one).
namespace nccl {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclChar = 0, ncclUint8 = 1, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 };

struct Api {
  int (*GetUniqueId)(ncclUniqueId*);
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  int (*CommDestroy)(ncclComm_t);
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(int);
  bool ok = false;
};

static Api& api()
{
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* lib           = nullptr;
    for (auto n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) return;
    auto sym = [&](const char* s) { return dlsym(lib, s); };
    a.GetUniqueId    = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank   = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy    = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.AllReduce      = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
    a.Broadcast      = reinterpret_cast<decltype(a.Broadcast)>(sym("ncclBroadcast"));
    a.AllGather      = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.Broadcast &&
           a.AllGather && a.GetErrorString;
  });
  if (!a.ok) throw Error(CUML_B200_NCCL_ERROR, "libnccl.so.2 could not be loaded (multi-GPU k-means needs NCCL)");
  return a;
}

#define CB2_NCCL(call)                                                                        \
  do {                                                                                        \
    int r__ = (call);                                                                         \
    if (r__ != ncclSuccess)                                                                   \
      throw Error(CUML_B200_NCCL_ERROR, std::string("NCCL error: ") + api().GetErrorString(r__) + \
                                          " (" #call ")");                                    \
  } while (0)

static ncclComm_t comm_of(Handle& h)
{
  if (!h.comm) throw Error(CUML_B200_INVALID_ARGUMENT, "handle has no NCCL communicator (n_ranks > 1 requires one)");
  return reinterpret_cast<ncclComm_t>(h.comm);
}

void allreduce_sum_f64(Handle& h, double* buf, size_t count)
{
  if (h.n_ranks <= 1) return;
  CB2_NCCL(api().AllReduce(buf, buf, count, ncclFloat64, ncclSum, comm_of(h), h.stream));
}
void allreduce_max_f64(Handle& h, double* buf, size_t count)
{
  if (h.n_ranks <= 1) return;
  CB2_NCCL(api().AllReduce(buf, buf, count, ncclFloat64, ncclMax, comm_of(h), h.stream));
}
void broadcast_bytes(Handle& h, void* buf, size_t bytes, int root)
{
  if (h.n_ranks <= 1) return;
  CB2_NCCL(api().Broadcast(buf, buf, bytes, ncclUint8, root, comm_of(h), h.stream));
}
void allgather_bytes(Handle& h, const void* send, void* recv, size_t bytes_per_rank)
{
  if (h.n_ranks <= 1) {
    if (send != recv) CB2_CUDA(cudaMemcpyAsync(recv, send, bytes_per_rank, cudaMemcpyDeviceToDevice, h.stream));
    return;
  }
  CB2_NCCL(api().AllGather(send, recv, bytes_per_rank, ncclUint8, comm_of(h), h.stream));
}
void unique_id(void* out128)
{
  ncclUniqueId id;
  CB2_NCCL(api().GetUniqueId(&id));
  std::memcpy(out128, &id, 128);
}
void init_rank(Handle& h, const void* id128, int rank, int n_ranks)
{
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  ncclComm_t c = nullptr;
  CB2_CUDA(cudaSetDevice(h.device));
  CB2_NCCL(api().CommInitRank(&c, n_ranks, id, rank));
  h.comm     = c;
  h.own_comm = true;
  h.rank     = rank;
  h.n_ranks  = n_ranks;
}
void destroy(Handle& h)
{
  if (h.comm && h.own_comm) api().CommDestroy(reinterpret_cast<ncclComm_t>(h.comm));
  h.comm = nullptr;
}
}  // namespace nccl

Handle* make_handle(void* stream, void* comm, int rank, int n_ranks)
{
  auto h = std::make_unique<Handle>();
  CB2_CUDA(cudaGetDevice(&h->device));
  cudaDeviceProp prop{};
  CB2_CUDA(cudaGetDeviceProperties(&prop, h->device));
  h->sm_count   = prop.multiProcessorCount;
  h->smem_optin = prop.sharedMemPerBlockOptin;
  h->cc_major   = prop.major;
  h->cc_minor   = prop.minor;
  if (stream) {
    h->stream = reinterpret_cast<cudaStream_t>(stream);
  } else {
    CB2_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
  }
  h->comm    = comm;
  h->rank    = rank;
  h->n_ranks = n_ranks < 1 ? 1 : n_ranks;
  {
    // keep up to 8 GB of freed work buffers (labels, partial tables, the per-row buffers of k-means|| seeding: ~4 GB at
    // C5) cached in the stream-ordered pool between calls instead of returning them to the driver at every
    // synchronisation (with 2 GB the C5 seeding time swung between 0.24 and 0.83 s from run to run)
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, h->device) == cudaSuccess && pool) {
      uint64_t thr = 0;
      if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr) == cudaSuccess && thr < (uint64_t(1) << 33)) {
        thr = uint64_t(1) << 33;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
      }
    }
    cudaGetLastError();
  }
  CB2_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h->pinned), 64 * sizeof(double)));
  CB2_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->fin_scratch), (Handle::FIN_BLOCKS + 2) * sizeof(double)));
  CB2_CUDA(cudaMemset(h->fin_scratch, 0, (Handle::FIN_BLOCKS + 2) * sizeof(double)));
  return h.release();
}

void free_handle(Handle* h)
{
  if (!h) return;
  cudaStreamSynchronize(h->stream);
  h->step_cache.reset();
  if (h->aux_stream) {
    cudaStreamSynchronize(h->aux_stream);
    cudaStreamDestroy(h->aux_stream);
  }
  nccl::destroy(*h);
  peer::destroy(*h);
  for (auto& v : {&h->fused_events, &h->update_events, &h->event_pool})
    for (auto& e : *v) {
      cudaEventDestroy(e.a);
      cudaEventDestroy(e.b);
    }
  if (h->pinned) cudaFreeHost(h->pinned);
  if (h->fin_scratch) cudaFree(h->fin_scratch);
  if (h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
}

}  // namespace cb2
