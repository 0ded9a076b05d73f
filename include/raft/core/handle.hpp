// Minimal stand-in for raft::handle_t: the sliver the k-means path uses -- a CUDA stream and,
// for the partition-list fit, an NCCL communicator injected by the caller
// (reference cpp/include/cuml/cluster/kmeans.hpp:11-13,86-90; wiki/cpp/DEVELOPER_GUIDE.md:391-449).
// A maintainer wiring this engine into libcuml keeps the real raft::handle_t and passes
// handle.get_stream() / the ncclComm_t of handle.get_comms() to cuml_b200_handle_create().
#pragma once
#include <cuml_b200/kmeans_c.h>

#include <stdexcept>
#include <string>

namespace raft {

class handle_t {
 public:
  // stream: cudaStream_t (nullptr = handle-owned stream); comm: ncclComm_t or nullptr
  explicit handle_t(void* stream = nullptr, void* nccl_comm = nullptr, int rank = 0, int n_ranks = 1)
  {
    if (cuml_b200_handle_create(&h_, stream, nccl_comm, rank, n_ranks) != CUML_B200_SUCCESS)
      throw std::runtime_error(std::string("raft::handle_t: ") + cuml_b200_last_error());
  }
  ~handle_t() { cuml_b200_handle_destroy(h_); }
  handle_t(const handle_t&)            = delete;
  handle_t& operator=(const handle_t&) = delete;

  void* get_stream() const { return cuml_b200_handle_stream(h_); }
  void sync_stream() const
  {
    if (cuml_b200_handle_sync(h_) != CUML_B200_SUCCESS) throw std::runtime_error(cuml_b200_last_error());
  }
  // one-process-per-GPU bring-up without raft::comms: all ranks call this with rank 0's id
  void init_nccl(const void* unique_id_128_bytes, int rank, int n_ranks)
  {
    if (cuml_b200_handle_init_comm(h_, unique_id_128_bytes, rank, n_ranks) != CUML_B200_SUCCESS)
      throw std::runtime_error(cuml_b200_last_error());
  }
  // the library's peer-memory communicator instead of NCCL: every rank creates its window (64-byte CUDA IPC handle out),
  // the caller gathers the n_ranks handles in rank order by any means, every rank attaches
  void peer_window_create(int n_ranks, void* ipc_handle_out_64_bytes)
  {
    if (cuml_b200_peer_window_create(h_, n_ranks, 0, ipc_handle_out_64_bytes) != CUML_B200_SUCCESS)
      throw std::runtime_error(cuml_b200_last_error());
  }
  void peer_window_attach(const void* all_handles, int rank, int n_ranks)
  {
    if (cuml_b200_peer_window_attach(h_, all_handles, rank, n_ranks) != CUML_B200_SUCCESS)
      throw std::runtime_error(cuml_b200_last_error());
  }
  cuml_b200_handle_t* c_handle() const { return h_; }

 private:
  cuml_b200_handle_t* h_ = nullptr;
};

}  // namespace raft
