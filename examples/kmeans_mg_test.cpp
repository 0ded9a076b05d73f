// C++ consumer of ML::kmeans::{fit,predict} on a handle that carries an NCCL communicator -- the role of the
// reference's multi-GPU gtest cpp/tests/mg/kmeans_test.cu:50-195: blobs in [-10, 10]^d (seed 1234), n_init = 5,
// rng seed 1, oversampling_factor = 1, tol 1e-4, weighted (all-ones weights) and unweighted, float and double, the
// eight (n_row, n_col, n_clusters) inputs of :167-183, pass criterion adjusted Rand index >= 0.99 against the
// generating labels (:144-156, :186-189).  Like the reference's test the default run injects a ONE-rank communicator
// (ncclCommInitAll(&comm, 1, {0}) there, cuml_b200_handle_init_comm(id, 0, 1) here); `kmeans_mg_test N` (N >= 2)
// additionally runs N rank processes, one GPU each, on contiguous row shards of the same inputs.
// `kmeans_mg_test N peer` does the same over the library's peer-memory communicator (CUDA IPC windows instead of NCCL;
// the IPC handles travel through files like the unique id), `kmeans_mg_test N shared` puts all N ranks on device 0
// (peer-memory communicator; NCCL refuses two ranks on one device) -- the multi-rank run a one-GPU box can do.
//   g++ -O2 -std=c++17 -Iinclude -I/usr/local/cuda/include examples/kmeans_mg_test.cpp -Lcuml_b200/lib -lcuml_b200
//       -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/cuml_b200/lib -o kmeans_mg_test        (one command line)
#include <cuda_runtime.h>
#include <cuml/cluster/kmeans.hpp>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#define CHECK(call)                                                                                  \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess) {                                                                        \
      std::fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); \
      std::exit(2);                                                                                  \
    }                                                                                                \
  } while (0)

struct Inputs {
  int n_row, n_col, n_clusters;
  double tol;
  bool weighted;
};
static const Inputs kInputs[] = {{1000, 32, 5, 0.0001, true},    {1000, 32, 5, 0.0001, false},
                                 {1000, 100, 20, 0.0001, true},  {1000, 100, 20, 0.0001, false},
                                 {10000, 32, 10, 0.0001, true},  {10000, 32, 10, 0.0001, false},
                                 {10000, 100, 50, 0.0001, true}, {10000, 100, 50, 0.0001, false}};

// adjusted Rand index of two labelings with values in [0, k)
static double adjusted_rand_index(const std::vector<int>& a, const std::vector<int>& b, int k)
{
  std::vector<double> table(static_cast<size_t>(k) * k, 0.0), ra(k, 0.0), rb(k, 0.0);
  for (size_t i = 0; i < a.size(); ++i) {
    table[static_cast<size_t>(a[i]) * k + b[i]] += 1.0;
    ra[a[i]] += 1.0;
    rb[b[i]] += 1.0;
  }
  auto c2 = [](double x) { return x * (x - 1.0) * 0.5; };
  double sum_ij = 0.0, sum_a = 0.0, sum_b = 0.0;
  for (double v : table) sum_ij += c2(v);
  for (double v : ra) sum_a += c2(v);
  for (double v : rb) sum_b += c2(v);
  const double total = c2(static_cast<double>(a.size()));
  const double expected = sum_a * sum_b / total, max_index = 0.5 * (sum_a + sum_b);
  return max_index == expected ? 1.0 : (sum_ij - expected) / (max_index - expected);
}

// one input on this rank's contiguous row shard; returns the shard's ARI
template <typename T>
static double run_case(const raft::handle_t& handle, const Inputs& in, int rank, int n_ranks)
{
  const int n = in.n_row, d = in.n_col, k = in.n_clusters;
  // the same blobs on every rank (seed 1234), of which the rank keeps rows [lo, hi)
  std::mt19937_64 gen(1234ull);
  std::uniform_real_distribution<double> box(-10.0, 10.0);
  std::normal_distribution<double> noise(0.0, 1.0);
  std::vector<double> centres(static_cast<size_t>(k) * d);
  for (auto& c : centres) c = box(gen);
  std::vector<T> X(static_cast<size_t>(n) * d);
  std::vector<int> truth(n);
  for (int i = 0; i < n; ++i) {
    truth[i] = static_cast<int>(gen() % static_cast<unsigned long long>(k));
    for (int j = 0; j < d; ++j)
      X[static_cast<size_t>(i) * d + j] = static_cast<T>(centres[static_cast<size_t>(truth[i]) * d + j] + noise(gen));
  }
  const int base = n / n_ranks, rem = n % n_ranks;
  const int lo = rank * base + (rank < rem ? rank : rem), rows = base + (rank < rem ? 1 : 0);

  T *d_X = nullptr, *d_C = nullptr, *d_w = nullptr;
  int* d_labels = nullptr;
  CHECK(cudaMalloc(&d_X, sizeof(T) * static_cast<size_t>(rows) * d));
  CHECK(cudaMalloc(&d_C, sizeof(T) * static_cast<size_t>(k) * d));
  CHECK(cudaMalloc(&d_labels, sizeof(int) * rows));
  CHECK(cudaMemcpy(d_X, X.data() + static_cast<size_t>(lo) * d, sizeof(T) * static_cast<size_t>(rows) * d,
                   cudaMemcpyHostToDevice));
  if (in.weighted) {
    std::vector<T> ones(rows, T(1));
    CHECK(cudaMalloc(&d_w, sizeof(T) * rows));
    CHECK(cudaMemcpy(d_w, ones.data(), sizeof(T) * rows, cudaMemcpyHostToDevice));
  }

  ML::kmeans::KMeansParams params;
  params.n_clusters          = k;
  params.tol                 = in.tol;
  params.n_init              = 5;
  params.rng_state.seed      = 1;
  params.oversampling_factor = 1;
  T inertia  = 0;
  int n_iter = 0;
  ML::kmeans::fit(handle, params, d_X, rows, d, d_w, d_C, inertia, n_iter);
  ML::kmeans::predict(handle, params, d_C, d_X, rows, d, d_w, true, d_labels, inertia);
  handle.sync_stream();

  std::vector<int> labels(rows);
  CHECK(cudaMemcpy(labels.data(), d_labels, sizeof(int) * rows, cudaMemcpyDeviceToHost));
  const double score = adjusted_rand_index(std::vector<int>(truth.begin() + lo, truth.begin() + lo + rows), labels, k);
  cudaFree(d_X);
  cudaFree(d_C);
  cudaFree(d_labels);
  cudaFree(d_w);
  return score;
}

// publish `bytes` atomically at `path`: write a temporary, then rename
static bool put_file(const std::string& path, const void* data, size_t bytes)
{
  const std::string tmp = path + ".tmp";
  FILE* f = std::fopen(tmp.c_str(), "wb");
  if (!f || std::fwrite(data, 1, bytes, f) != bytes) return false;
  std::fclose(f);
  return std::rename(tmp.c_str(), path.c_str()) == 0;
}
static bool get_file(const std::string& path, void* data, size_t bytes)
{
  FILE* f = nullptr;
  for (int tries = 0; tries < 600 && !(f = std::fopen(path.c_str(), "rb")); ++tries)
    std::this_thread::sleep_for(std::chrono::milliseconds(100));
  if (!f || std::fread(data, 1, bytes, f) != bytes) return false;
  std::fclose(f);
  return true;
}

// all inputs, float and double, on rank `rank` of `n_ranks`; the unique id comes from `id_path` (written by rank 0).
// mode: "nccl" (one device per rank), "peer" (one device per rank, peer-memory communicator), "shared" (all on device 0)
static int run_rank(int rank, int n_ranks, const std::string& id_path, const std::string& mode = "nccl")
{
  CHECK(cudaSetDevice(mode == "shared" ? 0 : rank));
  if (mode != "nccl" && n_ranks > 1) {
    int failed = 0;
    try {
      raft::handle_t handle;
      unsigned char mine[64];
      std::vector<unsigned char> all(static_cast<size_t>(64) * n_ranks);
      handle.peer_window_create(n_ranks, mine);
      if (!put_file(id_path + ".ipc" + std::to_string(rank), mine, 64)) return 2;
      for (int r = 0; r < n_ranks; ++r)
        if (!get_file(id_path + ".ipc" + std::to_string(r), all.data() + static_cast<size_t>(64) * r, 64)) return 2;
      handle.peer_window_attach(all.data(), rank, n_ranks);
      // every rank has mapped every window before anyone starts exchanging
      if (!put_file(id_path + ".att" + std::to_string(rank), mine, 1)) return 2;
      unsigned char one;
      for (int r = 0; r < n_ranks; ++r)
        if (!get_file(id_path + ".att" + std::to_string(r), &one, 1)) return 2;
      for (const Inputs& in : kInputs) {
        const double sf = run_case<float>(handle, in, rank, n_ranks);
        const double sd = run_case<double>(handle, in, rank, n_ranks);
        std::printf("[rank %d/%d %s] %5d x %3d k %2d %-10s ARI float %.4f double %.4f\n", rank, n_ranks, mode.c_str(),
                    in.n_row, in.n_col, in.n_clusters, in.weighted ? "weighted" : "unweighted", sf, sd);
        if (!(sf >= 0.99) || !(sd >= 0.99)) ++failed;
      }
    } catch (const std::exception& e) {
      std::fprintf(stderr, "[rank %d] exception: %s\n", rank, e.what());
      return 2;
    }
    return failed ? 1 : 0;
  }
  unsigned char id[128];
  if (rank == 0) {
    if (cuml_b200_nccl_unique_id(id) != CUML_B200_SUCCESS) {
      std::fprintf(stderr, "nccl unique id: %s\n", cuml_b200_last_error());
      return 2;
    }
    if (!id_path.empty()) {   // publish atomically: write a temporary, then rename
      const std::string tmp = id_path + ".tmp";
      FILE* f = std::fopen(tmp.c_str(), "wb");
      if (!f || std::fwrite(id, 1, sizeof(id), f) != sizeof(id)) return 2;
      std::fclose(f);
      if (std::rename(tmp.c_str(), id_path.c_str()) != 0) return 2;
    }
  } else {
    FILE* f = nullptr;
    for (int tries = 0; tries < 600 && !(f = std::fopen(id_path.c_str(), "rb")); ++tries)
      std::this_thread::sleep_for(std::chrono::milliseconds(100));
    if (!f || std::fread(id, 1, sizeof(id), f) != sizeof(id)) return 2;
    std::fclose(f);
  }
  int failed = 0;
  try {
    raft::handle_t handle;                 // handle-owned stream
    handle.init_nccl(id, rank, n_ranks);   // the injected communicator (build_comms_nccl_only in the reference test)
    for (const Inputs& in : kInputs) {
      const double sf = run_case<float>(handle, in, rank, n_ranks);
      const double sd = run_case<double>(handle, in, rank, n_ranks);
      std::printf("[rank %d/%d] %5d x %3d k %2d %-10s ARI float %.4f double %.4f\n", rank, n_ranks, in.n_row, in.n_col,
                  in.n_clusters, in.weighted ? "weighted" : "unweighted", sf, sd);
      if (!(sf >= 0.99) || !(sd >= 0.99)) ++failed;
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "[rank %d] exception: %s\n", rank, e.what());
    return 2;
  }
  return failed ? 1 : 0;
}

int main(int argc, char** argv)
{
  if (argc >= 5 && std::strcmp(argv[1], "--rank") == 0)   // a rank process started by the launcher below
    return run_rank(std::atoi(argv[2]), std::atoi(argv[3]), argv[4], argc >= 6 ? argv[5] : "nccl");
  const int n_ranks      = argc > 1 ? std::atoi(argv[1]) : 1;
  const std::string mode = argc > 2 ? argv[2] : "nccl";
  if (n_ranks <= 1) {
    const int rc = run_rank(0, 1, "");
    std::printf(rc == 0 ? "PASSED\n" : "FAILED\n");
    return rc;
  }
  // one process per GPU (no CUDA call before the fork: the launcher itself never touches the device)
  const std::string id_path = "/tmp/cuml_b200_mg_test_id." + std::to_string(static_cast<long>(getpid()));
  std::vector<pid_t> pids;
  for (int r = 0; r < n_ranks; ++r) {
    const pid_t pid = fork();
    if (pid == 0) {
      const std::string rs = std::to_string(r), ns = std::to_string(n_ranks);
      execl(argv[0], argv[0], "--rank", rs.c_str(), ns.c_str(), id_path.c_str(), mode.c_str(), static_cast<char*>(nullptr));
      std::perror("execl");
      _exit(127);
    }
    pids.push_back(pid);
  }
  int bad = 0;
  for (pid_t pid : pids) {
    int status = 0;
    waitpid(pid, &status, 0);
    if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) ++bad;
  }
  std::remove(id_path.c_str());
  for (int r = 0; r < n_ranks; ++r) {
    std::remove((id_path + ".ipc" + std::to_string(r)).c_str());
    std::remove((id_path + ".att" + std::to_string(r)).c_str());
  }
  std::printf(bad == 0 ? "PASSED\n" : "FAILED\n");
  return bad == 0 ? 0 : 1;
}
