// C++ consumer of ML::kmeans::{fit,predict} on the shapes of the reference's own C++ benchmark -- the role of
// cpp/bench/sg/kmeans.cu:43-61 (fit + predict timed together) with its inputs (:82-121): blobs in [-10, 10]^d,
// (n, d) in {160k, 320k, 640k} x 64, 80k x 500, 160k x 2000; k in {8, 16, 32}; init = k-means||, max_iter 300,
// tol 1e-4; timed with CUDA events after an L2 flush (cpp/bench/common/ml_benchmark.hpp:32-81).
//   g++ -O2 -std=c++17 -Iinclude examples/kmeans_bench.cpp -Lcuml_b200/lib -lcuml_b200 \
//       -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/cuml_b200/lib -o kmeans_bench
//   ./kmeans_bench [max_rows]        (max_rows caps n, for a quick run)
#include <cuda_runtime.h>
#include <cuml/cluster/kmeans.hpp>

#include <cstdio>
#include <cstdlib>
#include <random>
#include <utility>
#include <vector>

#define CHECK(call)                                                                               \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess) {                                                                     \
      std::fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); \
      return 1;                                                                                   \
    }                                                                                             \
  } while (0)

int main(int argc, char** argv)
{
  const long max_rows = argc > 1 ? std::atol(argv[1]) : 0;
  const std::vector<std::pair<int, int>> rowcols = {{160000, 64}, {320000, 64}, {640000, 64}, {80000, 500}, {160000, 2000}};
  const std::vector<int> nclasses                = {8, 16, 32};

  // a buffer larger than L2 (126 MB), rewritten before every timed call
  const size_t flush_bytes = size_t(256) << 20;
  void* flush = nullptr;
  CHECK(cudaMalloc(&flush, flush_bytes));

  raft::handle_t handle;
  cudaStream_t stream = static_cast<cudaStream_t>(handle.get_stream());
  cudaEvent_t e0, e1;
  CHECK(cudaEventCreate(&e0));
  CHECK(cudaEventCreate(&e1));

  std::printf("%9s %6s %4s %10s %8s %14s\n", "rows", "cols", "k", "fit+pred ms", "n_iter", "inertia");
  for (auto rc : rowcols) {
    int n       = rc.first;
    const int d = rc.second;
    if (max_rows > 0 && n > max_rows) n = static_cast<int>(max_rows);
    for (int k : nclasses) {
      // blobs: k centres ~ U(-10, 10)^d, unit-variance isotropic noise, seed 12345 (kmeans.cu:87-91)
      std::mt19937_64 gen(12345ull);
      std::uniform_real_distribution<float> box(-10.f, 10.f);
      std::normal_distribution<float> noise(0.f, 1.f);
      std::vector<float> centres(static_cast<size_t>(k) * d);
      for (auto& c : centres) c = box(gen);
      std::vector<float> h_X(static_cast<size_t>(n) * d);
      for (int i = 0; i < n; ++i) {
        const int c = static_cast<int>(gen() % static_cast<unsigned long long>(k));
        for (int j = 0; j < d; ++j) h_X[static_cast<size_t>(i) * d + j] = centres[static_cast<size_t>(c) * d + j] + noise(gen);
      }
      float *d_X = nullptr, *d_C = nullptr;
      int* d_labels = nullptr;
      CHECK(cudaMalloc(&d_X, sizeof(float) * h_X.size()));
      CHECK(cudaMalloc(&d_C, sizeof(float) * k * d));
      CHECK(cudaMalloc(&d_labels, sizeof(int) * n));
      CHECK(cudaMemcpy(d_X, h_X.data(), sizeof(float) * h_X.size(), cudaMemcpyHostToDevice));

      ML::kmeans::KMeansParams params;
      params.n_clusters     = k;
      params.init           = ML::kmeans::KMeansParams::InitMethod::KMeansPlusPlus;
      params.max_iter       = 300;
      params.tol            = 1e-4;
      params.metric         = ML::distance::DistanceType::L2Expanded;
      params.rng_state.seed = 12345ull;
      float inertia = 0.f, best_ms = 0.f;
      int n_iter = 0;
      for (int rep = 0; rep < 3; ++rep) {        // first repetition is the warm-up
        CHECK(cudaMemsetAsync(flush, rep, flush_bytes, stream));
        CHECK(cudaEventRecord(e0, stream));
        ML::kmeans::fit(handle, params, d_X, n, d, nullptr, d_C, inertia, n_iter);
        ML::kmeans::predict(handle, params, d_C, d_X, n, d, nullptr, true, d_labels, inertia);
        CHECK(cudaEventRecord(e1, stream));
        CHECK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep == 1 || (rep > 1 && ms < best_ms)) best_ms = ms;
      }
      std::printf("%9d %6d %4d %10.3f %8d %14.6g\n", n, d, k, best_ms, n_iter, static_cast<double>(inertia));
      cudaFree(d_X);
      cudaFree(d_C);
      cudaFree(d_labels);
    }
  }
  cudaFree(flush);
  return 0;
}
