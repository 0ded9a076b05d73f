// Minimal C++ consumer of ML::kmeans::{fit,predict} -- the role of the reference example
// (cpp/examples/kmeans/kmeans_example.cpp:56-253) and its known-answer check (:110-113,172-191):
// rows (1,1),(3,4),(1,2),(2,3), k=2, init=Array from rows 0 and 1, tol 0.05 ->
// labels {0,1,0,1}, centroids {1,1.5, 2.5,3.5}.
//   g++ -std=c++17 -Iinclude examples/kmeans_example.cpp -Lcuml_b200/lib -lcuml_b200 \
//       -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/cuml_b200/lib -o kmeans_example
#include <cuda_runtime.h>
#include <cuml/cluster/kmeans.hpp>

#include <cmath>
#include <cstdio>
#include <vector>

int main()
{
  const int n = 4, d = 2, k = 2;
  std::vector<double> h_X = {1.0, 1.0, 3.0, 4.0, 1.0, 2.0, 2.0, 3.0};
  std::vector<double> h_C = {1.0, 1.0, 3.0, 4.0};
  double *d_X, *d_C;
  int* d_labels;
  cudaMalloc(&d_X, sizeof(double) * n * d);
  cudaMalloc(&d_C, sizeof(double) * k * d);
  cudaMalloc(&d_labels, sizeof(int) * n);
  cudaMemcpy(d_X, h_X.data(), sizeof(double) * n * d, cudaMemcpyHostToDevice);
  cudaMemcpy(d_C, h_C.data(), sizeof(double) * k * d, cudaMemcpyHostToDevice);

  raft::handle_t handle;
  ML::kmeans::KMeansParams params;
  params.n_clusters = k;
  params.init       = ML::kmeans::KMeansParams::InitMethod::Array;
  params.metric     = ML::distance::DistanceType::L2SqrtExpanded;
  params.tol        = 0.05;
  params.max_iter   = 300;
  double inertia = 0;
  int n_iter     = 0;
  ML::kmeans::fit(handle, params, d_X, n, d, nullptr, d_C, inertia, n_iter);
  ML::kmeans::predict(handle, params, d_C, d_X, n, d, nullptr, true, d_labels, inertia);
  handle.sync_stream();

  std::vector<int> labels(n);
  cudaMemcpy(labels.data(), d_labels, sizeof(int) * n, cudaMemcpyDeviceToHost);
  cudaMemcpy(h_C.data(), d_C, sizeof(double) * k * d, cudaMemcpyDeviceToHost);
  const int want_l[4]    = {0, 1, 0, 1};
  const double want_c[4] = {1.0, 1.5, 2.5, 3.5};
  bool ok = true;
  for (int i = 0; i < n; ++i) ok = ok && labels[i] == want_l[i];
  for (int i = 0; i < k * d; ++i) ok = ok && std::fabs(h_C[i] - want_c[i]) <= 1e-12 * std::fabs(want_c[i]);
  std::printf("labels %d %d %d %d  centroids %.3f %.3f %.3f %.3f  inertia %.4f n_iter %d  %s\n", labels[0], labels[1],
              labels[2], labels[3], h_C[0], h_C[1], h_C[2], h_C[3], inertia, n_iter, ok ? "PASSED" : "FAILED");
  cudaFree(d_X);
  cudaFree(d_C);
  cudaFree(d_labels);
  return ok ? 0 : 1;
}
