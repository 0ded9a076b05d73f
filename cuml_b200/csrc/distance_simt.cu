// Generic CUDA-core distance kernels: nearest-centroid assignment, min-distance update (seeding)
// and the dense distance matrix (transform).  These serve shapes the tcgen05 engine does not
// take (n_features % 4 != 0, fp64) and the seeding/transform paths; the Lloyd hot path for fp32
// is fused_l2_argmin_sm100.cu.
//
// Roles replaced (reference call sites): cuvs fusedL2NN / pairwise_distance reached from
// cpp/src/kmeans/kmeans_predict.cu:41-42 and cpp/src/kmeans/kmeans_transform.cu:32.
#include <type_traits>
#include "kernels.cuh"

namespace cb2 {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;  // 256 threads, 4x4 micro-tile

enum { MODE_ASSIGN = 0, MODE_MINUPDATE = 1, MODE_MATRIX = 2 };

template <typename T>
struct Inf;
template <>
struct Inf<float> {
  static __device__ float v() { return __int_as_float(0x7f800000); }
};
template <>
struct Inf<double> {
  static __device__ double v() { return __longlong_as_double(0x7ff0000000000000LL); }
};

template <typename T, int MODE>
__global__ void __launch_bounds__(256) pairwise_kernel(const T* __restrict__ X, int64_t n, int d,
                                                       const T* __restrict__ C, int k,
                                                       const T* __restrict__ cnorm,
                                                       int32_t* __restrict__ labels, T* __restrict__ mind,
                                                       T* __restrict__ out, int take_sqrt)
{
  __shared__ T Xs[BK][BM + 4];
  __shared__ T Cs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx  = tid % 16;  // column group
  const int ty  = tid / 16;  // row group
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * BM;

  T best[TM];
  int bidx[TM];
  T xn[TM];
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    best[i] = Inf<T>::v();
    bidx[i] = 0;
    xn[i]   = T(0);
  }

  const int n_tiles = (k + BN - 1) / BN;
  for (int nt = 0; nt < n_tiles; ++nt) {
    T acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = T(0);
    if (nt == 0) {
#pragma unroll
      for (int i = 0; i < TM; ++i) xn[i] = T(0);
    }

    for (int k0 = 0; k0 < d; k0 += BK) {
      // cooperative load: 64x16 elements of X and of C, transposed into [kk][row]
#pragma unroll
      for (int l = 0; l < (BM * BK) / 256; ++l) {
        int e  = tid + l * 256;
        int r  = e / BK;
        int kk = e % BK;
        int64_t gr = row0 + r;
        int gc     = k0 + kk;
        Xs[kk][r]  = (gr < n && gc < d) ? X[gr * d + gc] : T(0);
        int cr     = nt * BN + r;
        Cs[kk][r]  = (cr < k && gc < d) ? C[static_cast<int64_t>(cr) * d + gc] : T(0);
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        T a[TM], b[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = Xs[kk][ty * TM + i];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Cs[kk][tx * TN + j];
        if (nt == 0) {
#pragma unroll
          for (int i = 0; i < TM; ++i) xn[i] += a[i] * a[i];
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] += a[i] * b[j];
      }
      __syncthreads();
    }

    if (MODE == MODE_MATRIX) {
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        int64_t gr = row0 + ty * TM + i;
        if (gr >= n) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          int gc = nt * BN + tx * TN + j;
          if (gc < k) {
            T v = xn[i] + cnorm[gc] - T(2) * acc[i][j];
            v   = v < T(0) ? T(0) : v;
            out[gr * k + gc] = take_sqrt ? sqrt(v) : v;
          }
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        int gc = nt * BN + tx * TN + j;
        if (gc < k) {
          T cn = cnorm[gc];
#pragma unroll
          for (int i = 0; i < TM; ++i) {
            T v = cn - T(2) * acc[i][j];
            if (v < best[i]) {  // strict <, ascending index => first minimum
              best[i] = v;
              bidx[i] = gc;
            }
          }
        }
      }
    }
  }

  if (MODE != MODE_MATRIX) {
    // reduce across the 16 threads (consecutive lanes) that share a row group
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      T v   = best[i];
      int b = bidx[i];
#pragma unroll
      for (int off = 8; off >= 1; off >>= 1) {
        T ov   = __shfl_xor_sync(0xffffffffu, v, off);
        int ob = __shfl_xor_sync(0xffffffffu, b, off);
        if (ov < v || (ov == v && ob < b)) {
          v = ov;
          b = ob;
        }
      }
      int64_t gr = row0 + ty * TM + i;
      if (tx == 0 && gr < n) {
        T dist = xn[i] + v;
        dist   = dist < T(0) ? T(0) : dist;
        if (MODE == MODE_ASSIGN) {
          if (labels) labels[gr] = b;
          if (mind) mind[gr] = dist;
        } else {
          T old    = mind[gr];
          mind[gr] = dist < old ? dist : old;
        }
      }
    }
  }
}

template <typename T>
__global__ void row_norms_kernel(const T* __restrict__ A, int64_t rows, int d, T* __restrict__ out)
{
  // one warp per row; fp64 accumulation so the norm is the correctly rounded one
  int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / 32;
  int lane    = threadIdx.x % 32;
  if (row >= rows) return;
  double s = 0.0;
  for (int c = lane; c < d; c += 32) {
    double v = static_cast<double>(A[row * d + c]);
    s += v * v;
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) out[row] = static_cast<T>(s);
}

// fp32 rows of 4..128 features (a multiple of 4, 16-byte aligned): LPR lanes per row, one float4 per lane and step,
// so a warp instruction covers 32 / LPR rows with coalesced 16-byte loads (the warp-per-row kernel above spends a
// whole warp and five shuffles on a 64-byte row: 14.6 ms for C5's 200M x 16 matrix, ~0.9 TB/s).  fp64 accumulation.
template <int LPR>
__global__ void row_norms_vec4_kernel(const float* __restrict__ A, int64_t rows, int d, float* __restrict__ out)
{
  const int64_t gtid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t row  = gtid / LPR;
  const int sub      = static_cast<int>(gtid % LPR);
  double s = 0.0;
  if (row < rows) {
    const float* a = A + row * d;
    for (int c = sub * 4; c < d; c += LPR * 4) {
      const float4 v = *reinterpret_cast<const float4*>(a + c);
      s += static_cast<double>(v.x) * v.x + static_cast<double>(v.y) * v.y + static_cast<double>(v.z) * v.z +
           static_cast<double>(v.w) * v.w;
    }
  }
#pragma unroll
  for (int off = LPR / 2; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (sub == 0 && row < rows) out[row] = static_cast<float>(s);
}

}  // namespace

template <typename T>
void row_norms(Handle& h, const T* A, int64_t rows, int d, T* out)
{
  if (rows == 0) return;
  if constexpr (std::is_same<T, float>::value) {
    if (d % 4 == 0 && d <= 128 && reinterpret_cast<uintptr_t>(A) % 16 == 0) {
      const int chunks = d / 4;
      auto launch = [&](auto kern, int lpr) {
        kern<<<static_cast<unsigned>(ceil_div(rows * lpr, 256)), 256, 0, h.stream>>>(A, rows, d, out);
      };
      if (chunks <= 1) launch(row_norms_vec4_kernel<1>, 1);
      else if (chunks <= 2) launch(row_norms_vec4_kernel<2>, 2);
      else if (chunks <= 4) launch(row_norms_vec4_kernel<4>, 4);
      else if (chunks <= 8) launch(row_norms_vec4_kernel<8>, 8);
      else if (chunks <= 16) launch(row_norms_vec4_kernel<16>, 16);
      else launch(row_norms_vec4_kernel<32>, 32);
      CB2_CHECK_LAUNCH();
      return;
    }
  }
  int64_t threads = rows * 32;
  row_norms_kernel<T><<<static_cast<unsigned>(ceil_div(threads, 256)), 256, 0, h.stream>>>(A, rows, d, out);
  CB2_CHECK_LAUNCH();
}

template <typename T>
void simt_assign(Handle& h, const T* X, int64_t n, int d, const T* C, int k, const T* cnorm,
                 int32_t* labels, T* mind)
{
  if (n == 0) return;
  pairwise_kernel<T, MODE_ASSIGN><<<static_cast<unsigned>(ceil_div(n, BM)), 256, 0, h.stream>>>(
    X, n, d, C, k, cnorm, labels, mind, nullptr, 0);
  CB2_CHECK_LAUNCH();
}

template <typename T>
void simt_min_update(Handle& h, const T* X, int64_t n, int d, const T* C, int k, const T* cnorm, T* mind)
{
  if (n == 0) return;
  pairwise_kernel<T, MODE_MINUPDATE><<<static_cast<unsigned>(ceil_div(n, BM)), 256, 0, h.stream>>>(
    X, n, d, C, k, cnorm, nullptr, mind, nullptr, 0);
  CB2_CHECK_LAUNCH();
}

template <typename T>
void simt_transform(Handle& h, const T* X, int64_t n, int d, const T* C, int k, const T* cnorm,
                    T* out, bool take_sqrt)
{
  if (n == 0) return;
  pairwise_kernel<T, MODE_MATRIX><<<static_cast<unsigned>(ceil_div(n, BM)), 256, 0, h.stream>>>(
    X, n, d, C, k, cnorm, nullptr, nullptr, out, take_sqrt ? 1 : 0);
  CB2_CHECK_LAUNCH();
}

#define INST(T)                                                                                         \
  template void row_norms<T>(Handle&, const T*, int64_t, int, T*);                                      \
  template void simt_assign<T>(Handle&, const T*, int64_t, int, const T*, int, const T*, int32_t*, T*); \
  template void simt_min_update<T>(Handle&, const T*, int64_t, int, const T*, int, const T*, T*);       \
  template void simt_transform<T>(Handle&, const T*, int64_t, int, const T*, int, const T*, T*, bool);
INST(float)
INST(double)
#undef INST

}  // namespace cb2
