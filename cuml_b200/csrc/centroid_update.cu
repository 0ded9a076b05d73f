// M-step of the Lloyd iteration: per-cluster weighted sums, cluster weights and the exact
// inertia sum_i w_i ||x_i - c_label(i)||^2, in ONE pass over X, with no global atomics on the
// hot tile.
//
// Roles replaced (cuVS side, reached from reference cpp/src/kmeans/kmeans_fit.cu:58-59,153-154):
// reduce_rows_by_key (centroid sums), reduce_cols_by_key (cluster weights), computeClusterCost
// (inertia), the divide / keep-old-centroid step and the squared centroid shift.
//
// Design: every CTA owns a private [k x DS] table of partial sums in shared memory (DS = a
// column slice of the feature dimension chosen so the table fits).  Inside the CTA each WARP
// owns a disjoint sub-range of the slice's columns, so two warps never touch the same table
// cell; rows that a warp processes in the same instruction and that share a label are ordered
// with __match_any_sync (segmented, in-warp).  X is read with coalesced 16-byte loads, four row
// batches in flight per lane.  CTA tables are written once to a partials buffer and summed in a
// fixed order in fp64 (deterministic; bitwise identical on every rank after the all-reduce).
#include "kernels.cuh"

namespace cb2 {

namespace {

template <typename T, int VEC>
struct Vec;
template <>
struct Vec<float, 4> {
  using type = float4;
};
template <>
struct Vec<float, 1> {
  using type = float;
};
template <>
struct Vec<double, 2> {
  using type = double2;
};
template <>
struct Vec<double, 1> {
  using type = double;
};

template <typename T, int VEC>
__device__ __forceinline__ void load_vec(const T* p, T (&v)[VEC])
{
  using V = typename Vec<T, VEC>::type;
  V t     = *reinterpret_cast<const V*>(p);
  const T* s = reinterpret_cast<const T*>(&t);
#pragma unroll
  for (int i = 0; i < VEC; ++i) v[i] = s[i];
}

constexpr int UNROLL = 4;

// grid: (row_blocks, slices).  block: warps*32 threads.  dynamic smem: k*ds*sizeof(T) + k*sizeof(T)
template <typename T, int VEC, bool SUMS>
__global__ void __launch_bounds__(256) accumulate_kernel(
  const T* __restrict__ X, int64_t n, int d, const int32_t* __restrict__ labels, const T* __restrict__ w,
  const T* __restrict__ C_old, int k, int ds, int cw, int64_t rows_per_block, T* __restrict__ partial_S,
  T* __restrict__ partial_W, double* __restrict__ partial_I)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* tab  = reinterpret_cast<T*>(smem_raw);      // [k][ds]
  T* wtab = tab + static_cast<size_t>(k) * ds;   // [k]
  __shared__ double red[8];

  const int warp   = threadIdx.x / 32;
  const int lane   = threadIdx.x % 32;
  const int nwarps = blockDim.x / 32;
  const int slice  = blockIdx.y;
  const int cs     = slice * ds;                    // first column of the slice
  const int ce     = min(d, cs + ds);               // one past the last column
  const bool count_here = (slice == 0);

  if (SUMS) {
    for (int i = threadIdx.x; i < k * ds; i += blockDim.x) tab[i] = T(0);
    for (int i = threadIdx.x; i < k; i += blockDim.x) wtab[i] = T(0);
    __syncthreads();
  }

  // columns owned by this warp
  const int c0 = cs + warp * cw;
  const int c1 = min(ce, c0 + cw);
  const int width = max(0, c1 - c0);
  // lanes per row: smallest power of two covering width/VEC (<= 32)
  int L = 1;
  while (L < 32 && L * VEC < width) L <<= 1;
  const int R   = 32 / L;            // rows per warp instruction
  const int g   = lane / L;          // row group of this lane
  const int lr  = lane % L;          // lane within the row
  const unsigned below = (g == 0) ? 0u : ((1u << (g * L)) - 1u);  // lanes of earlier row groups

  const int64_t r_begin = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  const int64_t r_end   = min(n, r_begin + rows_per_block);

  double inertia = 0.0;

  if (width > 0) {
    for (int64_t rb = r_begin; rb < r_end; rb += static_cast<int64_t>(R) * UNROLL) {
      int lab[UNROLL];
      T wv[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        int64_t r = rb + static_cast<int64_t>(u) * R + g;
        bool ok   = r < r_end;
        lab[u]    = ok ? labels[r] : (-1 - g);
        wv[u]     = (ok && w) ? w[r] : T(1);
      }
      // each lane may own several VEC chunks when the warp's width exceeds 32*VEC columns; the
      // trip count is warp-uniform so the full-mask warp primitives below are legal
      const int n_chunks = (width + L * VEC - 1) / (L * VEC);
      for (int ch = 0; ch < n_chunks; ++ch) {
        const int cb      = c0 + (ch * L + lr) * VEC;
        const bool col_ok = cb < c1;
        T xv[UNROLL][VEC];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          int64_t r = rb + static_cast<int64_t>(u) * R + g;
          if (r < r_end && col_ok) {
            if (VEC > 1 && cb + VEC <= c1) {
              load_vec<T, VEC>(X + r * d + cb, xv[u]);
            } else {
#pragma unroll
              for (int i = 0; i < VEC; ++i) xv[u][i] = (cb + i < c1) ? X[r * d + cb + i] : T(0);
            }
          } else {
#pragma unroll
            for (int i = 0; i < VEC; ++i) xv[u][i] = T(0);
          }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          int64_t r  = rb + static_cast<int64_t>(u) * R + g;
          bool ok    = (r < r_end) && col_ok;
          int lb     = lab[u];
          // exact distance to the centroid this row was assigned to (the one it was labelled with)
          if (ok) {
            T part = T(0);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
              if (cb + i < c1) {
                T df = xv[u][i] - C_old[static_cast<int64_t>(lb) * d + cb + i];
                part += df * df;
              }
            }
            inertia += static_cast<double>(part * wv[u]);
          }
          if (SUMS) {
            int rank = 0, maxrank = 0;
            if (R > 1) {
              unsigned peers = __match_any_sync(0xffffffffu, lb);
              rank           = __popc(peers & below) / L;
              maxrank        = __reduce_max_sync(0xffffffffu, rank);
            }
            for (int rr = 0; rr <= maxrank; ++rr) {
              if (ok && rank == rr) {
                T* cell = tab + static_cast<size_t>(lb) * ds + (cb - cs);
#pragma unroll
                for (int i = 0; i < VEC; ++i)
                  if (cb + i < c1) cell[i] += xv[u][i] * wv[u];
                if (count_here && warp == 0 && lr == 0 && ch == 0) wtab[lb] += wv[u];
              }
              if (R > 1) __syncwarp();
            }
          }
        }
      }
    }
  }

  // block reduction of the inertia partial (fixed order)
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) inertia += __shfl_xor_sync(0xffffffffu, inertia, off);
  if (lane == 0) red[warp] = inertia;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < nwarps; ++i) s += red[i];
    partial_I[static_cast<size_t>(blockIdx.x) * gridDim.y + slice] = s;
  }
  if (SUMS) {
    T* outS = partial_S + static_cast<size_t>(blockIdx.x) * k * d;
    const int wcols = ce - cs;
    for (int i = threadIdx.x; i < k * wcols; i += blockDim.x) {
      int j = i / wcols, c = i % wcols;
      outS[static_cast<size_t>(j) * d + cs + c] = tab[static_cast<size_t>(j) * ds + c];
    }
    if (count_here) {
      T* outW = partial_W + static_cast<size_t>(blockIdx.x) * k;
      for (int i = threadIdx.x; i < k; i += blockDim.x) outW[i] = wtab[i];
    }
  }
}

// exact inertia sum_i w_i ||x_i - c_label(i)||^2 (difference form), rows split across warps, the
// (small, hot) centroid table gathered through L1.  One fp64 partial per block.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) inertia_kernel(const T* __restrict__ X, int64_t n, int d,
                                                      const int32_t* __restrict__ labels, const T* __restrict__ w,
                                                      const T* __restrict__ C, double* __restrict__ partial)
{
  __shared__ double red[8];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  int L = 1;
  while (L < 32 && L * VEC < d) L <<= 1;
  const int R  = 32 / L;
  const int g  = lane / L;
  const int lr = lane % L;
  const int64_t gwarp  = static_cast<int64_t>(blockIdx.x) * (blockDim.x / 32) + warp;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * (blockDim.x / 32);
  const int n_chunks   = (d + L * VEC - 1) / (L * VEC);
  double acc = 0.0;
  for (int64_t rb = gwarp * R * UNROLL; rb < n; rb += nwarps * R * UNROLL) {
    int lab[UNROLL];
    T wv[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t r = rb + static_cast<int64_t>(u) * R + g;
      lab[u] = (r < n) ? labels[r] : 0;
      wv[u]  = (r < n) ? (w ? w[r] : T(1)) : T(0);
    }
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int cb = (ch * L + lr) * VEC;
      T xv[UNROLL][VEC], cv[UNROLL][VEC];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int64_t r = rb + static_cast<int64_t>(u) * R + g;
        const bool ok   = (r < n) && (cb < d);
        if (ok && VEC > 1 && cb + VEC <= d) {
          load_vec<T, VEC>(X + r * d + cb, xv[u]);
          load_vec<T, VEC>(C + static_cast<int64_t>(lab[u]) * d + cb, cv[u]);
        } else {
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            const bool in = ok && (cb + i < d);
            xv[u][i] = in ? X[r * d + cb + i] : T(0);
            cv[u][i] = in ? C[static_cast<int64_t>(lab[u]) * d + cb + i] : T(0);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        T part = T(0);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const T df = xv[u][i] - cv[u][i];
          part += df * df;
        }
        acc += static_cast<double>(part * wv[u]);
      }
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < blockDim.x / 32; ++i) s += red[i];
    partial[blockIdx.x] = s;
  }
}

__global__ void sum_partials_kernel(const double* __restrict__ partial, int m, double* __restrict__ cell, int accumulate_into)
{
  // single thread block, fixed order
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) s += partial[i];
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < blockDim.x / 32; ++i) t += red[i];
    *cell = accumulate_into ? *cell + t : t;
  }
}

// huge-k fallback (table does not fit even one VEC-wide slice): global fp64 atomics
template <typename T>
__global__ void accumulate_atomic_kernel(const T* __restrict__ X, int64_t n, int d,
                                         const int32_t* __restrict__ labels, const T* __restrict__ w,
                                         const T* __restrict__ C_old, int k, double* __restrict__ packed, int sums)
{
  int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / 32;
  int lane    = threadIdx.x % 32;
  if (row >= n) return;
  int lb    = labels[row];
  double wv = w ? static_cast<double>(w[row]) : 1.0;
  double part = 0.0;
  for (int c = lane; c < d; c += 32) {
    double x  = static_cast<double>(X[row * d + c]);
    double df = x - static_cast<double>(C_old[static_cast<int64_t>(lb) * d + c]);
    part += df * df;
    if (sums) atomicAdd(packed + static_cast<int64_t>(lb) * d + c, x * wv);
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
  if (lane == 0) {
    atomicAdd(packed + static_cast<int64_t>(k) * d + k, part * wv);
    if (sums) atomicAdd(packed + static_cast<int64_t>(k) * d + lb, wv);
  }
}

// packed[e] (+)= sum_b partial[b][e], fixed order, fp64
template <typename T>
__global__ void reduce_partials_kernel(const T* __restrict__ partial_S, const T* __restrict__ partial_W,
                                       const double* __restrict__ partial_I, int row_blocks, int slices, int k,
                                       int d, double* __restrict__ packed, int accumulate_into, int sums)
{
  const int64_t kd    = static_cast<int64_t>(k) * d;
  const int64_t total = kd + k + 1;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    double s = 0.0;
    if (e < kd) {
      if (!sums) continue;
      for (int b = 0; b < row_blocks; ++b) s += static_cast<double>(partial_S[static_cast<int64_t>(b) * kd + e]);
    } else if (e < kd + k) {
      if (!sums) continue;
      for (int b = 0; b < row_blocks; ++b) s += static_cast<double>(partial_W[static_cast<int64_t>(b) * k + (e - kd)]);
    } else {
      for (int b = 0; b < row_blocks * slices; ++b) s += partial_I[b];
    }
    packed[e] = accumulate_into ? packed[e] + s : s;
  }
}

// C_new = S / W (W > 0) else C_old; shift2 = sum (C_new - C_old)^2.  Multi-block (round 1's single 1024-thread block
// took 20.8 us at k*d = 16 384, a fifth of the non-kernel tail of an 8-GPU iteration): every block writes its partial of
// the squared shift, the last block to finish adds the partials in block order -- deterministic, no float atomics.
template <typename T>
__global__ void __launch_bounds__(256) finalize_kernel(const double* __restrict__ packed, T* __restrict__ C, int k,
                                                       int d, double* __restrict__ shift2_out,
                                                       double* __restrict__ block_shift, unsigned* __restrict__ done)
{
  __shared__ double red[8];
  __shared__ bool last;
  const int64_t kd   = static_cast<int64_t>(k) * d;
  const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
  double acc         = 0.0;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < kd; e += step) {
    int j      = static_cast<int>(e / d);
    double wj  = packed[kd + j];
    T old      = C[e];
    T nw       = old;
    if (wj > 0.0) nw = static_cast<T>(packed[e] / wj);  // empty cluster keeps its previous centroid
    double df = static_cast<double>(nw) - static_cast<double>(old);
    acc += df * df;
    C[e] = nw;
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
    *done = 0;   // the next finalize on this stream starts after this kernel
  }
}

template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ X, int d, const int64_t* __restrict__ idx, int m,
                                   T* __restrict__ out)
{
  int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= static_cast<int64_t>(m) * d) return;
  int r = static_cast<int>(e / d), c = static_cast<int>(e % d);
  out[e] = X[idx[r] * d + c];
}

template <typename T>
__global__ void sum_kernel(const T* __restrict__ w, int64_t n, double* __restrict__ out)
{
  __shared__ double red[32];
  double s = 0.0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    s += static_cast<double>(w[i]);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < blockDim.x / 32; ++i) t += red[i];
    out[blockIdx.x] = t;
  }
}

__global__ void widen_labels_kernel(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out)
{
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

template <typename T>
__global__ void weighted_hist_kernel(const int32_t* __restrict__ labels, const T* __restrict__ w, int64_t n, int k,
                                     double* __restrict__ out)
{
  // fallback for tables that do not fit shared memory: global fp64 atomics
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    atomicAdd(out + labels[i], w ? static_cast<double>(w[i]) : 1.0);
}

// candidate weights of k-means|| (rows per candidate, or their weight sums): a per-block table in shared memory --
// integer counts when unweighted (native shared-memory atomics, exact), fp64 otherwise -- flushed with one global
// atomic per non-empty bin.  (The global-atomic kernel above took 9.8 ms for 200M labels over ~640 bins.)
template <typename T, bool HAS_W>
__global__ void weighted_hist_smem_kernel(const int32_t* __restrict__ labels, const T* __restrict__ w, int64_t n, int k,
                                          double* __restrict__ out)
{
  extern __shared__ unsigned char hist_raw[];
  double* hd         = reinterpret_cast<double*>(hist_raw);
  unsigned int* hc   = reinterpret_cast<unsigned int*>(hist_raw);
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    if (HAS_W) hd[j] = 0.0;
    else hc[j] = 0u;
  }
  __syncthreads();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (HAS_W) atomicAdd(hd + labels[i], static_cast<double>(w[i]));
    else atomicAdd(hc + labels[i], 1u);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const double v = HAS_W ? hd[j] : static_cast<double>(hc[j]);
    if (v != 0.0) atomicAdd(out + j, v);
  }
}

}  // namespace

template <typename T>
void update_plan(Handle& h, int64_t n_max, int d, int k, UpdateWorkspace<T>& ws)
{
  constexpr int VECW = (sizeof(T) == 4) ? 4 : 2;
  ws.k = k;
  ws.d = d;
  const size_t budget = (h.smem_optin ? h.smem_optin : 99 * 1024) - 2048;
  // widest slice (multiple of 8 columns when possible) whose [k x ds] table (+k weights) fits
  int64_t max_cols = (static_cast<int64_t>(budget) / static_cast<int64_t>(sizeof(T)) - k) / k;
  ws.atomic_path   = max_cols < 1;
  if (ws.atomic_path) {
    ws.row_blocks = ws.slices = 0;
    return;
  }
  int ds = d;
  if (max_cols < d) {
    ds = static_cast<int>(max_cols);
    if (ds >= 8) ds -= ds % 8;
    else if (ds >= VECW) ds -= ds % VECW;
  }
  // prefer >= 2 resident CTAs per SM when the table is small
  ws.ds     = ds;
  ws.slices = static_cast<int>(ceil_div(d, ds));
  // warps: each owns >= 8 columns (one 32-byte sector per row) when the slice allows it
  int warps = ds / 8;
  if (warps < 1) warps = 1;
  if (warps > 8) warps = 8;
  ws.warps = warps;
  ws.smem  = (static_cast<size_t>(k) * ds + k) * sizeof(T);
  int per_sm = static_cast<int>(std::min<size_t>(16, (h.smem_optin ? 220 * 1024 : 96 * 1024) / std::max<size_t>(ws.smem, 1024)));
  int thread_limit = 2048 / (warps * 32);
  per_sm = std::max(1, std::min(per_sm, thread_limit));
  int64_t want_blocks = static_cast<int64_t>(h.sm_count) * per_sm;
  int64_t rb          = std::max<int64_t>(1, want_blocks / ws.slices);
  // never more row blocks than 1 per 256 rows; bound the partials traffic to ~1/4 of X
  rb = std::min<int64_t>(rb, std::max<int64_t>(1, ceil_div(n_max, 256)));
  rb = std::min<int64_t>(rb, std::max<int64_t>(1, n_max / (4 * static_cast<int64_t>(k)) + 1));
  ws.row_blocks = static_cast<int>(rb);
  ws.partial_S.alloc(static_cast<size_t>(ws.row_blocks) * k * d, h.stream);
  ws.partial_W.alloc(static_cast<size_t>(ws.row_blocks) * k, h.stream);
  ws.partial_I.alloc(static_cast<size_t>(ws.row_blocks) * ws.slices, h.stream);
}

template <typename T, int VEC, bool SUMS>
static void launch_accumulate(Handle& h, UpdateWorkspace<T>& ws, const T* X, int64_t n, int d,
                              const int32_t* labels, const T* w, const T* C_old, int k, int row_blocks)
{
  auto kern = accumulate_kernel<T, VEC, SUMS>;
  size_t smem = SUMS ? ws.smem : 0;
  if (smem > 48 * 1024) CB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int cw = static_cast<int>(ceil_div(ws.ds, ws.warps));
  cw     = static_cast<int>(ceil_div(cw, VEC)) * VEC;
  int64_t rpb = ceil_div(n, row_blocks);
  dim3 grid(row_blocks, ws.slices);
  kern<<<grid, ws.warps * 32, smem, h.stream>>>(X, n, d, labels, w, C_old, k, ws.ds, cw, rpb, ws.partial_S.get(),
                                                ws.partial_W.get(), ws.partial_I.get());
  CB2_CHECK_LAUNCH();
}

template <typename T>
void update_accumulate(Handle& h, UpdateWorkspace<T>& ws, const T* X, int64_t n, int d, const int32_t* labels,
                       const T* w, const T* C_old, int k, double* packed, bool accumulate_into, bool sums)
{
  const int64_t total = static_cast<int64_t>(k) * d + k + 1;
  if (ws.atomic_path) {
    if (!accumulate_into) CB2_CUDA(cudaMemsetAsync(packed, 0, total * sizeof(double), h.stream));
    if (n > 0) {
      accumulate_atomic_kernel<T><<<static_cast<unsigned>(ceil_div(n * 32, 256)), 256, 0, h.stream>>>(
        X, n, d, labels, w, C_old, k, packed, sums ? 1 : 0);
      CB2_CHECK_LAUNCH();
    }
    return;
  }
  if (n == 0) {
    if (!accumulate_into) CB2_CUDA(cudaMemsetAsync(packed, 0, total * sizeof(double), h.stream));
    return;
  }
  constexpr int VECW = (sizeof(T) == 4) ? 4 : 2;
  const bool vec_ok  = (d % VECW == 0) && (ws.ds % VECW == 0) &&
                      (reinterpret_cast<uintptr_t>(X) % (VECW * sizeof(T)) == 0);
  int row_blocks = static_cast<int>(std::min<int64_t>(ws.row_blocks, std::max<int64_t>(1, ceil_div(n, 64))));
  if (vec_ok) {
    if (sums) launch_accumulate<T, VECW, true>(h, ws, X, n, d, labels, w, C_old, k, row_blocks);
    else launch_accumulate<T, VECW, false>(h, ws, X, n, d, labels, w, C_old, k, row_blocks);
  } else {
    if (sums) launch_accumulate<T, 1, true>(h, ws, X, n, d, labels, w, C_old, k, row_blocks);
    else launch_accumulate<T, 1, false>(h, ws, X, n, d, labels, w, C_old, k, row_blocks);
  }
  int threads = 256;
  int blocks  = static_cast<int>(std::min<int64_t>(1024, ceil_div(total, threads)));
  reduce_partials_kernel<T><<<blocks, threads, 0, h.stream>>>(ws.partial_S.get(), ws.partial_W.get(),
                                                              ws.partial_I.get(), row_blocks, ws.slices, k, d, packed,
                                                              accumulate_into ? 1 : 0, sums ? 1 : 0);
  CB2_CHECK_LAUNCH();
}

template <typename T>
void compute_inertia(Handle& h, const T* X, int64_t n, int d, const int32_t* labels, const T* w, const T* C,
                     double* cell, bool accumulate_into)
{
  if (n == 0) {
    if (!accumulate_into) CB2_CUDA(cudaMemsetAsync(cell, 0, sizeof(double), h.stream));
    return;
  }
  constexpr int VECW = (sizeof(T) == 4) ? 4 : 2;
  const bool vec_ok  = (d % VECW == 0) && (reinterpret_cast<uintptr_t>(X) % (VECW * sizeof(T)) == 0) &&
                      (reinterpret_cast<uintptr_t>(C) % (VECW * sizeof(T)) == 0);
  const int blocks = static_cast<int>(std::min<int64_t>(static_cast<int64_t>(h.sm_count) * 8, ceil_div(n, 64)));
  DevBuf<double> partial(blocks, h.stream);
  if (vec_ok) inertia_kernel<T, VECW><<<blocks, 256, 0, h.stream>>>(X, n, d, labels, w, C, partial.get());
  else inertia_kernel<T, 1><<<blocks, 256, 0, h.stream>>>(X, n, d, labels, w, C, partial.get());
  CB2_CHECK_LAUNCH();
  sum_partials_kernel<<<1, 256, 0, h.stream>>>(partial.get(), blocks, cell, accumulate_into ? 1 : 0);
  CB2_CHECK_LAUNCH();
}

template <typename T>
void finalize_centroids(Handle& h, const double* packed, T* C, int k, int d, double* shift2_out)
{
  const int64_t kd      = static_cast<int64_t>(k) * d;
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(Handle::FIN_BLOCKS, std::max<int64_t>(1, kd / 512)));
  finalize_kernel<T><<<blocks, 256, 0, h.stream>>>(packed, C, k, d, shift2_out, h.fin_scratch,
                                                   reinterpret_cast<unsigned*>(h.fin_scratch + Handle::FIN_BLOCKS));
  CB2_CHECK_LAUNCH();
}

template <typename T>
void gather_rows(Handle& h, const T* X, int d, const int64_t* idx_dev, int m, T* out)
{
  if (m == 0) return;
  int64_t total = static_cast<int64_t>(m) * d;
  gather_rows_kernel<T><<<static_cast<unsigned>(ceil_div(total, 256)), 256, 0, h.stream>>>(X, d, idx_dev, m, out);
  CB2_CHECK_LAUNCH();
}

template <typename T>
double sum_weights(Handle& h, const T* w, int64_t n)
{
  if (n == 0) return 0.0;
  const int blocks = 256;
  DevBuf<double> part(blocks, h.stream);
  sum_kernel<T><<<blocks, 256, 0, h.stream>>>(w, n, part.get());
  CB2_CHECK_LAUNCH();
  std::vector<double> host(blocks);
  CB2_CUDA(cudaMemcpyAsync(host.data(), part.get(), blocks * sizeof(double), cudaMemcpyDeviceToHost, h.stream));
  CB2_CUDA(cudaStreamSynchronize(h.stream));
  double s = 0.0;
  for (double v : host) s += v;
  return s;
}

void labels_to_i64(Handle& h, const int32_t* in, int64_t n, int64_t* out)
{
  if (n == 0) return;
  widen_labels_kernel<<<static_cast<unsigned>(ceil_div(n, 256)), 256, 0, h.stream>>>(in, n, out);
  CB2_CHECK_LAUNCH();
}

template <typename T>
void weighted_histogram(Handle& h, const int32_t* labels, const T* w, int64_t n, int k, double* out)
{
  if (n == 0) return;
  int blocks = static_cast<int>(std::min<int64_t>(h.sm_count * 8, ceil_div(n, 256)));
  const size_t bytes = static_cast<size_t>(k) * (w ? sizeof(double) : sizeof(unsigned int));
  if (bytes <= 40 * 1024) {   // (a block sees n / blocks rows: its integer counts stay far below 2^32)
    if (w) weighted_hist_smem_kernel<T, true><<<blocks, 256, bytes, h.stream>>>(labels, w, n, k, out);
    else weighted_hist_smem_kernel<T, false><<<blocks, 256, bytes, h.stream>>>(labels, w, n, k, out);
  } else {
    weighted_hist_kernel<T><<<blocks, 256, 0, h.stream>>>(labels, w, n, k, out);
  }
  CB2_CHECK_LAUNCH();
}

#define INST(T)                                                                                              \
  template void update_plan<T>(Handle&, int64_t, int, int, UpdateWorkspace<T>&);                             \
  template void update_accumulate<T>(Handle&, UpdateWorkspace<T>&, const T*, int64_t, int, const int32_t*,   \
                                     const T*, const T*, int, double*, bool, bool);                          \
  template void finalize_centroids<T>(Handle&, const double*, T*, int, int, double*);                        \
  template void compute_inertia<T>(Handle&, const T*, int64_t, int, const int32_t*, const T*, const T*,      \
                                   double*, bool);                                                           \
  template void gather_rows<T>(Handle&, const T*, int, const int64_t*, int, T*);                             \
  template double sum_weights<T>(Handle&, const T*, int64_t);                                                \
  template void weighted_histogram<T>(Handle&, const int32_t*, const T*, int64_t, int, double*);
INST(float)
INST(double)
#undef INST

}  // namespace cb2
