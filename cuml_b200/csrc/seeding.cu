// Centroid seeding: random rows, sequential k-means++ (D^2 sampling with greedy trials) and
// k-means|| (distance-weighted Bernoulli over-sampling + weighted reduction to k).
//
// Roles replaced (cuVS side, selected by the reference through KMeansParams::init and
// oversampling_factor, python/cuml/cuml/cluster/kmeans.pyx:66-77): initRandom, kmeansPlusPlus,
// initScalableKMeansPlusPlus (+ sampleCentroids, countSamplesInCluster).  The multi-rank
// random-init split follows python/cuml/cuml/cluster/kmeans_mg.py:63-81.
//
// All random draws come from a counter-based Philox4x32-10 keyed by (seed, stream, GLOBAL row
// index), so a row-sharded run draws exactly the numbers the single-GPU run draws.
#include <limits>
#include <random>
#include <set>

#include "lloyd.cuh"

namespace cb2 {

namespace {

struct Philox {
  __host__ __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1)
  {
    const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c[0];
    const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c[2];
    const uint32_t hi0 = static_cast<uint32_t>(p0 >> 32), lo0 = static_cast<uint32_t>(p0);
    const uint32_t hi1 = static_cast<uint32_t>(p1 >> 32), lo1 = static_cast<uint32_t>(p1);
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  // 4 x 32 random bits for (seed, stream, index)
  __host__ __device__ static inline void gen(uint64_t seed, uint64_t stream, uint64_t index, uint32_t (&out)[4])
  {
    uint32_t c[4] = {static_cast<uint32_t>(index), static_cast<uint32_t>(index >> 32), static_cast<uint32_t>(stream),
                     static_cast<uint32_t>(stream >> 32)};
    uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      round(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
  }
  // uniform in (0, 1)
  __host__ __device__ static inline double u01(uint32_t a, uint32_t b)
  {
    const uint64_t v = (static_cast<uint64_t>(a) << 21) ^ static_cast<uint64_t>(b >> 11);  // 53 bits
    return (static_cast<double>(v & ((1ull << 53) - 1)) + 0.5) * (1.0 / 9007199254740992.0);
  }
};

// ---- k-means||: Bernoulli selection with p_i = min(1, l * w_i * d_i / phi) ----------------------
template <typename T>
__global__ void bernoulli_select_kernel(const T* __restrict__ mind, const T* __restrict__ w, int64_t n,
                                        int64_t global_offset, uint64_t seed, uint64_t stream, double scale /* l/phi */,
                                        int64_t* __restrict__ selected, int* __restrict__ count, int capacity)
{
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t r[4];
  Philox::gen(seed, stream, static_cast<uint64_t>(global_offset + i), r);
  const double u  = Philox::u01(r[0], r[1]);
  const double pr = scale * static_cast<double>(mind[i]) * (w ? static_cast<double>(w[i]) : 1.0);
  if (u < pr) {
    int slot = atomicAdd(count, 1);
    if (slot < capacity) selected[slot] = i;
  }
}

template <typename T>
__global__ void weighted_sum_kernel(const T* __restrict__ v, const T* __restrict__ w, int64_t n, double* __restrict__ out)
{
  __shared__ double red[32];
  double s = 0.0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    s += static_cast<double>(v[i]) * (w ? static_cast<double>(w[i]) : 1.0);
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

template <typename T>
__global__ void fill_kernel(T* p, int64_t n, T v)
{
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---- k-means++: draw `trials` rows with probability ~ w_i * mind_i (exponential race) -----------
template <typename T>
__global__ void kpp_sample_kernel(const T* __restrict__ mind, const T* __restrict__ w, int64_t n, uint64_t seed,
                                  uint64_t stream, int trials, unsigned long long* __restrict__ best /*[trials]*/)
{
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double wt = static_cast<double>(mind[i]) * (w ? static_cast<double>(w[i]) : 1.0);
  if (!(wt > 0.0)) return;
  for (int t = 0; t < trials; t += 2) {
    uint32_t r[4];
    Philox::gen(seed, stream + static_cast<uint64_t>(t / 2), static_cast<uint64_t>(i), r);
    for (int s = 0; s < 2 && t + s < trials; ++s) {
      const double u   = Philox::u01(r[2 * s], r[2 * s + 1]);
      const float key  = static_cast<float>(-log(u) / wt);  // Exp(wt): the minimum is distributed ~ wt
      const unsigned long long packed =
        (static_cast<unsigned long long>(__float_as_uint(key)) << 32) | static_cast<unsigned long long>(i & 0xffffffffu);
      atomicMin(best + t + s, packed);
    }
  }
}

// cost_t = sum_i w_i min(mind_i, ||x_i - cand_t||^2), one warp per row
template <typename T>
__global__ void kpp_trial_cost_kernel(const T* __restrict__ X, const T* __restrict__ w, const T* __restrict__ mind,
                                      int64_t n, int d, const T* __restrict__ cand /*[trials,d]*/, int trials,
                                      double* __restrict__ cost /*[trials]*/)
{
  extern __shared__ double sh_cost[];
  for (int t = threadIdx.x; t < trials; t += blockDim.x) sh_cost[t] = 0.0;
  __syncthreads();
  const int lane     = threadIdx.x % 32;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / 32;
  const int64_t nw   = (static_cast<int64_t>(gridDim.x) * blockDim.x) / 32;
  for (int64_t i = warp; i < n; i += nw) {
    const double wi = w ? static_cast<double>(w[i]) : 1.0;
    const double mi = static_cast<double>(mind[i]);
    for (int t = 0; t < trials; ++t) {
      double s = 0.0;
      for (int c = lane; c < d; c += 32) {
        double df = static_cast<double>(X[i * d + c]) - static_cast<double>(cand[static_cast<int64_t>(t) * d + c]);
        s += df * df;
      }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (lane == 0) atomicAdd(&sh_cost[t], wi * (s < mi ? s : mi));
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < trials; t += blockDim.x) atomicAdd(cost + t, sh_cost[t]);
}

// mind_i = min(mind_i, ||x_i - c||^2), one warp per row
template <typename T>
__global__ void kpp_update_kernel(const T* __restrict__ X, int64_t n, int d, const T* __restrict__ c,
                                  T* __restrict__ mind)
{
  const int lane     = threadIdx.x % 32;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / 32;
  if (warp >= n) return;
  double s = 0.0;
  for (int j = lane; j < d; j += 32) {
    double df = static_cast<double>(X[warp * d + j]) - static_cast<double>(c[j]);
    s += df * df;
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) {
    T v = static_cast<T>(s);
    if (v < mind[warp]) mind[warp] = v;
  }
}

// mind_i = [min(mind_i, ] max(||x_i||^2 + 2 best_i, 0) [)]  -- best_i = 1/2||c||^2 - x_i.c of the nearest new candidate
// (the winning value the fused tensor-core kernel stores next to the label)
__global__ void min_from_best_kernel(const float* __restrict__ xn, const float* __restrict__ best,
                                     float* __restrict__ mind, int64_t n, int fresh)
{
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = fmaxf(xn[i] + 2.0f * best[i], 0.0f);
  mind[i]       = fresh ? v : fminf(mind[i], v);
}

template <typename T>
double weighted_total(Handle& h, const T* v, const T* w, int64_t n)
{
  const int blocks = 256;
  DevBuf<double> part(blocks, h.stream);
  weighted_sum_kernel<T><<<blocks, 256, 0, h.stream>>>(v, w, n, part.get());
  CB2_CHECK_LAUNCH();
  std::vector<double> host(blocks);
  CB2_CUDA(cudaMemcpyAsync(host.data(), part.get(), blocks * sizeof(double), cudaMemcpyDeviceToHost, h.stream));
  CB2_CUDA(cudaStreamSynchronize(h.stream));
  double s = 0.0;
  for (double x : host) s += x;
  return s;
}

// sum over ranks of a host double (through a device cell)
double allreduce_host(Handle& h, double v)
{
  if (h.n_ranks <= 1) return v;
  DevBuf<double> cell(1, h.stream);
  CB2_CUDA(cudaMemcpyAsync(cell.get(), &v, sizeof(double), cudaMemcpyHostToDevice, h.stream));
  comms::allreduce_sum_f64(h, cell.get(), 1);
  CB2_CUDA(cudaMemcpyAsync(&v, cell.get(), sizeof(double), cudaMemcpyDeviceToHost, h.stream));
  CB2_CUDA(cudaStreamSynchronize(h.stream));
  return v;
}

// locate local row `r` (index into the concatenation of the partitions)
template <typename T>
const T* local_row(const std::vector<Part<T>>& parts, int d, int64_t r)
{
  for (auto& p : parts) {
    if (r < p.n) return p.X + r * d;
    r -= p.n;
  }
  throw Error(CUML_B200_INTERNAL_ERROR, "row index out of range in seeding");
}

// gather local rows (sorted local indices) into out [m, d]
template <typename T>
void gather_local(Handle& h, const std::vector<Part<T>>& parts, int d, const std::vector<int64_t>& rows, T* out)
{
  for (size_t i = 0; i < rows.size(); ++i)
    CB2_CUDA(cudaMemcpyAsync(out + i * d, local_row(parts, d, rows[i]), sizeof(T) * d, cudaMemcpyDeviceToDevice,
                             h.stream));
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Random rows.  Multi-rank split as the reference's preflight documents it (kmeans_mg.py:63-81):
// min(G,k) ranks each contribute k // min(G,k) rows, rank 0 also the remainder.
template <typename T>
void init_random(SeedContext<T>& ctx, int k, T* C)
{
  Handle& h     = ctx.h;
  const int G   = h.n_ranks;
  const int S   = std::min(G, k);
  const int per = k / S, rem = k % S;
  const int mine  = (h.rank < S) ? per + (h.rank == 0 ? rem : 0) : 0;
  const int m_max = per + rem;
  CB2_EXPECTS(ctx.n_local >= mine,
              "init='random' requires rank " + std::to_string(h.rank) + " to sample " + std::to_string(mine) +
                " initial centroid(s), but this rank only has " + std::to_string(ctx.n_local) + " row(s)");
  // Floyd's algorithm: `mine` distinct local rows
  std::mt19937_64 gen(ctx.seed * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull * static_cast<uint64_t>(h.rank + 1));
  std::set<int64_t> chosen;
  for (int64_t j = ctx.n_local - mine; j < ctx.n_local; ++j) {
    int64_t t = static_cast<int64_t>(gen() % static_cast<uint64_t>(j + 1));
    if (!chosen.insert(t).second) chosen.insert(j);
  }
  std::vector<int64_t> rows(chosen.begin(), chosen.end());
  DevBuf<T> send(static_cast<size_t>(std::max(m_max, 1)) * ctx.d, h.stream);
  CB2_CUDA(cudaMemsetAsync(send.get(), 0, send.n * sizeof(T), h.stream));
  gather_local(h, ctx.parts, ctx.d, rows, send.get());
  if (G == 1) {
    CB2_CUDA(cudaMemcpyAsync(C, send.get(), sizeof(T) * k * ctx.d, cudaMemcpyDeviceToDevice, h.stream));
    return;
  }
  DevBuf<T> recv(static_cast<size_t>(m_max) * ctx.d * G, h.stream);
  comms::allgather_bytes(h, send.get(), recv.get(), sizeof(T) * m_max * ctx.d);
  int out = 0;
  for (int r = 0; r < S; ++r) {
    const int cnt = per + (r == 0 ? rem : 0);
    CB2_CUDA(cudaMemcpyAsync(C + static_cast<size_t>(out) * ctx.d, recv.get() + static_cast<size_t>(r) * m_max * ctx.d,
                             sizeof(T) * cnt * ctx.d, cudaMemcpyDeviceToDevice, h.stream));
    out += cnt;
  }
}

// ------------------------------------------------------------------------------------------------
// Sequential k-means++ on ONE device-resident array (full data on a single rank, or the k-means||
// candidate set).  Greedy variant: 2 + floor(ln k) trial draws per centre, keep the one that
// lowers the potential most.
template <typename T>
static void kmeans_pp_array(Handle& h, const T* X, const T* w, int64_t n, int d, int k, uint64_t seed, T* C)
{
  CB2_EXPECTS(n >= k, "k-means++ needs at least n_clusters rows");
  // the sampling kernel packs (float key, 32-bit row index) into one 64-bit atomicMin word
  CB2_EXPECTS(n < (int64_t(1) << 32), "sequential k-means++ supports fewer than 2^32 rows per array; use init='k-means||'");
  const int trials = 2 + static_cast<int>(std::floor(std::log(static_cast<double>(k))));
  DevBuf<T> mind(n, h.stream);
  fill_kernel<T><<<static_cast<unsigned>(ceil_div(n, 256)), 256, 0, h.stream>>>(mind.get(), n, T(1));
  CB2_CHECK_LAUNCH();
  DevBuf<unsigned long long> best(trials, h.stream);
  DevBuf<double> cost(trials, h.stream);
  DevBuf<T> cand(static_cast<size_t>(trials) * d, h.stream);
  std::vector<unsigned long long> best_h(trials);
  std::vector<double> cost_h(trials);
  const unsigned row_blocks  = static_cast<unsigned>(ceil_div(n, 256));
  const unsigned warp_blocks = static_cast<unsigned>(std::min<int64_t>(ceil_div(n * 32, 256), h.sm_count * 16));

  // first centre: one draw ~ w (mind == 1 everywhere)
  for (int c = 0; c < k; ++c) {
    const int tr = (c == 0) ? 1 : trials;
    CB2_CUDA(cudaMemsetAsync(best.get(), 0xff, sizeof(unsigned long long) * trials, h.stream));
    kpp_sample_kernel<T><<<row_blocks, 256, 0, h.stream>>>(mind.get(), w, n, seed, 1000003ull * (c + 1), tr, best.get());
    CB2_CHECK_LAUNCH();
    CB2_CUDA(cudaMemcpyAsync(best_h.data(), best.get(), sizeof(unsigned long long) * tr, cudaMemcpyDeviceToHost, h.stream));
    CB2_CUDA(cudaStreamSynchronize(h.stream));
    int chosen_t = 0;
    std::vector<int64_t> idx(tr);
    for (int t = 0; t < tr; ++t) {
      idx[t] = (best_h[t] == ~0ull) ? static_cast<int64_t>(c % n) : static_cast<int64_t>(best_h[t] & 0xffffffffull);
      CB2_CUDA(cudaMemcpyAsync(cand.get() + static_cast<size_t>(t) * d, X + idx[t] * d, sizeof(T) * d,
                               cudaMemcpyDeviceToDevice, h.stream));
    }
    if (tr > 1) {
      CB2_CUDA(cudaMemsetAsync(cost.get(), 0, sizeof(double) * tr, h.stream));
      kpp_trial_cost_kernel<T><<<warp_blocks, 256, sizeof(double) * tr, h.stream>>>(X, w, mind.get(), n, d, cand.get(),
                                                                                   tr, cost.get());
      CB2_CHECK_LAUNCH();
      CB2_CUDA(cudaMemcpyAsync(cost_h.data(), cost.get(), sizeof(double) * tr, cudaMemcpyDeviceToHost, h.stream));
      CB2_CUDA(cudaStreamSynchronize(h.stream));
      for (int t = 1; t < tr; ++t)
        if (cost_h[t] < cost_h[chosen_t]) chosen_t = t;
    }
    CB2_CUDA(cudaMemcpyAsync(C + static_cast<size_t>(c) * d, cand.get() + static_cast<size_t>(chosen_t) * d,
                             sizeof(T) * d, cudaMemcpyDeviceToDevice, h.stream));
    if (c == 0) {
      fill_kernel<T><<<row_blocks, 256, 0, h.stream>>>(mind.get(), n, std::numeric_limits<T>::max());
      CB2_CHECK_LAUNCH();
    }
    kpp_update_kernel<T><<<static_cast<unsigned>(ceil_div(n * 32, 256)), 256, 0, h.stream>>>(
      X, n, d, C + static_cast<size_t>(c) * d, mind.get());
    CB2_CHECK_LAUNCH();
  }
}

template <typename T>
void init_kmeans_plus_plus(SeedContext<T>& ctx, int k, T* C)
{
  CB2_EXPECTS(ctx.h.n_ranks == 1, "init='k-means++' or oversampling_factor=0 not supported for multi-GPU KMeans");
  CB2_EXPECTS(ctx.parts.size() == 1, "sequential k-means++ needs a single contiguous partition");
  kmeans_pp_array<T>(ctx.h, ctx.parts[0].X, ctx.parts[0].w, ctx.parts[0].n, ctx.d, k, ctx.seed, C);
}

// ------------------------------------------------------------------------------------------------
// k-means||  (Bahmani et al.): rounds of distance-weighted Bernoulli over-sampling, then the
// candidates are weighted by the mass they attract and reduced to k with weighted k-means++ +
// Lloyd on the (small) candidate set.
template <typename T>
void init_scalable(SeedContext<T>& ctx, const cuml_b200_kmeans_params_t& params, T* C)
{
  Handle& h    = ctx.h;
  const int d  = ctx.d;
  const int k  = params.n_clusters;
  const double ell = params.oversampling_factor * k;

  // per-partition min-distance buffers
  std::vector<DevBuf<T>> mind(ctx.parts.size());
  for (size_t p = 0; p < ctx.parts.size(); ++p) mind[p].alloc(std::max<int64_t>(ctx.parts[p].n, 1), h.stream);

  std::vector<T> none;
  size_t cap = static_cast<size_t>(std::max<double>(k, ell) * 10 + 1024);
  DevBuf<T> cand(cap * d, h.stream);
  int m = 0;

  // 1. first centre: one uniformly random GLOBAL row, broadcast from its owner
  {
    uint32_t r[4];
    Philox::gen(ctx.seed, 0x5eedull, 0, r);
    const int64_t g0 = static_cast<int64_t>(((static_cast<uint64_t>(r[0]) << 32) | r[1]) % static_cast<uint64_t>(ctx.n_global));
    // owner = rank whose [offset, offset + n_local) contains g0; everyone learns offsets via allgather
    std::vector<int64_t> offs(h.n_ranks + 1, 0);
    if (h.n_ranks > 1) {
      DevBuf<int64_t> sendb(1, h.stream), recvb(h.n_ranks, h.stream);
      CB2_CUDA(cudaMemcpyAsync(sendb.get(), &ctx.n_local, sizeof(int64_t), cudaMemcpyHostToDevice, h.stream));
      comms::allgather_bytes(h, sendb.get(), recvb.get(), sizeof(int64_t));
      std::vector<int64_t> cnt(h.n_ranks);
      CB2_CUDA(cudaMemcpyAsync(cnt.data(), recvb.get(), sizeof(int64_t) * h.n_ranks, cudaMemcpyDeviceToHost, h.stream));
      CB2_CUDA(cudaStreamSynchronize(h.stream));
      for (int i = 0; i < h.n_ranks; ++i) offs[i + 1] = offs[i] + cnt[i];
    } else {
      offs[1] = ctx.n_local;
    }
    int owner = 0;
    while (owner + 1 < h.n_ranks && g0 >= offs[owner + 1]) ++owner;
    if (owner == h.rank)
      CB2_CUDA(cudaMemcpyAsync(cand.get(), local_row(ctx.parts, d, g0 - offs[owner]), sizeof(T) * d,
                               cudaMemcpyDeviceToDevice, h.stream));
    comms::broadcast_bytes(h, cand.get(), sizeof(T) * d, owner);
    m = 1;
  }

  DevBuf<T> cn(cap, h.stream);
  // Default (CUML_B200_SEED_TC=0 restores the CUDA-core kernel; k-means|| at C5 0.62 -> 0.26 s): the min-distance updates of the rounds run on the fused tensor-core
  // kernel (its DIST = 3 epilogue stores the winning value next to the label) instead of the CUDA-core kernel:
  // mind = ||x||^2 + 2 (1/2||c||^2 - x.c).  ||x||^2 is computed once per partition.
  bool seed_tc = false;
  std::vector<DevBuf<float>> xn_tc, best_tc;
  std::vector<DevBuf<int32_t>> lab_tc;
  TcCentroids cen_tc;
  if constexpr (std::is_same<T, float>::value) {
    static const bool env_on = env_flag("CUML_B200_SEED_TC", true);
    seed_tc = env_on && h.cc_major == 10 && engine_from_env(ctx.engine) != ENGINE_SIMT;
    for (auto& pt : ctx.parts)
      if (pt.n > 0 && reinterpret_cast<uintptr_t>(pt.X) % 16 != 0) seed_tc = false;
    if (seed_tc) {
      xn_tc.resize(ctx.parts.size());
      best_tc.resize(ctx.parts.size());
      lab_tc.resize(ctx.parts.size());
      for (size_t p = 0; p < ctx.parts.size(); ++p) {
        const int64_t np = std::max<int64_t>(ctx.parts[p].n, 1);
        xn_tc[p].alloc(np, h.stream);
        best_tc[p].alloc(np, h.stream);
        lab_tc[p].alloc(np, h.stream);
        row_norms<float>(h, ctx.parts[p].X, ctx.parts[p].n, d, xn_tc[p].get());
      }
    }
  }
  auto update_min = [&](int first, int count, bool fresh) {
    if constexpr (std::is_same<T, float>::value) {
      if (seed_tc && tc_best_supported(h, d, count)) {
        const float* cnew = cand.get() + static_cast<size_t>(first) * d;
        tc_prepare(h, cnew, count, d, cen_tc);
        bool cn_ready = false;
        for (size_t p = 0; p < ctx.parts.size(); ++p) {
          auto& pt = ctx.parts[p];
          if (pt.n == 0) continue;
          tc_assign(h, pt.X, pt.n, d, count, cen_tc, lab_tc[p].get(), nullptr, nullptr, best_tc[p].get());
          // the row-packed kernel leaves the winning value of an odd last row to the caller
          const bool tail   = cen_tc.pack == 2 && (pt.n & 1);
          const int64_t nte = pt.n - (tail ? 1 : 0);
          if (nte > 0) {
            min_from_best_kernel<<<static_cast<unsigned>(ceil_div(nte, 256)), 256, 0, h.stream>>>(
              xn_tc[p].get(), best_tc[p].get(), mind[p].get(), nte, fresh ? 1 : 0);
            CB2_CHECK_LAUNCH();
          }
          if (tail) {
            if (!cn_ready) row_norms<T>(h, cnew, count, d, cn.get());
            cn_ready = true;
            if (fresh) simt_assign<T>(h, pt.X + nte * d, 1, d, cnew, count, cn.get(), nullptr, mind[p].get() + nte);
            else simt_min_update<T>(h, pt.X + nte * d, 1, d, cnew, count, cn.get(), mind[p].get() + nte);
          }
        }
        return;
      }
    }
    row_norms<T>(h, cand.get() + static_cast<size_t>(first) * d, count, d, cn.get());
    for (size_t p = 0; p < ctx.parts.size(); ++p) {
      auto& pt = ctx.parts[p];
      if (fresh)
        simt_assign<T>(h, pt.X, pt.n, d, cand.get() + static_cast<size_t>(first) * d, count, cn.get(), nullptr,
                       mind[p].get());
      else
        simt_min_update<T>(h, pt.X, pt.n, d, cand.get() + static_cast<size_t>(first) * d, count, cn.get(),
                           mind[p].get());
    }
  };
  auto potential = [&]() {
    double s = 0.0;
    for (size_t p = 0; p < ctx.parts.size(); ++p)
      if (ctx.parts[p].n > 0) s += weighted_total<T>(h, mind[p].get(), ctx.parts[p].w, ctx.parts[p].n);
    return allreduce_host(h, s);
  };

  update_min(0, 1, true);
  double phi = potential();

  // 2. over-sampling rounds
  int rounds = 0;
  if (phi > 0.0) rounds = std::max(0, std::min(8, static_cast<int>(std::ceil(std::log(phi)))));
  int sel_cap = static_cast<int>(std::min<double>(4.0 * ell + 4096.0, 1 << 24));
  DevBuf<int64_t> sel(sel_cap, h.stream);
  DevBuf<int> sel_count(1, h.stream);
  for (int round = 0; round < rounds && phi > 0.0; ++round) {
    // local Bernoulli draws, partition by partition (global row ids keep the stream shard-invariant)
    std::vector<int64_t> rows;  // local concatenated indices
    int64_t base = 0;
    for (size_t p = 0; p < ctx.parts.size(); ++p) {
      auto& pt = ctx.parts[p];
      if (pt.n > 0) {
        int cnt = 0;
        for (;;) {
          CB2_CUDA(cudaMemsetAsync(sel_count.get(), 0, sizeof(int), h.stream));
          bernoulli_select_kernel<T><<<static_cast<unsigned>(ceil_div(pt.n, 256)), 256, 0, h.stream>>>(
            mind[p].get(), pt.w, pt.n, ctx.row_offset + base, ctx.seed, 0x1000ull + round, ell / phi, sel.get(),
            sel_count.get(), sel_cap);
          CB2_CHECK_LAUNCH();
          CB2_CUDA(cudaMemcpyAsync(&cnt, sel_count.get(), sizeof(int), cudaMemcpyDeviceToHost, h.stream));
          CB2_CUDA(cudaStreamSynchronize(h.stream));
          if (cnt <= sel_cap) break;
          // more picks than slots: which ones were dropped depends on the atomic order, so grow the list and redo the
          // draw (same Philox counters => same picks) instead of keeping an arbitrary subset
          sel_cap = cnt;
          sel.alloc(static_cast<size_t>(sel_cap), h.stream);
        }
        std::vector<int64_t> part_rows(cnt);
        if (cnt) {
          CB2_CUDA(cudaMemcpyAsync(part_rows.data(), sel.get(), sizeof(int64_t) * cnt, cudaMemcpyDeviceToHost, h.stream));
          CB2_CUDA(cudaStreamSynchronize(h.stream));
        }
        std::sort(part_rows.begin(), part_rows.end());  // deterministic candidate order
        for (auto r : part_rows) rows.push_back(base + r);
      }
      base += pt.n;
    }
    // exchange: every rank appends every rank's picks in rank order
    int my_cnt = static_cast<int>(rows.size());
    std::vector<int> counts(h.n_ranks, my_cnt);
    if (h.n_ranks > 1) {
      DevBuf<int> sb(1, h.stream), rb(h.n_ranks, h.stream);
      CB2_CUDA(cudaMemcpyAsync(sb.get(), &my_cnt, sizeof(int), cudaMemcpyHostToDevice, h.stream));
      comms::allgather_bytes(h, sb.get(), rb.get(), sizeof(int));
      CB2_CUDA(cudaMemcpyAsync(counts.data(), rb.get(), sizeof(int) * h.n_ranks, cudaMemcpyDeviceToHost, h.stream));
      CB2_CUDA(cudaStreamSynchronize(h.stream));
    }
    int total = 0, mx = 0;
    for (int c : counts) {
      total += c;
      mx = std::max(mx, c);
    }
    if (total == 0) continue;
    if (static_cast<size_t>(m + total) > cap) {  // grow the candidate buffer
      size_t ncap = std::max(cap * 2, static_cast<size_t>(m + total));
      DevBuf<T> bigger(ncap * d, h.stream);
      CB2_CUDA(cudaMemcpyAsync(bigger.get(), cand.get(), sizeof(T) * m * d, cudaMemcpyDeviceToDevice, h.stream));
      cand = std::move(bigger);
      cn.alloc(ncap, h.stream);
      cap = ncap;
    }
    if (h.n_ranks == 1) {
      gather_local(h, ctx.parts, d, rows, cand.get() + static_cast<size_t>(m) * d);
    } else {
      DevBuf<T> sendb(static_cast<size_t>(std::max(mx, 1)) * d, h.stream);
      DevBuf<T> recvb(static_cast<size_t>(std::max(mx, 1)) * d * h.n_ranks, h.stream);
      CB2_CUDA(cudaMemsetAsync(sendb.get(), 0, sendb.n * sizeof(T), h.stream));
      gather_local(h, ctx.parts, d, rows, sendb.get());
      comms::allgather_bytes(h, sendb.get(), recvb.get(), sizeof(T) * mx * d);
      int out = m;
      for (int r = 0; r < h.n_ranks; ++r) {
        if (counts[r])
          CB2_CUDA(cudaMemcpyAsync(cand.get() + static_cast<size_t>(out) * d, recvb.get() + static_cast<size_t>(r) * mx * d,
                                   sizeof(T) * counts[r] * d, cudaMemcpyDeviceToDevice, h.stream));
        out += counts[r];
      }
      CB2_CUDA(cudaStreamSynchronize(h.stream));  // sendb/recvb go out of scope
    }
    update_min(m, total, false);
    m += total;
    phi = potential();
  }

  // 3. reduce the candidates to k
  if (m < k) {
    // too few candidates (tiny / degenerate data): top up with random rows
    // (k distinct rows are drawn; those that coincide with a candidate already chosen are skipped while others remain)
    DevBuf<T> extra(static_cast<size_t>(k) * d, h.stream);
    init_random<T>(ctx, k, extra.get());
    std::vector<T> hc(static_cast<size_t>(m) * d), he(static_cast<size_t>(k) * d);
    if (m) CB2_CUDA(cudaMemcpyAsync(hc.data(), cand.get(), sizeof(T) * m * d, cudaMemcpyDeviceToHost, h.stream));
    CB2_CUDA(cudaMemcpyAsync(he.data(), extra.get(), sizeof(T) * k * d, cudaMemcpyDeviceToHost, h.stream));
    CB2_CUDA(cudaStreamSynchronize(h.stream));
    std::vector<int> fresh, dup;
    for (int e = 0; e < k; ++e) {
      bool same = false;
      for (int c = 0; c < m && !same; ++c) same = std::memcmp(&he[static_cast<size_t>(e) * d], &hc[static_cast<size_t>(c) * d], sizeof(T) * d) == 0;
      (same ? dup : fresh).push_back(e);
    }
    fresh.insert(fresh.end(), dup.begin(), dup.end());   // duplicates only when nothing else is left
    CB2_CUDA(cudaMemcpyAsync(C, cand.get(), sizeof(T) * m * d, cudaMemcpyDeviceToDevice, h.stream));
    for (int j = 0; j < k - m; ++j)
      CB2_CUDA(cudaMemcpyAsync(C + static_cast<size_t>(m + j) * d, extra.get() + static_cast<size_t>(fresh[j]) * d,
                               sizeof(T) * d, cudaMemcpyDeviceToDevice, h.stream));
    CB2_CUDA(cudaStreamSynchronize(h.stream));
    return;
  }
  if (m == k) {
    CB2_CUDA(cudaMemcpyAsync(C, cand.get(), sizeof(T) * k * d, cudaMemcpyDeviceToDevice, h.stream));
    return;
  }
  // candidate weights = mass of the points each candidate attracts (countSamplesInCluster role)
  DevBuf<double> cw64(m, h.stream);
  CB2_CUDA(cudaMemsetAsync(cw64.get(), 0, sizeof(double) * m, h.stream));
  {
    LloydSolver<T> all(h, ctx.parts, d, m, ctx.engine);
    all.assign(cand.get());
    for (size_t pi = 0; pi < ctx.parts.size(); ++pi)
      weighted_histogram<T>(h, all.labels(pi), ctx.parts[pi].w, ctx.parts[pi].n, m, cw64.get());
    comms::allreduce_sum_f64(h, cw64.get(), m);
  }
  std::vector<double> cw_h(m);
  CB2_CUDA(cudaMemcpyAsync(cw_h.data(), cw64.get(), sizeof(double) * m, cudaMemcpyDeviceToHost, h.stream));
  CB2_CUDA(cudaStreamSynchronize(h.stream));
  std::vector<T> cw_t(m);
  for (int i = 0; i < m; ++i) cw_t[i] = static_cast<T>(cw_h[i]);
  DevBuf<T> cw(m, h.stream);
  CB2_CUDA(cudaMemcpyAsync(cw.get(), cw_t.data(), sizeof(T) * m, cudaMemcpyHostToDevice, h.stream));

  // weighted k-means++ then weighted Lloyd on the candidate set, on this rank only (no collectives)
  {
    SoloGuard solo(h);
    kmeans_pp_array<T>(h, cand.get(), cw.get(), m, d, k, ctx.seed ^ 0xC0FFEEull, C);
    std::vector<Part<T>> cp{Part<T>{cand.get(), static_cast<int64_t>(m), cw.get()}};
    LloydSolver<T> small(h, cp, d, k, ENGINE_SIMT);
    small.run(C, std::max(1, params.max_iter), params.tol);
  }
  // identical centroids on every rank by construction of the inputs; broadcast rank 0's to be safe
  comms::broadcast_bytes(h, C, sizeof(T) * k * d, 0);
  CB2_CUDA(cudaStreamSynchronize(h.stream));
}

#define INST(T)                                                                  \
  template void init_random<T>(SeedContext<T>&, int, T*);                        \
  template void init_kmeans_plus_plus<T>(SeedContext<T>&, int, T*);              \
  template void init_scalable<T>(SeedContext<T>&, const cuml_b200_kmeans_params_t&, T*);
INST(float)
INST(double)
#undef INST

}  // namespace cb2
