// M-step for fp32 on sm_100a: per-cluster weighted sums and cluster weights in ONE streaming pass
// over X, TMA-staged, with exclusive table ownership (no atomics of any kind on the hot tile).
//
// Roles replaced (cuVS side, reached from reference cpp/src/kmeans/kmeans_fit.cu:58-59,153-154):
// reduce_rows_by_key (centroid sums) and reduce_cols_by_key (cluster weights).
//
// Two kernels, chosen by tma_update_accumulate():
//   accumulate_owner_kernel  (n_features >= 32, k >= 16, [k x 32*VEC] table fits shared memory; the default
//                            for C2 / C3): consumer warp w owns the table rows of the clusters of class w
//                            (16 size-balanced classes); analyst warps counting-sort each tile's rows by
//                            class on a label ring that runs ahead of the X ring; consumers apply their rows
//                            with lane = column (contiguous, conflict-free table read-modify-write).
//                            HBM-bound: 5.8-5.9 TB/s at C3.
//   accumulate_tma_kernel    (short rows / small k, e.g. C5): lane = row, consumer w owns 8 columns of the
//                            table, rows of one instruction that share a label are ordered by ranks an
//                            analyst warp computes with __match_any_sync; X tile and table XOR-swizzled.
//                            Bound by shared-memory wavefronts (3.3-3.6 TB/s).
// Every kernel writes its table once per CTA to a partials buffer; reduce_partials_f32_kernel sums the
// partials in a fixed order in fp64 (deterministic, bitwise identical on every rank after the all-reduce).
// Memory parallelism comes from the TMA rings, not from occupancy (one CTA per SM).
#include <cstdlib>

#include "kernels.cuh"
#include "ptx.cuh"
#include "tensormap.cuh"

namespace cb2 {

namespace {

constexpr int MAX_NSTAGE = 8;

struct UpdParams {
  int64_t n;
  int64_t tiles_total;      // ceil(n / tr)
  int64_t tiles_per_block;
  int d, k, ds, tr;
  int nb;                   // 128-byte sub-slices per CTA slice (1 or 2) when ds >= 32
  int nstage;               // ring depth (2..8)
  int na;                   // analyst warps (tile t is analysed by analyst t % na)
  uint32_t sub_bytes;       // one sub-tile: tr * min(ds,32) * 4, multiple of 1024
  const int32_t* labels;    // padded: readable up to n + tr
  const float* w;           // or null
  float* partial_S;         // [row_blocks][k][d]
  float* partial_W;         // [row_blocks][k]
};

__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr)
{
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v)
{
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float lds32f(uint32_t addr)
{
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts32f(uint32_t addr, float v)
{
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ int lds32(uint32_t addr)
{
  int v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
// plain spin on an mbarrier phase (the bounded variant in ptx.cuh costs issue slots in hot loops)
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity)
{
  uint32_t spins = 0;
  while (!ptx::mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("cuml_b200: update-kernel mbarrier timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}

// NC = 16-byte chunks per sub-slice row: 8 (128B-swizzled, ds >= 32), 4 (64B), 2 (32B) or 1 (dense).
// Warp roles: 0 = TMA producer, 1..na = analysts (round-robin over tiles), then the consumers.  Lane = row: one warp instruction
// covers 32 consecutive rows.  The analyst computes, once per 32-row group, each row's rank among
// the rows of the group that share its label (match_any) and the group's maximum rank, and keeps the
// per-cluster weights; consumers then apply the group in (max rank + 1) conflict-free rounds with no
// warp-wide matching on their critical path.  Every consumer lane owns CPL 16-byte chunks of its row.
template <int NC, bool HAS_W>
__global__ void __launch_bounds__(416)
accumulate_tma_kernel(const __grid_constant__ CUtensorMap tm_x, const UpdParams p)
{
  constexpr int CPL      = NC >= 2 ? 2 : 1;                           // chunks per lane
  constexpr int SH       = NC == 8 ? 0 : (NC == 4 ? 1 : (NC == 2 ? 2 : 0));  // swizzle: chunk ^= (row >> SH) & (NC-1)
  constexpr uint32_t MSK = NC - 1;
  constexpr uint32_t ROWB = NC * 16;
  constexpr int CONS_PER_SUB = NC / CPL;                              // consumer warps per sub-slice
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t raw  = ptx::smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* g          = smem_dyn + (base - raw);
  // layout: stages (nb sub-tiles | labels | meta) | tables (nb) | wtab | barriers
  const uint32_t lab_bytes  = static_cast<uint32_t>(p.tr) * 4u;
  const uint32_t x_bytes    = static_cast<uint32_t>(p.nb) * p.sub_bytes;
  const uint32_t stage_full = x_bytes + ((2u * lab_bytes + 1023u) & ~1023u);
  const uint32_t tab_bytes  = static_cast<uint32_t>(p.k) * ROWB;       // one sub-slice table
  const uint32_t tab_u32    = base + p.nstage * stage_full;
  float* tab       = reinterpret_cast<float*>(g + p.nstage * stage_full);
  float* wtab      = tab + static_cast<size_t>(p.nb) * p.k * (NC * 4);     // [na][k4]: one weight table per analyst
  const int k4     = (p.k + 3) & ~3;
  uint64_t* bars   = reinterpret_cast<uint64_t*>(wtab + static_cast<size_t>(p.na) * k4);  // full | ready | empty
  const uint32_t bars_u32 = ptx::smem_u32(bars);
  const uint32_t B_FULL = 0, B_READY = MAX_NSTAGE * 8, B_EMPTY = 2 * MAX_NSTAGE * 8;

  const int warp    = threadIdx.x / 32;
  const int lane    = threadIdx.x % 32;
  const int ncons   = blockDim.x / 32 - 1 - p.na;
  const int slice   = blockIdx.y;
  const int cs      = slice * p.ds;

  for (int i = threadIdx.x; i < p.nb * p.k * NC * 4; i += blockDim.x) tab[i] = 0.0f;
  for (int i = threadIdx.x; i < p.na * k4; i += blockDim.x) wtab[i] = 0.0f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_NSTAGE; ++s) {
      ptx::mbar_init(bars_u32 + B_FULL + s * 8, 1);
      ptx::mbar_init(bars_u32 + B_READY + s * 8, 1);
      ptx::mbar_init(bars_u32 + B_EMPTY + s * 8, ncons);
    }
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tm_x);
  }
  __syncthreads();

  const int64_t t_begin = static_cast<int64_t>(blockIdx.x) * p.tiles_per_block;
  const int64_t t_end   = min(p.tiles_total, t_begin + p.tiles_per_block);

  if (warp == 0) {
    // ---------------- producer ----------------
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      for (int64_t t = t_begin; t < t_end; ++t) {
        mbar_wait_spin(bars_u32 + B_EMPTY + s * 8, ph ^ 1u);
        const uint32_t full = bars_u32 + B_FULL + s * 8;
        ptx::mbar_arrive_expect_tx(full, x_bytes + lab_bytes);
        const uint32_t dst = base + s * stage_full;
        const int64_t row0 = t * p.tr;
        for (int sb = 0; sb < p.nb; ++sb)
          ptx::tma_load_2d_hint(dst + sb * p.sub_bytes, &tm_x, cs + sb * 32, static_cast<int32_t>(row0), full,
                                ptx::kEvictFirst);
        bulk_load_1d(dst + x_bytes, p.labels + row0, lab_bytes, full);
        if (++s == static_cast<uint32_t>(p.nstage)) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp <= p.na) {
    // ---------------- analysts: ranks among equal labels per 32-row group; cluster weights ----------------
    const unsigned below = (1u << lane) - 1u;
    const bool counts    = (slice == 0);
    const int me         = warp - 1;
    float* wtab_me       = wtab + static_cast<size_t>(me) * k4;
    uint32_t s = 0, ph = 0;
    for (int64_t t = t_begin; t < t_end; ++t) {
      if (static_cast<int>((t - t_begin) % p.na) != me) {   // another analyst's tile
        if (++s == static_cast<uint32_t>(p.nstage)) { s = 0; ph ^= 1u; }
        continue;
      }
      mbar_wait_spin(bars_u32 + B_FULL + s * 8, ph);
      const uint32_t ls = base + s * stage_full + x_bytes;
      const uint32_t ms = ls + lab_bytes;
      const int64_t row0 = t * p.tr;
      const int64_t left = p.n - row0;
      const int valid    = left < p.tr ? static_cast<int>(left) : p.tr;
#pragma unroll 2
      for (int r = lane; r < p.tr; r += 32) {
        const bool ok        = r < valid;
        const int lb         = ok ? lds32(ls + r * 4) : ~lane;   // invalid rows: unique labels
        const unsigned peers = __match_any_sync(0xffffffffu, lb);
        const int rank       = __popc(peers & below);
        const int maxr       = __reduce_max_sync(0xffffffffu, rank);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(ms + r * 4), "r"(rank | (maxr << 8)) : "memory");
        if (counts) {
          if (!HAS_W) {
            // the last of each set of equal labels adds the set's size: distinct addresses, no conflict
            const int cnt = __popc(peers);
            if (ok && rank == cnt - 1) wtab_me[lb] += static_cast<float>(cnt);
          } else {
            const float wv = ok ? __ldg(p.w + row0 + r) : 0.0f;
            for (int rr = 0; rr <= maxr; ++rr) {
              if (ok && rank == rr) wtab_me[lb] += wv;
              __syncwarp();
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bars_u32 + B_READY + s * 8);
      if (++s == static_cast<uint32_t>(p.nstage)) { s = 0; ph ^= 1u; }
    }
  } else {
    // ---------------- consumers ----------------
    const int cw      = warp - 1 - p.na;
    const int sub     = cw / CONS_PER_SUB;                                   // 128-byte sub-slice
    const uint32_t j0 = static_cast<uint32_t>(cw % CONS_PER_SUB) * CPL;     // first owned logical chunk (even)
    const uint32_t tab_sub = tab_u32 + sub * tab_bytes;
    uint32_t s = 0, ph = 0;
    for (int64_t t = t_begin; t < t_end; ++t) {
      mbar_wait_spin(bars_u32 + B_FULL + s * 8, ph);     // TMA bytes landed
      mbar_wait_spin(bars_u32 + B_READY + s * 8, ph);    // ranks written
      const uint32_t st = base + s * stage_full;
      const uint32_t xs = st + sub * p.sub_bytes;
      const uint32_t ls = st + x_bytes;
      const uint32_t ms = ls + lab_bytes;
      const int64_t row0 = t * p.tr;
      const int64_t left = p.n - row0;
      const int valid    = left < p.tr ? static_cast<int>(left) : p.tr;
#pragma unroll 2
      for (int r = lane; r < p.tr; r += 32) {
        const bool ok    = r < valid;
        const int lb     = ok ? lds32(ls + r * 4) : 0;
        const int meta   = lds32(ms + r * 4);
        const uint32_t xa = xs + static_cast<uint32_t>(r) * ROWB + ((j0 ^ ((static_cast<uint32_t>(r) >> SH) & MSK)) << 4);
        float4 x0 = lds128(xa);
        float4 x1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (CPL == 2) x1 = lds128(xa ^ 16u);
        const int rank   = meta & 0xff;
        const int maxr   = meta >> 8;                       // warp-uniform
        if (HAS_W) {
          const float wv = ok ? __ldg(p.w + row0 + r) : 0.0f;
          x0.x *= wv; x0.y *= wv; x0.z *= wv; x0.w *= wv;
          x1.x *= wv; x1.y *= wv; x1.z *= wv; x1.w *= wv;
        }
        const uint32_t ca = tab_sub + static_cast<uint32_t>(lb) * ROWB + ((j0 ^ ((static_cast<uint32_t>(lb) >> SH) & MSK)) << 4);
        for (int rr = 0; rr <= maxr; ++rr) {
          if (ok && rank == rr) {
            float4 c = lds128(ca);
            float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
            if (CPL == 2) e = lds128(ca ^ 16u);
            c.x += x0.x; c.y += x0.y; c.z += x0.z; c.w += x0.w;
            sts128(ca, c);
            if (CPL == 2) {
              e.x += x1.x; e.y += x1.y; e.z += x1.z; e.w += x1.w;
              sts128(ca ^ 16u, e);
            }
          }
          if (maxr) __syncwarp();
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bars_u32 + B_EMPTY + s * 8);
      if (++s == static_cast<uint32_t>(p.nstage)) { s = 0; ph ^= 1u; }
    }
  }
  __syncthreads();
  float* outS      = p.partial_S + static_cast<size_t>(blockIdx.x) * p.k * p.d;
  const int scols  = NC * 4;                              // columns per sub-slice
  const int wcols  = min(p.nb * scols, p.d - cs);
  for (int i = threadIdx.x; i < p.k * wcols; i += blockDim.x) {
    const int j = i / wcols, c = i % wcols;
    const int sb = c / scols, cc = c % scols;
    const int pc = ((((cc >> 2) ^ ((j >> SH) & MSK)) << 2) | (cc & 3));
    outS[static_cast<size_t>(j) * p.d + cs + c] = tab[static_cast<size_t>(sb) * p.k * scols + static_cast<size_t>(j) * scols + pc];
  }
  if (slice == 0) {
    float* outW = p.partial_W + static_cast<size_t>(blockIdx.x) * p.k;
    for (int i = threadIdx.x; i < p.k; i += blockDim.x) {
      float acc = 0.0f;
      for (int a = 0; a < p.na; ++a) acc += wtab[static_cast<size_t>(a) * k4 + i];
      outW[i] = acc;
    }
  }
}

// ---- label-class ownership ---------------------------------------------------------------------------
// One [k x DS] fp32 table per CTA (DS = 32*VEC columns), NCONS consumer warps.  Consumer w exclusively owns
// the table rows of the clusters with (label % NCONS) == w, so no two warps ever touch the same cell and no
// atomics or rank bookkeeping are needed.  Per 32-row group every consumer reads the 32 labels (lane = row),
// ballots the rows of its class and then walks its rows one by one with lane = column: the X row segment,
// the table row read and the table row write are each ONE contiguous, conflict-free shared-memory access
// (the vectorised variant above pays ~3x the wavefronts for its 32 random table rows per instruction).
// Consecutive rows of one consumer that share a label are ordered by program order of the same lanes.
template <int VEC>
struct VecIO;
template <>
struct VecIO<1> {
  float v[1];
  __device__ __forceinline__ void load(uint32_t a) { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[0]) : "r"(a)); }
  __device__ __forceinline__ void store(uint32_t a) const { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v[0]) : "memory"); }
};
template <>
struct VecIO<2> {
  float v[2];
  __device__ __forceinline__ void load(uint32_t a) { asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "r"(a)); }
  __device__ __forceinline__ void store(uint32_t a) const { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v[0]), "f"(v[1]) : "memory"); }
};
template <>
struct VecIO<4> {
  float v[4];
  __device__ __forceinline__ void load(uint32_t a)
  {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a));
  }
  __device__ __forceinline__ void store(uint32_t a) const
  {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
  }
};

struct UpdParams7 {
  int64_t n;
  int64_t tiles_total;
  int64_t tiles_per_block;
  int d, k, tr, nstage, ncons;   // ncons: power of two
  int nl;                        // depth of the label / row-list ring (>= nstage)
  int dbg_skip;
  const int32_t* labels;
  const float* w;
  const uint8_t* cls_map;        // [k] label -> class (balanced by last iteration's cluster sizes), or null
  float* partial_S;
  float* partial_W;
};

constexpr int OWN_CONS  = 16;   // consumer warps = label classes
constexpr int OWN_NA    = 2;    // analyst warps (alternate tiles)
constexpr int OWN_MAXCH = 8;    // 32-row groups per tile (tile rows <= 256)

__device__ __forceinline__ uint2 lds64u(uint32_t addr)
{
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}

constexpr int OWN_MAXNL = 8;    // label ring depth

// class map for the next launch: clusters sorted by size (descending), dealt to the 16 classes in serpentine
// order, so every consumer warp gets about the same number of rows whatever the cluster-size skew
__global__ void __launch_bounds__(1024)
balance_classes_kernel(const double* __restrict__ W, int k, uint8_t* __restrict__ cls_map)
{
  extern __shared__ float bc_w[];
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const float v = static_cast<float>(W[j]);
    bc_w[j]       = (v == v && v >= 0.f && v < 3.0e38f) ? v : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const float me = bc_w[j];
    int rank = 0;
    for (int i = 0; i < k; ++i) {
      const float o = bc_w[i];
      rank += (o > me || (o == me && i < j)) ? 1 : 0;
    }
    const int q = rank / OWN_CONS, m = rank % OWN_CONS;
    cls_map[j]  = static_cast<uint8_t>((q & 1) ? (OWN_CONS - 1 - m) : m);
  }
}

// Warp roles: 0 = X producer; 1 = label producer; 2..3 = analysts; 4..19 = consumers.
//   Two rings: X tiles (nstage deep, 32..64 KB each) and labels + row lists (nl deep, ~2 KB each).  The
//   label ring runs ahead of the X ring, so the analysts' latency never sits on an X stage's turn-around.
//   analyst  : counting sort of a tile's rows by label class, once per tile: 4 ballots (one per class bit)
//              give lane j < 16 the row mask of class j; the rows land in `perm` (row << 16 | label), class
//              after class, in row order, with the class offsets next to it.  Also keeps the per-cluster
//              row counts of unweighted fits (integer shared-memory reductions).
//   consumer : owns the table rows of its class.  Walks its perm segment with lane = column: per row one
//              contiguous table read, add, write; the X segment of the next row and the next list entries are
//              already in flight.  Two rows per loop iteration (register rotation without moves).
// (A variant for 64-byte rows -- two list rows per warp instruction -- was parity-green but measured slower than the
// lane = row kernel at C5, 6.1 against 3.8 ms, and was removed; n_features 4 / 8 / 16 take accumulate_lanecol_kernel.)
template <int VEC, bool HAS_W>
__global__ void __launch_bounds__((2 + OWN_NA + OWN_CONS) * 32)
accumulate_owner_kernel(const __grid_constant__ CUtensorMap tm_x, const UpdParams7 p)
{
  constexpr int DS        = 32 * VEC;
  constexpr uint32_t ROWB = DS * 4;
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t raw  = ptx::smem_u32(smem_dyn);
  const uint32_t base = (raw + 127u) & ~127u;
  uint8_t* g          = smem_dyn + (base - raw);
  // X stages | label stages (labels | perm | wperm | class offsets) | table | wtab | class map | barriers
  const uint32_t x_bytes    = static_cast<uint32_t>(p.tr) * ROWB;
  const uint32_t lab_bytes  = static_cast<uint32_t>(p.tr) * 4u;
  const uint32_t perm_bytes = lab_bytes + 128u;     // read up to 3 entries past a segment
  const uint32_t l_bytes    = lab_bytes + 2u * perm_bytes + 128u;
  const uint32_t lbase      = base + p.nstage * x_bytes;
  const uint32_t tab_off    = p.nstage * x_bytes + p.nl * l_bytes;
  const uint32_t tab_u32    = base + tab_off;
  float* tab       = reinterpret_cast<float*>(g + tab_off);
  float* wtab      = tab + static_cast<size_t>(p.k) * DS;
  uint8_t* smap    = reinterpret_cast<uint8_t*>(wtab + ((p.k + 3) & ~3));
  uint64_t* bars   = reinterpret_cast<uint64_t*>(smap + ((p.k + 15) & ~15));
  const uint32_t bars_u32 = ptx::smem_u32(bars);
  const uint32_t wtab_u32 = ptx::smem_u32(wtab);
  const uint32_t smap_u32 = ptx::smem_u32(smap);
  const uint32_t B_FULLX = 0, B_EMPTYX = MAX_NSTAGE * 8, B_FULLL = 2 * MAX_NSTAGE * 8,
                 B_READYL = B_FULLL + OWN_MAXNL * 8, B_EMPTYL = B_READYL + OWN_MAXNL * 8;

  const int warp  = threadIdx.x / 32;
  const int lane  = threadIdx.x % 32;
  const int slice = blockIdx.y;
  const int cs    = slice * DS;
  const bool counts = (slice == 0);

  for (int i = threadIdx.x; i < p.k * DS; i += blockDim.x) tab[i] = 0.0f;
  for (int i = threadIdx.x; i < p.k; i += blockDim.x) wtab[i] = 0.0f;   // (int 0 == float 0)
  for (int i = threadIdx.x; i < p.k; i += blockDim.x) smap[i] = p.cls_map ? p.cls_map[i] : static_cast<uint8_t>(i & (OWN_CONS - 1));
  for (uint32_t i = threadIdx.x; i < p.nl * l_bytes / 4u; i += blockDim.x)   // list slots past a segment are read
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(lbase + i * 4u), "r"(0u) : "memory");   // (never used): keep in range
  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_NSTAGE; ++s) {
      ptx::mbar_init(bars_u32 + B_FULLX + s * 8, 1);
      ptx::mbar_init(bars_u32 + B_EMPTYX + s * 8, OWN_CONS);
    }
    for (int s = 0; s < OWN_MAXNL; ++s) {
      ptx::mbar_init(bars_u32 + B_FULLL + s * 8, 1);
      ptx::mbar_init(bars_u32 + B_READYL + s * 8, 1);
      ptx::mbar_init(bars_u32 + B_EMPTYL + s * 8, OWN_CONS);
    }
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&tm_x);
  }
  __syncthreads();

  const int64_t t_begin = static_cast<int64_t>(blockIdx.x) * p.tiles_per_block;
  const int64_t t_end   = min(p.tiles_total, t_begin + p.tiles_per_block);
  const bool prof       = (p.dbg_skip & 2) != 0;   // role-level cycle accounting (CUML_B200_UPD_SKIP=2)

  if (warp == 0) {
    // ---------------- X producer ----------------
    uint32_t s = 0, ph = 0;
    for (int64_t t = t_begin; t < t_end; ++t) {
      ptx::mbar_wait_park(bars_u32 + B_EMPTYX + s * 8, ph ^ 1u);
      if (ptx::elect_one()) {
        const uint32_t full = bars_u32 + B_FULLX + s * 8;
        ptx::mbar_arrive_expect_tx(full, x_bytes);
        ptx::tma_load_2d_hint(base + s * x_bytes, &tm_x, cs, static_cast<int32_t>(t * p.tr), full, ptx::kEvictFirst);
      }
      __syncwarp();
      if (++s == static_cast<uint32_t>(p.nstage)) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // ---------------- label producer ----------------
    uint32_t s = 0, ph = 0;
    for (int64_t t = t_begin; t < t_end; ++t) {
      ptx::mbar_wait_park(bars_u32 + B_EMPTYL + s * 8, ph ^ 1u);
      if (ptx::elect_one()) {
        const uint32_t full = bars_u32 + B_FULLL + s * 8;
        ptx::mbar_arrive_expect_tx(full, lab_bytes);
        bulk_load_1d(lbase + s * l_bytes, p.labels + t * p.tr, lab_bytes, full);
      }
      __syncwarp();
      if (++s == static_cast<uint32_t>(p.nl)) { s = 0; ph ^= 1u; }
    }
  } else if (warp < 2 + OWN_NA) {
    // ---------------- analysts ----------------
    const int aw      = warp - 2;
    const uint32_t lt = (1u << lane) - 1u;
    // lane j < 16 selects class j: bit b of a row's class must equal bit b of j
    const uint32_t nj0 = (lane & 1) ? 0u : 0xffffffffu, nj1 = (lane & 2) ? 0u : 0xffffffffu;
    const uint32_t nj2 = (lane & 4) ? 0u : 0xffffffffu, nj3 = (lane & 8) ? 0u : 0xffffffffu;
    const uint32_t lane_lt16 = lane < OWN_CONS ? 0xffffffffu : 0u;
    uint32_t s = 0, ph = 0;
    int turn = 0;
    long long c_wait = 0;
    const long long c_start = clock64();
    for (int64_t t = t_begin; t < t_end; ++t) {
      if (turn == aw) {
        const long long c0 = prof ? clock64() : 0;
        ptx::mbar_wait_park(bars_u32 + B_FULLL + s * 8, ph);
        if (prof) c_wait += clock64() - c0;
        const uint32_t ls   = lbase + s * l_bytes;
        const uint32_t perm = ls + lab_bytes;
        const uint32_t wprm = perm + perm_bytes;
        const uint32_t offs = wprm + perm_bytes;
        const int64_t row0  = t * p.tr;
        const int64_t left  = p.n - row0;
        const int valid     = left < p.tr ? static_cast<int>(left) : p.tr;
        int lbv[OWN_MAXCH], clv[OWN_MAXCH];
        float wv[OWN_MAXCH];
        uint32_t keep[OWN_MAXCH];   // lane j < 16: rows of class j in group c
#pragma unroll
        for (int c = 0; c < OWN_MAXCH; ++c) {
          lbv[c] = (c * 32 < p.tr) ? lds32(ls + (c * 32 + lane) * 4) : 0;
          if (HAS_W) wv[c] = (c * 32 + lane < valid) ? __ldg(p.w + row0 + c * 32 + lane) : 0.0f;
        }
#pragma unroll
        for (int c = 0; c < OWN_MAXCH; ++c) {
          clv[c] = 0;
          if (c * 32 < p.tr && c * 32 + lane < valid)   // labels past the end of the data are not labels
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(clv[c]) : "r"(smap_u32 + static_cast<uint32_t>(lbv[c])));
        }
        int tot = 0;
#pragma unroll
        for (int c = 0; c < OWN_MAXCH; ++c) {
          keep[c] = 0;
          if (c * 32 < p.tr) {
            // 4 ballots (one per class bit) + the valid mask give lane j < 16 the row mask of class j
            const bool ok     = (c * 32 + lane < valid) && !(p.dbg_skip & 1);
            const uint32_t vm = __ballot_sync(0xffffffffu, ok);
            const uint32_t b0 = __ballot_sync(0xffffffffu, (clv[c] & 1) != 0);
            const uint32_t b1 = __ballot_sync(0xffffffffu, (clv[c] & 2) != 0);
            const uint32_t b2 = __ballot_sync(0xffffffffu, (clv[c] & 4) != 0);
            const uint32_t b3 = __ballot_sync(0xffffffffu, (clv[c] & 8) != 0);
            keep[c] = vm & (b0 ^ nj0) & (b1 ^ nj1) & (b2 ^ nj2) & (b3 ^ nj3) & lane_lt16;
            tot += __popc(keep[c]);
            if (!HAS_W && counts && ok)
              asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(wtab_u32 + static_cast<uint32_t>(lbv[c]) * 4u), "r"(1) : "memory");
          }
        }
        int incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        int run = incl - tot;   // lane j: first slot of class j; lane 16: number of queued rows
        if (lane <= OWN_CONS) asm volatile("st.shared.b32 [%0], %1;" ::"r"(offs + lane * 4u), "r"(run) : "memory");
#pragma unroll
        for (int c = 0; c < OWN_MAXCH; ++c) {
          if (c * 32 < p.tr) {
            const bool ok      = (c * 32 + lane < valid) && !(p.dbg_skip & 1);
            const int cls      = clv[c] & (OWN_CONS - 1);
            const uint32_t mym = __shfl_sync(0xffffffffu, keep[c], cls);
            const int cbase    = __shfl_sync(0xffffffffu, run, cls);
            if (ok) {
              const uint32_t pos = static_cast<uint32_t>(cbase + __popc(mym & lt));
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(perm + pos * 4u),
                           "r"((static_cast<uint32_t>(c * 32 + lane) << 16) | static_cast<uint32_t>(lbv[c])) : "memory");
              if (HAS_W) asm volatile("st.shared.f32 [%0], %1;" ::"r"(wprm + pos * 4u), "f"(wv[c]) : "memory");
            }
            run += __popc(keep[c]);
          }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bars_u32 + B_READYL + s * 8);
      }
      if (++turn == OWN_NA) turn = 0;
      if (++s == static_cast<uint32_t>(p.nl)) { s = 0; ph ^= 1u; }
    }
    if (prof && blockIdx.x == 1 && blockIdx.y == 0 && lane == 0)
      printf("[upd v9 prof] analyst %d total %lld wait %lld\n", aw, clock64() - c_start, c_wait);
  } else {
    // ---------------- consumers ----------------
    const int cw          = warp - 2 - OWN_NA;
    const uint32_t coff   = static_cast<uint32_t>(lane) * (VEC * 4u);
    const uint32_t tab_me = tab_u32 + coff;
    uint32_t s = 0, ph = 0, sl = 0, phl = 0;
    long long c_wait = 0, c_rows = 0;
    const long long c_start = clock64();
    for (int64_t t = t_begin; t < t_end; ++t) {
      const long long c0 = prof ? clock64() : 0;
      ptx::mbar_wait_park(bars_u32 + B_READYL + sl * 8, phl);
      ptx::mbar_wait_park(bars_u32 + B_FULLX + s * 8, ph);
      if (prof) c_wait += clock64() - c0;
      const uint32_t xs   = base + s * x_bytes + coff;
      const uint32_t perm = lbase + sl * l_bytes + lab_bytes;
      const uint32_t wprm = perm + perm_bytes;
      const uint32_t offs = wprm + perm_bytes;
      const uint2 oo = lds64u(offs + (cw & ~1) * 4);
      int j          = (cw & 1) ? static_cast<int>(oo.y) : static_cast<int>(oo.x);
      const int o1   = (cw & 1) ? lds32(offs + cw * 4 + 4) : static_cast<int>(oo.y);
      c_rows += o1 - j;
    VecIO<VEC> x0, x1, tv;
    // apply one row: table row of label (e & 0xffff) += w * x
    auto apply = [&](uint32_t e, const VecIO<VEC>& x, float w, auto&& between) {
      const uint32_t l  = e & 0xffffu;
      const uint32_t ta = tab_me + l * ROWB;
      tv.load(ta);
      float cur = 0.0f;
      if (HAS_W && counts) cur = __int_as_float(lds32(wtab_u32 + l * 4u));
      between();
#pragma unroll
      for (int q = 0; q < VEC; ++q) tv.v[q] = HAS_W ? fmaf(x.v[q], w, tv.v[q]) : tv.v[q] + x.v[q];
      tv.store(ta);
      if (HAS_W && counts)   // every lane writes the same value to the same word: no divergence, no atomics
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(wtab_u32 + l * 4u), "f"(cur + w) : "memory");
    };
    if (j < o1 && (j & 1)) {   // odd head so that the pair loop reads aligned entry pairs
      const uint32_t e = static_cast<uint32_t>(lds32(perm + j * 4));
      float w = 1.0f;
      if (HAS_W) w = __int_as_float(lds32(wprm + j * 4));
      x0.load(xs + (e >> 16) * ROWB);
      apply(e, x0, w, [] {});
      ++j;
    }
    if (j < o1) {
      uint2 e = lds64u(perm + j * 4);
      uint2 w = make_uint2(0x3f800000u, 0x3f800000u);
      if (HAS_W) w = lds64u(wprm + j * 4);
      x0.load(xs + (e.x >> 16) * ROWB);
      for (; j < o1; j += 2) {
        const uint2 en = lds64u(perm + j * 4 + 8);     // entries j+2, j+3 (past the end: stale, in range)
        uint2 wn       = w;
        if (HAS_W) wn = lds64u(wprm + j * 4 + 8);
        apply(e.x, x0, __uint_as_float(w.x), [&] { x1.load(xs + (e.y >> 16) * ROWB); });
        if (j + 1 < o1) apply(e.y, x1, __uint_as_float(w.y), [&] { x0.load(xs + (en.x >> 16) * ROWB); });
        e = en;
        w = wn;
      }
    }
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(bars_u32 + B_EMPTYX + s * 8);
        ptx::mbar_arrive(bars_u32 + B_EMPTYL + sl * 8);
      }
      if (++s == static_cast<uint32_t>(p.nstage)) { s = 0; ph ^= 1u; }
      if (++sl == static_cast<uint32_t>(p.nl)) { sl = 0; phl ^= 1u; }
    }
    if (prof && blockIdx.x == 1 && blockIdx.y == 0 && lane == 0)
      printf("[upd v9 prof] consumer %d tiles %lld rows %lld total %lld wait %lld\n", cw,
             static_cast<long long>(t_end - t_begin), c_rows, clock64() - c_start, c_wait);
  }
  __syncthreads();
  float* outS     = p.partial_S + static_cast<size_t>(blockIdx.x) * p.k * p.d;
  const int wcols = min(DS, p.d - cs);
  for (int i = threadIdx.x; i < p.k * DS; i += blockDim.x) {
    const int j = i / DS, c = i % DS;
    if (c < wcols) outS[static_cast<size_t>(j) * p.d + cs + c] = tab[i];
  }
  if (slice == 0) {
    float* outW = p.partial_W + static_cast<size_t>(blockIdx.x) * p.k;
    for (int i = threadIdx.x; i < p.k; i += blockDim.x)
      outW[i] = HAS_W ? wtab[i] : static_cast<float>(reinterpret_cast<const int*>(wtab)[i]);
  }
}

// packed[e] (+)= sum_b partial[b][e]: fixed order, fp64.  32 elements x 8 partial-lanes per block.
__global__ void __launch_bounds__(256)
reduce_partials_f32_kernel(const float* __restrict__ partial_S, const float* __restrict__ partial_W, int row_blocks,
                           int k, int d, double* __restrict__ packed, int accumulate_into)
{
  __shared__ double red[8][33];
  const int64_t kd    = static_cast<int64_t>(k) * d;
  const int64_t total = kd + k;
  const int ex        = threadIdx.x % 32;
  const int pl        = threadIdx.x / 32;
  const int64_t e     = static_cast<int64_t>(blockIdx.x) * 32 + ex;
  double s = 0.0;
  if (e < total) {
    const float* src     = (e < kd) ? (partial_S + e) : (partial_W + (e - kd));
    const int64_t stride = (e < kd) ? kd : k;
    for (int b = pl; b < row_blocks; b += 8) s += static_cast<double>(src[static_cast<int64_t>(b) * stride]);
  }
  red[pl][ex] = s;
  __syncthreads();
  if (pl == 0 && e < total) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][ex];
    packed[e] = accumulate_into ? packed[e] + t : t;
  }
}

// ---- M-step for short rows (n_features 4 / 8 / 16, e.g. C5): lane = column, warp-private tables -------------------
// Default for these shapes (CUML_B200_UPD_LANECOL=0 restores the lane = row kernel): 3.65 -> 2.06 ms at C5 = 6.6 TB/s.
// The lane = row kernel above is bound by shared-memory wavefronts (32 random table rows per instruction).  Here one
// warp instruction covers R = 32 / D consecutive rows with lane = (sub-row r, column c), lane == r * D + c, so a batch
// of 32 * U consecutive floats of X is loaded with U perfectly coalesced 128-byte loads and needs no staging at all.
// Every warp owns a private table of k x 32 floats: row j holds R sub-tables side by side (sub-row r adds into columns
// [r * D, r * D + D)), i.e. lane l only ever touches bank l -- no bank conflicts, and rows of one instruction that
// share a label never touch the same cell, so there is nothing to match, rank or sort.  Per R rows: 1 shuffle (label),
// 1 address, LDS + FADD + STS.  Memory parallelism: two batches of U loads per lane in flight (register double buffer)
// in each of up to 16 warps.  Cluster weights: integer shared-memory counts per warp (exact) or, weighted, one more
// private [R][k] table updated by the c == 0 lanes.  The CTA folds its warps' tables in a fixed order into the same
// partials format as the other kernels (deterministic).
struct LaneColParams {
  int64_t n;
  int64_t rows_per_block;   // multiple of the batch (U * R rows)
  int k, warps;
  const float* X;
  const int32_t* labels;
  const float* w;
  float* partial_S;
  float* partial_W;
};

constexpr int LC_U = 32;   // loads per lane and batch (batch = 32 * LC_U floats = 4 KB of X)

template <int D, bool HAS_W>
__global__ void __launch_bounds__(512, 1)
accumulate_lanecol_kernel(const LaneColParams p)
{
  constexpr int R  = 32 / D;          // rows per warp instruction
  constexpr int BR = LC_U * R;        // rows per batch
  constexpr int NQ = BR / 32;         // label (weight) registers per lane and batch (== R)
  extern __shared__ float lc_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r    = lane / D;
  const size_t per_warp = static_cast<size_t>(p.k) * 32 + static_cast<size_t>(R) * p.k;   // floats
  float* tbl = lc_smem + per_warp * warp;          // [k][32]
  float* aux = tbl + static_cast<size_t>(p.k) * 32;   // HAS_W: [R][k] weights; else [k] int counts (rest unused)
  for (size_t i = threadIdx.x; i < per_warp * p.warps; i += blockDim.x) lc_smem[i] = 0.0f;
  __syncthreads();

  const int64_t row_begin = static_cast<int64_t>(blockIdx.x) * p.rows_per_block;
  const int64_t row_end   = min(p.n, row_begin + p.rows_per_block);
  const int64_t n_batches = row_end > row_begin ? (row_end - row_begin + BR - 1) / BR : 0;

  float xa[LC_U], xb[LC_U];
  int la[NQ], lb[NQ];
  float wa[NQ], wb[NQ];

  // batch b of this CTA: rows [row_begin + b * BR, +BR) clipped to row_end; rows past the end read as label -1, x = w = 0
  auto load = [&](float (&x)[LC_U], int (&lab)[NQ], float (&wv)[NQ], int64_t b) {
    const int64_t row0 = row_begin + b * BR;
    const float* src   = p.X + row0 * D;
    if (row0 + BR <= row_end) {
#pragma unroll
      for (int u = 0; u < LC_U; ++u) x[u] = __ldcs(src + u * 32 + lane);
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        lab[q] = __ldcs(p.labels + row0 + q * 32 + lane);
        if (HAS_W) wv[q] = __ldg(p.w + row0 + q * 32 + lane);
      }
    } else {
      const int64_t left = (row_end - row0) * D;   // floats of this batch that exist (<= 0: none)
#pragma unroll
      for (int u = 0; u < LC_U; ++u) x[u] = (u * 32 + lane < left) ? __ldcs(src + u * 32 + lane) : 0.0f;
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const bool ok = row0 + q * 32 + lane < row_end;
        lab[q] = ok ? __ldcs(p.labels + row0 + q * 32 + lane) : -1;
        if (HAS_W) wv[q] = ok ? __ldg(p.w + row0 + q * 32 + lane) : 0.0f;
      }
    }
  };
  const uint32_t cell0 = ptx::smem_u32(tbl) + static_cast<uint32_t>(lane) * 4u;   // this lane's column of table row 0
  auto apply = [&](const float (&x)[LC_U], const int (&lab)[NQ], const float (&wv)[NQ]) {
    int byte_off[NQ];   // label -> byte offset of its table row (rows past the end: row 0, they add zeros)
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      if (!HAS_W && lab[q] >= 0) atomicAdd(reinterpret_cast<int*>(aux) + lab[q], 1);   // warp-private exact counts
      byte_off[q] = max(lab[q], 0) * 128;
    }
#pragma unroll
    for (int u = 0; u < LC_U; ++u) {
      // step u covers batch rows u * R + r; their labels sit in register (u * R) / 32 of lane (u * R + r) % 32
      const int q         = (u * R) >> 5;
      const int src       = ((u * R) & 31) + r;
      const uint32_t cell = cell0 + static_cast<uint32_t>(__shfl_sync(0xffffffffu, byte_off[q], src));
      if (HAS_W) {
        const float wr = __shfl_sync(0xffffffffu, wv[q], src);
        sts32f(cell, fmaf(wr, x[u], lds32f(cell)));
        if (lane % D == 0) {
          float* wc = aux + r * p.k + (static_cast<uint32_t>(cell - cell0) >> 7);
          *wc += wr;
        }
      } else {
        sts32f(cell, lds32f(cell) + x[u]);
      }
    }
  };

  // warp w takes batches w, w + warps, ... of its CTA.  Every warp runs the same number of rounds (a batch past the
  // end loads nothing and adds zeros), so the loop is CTA-uniform and the shuffles need no re-convergence; the next
  // batch's loads are in flight while the current one is applied.
  const int64_t rounds = (n_batches + p.warps - 1) / p.warps;
  if (rounds > 0) load(xa, la, wa, warp);
  for (int64_t it = 0; it < rounds; it += 2) {
    if (it + 1 < rounds) load(xb, lb, wb, (it + 1) * p.warps + warp);
    apply(xa, la, wa);
    if (it + 1 < rounds) {
      if (it + 2 < rounds) load(xa, la, wa, (it + 2) * p.warps + warp);
      apply(xb, lb, wb);
    }
  }
  __syncthreads();

  // fold the warps' tables (fixed order: warp, then sub-row) into this CTA's partials
  float* outS = p.partial_S + static_cast<size_t>(blockIdx.x) * p.k * D;
  for (int e = threadIdx.x; e < p.k * D; e += blockDim.x) {
    const int j = e / D, c = e % D;
    float sum = 0.0f;
    for (int w = 0; w < p.warps; ++w) {
      const float* t = lc_smem + per_warp * w + j * 32 + c;
#pragma unroll
      for (int rr = 0; rr < R; ++rr) sum += t[rr * D];
    }
    outS[e] = sum;
  }
  float* outW = p.partial_W + static_cast<size_t>(blockIdx.x) * p.k;
  for (int j = threadIdx.x; j < p.k; j += blockDim.x) {
    float sum = 0.0f;
    int cnt   = 0;
    for (int w = 0; w < p.warps; ++w) {
      const float* a = lc_smem + per_warp * w + static_cast<size_t>(p.k) * 32;
      if (HAS_W) {
#pragma unroll
        for (int rr = 0; rr < R; ++rr) sum += a[rr * p.k + j];
      } else {
        cnt += reinterpret_cast<const int*>(a)[j];
      }
    }
    outW[j] = HAS_W ? sum : static_cast<float>(cnt);
  }
}

// warps whose private tables fit shared memory (0: the kernel does not apply)
static int lanecol_warps(const Handle& h, int d, int k)
{
  static const bool on = env_flag("CUML_B200_UPD_LANECOL", true);
  if (!on || (d != 4 && d != 8 && d != 16)) return 0;
  const size_t per_warp = (static_cast<size_t>(k) * 32 + static_cast<size_t>(32 / d) * k) * sizeof(float);
  const int warps       = static_cast<int>(std::min<size_t>(16, (h.smem_optin - 1024) / per_warp));
  return warps >= 4 ? warps : 0;
}

}  // namespace

struct TmaUpdatePlan {
  int ds = 0, tr = 0, warps = 0, slices = 0, ctas_per_sm = 0, nb = 1, nstage = 3, na = 2;
  size_t smem = 0;
  uint32_t sub_bytes = 0;
};

// Choose (sub-slices per CTA, CTAs per SM, ring depth, tile rows) to maximise the bytes the TMA ring keeps
// in flight per SM (the measured limiter: ring turn-around is ~2-3 us) with >= 8 consumer warps per SM
// when the table allows it.
static int tma_analysts()
{
  const char* e = std::getenv("CUML_B200_ANALYSTS");
  int na        = e ? std::atoi(e) : 2;
  return std::max(1, std::min(4, na));
}

static TmaUpdatePlan plan_tma_update(const Handle& h, int d, int k)
{
  TmaUpdatePlan best;
  double best_score = -1.0;
  if (d % 4 != 0) return best;
  const size_t sm_total = 228 * 1024;
  const int sub_cols    = d >= 32 ? 32 : d;
  const int cons_per_sub = std::max(1, sub_cols / 8);
  for (int nb = 1; nb <= ((d >= 64) ? 2 : 1); ++nb) {
    const int ds       = sub_cols * nb;
    const int na       = tma_analysts();
    const size_t table = (static_cast<size_t>(k) * ds + static_cast<size_t>(na) * ((k + 3) & ~3)) * 4;
    for (int per_sm = 1; per_sm <= 8; ++per_sm) {
      const size_t budget = std::min<size_t>(h.smem_optin, sm_total / per_sm - 1024);
      const size_t fixed  = table + 1024 + 3 * MAX_NSTAGE * 8 + 256;
      if (fixed + 2 * 5120 > budget) continue;
      for (int nstage = 2; nstage <= MAX_NSTAGE; ++nstage) {
        const size_t stage_budget = (budget - fixed) / nstage;
        if (stage_budget < 4096 + 1024) break;
        // rows per tile: tile = nb sub-tiles of tr x sub_cols floats + labels/meta (8 B per row, 1 KB granules)
        int tr = static_cast<int>((stage_budget - 1024) / (static_cast<size_t>(ds) * 4 + 8));
        tr     = std::min(tr, 256);
        const int gran = std::max(32, 1024 / (sub_cols * 4));
        tr -= tr % gran;
        if (tr < gran) continue;
        const size_t tile_bytes = static_cast<size_t>(tr) * ds * 4;
        if (tile_bytes < 4096) continue;
        const int cons_sm       = per_sm * nb * cons_per_sub;
        const double inflight   = static_cast<double>(per_sm) * nstage * tile_bytes;       // bytes per SM
        // consumers below 8 per SM are the limiter; beyond ~12 extra warps do not help
        // the ring alone reaches ~6.6 TB/s with >= 64 KB in flight per SM (measured with the consumers
        // disabled); the kernel itself is bound by shared-memory wavefronts at ~3.5 TB/s whatever the split,
        // so the score only has to avoid starving either side (A/B on C3 and C5: profiles/README.md)
        static const double cap_kb = std::getenv("CUML_B200_UPD_INFLIGHT_KB") ? std::atof(std::getenv("CUML_B200_UPD_INFLIGHT_KB")) : 96.0;
        static const int cons_cap  = std::getenv("CUML_B200_UPD_CONS_CAP") ? std::atoi(std::getenv("CUML_B200_UPD_CONS_CAP")) : 16;
        const double score = std::min(inflight, cap_kb * 1024) * std::min(cons_sm, cons_cap) / cons_cap +
                             (tile_bytes >= 8192 ? 1.0 : 0.0) + 1e-6 * inflight;
        if (score > best_score) {
          best_score       = score;
          best.ds          = ds;
          best.nb          = nb;
          best.tr          = tr;
          best.nstage      = nstage;
          best.na          = na;
          best.warps       = nb * cons_per_sub;
          best.slices      = static_cast<int>(ceil_div(d, ds));
          best.ctas_per_sm = per_sm;
          best.sub_bytes   = static_cast<uint32_t>(tr) * sub_cols * 4;
          const uint32_t lab = (static_cast<uint32_t>(tr) * 8 + 1023u) & ~1023u;
          best.smem = static_cast<size_t>(nstage) * (static_cast<size_t>(nb) * best.sub_bytes + lab) + table +
                      3 * MAX_NSTAGE * 8 + 1024 + 64;
        }
      }
    }
  }
  return best;
}

struct OwnerPlan {
  int vec = 0, tr = 0, nstage = 0, slices = 0, ncons = 16, nl = 0;
  size_t smem = 0;
};

// label-class kernel: needs n_features >= 32 and the [k x 32*VEC] table plus >= 64 KB of ring in one SM
static OwnerPlan plan_owner_update(const Handle& h, int d, int k)
{
  OwnerPlan best;
  int mode = 1;
  if (const char* e = std::getenv("CUML_B200_UPDATE_OWNER")) mode = std::atoi(e);
  if (mode == 0 || d < 32 || d % 4 != 0 || k < 16) return best;
  int force_vec = 0, force_tr = 0, force_nl = 0, force_ns = 0;
  if (const char* e = std::getenv("CUML_B200_OWNER_VEC")) force_vec = std::atoi(e);
  if (const char* e = std::getenv("CUML_B200_OWNER_TR")) force_tr = std::atoi(e);
  if (const char* e = std::getenv("CUML_B200_OWNER_NL")) force_nl = std::atoi(e);
  if (const char* e = std::getenv("CUML_B200_OWNER_NS")) force_ns = std::atoi(e);
  const size_t budget = h.smem_optin - 512;
  double best_score   = -1.0;
  for (int vec = 4; vec >= 1; vec >>= 1) {
    if (force_vec && vec != force_vec) continue;
    const int ds = 32 * vec;
    if (vec > 1 && ds > d) continue;   // do not read padding columns
    const size_t table = (static_cast<size_t>(k) * ds + ((k + 3) & ~3)) * 4 + ((k + 15) & ~15) + (2 * MAX_NSTAGE + 3 * OWN_MAXNL) * 8 + 256;
    if (table + 2 * 8192 > budget) continue;
    const int slices = static_cast<int>(ceil_div(d, ds));
    for (int tr = 64; tr <= 256; tr += 32) {
      if (force_tr && tr != force_tr) continue;
      const size_t stage = static_cast<size_t>(tr) * ds * 4;
      for (int nl = OWN_MAXNL; nl >= 4; --nl) {
        if (force_nl && nl != force_nl) continue;
        const size_t lists = nl * (static_cast<size_t>(tr) * 12 + 384);   // label ring
        if (table + lists + 2 * stage > budget) continue;
        int nstage = static_cast<int>(std::min<size_t>(MAX_NSTAGE, (budget - table - lists) / stage));
        nstage     = std::min(nstage, nl - 1);
        if (force_ns) nstage = std::min(nstage, force_ns);
        if (nstage < 2) continue;
        const double inflight = static_cast<double>(nstage) * tr * ds * 4;
        // measured on C3 / C2 (profiles/README.md): large tiles amortise the consumers' per-tile barrier
        // round trips; two 64 KB stages beat four 32 KB ones once the analysts run ahead on the label ring
        const double score = std::min(inflight, 64.0 * 1024) * (1.0 + 0.25 * vec) + 96.0 * tr + 16.0 * nl +
                             (d % ds == 0 ? 4096.0 : 0.0);
        if (score > best_score) {
          best_score  = score;
          best.vec    = vec;
          best.tr     = tr;
          best.nstage = nstage;
          best.slices = slices;
          best.smem   = table + lists + nstage * stage + 128;
          best.nl     = nl;
        }
      }
    }
  }
  best.ncons = OWN_CONS;
  return best;
}

bool tma_update_supported(const Handle& h, int d, int k)
{
  if (h.cc_major < 9 || d % 4 != 0) return false;
  if (d < 32 && (d & (d - 1)) != 0) return false;  // short rows: power-of-two widths only
  return plan_tma_update(h, d, k).ds > 0;
}

// sums + weights of one partition into packed[0 .. k*d+k) (the inertia cell is left untouched)
// label -> consumer-class map from the cluster weights of the previous iteration (work balance only: the sums do
// not depend on it).  Returns null when the label-class kernel is not the one planned for (d, k).
const uint8_t* tma_update_balance(Handle& h, const double* W, int d, int k, DevBuf<uint8_t>& map)
{
  if (plan_owner_update(h, d, k).vec == 0) return nullptr;
  if (map.n < static_cast<size_t>(k)) map.alloc(k, h.stream);
  balance_classes_kernel<<<1, 1024, static_cast<size_t>(k) * sizeof(float), h.stream>>>(W, k, map.get());
  CB2_CHECK_LAUNCH();
  return map.get();
}

void tma_update_reduce(Handle& h, const float* partial_S, const float* partial_W, int row_blocks, int k, int d,
                       double* packed, bool accumulate_into)
{
  const int64_t total = static_cast<int64_t>(k) * d + k;
  reduce_partials_f32_kernel<<<static_cast<unsigned>(ceil_div(total, 32)), 256, 0, h.stream>>>(
    partial_S, partial_W, row_blocks, k, d, packed, accumulate_into ? 1 : 0);
  CB2_CHECK_LAUNCH();
}

void tma_update_accumulate(Handle& h, const float* X, int64_t n, int d, const int32_t* labels_padded, const float* w,
                           int k, DevBuf<float>& partial_S, DevBuf<float>& partial_W, double* packed,
                           bool accumulate_into, const uint8_t* cls_map)
{
  const int64_t total = static_cast<int64_t>(k) * d + k;
  if (n == 0) {
    if (!accumulate_into) CB2_CUDA(cudaMemsetAsync(packed, 0, total * sizeof(double), h.stream));
    return;
  }
  if (const int lc_warps = lanecol_warps(h, d, k); lc_warps > 0) {
    // lane = column kernel for short rows (see accumulate_lanecol_kernel)
    const int rows_batch = LC_U * (32 / d);
    const int64_t batches = ceil_div(n, rows_batch);
    int64_t rb = std::min<int64_t>(h.sm_count, ceil_div(batches, lc_warps));
    rb         = std::min<int64_t>(rb, std::max<int64_t>(1, n / (16 * static_cast<int64_t>(k)) + 1));
    LaneColParams q{};
    q.n = n; q.k = k; q.warps = lc_warps;
    q.rows_per_block = ceil_div(batches, rb) * rows_batch;
    rb               = ceil_div(n, q.rows_per_block);
    if (partial_S.n < static_cast<size_t>(rb) * k * d) partial_S.alloc(static_cast<size_t>(rb) * k * d, h.stream);
    if (partial_W.n < static_cast<size_t>(rb) * k) partial_W.alloc(static_cast<size_t>(rb) * k, h.stream);
    q.X = X; q.labels = labels_padded; q.w = w; q.partial_S = partial_S.get(); q.partial_W = partial_W.get();
    const size_t smem = (static_cast<size_t>(k) * 32 + static_cast<size_t>(32 / d) * k) * sizeof(float) * lc_warps;
    if (std::getenv("CUML_B200_UPD_PLAN"))
      std::printf("[cuml_b200 update plan lanecol] d %d warps %d row blocks %lld rows/block %lld smem %zu\n", d, lc_warps,
                  static_cast<long long>(rb), static_cast<long long>(q.rows_per_block), smem);
    auto launch = [&](auto kern) {
      CB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(h.smem_optin)));
      kern<<<static_cast<unsigned>(rb), lc_warps * 32, smem, h.stream>>>(q);
    };
    const bool hw = w != nullptr;
    switch (d) {
      case 4: hw ? launch(accumulate_lanecol_kernel<4, true>) : launch(accumulate_lanecol_kernel<4, false>); break;
      case 8: hw ? launch(accumulate_lanecol_kernel<8, true>) : launch(accumulate_lanecol_kernel<8, false>); break;
      default: hw ? launch(accumulate_lanecol_kernel<16, true>) : launch(accumulate_lanecol_kernel<16, false>); break;
    }
    CB2_CHECK_LAUNCH();
    reduce_partials_f32_kernel<<<static_cast<unsigned>(ceil_div(total, 32)), 256, 0, h.stream>>>(
      partial_S.get(), partial_W.get(), static_cast<int>(rb), k, d, packed, accumulate_into ? 1 : 0);
    CB2_CHECK_LAUNCH();
    return;
  }
  if (const OwnerPlan op = plan_owner_update(h, d, k); op.vec > 0) {
    UpdParams7 q{};
    q.n = n; q.d = d; q.k = k; q.tr = op.tr; q.nstage = op.nstage; q.ncons = op.ncons; q.nl = op.nl;
    q.cls_map = cls_map;
    {
      const char* e = std::getenv("CUML_B200_UPD_SKIP");
      q.dbg_skip    = e ? std::atoi(e) : 0;
      if (std::getenv("CUML_B200_UPD_PLAN"))
        std::printf("[cuml_b200 update plan owner] vec %d tr %d nstage %d nl %d slices %d smem %zu\n", op.vec, op.tr,
                    op.nstage, op.nl, op.slices, op.smem);
    }
    q.tiles_total = ceil_div(n, op.tr);
    int64_t rb = std::max<int64_t>(1, static_cast<int64_t>(h.sm_count) / op.slices);
    rb         = std::min<int64_t>(rb, q.tiles_total);
    rb         = std::min<int64_t>(rb, std::max<int64_t>(1, n / (16 * static_cast<int64_t>(k)) + 1));
    q.tiles_per_block = ceil_div(q.tiles_total, rb);
    rb                = ceil_div(q.tiles_total, q.tiles_per_block);
    if (partial_S.n < static_cast<size_t>(rb) * k * d) partial_S.alloc(static_cast<size_t>(rb) * k * d, h.stream);
    if (partial_W.n < static_cast<size_t>(rb) * k) partial_W.alloc(static_cast<size_t>(rb) * k, h.stream);
    q.labels = labels_padded; q.w = w; q.partial_S = partial_S.get(); q.partial_W = partial_W.get();
    CUtensorMap tm = make_map_2d(X, static_cast<uint64_t>(d), static_cast<uint64_t>(n),
                                 static_cast<uint64_t>(d) * sizeof(float), static_cast<uint32_t>(32 * op.vec),
                                 static_cast<uint32_t>(op.tr), CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
    dim3 grid(static_cast<unsigned>(rb), static_cast<unsigned>(op.slices));
    const unsigned threads = (2 + OWN_NA + OWN_CONS) * 32;
    auto launch = [&](auto kern) {
      CB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(h.smem_optin)));
      kern<<<grid, threads, op.smem, h.stream>>>(tm, q);
    };
    const bool hw = w != nullptr;
    switch (op.vec) {
      case 4: hw ? launch(accumulate_owner_kernel<4, true>) : launch(accumulate_owner_kernel<4, false>); break;
      case 2: hw ? launch(accumulate_owner_kernel<2, true>) : launch(accumulate_owner_kernel<2, false>); break;
      default: hw ? launch(accumulate_owner_kernel<1, true>) : launch(accumulate_owner_kernel<1, false>); break;
    }
    CB2_CHECK_LAUNCH();
    reduce_partials_f32_kernel<<<static_cast<unsigned>(ceil_div(total, 32)), 256, 0, h.stream>>>(
      partial_S.get(), partial_W.get(), static_cast<int>(rb), k, d, packed, accumulate_into ? 1 : 0);
    CB2_CHECK_LAUNCH();
    return;
  }
  TmaUpdatePlan pl = plan_tma_update(h, d, k);
  CB2_EXPECTS(pl.ds > 0, "no shared-memory plan for the TMA centroid update");
  if (std::getenv("CUML_B200_UPD_PLAN"))
    std::printf("[cuml_b200 update plan v5] ds %d nb %d tr %d nstage %d slices %d ctas/sm %d warps %d smem %zu\n", pl.ds, pl.nb,
                pl.tr, pl.nstage, pl.slices, pl.ctas_per_sm, pl.warps, pl.smem);
  UpdParams p{};
  p.n           = n;
  p.d           = d;
  p.k           = k;
  p.ds          = pl.ds;
  p.tr          = pl.tr;
  p.nb          = pl.nb;
  p.nstage      = pl.nstage;
  p.na          = pl.na;
  p.sub_bytes   = pl.sub_bytes;
  p.tiles_total = ceil_div(n, pl.tr);
  int64_t row_blocks = std::max<int64_t>(1, static_cast<int64_t>(h.sm_count) * pl.ctas_per_sm / pl.slices);
  row_blocks         = std::min<int64_t>(row_blocks, p.tiles_total);
  // bound the partials traffic (row_blocks * k * d * 8 bytes) to ~1/8 of the X traffic
  row_blocks = std::min<int64_t>(row_blocks, std::max<int64_t>(1, n / (16 * static_cast<int64_t>(k)) + 1));
  p.tiles_per_block = ceil_div(p.tiles_total, row_blocks);
  row_blocks        = ceil_div(p.tiles_total, p.tiles_per_block);
  if (partial_S.n < static_cast<size_t>(row_blocks) * k * d) partial_S.alloc(static_cast<size_t>(row_blocks) * k * d, h.stream);
  if (partial_W.n < static_cast<size_t>(row_blocks) * k) partial_W.alloc(static_cast<size_t>(row_blocks) * k, h.stream);
  p.labels    = labels_padded;
  p.w         = w;
  p.partial_S = partial_S.get();
  p.partial_W = partial_W.get();

  const int sub_cols = pl.ds / pl.nb;
  const CUtensorMapSwizzle swz = sub_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : sub_cols == 16 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : sub_cols == 8  ? CU_TENSOR_MAP_SWIZZLE_32B
                                                 : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUtensorMap tm = make_map_2d(X, static_cast<uint64_t>(d), static_cast<uint64_t>(n),
                               static_cast<uint64_t>(d) * sizeof(float), static_cast<uint32_t>(sub_cols),
                               static_cast<uint32_t>(pl.tr), swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  dim3 grid(static_cast<unsigned>(row_blocks), static_cast<unsigned>(pl.slices));
  const unsigned threads = (pl.warps + 1 + pl.na) * 32;
  auto launch = [&](auto kern) {
    CB2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(h.smem_optin)));
    kern<<<grid, threads, pl.smem, h.stream>>>(tm, p);
  };
  const bool hw = w != nullptr;
  switch (sub_cols) {
    case 32: hw ? launch(accumulate_tma_kernel<8, true>) : launch(accumulate_tma_kernel<8, false>); break;
    case 16: hw ? launch(accumulate_tma_kernel<4, true>) : launch(accumulate_tma_kernel<4, false>); break;
    case 8: hw ? launch(accumulate_tma_kernel<2, true>) : launch(accumulate_tma_kernel<2, false>); break;
    default: hw ? launch(accumulate_tma_kernel<1, true>) : launch(accumulate_tma_kernel<1, false>); break;
  }
  CB2_CHECK_LAUNCH();
  reduce_partials_f32_kernel<<<static_cast<unsigned>(ceil_div(total, 32)), 256, 0, h.stream>>>(
    partial_S.get(), partial_W.get(), static_cast<int>(row_blocks), k, d, packed, accumulate_into ? 1 : 0);
  CB2_CHECK_LAUNCH();
}

}  // namespace cb2
