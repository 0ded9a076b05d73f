// Fused pairwise-L2 + argmin for sm_100a (the fusedL2NN role of the reference's cuVS backend,
// reached from cpp/src/kmeans/kmeans_fit.cu:58-59,153-154 and kmeans_predict.cu:41-42).
//
//   label_i = argmin_j ( 1/2 ||c_j||^2 - x_i . c_j )        (first minimum on ties)
//
// The x.c contraction runs on the 5th-gen tensor cores (tcgen05.mma, fp32 accumulators in TMEM) in split
// precision, hi = x rounded / truncated to tf32, lo = x - hi (exact in fp32):
//   tf32 + 2 bf16   x_hi.c_hi as kind::tf32, the two correction terms as kind::f16 with bf16 operands (half
//                   the MMA instructions per term); -1/2||c||^2 enters the accumulator through one extra
//                   K = 8 MMA of a ones tile with three tf32-exact pieces               (default everywhere)
//   3xTF32          x.c ~= x_lo.c_hi + x_hi.c_lo + x_hi.c_hi, all kind::tf32     (CUML_B200_BF16C=0, transform)
// Both give fp32-grade labels (tests/test_kmeans_gpu.py::test_tensor_core_dot_accuracy, label-gap tests).
//
// Persistent warp-specialised kernels, one CTA per SM:
//   fused_l2_argmin_2cta_kernel  k > 128: two CTAs of a TPC share M = 256, N = 256 MMAs (cta_group::2), each
//                                holds half of every centroid block; 8 converter warps, 16 epilogue warps
//   fused_l2_argmin_solo_kernel  k <= 128: the same roles in one CTA, row-owner epilogue for one centroid tile
//   fused_l2_argmin_tsp_kernel   n_features <= 32 with few clusters: X operand in tensor memory (converter writes TMEM
//                                columns, A-from-TMEM MMAs), optionally with the M-step fused in (one pass over X per
//                                Lloyd step)
//   fused_l2_argmin_kernel       round 1's 3xTF32 kernel: winning-value epilogue of the seeding rounds, 3xTF32 row-packed
// Roles: TMA producers for X and centroid K-blocks (128B / 64B / 32B-swizzled tiles), converter warps that
// split the raw X tile in shared memory, one MMA-issuing warp (uniform control flow, elect.sync lane), and the
// argmin epilogue (tcgen05.ld, thread = row, four (min, argmin) chains, parts merged through shared memory).
// Pipelines are mbarrier rings: X raw -> ready -> empty, centroids full/empty, and two or more TMEM
// accumulators (full/empty) so the argmin of one tile overlaps the MMAs of the next.
#include <cuda_bf16.h>

#include <cstdlib>

#include "kernels.cuh"
#include "ptx.cuh"
#include "tensormap.cuh"

namespace cb2 {

namespace {

// timed mbarrier wait: accumulates the cycles spent waiting (role-level profiling, CTA 0 only)
#define CB2_TIMED_WAIT(bar, parity, acc)            \
  do {                                              \
    if (p.dbg_clk) {                                \
      const long long t__ = clock64();              \
      ptx::mbar_wait_park((bar), (parity));         \
      (acc) += clock64() - t__;                     \
    } else {                                        \
      ptx::mbar_wait_park((bar), (parity));         \
    }                                               \
  } while (0)

// the same for kernels that take the profiling as a template parameter CLK (the default instantiations carry no
// trace of it)
#define CB2_WAIT_CLK(bar, parity, acc)              \
  do {                                              \
    if constexpr (CLK) {                            \
      const long long t__ = clock64();              \
      ptx::mbar_wait_park((bar), (parity));         \
      (acc) += clock64() - t__;                     \
    } else {                                        \
      ptx::mbar_wait_park((bar), (parity));         \
    }                                               \
  } while (0)

constexpr int TILE_M       = 128;
constexpr int KBLOCK       = 32;              // fp32 elements per 128-byte swizzle row
constexpr int KBLOCK_BYTES = TILE_M * 128;    // one K-block of an A tile: 16 KB
constexpr int NUM_THREADS  = 768;             // 24 warps: producers, MMA issuer, 4 converter, 16 epilogue
constexpr int PAIR_THREADS = 896;             // CTA-pair kernel: 4 more converter warps (24..27)
constexpr int EPI_THREADS  = 512;             // 16 epilogue warps: 4 TMEM lane quarters x 4 column parts
constexpr int MAX_STAGES   = 4;
constexpr int MAX_A_SLOTS  = 8;
constexpr int MAX_ACC      = 8;
constexpr int MAX_RAW      = 12;  // raw X K-block slots (16 KB each) of the A-in-TMEM variant   // TMEM accumulator stages (512 columns / BN, at most 8)
constexpr int A_SLOT_BYTES = 2 * KBLOCK_BYTES;  // hi then lo

// (slot, phase) of an mbarrier ring, advanced without integer division (a runtime '%' or '/' costs ~100
// dependent cycles on the GPU, and every role used to pay several per tile)
struct Ring {
  uint32_t slot = 0, phase = 0;
  __device__ __forceinline__ void advance(uint32_t n)
  {
    if (++slot == n) {
      slot = 0;
      phase ^= 1u;
    }
  }
  __device__ __forceinline__ void advance_by(uint32_t steps, uint32_t n)
  {
    slot += steps;
    while (slot >= n) {
      slot -= n;
      phase ^= 1u;
    }
  }
};

struct FusedParams {
  int64_t n;
  int64_t m_tiles;
  int k_tiles;    // k_pad / bn
  int d;          // true feature count (K steps past it are skipped)
  int kb;         // d_pad / 32
  int bn;         // centroids per accumulator tile (multiple of 32, <= 256)
  int a_slots;    // ring of X K-block slots (hi 16 KB + lo 16 KB each); >= kb when k_tiles > 1
  int b_stages;   // 2..4
  int b_resident; // all k_tiles*kb centroid blocks fit the B stages: load once, never release
  int n_acc;      // TMEM accumulator stages: min(MAX_ACC, 512 / bn)
  int a_stream;   // n_features > 128: X K-blocks are not kept across centroid tiles but re-streamed per tile
  int pack;       // 1, or 2: two consecutive X rows share one operand row; centroids are block-diagonal
  int k_sub;      // pack == 2: accumulator columns per packed group (bn == 2 * k_sub)
  int raw_slots;  // A-in-TMEM variant: raw X ring depth; a_slots then counts 64-column TMEM operand slots
  int a_col0;     // A-in-TMEM variant: first TMEM column of the operand slots (= n_acc * bn)
  uint32_t tmem_cols;
  const float* cnh;  // [k_pad] 1/2 ||c||^2, +inf for padding
  int32_t* labels;
  float* dbg_dots;   // optional [n, k_pad] dump of the x.c accumulators (tests only)
  long long* dbg_clk; // optional [16]: per-role (wait cycles, total cycles) of CTA 0 (env CUML_B200_DBG_CLK)
  int l2_ahead;      // CTA-pair kernel: row tiles prefetched into L2 ahead of the shared-memory ring
  int fold;          // CTA-pair kernel: -1/2||c||^2 enters the accumulator through one extra K=8 MMA (ones x pieces)
  // Distance-matrix mode (DIST kernels, ML::kmeans::transform) reuses fields that are idle there, so that the
  // parameter block of the hot E-step instantiations keeps its size (growing it cost them register spills):
  //   dbg_dots -> output [n, k_sub] | labels -> ||x_i||^2 (as float*) | k_sub -> true n_clusters (row pitch) |
  //   raw_slots -> 1: write sqrt(max(d, 0))
};

struct Barriers {
  uint64_t a_raw_full[MAX_A_SLOTS], a_ready[MAX_A_SLOTS], a_empty[MAX_A_SLOTS];
  uint64_t b_full[MAX_STAGES], b_empty[MAX_STAGES];
  uint64_t acc_full[MAX_ACC], acc_empty[MAX_ACC];
  uint64_t raw_full[MAX_RAW], raw_empty[MAX_RAW];   // A-in-TMEM variant: raw X ring in shared memory
  uint64_t cn_full;                                 // CTA-pair kernel: folded half-norm tiles have landed
  uint64_t lab_full[MAX_A_SLOTS];                   // fused M-step: the labels of the tile in X slot s are in shared memory
  uint32_t tmem_base;
};

// ---- fused M-step of the row-packed tensor-memory kernel (fused_l2_argmin_tsp_kernel<MSTEP>; n_features = 16, unweighted)
// One pass over X per Lloyd iteration for the HBM-bound small-d shapes: the raw fp32 tile is still in its X slot when
// the row-owner epilogue knows the tile's labels, so accumulate warps add the rows into warp-private [k_sub + 1 x 32]
// tables (lane = (data row of the packed pair, column): lane l only ever touches bank l) before the slot is released.
// The slot's empty barrier then counts the converter and the accumulate warps.  Labels travel through a per-slot
// shared-memory buffer (one word per operand row: table-row byte offset of the even data row | odd row << 16); cluster
// sizes are integer shared-memory counts kept by the epilogue threads (exact).  At the end the CTA folds its tables in
// a fixed order into the partials format of the M-step kernels (deterministic).
constexpr int TSP_ACC_WARPS = 8;   // A-in-TMEM kernel: warps 20..27 accumulate (the row-owner epilogue then has 3 groups)
__host__ __device__ inline size_t tsp_mstep_smem_bytes(int k_sub)
{
  return static_cast<size_t>(TSP_ACC_WARPS) * (k_sub + 1) * 32 * sizeof(float)   // private tables (+ a dummy row each)
         + static_cast<size_t>(k_sub) * 2 * sizeof(int)                           // counts
         + static_cast<size_t>(MAX_A_SLOTS) * TILE_M * sizeof(uint32_t)           // label words per X slot
         + 128;
}

// ===================== epilogue role (shared by all kernel variants) ===================================
// 16 warps: warp e handles TMEM lanes [32*(e&3), +32) (its rows) and column part (e>>2) of every
// accumulator tile (BN/4 columns, at least one 32-column chunk); thread = row.  Per row: four independent
// running (min, argmin) chains, merged with the first-minimum rule; the column parts are merged through
// shared memory.  The argmin is instruction-issue bound (4 instructions per distance), hence many warps.
// DIST: 0 = argmin epilogue, 1 = distance matrix (transform), 2 = distance matrix with the lane-pair store pattern,
//       3 = argmin epilogue that also stores the winning value 1/2||c||^2 - x.c per row (seeding: min distance update)
template <bool PAIR, int DIST = 0, bool FOLD1 = false, bool CLK = false>
__device__ __forceinline__ void epilogue_role(const FusedParams& p, Barriers* bars, float* cn_s, float* mrg_v,
                                              int* mrg_i, uint32_t tmem_base, int64_t first_row, int64_t row_stride,
                                              int64_t n_tiles_cta)
{
  const int et      = threadIdx.x - 256;      // 0..511
  const int ew      = et >> 5;                // epilogue warp 0..15
  const int lane    = et & 31;
  const int quarter = ew & 3;                 // TMEM lane quarter this warp may access
  const int part    = ew >> 2;                // column part of the accumulator tile
  const int rit     = quarter * 32 + lane;    // row within the 128-row tile
  const float inf   = __int_as_float(0x7f800000);
  const int nparts  = min(4, p.bn / 32);      // parts that own at least one 32-column chunk
  const int pcols   = p.bn / nparts;
  const int cbeg    = part * pcols;
  const int cend    = (part < nparts) ? cbeg + pcols : cbeg;
  // row packing: the accumulator columns [g*k_sub, (g+1)*k_sub) belong to data row pack*r + g
  const int ppg     = max(1, nparts / p.pack);                       // column parts per packed group
  const int grp     = (p.pack > 1 && part < nparts) ? part / ppg : 0;
  const int col0    = grp * p.k_sub;                                 // first accumulator column of my group
  const bool lead   = (part < nparts) && (part % ppg == 0);          // merges its group's parts, stores labels
  uint32_t acc_cnt  = 0;
  Ring racc;
  // 1/2||c||^2 of the next centroid tile is fetched one tile ahead (registers), so its global-load
  // latency never sits on the epilogue's critical path
  float pre = 0.f;
  auto fetch_cn = [&](int nt) { pre = __ldg(p.cnh + static_cast<int64_t>(nt) * p.bn + (et < p.bn ? et : 0)); };
  const bool fold = (PAIR || FOLD1) && p.fold;   // accumulator already holds x.c - 1/2||c||^2: pick the maximum
  long long w_acc = 0, w_merge = 0, t_hold = 0, t_start = 0;   // CLK only: waits for the accumulator / the merge
  if constexpr (CLK) t_start = clock64();                      // barriers, time an accumulator is held
  if (!fold) fetch_cn(0);
  if (p.k_tiles == 1 && !fold) {  // single centroid tile: stage the half norms once
    if (et < p.bn) cn_s[et] = pre;
    ptx::named_bar_sync(1, EPI_THREADS);
  }
  for (int64_t t = 0; t < n_tiles_cta; ++t) {
    float b0 = inf, b1 = inf, b2 = inf, b3 = inf;
    int i0 = 0, i1 = 0, i2 = 0, i3 = 0;
    for (int nt = 0; nt < p.k_tiles; ++nt, ++acc_cnt) {
      const uint32_t acc = racc.slot, pacc = racc.phase;
      racc.advance(p.n_acc);
      float* cn = cn_s + ((p.k_tiles == 1) ? 0 : (acc_cnt & 1u) * p.bn);
      if (p.k_tiles > 1 && !fold) {
        if (et < p.bn) cn[et] = pre;
        fetch_cn(nt + 1 == p.k_tiles ? 0 : nt + 1);
        ptx::named_bar_sync(1, EPI_THREADS);
      }
      CB2_WAIT_CLK(ptx::smem_u32(&bars->acc_full[acc]), pacc, w_acc);
      long long t_got = 0;
      if constexpr (CLK) t_got = clock64();
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * p.bn;
      const int jbase      = nt * p.bn;
      uint32_t r[32];
      for (int c0 = cbeg; c0 < cend; c0 += 32) {
        ptx::tmem_ld_32x32(taddr + c0, r);
        ptx::tmem_ld_wait();
        if (!DIST && p.dbg_dots) {
          const int64_t row = first_row + t * row_stride + rit;
          if (row < p.n) {
            float* o = p.dbg_dots + row * (static_cast<int64_t>(p.k_tiles) * p.bn) + jbase + c0;
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(r[j]);
          }
        }
        if (DIST == 2) {
          // store pattern of the CTA-pair kernel (1.2 -> 1.7 TB/s of output at the C4 shape): lanes 2i and 2i+1 swap half of every
          // 8-column group, so each store instruction writes 32 contiguous bytes (a whole sector) of ONE row per
          // lane pair instead of 16 bytes in two different rows.  Same values, same addresses, other lanes.
          const int64_t row   = first_row + t * row_stride + rit;
          const int dist_k    = p.k_sub;
          const bool odd      = (lane & 1) != 0;
          const int64_t row_e = row - (odd ? 1 : 0);          // the pair's even / odd rows (consecutive)
          const int64_t row_o = row_e + 1;
          const float xx      = (row < p.n) ? __ldg(reinterpret_cast<const float*>(p.labels) + row) : 0.0f;
          const float* cnc    = cn + c0;
          if ((dist_k & 3) == 0 && jbase + c0 + 32 <= dist_k) {   // warp-uniform: shuffles are legal
            float* oe = p.dbg_dots + row_e * static_cast<int64_t>(dist_k) + jbase + c0 + (odd ? 4 : 0);
            float* oo = p.dbg_dots + row_o * static_cast<int64_t>(dist_k) + jbase + c0 + (odd ? 4 : 0);
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              float a[4], b[4], rcv[4];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int c   = g8 * 8 + j;
                const float s = fold ? __uint_as_float(r[c]) : __uint_as_float(r[c]) - cnc[c];
                float v       = fmaxf(xx - 2.0f * s, 0.0f);
                if (p.raw_slots) v = sqrtf(v);
                if (j < 4) a[j] = v; else b[j - 4] = v;
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) rcv[j] = __shfl_xor_sync(0xffffffffu, odd ? a[j] : b[j], 1);
              // even lane: own a -> row_e, partner's a -> row_o;  odd lane: partner's b -> row_e, own b -> row_o
              const float4 ve = odd ? make_float4(rcv[0], rcv[1], rcv[2], rcv[3]) : make_float4(a[0], a[1], a[2], a[3]);
              const float4 vo = odd ? make_float4(b[0], b[1], b[2], b[3]) : make_float4(rcv[0], rcv[1], rcv[2], rcv[3]);
              if (row_e < p.n) *reinterpret_cast<float4*>(oe + g8 * 8) = ve;
              if (row_o < p.n) *reinterpret_cast<float4*>(oo + g8 * 8) = vo;
            }
          } else if (row < p.n) {
            float* o = p.dbg_dots + row * static_cast<int64_t>(dist_k) + jbase + c0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float s = fold ? __uint_as_float(r[j]) : __uint_as_float(r[j]) - cnc[j];
              float v       = fmaxf(xx - 2.0f * s, 0.0f);
              if (p.raw_slots) v = sqrtf(v);
              if (jbase + c0 + j < dist_k) o[j] = v;
            }
          }
          continue;
        }
        if (DIST == 1) {
          // transform: ||x - c||^2 = ||x||^2 - 2 (x.c - 1/2||c||^2); this thread holds 32 consecutive columns of its row
          const int64_t row = first_row + t * row_stride + rit;
          if (row < p.n) {
            const int dist_k = p.k_sub;
            const float xx = __ldg(reinterpret_cast<const float*>(p.labels) + row);
            float* o       = p.dbg_dots + row * static_cast<int64_t>(dist_k) + jbase + c0;
            const float* cnc = cn + c0;
            float dv[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float s = fold ? __uint_as_float(r[j]) : __uint_as_float(r[j]) - cnc[j];
              dv[j]         = fmaxf(xx - 2.0f * s, 0.0f);
              if (p.raw_slots) dv[j] = sqrtf(dv[j]);
            }
            if ((dist_k & 3) == 0 && jbase + c0 + 32 <= dist_k) {   // 16-byte aligned full chunk: vector stores
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(o + j) = make_float4(dv[j], dv[j + 1], dv[j + 2], dv[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (jbase + c0 + j < dist_k) o[j] = dv[j];
            }
          }
          continue;
        }
        const int jb = jbase + c0 - col0;
        if (fold) {
          // b* hold the NEGATED score so that "smaller wins" and the chain merge below stay the same
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float v0 = -__uint_as_float(r[q4 * 4 + 0]);
            const float v1 = -__uint_as_float(r[q4 * 4 + 1]);
            const float v2 = -__uint_as_float(r[q4 * 4 + 2]);
            const float v3 = -__uint_as_float(r[q4 * 4 + 3]);
            if (v0 < b0) { b0 = v0; i0 = jb + q4 * 4 + 0; }
            if (v1 < b1) { b1 = v1; i1 = jb + q4 * 4 + 1; }
            if (v2 < b2) { b2 = v2; i2 = jb + q4 * 4 + 2; }
            if (v3 < b3) { b3 = v3; i3 = jb + q4 * 4 + 3; }
          }
        } else {
          const float4* cn4 = reinterpret_cast<const float4*>(cn + c0);
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 c4 = cn4[q4];
            const float v0 = c4.x - __uint_as_float(r[q4 * 4 + 0]);
            const float v1 = c4.y - __uint_as_float(r[q4 * 4 + 1]);
            const float v2 = c4.z - __uint_as_float(r[q4 * 4 + 2]);
            const float v3 = c4.w - __uint_as_float(r[q4 * 4 + 3]);
            if (v0 < b0) { b0 = v0; i0 = jb + q4 * 4 + 0; }
            if (v1 < b1) { b1 = v1; i1 = jb + q4 * 4 + 1; }
            if (v2 < b2) { b2 = v2; i2 = jb + q4 * 4 + 2; }
            if (v3 < b3) { b3 = v3; i3 = jb + q4 * 4 + 3; }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&bars->acc_empty[acc]), 0));
        else ptx::mbar_arrive(ptx::smem_u32(&bars->acc_empty[acc]));
      }
      if constexpr (CLK) t_hold += clock64() - t_got;
    }
    if (DIST == 1 || DIST == 2) continue;   // distance-matrix mode: nothing to merge, no labels
    // merge the chains: smaller value wins, equal values -> smaller index (first minimum)
    if (b1 < b0 || (b1 == b0 && i1 < i0)) { b0 = b1; i0 = i1; }
    if (b3 < b2 || (b3 == b2 && i3 < i2)) { b2 = b3; i2 = i3; }
    if (b2 < b0 || (b2 == b0 && i2 < i0)) { b0 = b2; i0 = i2; }
    // merge the column parts of each group through shared memory
    if (part < nparts && !lead) {
      mrg_v[(part - 1) * TILE_M + rit] = b0;
      mrg_i[(part - 1) * TILE_M + rit] = i0;
    }
    if constexpr (CLK) {
      const long long t__ = clock64();
      ptx::named_bar_sync(2, EPI_THREADS);
      w_merge += clock64() - t__;
    } else {
      ptx::named_bar_sync(2, EPI_THREADS);
    }
    if (lead) {
      for (int q = 1; q < ppg; ++q) {
        const float ov = mrg_v[(part + q - 1) * TILE_M + rit];
        const int oi   = mrg_i[(part + q - 1) * TILE_M + rit];
        if (ov < b0 || (ov == b0 && oi < i0)) { b0 = ov; i0 = oi; }
      }
      const int64_t row = (first_row + t * row_stride + rit) * p.pack + grp;
      if (row < p.n) {
        p.labels[row] = i0;
        if (DIST == 3) p.dbg_dots[row] = b0;   // b0 = 1/2||c||^2 - x.c of the winner in both (folded / staged) modes
      }
    }
    if constexpr (CLK) {
      const long long t__ = clock64();
      ptx::named_bar_sync(3, EPI_THREADS);
      w_merge += clock64() - t__;
    } else {
      ptx::named_bar_sync(3, EPI_THREADS);  // mrg_* may be overwritten by the next tile
    }
  }
  if constexpr (CLK) {
    if (p.dbg_clk && blockIdx.x == 0 && et == 0) {
      p.dbg_clk[8] = w_acc; p.dbg_clk[9] = clock64() - t_start; p.dbg_clk[10] = t_hold; p.dbg_clk[11] = w_merge;
    }
  }
}

// ===================== row-owner epilogue (single centroid tile, BN <= 128) ============================================
// Measured at C5: fused kernel 6.9 -> 5.9 ms on the 3xTF32 kernel, 6.5 -> 4.9 ms on the tf32 + bf16 twin
// (profiles/r02_ab_table.txt).  With few clusters the epilogue above
// gives each of its 16 warps ONE 32-column chunk per tile and then pays two 512-thread named barriers and a shared-
// memory merge per tile: a per-tile latency chain (accumulator wait -> tcgen05.ld -> compare -> barrier -> merge ->
// store -> barrier) that all 16 warps walk together, one tile at a time.  Here the 16 warps form 4 groups of 4 (one
// warp per TMEM lane quarter); group g owns the row tiles t = g, g + 4, ... of its CTA, each thread scans ALL columns
// of its row (both packed groups when two data rows share an operand row) and stores the label itself: no merge, no
// named barriers, and four tiles' epilogues in flight on different accumulator stages (BN <= 128 leaves >= 4 stages).
// Only 4 warps arrive on an accumulator's empty barrier (the kernel initialises it with 4 in this mode).
template <bool FOLD1, bool MSTEP = false>
__device__ __forceinline__ void epilogue_role_rowown(const FusedParams& p, Barriers* bars, float* cn_s,
                                                     uint32_t tmem_base, int64_t first_row, int64_t row_stride,
                                                     int64_t n_tiles_cta, uint32_t* ms_labels = nullptr,
                                                     int* ms_counts = nullptr, int x_slots = 0, int lab_shift = 0)
{
  const int et      = threadIdx.x - 256;      // 0..511
  const int ew      = et >> 5;                // epilogue warp 0..15
  const int lane    = et & 31;
  const int quarter = ew & 3;                 // TMEM lane quarter this warp may access
  const int group   = ew >> 2;                // row tiles t = group, group + 4, ...
  const int rit     = quarter * 32 + lane;    // row within the 128-row tile
  const float inf   = __int_as_float(0x7f800000);
  const bool fold   = FOLD1 && p.fold;        // accumulator already holds x.c - 1/2||c||^2
  const int gcols   = p.pack > 1 ? p.k_sub : p.bn;   // accumulator columns per data row
  if (!fold) {   // stage the half norms once (single centroid tile)
    if (et < p.bn) cn_s[et] = __ldg(p.cnh + et);
    ptx::named_bar_sync(1, EPI_THREADS);
  }
  // A group must see EVERY phase of an accumulator barrier it waits on (a parity wait cannot tell phase k from k + 2),
  // so the number of groups divides the ring: 4 groups when n_acc >= 4 (a multiple of 4 or exactly 4..8 stages walked
  // with stride 4 -- n_acc is then 4 or 8), otherwise one group per accumulator stage and the remaining warps idle
  const int n_groups = p.n_acc >= 4 ? 4 : p.n_acc;
  if (group >= n_groups) return;
  Ring racc;                  // accumulator stage of tile t: t % n_acc, phase (t / n_acc) & 1
  racc.advance_by(static_cast<uint32_t>(group), p.n_acc);
  const uint32_t xs_n = static_cast<uint32_t>(x_slots > 0 ? x_slots : p.a_slots);
  Ring rslot;                 // MSTEP: X slot of tile t (one K-block per tile): t % (X ring depth)
  if (MSTEP) rslot.advance_by(static_cast<uint32_t>(group), xs_n);
  for (int64_t t = group; t < n_tiles_cta; t += n_groups, racc.advance_by(n_groups, p.n_acc)) {
    const uint32_t acc = racc.slot, pacc = racc.phase;
    ptx::mbar_wait_park(ptx::smem_u32(&bars->acc_full[acc]), pacc);
    ptx::tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * p.bn;
    uint32_t lab_word = 0;
    for (int g = 0; g < p.pack; ++g) {
      float b0 = inf, b1 = inf, b2 = inf, b3 = inf;
      int i0 = 0, i1 = 0, i2 = 0, i3 = 0;
      uint32_t r[32];
      for (int c0 = 0; c0 < gcols; c0 += 32) {
        const int col = g * gcols + c0;
        ptx::tmem_ld_32x32(taddr + col, r);
        ptx::tmem_ld_wait();
        if (fold) {
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float v0 = -__uint_as_float(r[q4 * 4 + 0]);
            const float v1 = -__uint_as_float(r[q4 * 4 + 1]);
            const float v2 = -__uint_as_float(r[q4 * 4 + 2]);
            const float v3 = -__uint_as_float(r[q4 * 4 + 3]);
            if (v0 < b0) { b0 = v0; i0 = c0 + q4 * 4 + 0; }
            if (v1 < b1) { b1 = v1; i1 = c0 + q4 * 4 + 1; }
            if (v2 < b2) { b2 = v2; i2 = c0 + q4 * 4 + 2; }
            if (v3 < b3) { b3 = v3; i3 = c0 + q4 * 4 + 3; }
          }
        } else {
          const float4* cn4 = reinterpret_cast<const float4*>(cn_s + col);
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 c4 = cn4[q4];
            const float v0 = c4.x - __uint_as_float(r[q4 * 4 + 0]);
            const float v1 = c4.y - __uint_as_float(r[q4 * 4 + 1]);
            const float v2 = c4.z - __uint_as_float(r[q4 * 4 + 2]);
            const float v3 = c4.w - __uint_as_float(r[q4 * 4 + 3]);
            if (v0 < b0) { b0 = v0; i0 = c0 + q4 * 4 + 0; }
            if (v1 < b1) { b1 = v1; i1 = c0 + q4 * 4 + 1; }
            if (v2 < b2) { b2 = v2; i2 = c0 + q4 * 4 + 2; }
            if (v3 < b3) { b3 = v3; i3 = c0 + q4 * 4 + 3; }
          }
        }
      }
      // merge the chains: smaller value wins, equal values -> smaller index (first minimum)
      if (b1 < b0 || (b1 == b0 && i1 < i0)) { b0 = b1; i0 = i1; }
      if (b3 < b2 || (b3 == b2 && i3 < i2)) { b2 = b3; i2 = i3; }
      if (b2 < b0 || (b2 == b0 && i2 < i0)) { b0 = b2; i0 = i2; }
      const int64_t row = (first_row + t * row_stride + rit) * p.pack + g;
      if (row < p.n) p.labels[row] = i0;
      if (MSTEP) {
        lab_word |= (static_cast<uint32_t>(i0) << lab_shift) << (16 * g);   // label, or its table-row byte offset
        if (row < p.n) atomicAdd(ms_counts + i0, 1);   // integer: exact and order-independent
      }
    }
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->acc_empty[acc]));
    if (MSTEP) {
      ms_labels[rslot.slot * TILE_M + rit] = lab_word;
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->lab_full[rslot.slot]));   // release: the words above are visible
      rslot.advance_by(n_groups, xs_n);
    }
  }
}

template <int DIST>
__global__ void __launch_bounds__(NUM_THREADS, 1)
fused_l2_argmin_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_hi,
                       const __grid_constant__ CUtensorMap tm_lo, const FusedParams p)
{
  extern __shared__ uint8_t smem_dyn[];
  // 128B-swizzled operand tiles need 1024-byte alignment
  const uint32_t raw_base = ptx::smem_u32(smem_dyn);
  const uint32_t base     = (raw_base + 1023u) & ~1023u;
  uint8_t* gbase          = smem_dyn + (base - raw_base);

  const uint32_t b_half_bytes  = static_cast<uint32_t>(p.bn) * 128u;
  const uint32_t b_stage_bytes = 2u * b_half_bytes;                                // hi then lo
  const uint32_t a_base  = base;
  const uint32_t b_base  = a_base + p.a_slots * A_SLOT_BYTES;
  const uint32_t cn_off  = p.a_slots * A_SLOT_BYTES + p.b_stages * b_stage_bytes;
  float* cn_s            = reinterpret_cast<float*>(gbase + cn_off);               // [2][bn]
  float* mrg_v           = cn_s + 2 * p.bn;                                         // [128] epilogue half merge
  int* mrg_i             = reinterpret_cast<int*>(mrg_v + 3 * TILE_M);              // [3][128]
  Barriers* bars         = reinterpret_cast<Barriers*>(gbase + cn_off + 2u * p.bn * sizeof(float) + 6u * TILE_M * 4u);

  const int warp = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;

  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_A_SLOTS; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bars->a_raw_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->a_ready[s]), 128);
      ptx::mbar_init(ptx::smem_u32(&bars->a_empty[s]), 1);
    }
    for (int s = 0; s < MAX_ACC; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bars->acc_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->acc_empty[s]), 16);   // one arrive per epilogue warp of the accumulator
    }
    for (int s = 0; s < MAX_STAGES; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bars->b_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->b_empty[s]), 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_x);
    ptx::prefetch_tmap(&tm_hi);
    ptx::prefetch_tmap(&tm_lo);
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(&bars->tmem_base), p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================== A producer: raw X K-blocks into the slot ring =====================
    {
      uint32_t a_cnt = 0;
      Ring ra;
      long long wcyc = 0;
      const long long tstart = clock64();
      const int a_reps = p.a_stream ? p.k_tiles : 1;   // streamed mode reloads the row tile per centroid tile
      for (int64_t tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
        for (int rep = 0; rep < a_reps; ++rep)
        for (int kbi = 0; kbi < p.kb; ++kbi, ++a_cnt) {
          const uint32_t sa = ra.slot, pa = ra.phase;
          ra.advance(p.a_slots);
          CB2_TIMED_WAIT(ptx::smem_u32(&bars->a_empty[sa]), pa ^ 1u, wcyc);
          if (ptx::elect_one()) {
            const uint32_t full = ptx::smem_u32(&bars->a_raw_full[sa]);
            ptx::mbar_arrive_expect_tx(full, KBLOCK_BYTES);
            ptx::tma_load_2d_hint(a_base + sa * A_SLOT_BYTES, &tm_x, kbi * KBLOCK, static_cast<int32_t>(tile * TILE_M),
                                  full, ptx::kEvictFirst);
          }
          __syncwarp();
        }
      }
      if (p.dbg_clk && blockIdx.x == 0 && lane == 0) { p.dbg_clk[0] = wcyc; p.dbg_clk[1] = clock64() - tstart; }
    }
  } else if (warp == 2) {
    // ===================== B producer: centroid hi/lo K-blocks =====================
    {
      uint32_t b_cnt = 0;
      Ring rb;
      for (int64_t tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
        if (p.b_resident && tile != blockIdx.x) break;  // resident centroids: loaded once per CTA
        for (int nt = 0; nt < p.k_tiles; ++nt) {
          for (int kbi = 0; kbi < p.kb; ++kbi, ++b_cnt) {
            const uint32_t sb = rb.slot, pb = rb.phase;
            rb.advance(p.b_stages);
            ptx::mbar_wait(ptx::smem_u32(&bars->b_empty[sb]), pb ^ 1u);
            if (ptx::elect_one()) {
              const uint32_t full = ptx::smem_u32(&bars->b_full[sb]);
              ptx::mbar_arrive_expect_tx(full, b_stage_bytes);
              const uint32_t dst = b_base + sb * b_stage_bytes;
              ptx::tma_load_2d_hint(dst, &tm_hi, kbi * KBLOCK, nt * p.bn, full, ptx::kEvictLast);
              ptx::tma_load_2d_hint(dst + b_half_bytes, &tm_lo, kbi * KBLOCK, nt * p.bn, full, ptx::kEvictLast);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== converter: raw -> (hi in place, lo) =====================
    const int ct = threadIdx.x - 128;  // 0..127
    uint32_t a_cnt = 0;
      Ring ra;
    long long wcyc = 0;
    const long long tstart = clock64();
    const int a_reps = p.a_stream ? p.k_tiles : 1;
    for (int64_t tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
      for (int rep = 0; rep < a_reps; ++rep)
      for (int kbi = 0; kbi < p.kb; ++kbi, ++a_cnt) {
        const uint32_t sa = ra.slot, pa = ra.phase;
        ra.advance(p.a_slots);
        CB2_TIMED_WAIT(ptx::smem_u32(&bars->a_raw_full[sa]), pa, wcyc);
        uint4* hi = reinterpret_cast<uint4*>(gbase + sa * A_SLOT_BYTES);
        uint4* lo = reinterpret_cast<uint4*>(gbase + sa * A_SLOT_BYTES + KBLOCK_BYTES);
        // 16-byte chunks the tensor core will read in this K-block: 2 per K=8 step that holds real columns
        const int live_chunks = 2 * min(4, (p.d - kbi * KBLOCK + 7) / 8);
#pragma unroll
        for (int i = 0; i < KBLOCK_BYTES / 16 / 128; ++i) {
          const int e = ct + i * 128;
          // logical chunk of this physical position under the 128B swizzle: pos ^ (row & 7)
          if (((e & 7) ^ ((e >> 3) & 7)) >= live_chunks) continue;
          uint4 v = hi[e];
          uint4 h, l;
          h.x = v.x & 0xffffe000u; h.y = v.y & 0xffffe000u; h.z = v.z & 0xffffe000u; h.w = v.w & 0xffffe000u;
          l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x));
          l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y));
          l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z));
          l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w));
          hi[e] = h;
          lo[e] = l;
        }
        ptx::fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
        ptx::mbar_arrive(ptx::smem_u32(&bars->a_ready[sa]));
      }
    }
    if (p.dbg_clk && blockIdx.x == 0 && ct == 0) { p.dbg_clk[2] = wcyc; p.dbg_clk[3] = clock64() - tstart; }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loop in uniform control flow (so descriptors live in uniform registers and
    // each tcgen05.mma is a single predicated instruction); one elected lane issues.
    {
      const uint32_t idesc = ptx::umma_idesc_tf32(TILE_M, p.bn);
      uint32_t b_cnt = 0, acc_cnt = 0;
      Ring ra_tile, ra_run, rb, racc;
      long long wacc = 0, wa = 0, wb = 0, tissue = 0, tcommit = 0, tc0 = 0;
      const long long tstart = clock64();
      for (int64_t tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ra_tile.advance_by(p.kb, p.a_slots)) {
        for (int nt = 0; nt < p.k_tiles; ++nt, ++acc_cnt) {
          const uint32_t acc = racc.slot, pacc = racc.phase;
          racc.advance(p.n_acc);
          Ring ra = p.a_stream ? ra_run : ra_tile;
          CB2_TIMED_WAIT(ptx::smem_u32(&bars->acc_empty[acc]), pacc ^ 1u, wacc);
          const uint32_t d_tmem = tmem_base + acc * p.bn;
          for (int kbi = 0; kbi < p.kb; ++kbi, ++b_cnt) {
            const uint32_t sa = ra.slot, pa = ra.phase;
            ra.advance(p.a_slots);
            if (nt == 0 || p.a_stream) CB2_TIMED_WAIT(ptx::smem_u32(&bars->a_ready[sa]), pa, wa);  // first use
            uint32_t sb = rb.slot;
            const uint32_t pb = rb.phase;
            rb.advance(p.b_stages);
            if (p.b_resident) {
              sb = nt * p.kb + kbi;
              if (tile == static_cast<int64_t>(blockIdx.x)) ptx::mbar_wait(ptx::smem_u32(&bars->b_full[sb]), 0u);
            } else {
              CB2_TIMED_WAIT(ptx::smem_u32(&bars->b_full[sb]), pb, wb);
            }
            ptx::tc_fence_after();
            const uint64_t da_hi = ptx::umma_desc_sw128(a_base + sa * A_SLOT_BYTES);
            const uint64_t da_lo = ptx::umma_desc_sw128(a_base + sa * A_SLOT_BYTES + KBLOCK_BYTES);
            const uint64_t db_hi = ptx::umma_desc_sw128(b_base + sb * b_stage_bytes);
            const uint64_t db_lo = ptx::umma_desc_sw128(b_base + sb * b_stage_bytes + b_half_bytes);
            const int nks = min(4, (p.d - kbi * KBLOCK + 7) / 8);  // K=8 steps that hold real columns
            const long long ti0 = p.dbg_clk ? clock64() : 0;
            if (ptx::elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                if (ks >= nks) break;
                const uint64_t adv = static_cast<uint64_t>(ks * 2);  // 8 tf32 = 32 bytes = 2 x 16B units
                // small terms first, then the dominant hi.hi term
                ptx::mma_tf32_ss(d_tmem, da_lo + adv, db_hi + adv, idesc, (kbi | ks) != 0 ? 1u : 0u);
                ptx::mma_tf32_ss(d_tmem, da_hi + adv, db_lo + adv, idesc, 1u);
                ptx::mma_tf32_ss(d_tmem, da_hi + adv, db_hi + adv, idesc, 1u);
              }
            }
            __syncwarp();
            const long long ti1 = p.dbg_clk ? clock64() : 0;
            tissue += ti1 - ti0;
            if (ptx::elect_one()) {
              if (!p.b_resident) ptx::mma_commit(ptx::smem_u32(&bars->b_empty[sb]));  // frees the B stage
              // last centroid tile: this X K-block is not needed again -> release its slot early so
              // the next row tile's load + hi/lo split overlaps the remaining K-blocks
              if (nt == p.k_tiles - 1 || p.a_stream) ptx::mma_commit(ptx::smem_u32(&bars->a_empty[sa]));
            }
            __syncwarp();
            tc0 = ti1;
          }
          ra_run = ra;
          if (ptx::elect_one()) ptx::mma_commit(ptx::smem_u32(&bars->acc_full[acc]));  // accumulator ready
          __syncwarp();
          if (p.dbg_clk) tcommit += clock64() - tc0;
        }
      }
      if (p.dbg_clk && blockIdx.x == 0 && lane == 0) {
        p.dbg_clk[4] = wacc; p.dbg_clk[5] = wa; p.dbg_clk[6] = wb; p.dbg_clk[7] = clock64() - tstart;
        (void)tissue; (void)tcommit;
      }
    }
  } else if (warp >= 8) {
    // ===================== epilogue: argmin over the accumulator =====================
    const int64_t n_mine = (p.m_tiles > static_cast<int64_t>(blockIdx.x))
                             ? (p.m_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    epilogue_role<false, DIST>(p, bars, cn_s, mrg_v, mrg_i, tmem_base, static_cast<int64_t>(blockIdx.x) * TILE_M,
                                 static_cast<int64_t>(gridDim.x) * TILE_M, n_mine);
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, p.tmem_cols);
}

// =================================================================================================
// CTA-pair variant (tcgen05 cta_group::2).  Two CTAs on one TPC work on a 256-row tile: each loads and
// splits its own 128 rows of X, each holds HALF of every centroid block (BN/2 rows), and the pair leader
// issues M=256 MMAs that read both CTAs' shared memory and write both CTAs' TMEM.  Per CTA this halves
// the shared-memory footprint and read bandwidth of the centroid operand, which buys a deeper X ring
// (the limiter of the single-CTA kernel at d=64, k=256).  Same pipelines as above; barriers that gate
// the MMA issuer live in the leader and are arrived on remotely, barriers released by MMA completion
// are signalled in both CTAs with a multicast tcgen05.commit.
// BF16C: the two correction terms x_lo.c_hi and x_hi.c_lo are computed with bf16 operands (kind::f16, K = 16:
// half the MMA instructions of a tf32 term).  hi is then the round-to-nearest tf32 of x, so |lo| <= 2^-12 |x|, and
// rounding lo and hi to bf16 perturbs each correction by <= 2^-9 relative: ~2^-20 |x||c| per product -- the size of
// the lo.lo term every 3xTF32 scheme drops.  Operand slot layout per 32-feature K-block:
//   [0, 16 KB) hi tf32, 128B swizzle | [16 KB, 24 KB) hi bf16, 64B swizzle | [24 KB, 32 KB) lo bf16, 64B swizzle
// TRUNC (the shipped E-step instantiation; C3 12.3 -> 11.9 ms): hi = x truncated to tf32, which is what the tensor core reads
// from the raw fp32 tile anyway, so the converter neither rounds nor rewrites the tile (fewer instructions on the
// issue-bound d = 64 path) at the price of |lo| <= 2^-11 |x| instead of 2^-12.
template <bool BF16C, int DIST = 0, bool TRUNC = false, bool CLK = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PAIR_THREADS, 1)
fused_l2_argmin_2cta_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_hi,
                            const __grid_constant__ CUtensorMap tm_lo, const __grid_constant__ CUtensorMap tm_lb,
                            const __grid_constant__ CUtensorMap tm_cn, const FusedParams p)
{
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t raw_base = ptx::smem_u32(smem_dyn);
  const uint32_t base     = (raw_base + 1023u) & ~1023u;
  uint8_t* gbase          = smem_dyn + (base - raw_base);

  const uint32_t cta_rank = ptx::cluster_ctarank();
  const bool leader       = cta_rank == 0;
  const int64_t pair      = blockIdx.x >> 1;
  const int64_t n_pairs   = gridDim.x >> 1;
  const int half_n        = p.bn / 2;                                   // centroid rows held by this CTA

  const uint32_t b_half_bytes  = static_cast<uint32_t>(half_n) * 128u;   // hi (or lo) rows of this CTA
  const uint32_t b_stage_bytes = 2u * b_half_bytes;                     // hi then lo
  const uint32_t a_base  = base;
  const uint32_t b_base  = a_base + p.a_slots * A_SLOT_BYTES;
  // fold tiles (p.fold): ones [128 rows x 8 tf32] then one [half_n rows x 8 tf32] tile of half-norm pieces per
  // centroid tile, 32-byte rows, 4 KB each
  constexpr uint32_t FOLD_TILE = TILE_M * 32u;
  const uint32_t fold_off = p.a_slots * A_SLOT_BYTES + p.b_stages * b_stage_bytes;
  const uint32_t fold_u32 = base + fold_off;
  const uint32_t cn_off   = fold_off + (p.fold ? (1u + p.k_tiles) * FOLD_TILE : 0u);
  float* cn_s            = reinterpret_cast<float*>(gbase + cn_off);    // [2][bn]
  float* mrg_v           = cn_s + 2 * p.bn;                              // [128] epilogue half merge
  int* mrg_i             = reinterpret_cast<int*>(mrg_v + 3 * TILE_M);   // [3][128]
  Barriers* bars         = reinterpret_cast<Barriers*>(gbase + cn_off + 2u * p.bn * sizeof(float) + 6u * TILE_M * 4u);

  const int warp = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;

  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_A_SLOTS; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bars->a_raw_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->a_ready[s]), 16);    // 8 converter warps x 2 CTAs (leader's copy is used)
      ptx::mbar_init(ptx::smem_u32(&bars->a_empty[s]), 1);
    }
    for (int s = 0; s < MAX_ACC; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bars->acc_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->acc_empty[s]), 32);  // 16 epilogue warps x 2 CTAs (leader's copy)
    }
    for (int s = 0; s < MAX_STAGES; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bars->b_full[s]), 1);      // leader's copy: expect_tx covers both CTAs
      ptx::mbar_init(ptx::smem_u32(&bars->b_empty[s]), 1);
    }
    ptx::mbar_init(ptx::smem_u32(&bars->cn_full), 1);
    ptx::fence_barrier_init();
  }
  if (p.fold) {   // all-ones A tile (identical 16-byte chunks: the 32B swizzle does not matter)
    float* ones = reinterpret_cast<float*>(gbase + fold_off);
    for (int i = threadIdx.x; i < TILE_M * 8; i += blockDim.x) ones[i] = 1.0f;
    ptx::fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_x);
    ptx::prefetch_tmap(&tm_hi);
    ptx::prefetch_tmap(&tm_lo);
  }
  if (warp == 1) {
    ptx::tmem_alloc_2cta(ptx::smem_u32(&bars->tmem_base), p.tmem_cols);
    ptx::tmem_relinquish_2cta();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / multicast commit
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const int64_t pair_tiles = (p.m_tiles + 1) / 2;   // 256-row tiles

  if (warp == 0) {
    // ===================== X producer (own 128 rows) =====================
    {
      uint32_t a_cnt = 0;
      Ring ra;
      long long w_empty = 0, t_start = 0;
      if constexpr (CLK) t_start = clock64();
      const int a_reps = p.a_stream ? p.k_tiles : 1;
      for (int64_t pt = pair; pt < pair_tiles; pt += n_pairs) {
        const int32_t row0 = static_cast<int32_t>(pt * 2 * TILE_M + cta_rank * TILE_M);
        // L2 prefetch of the row tile p.l2_ahead rounds ahead: the shared-memory ring only holds ~2 row tiles, too
        // few to cover the ~1.5 us HBM latency; with the tile already in L2 the ring turns around in time
        if (p.l2_ahead > 0) {
          const int64_t pf = pt + static_cast<int64_t>(p.l2_ahead) * n_pairs;
          if (pf < pair_tiles && ptx::elect_one()) {
            const int32_t prow = static_cast<int32_t>(pf * 2 * TILE_M + cta_rank * TILE_M);
            for (int kbi = 0; kbi < p.kb; ++kbi) ptx::tma_prefetch_l2_2d(&tm_x, kbi * KBLOCK, prow);
          }
          __syncwarp();
        }
        for (int rep = 0; rep < a_reps; ++rep)
        for (int kbi = 0; kbi < p.kb; ++kbi, ++a_cnt) {
          const uint32_t sa = ra.slot, pa = ra.phase;
          ra.advance(p.a_slots);
          CB2_WAIT_CLK(ptx::smem_u32(&bars->a_empty[sa]), pa ^ 1u, w_empty);
          if (ptx::elect_one()) {
            const uint32_t full = ptx::smem_u32(&bars->a_raw_full[sa]);
            ptx::mbar_arrive_expect_tx(full, KBLOCK_BYTES);
            ptx::tma_load_2d_hint(a_base + sa * A_SLOT_BYTES, &tm_x, kbi * KBLOCK, row0, full, ptx::kEvictFirst);
          }
          __syncwarp();
        }
      }
      if constexpr (CLK) {
        if (p.dbg_clk && blockIdx.x == 0 && lane == 0) { p.dbg_clk[0] = w_empty; p.dbg_clk[1] = clock64() - t_start; }
      }
    }
  } else if (warp == 2) {
    // ===================== centroid producer (own half of every block) =====================
    {
      uint32_t b_cnt = 0;
      Ring rb;
      if (p.fold && pair < pair_tiles) {
        // half-norm pieces of every centroid tile (this CTA's half of the rows), resident for the whole kernel
        if (ptx::elect_one()) {
          const uint32_t full_local  = ptx::smem_u32(&bars->cn_full);
          const uint32_t full_leader = ptx::mapa(full_local, 0);
          if (leader) ptx::mbar_arrive_expect_tx(full_local, 2u * p.k_tiles * FOLD_TILE);
          for (int nt = 0; nt < p.k_tiles; ++nt)
            ptx::tma_load_2d_2cta(fold_u32 + (1u + nt) * FOLD_TILE, &tm_cn, 0,
                                  nt * p.bn + static_cast<int32_t>(cta_rank) * half_n, full_leader, ptx::kEvictLast);
        }
        __syncwarp();
      }
      for (int64_t pt = pair; pt < pair_tiles; pt += n_pairs) {
        if (p.b_resident && pt != pair) break;
        for (int nt = 0; nt < p.k_tiles; ++nt) {
          for (int kbi = 0; kbi < p.kb; ++kbi, ++b_cnt) {
            const uint32_t sb = rb.slot, pb = rb.phase;
            rb.advance(p.b_stages);
            ptx::mbar_wait_park(ptx::smem_u32(&bars->b_empty[sb]), pb ^ 1u);
            if (ptx::elect_one()) {
              const uint32_t full_local  = ptx::smem_u32(&bars->b_full[sb]);
              const uint32_t full_leader = ptx::mapa(full_local, 0);
              if (leader) ptx::mbar_arrive_expect_tx(full_local, 2u * b_stage_bytes);  // bytes of BOTH CTAs
              const uint32_t dst = b_base + sb * b_stage_bytes;
              const int32_t crow = nt * p.bn + static_cast<int32_t>(cta_rank) * half_n;
              ptx::tma_load_2d_2cta(dst, &tm_hi, kbi * KBLOCK, crow, full_leader, ptx::kEvictLast);
              if (BF16C) {   // tm_lo = bf16 hi, tm_lb = bf16 lo: two half-size tiles
                ptx::tma_load_2d_2cta(dst + b_half_bytes, &tm_lo, kbi * KBLOCK, crow, full_leader, ptx::kEvictLast);
                ptx::tma_load_2d_2cta(dst + b_half_bytes + b_half_bytes / 2, &tm_lb, kbi * KBLOCK, crow, full_leader,
                                      ptx::kEvictLast);
              } else {
                ptx::tma_load_2d_2cta(dst + b_half_bytes, &tm_lo, kbi * KBLOCK, crow, full_leader, ptx::kEvictLast);
              }
            }
            __syncwarp();
          }
        }
      }
    }
  } else if ((warp >= 4 && warp < 8) || warp >= 24) {
    // ===================== converter (8 warps: 4..7 and 24..27) =====================
    const int ct = warp >= 24 ? threadIdx.x - 768 + 128 : threadIdx.x - 128;   // 0..255
    // this thread's chunks are ct + 256 i: 32 rows apart, so the logical chunk and the swizzle phases are fixed
    const int conv_row       = ct >> 3;
    const int conv_lc        = (ct & 7) ^ (conv_row & 7);     // logical 16-byte chunk: features [4 lc, 4 lc + 4)
    const uint32_t conv_off  = static_cast<uint32_t>(conv_row) * 64u +
                               ((static_cast<uint32_t>(conv_lc >> 1) ^ ((conv_row >> 1) & 3u)) << 4) +
                               (static_cast<uint32_t>(conv_lc & 1) << 3);
    uint32_t a_cnt = 0;
      Ring ra;
    long long w_raw = 0, t_start = 0;
    if constexpr (CLK) t_start = clock64();
    const int a_reps = p.a_stream ? p.k_tiles : 1;
    for (int64_t pt = pair; pt < pair_tiles; pt += n_pairs) {
      for (int rep = 0; rep < a_reps; ++rep)
      for (int kbi = 0; kbi < p.kb; ++kbi, ++a_cnt) {
        const uint32_t sa = ra.slot, pa = ra.phase;
        ra.advance(p.a_slots);
        CB2_WAIT_CLK(ptx::smem_u32(&bars->a_raw_full[sa]), pa, w_raw);
        uint4* hi = reinterpret_cast<uint4*>(gbase + sa * A_SLOT_BYTES);
        uint4* lo = reinterpret_cast<uint4*>(gbase + sa * A_SLOT_BYTES + KBLOCK_BYTES);
        const int rem_f       = p.d - kbi * KBLOCK;
        const int live_chunks = BF16C ? 4 * min(2, (rem_f + 15) / 16) : 2 * min(4, (rem_f + 7) / 8);
        uint8_t* hb = gbase + sa * A_SLOT_BYTES + KBLOCK_BYTES;                    // bf16 hi tile (8 KB)
        uint8_t* lb = hb + KBLOCK_BYTES / 2;                                       // bf16 lo tile (8 KB)
        constexpr int CPT = KBLOCK_BYTES / 16 / 256;   // 16-byte chunks per thread
        // all loads first: the chunks are independent, one shared-memory latency instead of CPT
        uint4 v[CPT];
#pragma unroll
        for (int i = 0; i < CPT; ++i) v[i] = hi[ct + i * 256];
        if (conv_lc < live_chunks) {   // chunk ct + 256 i keeps the logical chunk and the swizzle phase of chunk ct
#pragma unroll
          for (int i = 0; i < CPT; ++i) {
            const int e = ct + i * 256;
            uint4 h, l;
            if (BF16C && !TRUNC) {   // nearest tf32 (ties away from zero): |lo| <= 2^-12 |x|
              h.x = (v[i].x + 0x1000u) & 0xffffe000u;
              h.y = (v[i].y + 0x1000u) & 0xffffe000u;
              h.z = (v[i].z + 0x1000u) & 0xffffe000u;
              h.w = (v[i].w + 0x1000u) & 0xffffe000u;
            } else {
              h.x = v[i].x & 0xffffe000u; h.y = v[i].y & 0xffffe000u; h.z = v[i].z & 0xffffe000u; h.w = v[i].w & 0xffffe000u;
            }
            l.x = __float_as_uint(__uint_as_float(v[i].x) - __uint_as_float(h.x));
            l.y = __float_as_uint(__uint_as_float(v[i].y) - __uint_as_float(h.y));
            l.z = __float_as_uint(__uint_as_float(v[i].z) - __uint_as_float(h.z));
            l.w = __float_as_uint(__uint_as_float(v[i].w) - __uint_as_float(h.w));
            if (!TRUNC) hi[e] = h;
            if (BF16C) {
              // 4 features -> 8 bytes of the 64-byte bf16 row (64B swizzle: 16-byte chunk ^= (row / 2) % 4)
              const uint32_t off = conv_off + static_cast<uint32_t>(i) * (32u * 64u);
              const __nv_bfloat162 h01 = __floats2bfloat162_rn(__uint_as_float(h.x), __uint_as_float(h.y));
              const __nv_bfloat162 h23 = __floats2bfloat162_rn(__uint_as_float(h.z), __uint_as_float(h.w));
              const __nv_bfloat162 l01 = __floats2bfloat162_rn(__uint_as_float(l.x), __uint_as_float(l.y));
              const __nv_bfloat162 l23 = __floats2bfloat162_rn(__uint_as_float(l.z), __uint_as_float(l.w));
              *reinterpret_cast<uint2*>(hb + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
              *reinterpret_cast<uint2*>(lb + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
            } else {
              lo[e] = l;
            }
          }
        }
        ptx::fence_proxy_async_smem();  // generic-proxy writes to this CTA's operand tiles -> async proxy (pair MMA)
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&bars->a_ready[sa]), 0));
      }
    }
    if constexpr (CLK) {
      if (p.dbg_clk && blockIdx.x == 0 && ct == 0) { p.dbg_clk[2] = w_raw; p.dbg_clk[3] = clock64() - t_start; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pair leader only; warp-uniform loop, one elected lane issues) ====
    if (leader) {
      const uint32_t idesc   = ptx::umma_idesc_tf32(2 * TILE_M, p.bn);
      const uint32_t idesc16 = ptx::umma_idesc_bf16(2 * TILE_M, p.bn);
      (void)idesc16;
      uint32_t b_cnt = 0, acc_cnt = 0;
      Ring ra_tile, ra_run, rb, racc;
      long long w_acc = 0, w_a = 0, w_b = 0, t_start = 0;
      if constexpr (CLK) t_start = clock64();
      if (p.fold && pair < pair_tiles) ptx::mbar_wait_park(ptx::smem_u32(&bars->cn_full), 0u);
      uint32_t last_kind = 0;   // kind of the last MMA issued: 0 tf32, 1 f16 (BF16C issue order, see below)
      for (int64_t pt = pair; pt < pair_tiles; pt += n_pairs, ra_tile.advance_by(p.kb, p.a_slots)) {
        for (int nt = 0; nt < p.k_tiles; ++nt, ++acc_cnt) {
          const uint32_t acc = racc.slot, pacc = racc.phase;
          racc.advance(p.n_acc);
          Ring ra = p.a_stream ? ra_run : ra_tile;
          CB2_WAIT_CLK(ptx::smem_u32(&bars->acc_empty[acc]), pacc ^ 1u, w_acc);
          const uint32_t d_tmem = tmem_base + acc * p.bn;
          // Issue order (BF16C): a change of MMA kind (tf32 <-> f16) costs the tensor pipe ~135 cycles (measured,
          // tools/micro/mma_rate_pair.cu: 178 cycles per MMA in the alternating mix against 142 / 147 alone), so every
          // K-block starts with the kind the previous one ended with, and the tf32 fold MMA rides in front of the
          // accumulator's first tf32 group: one change per K-block instead of two.
          bool fold_pending = p.fold != 0;
          uint32_t acc_on   = 0u;     // 0 for the first MMA into this accumulator
          if (p.fold && !BF16C) {
            ptx::tc_fence_after();
            if (ptx::elect_one())
              ptx::mma_tf32_ss_2cta(d_tmem, ptx::umma_desc_sw32(fold_u32), ptx::umma_desc_sw32(fold_u32 + (1u + nt) * FOLD_TILE),
                                    idesc, 0u);
            __syncwarp();
            fold_pending = false;
            acc_on       = 1u;
          }
          for (int kbi = 0; kbi < p.kb; ++kbi, ++b_cnt) {
            const uint32_t sa = ra.slot, pa = ra.phase;
            ra.advance(p.a_slots);
            if (nt == 0 || p.a_stream) CB2_WAIT_CLK(ptx::smem_u32(&bars->a_ready[sa]), pa, w_a);
            uint32_t sb = rb.slot;
            const uint32_t pb = rb.phase;
            rb.advance(p.b_stages);
            if (p.b_resident) {
              sb = nt * p.kb + kbi;
              if (pt == pair) CB2_WAIT_CLK(ptx::smem_u32(&bars->b_full[sb]), 0u, w_b);
            } else {
              CB2_WAIT_CLK(ptx::smem_u32(&bars->b_full[sb]), pb, w_b);
            }
            ptx::tc_fence_after();
            const uint64_t da_hi = ptx::umma_desc_sw128(a_base + sa * A_SLOT_BYTES);
            const uint64_t da_lo = ptx::umma_desc_sw128(a_base + sa * A_SLOT_BYTES + KBLOCK_BYTES);
            const uint64_t db_hi = ptx::umma_desc_sw128(b_base + sb * b_stage_bytes);
            const uint64_t db_lo = ptx::umma_desc_sw128(b_base + sb * b_stage_bytes + b_half_bytes);
            const int nks = min(4, (p.d - kbi * KBLOCK + 7) / 8);
            const bool tf_first = last_kind == 0;
            if (BF16C) last_kind ^= 1u;   // every K-block issues both kinds, so it ends on the other one
            if (ptx::elect_one()) {
              if (BF16C) {
                // corrections first (bf16, K = 16), then the tf32 main term
                const uint32_t a_hb = a_base + sa * A_SLOT_BYTES + KBLOCK_BYTES;
                const uint32_t b_hb = b_base + sb * b_stage_bytes + b_half_bytes;
                const uint64_t da_hb = ptx::umma_desc_sw64(a_hb), da_lb = ptx::umma_desc_sw64(a_hb + KBLOCK_BYTES / 2);
                const uint64_t db_hb = ptx::umma_desc_sw64(b_hb), db_lb = ptx::umma_desc_sw64(b_hb + b_half_bytes / 2);
                const int nk16 = (nks + 1) / 2;
#pragma unroll
                for (int ph = 0; ph < 2; ++ph) {
                  if ((ph == 0) == tf_first) {   // tf32 main term (preceded by the fold MMA of a new accumulator)
                    if (fold_pending) {
                      ptx::mma_tf32_ss_2cta(d_tmem, ptx::umma_desc_sw32(fold_u32),
                                            ptx::umma_desc_sw32(fold_u32 + (1u + nt) * FOLD_TILE), idesc, acc_on);
                      acc_on = 1u;
                    }
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                      if (ks >= nks) break;
                      const uint64_t adv = static_cast<uint64_t>(ks * 2);
                      ptx::mma_tf32_ss_2cta(d_tmem, da_hi + adv, db_hi + adv, idesc, acc_on);
                      acc_on = 1u;
                    }
                  } else {                        // the two bf16 correction terms (K = 16)
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                      if (ks >= nk16) break;
                      const uint64_t adv = static_cast<uint64_t>(ks * 2);
                      ptx::mma_f16_ss_2cta(d_tmem, da_lb + adv, db_hb + adv, idesc16, acc_on);
                      ptx::mma_f16_ss_2cta(d_tmem, da_hb + adv, db_lb + adv, idesc16, 1u);
                      acc_on = 1u;
                    }
                  }
                }
              } else {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  if (ks >= nks) break;
                  const uint64_t adv = static_cast<uint64_t>(ks * 2);
                  ptx::mma_tf32_ss_2cta(d_tmem, da_lo + adv, db_hi + adv, idesc, acc_on);
                  ptx::mma_tf32_ss_2cta(d_tmem, da_hi + adv, db_lo + adv, idesc, 1u);
                  ptx::mma_tf32_ss_2cta(d_tmem, da_hi + adv, db_hi + adv, idesc, 1u);
                  acc_on = 1u;
                }
              }
              if (!p.b_resident) ptx::mma_commit_2cta(ptx::smem_u32(&bars->b_empty[sb]), 3);
              if (nt == p.k_tiles - 1 || p.a_stream) ptx::mma_commit_2cta(ptx::smem_u32(&bars->a_empty[sa]), 3);
            }
            __syncwarp();
            fold_pending = false;   // issued with the first tf32 group
            acc_on       = 1u;
          }
          ra_run = ra;
          if (ptx::elect_one()) ptx::mma_commit_2cta(ptx::smem_u32(&bars->acc_full[acc]), 3);
          __syncwarp();
        }
      }
      if constexpr (CLK) {
        if (p.dbg_clk && blockIdx.x == 0 && lane == 0) {
          p.dbg_clk[4] = w_acc; p.dbg_clk[5] = w_a; p.dbg_clk[6] = w_b; p.dbg_clk[7] = clock64() - t_start;
        }
      }
    }
  } else if (warp >= 8 && warp < 24) {
    // ===================== epilogue (own 128 rows of the pair tile) =====================
    const int64_t n_mine = (pair_tiles > pair) ? (pair_tiles - pair + n_pairs - 1) / n_pairs : 0;
    epilogue_role<true, DIST, false, CLK>(p, bars, cn_s, mrg_v, mrg_i, tmem_base, pair * 2 * TILE_M + cta_rank * TILE_M,
                        n_pairs * 2 * TILE_M, n_mine);
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // the peer may still be reading this CTA's shared memory / barriers
  ptx::tc_fence_after();
  if (warp == 1) ptx::tmem_dealloc_2cta(tmem_base, p.tmem_cols);
}

// Single-CTA twin of the CTA-pair kernel: the same roles, bf16 correction terms and folded half norms for k <= 128,
// with cta_group::1 instructions and local barriers, plus the row-owner epilogue (ROWOWN) for one centroid tile.  It
// replaced fused_l2_argmin_kernel as the E-step of these shapes in round 2 (profiles/r02_ab_table.txt).
template <bool BF16C, int DIST = 0, bool TRUNC = false, bool ROWOWN = false>
__global__ void __launch_bounds__(PAIR_THREADS, 1)
fused_l2_argmin_solo_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_hi,
                            const __grid_constant__ CUtensorMap tm_lo, const __grid_constant__ CUtensorMap tm_lb,
                            const __grid_constant__ CUtensorMap tm_cn, const FusedParams p)
{
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t raw_base = ptx::smem_u32(smem_dyn);
  const uint32_t base     = (raw_base + 1023u) & ~1023u;
  uint8_t* gbase          = smem_dyn + (base - raw_base);

  const uint32_t cta_rank = 0;
  const bool leader       = cta_rank == 0;
  const int64_t pair      = blockIdx.x;
  const int64_t n_pairs   = gridDim.x;
  const int half_n        = p.bn;                                       // centroid rows held by this CTA (all)

  const uint32_t b_half_bytes  = static_cast<uint32_t>(half_n) * 128u;   // hi (or lo) rows of this CTA
  const uint32_t b_stage_bytes = 2u * b_half_bytes;                     // hi then lo
  const uint32_t a_base  = base;
  const uint32_t b_base  = a_base + p.a_slots * A_SLOT_BYTES;
  // fold tiles (p.fold): ones [128 rows x 8 tf32] then one [half_n rows x 8 tf32] tile of half-norm pieces per
  // centroid tile, 32-byte rows, 4 KB each
  constexpr uint32_t FOLD_TILE = TILE_M * 32u;
  const uint32_t fold_off = p.a_slots * A_SLOT_BYTES + p.b_stages * b_stage_bytes;
  const uint32_t fold_u32 = base + fold_off;
  const uint32_t cn_off   = fold_off + (p.fold ? (1u + p.k_tiles) * FOLD_TILE : 0u);
  float* cn_s            = reinterpret_cast<float*>(gbase + cn_off);    // [2][bn]
  float* mrg_v           = cn_s + 2 * p.bn;                              // [128] epilogue half merge
  int* mrg_i             = reinterpret_cast<int*>(mrg_v + 3 * TILE_M);   // [3][128]
  Barriers* bars         = reinterpret_cast<Barriers*>(gbase + cn_off + 2u * p.bn * sizeof(float) + 6u * TILE_M * 4u);
  constexpr int NCONV    = 256;   // converter threads

  const int warp = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;

  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_A_SLOTS; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bars->a_raw_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->a_ready[s]), NCONV / 32);            // converter warps
      ptx::mbar_init(ptx::smem_u32(&bars->a_empty[s]), 1);                     // MMA commit
    }
    for (int s = 0; s < MAX_ACC; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bars->acc_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->acc_empty[s]), ROWOWN ? 4 : 16);  // epilogue warps per accumulator
    }
    for (int s = 0; s < MAX_STAGES; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bars->b_full[s]), 1);      // leader's copy: expect_tx covers both CTAs
      ptx::mbar_init(ptx::smem_u32(&bars->b_empty[s]), 1);
    }
    ptx::mbar_init(ptx::smem_u32(&bars->cn_full), 1);
    ptx::fence_barrier_init();
  }
  if (p.fold) {   // all-ones A tile (identical 16-byte chunks: the 32B swizzle does not matter)
    float* ones = reinterpret_cast<float*>(gbase + fold_off);
    for (int i = threadIdx.x; i < TILE_M * 8; i += blockDim.x) ones[i] = 1.0f;
    ptx::fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_x);
    ptx::prefetch_tmap(&tm_hi);
    ptx::prefetch_tmap(&tm_lo);
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(&bars->tmem_base), p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const int64_t pair_tiles = p.m_tiles;             // 128-row tiles

  if (warp == 0) {
    // ===================== X producer (own 128 rows) =====================
    {
      uint32_t a_cnt = 0;
      Ring ra;
      const int a_reps = p.a_stream ? p.k_tiles : 1;
      for (int64_t pt = pair; pt < pair_tiles; pt += n_pairs) {
        const int32_t row0 = static_cast<int32_t>(pt * TILE_M);
        // L2 prefetch of the row tile p.l2_ahead rounds ahead: the shared-memory ring only holds ~2 row tiles, too
        // few to cover the ~1.5 us HBM latency; with the tile already in L2 the ring turns around in time
        if (p.l2_ahead > 0) {
          const int64_t pf = pt + static_cast<int64_t>(p.l2_ahead) * n_pairs;
          if (pf < pair_tiles && ptx::elect_one()) {
            const int32_t prow = static_cast<int32_t>(pf * TILE_M);
            for (int kbi = 0; kbi < p.kb; ++kbi) ptx::tma_prefetch_l2_2d(&tm_x, kbi * KBLOCK, prow);
          }
          __syncwarp();
        }
        for (int rep = 0; rep < a_reps; ++rep)
        for (int kbi = 0; kbi < p.kb; ++kbi, ++a_cnt) {
          const uint32_t sa = ra.slot, pa = ra.phase;
          ra.advance(p.a_slots);
          ptx::mbar_wait_park(ptx::smem_u32(&bars->a_empty[sa]), pa ^ 1u);
          if (ptx::elect_one()) {
            const uint32_t full = ptx::smem_u32(&bars->a_raw_full[sa]);
            ptx::mbar_arrive_expect_tx(full, KBLOCK_BYTES);
            ptx::tma_load_2d_hint(a_base + sa * A_SLOT_BYTES, &tm_x, kbi * KBLOCK, row0, full, ptx::kEvictFirst);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 2) {
    // ===================== centroid producer (own half of every block) =====================
    {
      uint32_t b_cnt = 0;
      Ring rb;
      if (p.fold && pair < pair_tiles) {
        // half-norm pieces of every centroid tile (this CTA's half of the rows), resident for the whole kernel
        if (ptx::elect_one()) {
          const uint32_t full_local  = ptx::smem_u32(&bars->cn_full);
          const uint32_t full_leader = full_local;
          ptx::mbar_arrive_expect_tx(full_local, static_cast<uint32_t>(p.k_tiles) * static_cast<uint32_t>(half_n) * 32u);
          for (int nt = 0; nt < p.k_tiles; ++nt)
            ptx::tma_load_2d_hint(fold_u32 + (1u + nt) * FOLD_TILE, &tm_cn, 0,
                                  nt * p.bn + static_cast<int32_t>(cta_rank) * half_n, full_leader, ptx::kEvictLast);
        }
        __syncwarp();
      }
      for (int64_t pt = pair; pt < pair_tiles; pt += n_pairs) {
        if (p.b_resident && pt != pair) break;
        for (int nt = 0; nt < p.k_tiles; ++nt) {
          for (int kbi = 0; kbi < p.kb; ++kbi, ++b_cnt) {
            const uint32_t sb = rb.slot, pb = rb.phase;
            rb.advance(p.b_stages);
            ptx::mbar_wait_park(ptx::smem_u32(&bars->b_empty[sb]), pb ^ 1u);
            if (ptx::elect_one()) {
              const uint32_t full_local  = ptx::smem_u32(&bars->b_full[sb]);
              const uint32_t full_leader = full_local;
              ptx::mbar_arrive_expect_tx(full_local, b_stage_bytes);
              const uint32_t dst = b_base + sb * b_stage_bytes;
              const int32_t crow = nt * p.bn + static_cast<int32_t>(cta_rank) * half_n;
              ptx::tma_load_2d_hint(dst, &tm_hi, kbi * KBLOCK, crow, full_leader, ptx::kEvictLast);
              if (BF16C) {   // tm_lo = bf16 hi, tm_lb = bf16 lo: two half-size tiles
                ptx::tma_load_2d_hint(dst + b_half_bytes, &tm_lo, kbi * KBLOCK, crow, full_leader, ptx::kEvictLast);
                ptx::tma_load_2d_hint(dst + b_half_bytes + b_half_bytes / 2, &tm_lb, kbi * KBLOCK, crow, full_leader,
                                      ptx::kEvictLast);
              } else {
                ptx::tma_load_2d_hint(dst + b_half_bytes, &tm_lo, kbi * KBLOCK, crow, full_leader, ptx::kEvictLast);
              }
            }
            __syncwarp();
          }
        }
      }
    }
  } else if ((warp >= 4 && warp < 8) || warp >= 24) {
    // ===================== converter (8 warps: 4..7 and 24..27) =====================
    const int ct = warp >= 24 ? threadIdx.x - 768 + 128 : threadIdx.x - 128;   // 0..NCONV-1
    // this thread's chunks are ct + NCONV i: NCONV / 8 rows apart, so the logical chunk and the swizzle phases are fixed
    const int conv_row       = ct >> 3;
    const int conv_lc        = (ct & 7) ^ (conv_row & 7);     // logical 16-byte chunk: features [4 lc, 4 lc + 4)
    const uint32_t conv_off  = static_cast<uint32_t>(conv_row) * 64u +
                               ((static_cast<uint32_t>(conv_lc >> 1) ^ ((conv_row >> 1) & 3u)) << 4) +
                               (static_cast<uint32_t>(conv_lc & 1) << 3);
    uint32_t a_cnt = 0;
      Ring ra;
    const int a_reps = p.a_stream ? p.k_tiles : 1;
    for (int64_t pt = pair; pt < pair_tiles; pt += n_pairs) {
      for (int rep = 0; rep < a_reps; ++rep)
      for (int kbi = 0; kbi < p.kb; ++kbi, ++a_cnt) {
        const uint32_t sa = ra.slot, pa = ra.phase;
        ra.advance(p.a_slots);
        ptx::mbar_wait_park(ptx::smem_u32(&bars->a_raw_full[sa]), pa);
        uint4* hi = reinterpret_cast<uint4*>(gbase + sa * A_SLOT_BYTES);
        uint4* lo = reinterpret_cast<uint4*>(gbase + sa * A_SLOT_BYTES + KBLOCK_BYTES);
        const int rem_f       = p.d - kbi * KBLOCK;
        const int live_chunks = BF16C ? 4 * min(2, (rem_f + 15) / 16) : 2 * min(4, (rem_f + 7) / 8);
        uint8_t* hb = gbase + sa * A_SLOT_BYTES + KBLOCK_BYTES;                    // bf16 hi tile (8 KB)
        uint8_t* lb = hb + KBLOCK_BYTES / 2;                                       // bf16 lo tile (8 KB)
        constexpr int CPT = KBLOCK_BYTES / 16 / NCONV;   // 16-byte chunks per thread
        // all loads first: the chunks are independent, one shared-memory latency instead of CPT
        uint4 v[CPT];
#pragma unroll
        for (int i = 0; i < CPT; ++i) v[i] = hi[ct + i * NCONV];
        if (conv_lc < live_chunks) {   // chunk ct + NCONV i keeps the logical chunk and the swizzle phase of chunk ct
#pragma unroll
          for (int i = 0; i < CPT; ++i) {
            const int e = ct + i * NCONV;
            uint4 h, l;
            if (BF16C && !TRUNC) {   // nearest tf32 (ties away from zero): |lo| <= 2^-12 |x|
              h.x = (v[i].x + 0x1000u) & 0xffffe000u;
              h.y = (v[i].y + 0x1000u) & 0xffffe000u;
              h.z = (v[i].z + 0x1000u) & 0xffffe000u;
              h.w = (v[i].w + 0x1000u) & 0xffffe000u;
            } else {
              h.x = v[i].x & 0xffffe000u; h.y = v[i].y & 0xffffe000u; h.z = v[i].z & 0xffffe000u; h.w = v[i].w & 0xffffe000u;
            }
            l.x = __float_as_uint(__uint_as_float(v[i].x) - __uint_as_float(h.x));
            l.y = __float_as_uint(__uint_as_float(v[i].y) - __uint_as_float(h.y));
            l.z = __float_as_uint(__uint_as_float(v[i].z) - __uint_as_float(h.z));
            l.w = __float_as_uint(__uint_as_float(v[i].w) - __uint_as_float(h.w));
            if (!TRUNC) hi[e] = h;
            if (BF16C) {
              // 4 features -> 8 bytes of the 64-byte bf16 row (64B swizzle: 16-byte chunk ^= (row / 2) % 4)
              const uint32_t off = conv_off + static_cast<uint32_t>(i) * (static_cast<uint32_t>(NCONV / 8) * 64u);
              const __nv_bfloat162 h01 = __floats2bfloat162_rn(__uint_as_float(h.x), __uint_as_float(h.y));
              const __nv_bfloat162 h23 = __floats2bfloat162_rn(__uint_as_float(h.z), __uint_as_float(h.w));
              const __nv_bfloat162 l01 = __floats2bfloat162_rn(__uint_as_float(l.x), __uint_as_float(l.y));
              const __nv_bfloat162 l23 = __floats2bfloat162_rn(__uint_as_float(l.z), __uint_as_float(l.w));
              *reinterpret_cast<uint2*>(hb + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
              *reinterpret_cast<uint2*>(lb + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
            } else {
              lo[e] = l;
            }
          }
        }
        ptx::fence_proxy_async_smem();  // generic-proxy writes to this CTA's operand tiles -> async proxy (pair MMA)
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->a_ready[sa]));
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pair leader only; warp-uniform loop, one elected lane issues) ====
    if (leader) {
      const uint32_t idesc   = ptx::umma_idesc_tf32(TILE_M, p.bn);
      const uint32_t idesc16 = ptx::umma_idesc_bf16(TILE_M, p.bn);
      (void)idesc16;
      uint32_t b_cnt = 0, acc_cnt = 0;
      Ring ra_tile, ra_run, rb, racc;
      if (p.fold && pair < pair_tiles) ptx::mbar_wait_park(ptx::smem_u32(&bars->cn_full), 0u);
      uint32_t last_kind = 0;   // kind of the last MMA issued: 0 tf32, 1 f16 (issue order as in the pair kernel)
      for (int64_t pt = pair; pt < pair_tiles; pt += n_pairs, ra_tile.advance_by(p.kb, p.a_slots)) {
        for (int nt = 0; nt < p.k_tiles; ++nt, ++acc_cnt) {
          const uint32_t acc = racc.slot, pacc = racc.phase;
          racc.advance(p.n_acc);
          Ring ra = p.a_stream ? ra_run : ra_tile;
          ptx::mbar_wait_park(ptx::smem_u32(&bars->acc_empty[acc]), pacc ^ 1u);
          const uint32_t d_tmem = tmem_base + acc * p.bn;
          bool fold_pending = p.fold != 0;
          uint32_t acc_on   = 0u;     // 0 for the first MMA into this accumulator
          if (p.fold && !BF16C) {
            ptx::tc_fence_after();
            if (ptx::elect_one())
              ptx::mma_tf32_ss(d_tmem, ptx::umma_desc_sw32(fold_u32), ptx::umma_desc_sw32(fold_u32 + (1u + nt) * FOLD_TILE),
                                    idesc, 0u);
            __syncwarp();
            fold_pending = false;
            acc_on       = 1u;
          }
          for (int kbi = 0; kbi < p.kb; ++kbi, ++b_cnt) {
            const uint32_t sa = ra.slot, pa = ra.phase;
            ra.advance(p.a_slots);
            if (nt == 0 || p.a_stream) ptx::mbar_wait_park(ptx::smem_u32(&bars->a_ready[sa]), pa);
            uint32_t sb = rb.slot;
            const uint32_t pb = rb.phase;
            rb.advance(p.b_stages);
            if (p.b_resident) {
              sb = nt * p.kb + kbi;
              if (pt == pair) ptx::mbar_wait_park(ptx::smem_u32(&bars->b_full[sb]), 0u);
            } else {
              ptx::mbar_wait_park(ptx::smem_u32(&bars->b_full[sb]), pb);
            }
            ptx::tc_fence_after();
            const uint64_t da_hi = ptx::umma_desc_sw128(a_base + sa * A_SLOT_BYTES);
            const uint64_t da_lo = ptx::umma_desc_sw128(a_base + sa * A_SLOT_BYTES + KBLOCK_BYTES);
            const uint64_t db_hi = ptx::umma_desc_sw128(b_base + sb * b_stage_bytes);
            const uint64_t db_lo = ptx::umma_desc_sw128(b_base + sb * b_stage_bytes + b_half_bytes);
            const int nks = min(4, (p.d - kbi * KBLOCK + 7) / 8);
            const bool tf_first = last_kind == 0;
            if (BF16C) last_kind ^= 1u;
            if (ptx::elect_one()) {
              if (BF16C) {
                // corrections first (bf16, K = 16), then the tf32 main term
                const uint32_t a_hb = a_base + sa * A_SLOT_BYTES + KBLOCK_BYTES;
                const uint32_t b_hb = b_base + sb * b_stage_bytes + b_half_bytes;
                const uint64_t da_hb = ptx::umma_desc_sw64(a_hb), da_lb = ptx::umma_desc_sw64(a_hb + KBLOCK_BYTES / 2);
                const uint64_t db_hb = ptx::umma_desc_sw64(b_hb), db_lb = ptx::umma_desc_sw64(b_hb + b_half_bytes / 2);
                const int nk16 = (nks + 1) / 2;
#pragma unroll
                for (int ph = 0; ph < 2; ++ph) {
                  if ((ph == 0) == tf_first) {   // tf32 main term (preceded by the fold MMA of a new accumulator)
                    if (fold_pending) {
                      ptx::mma_tf32_ss(d_tmem, ptx::umma_desc_sw32(fold_u32), ptx::umma_desc_sw32(fold_u32 + (1u + nt) * FOLD_TILE),
                                       idesc, acc_on);
                      acc_on = 1u;
                    }
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                      if (ks >= nks) break;
                      const uint64_t adv = static_cast<uint64_t>(ks * 2);
                      ptx::mma_tf32_ss(d_tmem, da_hi + adv, db_hi + adv, idesc, acc_on);
                      acc_on = 1u;
                    }
                  } else {                        // the two bf16 correction terms (K = 16)
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                      if (ks >= nk16) break;
                      const uint64_t adv = static_cast<uint64_t>(ks * 2);
                      ptx::mma_f16_ss(d_tmem, da_lb + adv, db_hb + adv, idesc16, acc_on);
                      ptx::mma_f16_ss(d_tmem, da_hb + adv, db_lb + adv, idesc16, 1u);
                      acc_on = 1u;
                    }
                  }
                }
              } else {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  if (ks >= nks) break;
                  const uint64_t adv = static_cast<uint64_t>(ks * 2);
                  ptx::mma_tf32_ss(d_tmem, da_lo + adv, db_hi + adv, idesc, acc_on);
                  ptx::mma_tf32_ss(d_tmem, da_hi + adv, db_lo + adv, idesc, 1u);
                  ptx::mma_tf32_ss(d_tmem, da_hi + adv, db_hi + adv, idesc, 1u);
                  acc_on = 1u;
                }
              }
              if (!p.b_resident) ptx::mma_commit(ptx::smem_u32(&bars->b_empty[sb]));
              if (nt == p.k_tiles - 1 || p.a_stream) ptx::mma_commit(ptx::smem_u32(&bars->a_empty[sa]));
            }
            __syncwarp();
            fold_pending = false;
            acc_on       = 1u;
          }
          ra_run = ra;
          if (ptx::elect_one()) ptx::mma_commit(ptx::smem_u32(&bars->acc_full[acc]));
          __syncwarp();
        }
      }
    }
  } else if (warp >= 8 && warp < 24) {
    // ===================== epilogue (own 128 rows of the pair tile) =====================
    const int64_t n_mine = (pair_tiles > pair) ? (pair_tiles - pair + n_pairs - 1) / n_pairs : 0;
    if constexpr (ROWOWN)
      epilogue_role_rowown<true>(p, bars, cn_s, tmem_base, pair * TILE_M, n_pairs * TILE_M, n_mine);
    else
      epilogue_role<false, DIST, true>(p, bars, cn_s, mrg_v, mrg_i, tmem_base, pair * TILE_M, n_pairs * TILE_M, n_mine);
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, p.tmem_cols);
}

// =================================================================================================
// Row-packed small-d kernel with the X operand in tensor memory (n_features = 16: two data rows per operand row,
// block-diagonal centroid operands, one accumulator tile of BN = 2 k_sub <= 128 columns, centroids resident).
// Why: the shared-memory twin above is bound by shared-memory bandwidth at this shape (profiles/
// r02_role_skip_experiments.txt) -- per 256-data-row tile its nine MMAs fetch 72 KB of operands and the converter
// writes 16 KB of bf16 tiles next to the 16 KB TMA writes and the 16 KB converter reads.  Here the converter (thread =
// operand row, one warp per TMEM lane quarter) writes the raw fp32 words (the tensor core truncates them to tf32) and
// the bf16 hi / lo pairs of its row straight into 64 TMEM columns (tcgen05.st), the eight data MMAs read A from
// tensor memory (N/2-cycle floor instead of 43 + N/2, tools/micro/mma_rate.cu) and only B (4 KB each) from shared
// memory; the -1/2||c||^2 fold stays a shared-memory MMA of the ones tile.  The raw X tiles ride a deep ring of 16 KB
// slots.  TMEM: n_acc accumulators of BN columns, then two 64-column operand stages [raw 32 | bf16 hi 16 | bf16 lo 16].
// MSTEP: the fused M-step (accumulate warps 20..27 add the raw tile into private tables once the
// row-owner epilogue has published the tile's labels; the slot's empty barrier counts converter and accumulate warps).
// In MSTEP mode idle parameter fields carry the outputs: dbg_dots -> partial_S, cnh -> partial_W, a_stream -> true k.
template <bool MSTEP>
__global__ void __launch_bounds__(PAIR_THREADS, 1)
fused_l2_argmin_tsp_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_hi,
                           const __grid_constant__ CUtensorMap tm_hb, const __grid_constant__ CUtensorMap tm_lb,
                           const __grid_constant__ CUtensorMap tm_cn, const FusedParams p)
{
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t raw_base = ptx::smem_u32(smem_dyn);
  const uint32_t base     = (raw_base + 1023u) & ~1023u;
  uint8_t* gbase          = smem_dyn + (base - raw_base);

  const int64_t cta    = blockIdx.x;
  const int64_t n_ctas = gridDim.x;
  const int64_t tiles  = p.m_tiles;                                      // tiles of 128 operand rows (256 data rows)

  const uint32_t x_base   = base;                                        // raw ring, 16 KB slots
  const uint32_t b_off    = static_cast<uint32_t>(p.raw_slots) * KBLOCK_BYTES;
  const uint32_t b_base   = base + b_off;                                // tf32 hi [bn x 128 B] | bf16 hi [bn x 64 B] | bf16 lo
  const uint32_t b_hi_sz  = static_cast<uint32_t>(p.bn) * 128u;
  constexpr uint32_t FOLD_TILE = TILE_M * 32u;                           // ones [128 x 8 tf32], then the half-norm pieces
  const uint32_t fold_off = b_off + 2u * b_hi_sz;
  const uint32_t fold_u32 = base + fold_off;
  Barriers* bars          = reinterpret_cast<Barriers*>(gbase + fold_off + 2u * FOLD_TILE);
  float* ms_tab           = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + ((sizeof(Barriers) + 127) & ~size_t(127)));
  const int ms_k          = p.pack > 1 ? p.k_sub : p.bn;   // accumulator columns per data row = table rows
  const int ms_rows       = ms_k + 1;      // private table rows per accumulate warp: the clusters + one dummy row
  int* ms_counts          = reinterpret_cast<int*>(ms_tab + static_cast<size_t>(TSP_ACC_WARPS) * ms_rows * 32);
  uint32_t* ms_labels     = reinterpret_cast<uint32_t*>(ms_counts + 2 * ms_k);

  const int warp = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;

  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_RAW; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bars->raw_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->raw_empty[s]), MSTEP ? 4 + TSP_ACC_WARPS : 4);   // converter (+ accumulate) warps
    }
    for (int s = 0; s < MAX_A_SLOTS; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bars->a_ready[s]), 4);                  // the 4 converter warps
      ptx::mbar_init(ptx::smem_u32(&bars->a_empty[s]), 1);                  // MMA commit
      ptx::mbar_init(ptx::smem_u32(&bars->lab_full[s]), 4);                 // the 4 epilogue warps of a tile's group
    }
    for (int s = 0; s < MAX_ACC; ++s) {
      ptx::mbar_init(ptx::smem_u32(&bars->acc_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&bars->acc_empty[s]), 4);               // row-owner epilogue: 4 warps per tile
    }
    ptx::mbar_init(ptx::smem_u32(&bars->b_full[0]), 1);
    ptx::mbar_init(ptx::smem_u32(&bars->cn_full), 1);
    ptx::fence_barrier_init();
  }
  {   // all-ones A tile of the fold MMA (identical 16-byte chunks: the 32B swizzle does not matter)
    float* ones = reinterpret_cast<float*>(gbase + fold_off);
    for (int i = threadIdx.x; i < TILE_M * 8; i += blockDim.x) ones[i] = 1.0f;
    ptx::fence_proxy_async_smem();
  }
  if (MSTEP) {
    for (int i = threadIdx.x; i < TSP_ACC_WARPS * ms_rows * 32 + 2 * ms_k; i += blockDim.x) ms_tab[i] = 0.0f;   // tables + counts
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_x);
    ptx::prefetch_tmap(&tm_hi);
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(&bars->tmem_base), p.tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================== raw X producer =====================
    Ring rr;
    for (int64_t t = cta; t < tiles; t += n_ctas) {
      const int32_t row0 = static_cast<int32_t>(t * TILE_M);
      const uint32_t rs = rr.slot, rp = rr.phase;
      rr.advance(p.raw_slots);
      ptx::mbar_wait_park(ptx::smem_u32(&bars->raw_empty[rs]), rp ^ 1u);
      if (ptx::elect_one()) {
        const uint32_t full = ptx::smem_u32(&bars->raw_full[rs]);
        ptx::mbar_arrive_expect_tx(full, KBLOCK_BYTES);
        ptx::tma_load_2d_hint(x_base + rs * KBLOCK_BYTES, &tm_x, 0, row0, full, ptx::kEvictFirst);
      }
      __syncwarp();
    }
  } else if (warp == 2) {
    // ===================== centroid operands + half-norm pieces: once, resident =====================
    if (cta < tiles && ptx::elect_one()) {
      const uint32_t bfull = ptx::smem_u32(&bars->b_full[0]);
      ptx::mbar_arrive_expect_tx(bfull, 2u * b_hi_sz);
      ptx::tma_load_2d_hint(b_base, &tm_hi, 0, 0, bfull, ptx::kEvictLast);
      ptx::tma_load_2d_hint(b_base + b_hi_sz, &tm_hb, 0, 0, bfull, ptx::kEvictLast);
      ptx::tma_load_2d_hint(b_base + b_hi_sz + b_hi_sz / 2, &tm_lb, 0, 0, bfull, ptx::kEvictLast);
      const uint32_t cfull = ptx::smem_u32(&bars->cn_full);
      ptx::mbar_arrive_expect_tx(cfull, static_cast<uint32_t>(p.bn) * 32u);
      ptx::tma_load_2d_hint(fold_u32 + FOLD_TILE, &tm_cn, 0, 0, cfull, ptx::kEvictLast);
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 8) {
    // ===================== converter: operand row (shared) -> raw | bf16 hi | bf16 lo columns (tensor memory) ======
    const int quarter        = warp & 3;
    const int row            = quarter * 32 + lane;                     // operand row of the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    Ring rr, ra;
    for (int64_t t = cta; t < tiles; t += n_ctas) {
      const uint32_t rs = rr.slot, rp = rr.phase;
      rr.advance(p.raw_slots);
      const uint32_t as = ra.slot, ap = ra.phase;
      ra.advance(p.a_slots);
      ptx::mbar_wait_park(ptx::smem_u32(&bars->raw_full[rs]), rp);
      ptx::mbar_wait_park(ptx::smem_u32(&bars->a_empty[as]), ap ^ 1u);   // the MMAs that read this operand stage retired
      ptx::tc_fence_after();
      const uint4* src    = reinterpret_cast<const uint4*>(gbase + rs * KBLOCK_BYTES + row * 128);
      const uint32_t acol = tmem_base + lane_addr + p.a_col0 + as * 64;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {   // 16 features (one data row of the packed pair) at a time: registers
        uint32_t w[16], hb[8], lb[8];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 v = src[(hf * 4 + c) ^ (row & 7)];                  // 128B swizzle: chunk ^= row & 7
          w[c * 4 + 0] = v.x; w[c * 4 + 1] = v.y; w[c * 4 + 2] = v.z; w[c * 4 + 3] = v.w;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float x0 = __uint_as_float(w[2 * e]), x1 = __uint_as_float(w[2 * e + 1]);
          const float h0 = __uint_as_float(w[2 * e] & 0xffffe000u), h1 = __uint_as_float(w[2 * e + 1] & 0xffffe000u);
          const __nv_bfloat162 hh = __floats2bfloat162_rn(h0, h1);        // .x (low half) = the lower k
          const __nv_bfloat162 ll = __floats2bfloat162_rn(x0 - h0, x1 - h1);
          hb[e] = *reinterpret_cast<const uint32_t*>(&hh);
          lb[e] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        ptx::tmem_st_32x16(acol + hf * 16, w);
        ptx::tmem_st_32x8(acol + 32 + hf * 8, hb);
        ptx::tmem_st_32x8(acol + 48 + hf * 8, lb);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(ptx::smem_u32(&bars->raw_empty[rs]));           // converter is done with the raw slot
        ptx::mbar_arrive(ptx::smem_u32(&bars->a_ready[as]));
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    const uint32_t idesc   = ptx::umma_idesc_tf32(TILE_M, p.bn);
    const uint32_t idesc16 = ptx::umma_idesc_bf16(TILE_M, p.bn);
    const uint64_t db_hi   = ptx::umma_desc_sw128(b_base);
    const uint64_t db_hb   = ptx::umma_desc_sw64(b_base + b_hi_sz);
    const uint64_t db_lb   = ptx::umma_desc_sw64(b_base + b_hi_sz + b_hi_sz / 2);
    const uint64_t d_ones  = ptx::umma_desc_sw32(fold_u32);
    const uint64_t d_cn    = ptx::umma_desc_sw32(fold_u32 + FOLD_TILE);
    Ring ra, racc;
    bool tf_first = true;   // a change of MMA kind costs the tensor pipe: every tile starts with the kind the last one ended on
    if (cta < tiles) {
      ptx::mbar_wait_park(ptx::smem_u32(&bars->b_full[0]), 0u);
      ptx::mbar_wait_park(ptx::smem_u32(&bars->cn_full), 0u);
    }
    for (int64_t t = cta; t < tiles; t += n_ctas, tf_first = !tf_first) {
      const uint32_t acc = racc.slot, pacc = racc.phase;
      racc.advance(p.n_acc);
      const uint32_t as = ra.slot, ap = ra.phase;
      ra.advance(p.a_slots);
      ptx::mbar_wait_park(ptx::smem_u32(&bars->acc_empty[acc]), pacc ^ 1u);
      ptx::mbar_wait_park(ptx::smem_u32(&bars->a_ready[as]), ap);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * p.bn;
      const uint32_t a_raw  = tmem_base + p.a_col0 + as * 64;
      const uint32_t a_hb   = a_raw + 32, a_lb = a_raw + 48;
      if (ptx::elect_one()) {
        uint32_t acc_on = 0u;
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
          if ((ph == 0) == tf_first) {   // fold + tf32 main term
            ptx::mma_tf32_ss(d_tmem, d_ones, d_cn, idesc, acc_on);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              ptx::mma_tf32_ts(d_tmem, a_raw + ks * 8, db_hi + static_cast<uint64_t>(ks * 2), idesc, 1u);
            acc_on = 1u;
          } else {                        // the two bf16 correction terms (K = 16: 8 TMEM columns of A, 32 bytes of B)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              ptx::mma_f16_ts(d_tmem, a_lb + ks * 8, db_hb + static_cast<uint64_t>(ks * 2), idesc16, acc_on);
              ptx::mma_f16_ts(d_tmem, a_hb + ks * 8, db_lb + static_cast<uint64_t>(ks * 2), idesc16, 1u);
              acc_on = 1u;
            }
          }
        }
        ptx::mma_commit(ptx::smem_u32(&bars->a_empty[as]));
        ptx::mma_commit(ptx::smem_u32(&bars->acc_full[acc]));
      }
      __syncwarp();
    }
  } else if (MSTEP && warp >= 20) {
    // ===================== fused M-step: accumulate warps 20..27 (16 operand rows of every tile each) ===============
    const int aw         = warp - 20;
    // lane = (data row of the packed pair, column) -- or simply the column of an unpacked row: lane l only ever touches
    // bank l of its warp's private table
    uint8_t* tab         = reinterpret_cast<uint8_t*>(ms_tab + static_cast<size_t>(aw) * ms_rows * 32 + lane);
    // packed rows: the operand row is [data row 2r (d floats) | data row 2r+1 (d floats) | zero fill]; lanes past 2 d add
    // the zero fill into columns nobody reads
    const int d_row      = p.pack > 1 ? p.d / 2 : p.d;
    const uint32_t sel   = (p.pack > 1 && lane >= d_row && lane < 2 * d_row) ? 0x4432u : 0x4410u;   // this half's 16 bits
    const uint32_t xch   = static_cast<uint32_t>(lane >> 2);                       // logical 16-byte chunk of the lane's float
    const uint32_t xin   = static_cast<uint32_t>(lane & 3) * 4u;
    const uint32_t dummy = static_cast<uint32_t>(ms_k) * 128u;                     // byte offset of the row nobody reads
    Ring rr;
    for (int64_t t = cta; t < tiles; t += n_ctas) {
      const uint32_t rs = rr.slot, rp = rr.phase;
      rr.advance(p.raw_slots);
      ptx::mbar_wait_park(ptx::smem_u32(&bars->raw_full[rs]), rp);     // the raw tile (async-proxy writes) is visible
      ptx::mbar_wait_park(ptx::smem_u32(&bars->lab_full[rs]), rp);     // ... and so are its labels
      const uint32_t* lw = ms_labels + rs * TILE_M + aw * 16;          // (offset of row 2r's cluster | row 2r+1's << 16)
      const uint8_t* xs  = gbase + rs * KBLOCK_BYTES + static_cast<uint32_t>(aw * 16) * 128u;
      // Two operand rows per step, branch-free: when both rows (in this lane's half) belong to one cluster the second is
      // folded into the first in registers and its own update is pointed at the dummy row, so the two table
      // read-modify-writes of a step never alias (loads, adds, stores; ~10 instructions per row).  The kernel is bound
      // by instruction issue, so the count matters more than the latency of the chain.
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const uint32_t l0 = __byte_perm(lw[j], 0u, sel), w1 = __byte_perm(lw[j + 1], 0u, sel);   // broadcast loads
        const uint32_t r0 = static_cast<uint32_t>(j), r1 = r0 + 1u;
        const float x0 = *reinterpret_cast<const float*>(xs + r0 * 128u + ((xch ^ (r0 & 7u)) << 4) + xin);
        const float x1 = *reinterpret_cast<const float*>(xs + r1 * 128u + ((xch ^ (r1 & 7u)) << 4) + xin);
        const bool dup    = w1 == l0;
        const float a0    = x0 + (dup ? x1 : 0.0f);
        const uint32_t l1 = dup ? dummy : w1;
        float* p0 = reinterpret_cast<float*>(tab + l0);
        float* p1 = reinterpret_cast<float*>(tab + l1);
        const float t0 = *p0, t1 = *p1;
        *p0 = t0 + a0;
        *p1 = t1 + x1;
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bars->raw_empty[rs]));
    }
  } else if (warp >= 8 && warp < 24) {
    // ===================== row-owner epilogue =====================
    const int64_t n_mine = (tiles > cta) ? (tiles - cta + n_ctas - 1) / n_ctas : 0;
    epilogue_role_rowown<true, MSTEP>(p, bars, nullptr, tmem_base, cta * TILE_M, n_ctas * TILE_M, n_mine, ms_labels, ms_counts,
                                      p.raw_slots, 7);   // label words carry byte offsets of 128-byte table rows
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, p.tmem_cols);
  if (MSTEP) {
    // fold the private tables in a fixed order (accumulate warp 0..7, packed group 0 then 1) into this CTA's partials
    const int k_true = p.a_stream;
    const int d_true = p.pack > 1 ? p.d / 2 : p.d;
    float* out_S     = p.dbg_dots + static_cast<size_t>(blockIdx.x) * k_true * d_true;
    float* out_W     = const_cast<float*>(p.cnh) + static_cast<size_t>(blockIdx.x) * k_true;
    for (int e = threadIdx.x; e < k_true * d_true; e += blockDim.x) {
      const int j = e / d_true, c = e - j * d_true;
      float acc = 0.0f;
#pragma unroll
      for (int w = 0; w < TSP_ACC_WARPS; ++w) {
        const float* tw = ms_tab + (static_cast<size_t>(w) * ms_rows + j) * 32;
        acc += tw[c];
        if (p.pack > 1) acc += tw[d_true + c];
      }
      out_S[e] = acc;
    }
    for (int j = threadIdx.x; j < k_true; j += blockDim.x) out_W[j] = static_cast<float>(ms_counts[j]);
  }
}

// hi/lo split + half norms of the centroids into padded operand buffers
__global__ void prepare_centroids_kernel(const float* __restrict__ C, int k, int d, int k_pad, int d_pad,
                                         float* __restrict__ hi, float* __restrict__ lo, float* __restrict__ cnh,
                                         __nv_bfloat16* __restrict__ hb, __nv_bfloat16* __restrict__ lb,
                                         float* __restrict__ cnp)
{
  const int j    = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  if (j >= k_pad) return;
  double s = 0.0;
  for (int c = lane; c < d_pad; c += 32) {
    float v = (j < k && c < d) ? C[static_cast<int64_t>(j) * d + c] : 0.0f;
    const uint32_t u = __float_as_uint(v);
    float h = __uint_as_float(hb ? ((u + 0x1000u) & 0xffffe000u) : (u & 0xffffe000u));
    hi[static_cast<int64_t>(j) * d_pad + c] = h;
    lo[static_cast<int64_t>(j) * d_pad + c] = v - h;
    if (hb) {
      hb[static_cast<int64_t>(j) * d_pad + c] = __float2bfloat16_rn(h);
      lb[static_cast<int64_t>(j) * d_pad + c] = __float2bfloat16_rn(v - h);
    }
    s += static_cast<double>(v) * static_cast<double>(v);
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) cnh[j] = (j < k) ? static_cast<float>(0.5 * s) : __int_as_float(0x7f800000);
  if (cnp && lane < 8) {
    // -1/2||c||^2 as three pieces that are exact in tf32 (11 + 11 + 2 mantissa bits); padding rows can never win
    const float m  = (j < k) ? -static_cast<float>(0.5 * s) : -3.0e38f;
    const float p1 = __uint_as_float(__float_as_uint(m) & 0xffffe000u);
    const float r1 = m - p1;
    const float p2 = __uint_as_float(__float_as_uint(r1) & 0xffffe000u);
    const float p3 = r1 - p2;
    cnp[static_cast<int64_t>(j) * 8 + lane] = lane == 0 ? p1 : lane == 1 ? p2 : lane == 2 ? p3 : 0.0f;
  }
}

// Row-packed operands for n_features <= 16: two consecutive rows of X are read as ONE 128-byte operand
// row [x_2r | x_2r+1] (X is simply viewed as [n/2, 2d]), and the centroid operand becomes block-diagonal:
// B'[g*k_sub + j] = c_j placed at columns [g*d, g*d + d).  One 128-row MMA tile then covers 256 data rows with a
// fully used K-block (no out-of-bounds half), halving the per-row cost of the tile hand-offs that bound the
// small-d regime.  Accumulator columns [g*k_sub, (g+1)*k_sub) hold x_{2r+g} . c_j.
__global__ void prepare_centroids_packed_kernel(const float* __restrict__ C, int k, int d, int k_sub,
                                                float* __restrict__ hi, float* __restrict__ lo,
                                                float* __restrict__ cnh)
{
  const int jj   = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;   // row of B' in [0, 2*k_sub)
  const int lane = threadIdx.x % 32;
  if (jj >= 2 * k_sub) return;
  const int g = jj / k_sub, j = jj % k_sub;
  const int c = lane - g * d;                                          // feature index held by this column
  const float v = (j < k && c >= 0 && c < d) ? C[static_cast<int64_t>(j) * d + c] : 0.0f;
  const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  hi[static_cast<int64_t>(jj) * KBLOCK + lane] = h;
  lo[static_cast<int64_t>(jj) * KBLOCK + lane] = v - h;
  double s = static_cast<double>(v) * static_cast<double>(v);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) cnh[jj] = (j < k) ? static_cast<float>(0.5 * s) : __int_as_float(0x7f800000);
}

// The same block-diagonal operands for the tf32 + bf16 kernels (single-CTA twin, tensor-memory kernel): hi rounded to
// the nearest tf32, bf16 copies of hi / lo for the two correction terms and the tf32-exact pieces of -1/2||c||^2 the
// fold MMA adds to every accumulator column (row jj of the pieces tile = accumulator column jj).  A separate kernel,
// so the 3xTF32 row-packed path (CUML_B200_BF16C=0) keeps its measured binary.
__global__ void prepare_centroids_packed_v2_kernel(const float* __restrict__ C, int k, int d, int k_sub,
                                                   float* __restrict__ hi, float* __restrict__ lo,
                                                   float* __restrict__ cnh, __nv_bfloat16* __restrict__ hb,
                                                   __nv_bfloat16* __restrict__ lb, float* __restrict__ cnp)
{
  const int jj   = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;   // row of B' in [0, 2*k_sub)
  const int lane = threadIdx.x % 32;
  if (jj >= 2 * k_sub) return;
  const int g = jj / k_sub, j = jj % k_sub;
  const int c = lane - g * d;
  const float v = (j < k && c >= 0 && c < d) ? C[static_cast<int64_t>(j) * d + c] : 0.0f;
  const float h = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
  hi[static_cast<int64_t>(jj) * KBLOCK + lane] = h;
  lo[static_cast<int64_t>(jj) * KBLOCK + lane] = v - h;
  hb[static_cast<int64_t>(jj) * KBLOCK + lane] = __float2bfloat16_rn(h);
  lb[static_cast<int64_t>(jj) * KBLOCK + lane] = __float2bfloat16_rn(v - h);
  double s = static_cast<double>(v) * static_cast<double>(v);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) cnh[jj] = (j < k) ? static_cast<float>(0.5 * s) : __int_as_float(0x7f800000);
  if (cnp && lane < 8) {
    const float m  = (j < k) ? -static_cast<float>(0.5 * s) : -3.0e38f;
    const float p1 = __uint_as_float(__float_as_uint(m) & 0xffffe000u);
    const float r1 = m - p1;
    const float p2 = __uint_as_float(__float_as_uint(r1) & 0xffffe000u);
    const float p3 = r1 - p2;
    cnp[static_cast<int64_t>(jj) * 8 + lane] = lane == 0 ? p1 : lane == 1 ? p2 : lane == 2 ? p3 : 0.0f;
  }
}

// packing applies to short rows with few clusters (both halves of the block-diagonal operand fit N <= 256)
int pack_k_sub(int d, int k)
{
  const char* e = std::getenv("CUML_B200_PACK");
  if (e && std::atoi(e) == 0) return 0;
  if (d > 16 || d % 4 != 0 || k > 128) return 0;
  // 65..128 clusters: the packed operand would need a 256-column accumulator tile (no row-owner epilogue, no tensor-memory
  // kernel); one data row per operand row on the tensor-memory kernel measures 3.2 against 4.8 ms (100M x 16, k = 100)
  if (k > 64 && env_flag("CUML_B200_TSP", true)) return 0;
  return k <= 32 ? 32 : (k <= 64 ? 64 : 128);
}

struct TilePlan {
  int kb, bn, a_slots, b_stages, b_resident, a_stream, fold;
  size_t smem;
};

TilePlan plan_tiles(int d, int k, size_t smem_limit)
{
  TilePlan t{};
  t.kb = static_cast<int>(ceil_div(d, KBLOCK));
  // widest N tile (<= 256, multiple of 32) covering k with the least padding
  int bn0 = 256;
  if (k <= 32) bn0 = 32;
  else if (k <= 64) bn0 = 64;
  else if (k <= 128) bn0 = 128;
  auto bytes = [&](int bn_, int as_, int bs_) {
    return static_cast<size_t>(as_) * A_SLOT_BYTES + static_cast<size_t>(bs_) * 2 * bn_ * 128 +
           2 * bn_ * sizeof(float) + 6 * TILE_M * 4 + sizeof(Barriers) + 1024;
  };
  t.bn = 0;
  for (int bn = bn0; bn >= 32 && t.bn == 0; bn /= 2) {
    const int k_tiles = static_cast<int>(ceil_div(k, bn));
    t.a_stream        = (k_tiles > 1 && t.kb > 4) ? 1 : 0;        // wide rows: re-stream X per centroid tile
    const int a_min   = (k_tiles > 1 && !t.a_stream) ? t.kb : std::min(t.kb, 2);  // resident across N tiles
    // at least 3 B stages when a stage is short (N <= 128), 2 otherwise
    const int b_min = (bn <= 128) ? 3 : 2;
    if (bytes(bn, std::max(a_min, 2), b_min) > smem_limit) continue;
    t.bn       = bn;
    t.a_slots  = std::max(a_min, 2);
    t.b_stages = b_min;
    // resident centroids: when every (N tile, K block) fits its own B stage they are loaded once per CTA
    if (k_tiles * t.kb <= MAX_STAGES && bytes(bn, t.a_slots, k_tiles * t.kb) <= smem_limit) {
      t.b_stages   = k_tiles * t.kb;
      t.b_resident = 1;
    }
    // spend what is left: X slots first (two row tiles in flight, or >= ~48 KB of loads in flight when
    // the tile is small), then B depth
    const int a_want = std::min(MAX_A_SLOTS, std::max(2 * t.kb, 6));
    while (t.a_slots < a_want && bytes(bn, t.a_slots + 1, t.b_stages) <= smem_limit) ++t.a_slots;
    while (!t.b_resident && t.b_stages < MAX_STAGES && bytes(bn, t.a_slots, t.b_stages + 1) <= smem_limit) ++t.b_stages;
    t.smem = bytes(bn, t.a_slots, t.b_stages);
  }
  return t;
}

// CTA-pair plan: BN = 256 split across the pair (128 centroid rows per CTA), deeper X ring.
// half norms folded into the accumulator by one extra MMA (default on; CUML_B200_FOLD=0 restores the epilogue add)
bool use_cn_fold()
{
  const char* e = std::getenv("CUML_B200_FOLD");
  return e ? std::atoi(e) != 0 : true;
}

// the solo kernel folds the half norms when the tiles fit next to the single-CTA plan
bool solo_fold_fits(const TilePlan& t, int k, size_t smem_limit)
{
  const int k_tiles = static_cast<int>(ceil_div(k, t.bn));
  // the pieces tile of one centroid tile is bn rows x 32 bytes and must fit the 4 KB fold-tile stride (bn <= 128);
  // bn = 256 only happens here when the pair kernel is switched off for k > 128
  return use_cn_fold() && t.bn <= TILE_M && k_tiles <= 4 &&
         t.smem + static_cast<size_t>(1 + k_tiles) * TILE_M * 32 <= smem_limit;
}

TilePlan plan_tiles_2cta(int d, int k, size_t smem_limit)
{
  TilePlan t{};
  t.kb = static_cast<int>(ceil_div(d, KBLOCK));
  const int bn = 256;
  auto bytes = [&](int as_, int bs_) {
    return static_cast<size_t>(as_) * A_SLOT_BYTES + static_cast<size_t>(bs_) * bn * 128 + 2 * bn * sizeof(float) +
           6 * TILE_M * 4 + sizeof(Barriers) + 1024;
  };
  const int k_tiles = static_cast<int>(ceil_div(k, bn));
  t.fold            = (use_cn_fold() && k_tiles <= 4) ? 1 : 0;
  if (t.fold) smem_limit -= static_cast<size_t>(1 + k_tiles) * TILE_M * 32;
  t.a_stream        = (k_tiles > 1 && t.kb > 4) ? 1 : 0;
  const int a_min   = (k_tiles > 1 && !t.a_stream) ? t.kb : std::min(t.kb, 2);
  int b_stages      = t.a_stream ? 3 : 2;
  int resident      = 0;
  if (k_tiles * t.kb <= MAX_STAGES && bytes(std::max(a_min, 2), k_tiles * t.kb) <= smem_limit) {
    b_stages = k_tiles * t.kb;
    resident = 1;
  }
  if (bytes(std::max(a_min, 2), b_stages) > smem_limit) return t;  // bn stays 0: not available
  t.bn = bn;
  t.a_slots = std::max(a_min, 2);
  t.b_stages = b_stages;
  t.b_resident = resident;
  const int a_want = std::min(MAX_A_SLOTS, std::max(2 * t.kb, 6));
  while (t.a_slots < a_want && bytes(t.a_slots + 1, t.b_stages) <= smem_limit) ++t.a_slots;
  while (!t.b_resident && t.b_stages < MAX_STAGES && bytes(t.a_slots, t.b_stages + 1) <= smem_limit) ++t.b_stages;
  t.smem = bytes(t.a_slots, t.b_stages) + (t.fold ? static_cast<size_t>(1 + k_tiles) * TILE_M * 32 : 0);
  return t;
}

// CTA pairs pay off when the single-CTA plan is starved of X slots or cannot use N = 256
bool use_2cta(const Handle& h, int d, int k)
{
  const char* e = std::getenv("CUML_B200_2CTA");
  if (e) return std::atoi(e) != 0 && k > 128;
  if (k <= 128 || (h.sm_count % 2) != 0) return false;
  return plan_tiles_2cta(d, k, h.smem_optin).bn > 0;
}

// bf16 correction terms in the CTA-pair kernel (default on; CUML_B200_BF16C=0 restores pure 3xTF32)
bool use_bf16_corrections()
{
  const char* e = std::getenv("CUML_B200_BF16C");
  return e ? std::atoi(e) != 0 : true;
}

}  // namespace

bool tc_supported(int64_t d, int k)
{
  return d >= 4 && d % 4 == 0 && d <= 1024 && k >= 1 && k <= (1 << 20);
}

int tc_variant(const Handle& h, int d, int k)
{
  if (pack_k_sub(d, k)) return use_bf16_corrections() ? 5 : 1;
  if (use_2cta(h, d, k)) return use_bf16_corrections() ? 3 : 2;
  if (use_bf16_corrections()) return 5;
  return 1;
}

void tc_prepare(Handle& h, const float* C, int k, int d, TcCentroids& out, bool allow_bf16)
{
  if (const int k_sub = pack_k_sub(d, k)) {
    const int k_pad = 2 * k_sub;
    if (out.k_pad != k_pad || out.d_pad != KBLOCK || !out.hi.get()) {
      out.hi.alloc(static_cast<size_t>(k_pad) * KBLOCK, h.stream);
      out.lo.alloc(static_cast<size_t>(k_pad) * KBLOCK, h.stream);
      out.cnh.alloc(k_pad, h.stream);
      out.k_pad = k_pad;
      out.d_pad = KBLOCK;
    }
    out.block_n = k_pad;
    out.pack    = 2;
    out.k_sub   = k_sub;
    out.bf16c   = 0;
    out.fold    = 0;
    if (allow_bf16 && use_bf16_corrections()) {
      // tf32 + bf16 kernels on the row-packed operands (see prepare_centroids_packed_v2_kernel)
      const TilePlan t = plan_tiles(2 * d, k_pad, h.smem_optin);
      CB2_EXPECTS(t.bn == k_pad && t.kb == 1, "row-packed plan mismatch");
      out.bf16c = 1;
      out.fold  = solo_fold_fits(t, k_pad, h.smem_optin) ? 1 : 0;
      if (out.hb.n < static_cast<size_t>(k_pad) * KBLOCK) {
        out.hb.alloc(static_cast<size_t>(k_pad) * KBLOCK, h.stream);
        out.lb.alloc(static_cast<size_t>(k_pad) * KBLOCK, h.stream);
      }
      if (out.fold && out.cnp.n < static_cast<size_t>(k_pad) * 8) out.cnp.alloc(static_cast<size_t>(k_pad) * 8, h.stream);
      prepare_centroids_packed_v2_kernel<<<static_cast<unsigned>(ceil_div(k_pad, 8)), 256, 0, h.stream>>>(
        C, k, d, k_sub, out.hi.get(), out.lo.get(), out.cnh.get(), reinterpret_cast<__nv_bfloat16*>(out.hb.get()),
        reinterpret_cast<__nv_bfloat16*>(out.lb.get()), out.fold ? out.cnp.get() : nullptr);
      CB2_CHECK_LAUNCH();
      return;
    }
    prepare_centroids_packed_kernel<<<static_cast<unsigned>(ceil_div(k_pad, 8)), 256, 0, h.stream>>>(
      C, k, d, k_sub, out.hi.get(), out.lo.get(), out.cnh.get());
    CB2_CHECK_LAUNCH();
    return;
  }
  out.pack  = 1;
  out.k_sub = 0;
  const bool pair = use_2cta(h, d, k);
  TilePlan t = pair ? plan_tiles_2cta(d, k, h.smem_optin) : plan_tiles(d, k, h.smem_optin);
  CB2_EXPECTS(t.bn > 0, "tcgen05 k-means tile plan does not fit shared memory");
  const int d_pad = t.kb * KBLOCK;
  const int k_pad = static_cast<int>(ceil_div(k, t.bn)) * t.bn;
  if (out.k_pad != k_pad || out.d_pad != d_pad || !out.hi.get()) {
    out.hi.alloc(static_cast<size_t>(k_pad) * d_pad, h.stream);
    out.lo.alloc(static_cast<size_t>(k_pad) * d_pad, h.stream);
    out.cnh.alloc(k_pad, h.stream);
    out.k_pad = k_pad;
    out.d_pad = d_pad;
  }
  out.block_n = t.bn;
  const bool solo = !pair;
  out.bf16c   = ((pair || solo) && allow_bf16 && use_bf16_corrections()) ? 1 : 0;
  if (out.bf16c && out.hb.n < static_cast<size_t>(k_pad) * d_pad) {
    out.hb.alloc(static_cast<size_t>(k_pad) * d_pad, h.stream);
    out.lb.alloc(static_cast<size_t>(k_pad) * d_pad, h.stream);
  }
  out.fold = ((pair && t.fold) || (solo && solo_fold_fits(t, k, h.smem_optin))) ? 1 : 0;
  if (out.fold && out.cnp.n < static_cast<size_t>(k_pad) * 8) out.cnp.alloc(static_cast<size_t>(k_pad) * 8, h.stream);
  prepare_centroids_kernel<<<static_cast<unsigned>(ceil_div(k_pad, 8)), 256, 0, h.stream>>>(
    C, k, d, k_pad, d_pad, out.hi.get(), out.lo.get(), out.cnh.get(),
    out.bf16c ? reinterpret_cast<__nv_bfloat16*>(out.hb.get()) : nullptr,
    out.bf16c ? reinterpret_cast<__nv_bfloat16*>(out.lb.get()) : nullptr, out.fold ? out.cnp.get() : nullptr);
  CB2_CHECK_LAUNCH();
}

// label of ONE row against the packed operand buffers (the odd last row of a row-packed launch)
__global__ void assign_tail_row_kernel(const float* __restrict__ x, int d, int k, const float* __restrict__ hi,
                                       const float* __restrict__ lo, int32_t* __restrict__ label)
{
  // one warp; lane j, j+32, ... scans centroids; exact (x-c)^2 in fp32 with c = hi + lo
  float best = __int_as_float(0x7f800000);
  int bidx   = 0x7fffffff;
  for (int j = threadIdx.x; j < k; j += 32) {
    float s = 0.f;
    for (int c = 0; c < d; ++c) {
      const float cv = hi[j * KBLOCK + c] + lo[j * KBLOCK + c];
      const float df = x[c] - cv;
      s += df * df;
    }
    if (s < best) { best = s; bidx = j; }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, off);
    const int oi   = __shfl_xor_sync(0xffffffffu, bidx, off);
    if (ob < best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
  }
  if (threadIdx.x == 0) *label = bidx;
}

// fused M-step with an odd number of rows: the last row (labelled by assign_tail_row_kernel) is added to CTA 0's partials
__global__ void mstep_tail_row_kernel(const float* __restrict__ x, int d, const int32_t* __restrict__ label,
                                      float* __restrict__ S0, float* __restrict__ W0)
{
  const int j = *label;
  for (int c = threadIdx.x; c < d; c += 32) S0[static_cast<int64_t>(j) * d + c] += x[c];
  if (threadIdx.x == 0) W0[j] += 1.0f;
}

bool tc_transform_supported(const Handle& h, int64_t d, int k)
{
  // the distance-matrix epilogue exists for the unpacked shared-memory-operand kernels
  return h.cc_major == 10 && tc_supported(d, k) && !pack_k_sub(static_cast<int>(d), k);
}

bool tc_best_supported(const Handle& h, int d, int k)
{
  return h.cc_major == 10 && tc_supported(d, k);
}

// ---- row-packed kernel with the X operand in tensor memory (fused_l2_argmin_tsp_kernel) ------------------------------
// CUML_B200_TSP (read per call so that a test can switch it).
static bool use_tsp() { return env_flag("CUML_B200_TSP", true); }
// fused M-step: C5 4.6 ms per Lloyd step against 3.7 (E) + 2.1 (M) ms for two kernels.  (On the shared-memory-operand
// twin the same fusion measured SLOWER, 8.0 ms against 5.1 + 2.1 ms, and was removed: profiles/README.md.)
static bool fused_mstep_on() { return env_flag("CUML_B200_FUSED_MSTEP", true); }

struct TspPlan {
  int raw_slots, n_acc, a_col0;
  size_t smem;
};
static bool plan_tsp(const Handle& h, int d, int k, bool mstep, TspPlan& out)
{
  if (h.cc_major != 10 || !use_bf16_corrections() || !use_cn_fold()) return false;
  const int k_sub = pack_k_sub(d, k);
  int bn, ms_k;
  if (k_sub) {   // two data rows per operand row (n_features <= 16), block-diagonal centroids
    // fused M-step: its cost is per operand row, so it pays for rows of 8+ features (200M rows: 16 features 4.6 against
    // 7.0 ms, 8 features 4.8 against 4.8, 4 features 4.4 against 3.5 -- there the stand-alone M-step takes 0.65 ms)
    if (k_sub > 64 || (mstep && d < 8)) return false;
    bn   = 2 * k_sub;
    ms_k = k_sub;
  } else {       // one data row per operand row: 17..32 features in one K-block, one centroid tile of <= 128 columns
    if (d > KBLOCK || d % 4 != 0 || k > 128) return false;
    const TilePlan t = plan_tiles(d, k, h.smem_optin);
    if (t.bn == 0 || t.bn > 128 || t.kb != 1 || k > t.bn || !solo_fold_fits(t, k, h.smem_optin)) return false;
    bn   = t.bn;
    ms_k = bn;
    if (mstep && bn > 64) return false;   // eight private tables of bn + 1 rows
  }
  const size_t fixed = static_cast<size_t>(bn) * 256 + 2 * static_cast<size_t>(TILE_M) * 32 +
                       ((sizeof(Barriers) + 127) & ~size_t(127)) + (mstep ? tsp_mstep_smem_bytes(ms_k) : 0) + 1024;
  if (h.smem_optin < fixed + 3 * static_cast<size_t>(KBLOCK_BYTES)) return false;
  out.raw_slots = static_cast<int>(std::min<size_t>(mstep ? MAX_A_SLOTS : MAX_RAW, (h.smem_optin - fixed) / KBLOCK_BYTES));
  out.n_acc     = std::min(mstep ? 3 : 4, (512 - 128) / bn);   // fused M-step: warps 20..23 accumulate, 3 epilogue groups
  out.a_col0    = out.n_acc * bn;
  out.smem      = fixed + static_cast<size_t>(out.raw_slots) * KBLOCK_BYTES;
  return true;
}

bool tc_fused_update_supported(const Handle& h, int d, int k)
{
  TspPlan tp;
  return h.cc_major == 10 && fused_mstep_on() && use_tsp() && plan_tsp(h, d, k, true, tp);
}

void tc_assign(Handle& h, const float* X, int64_t n, int d, int k, const TcCentroids& cen, int32_t* labels,
               float* dbg_dots, const TcDistOut* dist, float* best_out, TcMstepOut* mstep)
{
  if (mstep) mstep->row_blocks = 0;
  if (n == 0) return;
  CB2_EXPECTS(!best_out || (!dbg_dots && !dist && labels), "best-value output excludes the debug dump and the distance matrix");
  CB2_EXPECTS(!dist || cen.pack == 1, "distance-matrix mode does not support row packing");
  CB2_EXPECTS(h.cc_major == 10, "the tcgen05 k-means engine needs an sm_100-class GPU (B200)");
  CB2_EXPECTS(reinterpret_cast<uintptr_t>(X) % 16 == 0, "X must be 16-byte aligned for TMA");
  // TMA tile coordinates are 32-bit: more rows than that would be mis-addressed silently
  CB2_EXPECTS(n < (int64_t(1) << 31), "the tcgen05 k-means engine takes at most 2^31 - 1 rows per array; pass the rows as several "
                                      "partitions (ML::kmeans::fit partition list / chunked predict)");
  if (cen.pack == 2) {
    // two rows per operand row: X viewed as [n/2, 2d]
    const int64_t n2 = n / 2;
    if (n2 > 0) {
      TilePlan t = plan_tiles(2 * d, cen.k_pad, h.smem_optin);
      CB2_EXPECTS(t.bn == cen.k_pad && t.kb == 1, "row-packed plan mismatch");
      FusedParams p{};
      p.n = 2 * n2; p.m_tiles = ceil_div(n2, TILE_M); p.k_tiles = 1; p.d = 2 * d; p.kb = 1; p.bn = t.bn;
      p.a_slots = t.a_slots; p.b_stages = t.b_stages; p.b_resident = t.b_resident;
      p.n_acc = std::min(MAX_ACC, 512 / t.bn); p.pack = 2; p.k_sub = cen.k_sub;
      uint32_t cols = 32;
      while (cols < static_cast<uint32_t>(p.n_acc * t.bn)) cols <<= 1;
      p.tmem_cols = cols;
      p.cnh = cen.cnh.get(); p.labels = labels; p.dbg_dots = best_out;   // DIST = 3: winning value per row
      CUtensorMap tm_x  = make_map_2d(X, static_cast<uint64_t>(2 * d), static_cast<uint64_t>(n2),
                                      static_cast<uint64_t>(2 * d) * sizeof(float), KBLOCK, TILE_M,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
      CUtensorMap tm_hi = make_map_2d(cen.hi.get(), KBLOCK, cen.k_pad, KBLOCK * sizeof(float), KBLOCK, t.bn,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
      CUtensorMap tm_lo = make_map_2d(cen.lo.get(), KBLOCK, cen.k_pad, KBLOCK * sizeof(float), KBLOCK, t.bn,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
      static PerDeviceOnce pk_attr;
      pk_attr.run(h.device, [&] {
        CB2_CUDA(cudaFuncSetAttribute(fused_l2_argmin_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(h.smem_optin)));
        CB2_CUDA(cudaFuncSetAttribute(fused_l2_argmin_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(h.smem_optin)));
      });
      // role-level cycle counters of CTA 0 (measurement aid, CUML_B200_DBG_CLK=1), as in the unpacked path below
      DevBuf<long long> clk;
      const bool want_clk = std::getenv("CUML_B200_DBG_CLK") != nullptr;
      if (want_clk) {
        clk.alloc(16, h.stream);
        CB2_CUDA(cudaMemsetAsync(clk.get(), 0, 16 * sizeof(long long), h.stream));
        p.dbg_clk = clk.get();
      }
      EventPair ev{};
      if (h.timing) ev = h.begin_event();
      const unsigned grid = static_cast<unsigned>(std::min<int64_t>(p.m_tiles, h.sm_count));
      TspPlan tp{};
      const bool tsp_mstep = mstep && fused_mstep_on();   // (an odd last row is added to CTA 0's partials below)
      if (use_tsp() && cen.bf16c && cen.fold && !best_out && plan_tsp(h, d, k, tsp_mstep, tp)) {
        // ---- X operand in tensor memory (fused_l2_argmin_tsp_kernel), optionally with the fused M-step ----
        p.raw_slots = tp.raw_slots; p.a_slots = 2; p.n_acc = tp.n_acc; p.a_col0 = tp.a_col0; p.tmem_cols = 512;
        p.fold = 1; p.b_resident = 1; p.b_stages = 1;
        if (tsp_mstep) {
          if (mstep->partial_S->n < static_cast<size_t>(grid) * k * d) mstep->partial_S->alloc(static_cast<size_t>(grid) * k * d, h.stream);
          if (mstep->partial_W->n < static_cast<size_t>(grid) * k) mstep->partial_W->alloc(static_cast<size_t>(grid) * k, h.stream);
          p.dbg_dots = mstep->partial_S->get();   // idle fields carry the M-step outputs (see the kernel's header)
          p.cnh      = mstep->partial_W->get();
          p.a_stream = k;
        }
        CUtensorMap tm_hb = make_map_2d(cen.hb.get(), KBLOCK, cen.k_pad, static_cast<uint64_t>(KBLOCK) * 2, KBLOCK, t.bn,
                                        CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                        CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
        CUtensorMap tm_lb = make_map_2d(cen.lb.get(), KBLOCK, cen.k_pad, static_cast<uint64_t>(KBLOCK) * 2, KBLOCK, t.bn,
                                        CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                        CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
        CUtensorMap tm_cn = make_map_2d(cen.cnp.get(), 8, cen.k_pad, 8 * sizeof(float), 8, t.bn, CU_TENSOR_MAP_SWIZZLE_32B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        static PerDeviceOnce tsp_attr;
        tsp_attr.run(h.device, [&] {
          CB2_CUDA(cudaFuncSetAttribute(fused_l2_argmin_tsp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(h.smem_optin)));
          CB2_CUDA(cudaFuncSetAttribute(fused_l2_argmin_tsp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(h.smem_optin)));
        });
        if (tsp_mstep) {
          fused_l2_argmin_tsp_kernel<true><<<grid, PAIR_THREADS, tp.smem, h.stream>>>(tm_x, tm_hi, tm_hb, tm_lb, tm_cn, p);
          mstep->row_blocks = static_cast<int>(grid);
        } else {
          fused_l2_argmin_tsp_kernel<false><<<grid, PAIR_THREADS, tp.smem, h.stream>>>(tm_x, tm_hi, tm_hb, tm_lb, tm_cn, p);
        }
      } else if (cen.bf16c && !best_out) {
        // single-CTA tf32 + bf16 kernel on the row-packed operands (CUML_B200_TSP=0, or 65..128 clusters).  (It has no best-value
        // epilogue; the 3xTF32 kernel below reads the same buffers: hi rounded to nearest is still tf32-exact and
        // lo = c - hi is exact.)
        p.fold = (cen.fold && solo_fold_fits(t, cen.k_pad, h.smem_optin)) ? 1 : 0;
        p.l2_ahead = 3;
        const size_t smem = t.smem + (p.fold ? static_cast<size_t>(1 + p.k_tiles) * TILE_M * 32 : 0);
        CUtensorMap tm_hb = make_map_2d(cen.hb.get(), KBLOCK, cen.k_pad, static_cast<uint64_t>(KBLOCK) * 2, KBLOCK, t.bn,
                                        CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                        CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
        CUtensorMap tm_lb = make_map_2d(cen.lb.get(), KBLOCK, cen.k_pad, static_cast<uint64_t>(KBLOCK) * 2, KBLOCK, t.bn,
                                        CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                        CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
        CUtensorMap tm_cn = tm_hi;
        if (p.fold)
          tm_cn = make_map_2d(cen.cnp.get(), 8, cen.k_pad, 8 * sizeof(float), 8, t.bn, CU_TENSOR_MAP_SWIZZLE_32B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        static PerDeviceOnce pk_solo_attr;
        pk_solo_attr.run(h.device, [&] {
          CB2_CUDA(cudaFuncSetAttribute(fused_l2_argmin_solo_kernel<true, 0, false>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(h.smem_optin)));
          CB2_CUDA(cudaFuncSetAttribute(fused_l2_argmin_solo_kernel<true, 0, false, true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(h.smem_optin)));
        });
        if (t.bn <= TILE_M)   // one centroid tile, >= 4 accumulator stages: row-owner epilogue
          fused_l2_argmin_solo_kernel<true, 0, false, true><<<grid, PAIR_THREADS, smem, h.stream>>>(tm_x, tm_hi, tm_hb, tm_lb, tm_cn, p);
        else
          fused_l2_argmin_solo_kernel<true, 0, false><<<grid, PAIR_THREADS, smem, h.stream>>>(tm_x, tm_hi, tm_hb, tm_lb, tm_cn, p);
      } else if (best_out) fused_l2_argmin_kernel<3><<<grid, NUM_THREADS, t.smem, h.stream>>>(tm_x, tm_hi, tm_lo, p);
      else fused_l2_argmin_kernel<0><<<grid, NUM_THREADS, t.smem, h.stream>>>(tm_x, tm_hi, tm_lo, p);   // 3xTF32 (CUML_B200_BF16C=0)
      CB2_CHECK_LAUNCH();
      if (h.timing) h.end_event(ev, true);
      if (want_clk) {
        long long hc[16];
        CB2_CUDA(cudaMemcpyAsync(hc, clk.get(), sizeof(hc), cudaMemcpyDeviceToHost, h.stream));
        CB2_CUDA(cudaStreamSynchronize(h.stream));
        std::printf("[cuml_b200 clk packed] tiles/CTA %lld | producer wait %lld / %lld | converter wait %lld / %lld | mma wait acc %lld a %lld "
                    "b %lld / %lld (cycles, CTA 0)\n",
                    static_cast<long long>((p.m_tiles + grid - 1) / grid), hc[0], hc[1], hc[2], hc[3], hc[4], hc[5], hc[6], hc[7]);
      }
    }
    if (n & 1) {   // (labels only: a caller that wants the winning value of an odd last row computes it itself)
      assign_tail_row_kernel<<<1, 32, 0, h.stream>>>(X + (n - 1) * d, d, k, cen.hi.get(), cen.lo.get(), labels + (n - 1));
      CB2_CHECK_LAUNCH();
      if (mstep && mstep->row_blocks > 0) {   // fused M-step: the odd row joins CTA 0's partial sums / counts
        mstep_tail_row_kernel<<<1, 32, 0, h.stream>>>(X + (n - 1) * d, d, labels + (n - 1), mstep->partial_S->get(),
                                                      mstep->partial_W->get());
        CB2_CHECK_LAUNCH();
      }
    }
    return;
  }
  const bool pair = use_2cta(h, d, k);
  TilePlan t = pair ? plan_tiles_2cta(d, k, h.smem_optin) : plan_tiles(d, k, h.smem_optin);
  CB2_EXPECTS(t.bn == cen.block_n, "centroid operand buffers were prepared for a different tile plan");

  FusedParams p{};
  p.n         = n;
  p.m_tiles   = ceil_div(n, TILE_M);
  p.k_tiles   = cen.k_pad / t.bn;
  p.d         = d;
  p.kb        = t.kb;
  p.bn        = t.bn;
  p.a_slots   = t.a_slots;
  p.b_stages  = t.b_stages;
  p.b_resident = t.b_resident;
  p.a_stream   = t.a_stream;
  p.pack       = 1;
  p.k_sub      = 0;
  p.n_acc = std::min(MAX_ACC, 512 / t.bn);
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(p.n_acc * t.bn)) cols <<= 1;
  p.tmem_cols = cols;
  p.cnh       = cen.cnh.get();
  p.labels    = labels;
  p.dbg_dots  = best_out ? best_out : dbg_dots;   // DIST = 3 instantiations store the winning value there
  if (dist) {   // see FusedParams: idle fields carry the distance-matrix arguments
    p.dbg_dots  = dist->out;
    p.labels    = reinterpret_cast<int32_t*>(const_cast<float*>(dist->xnorm));
    p.k_sub     = k;
    p.raw_slots = dist->sqrt ? 1 : 0;
  }
  {
    p.l2_ahead = 3;   // row tiles prefetched into L2 ahead of the ring (0 / 6 / 12 measured the same)
  }
  DevBuf<long long> clk;
  const bool want_clk = std::getenv("CUML_B200_DBG_CLK") != nullptr;
  if (want_clk) {
    clk.alloc(16, h.stream);
    CB2_CUDA(cudaMemsetAsync(clk.get(), 0, 16 * sizeof(long long), h.stream));
    p.dbg_clk = clk.get();
  }

  CUtensorMap tm_x  = make_map_2d(X, static_cast<uint64_t>(d), static_cast<uint64_t>(n),
                                  static_cast<uint64_t>(d) * sizeof(float), KBLOCK, TILE_M,
                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  const uint32_t b_box_rows = pair ? t.bn / 2 : t.bn;
  CUtensorMap tm_hi = make_map_2d(cen.hi.get(), cen.d_pad, cen.k_pad, static_cast<uint64_t>(cen.d_pad) * sizeof(float),
                                  KBLOCK, b_box_rows, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
  CUtensorMap tm_lo = make_map_2d(cen.lo.get(), cen.d_pad, cen.k_pad, static_cast<uint64_t>(cen.d_pad) * sizeof(float),
                                  KBLOCK, b_box_rows, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);

  static PerDeviceOnce attr_set;
  attr_set.run(h.device, [&] {
    CB2_CUDA(cudaFuncSetAttribute(fused_l2_argmin_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(h.smem_optin)));
    CB2_CUDA(cudaFuncSetAttribute(fused_l2_argmin_2cta_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(h.smem_optin)));
    CB2_CUDA(cudaFuncSetAttribute(fused_l2_argmin_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(h.smem_optin)));
    CB2_CUDA(cudaFuncSetAttribute(fused_l2_argmin_2cta_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(h.smem_optin)));
    CB2_CUDA(cudaFuncSetAttribute(fused_l2_argmin_2cta_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(h.smem_optin)));
    CB2_CUDA(cudaFuncSetAttribute(fused_l2_argmin_2cta_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(h.smem_optin)));
    CB2_CUDA(cudaFuncSetAttribute(fused_l2_argmin_2cta_kernel<true, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(h.smem_optin)));
  });
  EventPair ev{};
  if (h.timing) ev = h.begin_event();
  const long long grid_dbg = pair ? h.sm_count / 2 : h.sm_count;
  if (pair) {
    // one CTA pair per TPC; grid must be even (cluster dims 2x1x1 are compiled into the kernel)
    // the debug dump wants the bare x.c accumulators: no fold then (the plan keeps the space reserved)
    p.fold = (cen.fold && t.fold && !dbg_dots) ? 1 : 0;
    CUtensorMap tm_cn = tm_hi;
    if (p.fold)
      tm_cn = make_map_2d(cen.cnp.get(), 8, cen.k_pad, 8 * sizeof(float), 8, b_box_rows, CU_TENSOR_MAP_SWIZZLE_32B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    const int64_t pair_tiles = (p.m_tiles + 1) / 2;
    const unsigned grid = 2u * static_cast<unsigned>(std::min<int64_t>(pair_tiles, h.sm_count / 2));
    if (cen.bf16c) {
      CUtensorMap tm_hb = make_map_2d(cen.hb.get(), cen.d_pad, cen.k_pad, static_cast<uint64_t>(cen.d_pad) * 2, KBLOCK,
                                      b_box_rows, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
      CUtensorMap tm_lb = make_map_2d(cen.lb.get(), cen.d_pad, cen.k_pad, static_cast<uint64_t>(cen.d_pad) * 2, KBLOCK,
                                      b_box_rows, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
      if (best_out) {
        fused_l2_argmin_2cta_kernel<true, 3><<<grid, PAIR_THREADS, t.smem, h.stream>>>(tm_x, tm_hi, tm_hb, tm_lb, tm_cn, p);
      } else if (want_clk) {   // role-level cycle counters: a separate instantiation, printed below
        static PerDeviceOnce clk_attr;
        clk_attr.run(h.device, [&] {
          CB2_CUDA(cudaFuncSetAttribute(fused_l2_argmin_2cta_kernel<true, 0, true, true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(h.smem_optin)));
        });
        fused_l2_argmin_2cta_kernel<true, 0, true, true><<<grid, PAIR_THREADS, t.smem, h.stream>>>(tm_x, tm_hi, tm_hb, tm_lb, tm_cn, p);
      } else {   // the converter leaves the raw tile alone: the tensor core truncates it to tf32 (TRUNC)
        fused_l2_argmin_2cta_kernel<true, 0, true><<<grid, PAIR_THREADS, t.smem, h.stream>>>(tm_x, tm_hi, tm_hb, tm_lb, tm_cn, p);
      }
    } else if (best_out) {
      fused_l2_argmin_2cta_kernel<false, 3><<<grid, PAIR_THREADS, t.smem, h.stream>>>(tm_x, tm_hi, tm_lo, tm_lo, tm_cn, p);
    } else if (dist) {   // distance matrix (transform): 3xTF32, lane-pair stores (32-byte sectors)
      fused_l2_argmin_2cta_kernel<false, 2><<<grid, PAIR_THREADS, t.smem, h.stream>>>(tm_x, tm_hi, tm_lo, tm_lo, tm_cn, p);
    } else {
      fused_l2_argmin_2cta_kernel<false><<<grid, PAIR_THREADS, t.smem, h.stream>>>(tm_x, tm_hi, tm_lo, tm_lo, tm_cn, p);
    }
  } else if (!best_out) {
    // single-CTA twin of the pair kernel (see fused_l2_argmin_solo_kernel)
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>(p.m_tiles, h.sm_count));
    p.fold = (cen.fold && !dbg_dots && solo_fold_fits(t, k, h.smem_optin)) ? 1 : 0;
    const size_t smem = t.smem + (p.fold ? static_cast<size_t>(1 + p.k_tiles) * TILE_M * 32 : 0);
    CUtensorMap tm_cn = tm_hi;
    if (p.fold)
      tm_cn = make_map_2d(cen.cnp.get(), 8, cen.k_pad, 8 * sizeof(float), 8, b_box_rows, CU_TENSOR_MAP_SWIZZLE_32B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    static PerDeviceOnce solo_attr;
    solo_attr.run(h.device, [&] {
      CB2_CUDA(cudaFuncSetAttribute(fused_l2_argmin_solo_kernel<true, 0, false>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(h.smem_optin)));
      CB2_CUDA(cudaFuncSetAttribute(fused_l2_argmin_solo_kernel<false, 0, false>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(h.smem_optin)));
      CB2_CUDA(cudaFuncSetAttribute(fused_l2_argmin_solo_kernel<false, 1, false>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(h.smem_optin)));
      CB2_CUDA(cudaFuncSetAttribute(fused_l2_argmin_solo_kernel<true, 0, false, true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(h.smem_optin)));
    });
    if (cen.bf16c && !dist) {
      CUtensorMap tm_hb = make_map_2d(cen.hb.get(), cen.d_pad, cen.k_pad, static_cast<uint64_t>(cen.d_pad) * 2, KBLOCK,
                                      b_box_rows, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
      CUtensorMap tm_lb = make_map_2d(cen.lb.get(), cen.d_pad, cen.k_pad, static_cast<uint64_t>(cen.d_pad) * 2, KBLOCK,
                                      b_box_rows, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
      TspPlan tp{};
      const bool tsp_mstep = mstep && fused_mstep_on();
      if (use_tsp() && p.fold && !dbg_dots && p.k_tiles == 1 && plan_tsp(h, d, k, tsp_mstep, tp)) {
        // ---- 17..32 features, k <= 128: X operand in tensor memory (fused_l2_argmin_tsp_kernel, one data row per
        // operand row), optionally with the fused M-step ----
        p.raw_slots = tp.raw_slots; p.a_slots = 2; p.n_acc = tp.n_acc; p.a_col0 = tp.a_col0; p.tmem_cols = 512;
        p.b_resident = 1; p.b_stages = 1;
        if (tsp_mstep) {
          if (mstep->partial_S->n < static_cast<size_t>(grid) * k * d) mstep->partial_S->alloc(static_cast<size_t>(grid) * k * d, h.stream);
          if (mstep->partial_W->n < static_cast<size_t>(grid) * k) mstep->partial_W->alloc(static_cast<size_t>(grid) * k, h.stream);
          p.dbg_dots = mstep->partial_S->get();   // idle fields carry the M-step outputs (see the kernel's header)
          p.cnh      = mstep->partial_W->get();
          p.a_stream = k;
        }
        static PerDeviceOnce tspu_attr;
        tspu_attr.run(h.device, [&] {
          CB2_CUDA(cudaFuncSetAttribute(fused_l2_argmin_tsp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(h.smem_optin)));
          CB2_CUDA(cudaFuncSetAttribute(fused_l2_argmin_tsp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(h.smem_optin)));
        });
        if (tsp_mstep) {
          fused_l2_argmin_tsp_kernel<true><<<grid, PAIR_THREADS, tp.smem, h.stream>>>(tm_x, tm_hi, tm_hb, tm_lb, tm_cn, p);
          mstep->row_blocks = static_cast<int>(grid);
        } else {
          fused_l2_argmin_tsp_kernel<false><<<grid, PAIR_THREADS, tp.smem, h.stream>>>(tm_x, tm_hi, tm_hb, tm_lb, tm_cn, p);
        }
      } else if (p.k_tiles == 1 && t.bn <= TILE_M && !dbg_dots)   // row-owner epilogue
        fused_l2_argmin_solo_kernel<true, 0, false, true><<<grid, PAIR_THREADS, smem, h.stream>>>(tm_x, tm_hi, tm_hb, tm_lb, tm_cn, p);
      else
        fused_l2_argmin_solo_kernel<true, 0, false><<<grid, PAIR_THREADS, smem, h.stream>>>(tm_x, tm_hi, tm_hb, tm_lb, tm_cn, p);
    } else if (dist) {
      fused_l2_argmin_solo_kernel<false, 1, false><<<grid, PAIR_THREADS, smem, h.stream>>>(tm_x, tm_hi, tm_lo, tm_lo, tm_cn, p);
    } else {
      fused_l2_argmin_solo_kernel<false, 0, false><<<grid, PAIR_THREADS, smem, h.stream>>>(tm_x, tm_hi, tm_lo, tm_lo, tm_cn, p);
    }
  } else {
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>(p.m_tiles, h.sm_count));
    // winning-value epilogue (seeding: min-distance update) of the shapes the single-CTA kernels take
    fused_l2_argmin_kernel<3><<<grid, NUM_THREADS, t.smem, h.stream>>>(tm_x, tm_hi, tm_lo, p);
  }
  CB2_CHECK_LAUNCH();
  if (h.timing) h.end_event(ev, true);
  if (want_clk) {
    long long hc[16];
    CB2_CUDA(cudaMemcpyAsync(hc, clk.get(), sizeof(hc), cudaMemcpyDeviceToHost, h.stream));
    CB2_CUDA(cudaStreamSynchronize(h.stream));
    std::printf("[cuml_b200 clk] tiles/CTA %lld | producer wait %lld / %lld | converter wait %lld / %lld | mma wait acc %lld a %lld b %lld / %lld | "
                "epilogue wait %lld / %lld | epilogue hold %lld merge %lld (cycles, CTA 0)\n",
                static_cast<long long>((p.m_tiles + grid_dbg - 1) / grid_dbg), hc[0], hc[1], hc[2], hc[3], hc[4], hc[5], hc[6], hc[7], hc[8], hc[9], hc[10], hc[11]);
  }
}

}  // namespace cb2
