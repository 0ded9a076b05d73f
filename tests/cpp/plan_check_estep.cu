// CPU-only property check of the E-step tile planners (plan_tiles, plan_tiles_2cta, pack_k_sub): for every supported
// (n_features, n_clusters) on a grid the plan fits the 227 KB shared-memory limit and keeps the invariants the kernels
// rely on (ring depths within the barrier arrays, resident X slots when several centroid tiles reuse a row tile, ...).
// The planners live in an anonymous namespace, so the kernel source is included whole; nothing is launched.
#include "fused_l2_argmin_sm100.cu"
#include <cstdio>
using namespace cb2;
int main() {
  const size_t lim = 232448;
  int bad = 0, n = 0;
  for (int d = 4; d <= 1024; d += 4)
    for (int k : {1, 2, 7, 8, 16, 31, 32, 33, 64, 65, 100, 128, 129, 200, 255, 256, 257, 300, 511, 512, 513, 1000, 1024, 1030, 2048, 4096, 5000, 65536, 1 << 20}) {
      ++n;
      for (int pair = 0; pair < 2; ++pair) {
        if (pair && k <= 128) continue;
        TilePlan t = pair ? plan_tiles_2cta(d, k, lim) : plan_tiles(d, k, lim);
        if (t.bn == 0) { if (!pair) { std::printf("NOFIT d=%d k=%d pair=%d\n", d, k, pair); ++bad; } continue; }
        const int k_tiles = (k + t.bn - 1) / t.bn;
        bool ok = t.smem <= lim && t.a_slots >= 2 && t.a_slots <= MAX_A_SLOTS && t.b_stages >= 2 - 1 && t.b_stages <= MAX_STAGES &&
                  t.bn % 32 == 0 && t.bn <= 256 && t.kb == (d + 31) / 32;
        if (k_tiles > 1 && !t.a_stream) ok = ok && t.a_slots >= t.kb;
        if (t.b_resident) ok = ok && t.b_stages == k_tiles * t.kb;
        if (!ok) { std::printf("BAD d=%d k=%d pair=%d bn=%d a=%d b=%d res=%d stream=%d smem=%zu\n", d, k, pair, t.bn, t.a_slots, t.b_stages, t.b_resident, t.a_stream, t.smem); ++bad; }
      }
      // packed
      int ks = pack_k_sub(d, k);
      if (ks) { TilePlan t = plan_tiles(2 * d, 2 * ks, lim); if (t.bn != 2 * ks || t.kb != 1) { std::printf("PACK BAD d=%d k=%d\n", d, k); ++bad; } }
    }
  std::printf("checked %d shapes, bad %d\n", n, bad);
  return bad != 0;
}
