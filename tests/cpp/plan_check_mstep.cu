// CPU-only property check of the M-step planners (plan_owner_update, plan_tma_update): shared-memory budget, ring
// depths, thread counts and column coverage for every (n_features, n_clusters) on a grid.  Nothing is launched.
#include "centroid_update_tma.cu"
#include <cstdio>
using namespace cb2;
int main() {
  Handle h; h.sm_count = 148; h.smem_optin = 232448; h.cc_major = 10;
  int bad = 0, n = 0, owner = 0, tma = 0, none = 0;
  for (int d = 4; d <= 1024; d += 4)
    for (int k : {1, 2, 7, 8, 15, 16, 17, 31, 32, 33, 64, 65, 100, 128, 129, 200, 255, 256, 257, 300, 511, 512, 513, 1000, 1024, 1030, 1500, 1700, 2048, 4096, 5000, 65536}) {
      ++n;
      OwnerPlan op = plan_owner_update(h, d, k);
      if (op.vec > 0) {
        ++owner;
        bool ok = op.smem <= h.smem_optin && op.tr >= 64 && op.tr <= 256 && op.nstage >= 2 && op.nstage <= MAX_NSTAGE && op.nl > op.nstage && op.nl <= OWN_MAXNL &&
                  op.slices == (d + 32 * op.vec - 1) / (32 * op.vec) && 32 * op.vec <= d && k < 65536;
        if (!ok) { std::printf("OWNER BAD d=%d k=%d vec=%d tr=%d ns=%d nl=%d smem=%zu\n", d, k, op.vec, op.tr, op.nstage, op.nl, op.smem); ++bad; }
        continue;
      }
      if (!tma_update_supported(h, d, k)) { ++none; continue; }
      ++tma;
      TmaUpdatePlan pl = plan_tma_update(h, d, k);
      bool ok = pl.smem <= h.smem_optin && pl.smem * pl.ctas_per_sm <= 228 * 1024 && pl.tr >= 32 && pl.tr <= 256 && pl.nstage >= 2 && pl.nstage <= MAX_NSTAGE &&
                (pl.warps + 1 + pl.na) * 32 <= 1024 && pl.slices * pl.ds >= d;
      if (!ok) { std::printf("TMA BAD d=%d k=%d ds=%d nb=%d tr=%d ns=%d per_sm=%d warps=%d smem=%zu\n", d, k, pl.ds, pl.nb, pl.tr, pl.nstage, pl.ctas_per_sm, pl.warps, pl.smem); ++bad; }
    }
  std::printf("checked %d shapes: owner %d, tma %d, generic %d, bad %d\n", n, owner, tma, none, bad);
  return bad != 0;
}
