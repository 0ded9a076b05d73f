// cuml_b200 internal: launch wrappers implemented in the .cu files.
#pragma once
#include "common.cuh"

namespace cb2 {

// ---- generic SIMT distance kernels (fallback shapes, fp64, seeding, transform) -----------
// labels/mind may be null.  mind = max(0, ||x||^2 + min_j(||c_j||^2 - 2 x.c_j)).
template <typename T>
void simt_assign(Handle& h, const T* X, int64_t n, int d, const T* C, int k, const T* cnorm,
                 int32_t* labels, T* mind);
// min-update form used by seeding: mind[i] = min(mind[i], dist to nearest of C)
template <typename T>
void simt_min_update(Handle& h, const T* X, int64_t n, int d, const T* C, int k, const T* cnorm, T* mind);
template <typename T>
void simt_transform(Handle& h, const T* X, int64_t n, int d, const T* C, int k, const T* cnorm,
                    T* out, bool take_sqrt);
template <typename T>
void row_norms(Handle& h, const T* A, int64_t rows, int d, T* out);  // ||a_i||^2

// ---- tcgen05 fused distance + argmin (fp32 via 3xTF32) -------------------------------------
bool tc_supported(int64_t d, int k);
struct TcCentroids {           // per-iteration operand buffers (hi/lo split + half norms)
  DevBuf<float> hi, lo, cnh;
  DevBuf<uint16_t> hb, lb;   // bf16 copies of hi / lo (correction terms of the CTA-pair kernel)
  DevBuf<float> cnp;         // [k_pad][8]: -1/2||c||^2 as three tf32-exact pieces (folded into the MMA), zeros
  int fold = 0;              // cnp is valid and the plan reserves the fold tiles
  int bf16c = 0;             // hi is rounded to nearest tf32 and hb / lb are valid
  int k_pad = 0, d_pad = 0, block_n = 0;
  int pack = 1;   // rows of X packed side by side into one 128-byte operand row (2 when n_features <= 16)
  int k_sub = 0;  // centroid rows per packed group (k padded to 32/64/128) when pack == 2
};
void tc_prepare(Handle& h, const float* C, int k, int d, TcCentroids& out, bool allow_bf16 = true);
int tc_variant(const Handle& h, int d, int k);   // 1 one-CTA 3xTF32, 2 pair 3xTF32, 3 pair tf32+bf16, 5 one-CTA tf32+bf16 (shared-memory or tensor-memory X operand)
// distance-matrix mode of the same kernels (ML::kmeans::transform): out[i, j] = ||x_i - c_j||^2 (or its sqrt)
struct TcDistOut {
  float* out         = nullptr;   // [n, k] row-major
  const float* xnorm = nullptr;   // [n] ||x_i||^2
  int sqrt           = 0;
};
// fused E + M step (short rows: n_features = 16, k <= 64, unweighted, even row count): the E-step kernel also leaves
// this partition's centroid sums / counts as per-CTA partials [row_blocks][k][d] / [row_blocks][k]; row_blocks == 0 on
// return means the shape took the plain kernel and the caller runs the separate M-step
struct TcMstepOut {
  DevBuf<float>* partial_S = nullptr;
  DevBuf<float>* partial_W = nullptr;
  int row_blocks           = 0;
};
bool tc_fused_update_supported(const Handle& h, int d, int k);
// best_out (optional, [n]): the winning value 1/2||c_label||^2 - x.c per row, i.e. (min distance - ||x||^2) / 2
void tc_assign(Handle& h, const float* X, int64_t n, int d, int k, const TcCentroids& cen,
               int32_t* labels, float* dbg_dots = nullptr, const TcDistOut* dist = nullptr, float* best_out = nullptr,
               TcMstepOut* mstep = nullptr);
bool tc_best_supported(const Handle& h, int d, int k);
bool tc_transform_supported(const Handle& h, int64_t d, int k);

// ---- M-step: centroid sums / weights / exact inertia ----------------------------------------
template <typename T>
struct UpdateWorkspace {
  DevBuf<T> partial_S;        // [row_blocks][k][d]
  DevBuf<T> partial_W;        // [row_blocks][k]
  DevBuf<double> partial_I;   // [row_blocks * slices]
  int row_blocks = 0, slices = 0, ds = 0, warps = 0;
  size_t smem = 0;
  int k = 0, d = 0;
  bool atomic_path = false;
};
template <typename T>
void update_plan(Handle& h, int64_t n_max, int d, int k, UpdateWorkspace<T>& ws);
// accumulate one partition into packed (double [k*d + k + 1], S | W | inertia); `accumulate_into`
// false => packed is overwritten, true => added to (multi-partition).
template <typename T>
void update_accumulate(Handle& h, UpdateWorkspace<T>& ws, const T* X, int64_t n, int d, const int32_t* labels,
                       const T* w, const T* C_old, int k, double* packed, bool accumulate_into,
                       bool sums /* false: inertia only */);
// exact inertia of a labelling wrt C into *cell (fp64), deterministic
template <typename T>
void compute_inertia(Handle& h, const T* X, int64_t n, int d, const int32_t* labels, const T* w, const T* C,
                     double* cell, bool accumulate_into);
// fp32 TMA-staged sums/weights (centroid_update_tma.cu); labels must be readable up to n + 256
bool tma_update_supported(const Handle& h, int d, int k);
void tma_update_accumulate(Handle& h, const float* X, int64_t n, int d, const int32_t* labels_padded, const float* w,
                           int k, DevBuf<float>& partial_S, DevBuf<float>& partial_W, double* packed,
                           bool accumulate_into, const uint8_t* cls_map = nullptr);
const uint8_t* tma_update_balance(Handle& h, const double* W, int d, int k, DevBuf<uint8_t>& map);
// packed[0 .. k*d+k) (+)= fixed-order fp64 sum of per-CTA partials (the tail of every fp32 M-step kernel; also used alone
// after the fused E + M kernel)
void tma_update_reduce(Handle& h, const float* partial_S, const float* partial_W, int row_blocks, int k, int d,
                       double* packed, bool accumulate_into);
// C_new = S/W (W>0) else C_old; shift2 = sum (C_new-C_old)^2 (deterministic, one block)
template <typename T>
void finalize_centroids(Handle& h, const double* packed, T* C, int k, int d, double* shift2_out);

// ---- small utilities -----------------------------------------------------------------------
template <typename T>
void gather_rows(Handle& h, const T* X, int d, const int64_t* idx_dev, int m, T* out);
template <typename T>
double sum_weights(Handle& h, const T* w, int64_t n);  // host result (syncs)
void labels_to_i64(Handle& h, const int32_t* in, int64_t n, int64_t* out);
template <typename T>
void weighted_histogram(Handle& h, const int32_t* labels, const T* w, int64_t n, int k, double* out /*dev, zeroed*/);

}  // namespace cb2
