// Host-side TMA tensor-map construction through the driver entry point (no link-time libcuda).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace cb2 {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn()
{
  // resolved once (thread-safe initialisation of a function-local static)
  static const EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CB2_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess)
      throw Error(CUML_B200_CUDA_ERROR, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// 2-D fp32 row-major [rows, cols] tensor, box = [box_rows, box_cols]; out-of-bounds -> zeros
inline CUtensorMap make_map_2d(const void* base, uint64_t cols, uint64_t rows, uint64_t row_stride_bytes,
                               uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle swizzle,
                               CUtensorMapL2promotion promo,
                               CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_FLOAT32)
{
  CUtensorMap m;
  cuuint64_t dims[2]    = {cols, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2]     = {box_cols, box_rows};
  cuuint32_t estr[2]    = {1, 1};
  CUresult r = encode_tiled_fn()(&m, dtype, 2, const_cast<void*>(base), dims, strides, box,
                                 estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    throw Error(CUML_B200_CUDA_ERROR,
                "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
  return m;
}

}  // namespace cb2
