// tmap.cuh — host-side helper: encode a 3-D tiled CUtensorMap without linking libcuda
// (the encoder is fetched from the driver at run time).
#pragma once
#include <cuda.h>   // CUtensorMap (types only)

#include "common.cuh"

namespace sb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) {
    set_error("cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

// dims (d0 fastest), dense strides, box {b0, b1, 1}; out-of-bounds elements read as 0.
static inline int make_map_3d_ex(CUtensorMap* m, CUtensorMapDataType dt, int elt_bytes, const void* base,
                                 unsigned long long d0, unsigned long long d1, unsigned long long d2,
                                 unsigned b0, unsigned b1, CUtensorMapSwizzle swz,
                                 CUtensorMapL2promotion promo, const char* what,
                                 unsigned long long pitch0 = 0) {   // row pitch in elements (0: dense, = d0)
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return SB_ECUDA;
  cuuint64_t dims[3] = {d0, d1, d2};
  const unsigned long long p0 = pitch0 ? pitch0 : d0;
  cuuint64_t strides[2] = {p0 * (unsigned long long)elt_bytes, p0 * d1 * (unsigned long long)elt_bytes};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, dt, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(%s) failed with CUresult %d (dims %llu x %llu x %llu)", what,
              (int)r, d0, d1, d2);
    return SB_ECUDA;
  }
  return SB_OK;
}

}  // namespace sb
