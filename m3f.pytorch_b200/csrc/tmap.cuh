// Host-side CUtensorMap construction.  The two encoders live in libcuda; they are resolved at run time through
// cudaGetDriverEntryPoint so the shared library has no link-time dependency on the driver (it must load on the
// GPU-less build box, where only the symbol table is checked).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace m3t {

typedef CUresult (*PFN_tmapTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_tmapIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TmapApi {
  PFN_tmapTiled tiled = nullptr;
  PFN_tmapIm2col im2col = nullptr;
  int driver_version = 0;
  bool ok = false;
};

inline const TmapApi& tmap_api() {
  static TmapApi api = [] {
    TmapApi a;
    cudaDriverEntryPointQueryResult q;
    void* f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      a.tiled = reinterpret_cast<PFN_tmapTiled>(f);
    f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      a.im2col = reinterpret_cast<PFN_tmapIm2col>(f);
    cudaDriverGetVersion(&a.driver_version);
    a.ok = a.tiled && a.im2col;
    return a;
  }();
  return api;
}

// Drivers <= 13.1 mis-set one descriptor bit for tensors smaller than 128 KiB (same fix-up CUTLASS applies).
inline void tmap_small_tensor_fixup(CUtensorMap* tm, uint64_t total_bytes) {
  if (tmap_api().driver_version <= 13010 && total_bytes < 131072)
    reinterpret_cast<uint64_t*>(tm)[1] &= ~(1ull << 21);
}

inline CUtensorMapSwizzle swizzle_enum(int bytes) {
  return bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
         : bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
         : bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                       : CU_TENSOR_MAP_SWIZZLE_NONE;
}

// Generic tiled map over bf16 elements.  dims[0] is the contiguous dimension; strides_bytes[i] is the byte
// stride of dims[i+1].
inline int make_tmap_tiled_bf16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                                const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  const TmapApi& api = tmap_api();
  if (!api.ok) return -10;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  uint64_t span = 2;
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    if (gstr[i] % 16) return -11;
  }
  span = (rank > 1 ? strides_bytes[rank - 2] * dims[rank - 1] : dims[0] * 2);
  if (reinterpret_cast<uintptr_t>(base) % 16) return -12;
  CUresult r = api.tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr,
                         bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_enum(swizzle_bytes),
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "m3t: cuTensorMapEncodeTiled failed (%d) rank=%d dims=[%llu,%llu,..] box=[%u,%u,..]\n", (int)r,
            rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0],
            rank > 1 ? box[1] : 0);
    return -13;
  }
  tmap_small_tensor_fixup(tm, span);
  return 0;
}

// 2-D convenience: [outer][inner] row-major bf16 matrix with row stride ld (elements).
inline int make_tmap_2d_bf16(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t ld,
                             uint32_t box_inner, uint32_t box_outer, int swizzle_bytes = 128) {
  uint64_t dims[2] = {inner, outer};
  uint64_t str[1] = {ld * 2};
  uint32_t box[2] = {box_inner, box_outer};
  return make_tmap_tiled_bf16(tm, base, 2, dims, str, box, swizzle_bytes);
}

// im2col map over a channels-last bf16 activation tensor.  dims = (C, W, [H, [D,]] N), contiguous;
// lower/upper: bounding-box corners per spatial dim in (W, H, D) order; conv_stride likewise.
inline int make_tmap_im2col_bf16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                                 const int* lower, const int* upper, const int* conv_stride,
                                 uint32_t channels_per_pixel, uint32_t pixels_per_column, int swizzle_bytes = 128) {
  const TmapApi& api = tmap_api();
  if (!api.ok) return -10;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t es[5];
  uint64_t s = 2;
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    s *= dims[i];
    if (i + 1 < rank) {
      gstr[i] = s;
      if (s % 16) return -11;
    }
    es[i] = (i >= 1 && i <= rank - 2) ? (cuuint32_t)conv_stride[i - 1] : 1u;
  }
  if (reinterpret_cast<uintptr_t>(base) % 16) return -12;
  CUresult r = api.im2col(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim,
                          gstr, lower, upper, channels_per_pixel, pixels_per_column, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_enum(swizzle_bytes),
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "m3t: cuTensorMapEncodeIm2col failed (%d) rank=%d dims=[%llu,%llu,%llu,..] lower=%d upper=%d\n",
            (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
            lower[0], upper[0]);
    return -14;
  }
  tmap_small_tensor_fixup(tm, s);
  return 0;
}

}  // namespace m3t
