// Host-side tensor-map helpers shared by the tcgen05 training kernels: the driver's cuTensorMapEncode* entry points are
// resolved through the runtime (no link-time libcuda dependency).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace dpft {
namespace tmah {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Driver {
    EncodeTiledFn encode_tiled = nullptr;
    EncodeIm2colFn encode_im2col = nullptr;
    int driver_version = 0;
    int sm_count = 0;
};

inline Driver& driver() {
    static Driver d;
    return d;
}

inline int resolve_driver() {
    Driver& d = driver();
    if (d.encode_tiled && d.encode_im2col) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    int st = cuda_status(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q), "cuTensorMapEncodeTiled");
    if (st) return st;
    if (!fn) { set_error("cuTensorMapEncodeTiled not available in this driver"); return DPFT_ERR_UNSUPPORTED; }
    d.encode_tiled = (EncodeTiledFn)fn;
    fn = nullptr;
    st = cuda_status(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q), "cuTensorMapEncodeIm2col");
    if (st) return st;
    if (!fn) { set_error("cuTensorMapEncodeIm2col not available in this driver"); return DPFT_ERR_UNSUPPORTED; }
    d.encode_im2col = (EncodeIm2colFn)fn;
    cudaDriverGetVersion(&d.driver_version);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev);
    return 0;
}

inline CUtensorMapDataType dtype16(bool is_f16) { return is_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; }

// [outer][inner] row-major matrix of 16-bit elements, boxes of box_outer rows x box_inner elements, 128-byte swizzle
inline int encode_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t row_bytes, uint32_t box_inner,
                     uint32_t box_outer, bool is_f16) {
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = driver().encode_tiled(map, dtype16(is_f16), 2, const_cast<void*>(base), dims, strides, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(2d) failed with CUresult %d", (int)r); return DPFT_ERR_INVALID_ARGUMENT; }
    return 0;
}

// NHWC activation (B, H, W, C) as an im2col-mode map: boxes of `pixels` output pixels x `channels` channels of one filter tap
// (the tap goes in the instruction's offset operands), traversal stride = the convolution stride.
inline int encode_im2col(CUtensorMap* map, const void* x, int B, int H, int W, int C, int R, int S, int stride, int pad,
                         uint32_t channels, uint32_t pixels, bool is_f16) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    int lower[2] = {-pad, -pad};
    int upper[2] = {pad - (S - 1), pad - (R - 1)};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = driver().encode_im2col(map, dtype16(is_f16), 4, const_cast<void*>(x), dims, strides, lower, upper, channels, pixels,
                                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeIm2col failed with CUresult %d", (int)r); return DPFT_ERR_INVALID_ARGUMENT; }
    // Same small-tensor fix-up CUTLASS applies for drivers <= 13.1 (cute/atom/copy_traits_sm90_im2col.hpp).
    if (driver().driver_version <= 13010 && (uint64_t)B * H * W * C * 2 < 131072) reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);
    return 0;
}

}  // namespace tmah
}  // namespace dpft
