// Shared helpers for the dpft_b200 sm_100a kernels (device-side type traits, packed loads, error plumbing).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dpft_b200.h"

namespace dpft {

// ---- error plumbing (thread-local message behind dpft_last_error()) ------------------------------------------
void set_error(const char* fmt, ...);
int cuda_status(cudaError_t e, const char* what);  // 0 on success, else records + returns (int)e

// conv3x3_halo.cu: the 64 -> 64 channel 3x3 convolution over a shared-memory halo tile (dispatched from dpft_conv2d_nhwc)
bool conv3x3_halo_eligible(int H, int W, int Cin, int Cout, int R, int S, int stride, int pad, bool has_residual);
int conv3x3_halo_launch(const void* x, const void* w, const float* bias, void* y, int B, int H, int W, int relu, bool is_f16,
                        cudaStream_t stream);

#define DPFT_REQUIRE(cond, ...)                      \
    do {                                             \
        if (!(cond)) {                               \
            ::dpft::set_error(__VA_ARGS__);          \
            return DPFT_ERR_INVALID_ARGUMENT;        \
        }                                            \
    } while (0)

#define DPFT_LAUNCH_CHECK(what)                                         \
    do {                                                                \
        int _st = ::dpft::cuda_status(cudaGetLastError(), what);        \
        if (_st) return _st;                                            \
    } while (0)

// ---- arithmetic type traits ----------------------------------------------------------------------------------
template <typename T> struct AccOf { using type = float; };
template <> struct AccOf<double> { using type = double; };

template <typename T> __device__ __forceinline__ typename AccOf<T>::type to_acc(T v) { return v; }
template <> __device__ __forceinline__ float to_acc<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_acc<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_acc(typename AccOf<T>::type v) { return (T)v; }
template <> __device__ __forceinline__ __half from_acc<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_acc<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---- packed (vector) global loads/stores ---------------------------------------------------------------------
template <typename T, int VEC> struct alignas(sizeof(T) * VEC) Pack { T v[VEC]; };

template <int BYTES> struct RawOf;
template <> struct RawOf<2> { using type = unsigned short; };
template <> struct RawOf<4> { using type = unsigned int; };
template <> struct RawOf<8> { using type = uint2; };
template <> struct RawOf<16> { using type = uint4; };

// Read-only (ld.global.nc) load of VEC contiguous elements; p must be aligned to sizeof(T)*VEC.
template <typename T, int VEC> __device__ __forceinline__ Pack<T, VEC> ldg_pack(const T* p) {
    using Raw = typename RawOf<sizeof(T) * VEC>::type;
    union { Raw r; Pack<T, VEC> k; } u;
    u.r = __ldg(reinterpret_cast<const Raw*>(p));
    return u.k;
}
// Raw (untyped) variant + a scheduling pin: `pin_loaded` emits no instruction but makes the compiler treat the loaded
// registers as consumed/redefined at that point, so every conversion of the loaded bits is scheduled after it.  Used to
// keep a batch of independent gather loads back to back instead of interleaved with the unpacking of the first ones.
template <typename T, int VEC> __device__ __forceinline__ typename RawOf<sizeof(T) * VEC>::type ldg_raw(const T* p) {
    using Raw = typename RawOf<sizeof(T) * VEC>::type;
    return __ldg(reinterpret_cast<const Raw*>(p));
}
template <typename T, int VEC> __device__ __forceinline__ Pack<T, VEC> unpack_raw(const typename RawOf<sizeof(T) * VEC>::type& r) {
    using Raw = typename RawOf<sizeof(T) * VEC>::type;
    union { Raw rr; Pack<T, VEC> k; } u;
    u.rr = r;
    return u.k;
}
__device__ __forceinline__ void pin_loaded(uint4& r) { asm volatile("" : "+r"(r.x), "+r"(r.y), "+r"(r.z), "+r"(r.w)); }
__device__ __forceinline__ void pin_loaded(uint2& r) { asm volatile("" : "+r"(r.x), "+r"(r.y)); }
__device__ __forceinline__ void pin_loaded(unsigned int& r) { asm volatile("" : "+r"(r)); }
__device__ __forceinline__ void pin_loaded(unsigned short& r) { asm volatile("" : "+h"(r)); }
__device__ __forceinline__ void issue_barrier() { asm volatile("" ::: "memory"); }
template <typename Raw> __device__ __forceinline__ Raw raw_zero() { Raw r; memset(&r, 0, sizeof(Raw)); return r; }

template <typename T, int VEC> __device__ __forceinline__ void st_pack(T* p, const Pack<T, VEC>& k) {
    using Raw = typename RawOf<sizeof(T) * VEC>::type;
    union { Raw r; Pack<T, VEC> kk; } u;
    u.kk = k;
    *reinterpret_cast<Raw*>(p) = u.r;
}

// Vector reductions into global memory (red.global.add.{f32,v2.f32,v4.f32,f64}); sm_90+ has the vector forms.
template <int VEC> __device__ __forceinline__ void red_add(float* p, const float* v) {
    if constexpr (VEC % 4 == 0) {
#pragma unroll
        for (int k = 0; k < VEC; k += 4)
            atomicAdd(reinterpret_cast<float4*>(p + k), make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]));
    } else if constexpr (VEC == 2) {
        atomicAdd(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]));
    } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) atomicAdd(p + k, v[k]);
    }
}
template <int VEC> __device__ __forceinline__ void red_add(double* p, const double* v) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) atomicAdd(p + k, v[k]);
}

__host__ __device__ constexpr int ilog2_floor(int x) { return x <= 1 ? 0 : 1 + ilog2_floor(x >> 1); }

}  // namespace dpft
