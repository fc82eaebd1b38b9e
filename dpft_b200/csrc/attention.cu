// Decoder self-attention as a tcgen05 flash-attention kernel for sm_100a.
//
// Replaces the scaled-dot-product core of nn.MultiheadAttention in the reference's decoder layer
// (src/dprt/models/fusers/mpfusion.py:56-57 construction, :122-148 forward_self_attn: q = k = query + pos, v = query,
// need_weights=False, eval mode -> no attention dropout):
//
//     out[b, n, h, :] = softmax_j( scale * <q[b,n,h,:], k[b,j,h,:]> ) . v[b,j,h,:]
//
// One CTA per (128-query tile, head, sample), 128 threads, thread t <-> query row t <-> TMEM lane t.
//   S  = Q K^T   tcgen05.mma.cta_group::1.kind::f16, M = 128 queries, N = 128 keys, K = head dim (zero-padded to 16/32/64)
//   online softmax in registers (exp2 domain), P written to shared memory as the A operand of the second product
//   Oj = P V     M = 128, N = padded head dim, K = 128 keys; the running output is rescaled in registers
// Operands live in shared memory in the un-swizzled K-major core-matrix layout ([K chunk of 8][row][16 B]); the threads
// build them from global memory (fp32 or 16-bit, arbitrary token stride) so no tensor map is needed for these small tiles.
//
// "precise" mode (fp32 inputs): every operand x is split into two halves x = hi + lo (hi = f16(x), lo = f16(x - hi)) and
// each product is issued as three MMAs (hi.hi + hi.lo + lo.hi) into the same fp32 accumulator, which carries ~22 mantissa
// bits through the tensor cores: the result matches an fp32 reference to ~1e-6, so the kernel can stand in for the fp32
// module-by-module path (1e-3 bar) and not only for the 16-bit tier.
#include "common.cuh"
#include "tcgen05.cuh"

namespace dpft {
namespace {
using namespace tc;

constexpr int ATT_BM = 128;     // queries per CTA
constexpr int ATT_BN = 128;     // keys per block
constexpr float kLog2e = 1.4426950408889634f;

template <typename MT> __device__ __forceinline__ unsigned short to_bits(float x);
template <> __device__ __forceinline__ unsigned short to_bits<__half>(float x) {
    unsigned short r;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(x));
    return r;
}
template <> __device__ __forceinline__ unsigned short to_bits<__nv_bfloat16>(float x) {
    return __bfloat16_as_ushort(__float2bfloat16_rn(x));
}
template <typename MT> __device__ __forceinline__ float from_bits(unsigned short b);
template <> __device__ __forceinline__ float from_bits<__half>(unsigned short b) { return __half2float(__ushort_as_half(b)); }
template <> __device__ __forceinline__ float from_bits<__nv_bfloat16>(unsigned short b) {
    return __bfloat162float(__ushort_as_bfloat16(b));
}

// hi / lo halves of 8 consecutive values -> two 16-byte core-matrix rows
template <typename MT, bool SPLIT> __device__ __forceinline__ void pack8(const float* x, uint4& hi, uint4& lo) {
    unsigned short h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        h[i] = to_bits<MT>(x[i]);
        l[i] = SPLIT ? to_bits<MT>(x[i] - from_bits<MT>(h[i])) : (unsigned short)0;
    }
    hi = make_uint4(h[0] | (uint32_t)h[1] << 16, h[2] | (uint32_t)h[3] << 16, h[4] | (uint32_t)h[5] << 16, h[6] | (uint32_t)h[7] << 16);
    lo = make_uint4(l[0] | (uint32_t)l[1] << 16, l[2] | (uint32_t)l[3] << 16, l[4] | (uint32_t)l[5] << 16, l[6] | (uint32_t)l[7] << 16);
}

struct AttnParams {
    const void *q, *k, *v;
    void* out;
    int B, H, N, D;
    long long q_row, k_row, v_row;          // elements between consecutive tokens
    long long q_batch, k_batch, v_batch;    // elements between consecutive samples
    float scale_log2e;
};

// Shared-memory map (bytes).  Operand tile with R rows and K elements: [K/8][R][16 B]  ->  LBO = R*16, SBO = 128.
template <int DP, bool SPLIT> struct AttnSmem {
    static constexpr int kHalves = SPLIT ? 2 : 1;
    static constexpr int kQ = ATT_BM * DP * 2;            // one half of Q  [DP/8][128][16]
    static constexpr int kK = ATT_BN * DP * 2;            // one half of K  [DP/8][128][16]
    static constexpr int kV = DP * ATT_BN * 2;            // one half of V^T [128/8][DP][16]
    static constexpr int kP = ATT_BM * ATT_BN * 2;        // one half of P  [128/8][128][16]
    static constexpr int oQ = 0;
    static constexpr int oK = oQ + kHalves * kQ;
    static constexpr int oV = oK + kHalves * kK;
    static constexpr int oP = oV + kHalves * kV;
    static constexpr int oBar = oP + kHalves * kP;
    static constexpr int kTotal = oBar + 32;
};

// 16 bytes of T -> floats
template <typename T> struct Vec16 { static constexpr int kElems = 16 / sizeof(T); };
template <typename T> __device__ __forceinline__ void unpack16(const uint4& raw, float* x) {
    const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
    for (int i = 0; i < Vec16<T>::kElems; ++i) x[i] = to_acc<T>(e[i]);
}

template <typename T, bool SPLIT, int DP, bool VECLOAD>
__global__ void __launch_bounds__(128) flash_attn_fwd_kernel(const AttnParams prm) {
    using MT = typename std::conditional<std::is_same<T, __nv_bfloat16>::value, __nv_bfloat16, __half>::type;
    using L = AttnSmem<DP, SPLIT>;
    constexpr bool kIsF16 = std::is_same<MT, __half>::value;
    constexpr int kHalves = L::kHalves;
    constexpr int DC = DP / 8;                      // 16-byte chunks along the head dimension
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bar_s = reinterpret_cast<uint64_t*>(smem + L::oBar);
    uint64_t* bar_o = bar_s + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_s + 2);

    const int t = threadIdx.x, warp = t >> 5;
    const int n0 = blockIdx.x * ATT_BM, h = blockIdx.y, b = blockIdx.z;
    const int N = prm.N, D = prm.D;
    const T* qg = reinterpret_cast<const T*>(prm.q) + (long long)b * prm.q_batch + (long long)h * D;
    const T* kg = reinterpret_cast<const T*>(prm.k) + (long long)b * prm.k_batch + (long long)h * D;
    const T* vg = reinterpret_cast<const T*>(prm.v) + (long long)b * prm.v_batch + (long long)h * D;

    if (t == 0) {
        mbar_init(bar_s, 1);
        mbar_init(bar_o, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 256);      // columns [0,128) = S, [128, 128+DP) = P.V of the current key block
    // ---- operand tiles -> shared memory.  Rows past N and channels past D are zero.
    // VECLOAD (D a multiple of 16 bytes of T, aligned strides): consecutive threads fetch consecutive 16-byte pieces of a row
    // (coalesced), otherwise thread t walks row t element by element (the 2-channel heads of the shipped configuration).
    constexpr int VE = Vec16<T>::kElems;                      // elements per 16-byte global load
    auto load_rows = [&](const T* src, long long row_stride, int first, int off) {   // [128 rows][DP] -> [DP/8][128][16 B]
        if constexpr (VECLOAD) {
#pragma unroll 2
            for (int idx = t; idx < 128 * (DP / VE); idx += 128) {
                const int r = idx / (DP / VE), cc = idx - r * (DP / VE);
                const int j = first + r, d0 = cc * VE;
                uint4 raw = make_uint4(0u, 0u, 0u, 0u);
                if (j < N && d0 < D) raw = __ldg(reinterpret_cast<const uint4*>(src + (long long)j * row_stride + d0));
                uint8_t* dst = smem + off + (d0 >> 3) * (128 * 16) + r * 16 + (d0 & 7) * 2;
                if constexpr (std::is_same<T, MT>::value) {
                    *reinterpret_cast<uint4*>(dst) = raw;     // already the MMA type: a straight copy
                } else {
                    float x[VE];
                    unpack16<T>(raw, x);
                    unsigned short hi[VE], lo[VE];
#pragma unroll
                    for (int i = 0; i < VE; ++i) {
                        hi[i] = to_bits<MT>(x[i]);
                        lo[i] = SPLIT ? to_bits<MT>(x[i] - from_bits<MT>(hi[i])) : (unsigned short)0;
                    }
                    *reinterpret_cast<uint2*>(dst) = make_uint2(hi[0] | (uint32_t)hi[1] << 16, hi[2] | (uint32_t)hi[3] << 16);
                    if (SPLIT)
                        *reinterpret_cast<uint2*>(dst + 128 * DP * 2) =
                            make_uint2(lo[0] | (uint32_t)lo[1] << 16, lo[2] | (uint32_t)lo[3] << 16);
                }
            }
        } else {
            const int j = first + t;
#pragma unroll
            for (int c = 0; c < DC; ++c) {
                float x[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int d = c * 8 + i;
                    x[i] = (j < N && d < D) ? to_acc<T>(src[(long long)j * row_stride + d]) : 0.0f;
                }
                uint4 hi, lo;
                pack8<MT, SPLIT>(x, hi, lo);
                *reinterpret_cast<uint4*>(smem + off + c * (128 * 16) + t * 16) = hi;
                if (SPLIT) *reinterpret_cast<uint4*>(smem + off + 128 * DP * 2 + c * (128 * 16) + t * 16) = lo;
            }
        }
    };
    auto load_k = [&](int j0) { load_rows(kg, prm.k_row, j0, L::oK); };
    auto load_v = [&](int j0) {                     // V^T: row d, K index = key  ->  [128/8][DP][16 B]
        if constexpr (VECLOAD) {
#pragma unroll 2
            for (int idx = t; idx < 128 * (DP / VE); idx += 128) {
                const int r = idx / (DP / VE), cc = idx - r * (DP / VE);
                const int j = j0 + r, d0 = cc * VE;
                uint4 raw = make_uint4(0u, 0u, 0u, 0u);
                if (j < N && d0 < D) raw = __ldg(reinterpret_cast<const uint4*>(vg + (long long)j * prm.v_row + d0));
                float x[VE];
                unpack16<T>(raw, x);
                uint8_t* base = smem + L::oV + (r >> 3) * (DP * 16) + (r & 7) * 2;
#pragma unroll
                for (int i = 0; i < VE; ++i) {
                    const unsigned short hi = to_bits<MT>(x[i]);
                    *reinterpret_cast<unsigned short*>(base + (d0 + i) * 16) = hi;
                    if (SPLIT) *reinterpret_cast<unsigned short*>(base + L::kV + (d0 + i) * 16) = to_bits<MT>(x[i] - from_bits<MT>(hi));
                }
            }
        } else {
            const int j = j0 + t;
            uint8_t* base = smem + L::oV + (t >> 3) * (DP * 16) + (t & 7) * 2;
#pragma unroll 4
            for (int d = 0; d < DP; ++d) {
                const float x = (j < N && d < D) ? to_acc<T>(vg[(long long)j * prm.v_row + d]) : 0.0f;
                const unsigned short hi = to_bits<MT>(x);
                *reinterpret_cast<unsigned short*>(base + d * 16) = hi;
                if (SPLIT) *reinterpret_cast<unsigned short*>(base + L::kV + d * 16) = to_bits<MT>(x - from_bits<MT>(hi));
            }
        }
    };
    load_rows(qg, prm.q_row, n0, L::oQ);
    load_k(0);
    load_v(0);
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);      // this warp's TMEM lane quadrant
    const uint32_t sQ = smem_u32(smem + L::oQ), sK = smem_u32(smem + L::oK), sV = smem_u32(smem + L::oV), sP = smem_u32(smem + L::oP);
    constexpr uint32_t idesc_s = umma_idesc_16bit(ATT_BM, ATT_BN, kIsF16);
    constexpr uint32_t idesc_o = umma_idesc_16bit(ATT_BM, DP, kIsF16);

    float o[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) o[d] = 0.0f;
    float m_run = -INFINITY, l_run = 0.0f;
    uint32_t phase = 0;
    const int n_blocks = (N + ATT_BN - 1) / ATT_BN;

    for (int jb = 0; jb < n_blocks; ++jb) {
        const int j0 = jb * ATT_BN;
        // ---- S = Q K^T ----
        if (t == 0) {
            uint32_t acc = 0;
#pragma unroll
            for (int kk = 0; kk < DP / 16; ++kk) {
#pragma unroll
                for (int term = 0; term < (SPLIT ? 3 : 1); ++term) {
                    const uint32_t qa = sQ + (term == 2 ? L::kQ : 0) + kk * 2 * (ATT_BM * 16);
                    const uint32_t kb = sK + (term == 1 ? L::kK : 0) + kk * 2 * (ATT_BN * 16);
                    umma_bf16(tmem_base, umma_desc_noswizzle(qa, ATT_BM * 16, 128), umma_desc_noswizzle(kb, ATT_BN * 16, 128), idesc_s, acc);
                    acc = 1;
                }
            }
            umma_commit(bar_s);
        }
        mbar_wait(bar_s, phase);
        tcgen05_fence_after();
        // ---- online softmax over this block's 128 scores of row t ----
        float m_blk = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < ATT_BN / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(lane_addr + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (j0 + c * 32 + i < N) m_blk = fmaxf(m_blk, __uint_as_float(v[i]));
        }
        const float m_new = fmaxf(m_run, m_blk * prm.scale_log2e);     // scale > 0: max commutes with the scaling
        const float alpha = exp2f(m_run - m_new);                      // first block: exp2(-inf) = 0
        float l_blk = 0.0f;
#pragma unroll 1
        for (int c = 0; c < ATT_BN / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(lane_addr + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float p[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int j = j0 + c * 32 + g * 8 + i;
                    p[i] = j < N ? exp2f(__uint_as_float(v[g * 8 + i]) * prm.scale_log2e - m_new) : 0.0f;
                    l_blk += p[i];
                }
                uint4 hi, lo;
                pack8<MT, SPLIT>(p, hi, lo);
                const int chunk = c * 4 + g;
                *reinterpret_cast<uint4*>(smem + L::oP + chunk * (ATT_BM * 16) + t * 16) = hi;
                if (SPLIT) *reinterpret_cast<uint4*>(smem + L::oP + L::kP + chunk * (ATT_BM * 16) + t * 16) = lo;
            }
        }
        l_run = l_run * alpha + l_blk;
        m_run = m_new;
        fence_proxy_async();
        tcgen05_fence_before();
        __syncthreads();                              // P complete; every thread is done reading S
        // ---- Oj = P V ----
        if (t == 0) {
            tcgen05_fence_after();
            uint32_t acc = 0;
#pragma unroll
            for (int kk = 0; kk < ATT_BN / 16; ++kk) {
#pragma unroll
                for (int term = 0; term < (SPLIT ? 3 : 1); ++term) {
                    const uint32_t pa = sP + (term == 2 ? L::kP : 0) + kk * 2 * (ATT_BM * 16);
                    const uint32_t vb = sV + (term == 1 ? L::kV : 0) + kk * 2 * (DP * 16);
                    umma_bf16(tmem_base + ATT_BN, umma_desc_noswizzle(pa, ATT_BM * 16, 128), umma_desc_noswizzle(vb, DP * 16, 128), idesc_o, acc);
                    acc = 1;
                }
            }
            umma_commit(bar_o);
        }
        if (jb + 1 < n_blocks) load_k(j0 + ATT_BN);   // the K buffer is free (S is complete): overlap with the P.V product
        mbar_wait(bar_o, phase);
        tcgen05_fence_after();
#pragma unroll
        for (int c = 0; c < DP / 16; ++c) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(lane_addr + ATT_BN + c * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[c * 16 + i] = o[c * 16 + i] * alpha + __uint_as_float(v[i]);
        }
        if (jb + 1 < n_blocks) {
            load_v(j0 + ATT_BN);                      // V and P buffers are free now
            fence_proxy_async();
        }
        tcgen05_fence_before();
        __syncthreads();                              // operands of the next block visible; Oj consumed by every thread
        tcgen05_fence_after();
        phase ^= 1;
    }
    // ---- normalise and store out[b, n, h, :] ((B, N, H, D) contiguous) ----
    const int n = n0 + t;
    if (n < N) {
        const float inv = 1.0f / l_run;
        T* og = reinterpret_cast<T*>(prm.out) + (((long long)b * N + n) * prm.H + h) * D;
#pragma unroll
        for (int d = 0; d < DP; ++d)
            if (d < D) og[d] = from_acc<T>(o[d] * inv);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

template <typename T, bool SPLIT, int DP, bool VECLOAD> int launch_attn(const AttnParams& prm, cudaStream_t stream) {
    using L = AttnSmem<DP, SPLIT>;
    auto kern = flash_attn_fwd_kernel<T, SPLIT, DP, VECLOAD>;
    static bool configured = false;
    if (!configured) {
        int st = cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal),
                             "cudaFuncSetAttribute(flash_attn_fwd_kernel)");
        if (st) return st;
        configured = true;
    }
    dim3 grid((prm.N + ATT_BM - 1) / ATT_BM, prm.H, prm.B);
    kern<<<grid, 128, L::kTotal, stream>>>(prm);
    DPFT_LAUNCH_CHECK("flash_attn_fwd_kernel");
    return DPFT_OK;
}

template <typename T, bool SPLIT> int dispatch_dp(const AttnParams& prm, cudaStream_t stream) {
    // 16-byte row pieces: head dim, strides and base addresses must all be multiples of 16 bytes of T
    constexpr long long ve = 16 / sizeof(T);
    const bool vec = prm.D % ve == 0 && prm.q_row % ve == 0 && prm.k_row % ve == 0 && prm.v_row % ve == 0 &&
                     prm.q_batch % ve == 0 && prm.k_batch % ve == 0 && prm.v_batch % ve == 0 &&
                     (((uintptr_t)prm.q | (uintptr_t)prm.k | (uintptr_t)prm.v) & 15) == 0;
    if (prm.D <= 16) return vec ? launch_attn<T, SPLIT, 16, true>(prm, stream) : launch_attn<T, SPLIT, 16, false>(prm, stream);
    if (prm.D <= 32) return vec ? launch_attn<T, SPLIT, 32, true>(prm, stream) : launch_attn<T, SPLIT, 32, false>(prm, stream);
    return vec ? launch_attn<T, SPLIT, 64, true>(prm, stream) : launch_attn<T, SPLIT, 64, false>(prm, stream);
}

}  // namespace
}  // namespace dpft

using namespace dpft;

extern "C" int dpft_self_attention_forward(const void* q, const void* k, const void* v, void* out, int B, int H, int N, int D,
                                           long long q_row_stride, long long k_row_stride, long long v_row_stride,
                                           long long q_batch_stride, long long k_batch_stride, long long v_batch_stride,
                                           float scale, int dtype, int precise, void* stream) {
    DPFT_REQUIRE(q && k && v && out, "self_attention: null pointer");
    DPFT_REQUIRE(B > 0 && H > 0 && N > 0 && D > 0 && D <= 64, "self_attention: bad size B=%d H=%d N=%d D=%d (head dim <= 64)", B, H, N, D);
    DPFT_REQUIRE(B <= 65535 && H <= 65535, "self_attention: B and H must fit the grid (<= 65535)");
    DPFT_REQUIRE(scale > 0.0f, "self_attention: scale must be positive");
    DPFT_REQUIRE(q_row_stride >= (long long)H * D && k_row_stride >= (long long)H * D && v_row_stride >= (long long)H * D,
                 "self_attention: token strides must cover H*D elements");
    AttnParams prm{q, k, v, out, B, H, N, D, q_row_stride, k_row_stride, v_row_stride, q_batch_stride, k_batch_stride,
                   v_batch_stride, scale * kLog2e};
    cudaStream_t s = (cudaStream_t)stream;
    switch (dtype) {
        case DPFT_F32:
            return precise ? dispatch_dp<float, true>(prm, s) : dispatch_dp<float, false>(prm, s);
        case DPFT_F16:
            return dispatch_dp<__half, false>(prm, s);
        case DPFT_BF16:
            return dispatch_dp<__nv_bfloat16, false>(prm, s);
        default:
            set_error("self_attention: dtype must be DPFT_F32, DPFT_F16 or DPFT_BF16");
            return DPFT_ERR_INVALID_ARGUMENT;
    }
}
