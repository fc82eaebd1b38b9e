// Weight gradient of the backbone convolutions as a tcgen05 GEMM over the pixel dimension (sm_100a).
//
// Training counterpart of conv_tcgen05.cu: autograd's convolution-backward-weight for the torchvision Bottleneck convs the
// reference trains (src/dprt/models/backbones/resnet.py:54-55,101; optimiser step at src/dprt/training/trainer.py:125-133).
//
//   dW[n, r, s, c] += sum_m dY[m, n] * X[b, p*stride - pad + r, q*stride - pad + s, c]        m = (b, p, q) linearised
//
// As a GEMM per filter tap: D[co (M=128)][ci (N=BLOCK_N)] = sum over pixels (K).  Both operands are "MN-major" for the
// tensor core (channels contiguous, the reduction index = pixel is the slow dimension), which is exactly what TMA delivers:
//   A  dY tile   [64 pixels][64 channels] boxes of the [M, Cout] gradient matrix, two per stage (128 output channels);
//   B  X  tile   [64 pixels][64 channels] boxes through an im2col-mode tensor map (the tap in the offset operands) or, for
//                1x1/stride-1 layers, the plain [M, Cin] matrix; BLOCK_N / 64 per stage.
// 128-byte swizzle; UMMA descriptors: leading byte offset = 8 KB (next 64-channel box), stride byte offset = 1 KB (next
// group of 8 pixels), +2 KB per K=16 step.
// The pixel range is split across work items (split-K): a persistent CTA per SM walks items (split, co tile, tap, ci tile),
// accumulates in TMEM (double buffered) and adds its partial tile into the fp32 gradient with red.global.add.v4.f32.
//   warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue.
#include <cuda.h>

#include "common.cuh"
#include "tcgen05.cuh"
#include "tma_host.cuh"

namespace dpft {
namespace {

using namespace tc;

constexpr int WG_M = 128;            // output channels per tile
constexpr int WG_KPIX = 64;          // pixels per pipeline stage
constexpr int WG_BOX_BYTES = WG_KPIX * 128;   // one [64 pixels][64 channels] box
constexpr int WG_THREADS = 192;

struct WgradParams {
    int Mpix, P, Q;
    int Cout, Cin;
    int taps, taps_s;
    int stride, pad;
    int im2col;
    int is_f16;
    int co_tiles, ci_tiles, splits, kblocks_total, kb_per_split;
    float* dw;
};

template <int BLOCK_N, int STAGES> struct WgSmem {
    static constexpr int kABytes = 2 * WG_BOX_BYTES;
    static constexpr int kBBytes = (BLOCK_N / 64) * WG_BOX_BYTES;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarrierBytes = (2 * STAGES + 4) * 8 + 16;
    static constexpr int kTotal = STAGES * kStageBytes + kBarrierBytes;
};

// MN-major, 128-byte-swizzled operand: 64-element (128 B) rows along MN, 8-row (K) groups `sbo` bytes apart, the next
// 64-element MN block `lbo` bytes away.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                  // version = 1 (Blackwell)
    d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
    return d;
}

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x, const WgradParams prm) {
    using L = WgSmem<BLOCK_N, STAGES>;
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * L::kABytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * L::kStageBytes);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr uint32_t kTmemCols = 2 * BLOCK_N;

    if (warp == 0 && elect_one()) {
        prefetch_tmap(&tmap_dy);
        prefetch_tmap(&tmap_x);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4);
        }
        fence_barrier_init();
    } else if (warp == 1) {
        tmem_alloc(tmem_slot, kTmemCols);
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int inner_items = prm.co_tiles * prm.taps * prm.ci_tiles;
    const int num_items = inner_items * prm.splits;

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            const int pq = prm.P * prm.Q;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const int split = item / inner_items;
                int rem = item - split * inner_items;
                const int co_tile = rem / (prm.taps * prm.ci_tiles);
                rem -= co_tile * prm.taps * prm.ci_tiles;
                const int tap = rem / prm.ci_tiles, ci_tile = rem - tap * prm.ci_tiles;
                const int r = tap / prm.taps_s, s = tap - r * prm.taps_s;
                const int kb0 = split * prm.kb_per_split;
                const int kb1 = min(kb0 + prm.kb_per_split, prm.kblocks_total);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], L::kStageBytes);
                    uint8_t* a_dst = smem_a + stage * L::kABytes;
                    uint8_t* b_dst = smem_b + stage * L::kBBytes;
                    const int m0 = kb * WG_KPIX;
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        tma_load_2d(&tmap_dy, &full_bar[stage], a_dst + j * WG_BOX_BYTES, co_tile * WG_M + j * 64, m0);
                    if (prm.im2col) {
                        const int cn = m0 / pq;
                        const int rr = m0 - cn * pq;
                        const int p = rr / prm.Q, q = rr - p * prm.Q;
                        const int cw = q * prm.stride - prm.pad, ch = p * prm.stride - prm.pad;
#pragma unroll
                        for (int j = 0; j < BLOCK_N / 64; ++j)
                            tma_load_im2col_4d(&tmap_x, &full_bar[stage], b_dst + j * WG_BOX_BYTES, ci_tile * BLOCK_N + j * 64, cw, ch,
                                               cn, (uint16_t)s, (uint16_t)r);
                    } else {
#pragma unroll
                        for (int j = 0; j < BLOCK_N / 64; ++j)
                            tma_load_2d(&tmap_x, &full_bar[stage], b_dst + j * WG_BOX_BYTES, ci_tile * BLOCK_N + j * 64, m0);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ====================================== MMA issuer ======================================
        const uint32_t idesc = tc::umma_idesc_16bit(WG_M, BLOCK_N, prm.is_f16 != 0) | (1u << 15) | (1u << 16);   // A, B MN-major
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            const int split = item / inner_items;
            const int kb0 = split * prm.kb_per_split;
            const int kb1 = min(kb0 + prm.kb_per_split, prm.kblocks_total);
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tcgen05_fence_after();
                if (elect_one()) {
                    const uint32_t a_addr = smem_u32(smem_a + stage * L::kABytes);
                    const uint32_t b_addr = smem_u32(smem_b + stage * L::kBBytes);
#pragma unroll
                    for (int k = 0; k < WG_KPIX / 16; ++k) {
                        const uint64_t adesc = umma_desc_mn_sw128(a_addr + k * 2048, WG_BOX_BYTES, 1024);
                        const uint64_t bdesc = umma_desc_mn_sw128(b_addr + k * 2048, WG_BOX_BYTES, 1024);
                        umma_bf16(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (kb == kb1 - 1) umma_commit(&tmem_full[acc]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ======================================= epilogue =======================================
        const int quad = warp & 3;                    // TMEM lane quadrant this warp may read
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            const int split = item / inner_items;
            int rem = item - split * inner_items;
            const int co_tile = rem / (prm.taps * prm.ci_tiles);
            rem -= co_tile * prm.taps * prm.ci_tiles;
            const int tap = rem / prm.ci_tiles, ci_tile = rem - tap * prm.ci_tiles;
            mbar_wait(&tmem_full[acc], acc_phase);
            tcgen05_fence_after();
            const int co = co_tile * WG_M + quad * 32 + lane;
            float* row = prm.dw + ((size_t)co * prm.taps + tap) * prm.Cin + ci_tile * BLOCK_N;
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N + c * 32), v);
                tmem_ld_wait();
                if (co < prm.Cout) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        atomicAdd(reinterpret_cast<float4*>(row + c * 32 + 4 * j),
                                  make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                              __uint_as_float(v[4 * j + 3])));
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

template <int BLOCK_N, int STAGES>
int launch_wgrad(const CUtensorMap& tdy, const CUtensorMap& tx, const WgradParams& prm, cudaStream_t stream) {
    using L = WgSmem<BLOCK_N, STAGES>;
    static_assert(L::kTotal <= 232448, "shared memory budget of one CTA exceeded");
    auto kern = conv_wgrad_kernel<BLOCK_N, STAGES>;
    static bool configured = false;
    if (!configured) {
        int st = cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal),
                             "cudaFuncSetAttribute(conv_wgrad_kernel)");
        if (st) return st;
        configured = true;
    }
    const int items = prm.co_tiles * prm.taps * prm.ci_tiles * prm.splits;
    const int sms = tmah::driver().sm_count;
    const int grid = items < sms ? items : sms;
    kern<<<grid, WG_THREADS, L::kTotal, stream>>>(tdy, tx, prm);
    DPFT_LAUNCH_CHECK("conv_wgrad_kernel");
    return DPFT_OK;
}

}  // namespace
}  // namespace dpft

using namespace dpft;

extern "C" int dpft_conv2d_wgrad(const void* x, const void* dy, float* dw, int B, int H, int W, int Cin, int Cout, int R, int S,
                                 int stride, int pad, int splits, int dtype, void* stream) {
    DPFT_REQUIRE(x && dy && dw, "conv2d_wgrad: null pointer");
    DPFT_REQUIRE(dtype == DPFT_BF16 || dtype == DPFT_F16, "conv2d_wgrad: dtype must be DPFT_BF16 or DPFT_F16");
    const bool is_f16 = dtype == DPFT_F16;
    DPFT_REQUIRE(B > 0 && H > 0 && W > 0, "conv2d_wgrad: bad input size %dx%dx%d", B, H, W);
    DPFT_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "conv2d_wgrad: Cin=%d and Cout=%d must be multiples of 64", Cin, Cout);
    DPFT_REQUIRE(R >= 1 && S >= 1 && R <= 16 && S <= 16 && stride >= 1 && stride <= 8 && pad >= 0 && pad < 16,
                 "conv2d_wgrad: unsupported filter %dx%d stride %d pad %d", R, S, stride, pad);
    DPFT_REQUIRE((((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dw) & 15) == 0, "conv2d_wgrad: pointers must be 16-byte aligned");
    int st = tmah::resolve_driver();
    if (st) return st;
    const int P = (H + 2 * pad - R) / stride + 1, Q = (W + 2 * pad - S) / stride + 1;
    DPFT_REQUIRE(P > 0 && Q > 0, "conv2d_wgrad: empty output");
    WgradParams prm{};
    prm.Mpix = B * P * Q; prm.P = P; prm.Q = Q; prm.Cout = Cout; prm.Cin = Cin; prm.taps = R * S; prm.taps_s = S;
    prm.stride = stride; prm.pad = pad; prm.is_f16 = is_f16; prm.dw = dw;
    prm.im2col = !(R == 1 && S == 1 && stride == 1 && pad == 0);
    const int bn = Cin % 256 == 0 ? 256 : (Cin % 128 == 0 ? 128 : 64);
    prm.co_tiles = (Cout + WG_M - 1) / WG_M;
    prm.ci_tiles = Cin / bn;
    prm.kblocks_total = (prm.Mpix + WG_KPIX - 1) / WG_KPIX;
    const int base = prm.co_tiles * prm.taps * prm.ci_tiles;
    int want = splits > 0 ? splits : (2 * tmah::driver().sm_count + base - 1) / base;     // about two items per SM
    if (want < 1) want = 1;
    int per = (prm.kblocks_total + want - 1) / want;
    if (splits <= 0 && per < 8) per = 8;                      // amortise the atomic epilogue over a few k-blocks
    if (per > prm.kblocks_total) per = prm.kblocks_total;
    prm.kb_per_split = per;
    prm.splits = (prm.kblocks_total + per - 1) / per;

    CUtensorMap tdy, tx;
    st = tmah::encode_2d(&tdy, dy, (uint64_t)Cout, (uint64_t)prm.Mpix, (uint64_t)Cout * 2, 64, WG_KPIX, is_f16);
    if (st) return st;
    if (prm.im2col) st = tmah::encode_im2col(&tx, x, B, H, W, Cin, R, S, stride, pad, 64, WG_KPIX, is_f16);
    else st = tmah::encode_2d(&tx, x, (uint64_t)Cin, (uint64_t)prm.Mpix, (uint64_t)Cin * 2, 64, WG_KPIX, is_f16);
    if (st) return st;
    cudaStream_t s = (cudaStream_t)stream;
    if (bn == 256) return launch_wgrad<256, 4>(tdy, tx, prm, s);
    if (bn == 128) return launch_wgrad<128, 6>(tdy, tx, prm, s);
    return launch_wgrad<64, 8>(tdy, tx, prm, s);
}
