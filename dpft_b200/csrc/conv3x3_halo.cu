// 3x3 / stride 1 / pad 1 convolution with 64 input and 64 output channels (conv2 of the stage-1 Bottleneck blocks, reference
// src/dprt/models/backbones/resnet.py:54-55,101) as a tcgen05 implicit GEMM over a shared-memory HALO tile.
//
// The generic kernel (conv_tcgen05.cu) fetches one im2col tile per filter tap, i.e. it moves every input pixel nine times
// from L2 to shared memory; at 64 channels that traffic (518 MB per launch at 8 x 180 x 320) is what bounds the layer
// (7 TB/s of L2->SM traffic, tensor pipe 25 % active).  Here a CTA stages the input rows of a tile ONCE:
//
//   tile      2 output rows x 128 output columns of one image
//   A (smem)  4 input rows x 136 input columns x 64 channels = 128-byte pixel rows, ONE TMA load (4-d tiled map, box =
//             64 channels x 136 columns x 4 rows, 128B swizzle, zero fill outside the image = the padding)
//   tap       (dr, ds) is nothing but a start address: + dr * (136 * 128 B) + ds * 128 B into the K-major 128B-swizzled
//             tile (the swizzle is a function of the address bits, so the shifted start needs no descriptor base offset)
//   B (smem)  the whole [64 x 576] weight matrix, resident for the life of the (persistent) CTA: nine [64 filters x 64
//             channels] 128B-swizzled tiles, one per tap
//   MMA       per output row 9 taps x 4 k-steps of tcgen05.mma kind::f16 M=128 (columns) N=64 K=16 into 64 TMEM columns;
//             two rows per tile, accumulators double buffered (256 columns)
//   epilogue  eight warps (TMEM lane quadrant x output row): + bias, ReLU, 16-bit pack, 128 contiguous bytes per pixel
//
// warp 0 = TMA producer (A double buffered: the next tile lands while this one is multiplied), warp 1 = MMA issuer,
// warps 2-9 = epilogue.  L2->SM traffic drops from 9x to 2.1x of the input.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tcgen05.cuh"
#include "tma_host.cuh"

namespace dpft {
namespace {
using namespace tc;

constexpr int HC = 64;                               // channels in and out
constexpr int H_TH = 2, H_TW = 128;                  // output tile
constexpr int H_ROWS = H_TH + 2;                     // input rows per tile
constexpr int H_PW = 136;                            // input columns per tile (>= H_TW + 2, multiple of 8)
constexpr int H_ROW_BYTES = H_PW * HC * 2;           // one input row of the tile: 136 pixels x 128 B = 17 swizzle atoms
constexpr int H_A_BYTES = H_ROWS * H_ROW_BYTES;      // 69,632 = one TMA box
constexpr int H_B_TILE = HC * HC * 2;                // one tap: 64 filters x 128 B
constexpr int H_B_BYTES = 9 * H_B_TILE;              // 73,728
static_assert(H_ROW_BYTES % 1024 == 0, "image rows of the halo tile must start on a swizzle-atom boundary");
constexpr int H_THREADS = 320;
constexpr int H_SMEM = H_B_BYTES + 2 * H_A_BYTES + 64 * 4 + 16 * 8 + 16;

struct HaloParams {
    int B, H, W;
    int tiles_w, tiles_h;       // tiles per image row / column
    int relu, is_f16;
    const float* bias;
    void* y;
};

__global__ void __launch_bounds__(H_THREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const HaloParams prm) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_b = smem;
    uint8_t* s_a = smem + H_B_BYTES;
    float* s_bias = reinterpret_cast<float*>(s_a + 2 * H_A_BYTES);
    uint64_t* a_full = reinterpret_cast<uint64_t*>(s_bias + 64);
    uint64_t* a_empty = a_full + 2;
    uint64_t* w_bar = a_empty + 2;
    uint64_t* tmem_full = w_bar + 1;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_image = prm.tiles_w * prm.tiles_h;
    const int num_tiles = prm.B * tiles_per_image;

    if (warp == 0 && elect_one()) {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_w);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 8);
        }
        mbar_init(w_bar, 1);
        fence_barrier_init();
    } else if (warp == 1) {
        tmem_alloc(tmem_slot, 256);
    }
    if (threadIdx.x >= 64 && threadIdx.x < 128) s_bias[threadIdx.x - 64] = __ldg(prm.bias + threadIdx.x - 64);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (elect_one()) {
            mbar_expect_tx(w_bar, H_B_BYTES);
            for (int tap = 0; tap < 9; ++tap)                  // columns tap*64 .. +64 of the [64 x 576] weight matrix
                tma_load_2d(&tmap_w, w_bar, s_b + tap * H_B_TILE, tap * HC, 0);
            int buf = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int b = tile / tiles_per_image;
                const int rem = tile - b * tiles_per_image;
                const int th = rem / prm.tiles_w, tw = rem - th * prm.tiles_w;
                mbar_wait(&a_empty[buf], phase ^ 1);
                mbar_expect_tx(&a_full[buf], H_A_BYTES);
                tma_load_4d(&tmap_x, &a_full[buf], s_a + buf * H_A_BYTES, 0, tw * H_TW - 1, th * H_TH - 1, b);
                if (++buf == 2) { buf = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ====================================== MMA issuer ======================================
        const uint32_t idesc = umma_idesc_16bit(128, HC, prm.is_f16 != 0);
        int buf = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        mbar_wait(w_bar, 0);
        const uint32_t b0 = smem_u32(s_b);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            mbar_wait(&a_full[buf], phase);
            tcgen05_fence_after();
            if (elect_one()) {
                const uint32_t a0 = smem_u32(s_a + buf * H_A_BYTES);
#pragma unroll 1
                for (int r = 0; r < H_TH; ++r) {
                    const uint32_t tmem_d = tmem_base + acc * (H_TH * HC) + r * HC;
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap) {
                        const int dr = tap / 3, ds = tap - dr * 3;
                        // 128 pixel rows starting ds pixels into input row r + dr.  The 128B swizzle is a function of the
                        // shared-memory ADDRESS bits (TMA wrote the tile that way), so a start that is 128-byte but not
                        // 1024-byte aligned needs no base offset in the descriptor (measured: setting it breaks the result)
                        const uint64_t adesc = umma_desc_sw128(a0 + (r + dr) * H_ROW_BYTES + ds * 128);
                        const uint64_t bdesc = umma_desc_sw128(b0 + tap * H_B_TILE);
#pragma unroll
                        for (int kk = 0; kk < HC / 16; ++kk)           // +32 bytes per K = 16 step inside the 128-byte row
                            umma_bf16(tmem_d, adesc + 2 * kk, bdesc + 2 * kk, idesc, (tap | kk) ? 1u : 0u);
                    }
                }
                umma_commit(&a_empty[buf]);                    // the halo buffer is free once these MMAs retire
                umma_commit(&tmem_full[acc]);
            }
            __syncwarp();
            if (++buf == 2) { buf = 0; phase ^= 1; }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ======================================= epilogue =======================================
        const int quad = warp & 3;                        // TMEM lane quadrant = 32 output columns
        const int r = (warp - 2) >> 2;                    // output row of the tile
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int b = tile / tiles_per_image;
            const int rem = tile - b * tiles_per_image;
            const int th = rem / prm.tiles_w, tw = rem - th * prm.tiles_w;
            const int p = th * H_TH + r, q = tw * H_TW + quad * 32 + lane;
            mbar_wait(&tmem_full[acc], acc_phase);
            tcgen05_fence_after();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * (H_TH * HC) + r * HC + half * 32), v);
                tmem_ld_wait();
                if (p < prm.H && q < prm.W) {
                    uint16_t* o = reinterpret_cast<uint16_t*>(prm.y) + (((long long)b * prm.H + p) * prm.W + q) * HC + half * 32;
#pragma unroll
                    for (int o8 = 0; o8 < 4; ++o8) {
                        const float4 b0v = *reinterpret_cast<const float4*>(s_bias + half * 32 + o8 * 8);
                        const float4 b1v = *reinterpret_cast<const float4*>(s_bias + half * 32 + o8 * 8 + 4);
                        float f[8] = {__uint_as_float(v[8 * o8]) + b0v.x,     __uint_as_float(v[8 * o8 + 1]) + b0v.y,
                                      __uint_as_float(v[8 * o8 + 2]) + b0v.z, __uint_as_float(v[8 * o8 + 3]) + b0v.w,
                                      __uint_as_float(v[8 * o8 + 4]) + b1v.x, __uint_as_float(v[8 * o8 + 5]) + b1v.y,
                                      __uint_as_float(v[8 * o8 + 6]) + b1v.z, __uint_as_float(v[8 * o8 + 7]) + b1v.w};
                        if (prm.relu) {
#pragma unroll
                            for (int t = 0; t < 8; ++t) f[t] = fmaxf(f[t], 0.0f);
                        }
                        uint4 ov;
                        uint32_t* ow = reinterpret_cast<uint32_t*>(&ov);
                        if (prm.is_f16) {
#pragma unroll
                            for (int t = 0; t < 4; ++t)
                                asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(ow[t]) : "f"(f[2 * t + 1]), "f"(f[2 * t]));
                        } else {
#pragma unroll
                            for (int t = 0; t < 4; ++t)
                                asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(ow[t]) : "f"(f[2 * t + 1]), "f"(f[2 * t]));
                        }
                        reinterpret_cast<uint4*>(o)[o8] = ov;
                    }
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
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace

bool conv3x3_halo_eligible(int H, int W, int Cin, int Cout, int R, int S, int stride, int pad, bool has_residual) {
    static const int enabled = [] { const char* e = getenv("DPFT_CONV_HALO"); return (e && e[0] == '0') ? 0 : 1; }();
    // wide maps only: a 128-column tile on a narrower map multiplies zeros
    return enabled && R == 3 && S == 3 && stride == 1 && pad == 1 && Cin == HC && Cout == HC && !has_residual && W >= 96 && H >= 2;
}

int conv3x3_halo_launch(const void* x, const void* w, const float* bias, void* y, int B, int H, int W, int relu, bool is_f16,
                        cudaStream_t stream) {
    int st = tmah::resolve_driver();
    if (st) return st;
    CUtensorMap tx, tw;
    {   // activation (B, H, W, 64) as a 4-d tiled map, 128B swizzle: box = 64 channels x 136 columns x 4 rows x 1 image
        cuuint64_t dims[4] = {(cuuint64_t)HC, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)HC * 2, (cuuint64_t)W * HC * 2, (cuuint64_t)H * W * HC * 2};
        cuuint32_t box[4] = {(cuuint32_t)HC, (cuuint32_t)H_PW, (cuuint32_t)H_ROWS, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = tmah::driver().encode_tiled(&tx, tmah::dtype16(is_f16), 4, const_cast<void*>(x), dims, strides, box, estr,
                                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(halo activation) failed with CUresult %d", (int)r); return DPFT_ERR_INVALID_ARGUMENT; }
    }
    {   // weights [64][3][3][64] as a [64 x 576] matrix: box = 64 channels x 64 filters (one 128B-swizzled tile per tap)
        cuuint64_t dims[2] = {(cuuint64_t)9 * HC, (cuuint64_t)HC};
        cuuint64_t strides[1] = {(cuuint64_t)9 * HC * 2};
        cuuint32_t box[2] = {(cuuint32_t)HC, (cuuint32_t)HC};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = tmah::driver().encode_tiled(&tw, tmah::dtype16(is_f16), 2, const_cast<void*>(w), dims, strides, box, estr,
                                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(halo weights) failed with CUresult %d", (int)r); return DPFT_ERR_INVALID_ARGUMENT; }
    }
    static bool configured = false;
    if (!configured) {
        st = cuda_status(cudaFuncSetAttribute(conv3x3_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, H_SMEM),
                         "cudaFuncSetAttribute(conv3x3_halo_kernel)");
        if (st) return st;
        configured = true;
    }
    HaloParams prm{B, H, W, (W + H_TW - 1) / H_TW, (H + H_TH - 1) / H_TH, relu, is_f16 ? 1 : 0, bias, y};
    const long long tiles = (long long)B * prm.tiles_w * prm.tiles_h;
    const int sms = tmah::driver().sm_count;
    const int grid = (int)(tiles < sms ? tiles : sms);
    conv3x3_halo_kernel<<<grid, H_THREADS, H_SMEM, stream>>>(tx, tw, prm);
    DPFT_LAUNCH_CHECK("conv3x3_halo_kernel");
    return DPFT_OK;
}

}  // namespace dpft
