// 3x3 / stride 1 / pad 1 convolution with 64 input and 64 output channels (conv2 of the stage-1 Bottleneck blocks, reference
// src/dprt/models/backbones/resnet.py:54-55,101) as a tcgen05 implicit GEMM over a shared-memory HALO tile.
//
// The generic kernel (conv_tcgen05.cu) fetches one im2col tile per filter tap, i.e. it moves every input pixel nine times
// from L2 to shared memory; at 64 channels that traffic (518 MB per launch at 8 x 180 x 320) is what bounds the layer
// (7 TB/s of L2->SM traffic, tensor pipe 25 % active).  Here a CTA stages the input rows of a tile ONCE:
//
//   unit      one CTA = (image, 128-column strip, chunk of ~H / chunks output rows), about one unit per SM
//   A (smem)  a ring of 8 input rows x 136 input columns x 64 channels = 128-byte pixel rows; every row of the chunk is
//             loaded exactly once by ONE TMA op (4-d tiled map, box = 64 channels x 136 columns x 1 row, 128B swizzle,
//             zero fill outside the image = the padding), five rows ahead of the multiply
//   tap       (dr, ds) is nothing but a start address: ring row (j + dr) + ds * 128 B into the K-major 128B-swizzled
//             row (the swizzle is a function of the address bits, so the shifted start needs no descriptor base offset)
//   B (smem)  the whole [64 x 576] weight matrix, resident: nine [64 filters x 64 channels] 128B-swizzled tiles
//   MMA       per output row 9 taps x 4 k-steps of tcgen05.mma kind::f16 M=128 (columns) N=64 K=16 into one of four
//             64-column TMEM accumulators; the commit also releases the oldest input row of the window
//   epilogue  eight warps (TMEM lane quadrant x even / odd output row): + bias, ReLU, 16-bit pack, 128 contiguous bytes
//             per pixel
//
// warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-9 = epilogue.  L2->SM traffic drops from 9x to ~1.1x of the input.
// Measured (8 x 180 x 320, f16): 72.7 us (im2col kernel) -> 51.2 us; ncu: tensor pipe 47 % active, bound by the shared-
// memory operand reads of N = 64 MMAs (4 KB of A + 2 KB of B per 32-cycle MMA), no longer by L2 traffic.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tcgen05.cuh"
#include "tma_host.cuh"

namespace dpft {
namespace {
using namespace tc;

constexpr int HC = 64;                               // channels in and out
constexpr int H_TW = 128;                            // output columns per strip
constexpr int H_PW = 136;                            // input columns per strip (>= H_TW + 2, multiple of 8)
constexpr int H_ROW_BYTES = H_PW * HC * 2;           // one input row of a strip: 136 pixels x 128 B = 17 swizzle atoms = one TMA box
constexpr int H_RING = 8;                            // input rows in flight
constexpr int H_ACCS = 4;                            // output-row accumulators in TMEM (64 columns each)
constexpr int H_B_TILE = HC * HC * 2;                // one tap: 64 filters x 128 B
constexpr int H_B_BYTES = 9 * H_B_TILE;              // 73,728
static_assert(H_ROW_BYTES % 1024 == 0, "input rows must start on a swizzle-atom boundary");
constexpr int H_THREADS = 320;
constexpr int H_SMEM = H_B_BYTES + H_RING * H_ROW_BYTES + 64 * 4 + (2 * H_RING + 1 + 2 * H_ACCS) * 8 + 16;

struct HaloParams {
    int B, H, W;
    int strips;                 // 128-column strips per image row
    int chunks, chunk_rows;     // row chunks per strip, output rows per chunk
    int relu, is_f16;
    const float* bias;
    void* y;
};

// One CTA = one (image, column strip, row chunk): it streams the chunk's input rows through a ring (each row is loaded
// exactly once, + the two halo rows of the chunk) and produces one output row per 36 MMAs.
__global__ void __launch_bounds__(H_THREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const HaloParams prm) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_b = smem;
    uint8_t* s_a = smem + H_B_BYTES;
    float* s_bias = reinterpret_cast<float*>(s_a + H_RING * H_ROW_BYTES);
    uint64_t* row_full = reinterpret_cast<uint64_t*>(s_bias + 64);
    uint64_t* row_empty = row_full + H_RING;
    uint64_t* w_bar = row_empty + H_RING;
    uint64_t* tmem_full = w_bar + 1;
    uint64_t* tmem_empty = tmem_full + H_ACCS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + H_ACCS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = blockIdx.x;
    const int chunk = unit % prm.chunks;
    const int strip = (unit / prm.chunks) % prm.strips;
    const int b = unit / (prm.chunks * prm.strips);
    const int p_begin = chunk * prm.chunk_rows;
    const int n_out = min(prm.chunk_rows, prm.H - p_begin);          // output rows of this CTA (>= 1)
    const int n_in = n_out + 2;                                       // input rows p_begin - 1 .. p_begin + n_out

    if (warp == 0 && elect_one()) {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_w);
        for (int i = 0; i < H_RING; ++i) {
            mbar_init(&row_full[i], 1);
            mbar_init(&row_empty[i], 1);
        }
        for (int i = 0; i < H_ACCS; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4);             // the four warps (lane quadrants) that drain this accumulator
        }
        mbar_init(w_bar, 1);
        fence_barrier_init();
    } else if (warp == 1) {
        tmem_alloc(tmem_slot, H_ACCS * HC);
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
            for (int i = 0; i < n_in; ++i) {
                const int slot = i % H_RING;
                mbar_wait(&row_empty[slot], ((i / H_RING) & 1) ^ 1);
                mbar_expect_tx(&row_full[slot], H_ROW_BYTES);
                // input row p_begin - 1 + i, columns strip*128 - 1 .. +135; rows / columns outside the image arrive as zeros
                tma_load_4d(&tmap_x, &row_full[slot], s_a + slot * H_ROW_BYTES, 0, strip * H_TW - 1, p_begin - 1 + i, b);
            }
        }
    } else if (warp == 1) {
        // ====================================== MMA issuer ======================================
        const uint32_t idesc = umma_idesc_16bit(128, HC, prm.is_f16 != 0);
        mbar_wait(w_bar, 0);
        const uint32_t b0 = smem_u32(s_b), a0 = smem_u32(s_a);
        // input rows 0 and 1 of the chunk (the third row of each window is awaited inside the loop)
        mbar_wait(&row_full[0], 0);
        mbar_wait(&row_full[1], 0);
        for (int j = 0; j < n_out; ++j) {
            const int acc = j % H_ACCS;
            mbar_wait(&tmem_empty[acc], ((j / H_ACCS) & 1) ^ 1);
            mbar_wait(&row_full[(j + 2) % H_RING], ((j + 2) / H_RING) & 1);
            tcgen05_fence_after();
            if (elect_one()) {
                const uint32_t tmem_d = tmem_base + acc * HC;
#pragma unroll 1
                for (int dr = 0; dr < 3; ++dr) {
                    const uint32_t row = a0 + ((j + dr) % H_RING) * H_ROW_BYTES;
#pragma unroll
                    for (int ds = 0; ds < 3; ++ds) {
                        // 128 pixel rows starting ds pixels into input row j + dr.  The 128B swizzle is a function of the
                        // shared-memory ADDRESS bits (TMA wrote the row that way), so a start that is 128-byte but not
                        // 1024-byte aligned needs no base offset in the descriptor (measured: setting it breaks the result)
                        const uint64_t adesc = umma_desc_sw128(row + ds * 128);
                        const uint64_t bdesc = umma_desc_sw128(b0 + (dr * 3 + ds) * H_B_TILE);
#pragma unroll
                        for (int kk = 0; kk < HC / 16; ++kk)           // +32 bytes per K = 16 step inside the 128-byte row
                            umma_bf16(tmem_d, adesc + 2 * kk, bdesc + 2 * kk, idesc, (dr | ds | kk) ? 1u : 0u);
                    }
                }
                umma_commit(&tmem_full[acc]);
                umma_commit(&row_empty[j % H_RING]);           // input row j is not needed by any later output row
            }
            __syncwarp();
        }
    } else {
        // ======================================= epilogue =======================================
        // warps 2-5 drain the even output rows, warps 6-9 the odd ones; a warp owns TMEM lane quadrant warp % 4 = 32 columns
        const int quad = warp & 3;
        const int grp = (warp - 2) >> 2;
        const int q = strip * H_TW + quad * 32 + lane;
        for (int j = grp; j < n_out; j += 2) {
            const int acc = j % H_ACCS;
            const int p = p_begin + j;
            mbar_wait(&tmem_full[acc], (j / H_ACCS) & 1);
            tcgen05_fence_after();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * HC + half * 32), v);
                tmem_ld_wait();
                if (q < prm.W) {
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
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, H_ACCS * HC);
    }
}

}  // namespace

bool conv3x3_halo_eligible(int H, int W, int Cin, int Cout, int R, int S, int stride, int pad, bool has_residual) {
    static const int enabled = [] { const char* e = getenv("DPFT_CONV_HALO"); return (e && e[0] == '0') ? 0 : 1; }();
    // wide maps only: a 128-column tile on a narrower map multiplies zeros
    return enabled && R == 3 && S == 3 && stride == 1 && pad == 1 && Cin == HC && Cout == HC && !has_residual && W >= 96;
}

int conv3x3_halo_launch(const void* x, const void* w, const float* bias, void* y, int B, int H, int W, int relu, bool is_f16,
                        cudaStream_t stream) {
    int st = tmah::resolve_driver();
    if (st) return st;
    CUtensorMap tx, tw;
    {   // activation (B, H, W, 64) as a 4-d tiled map, 128B swizzle: box = 64 channels x 136 columns x 1 row x 1 image
        cuuint64_t dims[4] = {(cuuint64_t)HC, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)HC * 2, (cuuint64_t)W * HC * 2, (cuuint64_t)H * W * HC * 2};
        cuuint32_t box[4] = {(cuuint32_t)HC, (cuuint32_t)H_PW, 1, 1};
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
    // work units = (image, 128-column strip, row chunk): about one per SM, every chunk with the same number of rows
    const int sms = tmah::driver().sm_count;
    const int strips = (W + H_TW - 1) / H_TW;
    int chunks = (sms + B * strips / 2) / (B * strips);
    if (chunks < 1) chunks = 1;
    if (chunks > H) chunks = H;
    const int chunk_rows = (H + chunks - 1) / chunks;
    chunks = (H + chunk_rows - 1) / chunk_rows;
    HaloParams prm{B, H, W, strips, chunks, chunk_rows, relu, is_f16 ? 1 : 0, bias, y};
    const long long grid = (long long)B * strips * chunks;
    DPFT_REQUIRE(grid <= 0x7fffffffLL, "conv3x3 halo: too many work units");
    conv3x3_halo_kernel<<<(unsigned)grid, H_THREADS, H_SMEM, stream>>>(tx, tw, prm);
    DPFT_LAUNCH_CHECK("conv3x3_halo_kernel");
    return DPFT_OK;
}

}  // namespace dpft
