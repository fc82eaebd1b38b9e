// Backbone convolutions as tcgen05 implicit GEMMs for sm_100a.
//
// Replaces the cuDNN convolution + BatchNorm + ReLU (+ residual add) launches behind the ResNet bodies the
// reference builds at src/dprt/models/backbones/resnet.py:54-55,101 (torchvision Bottleneck blocks), in
// inference form: BatchNorm folded into the weights/bias on the host, activations NHWC bf16, fp32
// accumulation in TMEM.
//
//   D[m, n] = act( sum_k A[m, k] * Wt[n, k] + bias[n] (+ residual[m, n]) )
//   m = output pixel (b, p, q) linearised,  n = output channel,  k = (r, s, c) with c fastest.
//
// One persistent CTA per SM (or a CTA pair per 256-row tile, cta_group::2), 10 or 18 warps:
//   warp 0   TMA producer   A tile (128 pixels x 64 channels of one filter tap) by an im2col-mode tensor map
//                           (cp.async.bulk.tensor.4d...im2col; the tap (s, r) goes in the offset operands) or,
//                           for 1x1/stride-1 layers, a plain 2-d tiled map over the [M, Cin] activation matrix;
//                           B tile (BLOCK_N filters x 64) from the [Cout, R*S*Cin] weight matrix.  128B swizzle.
//   warp 1   MMA issuer     one elected lane issues tcgen05.mma.kind::f16 (M=128 or 256, N=BLOCK_N, K=16) x4 per stage;
//                           tcgen05.commit releases the stage / publishes the accumulator.
//   warps 2+ epilogue       EG (2 or 4) warps per TMEM lane quadrant: tcgen05.ld the fp32 accumulator, + bias (smem copy of
//                           the tile's bias), + residual (TMA-prefetched [32 x 64] sub-tiles), ReLU, cvt.satfinite pack into
//                           a 128B-swizzled staging buffer, TMA store.
// The accumulator is double-buffered in TMEM (2 x BLOCK_N columns) so the epilogue of tile i overlaps the
// main loop of tile i+1.  Stages: a ring of STAGES {A 16 KB, B BLOCK_N*128 B} buffers with full/empty mbarriers.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace dpft {
namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;          // bf16 elements = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
// warp 0 TMA, warp 1 MMA, then 4*EG epilogue warps: EG per TMEM lane quadrant (= per scheduler), each owning 64/EG of an item's columns
constexpr int threads_for(int eg) { return (2 + 4 * eg) * 32; }

enum ConvMode : int { kTiled2D = 0, kIm2col = 1 };

struct ConvParams {
    int M, N;               // output pixels, output channels
    int P, Q;               // output height, width
    int taps_s;             // filter width S (k-block -> (r, s))
    int cblocks;            // Cin / 64
    int kblocks;            // R * S * cblocks
    int stride, pad;
    int relu;
    int mode;
    int is_f16;            // activations / weights are IEEE half instead of bfloat16
    int m_tiles, n_tiles;
    const float* bias;      // [N]
    const __nv_bfloat16* residual;  // [M, N] or null
    __nv_bfloat16* out;     // [M, N]
    // FPN lateral form: fp32 output of the first 16 channels (+ nearest-upsampled coarser level)
    float* out_f32;         // [M, 16] or null
    const float* coarse;    // (B, Hc, Wc, 16) or null
    int Hc, Wc;
};

__device__ __forceinline__ int nearest_src(int dst, int in_size, int out_size) {
    // torch 'nearest': src = min(floor(dst * (float(in) / out)), in - 1)
    const float scale = (float)in_size / (float)out_size;
    const int s = (int)floorf((float)dst * scale);
    return s < in_size - 1 ? s : in_size - 1;
}

// kind::f16 instruction descriptor: D = f32, A = B = bf16 (format 1) or f16 (format 0), both K-major, M = 128, N = n.
__host__ __device__ constexpr uint32_t umma_idesc_16bit(int n, bool is_f16) {
    const uint32_t fmt = is_f16 ? 0u : 1u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}
using namespace tc;

constexpr int EPI_COLS = 64;                       // columns per epilogue item: one 128-byte swizzle row of 16-bit outputs
constexpr int EPI_BUF_BYTES = 32 * EPI_COLS * 2;   // one warp's [32 rows x 64 cols] sub-tile
constexpr int EPI_OUT_BUFS = 2;

// EPI_RES_BUFS = residual sub-tiles per warp (EPI_RES_BUFS - 1 prefetched + one in use).  Layers with a short K loop
// and a residual (the 1x1 "expand" convs) are bound by the residual/output streams, not by the operand pipeline, so
// they trade operand stages for a deeper residual prefetch.
template <int BLOCK_N, int STAGES, int EPI_RES_BUFS, int CG = 1> struct SmemLayout {
    static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
    static constexpr int kBBytes = (BLOCK_N / CG) * BLOCK_K * 2;          // a CTA pair holds half of the B tile each
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kEpiBytes = 4 * (EPI_RES_BUFS + EPI_OUT_BUFS) * EPI_BUF_BYTES;
    static constexpr int kBiasBytes = 4 * BLOCK_N * 4;                                      // tile bias, one copy per warp pair
    static constexpr int kBarrierBytes = (2 * STAGES + 4 + 4 * EPI_RES_BUFS) * 8 + 16;
    static constexpr int kTotal = STAGES * kStageBytes + kEpiBytes + kBiasBytes + kBarrierBytes;
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)),
                 "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}


// ---- 2-CTA (cta_group::2) helpers: the CTA pair of a cluster cooperates on a 256-row tile; CTA 0 issues the MMAs ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(const void* p, uint32_t rank) {
    uint32_t a;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(smem_u32(p)), "r"(rank));
    return a;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: the data lands in the issuing CTA's shared memory, the bytes are counted on `bar_cluster_addr`
// (the leader CTA's barrier).
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* map, uint32_t bar_cluster_addr, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
            "r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_2sm(const CUtensorMap* map, uint32_t bar_cluster_addr, void* dst, int c, int w,
                                                       int h, int n, uint16_t off_w, uint16_t off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2], {%7, %8};" ::"r"(smem_u32(dst)),
        "l"(map), "r"(bar_cluster_addr), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
        : "memory");
}
__device__ __forceinline__ void umma_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// commit of a CTA-pair MMA: arrives on the barrier at the same offset in both CTAs
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* slot_in_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t base, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}

// ---- the kernel -------------------------------------------------------------------------------------------------
// One launch serves one problem, or TWO problems of identical shape (gridDim.y = 2: the two radar views of the fusion model
// run the same ResNet-50 on equally sized inputs with different weights).  Their layers are latency-bound (~10 us for < 1 us of
// math), so a second problem in the same launch costs the first one nothing and halves the launches and the SM-time of the pair.
struct alignas(64) ConvLaunch {
    CUtensorMap ta[2], tb[2], td[2], tr[2];
    ConvParams prm[2];
};

template <int BLOCK_N, int STAGES, int EPI_RES_BUFS, int CG, int EG>
__global__ void __launch_bounds__(threads_for(EG), 1)
conv_gemm_kernel(const __grid_constant__ ConvLaunch launch_desc) {
    const CUtensorMap& tmap_a = launch_desc.ta[blockIdx.y];
    const CUtensorMap& tmap_b = launch_desc.tb[blockIdx.y];
    const CUtensorMap& tmap_d = launch_desc.td[blockIdx.y];
    const CUtensorMap& tmap_r = launch_desc.tr[blockIdx.y];
    const ConvParams& prm = launch_desc.prm[blockIdx.y];
    using L = SmemLayout<BLOCK_N, STAGES, EPI_RES_BUFS, CG>;
    extern __shared__ __align__(1024) uint8_t smem[];                    // 128-byte swizzle needs 1024-byte alignment
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * L::kABytes;
    uint8_t* smem_epi = smem + STAGES * L::kStageBytes;                  // per warp: residual + output sub-tiles
    float* smem_bias = reinterpret_cast<float*>(smem_epi + L::kEpiBytes);   // [4 pairs][BLOCK_N]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_epi + L::kEpiBytes + L::kBiasBytes);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* res_bar = tmem_empty + 2;                                  // [4 warps][EPI_RES_BUFS]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 4 * EPI_RES_BUFS);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr uint32_t kTmemCols = 2 * BLOCK_N;       // 128, 256 or 512: a power of two >= 32
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;       // position in the CTA pair
    const int cta_id = blockIdx.x / CG, num_ctas = gridDim.x / CG;   // the pair walks the tile list together

    if (warp == 0 && elect_one()) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        prefetch_tmap(&tmap_d);
        prefetch_tmap(&tmap_r);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4 * EG * CG);   // one arrive per epilogue warp (of both CTAs of a pair)
        }
        for (int i = 0; i < 4 * EPI_RES_BUFS; ++i) mbar_init(&res_bar[i], 1);
        fence_barrier_init();
    } else if (warp == 1) {
        if (CG == 2) tmem_alloc_2sm(tmem_slot, kTmemCols);
        else tmem_alloc(tmem_slot, kTmemCols);
    }
    tcgen05_fence_before();
    if (CG == 2) cluster_sync_all();                  // the peer's barriers must be initialised before anything signals them
    else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, tensor-map prefetch) touches no global
    // data, so it may overlap the tail of the previous kernel in the stream; the next kernel's CTAs may start their own set-up
    // as soon as SMs free up.  No global read or write happens before this wait.  (No-ops without the launch attribute.)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // CG == 2: a tile is 256 rows (this CTA owns rows rank*128 .. +128 of it) and the pair shares the B tile
    const int big_m_tiles = (prm.m_tiles + CG - 1) / CG;
    const int num_tiles = big_m_tiles * prm.n_tiles;

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t full_leader = CG == 2 ? map_to_cta(full_bar, 0) : 0;     // the pair's loads count on CTA 0's barrier
            for (int tile = cta_id; tile < num_tiles; tile += num_ctas) {
                const int m_tile = (tile / prm.n_tiles) * CG + (int)rank, n_tile = tile % prm.n_tiles;
                int cw = 0, ch = 0, cn = 0;           // first pixel of the tile in im2col tensor-map coordinates
                if (prm.mode == kIm2col) {
                    const int m0 = m_tile * BLOCK_M;
                    const int pq = prm.P * prm.Q;
                    cn = m0 / pq;
                    const int rem = m0 - cn * pq;
                    const int p = rem / prm.Q, q = rem - p * prm.Q;
                    cw = q * prm.stride - prm.pad;
                    ch = p * prm.stride - prm.pad;
                }
                for (int kb = 0; kb < prm.kblocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (rank == 0) mbar_expect_tx(&full_bar[stage], CG * L::kStageBytes);
                    void* a_dst = smem_a + stage * L::kABytes;
                    void* b_dst = smem_b + stage * L::kBBytes;
                    const int tap = kb / prm.cblocks;
                    const int c0 = (kb - tap * prm.cblocks) * BLOCK_K;
                    if (CG == 2) {
                        const uint32_t bar = full_leader + stage * 8;
                        if (prm.mode == kTiled2D) {
                            tma_load_2d_2sm(&tmap_a, bar, a_dst, c0, m_tile * BLOCK_M);
                        } else {
                            const int r = tap / prm.taps_s, s = tap - r * prm.taps_s;
                            tma_load_im2col_4d_2sm(&tmap_a, bar, a_dst, c0, cw, ch, cn, (uint16_t)s, (uint16_t)r);
                        }
                        tma_load_2d_2sm(&tmap_b, bar, b_dst, kb * BLOCK_K, n_tile * BLOCK_N + (int)rank * (BLOCK_N / 2));
                    } else {
                        if (prm.mode == kTiled2D) {
                            tma_load_2d(&tmap_a, &full_bar[stage], a_dst, c0, m_tile * BLOCK_M);
                        } else {
                            const int r = tap / prm.taps_s, s = tap - r * prm.taps_s;
                            tma_load_im2col_4d(&tmap_a, &full_bar[stage], a_dst, c0, cw, ch, cn, (uint16_t)s, (uint16_t)r);
                        }
                        tma_load_2d(&tmap_b, &full_bar[stage], b_dst, kb * BLOCK_K, n_tile * BLOCK_N);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ====================================== MMA issuer (leader CTA of a pair) ======================================
        const uint32_t idesc = tc::umma_idesc_16bit(BLOCK_M * CG, BLOCK_N, prm.is_f16 != 0);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = cta_id; tile < num_tiles; tile += num_ctas) {
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
            for (int kb = 0; kb < prm.kblocks; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tcgen05_fence_after();
                if (elect_one()) {
                    const uint64_t adesc = umma_desc_sw128(smem_u32(smem_a + stage * L::kABytes));
                    const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + stage * L::kBBytes));
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // +32 bytes per K=16 step inside the 128-byte swizzle row: start-address field += 2
                        if (CG == 2) umma_2sm(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
                        else umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    }
                    if (CG == 2) {
                        umma_commit_2sm(&empty_bar[stage]);          // frees the stage in both CTAs
                        if (kb == prm.kblocks - 1) umma_commit_2sm(&tmem_full[acc]);
                    } else {
                        umma_commit(&empty_bar[stage]);              // frees the smem stage when the MMAs retire
                        if (kb == prm.kblocks - 1) umma_commit(&tmem_full[acc]);
                    }
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 2) {
        // ======================================= epilogue =======================================
        // Eight warps: warps w and w+4 share TMEM lane quadrant w%4 (= 32 rows of the tile) and the same scheduler, and
        // walk it in [32 rows x 64 cols] "items"; within an item warp group A (warps 2-5) handles columns 0-31 and group
        // B (warps 6-9) columns 32-63, so two warps per scheduler hide each other's dependent-issue latency.  Output items
        // are staged in shared memory (128-byte swizzle) and written by TMA; residual items are prefetched by TMA several
        // items ahead.  The pair synchronises with a 64-thread named barrier around the staging buffer.
        const int quad = warp & 3;
        const int grp = (warp - 2) >> 2;              // 0 .. EG-1: which 64/EG columns of an item; 0 issues the stores, 1 the residual loads
        constexpr int WCOLS = EPI_COLS / EG;          // columns per warp and item (32 or 16)
        constexpr int WCHUNKS = WCOLS / 8;            // 16-byte chunks per warp and row
        uint8_t* my_smem = smem_epi + quad * (EPI_RES_BUFS + EPI_OUT_BUFS) * EPI_BUF_BYTES;
        uint8_t* res_buf = my_smem;
        uint8_t* out_buf = my_smem + EPI_RES_BUFS * EPI_BUF_BYTES;
        uint64_t* my_res_bar = res_bar + quad * EPI_RES_BUFS;
        constexpr int kChunks = BLOCK_N / EPI_COLS;
        const bool has_res = prm.residual != nullptr;
        const int my_tiles = cta_id < num_tiles ? (num_tiles - 1 - cta_id) / num_ctas + 1 : 0;
        const uint32_t empty_leader = CG == 2 ? map_to_cta(tmem_empty, 0) : 0;
        const int total_items = my_tiles * kChunks;
        const bool issuer = grp == 0 && lane == 0;          // issues the TMA stores of the pair
        const bool loader = grp == 1 && lane == 0;          // issues the residual TMA loads of the pair
        auto pair_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(2 + quad), "n"(32 * EG) : "memory"); };

        auto prefetch_residual = [&](int item) {
            if (!has_res || item >= total_items) return;
            const int tile = cta_id + (item / kChunks) * num_ctas;
            const int m_tile = (tile / prm.n_tiles) * CG + (int)rank, n_tile = tile % prm.n_tiles;
            const int b = item % EPI_RES_BUFS;
            if (loader) {
                mbar_expect_tx(&my_res_bar[b], EPI_BUF_BYTES);
                tma_load_2d(&tmap_r, &my_res_bar[b], res_buf + b * EPI_BUF_BYTES, n_tile * BLOCK_N + (item % kChunks) * EPI_COLS,
                            m_tile * BLOCK_M + quad * 32);
            }
        };

        int acc = 0;
        uint32_t acc_phase = 0;
        int item = 0;
#pragma unroll
        for (int i = 0; i < EPI_RES_BUFS - 1; ++i) prefetch_residual(i);
        const int sw = lane & 7;                      // swizzle phase of this thread's row
        for (int tile = cta_id; tile < num_tiles; tile += num_ctas) {
            const int m_tile = (tile / prm.n_tiles) * CG + (int)rank, n_tile = tile % prm.n_tiles;
            mbar_wait(&tmem_full[acc], acc_phase);
            tcgen05_fence_after();
            const int n0 = n_tile * BLOCK_N;
            // this tile's bias -> the pair's shared-memory copy (the previous tile's readers are past their last pair_sync)
            float* bias_s = smem_bias + quad * BLOCK_N;
            for (int i = (grp * 32 + lane) * 4; i < BLOCK_N; i += 32 * EG * 4)
                *reinterpret_cast<float4*>(bias_s + i) = __ldg(reinterpret_cast<const float4*>(prm.bias + n0 + i));
            pair_sync();
            if (prm.out_f32) {
                // FPN lateral: inner = conv1x1 + bias (+ top-down), 16 fp32 channels per pixel, written directly (group A)
                const long long m = (long long)m_tile * BLOCK_M + quad * 32 + lane;
                if (grp == 0) {
                    uint32_t v[16];
                    tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_N), v);
                    tmem_ld_wait();
                    if (m < prm.M) {
                        const float4* cp = nullptr;
                        if (prm.coarse) {
                            const int pq = prm.P * prm.Q;
                            const int bi = (int)(m / pq);
                            const int rem = (int)(m - (long long)bi * pq);
                            const int p = rem / prm.Q, q = rem - p * prm.Q;
                            const int hc = nearest_src(p, prm.Hc, prm.P), wc = nearest_src(q, prm.Wc, prm.Q);
                            cp = reinterpret_cast<const float4*>(prm.coarse + (((long long)bi * prm.Hc + hc) * prm.Wc + wc) * 16);
                        }
                        float4* o = reinterpret_cast<float4*>(prm.out_f32 + m * 16);
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            const float4 bv = *reinterpret_cast<const float4*>(bias_s + 4 * j4);
                            float4 r = make_float4(__uint_as_float(v[4 * j4]) + bv.x, __uint_as_float(v[4 * j4 + 1]) + bv.y,
                                                   __uint_as_float(v[4 * j4 + 2]) + bv.z, __uint_as_float(v[4 * j4 + 3]) + bv.w);
                            if (cp) {
                                const float4 c = __ldg(cp + j4);
                                r.x += c.x; r.y += c.y; r.z += c.z; r.w += c.w;
                            }
                            o[j4] = r;
                        }
                    }
                }
            } else {
#pragma unroll 1
                for (int c = 0; c < kChunks; ++c, ++item) {
                    uint32_t v[WCOLS];
                    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) +
                                           (uint32_t)(acc * BLOCK_N + c * EPI_COLS + grp * WCOLS);
                    if constexpr (EG == 2) tmem_ld_32x32b_x32(taddr, v);
                    else tmem_ld_32x32b_x16(taddr, v);
                    tmem_ld_wait();
                    const float4* bias4 = reinterpret_cast<const float4*>(bias_s + c * EPI_COLS + grp * WCOLS);
                    const uint8_t* rrow = res_buf + (item % EPI_RES_BUFS) * EPI_BUF_BYTES + lane * 128;
                    if (has_res) mbar_wait(&my_res_bar[item % EPI_RES_BUFS], (uint32_t)((item / EPI_RES_BUFS) & 1));
                    // the staging buffer of item-2 must have been read by its TMA store before it is overwritten
                    if (issuer) bulk_wait_read<EPI_OUT_BUFS - 1>();
                    pair_sync();
                    uint8_t* orow = out_buf + (item % EPI_OUT_BUFS) * EPI_BUF_BYTES + lane * 128;
#pragma unroll
                    for (int jj = 0; jj < WCHUNKS; ++jj) {         // 8 columns = one 16-byte chunk of the row
                        const int j = grp * WCHUNKS + jj;           // chunk index within the 128-byte row
                        float f[8];
                        const float4 b0 = bias4[2 * jj], b1 = bias4[2 * jj + 1];
                        f[0] = __uint_as_float(v[8 * jj]) + b0.x;     f[1] = __uint_as_float(v[8 * jj + 1]) + b0.y;
                        f[2] = __uint_as_float(v[8 * jj + 2]) + b0.z; f[3] = __uint_as_float(v[8 * jj + 3]) + b0.w;
                        f[4] = __uint_as_float(v[8 * jj + 4]) + b1.x; f[5] = __uint_as_float(v[8 * jj + 5]) + b1.y;
                        f[6] = __uint_as_float(v[8 * jj + 6]) + b1.z; f[7] = __uint_as_float(v[8 * jj + 7]) + b1.w;
                        const int phys = (j ^ sw) << 4;             // 128-byte swizzle: chunk index xor (row % 8)
                        if (has_res) {
                            const uint4 rv = *reinterpret_cast<const uint4*>(rrow + phys);
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const float2 rf = prm.is_f16 ? __half22float2(reinterpret_cast<const __half2*>(&rv)[t])
                                                             : __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(&rv)[t]);
                                f[2 * t] += rf.x;
                                f[2 * t + 1] += rf.y;
                            }
                        }
                        if (prm.relu) {
#pragma unroll
                            for (int t = 0; t < 8; ++t) f[t] = fmaxf(f[t], 0.0f);
                        }
                        uint4 ov;                   // cvt.rn.satfinite packs two floats and clamps to the finite range
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
                        *reinterpret_cast<uint4*>(orow + phys) = ov;
                    }
                    fence_proxy_async();              // make the generic-proxy smem writes visible to the TMA engine
                    pair_sync();                      // both halves staged, this item's residual rows consumed
                    if (issuer) {
                        tma_store_2d(&tmap_d, out_buf + (item % EPI_OUT_BUFS) * EPI_BUF_BYTES, n0 + c * EPI_COLS,
                                     m_tile * BLOCK_M + quad * 32);
                        bulk_commit();
                    }
                    prefetch_residual(item + EPI_RES_BUFS - 1);   // refills the buffer consumed by item-1
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 2) mbar_arrive_cluster(empty_leader + acc * 8);   // the leader's MMA warp waits for both CTAs
                else mbar_arrive(&tmem_empty[acc]);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (issuer) bulk_wait_read<0>();              // smem must stay valid until the last stores have read it
    }

    tcgen05_fence_before();
    if (CG == 2) cluster_sync_all();                  // neither CTA may leave while the other can still signal it
    else __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        if (CG == 2) tmem_dealloc_2sm(tmem_base, kTmemCols);
        else tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---- weight-stationary variant for the 1x1 "expand" convolutions with a residual (conv3 of every Bottleneck) -------------
// These layers have a short K (Cin = 64 .. 256) and a wide N (4 x Cin): per 128 x 256 output tile the generic kernel moves
// A 128 x K, B 256 x K, the residual and the output, i.e. for K = 256 it re-reads 128 KB of weights for 64 KB of output, and
// on the 900-tile stage-3 layer the sum of operand + residual + output traffic (~290 MB in 34 us) runs at the chip's L2
// throughput, not at the tensor pipe (26 % active) or at HBM.  Here a CTA keeps its n-tile for its whole life (the grid is a
// multiple of n_tiles), loads that weight slice ONCE into shared memory and streams only A; the epilogue works IN PLACE on a
// ring of [32 x 64] buffers per TMEM lane quadrant: the residual sub-tile lands in a buffer by TMA, the warps add accumulator
// + bias, apply the ReLU and overwrite it with the packed outputs, the TMA store reads it, and the same thread refills it
// with the residual NB - 1 items ahead.  One 32*EG-thread barrier per item instead of two; the bias is staged once per CTA.
template <int BLOCK_N, int KB, int A_STAGES, int NB> struct WsLayout {
    static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
    static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
    static constexpr int kBResident = KB * kBBytes;
    static constexpr int kBufBytes = 4 * NB * EPI_BUF_BYTES;
    static constexpr int kBiasBytes = BLOCK_N * 4;
    static constexpr int kBarrierBytes = (1 + 2 * A_STAGES + 4 + 4 * NB) * 8 + 16;
    static constexpr int kTotal = kBResident + A_STAGES * kABytes + kBufBytes + kBiasBytes + kBarrierBytes;
};

template <int BLOCK_N, int KB, int A_STAGES, int NB, int EG>
__global__ void __launch_bounds__(threads_for(EG), 1)
conv_expand_ws_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                      const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_r,
                      const ConvParams prm) {
    using L = WsLayout<BLOCK_N, KB, A_STAGES, NB>;
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t* smem_b = smem;                                              // [KB][BLOCK_N x 64] resident weight slice
    uint8_t* smem_a = smem + L::kBResident;                              // [A_STAGES][128 x 64]
    uint8_t* smem_buf = smem_a + A_STAGES * L::kABytes;                  // [4 quadrants][NB][32 x 64]
    float* smem_bias = reinterpret_cast<float*>(smem_buf + L::kBufBytes);
    uint64_t* b_full = reinterpret_cast<uint64_t*>(smem_buf + L::kBufBytes + L::kBiasBytes);
    uint64_t* full_bar = b_full + 1;
    uint64_t* empty_bar = full_bar + A_STAGES;
    uint64_t* tmem_full = empty_bar + A_STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* res_bar = tmem_empty + 2;                                  // [4 quadrants][NB]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 4 * NB);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr uint32_t kTmemCols = 2 * BLOCK_N;
    const int cta_id = blockIdx.x, num_ctas = gridDim.x;                 // num_ctas % n_tiles == 0 (host): n_tile is fixed per CTA
    const int n_tile = cta_id % prm.n_tiles;
    const int num_tiles = prm.m_tiles * prm.n_tiles;

    if (warp == 0 && elect_one()) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        prefetch_tmap(&tmap_d);
        prefetch_tmap(&tmap_r);
        mbar_init(b_full, 1);
        for (int i = 0; i < A_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4 * EG);
        }
        for (int i = 0; i < 4 * NB; ++i) mbar_init(&res_bar[i], 1);
        fence_barrier_init();
    } else if (warp == 1) {
        tmem_alloc(tmem_slot, kTmemCols);
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == 0) {
        // ===================================== TMA producer: weights once, then A =====================================
        if (elect_one()) {
            mbar_expect_tx(b_full, (uint32_t)(prm.kblocks * L::kBBytes));
            for (int kb = 0; kb < prm.kblocks; ++kb)
                tma_load_2d(&tmap_b, b_full, smem_b + kb * L::kBBytes, kb * BLOCK_K, n_tile * BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cta_id; tile < num_tiles; tile += num_ctas) {
                const int m_tile = tile / prm.n_tiles;
                for (int kb = 0; kb < prm.kblocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], L::kABytes);
                    tma_load_2d(&tmap_a, &full_bar[stage], smem_a + stage * L::kABytes, kb * BLOCK_K, m_tile * BLOCK_M);
                    if (++stage == A_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ====================================== MMA issuer ======================================
        const uint32_t idesc = tc::umma_idesc_16bit(BLOCK_M, BLOCK_N, prm.is_f16 != 0);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        mbar_wait(b_full, 0);
        for (int tile = cta_id; tile < num_tiles; tile += num_ctas) {
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
            for (int kb = 0; kb < prm.kblocks; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tcgen05_fence_after();
                if (elect_one()) {
                    const uint64_t adesc = umma_desc_sw128(smem_u32(smem_a + stage * L::kABytes));
                    const uint64_t bdesc = umma_desc_sw128(smem_u32(smem_b + kb * L::kBBytes));
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                        umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    umma_commit(&empty_bar[stage]);
                    if (kb == prm.kblocks - 1) umma_commit(&tmem_full[acc]);
                }
                __syncwarp();
                if (++stage == A_STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ======================================= epilogue (in place) =======================================
        const int quad = warp & 3;
        const int grp = (warp - 2) >> 2;
        constexpr int WCOLS = EPI_COLS / EG;
        constexpr int WCHUNKS = WCOLS / 8;
        constexpr int kChunks = BLOCK_N / EPI_COLS;
        uint8_t* my_buf = smem_buf + quad * NB * EPI_BUF_BYTES;
        uint64_t* my_res_bar = res_bar + quad * NB;
        const int my_tiles = cta_id < num_tiles ? (num_tiles - 1 - cta_id) / num_ctas + 1 : 0;
        const int total_items = my_tiles * kChunks;
        const bool issuer = grp == 0 && lane == 0;          // issues the stores AND the residual loads of its quadrant
        auto quad_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(2 + quad), "n"(32 * EG) : "memory"); };
        const int n0 = n_tile * BLOCK_N;

        auto load_residual = [&](int item) {                // issuer only
            if (item >= total_items) return;
            const int m_tile = (cta_id + (item / kChunks) * num_ctas) / prm.n_tiles;
            const int b = item % NB;
            mbar_expect_tx(&my_res_bar[b], EPI_BUF_BYTES);
            tma_load_2d(&tmap_r, &my_res_bar[b], my_buf + b * EPI_BUF_BYTES, n0 + (item % kChunks) * EPI_COLS,
                        m_tile * BLOCK_M + quad * 32);
        };
        if (issuer) {
#pragma unroll
            for (int i = 0; i < NB; ++i) load_residual(i);
        }
        // the CTA's bias slice, once (n_tile never changes): every epilogue thread copies a part, all of them meet at barrier 1
        for (int i = (warp - 2) * 32 + lane; i < BLOCK_N; i += 128 * EG) smem_bias[i] = __ldg(prm.bias + n0 + i);
        asm volatile("bar.sync 1, %0;" ::"n"(128 * EG) : "memory");

        int acc = 0;
        uint32_t acc_phase = 0;
        int item = 0;
        const int sw = lane & 7;
        for (int tile = cta_id; tile < num_tiles; tile += num_ctas) {
            const int m_tile = tile / prm.n_tiles;
            mbar_wait(&tmem_full[acc], acc_phase);
            tcgen05_fence_after();
#pragma unroll 1
            for (int c = 0; c < kChunks; ++c, ++item) {
                uint32_t v[WCOLS];
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) +
                                       (uint32_t)(acc * BLOCK_N + c * EPI_COLS + grp * WCOLS);
                if constexpr (EG == 2) tmem_ld_32x32b_x32(taddr, v);
                else tmem_ld_32x32b_x16(taddr, v);
                tmem_ld_wait();
                const int b = item % NB;
                const float4* bias4 = reinterpret_cast<const float4*>(smem_bias + c * EPI_COLS + grp * WCOLS);
                uint8_t* row = my_buf + b * EPI_BUF_BYTES + lane * 128;
                mbar_wait(&my_res_bar[b], (uint32_t)((item / NB) & 1));      // the residual sub-tile has landed in buffer b
#pragma unroll
                for (int jj = 0; jj < WCHUNKS; ++jj) {
                    const int j = grp * WCHUNKS + jj;
                    const int phys = (j ^ sw) << 4;
                    float f[8];
                    const float4 b0 = bias4[2 * jj], b1 = bias4[2 * jj + 1];
                    f[0] = __uint_as_float(v[8 * jj]) + b0.x;     f[1] = __uint_as_float(v[8 * jj + 1]) + b0.y;
                    f[2] = __uint_as_float(v[8 * jj + 2]) + b0.z; f[3] = __uint_as_float(v[8 * jj + 3]) + b0.w;
                    f[4] = __uint_as_float(v[8 * jj + 4]) + b1.x; f[5] = __uint_as_float(v[8 * jj + 5]) + b1.y;
                    f[6] = __uint_as_float(v[8 * jj + 6]) + b1.z; f[7] = __uint_as_float(v[8 * jj + 7]) + b1.w;
                    const uint4 rv = *reinterpret_cast<const uint4*>(row + phys);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const float2 rf = prm.is_f16 ? __half22float2(reinterpret_cast<const __half2*>(&rv)[t])
                                                     : __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(&rv)[t]);
                        f[2 * t] += rf.x;
                        f[2 * t + 1] += rf.y;
                    }
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
                    *reinterpret_cast<uint4*>(row + phys) = ov;      // same 16 bytes this thread just read: in place
                }
                fence_proxy_async();
                quad_sync();                          // every warp of the quadrant has overwritten its columns of buffer b
                if (issuer) {
                    tma_store_2d(&tmap_d, my_buf + b * EPI_BUF_BYTES, n0 + c * EPI_COLS, m_tile * BLOCK_M + quad * 32);
                    bulk_commit();
                    if (item >= 1) {
                        bulk_wait_read<1>();          // the store of item - 1 has read its buffer: refill it NB - 1 items ahead
                        load_residual(item - 1 + NB);
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (issuer) bulk_wait_read<0>();
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---- host side: tensor maps through the driver entry points (no link-time libcuda dependency) -----------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn g_encode_tiled = nullptr;
EncodeIm2colFn g_encode_im2col = nullptr;
int g_driver_version = 0;
int g_sm_count = 0;
thread_local int t_max_ctas = 0;      // per-call cap on the persistent grid (dpft_conv2d_nhwc_ex), 0 = one CTA per SM

int cta_budget() { return (t_max_ctas > 0 && t_max_ctas < g_sm_count) ? t_max_ctas : g_sm_count; }

int resolve_driver() {
    if (g_encode_tiled && g_encode_im2col) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    int st = cuda_status(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q), "cuTensorMapEncodeTiled");
    if (st) return st;
    if (!fn) { set_error("cuTensorMapEncodeTiled not available in this driver"); return DPFT_ERR_UNSUPPORTED; }
    g_encode_tiled = (EncodeTiledFn)fn;
    fn = nullptr;
    st = cuda_status(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q), "cuTensorMapEncodeIm2col");
    if (st) return st;
    if (!fn) { set_error("cuTensorMapEncodeIm2col not available in this driver"); return DPFT_ERR_UNSUPPORTED; }
    g_encode_im2col = (EncodeIm2colFn)fn;
    cudaDriverGetVersion(&g_driver_version);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    return 0;
}

int encode_2d(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t row_bytes, uint32_t box_inner,
              uint32_t box_outer, bool is_f16) {
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(map, is_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(2d) failed with CUresult %d", (int)r); return DPFT_ERR_INVALID_ARGUMENT; }
    return 0;
}

// Programmatic dependent launch of the convolution chain; DPFT_CONV_PDL=0 in the environment turns it off (A/B timing).
bool use_pdl() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DPFT_CONV_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

// dpft_conv2d_nhwc_pair: the first problem is RECORDED (its tensor maps and parameters, and which kernel variant the
// dispatch chose), the second one is launched together with it.  Both take the same dispatch decisions (identical shapes).
enum PairMode : int { kPairOff = 0, kPairRecord = 1, kPairLaunch = 2 };
thread_local int t_pair_mode = kPairOff;
thread_local ConvLaunch t_pair_desc;
thread_local const void* t_pair_kernel = nullptr;

template <int BLOCK_N, int STAGES, int RES_BUFS, int CG = 1, int EG = 2>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const CUtensorMap& tr, const ConvParams& prm,
           cudaStream_t stream) {
    using L = SmemLayout<BLOCK_N, STAGES, RES_BUFS, CG>;
    static_assert(L::kTotal <= 232448, "shared memory budget of one CTA exceeded");
    auto kern = conv_gemm_kernel<BLOCK_N, STAGES, RES_BUFS, CG, EG>;
    if (t_pair_mode == kPairRecord) {
        t_pair_desc.ta[0] = ta; t_pair_desc.tb[0] = tb; t_pair_desc.td[0] = td; t_pair_desc.tr[0] = tr; t_pair_desc.prm[0] = prm;
        t_pair_kernel = (const void*)kern;
        return DPFT_OK;
    }
    int groups = 1;
    if (t_pair_mode == kPairLaunch) {
        if (t_pair_kernel != (const void*)kern) { set_error("conv2d_pair: the two problems chose different kernels"); return DPFT_ERR_INVALID_ARGUMENT; }
        groups = 2;
    }
    ConvLaunch desc;
    if (groups == 2) {
        desc = t_pair_desc;
        desc.ta[1] = ta; desc.tb[1] = tb; desc.td[1] = td; desc.tr[1] = tr; desc.prm[1] = prm;
    } else {
        desc.ta[0] = ta; desc.tb[0] = tb; desc.td[0] = td; desc.tr[0] = tr; desc.prm[0] = prm;
        desc.ta[1] = ta; desc.tb[1] = tb; desc.td[1] = td; desc.tr[1] = tr; desc.prm[1] = prm;
    }
    static bool configured = false;
    if (!configured) {
        int st = cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  L::kTotal), "cudaFuncSetAttribute(conv_gemm_kernel)");
        if (st) return st;
        configured = true;
    }
    const int tiles = ((prm.m_tiles + CG - 1) / CG) * prm.n_tiles;
    const int budget = cta_budget() / groups;                    // the two problems of a pair share the SMs
    const int max_ctas = budget / CG > 0 ? budget / CG : 1;
    const int grid = CG * (tiles < max_ctas ? tiles : max_ctas);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid, groups);
    cfg.blockDim = dim3(threads_for(EG));
    cfg.dynamicSmemBytes = L::kTotal;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int n_attr = 0;
    if (CG == 2) {
        attr[n_attr].id = cudaLaunchAttributeClusterDimension;
        attr[n_attr].val.clusterDim.x = CG;
        attr[n_attr].val.clusterDim.y = 1;
        attr[n_attr].val.clusterDim.z = 1;
        ++n_attr;
    }
    if (use_pdl()) {       // overlap this kernel's set-up with the previous kernel's tail (griddepcontrol.wait in the kernel)
        attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
        ++n_attr;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n_attr;
    int st = cuda_status(cudaLaunchKernelEx(&cfg, kern, desc), "cudaLaunchKernelEx(conv_gemm_kernel)");
    if (st) return st;
    DPFT_LAUNCH_CHECK("conv_gemm_kernel");
    return DPFT_OK;
}

template <int BLOCK_N, int KB, int A_STAGES, int NB, int EG = 4>
int launch_ws(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const CUtensorMap& tr, const ConvParams& prm,
              cudaStream_t stream) {
    if (t_pair_mode != kPairOff) { set_error("conv2d_pair: layer is served by the weight-stationary kernel"); return DPFT_ERR_UNSUPPORTED; }
    using L = WsLayout<BLOCK_N, KB, A_STAGES, NB>;
    static_assert(L::kTotal <= 232448, "shared memory budget of one CTA exceeded");
    auto kern = conv_expand_ws_kernel<BLOCK_N, KB, A_STAGES, NB, EG>;
    static bool configured = false;
    if (!configured) {
        int st = cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal),
                             "cudaFuncSetAttribute(conv_expand_ws_kernel)");
        if (st) return st;
        configured = true;
    }
    const int tiles = prm.m_tiles * prm.n_tiles;
    int grid = (cta_budget() / prm.n_tiles) * prm.n_tiles;    // a multiple of n_tiles: tile -> n_tile is constant per CTA
    if (grid < prm.n_tiles) grid = prm.n_tiles;
    if (grid > tiles) grid = tiles;                           // (tiles is a multiple of n_tiles as well)
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads_for(EG));
    cfg.dynamicSmemBytes = L::kTotal;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    int n_attr = 0;
    if (use_pdl()) {
        attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
        ++n_attr;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n_attr;
    int st = cuda_status(cudaLaunchKernelEx(&cfg, kern, ta, tb, td, tr, prm), "cudaLaunchKernelEx(conv_expand_ws_kernel)");
    if (st) return st;
    DPFT_LAUNCH_CHECK("conv_expand_ws_kernel");
    return DPFT_OK;
}

// DPFT_CONV_WS=0 in the environment keeps the expand layers on the generic kernel (A/B timing)
bool use_ws() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DPFT_CONV_WS");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

int pick_block_n(int Cout, int m_tiles, int want) {
    if (want == 64 || want == 128 || want == 256) return (Cout % want == 0) ? want : 64;
    // largest tile that still gives three quarters of the SMs a tile (measured on the stage-4 layers, 58 m-tiles: 256-wide
    // tiles on 116 CTAs beat 128-wide ones on 148; profiles/r01_conv_layers_v7_sweep.jsonl); small problems prefer more, smaller tiles
    const int cands[3] = {256, 128, 64};
    for (int i = 0; i < 3; ++i) {
        const int bn = cands[i];
        if (Cout % bn) continue;
        if ((long long)m_tiles * (Cout / bn) * 4 >= (long long)g_sm_count * 3 || bn == 64) return bn;
    }
    return 64;
}

}  // namespace
}  // namespace dpft

using namespace dpft;

extern "C" int dpft_fpn_lateral_forward(const void* x, const void* w, const float* bias, const float* coarse, int Hc, int Wc,
                                        float* out, int B, int H, int W, int Cin, int dtype, void* stream) {
    DPFT_REQUIRE(dtype == DPFT_BF16 || dtype == DPFT_F16, "fpn_lateral: dtype must be DPFT_BF16 or DPFT_F16");
    const bool is_f16 = dtype == DPFT_F16;
    DPFT_REQUIRE(x && w && bias && out, "fpn_lateral: null pointer");
    DPFT_REQUIRE(B > 0 && H > 0 && W > 0 && Cin % 64 == 0, "fpn_lateral: bad size B=%d H=%d W=%d Cin=%d", B, H, W, Cin);
    DPFT_REQUIRE(coarse == nullptr || (Hc > 0 && Wc > 0), "fpn_lateral: bad coarse size");
    t_max_ctas = 0;
    t_pair_mode = kPairOff;
    int st = resolve_driver();
    if (st) return st;
    ConvParams prm{};
    prm.M = B * H * W; prm.N = 64; prm.P = H; prm.Q = W; prm.taps_s = 1; prm.cblocks = Cin / 64; prm.kblocks = prm.cblocks;
    prm.stride = 1; prm.pad = 0; prm.relu = 0; prm.mode = kTiled2D; prm.bias = bias;
    prm.m_tiles = (prm.M + BLOCK_M - 1) / BLOCK_M; prm.n_tiles = 1;
    prm.out_f32 = out; prm.coarse = coarse; prm.Hc = Hc; prm.Wc = Wc; prm.is_f16 = is_f16;
    CUtensorMap ta, tb;
    st = encode_2d(&ta, x, (uint64_t)Cin, (uint64_t)prm.M, (uint64_t)Cin * 2, BLOCK_K, BLOCK_M, is_f16);
    if (st) return st;
    st = encode_2d(&tb, w, (uint64_t)Cin, 64, (uint64_t)Cin * 2, BLOCK_K, 64, is_f16);
    if (st) return st;
    return launch<64, 6, 3, 1>(ta, tb, ta, ta, prm, (cudaStream_t)stream);   // d / r maps unused in the fp32 lateral form
}

static int conv2d_nhwc_impl(const void* x, const void* w, const float* bias, const void* residual, void* y,
                            int B, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad, int relu,
                            int block_n, int cluster_mode, int dtype, void* stream);

extern "C" int dpft_conv2d_nhwc(const void* x, const void* w, const float* bias, const void* residual, void* y,
                                int B, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad, int relu,
                                int block_n, int cluster_mode, int dtype, void* stream) {
    t_max_ctas = 0;
    return conv2d_nhwc_impl(x, w, bias, residual, y, B, H, W, Cin, Cout, R, S, stride, pad, relu, block_n, cluster_mode, dtype, stream);
}

extern "C" int dpft_conv2d_nhwc_ex(const void* x, const void* w, const float* bias, const void* residual, void* y,
                                   int B, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad, int relu,
                                   int block_n, int cluster_mode, int dtype, int max_ctas, void* stream) {
    DPFT_REQUIRE(max_ctas >= 0, "conv2d: max_ctas=%d", max_ctas);
    t_max_ctas = max_ctas;
    const int st = conv2d_nhwc_impl(x, w, bias, residual, y, B, H, W, Cin, Cout, R, S, stride, pad, relu, block_n, cluster_mode, dtype, stream);
    t_max_ctas = 0;
    return st;
}

extern "C" int dpft_conv2d_nhwc_pair(const void* x0, const void* w0, const float* bias0, const void* residual0, void* y0,
                                     const void* x1, const void* w1, const float* bias1, const void* residual1, void* y1,
                                     int B, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad, int relu,
                                     int dtype, int max_ctas, void* stream) {
    DPFT_REQUIRE(max_ctas >= 0, "conv2d_pair: max_ctas=%d", max_ctas);
    DPFT_REQUIRE((residual0 == nullptr) == (residual1 == nullptr), "conv2d_pair: both problems need a residual, or neither");
    t_max_ctas = max_ctas;
    t_pair_mode = kPairRecord;
    int st = conv2d_nhwc_impl(x0, w0, bias0, residual0, y0, B, H, W, Cin, Cout, R, S, stride, pad, relu, 0, 0, dtype, stream);
    if (st == DPFT_OK) {
        t_pair_mode = kPairLaunch;
        st = conv2d_nhwc_impl(x1, w1, bias1, residual1, y1, B, H, W, Cin, Cout, R, S, stride, pad, relu, 0, 0, dtype, stream);
    }
    t_pair_mode = kPairOff;
    t_max_ctas = 0;
    return st;
}

static int conv2d_nhwc_impl(const void* x, const void* w, const float* bias, const void* residual, void* y,
                            int B, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad, int relu,
                            int block_n, int cluster_mode, int dtype, void* stream) {
    DPFT_REQUIRE(x && w && bias && y, "conv2d: null pointer");
    DPFT_REQUIRE(dtype == DPFT_BF16 || dtype == DPFT_F16, "conv2d: dtype must be DPFT_BF16 or DPFT_F16");
    const bool is_f16 = dtype == DPFT_F16;
    DPFT_REQUIRE(B > 0 && H > 0 && W > 0, "conv2d: bad input size %dx%dx%d", B, H, W);
    DPFT_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "conv2d: Cin=%d and Cout=%d must be multiples of 64", Cin, Cout);
    DPFT_REQUIRE(R >= 1 && S >= 1 && R <= 16 && S <= 16 && stride >= 1 && stride <= 8 && pad >= 0 && pad < 16,
                 "conv2d: unsupported filter %dx%d stride %d pad %d", R, S, stride, pad);
    DPFT_REQUIRE((((uintptr_t)x | (uintptr_t)w | (uintptr_t)y | (uintptr_t)residual) & 15) == 0, "conv2d: pointers must be 16-byte aligned");
    int st = resolve_driver();
    if (st) return st;
    const int P = (H + 2 * pad - R) / stride + 1, Q = (W + 2 * pad - S) / stride + 1;
    DPFT_REQUIRE(P > 0 && Q > 0, "conv2d: empty output");
    // 64 -> 64 channel 3x3 layers on wide maps: halo-tile kernel (input staged once instead of once per tap)
    if (block_n == 0 && cluster_mode == 0 && conv3x3_halo_eligible(H, W, Cin, Cout, R, S, stride, pad, residual != nullptr)) {
        if (t_pair_mode != kPairOff) { set_error("conv2d_pair: layer is served by the halo kernel"); return DPFT_ERR_UNSUPPORTED; }
        return conv3x3_halo_launch(x, w, bias, y, B, H, W, relu, is_f16, (cudaStream_t)stream);
    }
    ConvParams prm{};
    prm.M = B * P * Q; prm.N = Cout; prm.P = P; prm.Q = Q; prm.taps_s = S; prm.cblocks = Cin / 64;
    prm.kblocks = R * S * prm.cblocks; prm.stride = stride; prm.pad = pad; prm.relu = relu; prm.is_f16 = is_f16;
    prm.bias = bias; prm.residual = (const __nv_bfloat16*)residual; prm.out = (__nv_bfloat16*)y;
    prm.m_tiles = (prm.M + BLOCK_M - 1) / BLOCK_M;
    const int bn = pick_block_n(Cout, prm.m_tiles, block_n);
    prm.n_tiles = Cout / bn;

    CUtensorMap ta, tb;
    const bool pointwise = (R == 1 && S == 1 && stride == 1 && pad == 0);
    if (pointwise) {
        prm.mode = kTiled2D;
        st = encode_2d(&ta, x, (uint64_t)Cin, (uint64_t)prm.M, (uint64_t)Cin * 2, BLOCK_K, BLOCK_M, is_f16);
        if (st) return st;
    } else {
        prm.mode = kIm2col;
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
        int lower[2] = {-pad, -pad};
        int upper[2] = {pad - (S - 1), pad - (R - 1)};
        cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
        CUresult r = g_encode_im2col(&ta, is_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, lower, upper,
                                     BLOCK_K, BLOCK_M, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeIm2col failed with CUresult %d", (int)r); return DPFT_ERR_INVALID_ARGUMENT; }
        // Same small-tensor fix-up CUTLASS applies for drivers <= 13.1 (cute/atom/copy_traits_sm90_im2col.hpp).
        if (g_driver_version <= 13010 && (uint64_t)B * H * W * Cin * 2 < 131072)
            reinterpret_cast<uint64_t*>(&ta)[1] &= ~(1ull << 21);
    }
    // CTA pairs (cta_group::2) for the operand-bound layers: long K loop and enough 256-row tiles; each CTA then loads only
    // half of the B tile.  cluster_mode: 0 = choose, 1 = single CTA, 2 = force pairs (tests).
    const bool pairs = bn >= 128 && (cluster_mode == 2 || (cluster_mode == 0 && prm.kblocks >= 8 &&
                       (long long)((prm.m_tiles + 1) / 2) * prm.n_tiles >= g_sm_count / 4));
    st = encode_2d(&tb, w, (uint64_t)R * S * Cin, (uint64_t)Cout, (uint64_t)R * S * Cin * 2, BLOCK_K, pairs ? bn / 2 : bn, is_f16);
    if (st) return st;
    // output / residual as [M, Cout] matrices, written / read in [32 rows x 64 cols] boxes (128-byte swizzle)
    CUtensorMap td, tr;
    st = encode_2d(&td, y, (uint64_t)Cout, (uint64_t)prm.M, (uint64_t)Cout * 2, EPI_COLS, 32, is_f16);
    if (st) return st;
    st = encode_2d(&tr, residual ? residual : y, (uint64_t)Cout, (uint64_t)prm.M, (uint64_t)Cout * 2, EPI_COLS, 32, is_f16);
    if (st) return st;
    cudaStream_t s = (cudaStream_t)stream;
    // weight-stationary kernel for the expand 1x1 + residual layers with enough tiles to amortise the resident weight slice
    // (cluster_mode 3 forces it: tests), K = 128 .. 256.  Measured on the camera ResNet-101 at bs 8 (profiles/r02_conv_expand_*):
    // what bounds these layers is bytes in flight per SM (operand and residual latency), not L2 or HBM bandwidth — 128-wide
    // n-tiles leave room for six to eight A stages next to the resident weights and a four-deep in-place ring: stage-3 conv3
    // 37.8 -> 31.7 us, stage-2 conv3 52.2 -> 49.2 us.  K = 64 (stage 1) is HBM-bound at 0.9 of the copy peak on the generic
    // kernel and stays there (the resident slice buys nothing: 96 vs 90 us).
    if (pointwise && residual != nullptr && Cout % 256 == 0 && prm.kblocks <= 4 && (block_n == 0 || block_n == 256) &&
        (cluster_mode == 3 || (cluster_mode == 0 && use_ws() && prm.kblocks >= 2 &&
                               (long long)prm.m_tiles * (Cout / 256) >= 2LL * g_sm_count))) {
        // DPFT_WS_VARIANT=10 (tools/conv_bench.py) keeps the first, 256-wide split for A/B timing; the other splits that were
        // measured (A stages 2..8, ring depth 2..6, two epilogue warps per quadrant) are in profiles/r02_conv_expand_ws_variants.txt
        static const int ws_variant = [] { const char* e = getenv("DPFT_WS_VARIANT"); return e ? atoi(e) : 0; }();
        const int wbn = (ws_variant == 10 || prm.kblocks == 1) ? 256 : 128;
        prm.n_tiles = Cout / wbn;
        if (prm.n_tiles <= g_sm_count) {
            st = encode_2d(&tb, w, (uint64_t)Cin, (uint64_t)Cout, (uint64_t)Cin * 2, BLOCK_K, wbn, is_f16);
            if (st) return st;
            if (prm.kblocks == 1) return launch_ws<256, 1, 4, 8>(ta, tb, td, tr, prm, s);
            if (ws_variant == 10) return prm.kblocks <= 2 ? launch_ws<256, 2, 4, 6>(ta, tb, td, tr, prm, s)
                                                          : launch_ws<256, 4, 2, 4>(ta, tb, td, tr, prm, s);
            if (prm.kblocks <= 2) return launch_ws<128, 2, 8, 4>(ta, tb, td, tr, prm, s);
            return launch_ws<128, 4, 6, 4>(ta, tb, td, tr, prm, s);
        }
        prm.n_tiles = Cout / bn;
    }
    DPFT_REQUIRE(cluster_mode != 3, "conv2d: cluster_mode 3 (weight-stationary) needs a 1x1 stride-1 layer with a residual, Cout %% 256 == 0, Cin <= 256");
    if (pairs) {
        cudaStream_t s2 = (cudaStream_t)stream;
        if (residual != nullptr && prm.kblocks <= 4) {   // forced pairs on an expand layer (cluster_mode 2): the stream-tuned split
            if (bn == 256) return launch<256, 2, 5, 2, 4>(ta, tb, td, tr, prm, s2);
            return launch<128, 2, 7, 2, 4>(ta, tb, td, tr, prm, s2);
        }
        // (four epilogue warps per lane quadrant were tried here as well: no gain, 37.9 -> 38.9 us on s3_conv2)
        // layers without a residual do not need the residual sub-tile ring: it is traded for operand stages (more bytes in
        // flight per SM; s3_conv2 38.9 -> 36.9 us, s4_conv2 38.9 -> 35.8 us, s4_conv1 25.6 -> 23.5 us, whole step -1.7 %,
        // profiles/r02_conv_deep_variants.txt).  DPFT_CONV_DEEP_VARIANT=0 restores the 4 / 6-stage split for A/B timing
        static const int deep_variant = [] { const char* e = getenv("DPFT_CONV_DEEP_VARIANT"); return e ? atoi(e) : 1; }();
        if (deep_variant == 1 && residual == nullptr) {
            if (bn == 256) return launch<256, 5, 1, 2>(ta, tb, td, tr, prm, s2);
            return launch<128, 7, 1, 2>(ta, tb, td, tr, prm, s2);
        }
        if (bn == 256) return launch<256, 4, 3, 2>(ta, tb, td, tr, prm, s2);
        return launch<128, 6, 3, 2>(ta, tb, td, tr, prm, s2);
    }
    const bool stream_bound = residual != nullptr && prm.kblocks <= 4;   // 1x1 expand convs: deep residual prefetch
    // Stream-bound layers run FOUR epilogue warps per TMEM lane quadrant (16 columns of an item each, 18 warps per CTA): the
    // epilogue is a chain of dependent short operations (TMEM load, smem residual, pack, staging store), and four warps
    // per scheduler hide its latency better than two (s1_conv3 97 -> 89 us = 0.91 of the HBM peak, whole step -1.7 %).
    if (stream_bound) {
        if (bn == 256) return launch<256, 2, 5, 1, 4>(ta, tb, td, tr, prm, s);
        if (bn == 128) return launch<128, 2, 7, 1, 4>(ta, tb, td, tr, prm, s);
        return launch<64, 3, 7, 1, 4>(ta, tb, td, tr, prm, s);
    }
    if (bn == 256) return launch<256, 3, 2>(ta, tb, td, tr, prm, s);
    if (bn == 128) return launch<128, 4, 3>(ta, tb, td, tr, prm, s);
    return launch<64, 6, 3>(ta, tb, td, tr, prm, s);
}
