// Dense feature path around the tcgen05 bottleneck convolutions, for sm_100a (inference):
//
//   stem_conv7x7     raw fp32 NHWC input -> 7x7/2 conv (+ folded BatchNorm, + folded 1x1 adjustment conv for the
//                    6-channel radar cubes) + ReLU -> bf16 NHWC, 64 channels
//                    (reference src/dprt/models/backbones/resnet.py:98-101: adjustment_layer, conv1, bn1, relu)
//   maxpool3x3s2     torchvision ResNet maxpool (kernel 3, stride 2, padding 1) on bf16 NHWC
//   fpn_output       the FPN output stage of one level (reference src/dprt/models/necks/fpn.py:70-83 over
//                    torchvision FeaturePyramidNetwork): 3x3 conv (16 -> 16, zero padding) + bias, fused with the
//                    sinusoidal positional embedding (src/dprt/models/embeddings/sinusoidal.py:107-108) and written
//                    straight into the view's feature pyramid buffer (B, S, 16) — the tensor the reference builds with
//                    torch.cat in every decoder layer (src/dprt/models/fusers/mpfusion.py:179).
//                    For the finest level (the raw input, "skip link" dprt.py:222-225) the lateral 1x1 conv and the
//                    top-down nearest-upsample-add are computed on the fly into the shared-memory halo tile, so the
//                    full-resolution 16-channel "inner" map never touches HBM.
//
// All three are HBM-bound (N = 16 output channels, tiny K): coalesced 16-byte accesses, halo tiles in shared memory.
#include <type_traits>

#include <stdlib.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace dpft {
namespace {

// --------------------------------------------------------------------------------------------------------- stem
constexpr int STEM_TH = 8, STEM_TW = 16;             // output tile
constexpr int STEM_PH = STEM_TH * 2 + 5, STEM_PW = STEM_TW * 2 + 5;   // input patch
constexpr int STEM_COUT = 64;

template <int CIN, typename OT>
__global__ void __launch_bounds__(256)
stem_conv7x7_kernel(const float* __restrict__ x, const float* __restrict__ w /* [49*CIN][64] */,
                    const float* __restrict__ bias, OT* __restrict__ y, int H, int W, int P, int Q, float lo) {
    extern __shared__ __align__(16) float smem[];
    float* s_w = smem;                               // [49*CIN][64]
    float* s_x = smem + 49 * CIN * STEM_COUT;        // [PH][PW][CIN]
    const int b = blockIdx.z;
    const int p0 = blockIdx.y * STEM_TH, q0 = blockIdx.x * STEM_TW;
    const int tid = threadIdx.x;
    for (int i = tid; i < 49 * CIN * STEM_COUT / 4; i += 256)
        reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(w) + i);
    const int h_base = p0 * 2 - 3, w_base = q0 * 2 - 3;
    for (int i = tid; i < STEM_PH * STEM_PW * CIN; i += 256) {
        const int c = i % CIN;
        const int pw = (i / CIN) % STEM_PW;
        const int ph = i / (CIN * STEM_PW);
        const int hh = h_base + ph, ww = w_base + pw;
        float v = 0.0f;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = __ldg(x + (((long long)b * H + hh) * W + ww) * CIN + c);
        s_x[i] = v;
    }
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    const int half = warp & 1;                       // which 32 output channels
    const int pix = (warp >> 1) * 32 + lane;         // 0..127
    const int ty = pix / STEM_TW, tx = pix % STEM_TW;
    float acc[32];
#pragma unroll
    for (int o = 0; o < 32; ++o) acc[o] = __ldg(bias + half * 32 + o);
    for (int r = 0; r < 7; ++r) {
        for (int s = 0; s < 7; ++s) {
            const float* xp = s_x + ((2 * ty + r) * STEM_PW + 2 * tx + s) * CIN;
            const float* wp = s_w + ((r * 7 + s) * CIN) * STEM_COUT + half * 32;
#pragma unroll
            for (int c = 0; c < CIN; ++c) {
                const float xv = xp[c];
                const float4* w4 = reinterpret_cast<const float4*>(wp + c * STEM_COUT);
#pragma unroll
                for (int o4 = 0; o4 < 8; ++o4) {
                    const float4 wv = w4[o4];
                    acc[4 * o4] = fmaf(xv, wv.x, acc[4 * o4]);
                    acc[4 * o4 + 1] = fmaf(xv, wv.y, acc[4 * o4 + 1]);
                    acc[4 * o4 + 2] = fmaf(xv, wv.z, acc[4 * o4 + 2]);
                    acc[4 * o4 + 3] = fmaf(xv, wv.w, acc[4 * o4 + 3]);
                }
            }
        }
    }
    const int p = p0 + ty, q = q0 + tx;
    if (p < P && q < Q) {
        OT* o = y + (((long long)b * P + p) * Q + q) * STEM_COUT + half * 32;
#pragma unroll
        for (int o8 = 0; o8 < 4; ++o8) {
            uint4 pk;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float a0 = fmaxf(acc[8 * o8 + 2 * t], lo), a1 = fmaxf(acc[8 * o8 + 2 * t + 1], lo);   // lo = 0: ReLU
                if constexpr (sizeof(OT) == 2 && std::is_same<OT, __half>::value)
                    reinterpret_cast<__half2*>(&pk)[t] = __floats2half2_rn(fminf(a0, 65504.0f), fminf(a1, 65504.0f));
                else
                    reinterpret_cast<__nv_bfloat162*>(&pk)[t] = __floats2bfloat162_rn(a0, a1);
            }
            reinterpret_cast<uint4*>(o)[o8] = pk;
        }
    }
}

// ------------------------------------------------------------------------------------------------------ maxpool
template <typename T2>   // __nv_bfloat162 or __half2
__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y, int B, int H, int W, int C8,
                    int P, int Q) {
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long total = (long long)B * P * Q * C8;
    if (idx >= total) return;
    const int c8 = (int)(idx % C8);
    const int q = (int)((idx / C8) % Q);
    const int p = (int)((idx / ((long long)C8 * Q)) % P);
    const int b = (int)(idx / ((long long)C8 * Q * P));
    T2 m[4];
    T2 ninf;
    if constexpr (std::is_same<T2, __half2>::value) ninf = __floats2half2_rn(-INFINITY, -INFINITY);
    else ninf = __floats2bfloat162_rn(-INFINITY, -INFINITY);
#pragma unroll
    for (int t = 0; t < 4; ++t) m[t] = ninf;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int hh = 2 * p - 1 + r;
        if (hh < 0 || hh >= H) continue;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const int ww = 2 * q - 1 + s;
            if (ww < 0 || ww >= W) continue;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + (((long long)b * H + hh) * W + ww) * C8 * 8) + c8);
            const T2* vb = reinterpret_cast<const T2*>(&v);
#pragma unroll
            for (int t = 0; t < 4; ++t) m[t] = __hmax2(m[t], vb[t]);
        }
    }
    uint4 o;
    T2* ob = reinterpret_cast<T2*>(&o);
#pragma unroll
    for (int t = 0; t < 4; ++t) ob[t] = m[t];
    reinterpret_cast<uint4*>(y + (((long long)b * P + p) * Q + q) * C8 * 8)[c8] = o;
}

// --------------------------------------------------------------------------------------------------- FPN output
constexpr int FPN_TH = 8, FPN_TW = 32;
constexpr int FPN_PH = FPN_TH + 2, FPN_PW = FPN_TW + 2;
constexpr int FC = 16;

struct FpnOutParams {
    const float* inner;        // (B, H, W, 16) fp32 — levels fed by the lateral GEMM; null when from_raw
    const float* raw;          // (B, H, W, CIN) fp32 raw input (finest level with skip link)
    const float* lat_w;        // [16][CIN] lateral weights of the raw level
    const float* lat_b;        // [16]
    const float* coarse;       // (B, Hc, Wc, 16) fp32 inner map of the next coarser level (top-down), or null
    const float* w;            // [9][16 out][16 in] 3x3 weights
    const uint4* w_packed;     // f16 UMMA image of w (tensor-core kernel), 4608 B
    const float* bias;         // [16]
    const float* pos_y;        // (H, 16)
    const float* pos_x;        // (W, 16)
    void* pyramid;             // (B, S, 16) fp32 or f16
    int pyramid_f16;
    long long S, start;
    int H, W, Hc, Wc;
    int raw_u8;                // raw is (B, H, W, CIN) uint8 (what an image decoder produces): converted on load, values exact
    float lat_c[16 * 6];       // the raw level's lateral weights by value (kernel parameter = constant bank): the column builder's
                               // FMAs then take them as constant operands instead of 12 shared-memory loads per halo entry
};

// element idx of the raw input as fp32
__device__ __forceinline__ float raw_ld(const FpnOutParams& prm, long long idx) {
    return prm.raw_u8 ? (float)__ldg(reinterpret_cast<const unsigned char*>(prm.raw) + idx) : __ldg(prm.raw + idx);
}

// one pyramid row (16 channels of one pixel): 64 B as fp32 or 32 B as f16 (saturating)
__device__ __forceinline__ void store_pyramid_row(void* pyramid, long long pixel, const float* v, bool f16) {
    if (f16) {
        uint4 pk[2];
        uint32_t* w = reinterpret_cast<uint32_t*>(pk);
#pragma unroll
        for (int t = 0; t < 8; ++t) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(w[t]) : "f"(v[2 * t + 1]), "f"(v[2 * t]));
        uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(pyramid) + pixel * FC);
        o[0] = pk[0];
        o[1] = pk[1];
    } else {
        float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(pyramid) + pixel * FC);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) o[c4] = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
    }
}

__device__ __forceinline__ int nearest_src(int dst, int in_size, int out_size) {
    // torch 'nearest': src = min(floor(dst * (float(in) / out)), in - 1)
    const float scale = (float)in_size / (float)out_size;
    const int s = (int)floorf((float)dst * scale);
    return s < in_size - 1 ? s : in_size - 1;
}

template <int CIN>   // CIN == 0: inner map comes from global memory
__global__ void __launch_bounds__(256)
fpn_output_kernel(const FpnOutParams prm) {
    __shared__ float4 s_in[4][FPN_PH * FPN_PW];     // channel planes of the halo tile
    __shared__ __align__(16) float s_w[9 * FC * FC];
    __shared__ float s_lat[FC * (CIN > 0 ? CIN : 1) + FC];
    const int b = blockIdx.z;
    const int p0 = blockIdx.y * FPN_TH, q0 = blockIdx.x * FPN_TW;
    const int tid = threadIdx.x;
    const int H = prm.H, W = prm.W;
    for (int i = tid; i < 9 * FC * FC / 4; i += 256)
        reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(prm.w) + i);
    if (CIN > 0) {
        for (int i = tid; i < FC * CIN; i += 256) s_lat[i] = __ldg(prm.lat_w + i);
        if (tid < FC) s_lat[FC * CIN + tid] = __ldg(prm.lat_b + tid);
        __syncthreads();
    }
    // halo tile of the "inner" map (zero outside the image: the 3x3 conv pads with zeros)
    for (int i = tid; i < FPN_PH * FPN_PW; i += 256) {
        const int ph = i / FPN_PW, pw = i % FPN_PW;
        const int hh = p0 - 1 + ph, ww = q0 - 1 + pw;
        float v[FC];
#pragma unroll
        for (int c = 0; c < FC; ++c) v[c] = 0.0f;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
            if (CIN > 0) {
                float xin[CIN > 0 ? CIN : 1];
                const long long xi = (((long long)b * H + hh) * W + ww) * CIN;
#pragma unroll
                for (int c = 0; c < CIN; ++c) xin[c] = raw_ld(prm, xi + c);
#pragma unroll
                for (int o = 0; o < FC; ++o) {
                    float a = s_lat[FC * CIN + o];
#pragma unroll
                    for (int c = 0; c < CIN; ++c) a = fmaf(s_lat[o * CIN + c], xin[c], a);
                    v[o] = a;
                }
                if (prm.coarse) {
                    const int hc = nearest_src(hh, prm.Hc, H), wc = nearest_src(ww, prm.Wc, W);
                    const float4* cp = reinterpret_cast<const float4*>(prm.coarse + (((long long)b * prm.Hc + hc) * prm.Wc + wc) * FC);
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const float4 f = __ldg(cp + c4);
                        v[4 * c4] += f.x; v[4 * c4 + 1] += f.y; v[4 * c4 + 2] += f.z; v[4 * c4 + 3] += f.w;
                    }
                }
            } else {
                const float4* ip = reinterpret_cast<const float4*>(prm.inner + (((long long)b * H + hh) * W + ww) * FC);
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const float4 f = __ldg(ip + c4);
                    v[4 * c4] = f.x; v[4 * c4 + 1] = f.y; v[4 * c4 + 2] = f.z; v[4 * c4 + 3] = f.w;
                }
            }
        }
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) s_in[c4][i] = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
    }
    __syncthreads();

    const int ty = tid / FPN_TW, tx = tid % FPN_TW;
    float acc[FC];
#pragma unroll
    for (int o = 0; o < FC; ++o) acc[o] = __ldg(prm.bias + o);
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
        const int r = tap / 3, s = tap % 3;
        const int pi = (ty + r) * FPN_PW + tx + s;
        float in[FC];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
            const float4 f = s_in[c4][pi];
            in[4 * c4] = f.x; in[4 * c4 + 1] = f.y; in[4 * c4 + 2] = f.z; in[4 * c4 + 3] = f.w;
        }
        const float4* w4 = reinterpret_cast<const float4*>(s_w + tap * FC * FC);
#pragma unroll
        for (int o = 0; o < FC; ++o) {
            float a = acc[o];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float4 wv = w4[o * 4 + c4];
                a = fmaf(wv.x, in[4 * c4], a);
                a = fmaf(wv.y, in[4 * c4 + 1], a);
                a = fmaf(wv.z, in[4 * c4 + 2], a);
                a = fmaf(wv.w, in[4 * c4 + 3], a);
            }
            acc[o] = a;
        }
    }
    const int p = p0 + ty, q = q0 + tx;
    if (p < H && q < W) {
        const float4* px = reinterpret_cast<const float4*>(prm.pos_x + (long long)q * FC);
        const float4* py = reinterpret_cast<const float4*>(prm.pos_y + (long long)p * FC);
        float outv[FC];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
            const float4 a = __ldg(px + c4), c = __ldg(py + c4);
            // reference order: feat += pos_x; feat += pos_y (sinusoidal.py:107-108)
            outv[4 * c4] = (acc[4 * c4] + a.x) + c.x;
            outv[4 * c4 + 1] = (acc[4 * c4 + 1] + a.y) + c.y;
            outv[4 * c4 + 2] = (acc[4 * c4 + 2] + a.z) + c.z;
            outv[4 * c4 + 3] = (acc[4 * c4 + 3] + a.w) + c.w;
        }
        store_pyramid_row(prm.pyramid, (long long)b * prm.S + prm.start + (long long)p * W + q, outv, prm.pyramid_f16 != 0);
    }
}


// ------------------------------------------------------------------------------- FPN output on the tensor cores
// Same computation as fpn_output_kernel, as an implicit GEMM per output row: M = 128 pixels of one image row,
// N = 16 output channels, K = 9 taps x 16 input channels.  The inner halo tile ((TH+2) rows x 130 pixels x 16 channels)
// is built by the threads in f16 and laid out as the canonical un-swizzled K-major UMMA operand: two channel-half
// planes per row, 16 bytes per pixel, so that for tap (dr, ds) the A operand of output row r is simply the smem
// window starting at inner row r+dr, pixel ds (start address + ds*16 B; 8-pixel core matrices 128 B apart, the two
// K halves one plane apart).  One thread issues the TH x 9 tcgen05.mma (each output row has its own 16 TMEM columns
// and its own mbarrier); 128 threads then read row after row from TMEM, add bias + positional embedding and write
// 64 contiguous bytes per pixel.
constexpr int TC_TH = 8, TC_TW = 128;
constexpr int TC_PW = 136;                           // padded pixels per inner row (>= TC_TW + 2)
constexpr int TC_PLANE = TC_PW * 16;                 // bytes of one channel-half plane of one inner row
constexpr int TC_ROWB = 2 * TC_PLANE;
constexpr int TC_A_BYTES = (TC_TH + 2) * TC_ROWB;    // 43,520
constexpr int TC_B_BYTES = 9 * 512;                  // 9 taps x [2 K-halves][2 N-groups][8 rows][16 B]

constexpr int TC_THREADS = 256;

// w[tap][o][c] fp32 -> f16 UMMA image [tap][khalf][ngroup][8 rows][8 elems], once per model
__global__ void fpn_pack_weights_kernel(const float* __restrict__ w, __half* __restrict__ packed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 9 * 16 * 16) return;
    const int c = i & 15, o = (i >> 4) & 15, tap = i >> 8;
    packed[tap * 256 + (c >> 3) * 128 + (o >> 3) * 64 + (o & 7) * 8 + (c & 7)] = __float2half_rn(w[i]);
}

template <int CIN>   // CIN == 0: inner map comes from global memory
__global__ void __launch_bounds__(TC_THREADS)
fpn_output_tc_kernel(const FpnOutParams prm) {
    extern __shared__ __align__(128) uint8_t tsm[];
    uint8_t* s_a = tsm;
    uint8_t* s_b = tsm + TC_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_b + TC_B_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + TC_TH);
    float* s_lat = reinterpret_cast<float*>(tmem_slot + 4);          // [16][CIN] + [16]
    float* s_py = s_lat + FC * 6 + FC;                                // [TC_TH][16]: pos_y rows of this tile (+ bias)
    const int b = blockIdx.z;
    const int p0 = blockIdx.y * TC_TH, q0 = blockIdx.x * TC_TW;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int H = prm.H, W = prm.W;

    if (warp == 0) {
        tc::tmem_alloc(tmem_slot, 128);
    } else if (tid == 32) {
        for (int r = 0; r < TC_TH; ++r) tc::mbar_init(&bars[r], 1);
        tc::fence_barrier_init();
    }
    for (int i = tid; i < TC_B_BYTES / 16; i += TC_THREADS) reinterpret_cast<uint4*>(s_b)[i] = __ldg(prm.w_packed + i);
    if (tid < TC_TH * FC) {                      // pos_y[p] + bias for the tile's rows: the epilogue reads them from shared memory
        const int r = tid / FC, c = tid - r * FC;
        s_py[tid] = (p0 + r < H ? __ldg(prm.pos_y + (long long)(p0 + r) * FC + c) : 0.0f) + __ldg(prm.bias + c);
    }
    if (CIN > 0) {
        for (int i = tid; i < FC * CIN; i += TC_THREADS) s_lat[i] = __ldg(prm.lat_w + i);
        if (tid < FC) s_lat[FC * CIN + tid] = __ldg(prm.lat_b + tid);
        __syncthreads();
    }
    // inner halo tile (zero outside the image), f16
    // (batches of TC_BATCH entries per thread: every global load of a batch is issued before any is consumed)
    constexpr int TC_ENTRIES = (TC_TH + 2) * TC_PW;
    constexpr int TC_BATCH = 3;
    for (int base = tid; base < TC_ENTRIES; base += TC_BATCH * TC_THREADS) {
        float xin[TC_BATCH][CIN > 0 ? CIN : 1];
        float4 cf[TC_BATCH][4];
        bool okv[TC_BATCH];
#pragma unroll
        for (int u = 0; u < TC_BATCH; ++u) {
            const int i = base + u * TC_THREADS;
            const int rr = i / TC_PW, px = i - rr * TC_PW;
            const int hh = p0 - 1 + rr, ww = q0 - 1 + px;
            const bool ok = i < TC_ENTRIES && px < TC_TW + 2 && hh >= 0 && hh < H && ww >= 0 && ww < W;
            okv[u] = ok;
            const int hs = ok ? hh : 0, ws = ok ? ww : 0;
            if (CIN > 0) {
                const long long xi = (((long long)b * H + hs) * W + ws) * CIN;
#pragma unroll
                for (int c = 0; c < CIN; ++c) xin[u][c] = ok ? raw_ld(prm, xi + c) : 0.0f;
                if (prm.coarse) {
                    const int hc = nearest_src(hs, prm.Hc, H), wc = nearest_src(ws, prm.Wc, W);
                    const float4* cp = reinterpret_cast<const float4*>(prm.coarse + (((long long)b * prm.Hc + hc) * prm.Wc + wc) * FC);
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) cf[u][c4] = ok ? __ldg(cp + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
                } else {
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) cf[u][c4] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {
                const float4* ip = reinterpret_cast<const float4*>(prm.inner + (((long long)b * H + hs) * W + ws) * FC);
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) cf[u][c4] = ok ? __ldg(ip + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int u = 0; u < TC_BATCH; ++u) {
            const int i = base + u * TC_THREADS;
            if (i >= TC_ENTRIES) continue;
            const int rr = i / TC_PW, px = i - rr * TC_PW;
            float v[FC];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                v[4 * c4] = cf[u][c4].x; v[4 * c4 + 1] = cf[u][c4].y; v[4 * c4 + 2] = cf[u][c4].z; v[4 * c4 + 3] = cf[u][c4].w;
            }
            if (CIN > 0 && okv[u]) {
#pragma unroll
                for (int o = 0; o < FC; ++o) {
                    float a = s_lat[FC * CIN + o];
#pragma unroll
                    for (int c = 0; c < CIN; ++c) a = fmaf(s_lat[o * CIN + c], xin[u][c], a);
                    v[o] += a;
                }
            }
            uint4 pk[2];
            uint32_t* pw = reinterpret_cast<uint32_t*>(pk);
#pragma unroll
            for (int t = 0; t < 8; ++t) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pw[t]) : "f"(v[2 * t + 1]), "f"(v[2 * t]));
            *reinterpret_cast<uint4*>(s_a + rr * TC_ROWB + px * 16) = pk[0];
            *reinterpret_cast<uint4*>(s_a + rr * TC_ROWB + TC_PLANE + px * 16) = pk[1];
        }
    }
    tc::fence_proxy_async();                 // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (tid == 0) {
        constexpr uint32_t idesc = tc::umma_idesc_16bit(128, 16, true);
        const uint32_t a0 = tc::smem_u32(s_a), b0 = tc::smem_u32(s_b);
        for (int r = 0; r < TC_TH; ++r) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int dr = tap / 3, ds = tap % 3;
                const uint64_t adesc = tc::umma_desc_noswizzle(a0 + (r + dr) * TC_ROWB + ds * 16, TC_PLANE, 128);
                const uint64_t bdesc = tc::umma_desc_noswizzle(b0 + tap * 512, 256, 128);
                tc::umma_bf16(tmem_base + r * 16, adesc, bdesc, idesc, tap ? 1u : 0u);
            }
            tc::umma_commit(&bars[r]);
        }
    }
    __syncwarp();
    // epilogue: all eight warps; warps w and w+4 share TMEM lane quadrant w%4 (pixels q0 + 32*(w%4) + lane) and take the
    // first / second half of the tile's rows
    const int lane_px = (warp & 3) * 32 + (tid & 31);
    const int q = q0 + lane_px;
    float4 px4[4];
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4)
        px4[c4] = q < W ? __ldg(reinterpret_cast<const float4*>(prm.pos_x + (long long)q * FC) + c4) : make_float4(0, 0, 0, 0);
    const int r_begin = (warp >> 2) * (TC_TH / 2);
    for (int r = r_begin; r < r_begin + TC_TH / 2; ++r) {
        tc::mbar_wait(&bars[r], 0);
        tc::tcgen05_fence_after();
        uint32_t v[16];
        tc::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(r * 16), v);
        tc::tmem_ld_wait();
        const int p = p0 + r;
        if (p < H && q < W) {
            const float4* py = reinterpret_cast<const float4*>(s_py + r * FC);     // pos_y + bias
            float outv[FC];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float4 c = py[c4];
                outv[4 * c4] = (__uint_as_float(v[4 * c4]) + c.x) + px4[c4].x;
                outv[4 * c4 + 1] = (__uint_as_float(v[4 * c4 + 1]) + c.y) + px4[c4].y;
                outv[4 * c4 + 2] = (__uint_as_float(v[4 * c4 + 2]) + c.z) + px4[c4].z;
                outv[4 * c4 + 3] = (__uint_as_float(v[4 * c4 + 3]) + c.w) + px4[c4].w;
            }
            store_pyramid_row(prm.pyramid, (long long)b * prm.S + prm.start + (long long)p * W + q, outv, prm.pyramid_f16 != 0);
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc::tcgen05_fence_after();
        tc::tmem_dealloc(tmem_base, 128);
    }
}


// ----------------------------------------------------------- FPN output, raw level: column-owning tile builder (v2)
// Default raw-level builder since round 2 (impl 3; DPFT_FPN_BUILD=1 selects fpn_output_tc_kernel for A/B): validated on B200
// (tests/test_features_gpu.py column-builder cases + every whole-model golden case), bench step 4.841 -> 4.802 ms
// (profiles/r02_ab_validated_paths.txt).  Same GEMM, same shared-memory operand image and
// the same epilogue as fpn_output_tc_kernel<CIN>; only the construction of the inner halo tile differs.  ncu on the
// 8x720x1280 camera level (gpurun_out/fpn_out_cam.raw.csv): 102.7 M warp instructions, issue slots 39 % active, two CTAs
// per SM (96 registers) -> 243 us for 324 MB, i.e. the kernel is bound by the builders' instruction stream (~230 per halo
// entry), not by HBM (50 us).  Here
//   * a thread owns one tile COLUMN and walks five of the tile's ten rows: no per-entry division, the column's validity,
//     coarse column and raw pointer are computed once;
//   * the top-down term is staged once per CTA: the (<= FB_CR x FB_CC) patch of the coarser inner map the tile touches,
//     with the lateral bias already added, goes to shared memory with coalesced loads; a thread re-reads its 16 values only
//     when the nearest-neighbour source row changes (every fourth row) and they START the FMA chain
//     (v = w2*x2 + (w1*x1 + (w0*x0 + (coarse + bias)))), so the 16 adds and 4 global loads per entry disappear;
//   * three CTAs per SM (launch bound), 58 KB of shared memory each.
constexpr int FB_CR = 5, FB_CC = 36;                 // rows / columns of the staged coarse patch (host checks the spans)
constexpr int FB_ROWS = (TC_TH + 2) / 2;             // rows per column-owning thread
static_assert(TC_THREADS == 2 * TC_TW && (TC_TH + 2) % 2 == 0, "two row groups of TC_TW column threads");

__device__ __forceinline__ int nearest_src_scaled(int dst, float scale, int in_size) {
    const int s = (int)floorf((float)dst * scale);   // same arithmetic as nearest_src (scale = float(in) / float(out))
    return s < in_size - 1 ? s : in_size - 1;
}

template <int CIN, int ROWS, bool LATC>
__device__ __forceinline__ void fpn_build_column(const FpnOutParams& prm, uint8_t* s_a, const float* s_lat, const float* s_cb,
                                                 int b, int p0, int q0, int px, int rr0, float scale_h, float scale_w,
                                                 int hc0, int wc0) {
    const int H = prm.H, W = prm.W;
    const int ww = q0 - 1 + px;
    const bool col_ok = ww >= 0 && ww < W;
    const int ws = col_ok ? ww : 0;
    const int wcl = nearest_src_scaled(ws, scale_w, prm.Wc) - wc0;
    float xin[ROWS][CIN];
    bool ok[ROWS];
#pragma unroll
    for (int u = 0; u < ROWS; ++u) {                 // every raw load of the column is issued before any is consumed
        const int hh = p0 - 1 + rr0 + u;
        ok[u] = col_ok && hh >= 0 && hh < H;
        const long long xi = (((long long)b * H + (ok[u] ? hh : 0)) * W + ws) * CIN;
#pragma unroll
        for (int c = 0; c < CIN; ++c) xin[u][c] = ok[u] ? raw_ld(prm, xi + c) : 0.0f;
    }
    float cb[FC];
#pragma unroll
    for (int o = 0; o < FC; ++o) cb[o] = 0.0f;
    int cur = -1;
#pragma unroll
    for (int u = 0; u < ROWS; ++u) {
        const int rr = rr0 + u;
        uint4 pk[2] = {make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u)};      // zero outside the image (conv padding)
        if (ok[u]) {
            const int hcl = nearest_src_scaled(p0 - 1 + rr, scale_h, prm.Hc) - hc0;
            if (hcl != cur) {                        // coarse + lateral bias of this (coarse row, coarse column)
                const float4* cp = reinterpret_cast<const float4*>(s_cb + (hcl * FB_CC + wcl) * FC);
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const float4 f = cp[c4];
                    cb[4 * c4] = f.x; cb[4 * c4 + 1] = f.y; cb[4 * c4 + 2] = f.z; cb[4 * c4 + 3] = f.w;
                }
                cur = hcl;
            }
            float v[FC];
#pragma unroll
            for (int o = 0; o < FC; ++o) {
                float a = cb[o];
#pragma unroll
                for (int c = 0; c < CIN; ++c) a = fmaf(LATC ? prm.lat_c[o * CIN + c] : s_lat[o * CIN + c], xin[u][c], a);
                v[o] = a;
            }
            uint32_t* pw = reinterpret_cast<uint32_t*>(pk);
#pragma unroll
            for (int t = 0; t < 8; ++t) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(pw[t]) : "f"(v[2 * t + 1]), "f"(v[2 * t]));
        }
        *reinterpret_cast<uint4*>(s_a + rr * TC_ROWB + px * 16) = pk[0];
        *reinterpret_cast<uint4*>(s_a + rr * TC_ROWB + TC_PLANE + px * 16) = pk[1];
    }
}

template <int CIN, bool LATC>
__global__ void __launch_bounds__(TC_THREADS, 3)
fpn_output_tc2_kernel(const __grid_constant__ FpnOutParams prm) {
    static_assert(CIN > 0, "the column builder is for the raw level");
    extern __shared__ __align__(128) uint8_t tsm[];
    uint8_t* s_a = tsm;
    uint8_t* s_b = tsm + TC_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_b + TC_B_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + TC_TH);
    float* s_lat = reinterpret_cast<float*>(tmem_slot + 4);          // [16][CIN]
    float* s_py = s_lat + FC * 6 + FC;                                // [TC_TH][16]: pos_y rows of this tile (+ bias)
    float* s_cb = s_py + TC_TH * FC;                                  // [FB_CR][FB_CC][16]: coarse patch + lateral bias
    const int b = blockIdx.z;
    const int p0 = blockIdx.y * TC_TH, q0 = blockIdx.x * TC_TW;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int H = prm.H, W = prm.W;

    if (warp == 0) {
        tc::tmem_alloc(tmem_slot, 128);
    } else if (tid == 32) {
        for (int r = 0; r < TC_TH; ++r) tc::mbar_init(&bars[r], 1);
        tc::fence_barrier_init();
    }
    for (int i = tid; i < TC_B_BYTES / 16; i += TC_THREADS) reinterpret_cast<uint4*>(s_b)[i] = __ldg(prm.w_packed + i);
    if (tid < TC_TH * FC) {
        const int r = tid / FC, c = tid - r * FC;
        s_py[tid] = (p0 + r < H ? __ldg(prm.pos_y + (long long)(p0 + r) * FC + c) : 0.0f) + __ldg(prm.bias + c);
    }
    for (int i = tid; i < FC * CIN; i += TC_THREADS) s_lat[i] = __ldg(prm.lat_w + i);
    // the patch of the coarser inner map under this tile (nearest-neighbour sources of rows p0-1 .. p0+TC_TH and columns
    // q0-1 .. q0+TC_TW, clamped to the image), lateral bias added
    const float scale_h = (float)prm.Hc / (float)H, scale_w = (float)prm.Wc / (float)W;
    const int hc0 = nearest_src_scaled(p0 > 0 ? p0 - 1 : 0, scale_h, prm.Hc);
    const int wc0 = nearest_src_scaled(q0 > 0 ? q0 - 1 : 0, scale_w, prm.Wc);
    for (int i = tid; i < FB_CR * FB_CC * 4; i += TC_THREADS) {
        const int cell = i >> 2, c4 = i & 3;
        const int cr = cell / FB_CC, cc = cell - cr * FB_CC;
        const int hc = hc0 + cr < prm.Hc ? hc0 + cr : prm.Hc - 1;
        const int wc = wc0 + cc < prm.Wc ? wc0 + cc : prm.Wc - 1;
        float4 f = __ldg(reinterpret_cast<const float4*>(prm.coarse + (((long long)b * prm.Hc + hc) * prm.Wc + wc) * FC) + c4);
        const float4 lb = __ldg(reinterpret_cast<const float4*>(prm.lat_b) + c4);
        f.x += lb.x; f.y += lb.y; f.z += lb.z; f.w += lb.w;
        reinterpret_cast<float4*>(s_cb)[i] = f;
    }
    __syncthreads();
    // inner halo tile, f16: thread (group g, column c) builds rows g*FB_ROWS .. of column c; the two right halo columns
    // (TC_TW, TC_TW + 1) are one more entry for the first 2 * (TC_TH + 2) threads.  Entries beyond TC_TW + 1 are never read
    // by the MMAs (tap ds reads entries ds .. ds + 127).
    fpn_build_column<CIN, FB_ROWS, LATC>(prm, s_a, s_lat, s_cb, b, p0, q0, tid & (TC_TW - 1), (tid >> 7) * FB_ROWS, scale_h, scale_w, hc0, wc0);
    if (tid < 2 * (TC_TH + 2))
        fpn_build_column<CIN, 1, LATC>(prm, s_a, s_lat, s_cb, b, p0, q0, TC_TW + (tid & 1), tid >> 1, scale_h, scale_w, hc0, wc0);
    tc::fence_proxy_async();                 // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (tid == 0) {
        constexpr uint32_t idesc = tc::umma_idesc_16bit(128, 16, true);
        const uint32_t a0 = tc::smem_u32(s_a), b0 = tc::smem_u32(s_b);
        for (int r = 0; r < TC_TH; ++r) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int dr = tap / 3, ds = tap % 3;
                const uint64_t adesc = tc::umma_desc_noswizzle(a0 + (r + dr) * TC_ROWB + ds * 16, TC_PLANE, 128);
                const uint64_t bdesc = tc::umma_desc_noswizzle(b0 + tap * 512, 256, 128);
                tc::umma_bf16(tmem_base + r * 16, adesc, bdesc, idesc, tap ? 1u : 0u);
            }
            tc::umma_commit(&bars[r]);
        }
    }
    __syncwarp();
    // epilogue: as fpn_output_tc_kernel
    const int lane_px = (warp & 3) * 32 + (tid & 31);
    const int q = q0 + lane_px;
    float4 px4[4];
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4)
        px4[c4] = q < W ? __ldg(reinterpret_cast<const float4*>(prm.pos_x + (long long)q * FC) + c4) : make_float4(0, 0, 0, 0);
    const int r_begin = (warp >> 2) * (TC_TH / 2);
    for (int r = r_begin; r < r_begin + TC_TH / 2; ++r) {
        tc::mbar_wait(&bars[r], 0);
        tc::tcgen05_fence_after();
        uint32_t v[16];
        tc::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(r * 16), v);
        tc::tmem_ld_wait();
        const int p = p0 + r;
        if (p < H && q < W) {
            const float4* py = reinterpret_cast<const float4*>(s_py + r * FC);     // pos_y + bias
            float outv[FC];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float4 c = py[c4];
                outv[4 * c4] = (__uint_as_float(v[4 * c4]) + c.x) + px4[c4].x;
                outv[4 * c4 + 1] = (__uint_as_float(v[4 * c4 + 1]) + c.y) + px4[c4].y;
                outv[4 * c4 + 2] = (__uint_as_float(v[4 * c4 + 2]) + c.z) + px4[c4].z;
                outv[4 * c4 + 3] = (__uint_as_float(v[4 * c4 + 3]) + c.w) + px4[c4].w;
            }
            store_pyramid_row(prm.pyramid, (long long)b * prm.S + prm.start + (long long)p * W + q, outv, prm.pyramid_f16 != 0);
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc::tcgen05_fence_after();
        tc::tmem_dealloc(tmem_base, 128);
    }
}

// Worst-case extent of the coarse patch under one tile: `n` consecutive fine positions map to at most floor((n-1)*in/out)+2
// consecutive coarse positions (nearest-neighbour indices are floor(dst * in/out), clamped); one more for the float
// rounding of dst * scale at inexact ratios.
inline bool fpn_column_builder_eligible(int H, int W, int Hc, int Wc, bool has_coarse) {
    if (!has_coarse || Hc <= 0 || Wc <= 0 || Hc > H || Wc > W) return false;
    const long long span_r = ((long long)(TC_TH + 1) * Hc) / H + 3, span_c = ((long long)(TC_TW + 1) * Wc) / W + 3;
    return span_r <= FB_CR && span_c <= FB_CC;
}


// --------------------------------------------------------------------------------------- stem on the tensor cores
// conv 7x7 / stride 2 as an implicit GEMM per output row: M = 128 output pixels, N = 64 channels, K = 7 filter rows x
// 8 taps (7 + one zero tap) x 8 channels (Cin + zero padding).  The stride-2 window is made contiguous by splitting
// every input row into an even-column and an odd-column plane of 16-byte (8 x f16) entries: for tap s the operand of
// output pixels q0..q0+127 is plane (s & 1) starting at entry (s >> 1), so taps (2j, 2j+1) form one K = 16 MMA whose
// two K halves are exactly one plane apart (LBO) — again the canonical un-swizzled K-major UMMA layout, built by the
// threads while they convert the raw fp32 input.  Operands are always f16 (0..255 inputs are exact); 4 output rows per
// CTA, 64 TMEM columns and one mbarrier per row.
constexpr int ST_TH = 4, ST_TW = 128;
constexpr int ST_ROWS = 2 * ST_TH + 5;               // input rows per tile
constexpr int ST_PLANE_ENTRIES = 136;                // >= 128 + 4
constexpr int ST_PLANE = ST_PLANE_ENTRIES * 16;
constexpr int ST_ROWB = 2 * ST_PLANE;
constexpr int ST_A_BYTES = ST_ROWS * ST_ROWB;        // 56,576
constexpr int ST_B_BYTES = 7 * 4 * 2048;             // (r, tap pair) x [2 K-halves][8 N-groups][8 rows][16 B] = 57,344

constexpr int ST_THREADS = 256;                      // all build the operands; warps 0-3 own the 128 TMEM lanes

// w[r][s][c][o] fp32 -> f16 UMMA image [(r*4 + s/2)][s&1][o/8][o%8][c (8, zero padded)] (zero 8th tap), once per model
template <int CIN>
__global__ void stem_pack_weights_kernel(const float* __restrict__ w, __half* __restrict__ packed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // one packed element
    if (i >= ST_B_BYTES / 2) return;
    const int c = i & 7, orow = (i >> 3) & 7, og = (i >> 6) & 7, kh = (i >> 9) & 1, rj = i >> 10;
    const int r = rj >> 2, s2 = (rj & 3) * 2 + kh, o = og * 8 + orow;
    float v = 0.0f;
    if (s2 < 7 && c < CIN) v = w[((r * 7 + s2) * CIN + c) * 64 + o];
    packed[i] = __float2half_rn(v);
}

// 4-channels-per-tap image of the streaming kernel (Cin <= 4): a 16-byte K chunk holds TWO taps of the same column parity,
// [tap sa: ch 0..3 | tap sb: ch 0..3] with (sa, sb) = (0,2), (1,3), (4,6), (5,-) for chunks 0..3 of a filter row, so a K = 16
// MMA covers four taps and a filter row takes two MMAs instead of four.  Layout [(r*2 + t)][khalf][o/8][o%8][8 elems].
constexpr int ST_B4_BYTES = 7 * 2 * 2048;            // 28,672
template <int CIN>
__global__ void stem_pack_weights4_kernel(const float* __restrict__ w, __half* __restrict__ packed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // one packed element
    if (i >= ST_B4_BYTES / 2) return;
    const int k = i & 7, orow = (i >> 3) & 7, og = (i >> 6) & 7, kh = (i >> 9) & 1, rt = i >> 10;
    const int r = rt >> 1, chunk = (rt & 1) * 2 + kh, o = og * 8 + orow;
    const int sa = chunk == 0 ? 0 : chunk == 1 ? 1 : chunk == 2 ? 4 : 5;
    const int tap = k < 4 ? sa : sa + 2, c = k & 3;
    float v = 0.0f;
    if (tap < 7 && c < CIN) v = w[((r * 7 + tap) * CIN + c) * 64 + o];
    packed[i] = __float2half_rn(v);
}

template <int CIN, typename OT>
__global__ void __launch_bounds__(ST_THREADS)
stem_tc_kernel(const float* __restrict__ x, const uint4* __restrict__ w_packed, const float* __restrict__ bias,
               OT* __restrict__ y, int H, int W, int P, int Q, float lo) {
    extern __shared__ __align__(128) uint8_t tsm[];
    uint8_t* s_a = tsm;
    uint8_t* s_b = tsm + ST_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_b + ST_B_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + ST_TH + 2);   // (+2 keeps s_bias 16-byte aligned)
    float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);         // [64]
    const int b = blockIdx.z;
    const int p0 = blockIdx.y * ST_TH, q0 = blockIdx.x * ST_TW;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (warp == 0) {
        tc::tmem_alloc(tmem_slot, ST_TH * 64);
    } else if (tid == 32) {
        for (int r = 0; r <= ST_TH; ++r) tc::mbar_init(&bars[r], 1);
        tc::fence_barrier_init();
        // weights: the pre-packed f16 UMMA image, one bulk async copy (57 KB) that lands while the threads build the input tile
        tc::mbar_expect_tx(&bars[ST_TH], ST_B_BYTES);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(s_b)),
                     "l"(w_packed), "r"((uint32_t)ST_B_BYTES), "r"(tc::smem_u32(&bars[ST_TH]))
                     : "memory");
    }
    if (tid < STEM_COUT) s_bias[tid] = __ldg(bias + tid);
    // input tile: rows 2*p0-3 .. 2*p0-3+ST_ROWS-1, columns 2*q0-3 .. (+2*ST_PLANE_ENTRIES-1), even/odd planes
    const int h_base = 2 * p0 - 3, w_base = 2 * q0 - 3;
    // (batches of ST_BATCH entries per thread: all global loads of a batch are issued before any is converted)
    constexpr int ST_ENTRIES = ST_ROWS * 2 * ST_PLANE_ENTRIES;
    constexpr int ST_BATCH = 7;
    for (int base = tid; base < ST_ENTRIES; base += ST_BATCH * ST_THREADS) {
        float v[ST_BATCH][CIN];
#pragma unroll
        for (int u = 0; u < ST_BATCH; ++u) {
            const int i = base + u * ST_THREADS;
            const int rr = i / (2 * ST_PLANE_ENTRIES);
            const int xl = i - rr * (2 * ST_PLANE_ENTRIES);      // local column
            const int hh = h_base + rr, ww = w_base + xl;
            const bool ok = i < ST_ENTRIES && hh >= 0 && hh < H && ww >= 0 && ww < W;
            const float* xp = x + (((long long)b * H + (ok ? hh : 0)) * W + (ok ? ww : 0)) * CIN;
#pragma unroll
            for (int c = 0; c < CIN; ++c) v[u][c] = ok ? __ldg(xp + c) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < ST_BATCH; ++u) {
            const int i = base + u * ST_THREADS;
            if (i < ST_ENTRIES) {
                const int rr = i / (2 * ST_PLANE_ENTRIES);
                const int xl = i - rr * (2 * ST_PLANE_ENTRIES);
                __align__(16) __half hv[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) hv[c] = __float2half_rn(c < CIN ? v[u][c < CIN ? c : 0] : 0.0f);
                *reinterpret_cast<uint4*>(s_a + rr * ST_ROWB + (xl & 1) * ST_PLANE + (xl >> 1) * 16) = *reinterpret_cast<uint4*>(hv);
            }
        }
    }
    tc::fence_proxy_async();
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (tid == 0) {
        tc::mbar_wait(&bars[ST_TH], 0);              // the weight image has landed (async proxy -> async proxy: no fence needed)
        constexpr uint32_t idesc = tc::umma_idesc_16bit(128, 64, true);
        const uint32_t a0 = tc::smem_u32(s_a), b0 = tc::smem_u32(s_b);
        for (int pr = 0; pr < ST_TH; ++pr) {
#pragma unroll 1
            for (int r = 0; r < 7; ++r) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint64_t adesc = tc::umma_desc_noswizzle(a0 + (2 * pr + r) * ST_ROWB + j * 16, ST_PLANE, 128);
                    const uint64_t bdesc = tc::umma_desc_noswizzle(b0 + (r * 4 + j) * 2048, 1024, 128);
                    tc::umma_bf16(tmem_base + pr * 64, adesc, bdesc, idesc, (r | j) ? 1u : 0u);
                }
            }
            tc::umma_commit(&bars[pr]);
        }
    }
    __syncwarp();
    // epilogue: all eight warps; warps w and w+4 share TMEM lane quadrant w%4 and take the first / second half of the rows
    const int q = q0 + (warp & 3) * 32 + (tid & 31);
    const int pr_begin = (warp >> 2) * (ST_TH / 2);
    for (int pr = pr_begin; pr < pr_begin + ST_TH / 2; ++pr) {
        tc::mbar_wait(&bars[pr], 0);
        tc::tcgen05_fence_after();
        const int p = p0 + pr;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t v[32];
            tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(pr * 64 + half * 32), v);
            tc::tmem_ld_wait();
            if (p < P && q < Q) {
                OT* o = y + (((long long)b * P + p) * Q + q) * STEM_COUT + half * 32;
#pragma unroll
                for (int o8 = 0; o8 < 4; ++o8) {
                    uint4 pk;
                    const float4 b0v = *reinterpret_cast<const float4*>(s_bias + half * 32 + o8 * 8);
                    const float4 b1v = *reinterpret_cast<const float4*>(s_bias + half * 32 + o8 * 8 + 4);
                    const float bb[8] = {b0v.x, b0v.y, b0v.z, b0v.w, b1v.x, b1v.y, b1v.z, b1v.w};
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const float a0f = fmaxf(__uint_as_float(v[8 * o8 + 2 * t]) + bb[2 * t], lo);      // lo = 0: ReLU
                        const float a1f = fmaxf(__uint_as_float(v[8 * o8 + 2 * t + 1]) + bb[2 * t + 1], lo);
                        if constexpr (std::is_same<OT, __half>::value)
                            reinterpret_cast<__half2*>(&pk)[t] = __floats2half2_rn(fminf(a0f, 65504.0f), fminf(a1f, 65504.0f));
                        else
                            reinterpret_cast<__nv_bfloat162*>(&pk)[t] = __floats2bfloat162_rn(a0f, a1f);
                    }
                    reinterpret_cast<uint4*>(o)[o8] = pk;
                }
            }
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc::tcgen05_fence_after();
        tc::tmem_dealloc(tmem_base, ST_TH * 64);
    }
}

// ---- the same stem as a ROW-STREAMING pipeline --------------------------------------------------------------------------
// stem_tc_kernel builds 13 input rows to produce 4 output rows (3.25 input rows per output row instead of 2) and runs its
// phases — weights, tile build, multiply, drain — one after the other in a short-lived CTA.  Here a CTA owns (image, 128-column
// strip, chunk of ~P / chunks output rows) and its warps specialise:
//   warps 4-7  builders: warp w converts input rows i = w, w + 4, ... (fp32 -> f16, even / odd column planes) into a ring of
//              12 rows; every input row of the chunk is built once
//   warp 8     the 57 KB weight image by one bulk async copy, then per output row 7 x 4 tcgen05.mma (M=128, N=64, K=16) on
//              ring rows 2j .. 2j+6 into one of four TMEM accumulators; the commit releases rows 2j and 2j+1
//   warps 0-3  drain (TMEM lane quadrant = 32 output columns): + bias, clamp, 16-bit pack, 128 contiguous bytes per pixel
// Two CTAs fit per SM (110 KB of shared memory, 256 TMEM columns each).
constexpr int SS_RING = 12;
constexpr int SS_ACCS = 4;
constexpr int SS_THREADS = 288;
constexpr int SS_SMEM = ST_B_BYTES + SS_RING * ST_ROWB + STEM_COUT * 4 + (2 * SS_RING + 1 + 2 * SS_ACCS) * 8 + 16;
// Cin == 3 uses the 4-channels-per-tap operands (stem_pack_weights4_kernel): an A entry is [pixel x: 4 ch | pixel x + 2: 4 ch],
// i.e. every pixel is written into two entries of its parity plane; the MMA count per output row drops from 28 to 14.

template <int CIN, typename OT, typename IT>
__global__ void __launch_bounds__(SS_THREADS)
stem_stream_kernel(const IT* __restrict__ x, const uint4* __restrict__ w_packed, const float* __restrict__ bias,
                   OT* __restrict__ y, int H, int W, int P, int Q, float lo, int strips, int chunks, int chunk_rows) {
    extern __shared__ __align__(128) uint8_t tsm[];
    uint8_t* s_b = tsm;
    uint8_t* s_a = tsm + ST_B_BYTES;
    float* s_bias = reinterpret_cast<float*>(s_a + SS_RING * ST_ROWB);
    uint64_t* row_full = reinterpret_cast<uint64_t*>(s_bias + STEM_COUT);
    uint64_t* row_empty = row_full + SS_RING;
    uint64_t* w_bar = row_empty + SS_RING;
    uint64_t* tmem_full = w_bar + 1;
    uint64_t* tmem_empty = tmem_full + SS_ACCS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + SS_ACCS);

    constexpr bool PACK4 = CIN <= 4;
    constexpr int W_BYTES = PACK4 ? ST_B4_BYTES : ST_B_BYTES;
    constexpr int MMAS_PER_ROW = PACK4 ? 2 : 4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int unit = blockIdx.x;
    const int chunk = unit % chunks;
    const int strip = (unit / chunks) % strips;
    const int b = unit / (chunks * strips);
    const int p_begin = chunk * chunk_rows, q0 = strip * ST_TW;
    const int n_out = min(chunk_rows, P - p_begin);          // output rows of this CTA (>= 1)
    const int n_in = 2 * n_out + 5;                            // input rows 2*p_begin - 3 .. 2*(p_begin + n_out - 1) + 3
    const int h_base = 2 * p_begin - 3, w_base = 2 * q0 - 3;

    if (warp == 8) {
        tc::tmem_alloc(tmem_slot, SS_ACCS * 64);
    } else if (tid == 0) {
        for (int i = 0; i < SS_RING; ++i) {
            tc::mbar_init(&row_full[i], 1);
            tc::mbar_init(&row_empty[i], 1);
        }
        for (int i = 0; i < SS_ACCS; ++i) {
            tc::mbar_init(&tmem_full[i], 1);
            tc::mbar_init(&tmem_empty[i], 4);
        }
        tc::mbar_init(w_bar, 1);
        tc::fence_barrier_init();
    }
    if (tid < STEM_COUT) s_bias[tid] = __ldg(bias + tid);
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 4 && warp < 8) {
        // ================================ builders: one input row per warp and turn ================================
        constexpr int ENTRIES = 2 * ST_PLANE_ENTRIES;        // 272 column entries per row (even + odd plane)
        constexpr int PER_LANE = (ENTRIES + 31) / 32;        // 9
        for (int i = warp - 4; i < n_in; i += 4) {
            const int slot = i % SS_RING;
            tc::mbar_wait(&row_empty[slot], ((i / SS_RING) & 1) ^ 1);
            const int hh = h_base + i;
            const bool row_ok = hh >= 0 && hh < H;
            const IT* xrow = x + ((long long)b * H + (row_ok ? hh : 0)) * W * CIN;     // IT = float, or uint8 frames (exact in f16)
            float v[PER_LANE][CIN];
#pragma unroll
            for (int u = 0; u < PER_LANE; ++u) {             // all global loads of the row first
                const int xl = lane + u * 32;
                const int ww = w_base + xl;
                const bool ok = row_ok && xl < ENTRIES && ww >= 0 && ww < W;
                const IT* xp = xrow + (long long)(ok ? ww : 0) * CIN;
#pragma unroll
                for (int c = 0; c < CIN; ++c) v[u][c] = ok ? (float)__ldg(xp + c) : 0.0f;
            }
            uint8_t* dst = s_a + slot * ST_ROWB;
#pragma unroll
            for (int u = 0; u < PER_LANE; ++u) {
                const int xl = lane + u * 32;
                if (xl < ENTRIES) {
                    if constexpr (PACK4) {
                        __align__(8) __half hv[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) hv[c] = __float2half_rn(c < CIN ? v[u][c < CIN ? c : 0] : 0.0f);
                        uint8_t* e = dst + (xl & 1) * ST_PLANE + (xl >> 1) * 16;
                        *reinterpret_cast<uint2*>(e) = *reinterpret_cast<uint2*>(hv);                 // first tap of entry xl >> 1
                        if (xl >= 2) *reinterpret_cast<uint2*>(e - 8) = *reinterpret_cast<uint2*>(hv); // second tap of the entry before
                    } else {
                        __align__(16) __half hv[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) hv[c] = __float2half_rn(c < CIN ? v[u][c < CIN ? c : 0] : 0.0f);
                        *reinterpret_cast<uint4*>(dst + (xl & 1) * ST_PLANE + (xl >> 1) * 16) = *reinterpret_cast<uint4*>(hv);
                    }
                }
            }
            tc::fence_proxy_async();                          // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&row_full[slot]);
        }
    } else if (warp == 8) {
        // ====================================== weights + MMA issuer ======================================
        if (lane == 0) {
            // the 4-channel image follows the 8-channel one in the packed buffer
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(w_packed) + (PACK4 ? ST_B_BYTES : 0);
            tc::mbar_expect_tx(w_bar, W_BYTES);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(s_b)),
                         "l"(wsrc), "r"((uint32_t)W_BYTES), "r"(tc::smem_u32(w_bar))
                         : "memory");
        }
        tc::mbar_wait(w_bar, 0);
        constexpr uint32_t idesc = tc::umma_idesc_16bit(128, 64, true);
        const uint32_t a0 = tc::smem_u32(s_a), b0 = tc::smem_u32(s_b);
        for (int j = 0; j < n_out; ++j) {
            const int acc = j % SS_ACCS;
            tc::mbar_wait(&tmem_empty[acc], ((j / SS_ACCS) & 1) ^ 1);
            for (int r = (j == 0 ? 0 : 5); r < 7; ++r) {     // rows 2j .. 2j+4 were awaited by the previous output row
                const int i = 2 * j + r;
                tc::mbar_wait(&row_full[i % SS_RING], (i / SS_RING) & 1);
            }
            tc::tcgen05_fence_after();
            if (tc::elect_one()) {
#pragma unroll 1
                for (int r = 0; r < 7; ++r) {
                    const uint32_t row = a0 + ((2 * j + r) % SS_RING) * ST_ROWB;
#pragma unroll
                    for (int t = 0; t < MMAS_PER_ROW; ++t) {
                        // 8-channel entries: taps (2t, 2t+1) start t entries in; 4-channel entries: taps (0,2 | 1,3) start at
                        // entry 0 and taps (4,6 | 5,-) two entries in
                        const uint64_t adesc = tc::umma_desc_noswizzle(row + (PACK4 ? 2 * t : t) * 16, ST_PLANE, 128);
                        const uint64_t bdesc = tc::umma_desc_noswizzle(b0 + (r * MMAS_PER_ROW + t) * 2048, 1024, 128);
                        tc::umma_bf16(tmem_base + acc * 64, adesc, bdesc, idesc, (r | t) ? 1u : 0u);
                    }
                }
                tc::umma_commit(&tmem_full[acc]);
                tc::umma_commit(&row_empty[(2 * j) % SS_RING]);       // input rows 2j and 2j+1 are not read again
                tc::umma_commit(&row_empty[(2 * j + 1) % SS_RING]);
            }
            __syncwarp();
        }
    } else {
        // =========================================== drain ===========================================
        const int q = q0 + warp * 32 + lane;
        for (int j = 0; j < n_out; ++j) {
            const int acc = j % SS_ACCS;
            const int p = p_begin + j;
            tc::mbar_wait(&tmem_full[acc], (j / SS_ACCS) & 1);
            tc::tcgen05_fence_after();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t v[32];
                tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * 64 + half * 32), v);
                tc::tmem_ld_wait();
                if (q < Q) {
                    OT* o = y + (((long long)b * P + p) * Q + q) * STEM_COUT + half * 32;
#pragma unroll
                    for (int o8 = 0; o8 < 4; ++o8) {
                        uint4 pk;
                        const float4 b0v = *reinterpret_cast<const float4*>(s_bias + half * 32 + o8 * 8);
                        const float4 b1v = *reinterpret_cast<const float4*>(s_bias + half * 32 + o8 * 8 + 4);
                        const float bb[8] = {b0v.x, b0v.y, b0v.z, b0v.w, b1v.x, b1v.y, b1v.z, b1v.w};
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const float a0f = fmaxf(__uint_as_float(v[8 * o8 + 2 * t]) + bb[2 * t], lo);      // lo = 0: ReLU
                            const float a1f = fmaxf(__uint_as_float(v[8 * o8 + 2 * t + 1]) + bb[2 * t + 1], lo);
                            if constexpr (std::is_same<OT, __half>::value)
                                reinterpret_cast<__half2*>(&pk)[t] = __floats2half2_rn(fminf(a0f, 65504.0f), fminf(a1f, 65504.0f));
                            else
                                reinterpret_cast<__nv_bfloat162*>(&pk)[t] = __floats2bfloat162_rn(a0f, a1f);
                        }
                        reinterpret_cast<uint4*>(o)[o8] = pk;
                    }
                }
            }
            tc::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tmem_empty[acc]);
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc::tcgen05_fence_after();
        tc::tmem_dealloc(tmem_base, SS_ACCS * 64);
    }
}

template <int CIN, typename OT, typename IT>
static int launch_stem_stream(const IT* x, const void* w_packed, const float* bias, void* y, int B, int H, int W, int P, int Q,
                              float lo, cudaStream_t s) {
    auto kern = stem_stream_kernel<CIN, OT, IT>;
    static bool configured = false;
    if (!configured) {
        int st = cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SS_SMEM), "stem stream attr");
        if (st) return st;
        configured = true;
    }
    // work units = (image, 128-column strip, row chunk): about two per SM (two CTAs are resident per SM)
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int strips = (Q + ST_TW - 1) / ST_TW;
    int chunks = (2 * sms + B * strips / 2) / (B * strips);
    if (chunks < 1) chunks = 1;
    if (chunks > P) chunks = P;
    const int chunk_rows = (P + chunks - 1) / chunks;
    chunks = (P + chunk_rows - 1) / chunk_rows;
    kern<<<B * strips * chunks, SS_THREADS, SS_SMEM, s>>>(x, (const uint4*)w_packed, bias, (OT*)y, H, W, P, Q, lo, strips, chunks, chunk_rows);
    return 0;
}

template <int CIN, typename OT>
static int launch_stem_tc(const float* x, const void* w_packed, const float* bias, void* y, int B, int H, int W, int P, int Q,
                          float lo, cudaStream_t s) {
    static const int stream_mode = [] { const char* e = getenv("DPFT_STEM_STREAM"); return (e && e[0] == '0') ? 0 : 1; }();
    if (stream_mode) return launch_stem_stream<CIN, OT, float>(x, w_packed, bias, y, B, H, W, P, Q, lo, s);
    auto kern = stem_tc_kernel<CIN, OT>;
    const size_t smem = ST_A_BYTES + ST_B_BYTES + (ST_TH + 2) * 8 + 16 + STEM_COUT * sizeof(float);
    static bool configured = false;
    if (!configured) {
        int st = cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "stem tc attr");
        if (st) return st;
        configured = true;
    }
    const dim3 grid((Q + ST_TW - 1) / ST_TW, (P + ST_TH - 1) / ST_TH, B);
    kern<<<grid, ST_THREADS, smem, s>>>(x, (const uint4*)w_packed, bias, (OT*)y, H, W, P, Q, lo);
    return 0;
}

}  // namespace
}  // namespace dpft

using namespace dpft;

template <int CIN, typename OT>
static int launch_stem(const float* x, const float* w, const float* bias, void* y, int H, int W, int P, int Q, float lo, dim3 grid,
                       size_t smem, cudaStream_t s) {
    auto kern = stem_conv7x7_kernel<CIN, OT>;
    static bool configured = false;
    if (!configured) {
        int st = cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "stem attr");
        if (st) return st;
        configured = true;
    }
    kern<<<grid, 256, smem, s>>>(x, w, bias, (OT*)y, H, W, P, Q, lo);
    return 0;
}

extern "C" int dpft_stem_pack_weights(const float* w, void* packed, int Cin, void* stream) {
    DPFT_REQUIRE(w && packed, "stem_pack_weights: null pointer");
    DPFT_REQUIRE(Cin == 3 || Cin == 6, "stem_pack_weights: Cin=%d (3 or 6 supported)", Cin);
    const int n = ST_B_BYTES / 2;
    if (Cin == 3) {
        stem_pack_weights_kernel<3><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, (__half*)packed);
        stem_pack_weights4_kernel<3><<<(ST_B4_BYTES / 2 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
            w, reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(packed) + ST_B_BYTES));
    } else {
        stem_pack_weights_kernel<6><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, (__half*)packed);
    }
    DPFT_LAUNCH_CHECK("stem_pack_weights_kernel");
    return DPFT_OK;
}

extern "C" int dpft_stem_conv7x7_forward(const float* x, const float* w, const void* w_packed, const float* bias, void* y,
                                         int B, int H, int W, int Cin, int dtype, int impl, void* stream) {
    return dpft_stem_conv7x7_forward_ex(x, w, w_packed, bias, y, B, H, W, Cin, dtype, impl, 1, stream);
}

extern "C" int dpft_stem_conv7x7_forward_ex(const float* x, const float* w, const void* w_packed, const float* bias, void* y,
                                            int B, int H, int W, int Cin, int dtype, int impl, int relu, void* stream) {
    const float lo = relu ? 0.0f : (dtype == DPFT_F16 ? -65504.0f : -3.0e38f);
    const bool x_u8 = (Cin & DPFT_RAW_U8) != 0;                    // x is (B, H, W, Cin) uint8
    Cin &= ~DPFT_RAW_U8;
    if (x_u8) {
        // uint8 frames go through the row-streaming tensor-core kernel only (the camera stem); other shapes convert first
        DPFT_REQUIRE(Cin == 3 && w_packed && impl != 1 && (W - 1) / 2 + 1 >= 64 && x && bias && y && B > 0 && H > 0,
                     "stem: uint8 input needs Cin = 3, the packed weights and an output width >= 64");
        DPFT_REQUIRE(dtype == DPFT_BF16 || dtype == DPFT_F16, "stem: output dtype must be DPFT_BF16 or DPFT_F16");
        const int P8 = (H - 1) / 2 + 1, Q8 = (W - 1) / 2 + 1;
        const unsigned char* x8 = reinterpret_cast<const unsigned char*>(x);
        const int st8 = dtype == DPFT_F16
            ? launch_stem_stream<3, __half, unsigned char>(x8, w_packed, bias, y, B, H, W, P8, Q8, lo, (cudaStream_t)stream)
            : launch_stem_stream<3, __nv_bfloat16, unsigned char>(x8, w_packed, bias, y, B, H, W, P8, Q8, lo, (cudaStream_t)stream);
        if (st8) return st8;
        DPFT_LAUNCH_CHECK("stem_stream_kernel<uint8>");
        return DPFT_OK;
    }
    DPFT_REQUIRE(impl != 2 || w_packed, "stem: the tensor-core kernel needs the packed weights (dpft_stem_pack_weights)");
    DPFT_REQUIRE(impl >= 0 && impl <= 2, "stem: impl must be 0 (auto), 1 (CUDA cores) or 2 (tensor cores)");
    DPFT_REQUIRE(x && w && bias && y, "stem: null pointer");
    DPFT_REQUIRE(dtype == DPFT_BF16 || dtype == DPFT_F16, "stem: output dtype must be DPFT_BF16 or DPFT_F16");
    DPFT_REQUIRE(Cin == 3 || Cin == 6, "stem: Cin=%d (3 or 6 supported)", Cin);
    DPFT_REQUIRE(B > 0 && H > 0 && W > 0, "stem: bad size");
    const int P = (H - 1) / 2 + 1, Q = (W - 1) / 2 + 1;
    const dim3 grid((Q + STEM_TW - 1) / STEM_TW, (P + STEM_TH - 1) / STEM_TH, B);
    const size_t smem = sizeof(float) * (49 * Cin * STEM_COUT + STEM_PH * STEM_PW * Cin);
    cudaStream_t s = (cudaStream_t)stream;
    int st;
    if (impl == 2 || (impl == 0 && Q >= 64 && w_packed)) {
        if (Cin == 3) st = dtype == DPFT_F16 ? launch_stem_tc<3, __half>(x, w_packed, bias, y, B, H, W, P, Q, lo, s)
                                             : launch_stem_tc<3, __nv_bfloat16>(x, w_packed, bias, y, B, H, W, P, Q, lo, s);
        else st = dtype == DPFT_F16 ? launch_stem_tc<6, __half>(x, w_packed, bias, y, B, H, W, P, Q, lo, s)
                                    : launch_stem_tc<6, __nv_bfloat16>(x, w_packed, bias, y, B, H, W, P, Q, lo, s);
        if (st) return st;
        DPFT_LAUNCH_CHECK("stem_tc_kernel");
        return DPFT_OK;
    }
    if (Cin == 3) st = dtype == DPFT_F16 ? launch_stem<3, __half>(x, w, bias, y, H, W, P, Q, lo, grid, smem, s)
                                         : launch_stem<3, __nv_bfloat16>(x, w, bias, y, H, W, P, Q, lo, grid, smem, s);
    else st = dtype == DPFT_F16 ? launch_stem<6, __half>(x, w, bias, y, H, W, P, Q, lo, grid, smem, s)
                                : launch_stem<6, __nv_bfloat16>(x, w, bias, y, H, W, P, Q, lo, grid, smem, s);
    if (st) return st;
    DPFT_LAUNCH_CHECK("stem_conv7x7_kernel");
    return DPFT_OK;
}

extern "C" int dpft_maxpool3x3s2_nhwc(const void* x, void* y, int B, int H, int W, int C, int dtype, void* stream) {
    DPFT_REQUIRE(dtype == DPFT_BF16 || dtype == DPFT_F16, "maxpool: dtype must be DPFT_BF16 or DPFT_F16");
    DPFT_REQUIRE(x && y, "maxpool: null pointer");
    DPFT_REQUIRE(C % 8 == 0 && B > 0 && H > 0 && W > 0, "maxpool: C=%d must be a multiple of 8", C);
    const int P = (H - 1) / 2 + 1, Q = (W - 1) / 2 + 1;
    const long long total = (long long)B * P * Q * (C / 8);
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (dtype == DPFT_F16)
        maxpool3x3s2_kernel<__half2><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)x, (uint16_t*)y, B, H, W, C / 8, P, Q);
    else
        maxpool3x3s2_kernel<__nv_bfloat162><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)x, (uint16_t*)y, B, H, W, C / 8, P, Q);
    DPFT_LAUNCH_CHECK("maxpool3x3s2_kernel");
    return DPFT_OK;
}

extern "C" int dpft_fpn_pack_weights(const float* w, void* packed, void* stream) {
    DPFT_REQUIRE(w && packed, "fpn_pack_weights: null pointer");
    fpn_pack_weights_kernel<<<9, 256, 0, (cudaStream_t)stream>>>(w, (__half*)packed);
    DPFT_LAUNCH_CHECK("fpn_pack_weights_kernel");
    return DPFT_OK;
}

extern "C" int dpft_fpn_output_forward(const float* inner, const float* raw, int raw_channels, const float* lat_w,
                                       const float* lat_b, const float* coarse, int Hc, int Wc, const float* w,
                                       const void* w_packed, const float* bias, const float* pos_y, const float* pos_x, void* pyramid,
                                       int pyramid_dtype, long long S, long long start, int B, int H, int W, int impl,
                                       void* stream) {
    return dpft_fpn_output_forward_ex(inner, raw, raw_channels, lat_w, nullptr, lat_b, coarse, Hc, Wc, w, w_packed, bias, pos_y, pos_x,
                                      pyramid, pyramid_dtype, S, start, B, H, W, impl, stream);
}

extern "C" int dpft_fpn_output_forward_ex(const float* inner, const float* raw, int raw_channels, const float* lat_w,
                                          const float* lat_w_host, const float* lat_b, const float* coarse, int Hc, int Wc,
                                          const float* w, const void* w_packed, const float* bias, const float* pos_y,
                                          const float* pos_x, void* pyramid, int pyramid_dtype, long long S, long long start, int B,
                                          int H, int W, int impl, void* stream) {
    DPFT_REQUIRE(pyramid_dtype == DPFT_F32 || pyramid_dtype == DPFT_F16, "fpn_output: pyramid dtype must be DPFT_F32 or DPFT_F16");
    DPFT_REQUIRE(w && bias && pos_y && pos_x && pyramid, "fpn_output: null pointer");
    DPFT_REQUIRE((inner != nullptr) != (raw != nullptr), "fpn_output: exactly one of inner / raw must be given");
    DPFT_REQUIRE(B > 0 && H > 0 && W > 0, "fpn_output: bad size");
    const int raw_u8 = (raw_channels & DPFT_RAW_U8) ? 1 : 0;       // the raw level handed over as uint8
    raw_channels &= ~DPFT_RAW_U8;
    FpnOutParams prm{inner, raw, lat_w, lat_b, coarse, w, (const uint4*)w_packed, bias, pos_y, pos_x, pyramid, pyramid_dtype == DPFT_F16 ? 1 : 0,
                     S, start, H, W, Hc, Wc, raw_u8};
    cudaStream_t s = (cudaStream_t)stream;
    DPFT_REQUIRE(impl >= 0 && impl <= 3, "fpn_output: impl must be 0 (auto), 1 (CUDA cores), 2 (tensor cores) or 3 (tensor cores, "
                 "column-owning tile builder)");
    if (!inner) {
        DPFT_REQUIRE(lat_w && lat_b, "fpn_output: lateral weights needed for the raw level");
        DPFT_REQUIRE(coarse == nullptr || (Hc > 0 && Wc > 0), "fpn_output: bad coarse size");
        DPFT_REQUIRE(raw_channels == 3 || raw_channels == 6, "fpn_output: raw_channels=%d (3 or 6 supported)", raw_channels);
    }
    DPFT_REQUIRE(impl < 2 || w_packed, "fpn_output: the tensor-core kernel needs the packed weights (dpft_fpn_pack_weights)");
    // raw-level tile builder: the column-owning fpn_output_tc2_kernel where eligible (impl 3 forces it, DPFT_FPN_BUILD=1 in the
    // environment keeps the automatic choice on fpn_output_tc_kernel for A/B timing)
    static const int build_mode = [] { const char* e = getenv("DPFT_FPN_BUILD"); return e ? atoi(e) : 2; }();
    const bool column_builder = !inner && fpn_column_builder_eligible(H, W, Hc, Wc, coarse != nullptr) && ((uintptr_t)lat_b & 15) == 0 &&
                                (impl == 3 || (impl == 0 && build_mode == 2 && W >= 96 && w_packed));
    DPFT_REQUIRE(impl != 3 || column_builder, "fpn_output: impl 3 needs the raw level with a coarser map of <= 1/%d x the size",
                 (TC_TH + 1) / (FB_CR - 3));
    if (column_builder) {
        const dim3 tgrid((W + TC_TW - 1) / TC_TW, (H + TC_TH - 1) / TC_TH, B);
        const size_t smem = TC_A_BYTES + TC_B_BYTES + TC_TH * 8 + 16 + sizeof(float) * (FC * 6 + FC + TC_TH * FC + FB_CR * FB_CC * FC);
        static bool configured2 = false;
        if (!configured2) {
            int st = cuda_status(cudaFuncSetAttribute(fpn_output_tc2_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "fpn tc2 attr");
            if (!st) st = cuda_status(cudaFuncSetAttribute(fpn_output_tc2_kernel<6, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "fpn tc2 attr");
            if (!st) st = cuda_status(cudaFuncSetAttribute(fpn_output_tc2_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "fpn tc2 attr");
            if (!st) st = cuda_status(cudaFuncSetAttribute(fpn_output_tc2_kernel<6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "fpn tc2 attr");
            if (st) return st;
            configured2 = true;
        }
        if (lat_w_host) {                       // lateral weights as kernel parameters (constant operands of the builder's FMAs)
            for (int i = 0; i < FC * raw_channels; ++i) prm.lat_c[i] = lat_w_host[i];
            if (raw_channels == 3) fpn_output_tc2_kernel<3, true><<<tgrid, TC_THREADS, smem, s>>>(prm);
            else fpn_output_tc2_kernel<6, true><<<tgrid, TC_THREADS, smem, s>>>(prm);
        } else if (raw_channels == 3) fpn_output_tc2_kernel<3, false><<<tgrid, TC_THREADS, smem, s>>>(prm);
        else fpn_output_tc2_kernel<6, false><<<tgrid, TC_THREADS, smem, s>>>(prm);
        DPFT_LAUNCH_CHECK("fpn_output_tc2_kernel");
        return DPFT_OK;
    }
    if (impl == 2 || (impl == 0 && W >= 96 && w_packed)) {          // wide levels: 128-pixel row strips on the tensor cores
        const dim3 tgrid((W + TC_TW - 1) / TC_TW, (H + TC_TH - 1) / TC_TH, B);
        const size_t smem = TC_A_BYTES + TC_B_BYTES + TC_TH * 8 + 16 + sizeof(float) * (FC * 6 + FC + TC_TH * FC);
        static bool configured = false;
        if (!configured) {
            int st = cuda_status(cudaFuncSetAttribute(fpn_output_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "fpn tc attr");
            if (!st) st = cuda_status(cudaFuncSetAttribute(fpn_output_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "fpn tc attr");
            if (!st) st = cuda_status(cudaFuncSetAttribute(fpn_output_tc_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "fpn tc attr");
            if (st) return st;
            configured = true;
        }
        if (inner) fpn_output_tc_kernel<0><<<tgrid, TC_THREADS, smem, s>>>(prm);
        else if (raw_channels == 3) fpn_output_tc_kernel<3><<<tgrid, TC_THREADS, smem, s>>>(prm);
        else fpn_output_tc_kernel<6><<<tgrid, TC_THREADS, smem, s>>>(prm);
        DPFT_LAUNCH_CHECK("fpn_output_tc_kernel");
        return DPFT_OK;
    }
    const dim3 grid((W + FPN_TW - 1) / FPN_TW, (H + FPN_TH - 1) / FPN_TH, B);
    if (inner) {
        fpn_output_kernel<0><<<grid, 256, 0, s>>>(prm);
    } else {
        if (raw_channels == 3) fpn_output_kernel<3><<<grid, 256, 0, s>>>(prm);
        else if (raw_channels == 6) fpn_output_kernel<6><<<grid, 256, 0, s>>>(prm);
        else { set_error("fpn_output: raw_channels=%d (3 or 6 supported)", raw_channels); return DPFT_ERR_UNSUPPORTED; }
    }
    DPFT_LAUNCH_CHECK("fpn_output_kernel");
    return DPFT_OK;
}
