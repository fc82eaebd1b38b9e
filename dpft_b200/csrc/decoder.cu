// Fused query decoder of the DPRT fuser for sm_100a (inference): one launch per iteration runs, for every
// (sample, view, query tile), the whole MLFusion layer of the reference (src/dprt/models/fusers/mpfusion.py:231-263):
//
//   self-attention (nn.MultiheadAttention, q = k = x + pos, v = x; :122-148) -> + x -> LayerNorm
//   reference-point projection of the current box centres into the view (:617-696, cart2spher
//     src/dprt/models/utils/transformations.py:71-120)
//   multi-scale deformable cross-attention (src/dprt/models/layers/ms_deform_attn.py:138-217):
//     sampling offsets + softmax attention weights from (x + pos), bilinear gather over the L levels,
//     head-weighted reduce, value projection, output projection -> + x -> LayerNorm
//   feed-forward (Linear -> activation -> Linear; mpfusion.py:210-229) -> + x -> LayerNorm
//
// and a second launch does the view reduction (Linear(V*C -> C), :438,:512) and the detection head MLPs
// (src/dprt/models/heads/detection.py:252-275) including the additive centre refinement (:273).
//
// Layout / mapping.  C = d_model = 16, M = 8 heads, D = 2 (the configuration every shipped DPFT config uses).
// Thread (q, j) owns channels (2j, 2j+1) of query q — which is exactly head j — so the attention head state, the
// layer-norm statistics (8-lane xor shuffles) and the 16-vector exchanges (one padded smem row per query) all
// stay inside 8 consecutive lanes.  The keys/values of all N queries of the sample live in shared memory
// (N*32 floats), the layer's weights too (one flat pre-packed image, bank-conflict padded on the host).
//
// The deformable gather is done "gather-then-project": value_proj is linear, so
//   sum_p a_p * (W_h x_p + b_h) = W_h (sum_p a_p x_p) + b_h * sum_p a_p * inb_p
// (inb_p = bilinear weight mass of the in-bounds corners; zero padding zeroes the bias too).  The kernel gathers
// the 16-channel feature row (64 B, two sectors) of each corner straight from the FPN pyramid and projects after
// the reduce; the dense value_proj / torch.cat passes over all S positions that the reference makes in every
// layer (ms_deform_attn.py:172, mpfusion.py:179) disappear.
#include <stdlib.h>

#include "common.cuh"

namespace dpft {
namespace {

constexpr int C = 16;          // d_model
constexpr int NH = 8;          // heads
constexpr int TQ = 32;         // queries per CTA
constexpr int kThreads = TQ * NH;
constexpr int kMaxLevels = 8;
constexpr int kMaxViews = 4;
constexpr int ROW = C + 1;     // padded smem row of a per-query 16-vector

struct ViewDesc {
    const void* pyramid;       // (B, S, 16) fp32 or f16, positional embedding already added
    const float* weights;      // packed layer image (see pack_layer in dpft_b200/decoder.py)
    const float* transform;    // (B, 4, 4)
    const float* projection;   // (B, 4, 4) (3x4 inputs are padded with the row [0 0 0 1])
    const float* shape_hw;     // (B, 2) original input (H, W) as float
    const int* use_transform;  // 1 element: transformation.any()
    long long S;
    int h[kMaxLevels], w[kMaxLevels];
    long long start[kMaxLevels];
};

struct LayerParams {
    ViewDesc view[kMaxViews];
    const float* query;        // (B, N, 16) or (N, 16) when query_batch_stride == 0
    const float* pos;          // (N, 16)
    const float* center;       // (B, N, 3) or (N, 3) when center_batch_stride == 0
    float* out;                // (B, V, N, 16)
    long long query_batch_stride, center_batch_stride;
    int B, V, N, d_ffn, act;   // act: 0 = ReLU, 1 = Mish, 2 = GELU(erf)
    int weight_floats;         // size of one packed layer image
};

// Offsets (in floats) into the packed layer image.  Must match dpft_b200/decoder.py::layer_layout.
template <int L, int P> struct LayerImage {
    static constexpr int LP = L * P;
    static constexpr int in_w = 0;                        // [48][16]
    static constexpr int in_b = in_w + 48 * C;            // [48]
    static constexpr int out_w = in_b + 48;               // [16][16]
    static constexpr int out_b = out_w + C * C;
    static constexpr int ln1_w = out_b + C, ln1_b = ln1_w + C;
    static constexpr int off_stride = LP * 2 * C + 4;     // per head block, +4 floats so heads hit distinct banks
    static constexpr int off_w = ln1_b + C;               // [8][LP*2][16] (+pad)
    static constexpr int off_b = off_w + NH * off_stride; // [8*LP*2]
    static constexpr int att_stride = LP * C + 4;
    static constexpr int att_w = off_b + NH * LP * 2;     // [8][LP][16] (+pad)
    static constexpr int att_b = att_w + NH * att_stride; // [8*LP]
    static constexpr int val_w = att_b + NH * LP;         // [16][16]
    static constexpr int val_b = val_w + C * C;
    static constexpr int prj_w = val_b + C;               // output_proj [16][16]
    static constexpr int prj_b = prj_w + C * C;
    static constexpr int ln2_w = prj_b + C, ln2_b = ln2_w + C;
    static constexpr int ln3_w = ln2_b + C, ln3_b = ln3_w + C;
    static constexpr int ffn = ln3_b + C;                 // ffn1_w [F][16], ffn1_b [F], ffn2_w [16][F], ffn2_b [16]
    static constexpr int fixed = ffn;
};

__device__ __forceinline__ float group8_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

__device__ __forceinline__ float activation(float x, int act) {
    if (act == 1) {   // Mish: x * tanh(softplus(x)), softplus with torch's threshold of 20
        const float sp = x > 20.0f ? x : log1pf(expf(x));
        return x * tanhf(sp);
    }
    if (act == 2) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
    return fmaxf(x, 0.0f);
}

// y = LayerNorm over the 16 channels held 2 per lane across 8 lanes (eps 1e-5, biased variance).
__device__ __forceinline__ void layer_norm2(float& a, float& b, const float* w, const float* bias, int j) {
    const float mean = group8_sum(a + b) * (1.0f / C);
    const float da = a - mean, db = b - mean;
    const float var = group8_sum(da * da + db * db) * (1.0f / C);
    const float inv = rsqrtf(var + 1e-5f);
    a = da * inv * w[2 * j] + bias[2 * j];
    b = db * inv * w[2 * j + 1] + bias[2 * j + 1];
}

// dot of a 16-float weight row (16-byte aligned, smem) with a 16-vector in registers
__device__ __forceinline__ float dot16(const float* wrow, const float* x) {
    const float4* w4 = reinterpret_cast<const float4*>(wrow);
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 w = w4[i];
        acc = fmaf(w.x, x[4 * i], acc);
        acc = fmaf(w.y, x[4 * i + 1], acc);
        acc = fmaf(w.z, x[4 * i + 2], acc);
        acc = fmaf(w.w, x[4 * i + 3], acc);
    }
    return acc;
}

// 16-channel pyramid row of one corner, as fp32 (4 x 16 B) or f16 (2 x 16 B)
template <typename PT> struct RowRegs;
template <> struct RowRegs<float> {
    float4 q[4];
    __device__ __forceinline__ void load(const void* base, long long pixel) {
        const float4* r = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + pixel * C);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) q[c4] = __ldg(r + c4);
    }
    __device__ __forceinline__ void accumulate(float w, float* agg) const {
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
            agg[4 * c4] = fmaf(w, q[c4].x, agg[4 * c4]);
            agg[4 * c4 + 1] = fmaf(w, q[c4].y, agg[4 * c4 + 1]);
            agg[4 * c4 + 2] = fmaf(w, q[c4].z, agg[4 * c4 + 2]);
            agg[4 * c4 + 3] = fmaf(w, q[c4].w, agg[4 * c4 + 3]);
        }
    }
};
template <> struct RowRegs<__half> {
    uint4 q[2];
    __device__ __forceinline__ void load(const void* base, long long pixel) {
        const uint4* r = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(base) + pixel * C);
        q[0] = __ldg(r);
        q[1] = __ldg(r + 1);
    }
    __device__ __forceinline__ void accumulate(float w, float* agg) const {
        const __half2* h = reinterpret_cast<const __half2*>(q);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const float2 f = __half22float2(h[t]);
            agg[2 * t] = fmaf(w, f.x, agg[2 * t]);
            agg[2 * t + 1] = fmaf(w, f.y, agg[2 * t + 1]);
        }
    }
};

// bilinear footprint of one sample on an (H, W) level: clamped corner pixels + weights (already times the attention weight)
struct Corners {
    long long px[4];
    float wt[4];
};
__device__ __forceinline__ Corners corners_of(float lx, float ly, float a, int H, int W) {
    Corners c;
    const float w_im = lx * (float)W - 0.5f;
    const float h_im = ly * (float)H - 0.5f;
    // Unconditional loads from clamped corner addresses (weight 0 where the corner or the whole sample is out of bounds):
    // the row requests of a sample issue back to back instead of behind four branches.
    const bool inside = h_im > -1.0f && w_im > -1.0f && h_im < (float)H && w_im < (float)W;
    const float hs = inside ? h_im : 0.0f, ws = inside ? w_im : 0.0f;
    const float hf = floorf(hs), wf = floorf(ws);
    const int h0 = (int)hf, w0 = (int)wf;
    const float lh = hs - hf, lw = ws - wf, hh = 1.0f - lh, hw = 1.0f - lw;
    const bool t_ok = inside && h0 >= 0, b_ok = inside && h0 + 1 <= H - 1;
    const bool l_ok = w0 >= 0, r_ok = w0 + 1 <= W - 1;
    c.wt[0] = t_ok && l_ok ? hh * hw * a : 0.0f;
    c.wt[1] = t_ok && r_ok ? hh * lw * a : 0.0f;
    c.wt[2] = b_ok && l_ok ? lh * hw * a : 0.0f;
    c.wt[3] = b_ok && r_ok ? lh * lw * a : 0.0f;
    const int h0c = min(max(h0, 0), H - 1), h1c = min(max(h0 + 1, 0), H - 1);
    const int w0c = min(max(w0, 0), W - 1), w1c = min(max(w0 + 1, 0), W - 1);
    c.px[0] = (long long)h0c * W + w0c; c.px[1] = (long long)h0c * W + w1c;
    c.px[2] = (long long)h1c * W + w0c; c.px[3] = (long long)h1c * W + w1c;
    return c;
}

template <int L, int P, typename PT>
__global__ void __launch_bounds__(kThreads)
decoder_layer_kernel(const LayerParams prm) {
    using IMG = LayerImage<L, P>;
    constexpr int LP = L * P;
    extern __shared__ __align__(16) float smem[];
    const int N = prm.N;
    float* s_w = smem;                                   // packed weights
    float* s_k = s_w + ((prm.weight_floats + 3) & ~3);    // [N][16]
    float* s_v = s_k + (size_t)N * C;                     // [N][16]
    float* s_x = s_v + (size_t)N * C;                     // [TQ][ROW] exchange rows
    float* s_h = s_x + TQ * ROW;                          // [TQ][d_ffn + 1] hidden rows

    const int tiles = (N + TQ - 1) / TQ;
    const int tile = blockIdx.x % tiles;
    const int v = (blockIdx.x / tiles) % prm.V;
    const int b = blockIdx.x / (tiles * prm.V);
    const ViewDesc& vd = prm.view[v];
    const int tid = threadIdx.x;
    const int j = tid & 7;                                // head / channel pair
    const int ql = tid >> 3;                              // query within the tile
    const int q = tile * TQ + ql;
    const bool q_ok = q < N;
    const int qc = q_ok ? q : N - 1;

    // ---- stage the layer weights -------------------------------------------------------------------------
    {
        const float4* src = reinterpret_cast<const float4*>(vd.weights);
        float4* dst = reinterpret_cast<float4*>(s_w);
        for (int i = tid; i < (prm.weight_floats + 3) / 4; i += kThreads) dst[i] = __ldg(src + i);
    }
    __syncthreads();

    // ---- keys / values of every query of this sample: k = Wk (x + pos) + bk, v = Wv x + bv -------------------
    const float* xq = prm.query + (long long)b * prm.query_batch_stride;
    for (int idx = tid; idx < N * 2; idx += kThreads) {
        const int n = idx >> 1, half = idx & 1;          // each thread makes 8 k and 8 v channels of one query
        float xv[C], xp[C];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(xq + (long long)n * C) + c4);
            const float4 p = __ldg(reinterpret_cast<const float4*>(prm.pos + (long long)n * C) + c4);
            xv[4 * c4] = a.x; xv[4 * c4 + 1] = a.y; xv[4 * c4 + 2] = a.z; xv[4 * c4 + 3] = a.w;
            xp[4 * c4] = a.x + p.x; xp[4 * c4 + 1] = a.y + p.y; xp[4 * c4 + 2] = a.z + p.z; xp[4 * c4 + 3] = a.w + p.w;
        }
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            const int ch = half * 8 + o;
            s_k[n * C + ch] = dot16(s_w + IMG::in_w + (C + ch) * C, xp) + s_w[IMG::in_b + C + ch];
            s_v[n * C + ch] = dot16(s_w + IMG::in_w + (2 * C + ch) * C, xv) + s_w[IMG::in_b + 2 * C + ch];
        }
    }

    // ---- this thread's query: state x (2 channels), full vectors through the exchange row --------------------
    float x0, x1;
    float vec[C];                                         // scratch 16-vector
    {
        const float2 a = __ldg(reinterpret_cast<const float2*>(xq + (long long)qc * C) + j);
        x0 = a.x; x1 = a.y;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
            const float4 a4 = __ldg(reinterpret_cast<const float4*>(xq + (long long)qc * C) + c4);
            const float4 p4 = __ldg(reinterpret_cast<const float4*>(prm.pos + (long long)qc * C) + c4);
            vec[4 * c4] = a4.x + p4.x; vec[4 * c4 + 1] = a4.y + p4.y; vec[4 * c4 + 2] = a4.z + p4.z; vec[4 * c4 + 3] = a4.w + p4.w;
        }
    }
    // q projection of head j, pre-scaled by 1/sqrt(D) * log2(e) for exp2f
    const float qscale = 0.70710678118654752f * 1.4426950408889634f;
    const float qa = (dot16(s_w + IMG::in_w + (2 * j) * C, vec) + s_w[IMG::in_b + 2 * j]) * qscale;
    const float qb = (dot16(s_w + IMG::in_w + (2 * j + 1) * C, vec) + s_w[IMG::in_b + 2 * j + 1]) * qscale;
    __syncthreads();                                      // s_k / s_v complete

    // ---- self-attention of head j over the N keys (online softmax) ---------------------------------------------
    // two passes over the keys in shared memory (the scores are two FMAs): first the row maximum, then exp2 / sums.  No
    // running-maximum branch and no rescaling in the loop, so four keys are in flight per trip instead of a dependent chain.
    float mx = -INFINITY, den = 0.0f, o0 = 0.0f, o1 = 0.0f;
    {
        const float2* kp = reinterpret_cast<const float2*>(s_k + 2 * j);
        const float2* vp = reinterpret_cast<const float2*>(s_v + 2 * j);
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
        int n = 0;
        for (; n + 4 <= N; n += 4) {
            const float2 k0 = kp[(n + 0) * (C / 2)], k1 = kp[(n + 1) * (C / 2)], k2 = kp[(n + 2) * (C / 2)], k3 = kp[(n + 3) * (C / 2)];
            m0 = fmaxf(m0, fmaf(qa, k0.x, qb * k0.y));
            m1 = fmaxf(m1, fmaf(qa, k1.x, qb * k1.y));
            m2 = fmaxf(m2, fmaf(qa, k2.x, qb * k2.y));
            m3 = fmaxf(m3, fmaf(qa, k3.x, qb * k3.y));
        }
        for (; n < N; ++n) {
            const float2 k0 = kp[n * (C / 2)];
            m0 = fmaxf(m0, fmaf(qa, k0.x, qb * k0.y));
        }
        mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        float d0 = 0.0f, d1 = 0.0f, a0 = 0.0f, a1 = 0.0f, b0 = 0.0f, b1 = 0.0f;
        n = 0;
        for (; n + 2 <= N; n += 2) {
            const float2 k0 = kp[n * (C / 2)], k1 = kp[(n + 1) * (C / 2)];
            const float2 v0 = vp[n * (C / 2)], v1 = vp[(n + 1) * (C / 2)];
            const float p0 = exp2f(fmaf(qa, k0.x, qb * k0.y) - mx);
            const float p1 = exp2f(fmaf(qa, k1.x, qb * k1.y) - mx);
            d0 += p0; d1 += p1;
            a0 = fmaf(p0, v0.x, a0); a1 = fmaf(p0, v0.y, a1);
            b0 = fmaf(p1, v1.x, b0); b1 = fmaf(p1, v1.y, b1);
        }
        for (; n < N; ++n) {
            const float2 k0 = kp[n * (C / 2)], v0 = vp[n * (C / 2)];
            const float p0 = exp2f(fmaf(qa, k0.x, qb * k0.y) - mx);
            d0 += p0;
            a0 = fmaf(p0, v0.x, a0); a1 = fmaf(p0, v0.y, a1);
        }
        den = d0 + d1; o0 = a0 + b0; o1 = a1 + b1;
    }
    {
        const float inv = 1.0f / den;
        s_x[ql * ROW + 2 * j] = o0 * inv;
        s_x[ql * ROW + 2 * j + 1] = o1 * inv;
    }
    __syncwarp();                                         // the 8 lanes of a query sit in one warp
#pragma unroll
    for (int c = 0; c < C; ++c) vec[c] = s_x[ql * ROW + c];
    __syncwarp();
    x0 += dot16(s_w + IMG::out_w + (2 * j) * C, vec) + s_w[IMG::out_b + 2 * j];
    x1 += dot16(s_w + IMG::out_w + (2 * j + 1) * C, vec) + s_w[IMG::out_b + 2 * j + 1];
    layer_norm2(x0, x1, s_w + IMG::ln1_w, s_w + IMG::ln1_b, j);

    // ---- reference point of this query in this view -------------------------------------------------------------
    float ref_u, ref_v;
    {
        const float* cp = prm.center + (long long)b * prm.center_batch_stride + (long long)qc * 3;
        float px = __ldg(cp), py = __ldg(cp + 1), pz = __ldg(cp + 2);
        if (__ldg(vd.use_transform) != 0) {
            const float* T = vd.transform + b * 16;
            const float tx = T[0] * px + T[1] * py + T[2] * pz + T[3];
            const float ty = T[4] * px + T[5] * py + T[6] * pz + T[7];
            const float tz = T[8] * px + T[9] * py + T[10] * pz + T[11];
            const float r = sqrtf(tx * tx + ty * ty + tz * tz);
            const float phi = atan2f(ty, tx);
            const float roh = asinf(r != 0.0f ? tz / r : 0.0f);
            px = r; py = phi * 57.29577951308232f; pz = roh * 57.29577951308232f;
        }
        const float* Pm = vd.projection + b * 16;
        float u = Pm[0] * px + Pm[1] * py + Pm[2] * pz + Pm[3];
        float w_ = Pm[8] * px + Pm[9] * py + Pm[10] * pz + Pm[11];
        float vv = Pm[4] * px + Pm[5] * py + Pm[6] * pz + Pm[7];
        if (w_ != 0.0f) { u = u / w_; vv = vv / w_; }
        u = u / __ldg(vd.shape_hw + b * 2 + 1);
        vv = vv / __ldg(vd.shape_hw + b * 2);
        ref_u = fminf(fmaxf(u, 0.0f), 1.0f);
        ref_v = fminf(fmaxf(vv, 0.0f), 1.0f);
    }

    // ---- deformable cross-attention of head j -----------------------------------------------------------------------
    s_x[ql * ROW + 2 * j] = x0 + __ldg(prm.pos + (long long)qc * C + 2 * j);
    s_x[ql * ROW + 2 * j + 1] = x1 + __ldg(prm.pos + (long long)qc * C + 2 * j + 1);
    __syncwarp();
#pragma unroll
    for (int c = 0; c < C; ++c) vec[c] = s_x[ql * ROW + c];
    __syncwarp();

    float logit[LP];
    float lmax = -INFINITY;
#pragma unroll
    for (int i = 0; i < LP; ++i) {
        logit[i] = dot16(s_w + IMG::att_w + j * IMG::att_stride + i * C, vec) + s_w[IMG::att_b + j * LP + i];
        lmax = fmaxf(lmax, logit[i]);
    }
    float lsum = 0.0f;
#pragma unroll
    for (int i = 0; i < LP; ++i) {
        logit[i] = expf(logit[i] - lmax);
        lsum += logit[i];
    }
    const float linv = 1.0f / lsum;

    float agg[C];
#pragma unroll
    for (int c = 0; c < C; ++c) agg[c] = 0.0f;
    float inb = 0.0f;                                     // attention-weighted in-bounds weight mass (bias term)
    const PT* pyr = reinterpret_cast<const PT*>(vd.pyramid) + (long long)b * vd.S * C;
    // f16 rows are half as wide: two samples (sixteen 16-byte requests) are kept in flight per trip, as with fp32 rows
    constexpr int SPT = sizeof(PT) == 2 ? 2 : 1;          // samples per trip
    static_assert(P % SPT == 0, "points per level must be even for the f16 pyramid path");
#pragma unroll
    for (int l = 0; l < L; ++l) {
        const int H = vd.h[l], W = vd.w[l];
        const PT* lvl = pyr + vd.start[l] * C;
#pragma unroll
        for (int p = 0; p < P; p += SPT) {
            Corners cs[SPT];
#pragma unroll
            for (int u = 0; u < SPT; ++u) {
                const int i = l * P + p + u;
                const float* orow = s_w + IMG::off_w + j * IMG::off_stride + (2 * i) * C;
                const float ox = dot16(orow, vec) + s_w[IMG::off_b + (j * LP + i) * 2];
                const float oy = dot16(orow + C, vec) + s_w[IMG::off_b + (j * LP + i) * 2 + 1];
                cs[u] = corners_of(ref_u + ox / (float)W, ref_v + oy / (float)H, logit[i] * linv, H, W);
            }
            RowRegs<PT> rows[SPT][4];
#pragma unroll
            for (int u = 0; u < SPT; ++u) {
#pragma unroll
                for (int k = 0; k < 4; ++k) rows[u][k].load(lvl, cs[u].px[k]);
            }
#pragma unroll
            for (int u = 0; u < SPT; ++u) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    rows[u][k].accumulate(cs[u].wt[k], agg);
                    inb += cs[u].wt[k];
                }
            }
        }
    }
    // value projection of head j after the reduce, then the output projection over all heads
    {
        const float v0 = dot16(s_w + IMG::val_w + (2 * j) * C, agg) + s_w[IMG::val_b + 2 * j] * inb;
        const float v1 = dot16(s_w + IMG::val_w + (2 * j + 1) * C, agg) + s_w[IMG::val_b + 2 * j + 1] * inb;
        s_x[ql * ROW + 2 * j] = v0;
        s_x[ql * ROW + 2 * j + 1] = v1;
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < C; ++c) vec[c] = s_x[ql * ROW + c];
    __syncwarp();
    x0 += dot16(s_w + IMG::prj_w + (2 * j) * C, vec) + s_w[IMG::prj_b + 2 * j];
    x1 += dot16(s_w + IMG::prj_w + (2 * j + 1) * C, vec) + s_w[IMG::prj_b + 2 * j + 1];
    layer_norm2(x0, x1, s_w + IMG::ln2_w, s_w + IMG::ln2_b, j);

    // ---- feed-forward ----------------------------------------------------------------------------------------------
    const int F = prm.d_ffn;
    const float* f1w = s_w + IMG::ffn;
    const float* f1b = f1w + F * C;
    const float* f2w = f1b + F;
    const float* f2b = f2w + C * F;
    s_x[ql * ROW + 2 * j] = x0;
    s_x[ql * ROW + 2 * j + 1] = x1;
    __syncwarp();
#pragma unroll
    for (int c = 0; c < C; ++c) vec[c] = s_x[ql * ROW + c];
    float* hrow = s_h + ql * (F + 1);
    for (int f = j; f < F; f += NH) hrow[f] = activation(dot16(f1w + f * C, vec) + f1b[f], prm.act);
    __syncwarp();
    float y0 = f2b[2 * j], y1 = f2b[2 * j + 1];
    for (int f = 0; f < F; ++f) {
        const float hv = hrow[f];
        y0 = fmaf(f2w[(2 * j) * F + f], hv, y0);
        y1 = fmaf(f2w[(2 * j + 1) * F + f], hv, y1);
    }
    x0 += y0; x1 += y1;
    layer_norm2(x0, x1, s_w + IMG::ln3_w, s_w + IMG::ln3_b, j);

    if (q_ok) {
        float* o = prm.out + (((long long)b * prm.V + v) * N + q) * C + 2 * j;
        *reinterpret_cast<float2*>(o) = make_float2(x0, x1);
    }
}

// ---- view reduction + detection head ---------------------------------------------------------------------------------
struct HeadParams {
    const float* views;        // (B, V, N, 16) decoder-layer outputs
    const float* weights;      // packed: reduction [16][V*16] (input index c*V + v), then 4 branches x {[16][16],[16][16],[k][16]}
    const float* center_in;    // (B, N, 3) or (N, 3)
    float* query_out;          // (B, N, 16)
    float* center_out;         // (B, N, 3)
    float* size_out;           // (B, N, 3) or null (intermediate iterations)
    float* angle_out;          // (B, N, 2) or null
    float* class_out;          // (B, N, n_cls) or null
    long long center_batch_stride;
    int B, V, N, n_cls, reduction;   // reduction: 0 = linear, 1 = mean, 2 = max
    int weight_floats;
};

__device__ __forceinline__ void mlp3(const float* w, const float* x, int k, float* out) {
    float h1[C], h2[C];
#pragma unroll
    for (int o = 0; o < C; ++o) h1[o] = fmaxf(dot16(w + o * C, x), 0.0f);
#pragma unroll
    for (int o = 0; o < C; ++o) h2[o] = fmaxf(dot16(w + C * C + o * C, h1), 0.0f);
    for (int o = 0; o < k; ++o) out[o] = dot16(w + 2 * C * C + o * C, h2);
}

__global__ void __launch_bounds__(128)
decoder_head_kernel(const HeadParams prm) {
    extern __shared__ __align__(16) float smem[];
    for (int i = threadIdx.x; i < prm.weight_floats; i += blockDim.x) smem[i] = __ldg(prm.weights + i);
    __syncthreads();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)prm.B * prm.N) return;
    const int n = (int)(idx % prm.N);
    const int b = (int)(idx / prm.N);
    const int V = prm.V;
    float x[C];
    const float* red = smem;
    if (prm.reduction == 0) {
#pragma unroll
        for (int o = 0; o < C; ++o) x[o] = 0.0f;
        for (int v = 0; v < V; ++v) {
            const float4* row = reinterpret_cast<const float4*>(prm.views + (((long long)b * V + v) * prm.N + n) * C);
            float xv[C];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float4 f = __ldg(row + c4);
                xv[4 * c4] = f.x; xv[4 * c4 + 1] = f.y; xv[4 * c4 + 2] = f.z; xv[4 * c4 + 3] = f.w;
            }
#pragma unroll
            for (int o = 0; o < C; ++o) {
                float acc = x[o];
#pragma unroll
                for (int c = 0; c < C; ++c) acc = fmaf(red[o * (V * C) + c * V + v], xv[c], acc);
                x[o] = acc;
            }
        }
    } else {
#pragma unroll
        for (int o = 0; o < C; ++o) x[o] = prm.reduction == 1 ? 0.0f : -INFINITY;
        for (int v = 0; v < V; ++v) {
            const float* row = prm.views + (((long long)b * V + v) * prm.N + n) * C;
#pragma unroll
            for (int o = 0; o < C; ++o) x[o] = prm.reduction == 1 ? x[o] + __ldg(row + o) / (float)V : fmaxf(x[o], __ldg(row + o));
        }
    }
    float* qo = prm.query_out + idx * C;
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4)
        reinterpret_cast<float4*>(qo)[c4] = make_float4(x[4 * c4], x[4 * c4 + 1], x[4 * c4 + 2], x[4 * c4 + 3]);

    const float* hw = smem + (prm.reduction == 0 ? C * V * C : 0);
    const int branch = 2 * C * C;                          // two hidden layers, then k x 16
    float o3[3];
    mlp3(hw, x, 3, o3);                                    // centre: Identity, + previous centre
    const float* cp = prm.center_in + (long long)b * prm.center_batch_stride + (long long)n * 3;
    prm.center_out[idx * 3] = o3[0] + __ldg(cp);
    prm.center_out[idx * 3 + 1] = o3[1] + __ldg(cp + 1);
    prm.center_out[idx * 3 + 2] = o3[2] + __ldg(cp + 2);
    if (prm.size_out) {
        hw += branch + 3 * C;
        mlp3(hw, x, 3, o3);                                // size: ReLU
        prm.size_out[idx * 3] = fmaxf(o3[0], 0.0f);
        prm.size_out[idx * 3 + 1] = fmaxf(o3[1], 0.0f);
        prm.size_out[idx * 3 + 2] = fmaxf(o3[2], 0.0f);
        hw += branch + 3 * C;
        mlp3(hw, x, 2, o3);                                // angle: Tanh
        prm.angle_out[idx * 2] = tanhf(o3[0]);
        prm.angle_out[idx * 2 + 1] = tanhf(o3[1]);
        hw += branch + 2 * C;
        float oc[8];
        mlp3(hw, x, prm.n_cls, oc);                        // class: Identity (logits)
        for (int k = 0; k < prm.n_cls; ++k) prm.class_out[idx * prm.n_cls + k] = oc[k];
    }
}

// Default since round 2 (bit-identical to decoder_head_kernel on B200: tests/test_decoder_head16_gpu.py; DPFT_HEAD_LANES=1 in the
// environment selects the one-thread-per-query kernel for A/B): the same reduction + head with SIXTEEN lanes per query instead of one thread per query.  decoder_head_kernel is a
// serial chain of ~3000 instructions per thread on 19 CTAs (B*N = 2400 threads) at the end of every decoder iteration:
// 19 us of pure latency, four times per forward.  Here lane o of a query owns output channel o of every layer (one dot16
// per layer instead of sixteen), the 16-vectors go through a padded shared-memory row exactly as in decoder_layer_kernel,
// and B*N/16 CTAs cover the GPU.  Every output is accumulated in the same order as in decoder_head_kernel (fmaf chains from
// zero, views outer / channels inner in the reduction), so the results are bit-identical.
constexpr int HQ = 16;                                 // queries per CTA (16 lanes each)

__global__ void __launch_bounds__(HQ * C)
decoder_head16_kernel(const HeadParams prm) {
    extern __shared__ __align__(16) float smem[];
    float* s_x = smem + ((prm.weight_floats + 3) & ~3);    // [HQ][ROW] exchange rows
    for (int i = threadIdx.x; i < prm.weight_floats; i += blockDim.x) smem[i] = __ldg(prm.weights + i);
    __syncthreads();
    const int o = threadIdx.x & (C - 1);                   // output channel owned by this lane
    const int ql = threadIdx.x >> 4;
    const long long total = (long long)prm.B * prm.N;
    long long idx = (long long)blockIdx.x * HQ + ql;
    const bool ok = idx < total;
    if (!ok) idx = total - 1;                              // keep the lanes alive for the warp-level exchanges; they never store
    const int n = (int)(idx % prm.N);
    const int b = (int)(idx / prm.N);
    const int V = prm.V;
    float* row = s_x + ql * ROW;
    float vec[C];

    float x;
    if (prm.reduction == 0) {
        const float* red = smem;
        x = 0.0f;
        for (int v = 0; v < V; ++v) {
            const float4* src = reinterpret_cast<const float4*>(prm.views + (((long long)b * V + v) * prm.N + n) * C);
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float4 f = __ldg(src + c4);
                vec[4 * c4] = f.x; vec[4 * c4 + 1] = f.y; vec[4 * c4 + 2] = f.z; vec[4 * c4 + 3] = f.w;
            }
#pragma unroll
            for (int c = 0; c < C; ++c) x = fmaf(red[o * (V * C) + c * V + v], vec[c], x);
        }
    } else {
        x = prm.reduction == 1 ? 0.0f : -INFINITY;
        for (int v = 0; v < V; ++v) {
            const float* src = prm.views + (((long long)b * V + v) * prm.N + n) * C;
            x = prm.reduction == 1 ? x + __ldg(src + o) / (float)V : fmaxf(x, __ldg(src + o));
        }
    }
    if (ok) prm.query_out[idx * C + o] = x;

    auto exchange = [&](float mine) {                      // every lane of the query gets the full 16-vector
        row[o] = mine;
        __syncwarp();
#pragma unroll
        for (int c = 0; c < C; ++c) vec[c] = row[c];
        __syncwarp();
    };
    float xin[C];
    exchange(x);
#pragma unroll
    for (int c = 0; c < C; ++c) xin[c] = vec[c];
    // one branch: Linear(16,16) -> ReLU -> Linear(16,16) -> ReLU -> Linear(16,k); lane o < k returns output o
    auto mlp = [&](const float* w, int k) -> float {
        exchange(fmaxf(dot16(w + o * C, xin), 0.0f));
        exchange(fmaxf(dot16(w + C * C + o * C, vec), 0.0f));
        return o < k ? dot16(w + 2 * C * C + o * C, vec) : 0.0f;
    };
    const float* hw = smem + (prm.reduction == 0 ? C * V * C : 0);
    const int branch = 2 * C * C;
    float r = mlp(hw, 3);                                  // centre: Identity, + previous centre
    if (ok && o < 3)
        prm.center_out[idx * 3 + o] = r + __ldg(prm.center_in + (long long)b * prm.center_batch_stride + (long long)n * 3 + o);
    if (prm.size_out) {
        hw += branch + 3 * C;
        r = mlp(hw, 3);                                    // size: ReLU
        if (ok && o < 3) prm.size_out[idx * 3 + o] = fmaxf(r, 0.0f);
        hw += branch + 3 * C;
        r = mlp(hw, 2);                                    // angle: Tanh
        if (ok && o < 2) prm.angle_out[idx * 2 + o] = tanhf(r);
        hw += branch + 2 * C;
        r = mlp(hw, prm.n_cls);                            // class: Identity (logits)
        if (ok && o < prm.n_cls) prm.class_out[idx * prm.n_cls + o] = r;
    }
}

template <int L, int P, typename PT>
int launch_layer(const LayerParams& prm, cudaStream_t stream) {
    const size_t smem = sizeof(float) * (((prm.weight_floats + 3) & ~3) + (size_t)prm.N * C * 2 + TQ * ROW + TQ * (prm.d_ffn + 1));
    DPFT_REQUIRE(smem <= 227 * 1024, "decoder: %zu bytes of shared memory needed (N=%d too large)", smem, prm.N);
    using IMG = LayerImage<L, P>;
    const int expected = IMG::fixed + prm.d_ffn * C + prm.d_ffn + C * prm.d_ffn + C;
    DPFT_REQUIRE(prm.weight_floats == expected, "decoder: packed layer image has %d floats, expected %d",
                 prm.weight_floats, expected);
    static size_t configured = 0;                      // raise the limit only when needed (never inside a graph replay)
    if (smem > configured) {
        auto kern_attr = decoder_layer_kernel<L, P, PT>;
        int st = cuda_status(cudaFuncSetAttribute(kern_attr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                             "cudaFuncSetAttribute(decoder_layer_kernel)");
        if (st) return st;
        configured = smem;
    }
    const int tiles = (prm.N + TQ - 1) / TQ;
    auto kern = decoder_layer_kernel<L, P, PT>;
    kern<<<prm.B * prm.V * tiles, kThreads, smem, stream>>>(prm);
    DPFT_LAUNCH_CHECK("decoder_layer_kernel");
    return DPFT_OK;
}

}  // namespace
}  // namespace dpft

using namespace dpft;

extern "C" int dpft_decoder_layer_forward(const dpft_decoder_view* views, int V, const float* query, long long query_batch_stride,
                                          const float* pos, const float* center, long long center_batch_stride, float* out,
                                          int B, int N, int L, int P, int d_ffn, int activation, int weight_floats,
                                          int pyramid_dtype, void* stream) {
    DPFT_REQUIRE(pyramid_dtype == DPFT_F32 || pyramid_dtype == DPFT_F16, "decoder_layer: pyramid dtype must be DPFT_F32 or DPFT_F16");
    DPFT_REQUIRE(views && query && pos && center && out, "decoder_layer: null pointer");
    DPFT_REQUIRE(V >= 1 && V <= kMaxViews, "decoder_layer: V=%d views (1..%d supported)", V, kMaxViews);
    DPFT_REQUIRE(B >= 1 && N >= 1 && d_ffn >= 1 && d_ffn % 4 == 0, "decoder_layer: bad sizes B=%d N=%d d_ffn=%d", B, N, d_ffn);
    DPFT_REQUIRE(L >= 1 && L <= kMaxLevels, "decoder_layer: L=%d levels (1..%d supported)", L, kMaxLevels);
    LayerParams prm{};
    for (int v = 0; v < V; ++v) {
        const dpft_decoder_view& s = views[v];
        DPFT_REQUIRE(s.pyramid && s.weights && s.transform && s.projection && s.shape_hw && s.use_transform,
                     "decoder_layer: null pointer in view %d", v);
        ViewDesc& d = prm.view[v];
        d.pyramid = s.pyramid; d.weights = s.weights; d.transform = s.transform; d.projection = s.projection;
        d.shape_hw = s.shape_hw; d.use_transform = s.use_transform; d.S = s.S;
        for (int l = 0; l < L; ++l) { d.h[l] = s.level_h[l]; d.w[l] = s.level_w[l]; d.start[l] = s.level_start[l]; }
    }
    prm.query = query; prm.pos = pos; prm.center = center; prm.out = out;
    prm.query_batch_stride = query_batch_stride; prm.center_batch_stride = center_batch_stride;
    prm.B = B; prm.V = V; prm.N = N; prm.d_ffn = d_ffn; prm.act = activation; prm.weight_floats = weight_floats;
    cudaStream_t s = (cudaStream_t)stream;
    if (P == 4) {
        const bool h = pyramid_dtype == DPFT_F16;
        switch (L) {
            case 1: return h ? launch_layer<1, 4, __half>(prm, s) : launch_layer<1, 4, float>(prm, s);
            case 2: return h ? launch_layer<2, 4, __half>(prm, s) : launch_layer<2, 4, float>(prm, s);
            case 3: return h ? launch_layer<3, 4, __half>(prm, s) : launch_layer<3, 4, float>(prm, s);
            case 4: return h ? launch_layer<4, 4, __half>(prm, s) : launch_layer<4, 4, float>(prm, s);
            case 5: return h ? launch_layer<5, 4, __half>(prm, s) : launch_layer<5, 4, float>(prm, s);
            default: break;
        }
    }
    set_error("decoder_layer: no fused instance for L=%d, P=%d (P=4 with L=1..5 are built)", L, P);
    return DPFT_ERR_UNSUPPORTED;
}

extern "C" int dpft_decoder_head_forward(const float* views, const float* weights, const float* center_in,
                                         long long center_batch_stride, float* query_out, float* center_out, float* size_out,
                                         float* angle_out, float* class_out, int B, int V, int N, int n_cls, int reduction,
                                         int weight_floats, void* stream) {
    // kernel selection: sixteen lanes per query unless DPFT_HEAD_LANES=1 (reduction | DPFT_HEAD_LANES16 forces it: tests)
    static const int env_lanes = [] { const char* e = getenv("DPFT_HEAD_LANES"); return e ? atoi(e) : 16; }();
    const bool lanes16 = (reduction & DPFT_HEAD_LANES16) != 0 || (env_lanes == 16 && (reduction & DPFT_HEAD_LANES1) == 0);
    reduction &= ~(DPFT_HEAD_LANES16 | DPFT_HEAD_LANES1);
    DPFT_REQUIRE(reduction >= 0 && reduction <= 2, "decoder_head: reduction=%d (0 linear, 1 mean, 2 max)", reduction);
    DPFT_REQUIRE(views && weights && center_in && query_out && center_out, "decoder_head: null pointer");
    DPFT_REQUIRE(!size_out == !angle_out && !size_out == !class_out, "decoder_head: size/angle/class outputs go together");
    DPFT_REQUIRE(n_cls >= 1 && n_cls <= 8, "decoder_head: n_cls=%d (1..8 supported)", n_cls);
    DPFT_REQUIRE(weight_floats * 4 <= 200 * 1024, "decoder_head: weights too large");
    HeadParams prm{views, weights, center_in, query_out, center_out, size_out, angle_out, class_out,
                   center_batch_stride, B, V, N, n_cls, reduction, weight_floats};
    const size_t smem = (size_t)weight_floats * 4;
    static size_t configured = 0;
    if (smem > configured) {
        int st = cuda_status(cudaFuncSetAttribute(decoder_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                             "cudaFuncSetAttribute(decoder_head_kernel)");
        if (st) return st;
        configured = smem;
    }
    const long long total = (long long)B * N;
    if (lanes16) {
        const size_t smem16 = (size_t)((weight_floats + 3) & ~3) * 4 + HQ * ROW * 4;
        static size_t configured16 = 0;
        if (smem16 > configured16) {
            int st = cuda_status(cudaFuncSetAttribute(decoder_head16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16),
                                 "cudaFuncSetAttribute(decoder_head16_kernel)");
            if (st) return st;
            configured16 = smem16;
        }
        decoder_head16_kernel<<<(unsigned)((total + HQ - 1) / HQ), HQ * C, smem16, (cudaStream_t)stream>>>(prm);
        DPFT_LAUNCH_CHECK("decoder_head16_kernel");
        return DPFT_OK;
    }
    decoder_head_kernel<<<(unsigned)((total + 127) / 128), 128, smem, (cudaStream_t)stream>>>(prm);
    DPFT_LAUNCH_CHECK("decoder_head_kernel");
    return DPFT_OK;
}
