// Training-mode companions of the tcgen05 convolutions (sm_100a): BatchNorm with batch statistics (forward and backward),
// ReLU / residual add folded into the normalisation passes, max-pool backward, zero insertion for strided data gradients
// and the per-step weight re-layout.  All NHWC, 16-bit activations (bf16 or f16), fp32 statistics; every kernel is a
// streaming pass bound by HBM bandwidth: 16-byte accesses, one thread owns 8 channels of a pixel and keeps its
// per-channel constants in registers across a grid-stride loop.
//
// Reference semantics: torchvision Bottleneck blocks in train() (src/dprt/models/backbones/resnet.py:54-55,101;
// norm_layer BatchNorm2d, config/kradar.json:86) under autograd (src/dprt/training/trainer.py:125-133).
#include "common.cuh"

namespace dpft {
namespace {

constexpr int TB = 256;
constexpr int RB = 512;           // threads per block of the reduction passes (<= 2 blocks per SM)

template <bool F16> struct H2;
template <> struct H2<true> {
    static __device__ __forceinline__ float2 unpack(uint32_t w) { return __half22float2(*reinterpret_cast<const __half2*>(&w)); }
    static __device__ __forceinline__ uint32_t pack(float a, float b) {
        uint32_t o;
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(o) : "f"(b), "f"(a));
        return o;
    }
};
template <> struct H2<false> {
    static __device__ __forceinline__ float2 unpack(uint32_t w) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w)); }
    static __device__ __forceinline__ uint32_t pack(float a, float b) {
        uint32_t o;
        asm("cvt.rn.satfinite.bf16x2.f32 %0, %1, %2;" : "=r"(o) : "f"(b), "f"(a));
        return o;
    }
};

template <bool F16> __device__ __forceinline__ void unpack8(const uint4& r, float* f) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&r);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const float2 v = H2<F16>::unpack(w[t]);
        f[2 * t] = v.x;
        f[2 * t + 1] = v.y;
    }
}
template <bool F16> __device__ __forceinline__ uint4 pack8(const float* f) {
    uint4 r;
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int t = 0; t < 4; ++t) w[t] = H2<F16>::pack(f[2 * t], f[2 * t + 1]);
    return r;
}

__device__ __forceinline__ uint4 ld_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// Block-level reduction of per-thread partial sums that belong to channel group (threadIdx.x % groups): NV values per thread
// (NV = 16: two quantities x 8 channels).  The block's result goes to row blockIdx.x of a [gridDim.x][2][C] partial buffer; a
// second small kernel sums the rows.  (Atomics on the 2C result addresses from ~1000 blocks serialise in L2: ~80 us per launch.)
template <int NV>
__device__ __forceinline__ void block_reduce_to_partials(const float (&v)[NV], int groups, float* partials, int C) {
    __shared__ float red[RB][NV + 1];
#pragma unroll
    for (int k = 0; k < NV; ++k) red[threadIdx.x][k] = v[k];
    __syncthreads();
    // thread t < groups * NV sums value k = t % NV of channel group t / NV over the RB / groups threads that own it
    for (int t = threadIdx.x; t < groups * NV; t += RB) {
        const int grp = t / NV, k = t - grp * NV;
        float acc = 0.0f;
        for (int u = grp; u < RB; u += groups) acc += red[u][k];
        partials[(size_t)blockIdx.x * 2 * C + (k / 8) * C + grp * 8 + (k % 8)] = acc;
    }
}

// column sums of a [nparts][n] partial buffer; blockDim = (32, CS_ROWS), one block per 32 columns.  These are the 420 tiny
// launches between the streaming passes of a training step: each thread's chain of dependent L2 round trips is what they
// cost, so the rows are spread over 32 threads per column (was 8: ~10 round trips per thread, ~8 us per launch).
constexpr int CS_ROWS = 32;
__device__ __forceinline__ float column_sum(const float* __restrict__ partials, int nparts, int n, int col) {
    __shared__ float part[CS_ROWS][33];
    float acc = 0.0f;
    if (col < n) {
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;                 // four independent loads in flight
        int r = threadIdx.y;
        for (; r + 3 * CS_ROWS < nparts; r += 4 * CS_ROWS) {
            a0 += partials[(size_t)r * n + col];
            a1 += partials[(size_t)(r + CS_ROWS) * n + col];
            a2 += partials[(size_t)(r + 2 * CS_ROWS) * n + col];
            a3 += partials[(size_t)(r + 3 * CS_ROWS) * n + col];
        }
        for (; r < nparts; r += CS_ROWS) a0 += partials[(size_t)r * n + col];
        acc = (a0 + a1) + (a2 + a3);
    }
    part[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    float tot = 0.0f;
#pragma unroll
    for (int r = 0; r < CS_ROWS; ++r) tot += part[r][threadIdx.x];
    __syncthreads();
    return tot;
}

// ---- BatchNorm forward ---------------------------------------------------------------------------------------------
// pass 1: per-channel sum and sum of squares of y (M, C)
template <bool F16>
__global__ void __launch_bounds__(RB) bn_stats_kernel(const uint4* __restrict__ y, float* __restrict__ partials, long long nvec, int groups) {
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.0f;
    const long long stride = (long long)gridDim.x * RB;
    long long i = (long long)blockIdx.x * RB + threadIdx.x;
    for (; i + 3 * stride < nvec; i += 4 * stride) {          // four independent 16-byte loads in flight per thread
        uint4 r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) r[u] = ld_stream(y + i + u * stride);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float f[8];
            unpack8<F16>(r[u], f);
#pragma unroll
            for (int k = 0; k < 8; ++k) { acc[k] += f[k]; acc[8 + k] = fmaf(f[k], f[k], acc[8 + k]); }
        }
    }
    for (; i < nvec; i += stride) {
        float f[8];
        unpack8<F16>(ld_stream(y + i), f);
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc[k] += f[k]; acc[8 + k] = fmaf(f[k], f[k], acc[8 + k]); }
    }
    block_reduce_to_partials<16>(acc, groups, partials, groups * 8);
}

// pass 2 (C threads): batch mean / inverse std, the affine form used by the apply pass, and the running statistics
// (momentum update with the unbiased variance, torch.nn.BatchNorm2d semantics).
__global__ void bn_finalize_kernel(const float* __restrict__ partials, int nparts, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* running_mean, float* running_var, float momentum, float eps,
                                   float count, int C, float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out) {
    const int c = blockIdx.x * 32 + threadIdx.x;
    const float sum = column_sum(partials, nparts, 2 * C, c);
    const float sumsq = column_sum(partials, nparts, 2 * C, C + c);
    if (c >= C || threadIdx.y != 0) return;
    const float mean = sum / count;
    const float var = fmaxf(sumsq / count - mean * mean, 0.0f);
    const float invstd = rsqrtf(var + eps);
    const float g = gamma ? gamma[c] : 1.0f, b = beta ? beta[c] : 0.0f;
    scale[c] = g * invstd;
    shift[c] = b - mean * g * invstd;
    mean_out[c] = mean;
    invstd_out[c] = invstd;
    if (running_mean) running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * mean;
    if (running_var) running_var[c] = (1.0f - momentum) * running_var[c] + momentum * var * (count > 1.0f ? count / (count - 1.0f) : 1.0f);
}

// pass 3: z = act(y * scale + shift (+ residual))
template <bool F16>
__global__ void __launch_bounds__(TB) bn_apply_kernel(const uint4* __restrict__ y, const float* __restrict__ scale,
                                                      const float* __restrict__ shift, const uint4* __restrict__ residual,
                                                      uint4* __restrict__ z, long long nvec, int groups, int relu) {
    const int c0 = (threadIdx.x % groups) * 8;
    float sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { sc[k] = scale[c0 + k]; sh[k] = shift[c0 + k]; }
    const long long stride = (long long)gridDim.x * TB;
    for (long long i = (long long)blockIdx.x * TB + threadIdx.x; i < nvec; i += 2 * stride) {
        const bool two = i + stride < nvec;
        uint4 ry[2], rr[2];
        ry[0] = ld_stream(y + i);
        if (two) ry[1] = ld_stream(y + i + stride);
        if (residual) {
            rr[0] = ld_stream(residual + i);
            if (two) rr[1] = ld_stream(residual + i + stride);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (u == 1 && !two) break;
            float f[8];
            unpack8<F16>(ry[u], f);
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = fmaf(f[k], sc[k], sh[k]);
            if (residual) {
                float g[8];
                unpack8<F16>(rr[u], g);
#pragma unroll
                for (int k = 0; k < 8; ++k) f[k] += g[k];
            }
            if (relu) {
#pragma unroll
                for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.0f);
            }
            z[i + u * stride] = pack8<F16>(f);
        }
    }
}

// ---- BatchNorm backward --------------------------------------------------------------------------------------------
// pass 1: g = dz * [z > 0] (ReLU mask when relu), sum_g[c] = sum g, sum_gx[c] = sum g * xhat, xhat = (y - mean) * invstd
template <bool F16>
__global__ void __launch_bounds__(RB) bn_bwd_reduce_kernel(const uint4* __restrict__ dz, const uint4* __restrict__ z,
                                                           const uint4* __restrict__ y, const float* __restrict__ mean,
                                                           const float* __restrict__ invstd, float* __restrict__ partials,
                                                           long long nvec, int groups, int relu) {
    const int c0 = (threadIdx.x % groups) * 8;
    float mu[8], is[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { mu[k] = mean[c0 + k]; is[k] = invstd[c0 + k]; }
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.0f;
    const long long stride = (long long)gridDim.x * RB;
    for (long long i = (long long)blockIdx.x * RB + threadIdx.x; i < nvec; i += 2 * stride) {
        const bool two = i + stride < nvec;
        uint4 rd[2], rz[2], ry[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (u == 1 && !two) break;
            rd[u] = ld_stream(dz + i + u * stride);
            ry[u] = ld_stream(y + i + u * stride);
            if (relu) rz[u] = ld_stream(z + i + u * stride);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (u == 1 && !two) break;
            float g[8], fy[8];
            unpack8<F16>(rd[u], g);
            unpack8<F16>(ry[u], fy);
            if (relu) {
                float fz[8];
                unpack8<F16>(rz[u], fz);
#pragma unroll
                for (int k = 0; k < 8; ++k) g[k] = fz[k] > 0.0f ? g[k] : 0.0f;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                acc[k] += g[k];
                acc[8 + k] = fmaf(g[k], (fy[k] - mu[k]) * is[k], acc[8 + k]);
            }
        }
    }
    block_reduce_to_partials<16>(acc, groups, partials, groups * 8);
}

// column sums of the partial rows -> sum_g | sum_gx, and the affine parameter gradients dgamma += sum_gx, dbeta += sum_g
__global__ void bn_bwd_sums_kernel(const float* __restrict__ partials, int nparts, int C, float* __restrict__ sum_g,
                                   float* __restrict__ sum_gx, float* dgamma, float* dbeta) {
    const int c = blockIdx.x * 32 + threadIdx.x;
    const float sg = column_sum(partials, nparts, 2 * C, c);
    const float sgx = column_sum(partials, nparts, 2 * C, C + c);
    if (c >= C || threadIdx.y != 0) return;
    sum_g[c] = sg;
    sum_gx[c] = sgx;
    if (dgamma) dgamma[c] += sgx;
    if (dbeta) dbeta[c] += sg;
}


// pass 2: dy = gamma * invstd * (g - sum_g / M - xhat * sum_gx / M); optionally stores g (the gradient that flows into
// the block's identity branch).
template <bool F16>
__global__ void __launch_bounds__(TB) bn_bwd_apply_kernel(const uint4* __restrict__ dz, const uint4* __restrict__ z,
                                                          const uint4* __restrict__ y, const float* __restrict__ mean,
                                                          const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                          const float* __restrict__ sum_g, const float* __restrict__ sum_gx,
                                                          uint4* __restrict__ dy, uint4* __restrict__ g_out, float inv_count,
                                                          long long nvec, int groups, int relu) {
    const int c0 = (threadIdx.x % groups) * 8;
    float mu[8], is[8], k1[8], k2[8], k3[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        mu[k] = mean[c0 + k];
        is[k] = invstd[c0 + k];
        const float gi = (gamma ? gamma[c0 + k] : 1.0f) * is[k];
        k1[k] = gi;                                          // dy = k1 * g - k2 - xhat * k3
        k2[k] = gi * sum_g[c0 + k] * inv_count;
        k3[k] = gi * sum_gx[c0 + k] * inv_count;
    }
    const long long stride = (long long)gridDim.x * TB;
    for (long long i = (long long)blockIdx.x * TB + threadIdx.x; i < nvec; i += 2 * stride) {
        const bool two = i + stride < nvec;
        uint4 rd[2], rz[2], ry[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (u == 1 && !two) break;
            rd[u] = ld_stream(dz + i + u * stride);
            ry[u] = ld_stream(y + i + u * stride);
            if (relu) rz[u] = ld_stream(z + i + u * stride);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (u == 1 && !two) break;
            float g[8], fy[8];
            unpack8<F16>(rd[u], g);
            unpack8<F16>(ry[u], fy);
            if (relu) {
                float fz[8];
                unpack8<F16>(rz[u], fz);
#pragma unroll
                for (int k = 0; k < 8; ++k) g[k] = fz[k] > 0.0f ? g[k] : 0.0f;
            }
            if (g_out) g_out[i + u * stride] = pack8<F16>(g);
            float o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = fmaf(k1[k], g[k], -k2[k]) - (fy[k] - mu[k]) * is[k] * k3[k];
            dy[i + u * stride] = pack8<F16>(o);
        }
    }
}

// ---- max-pool 3x3 / stride 2 / pad 1 backward (torch semantics: the whole gradient goes to the first maximum in scan order) ----
// One thread per input position and 8 channels.  With the pooled output at hand a position only has to look at its <= 4
// windows (maximum + gradient) and, where it equals the maximum, at the positions that precede it in the window's scan order.
template <bool F16>
__global__ void __launch_bounds__(TB) maxpool_bwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ pooled,
                                                         const uint4* __restrict__ dy, uint4* __restrict__ dx, int B, int H, int W,
                                                         int CG, int P, int Q) {
    const long long total = (long long)B * H * W * CG;
    for (long long i = (long long)blockIdx.x * TB + threadIdx.x; i < total; i += (long long)gridDim.x * TB) {
        const int cg = (int)(i % CG);
        long long t = i / CG;
        const int w = (int)(t % W);
        t /= W;
        const int h = (int)(t % H);
        const int b = (int)(t / H);
        float own[8], acc[8];
        unpack8<F16>(__ldg(x + i), own);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.0f;
        // windows (p, q) that contain (h, w): 2p - 1 <= h <= 2p + 1
        const int p_lo = h / 2, p_hi = min((h + 1) / 2, P - 1);
        const int q_lo = w / 2, q_hi = min((w + 1) / 2, Q - 1);
        for (int p = p_lo; p <= p_hi; ++p) {
            for (int q = q_lo; q <= q_hi; ++q) {
                const long long o = (((long long)b * P + p) * Q + q) * CG + cg;
                float mx[8], g[8];
                unpack8<F16>(__ldg(pooled + o), mx);
                unpack8<F16>(__ldg(dy + o), g);
                bool win[8];
                bool any = false;
#pragma unroll
                for (int k = 0; k < 8; ++k) { win[k] = own[k] == mx[k]; any |= win[k]; }
                if (any) {
                    const int r_me = h - (2 * p - 1), s_me = w - (2 * q - 1);
                    for (int r = 0; r <= r_me; ++r) {
                        const int hh = 2 * p - 1 + r;
                        if (hh < 0) continue;
                        const int s_end = r < r_me ? 3 : s_me;               // positions strictly before (r_me, s_me)
                        for (int sx = 0; sx < s_end; ++sx) {
                            const int ww = 2 * q - 1 + sx;
                            if (ww < 0 || ww >= W) continue;
                            float v[8];
                            unpack8<F16>(__ldg(x + (((long long)b * H + hh) * W + ww) * CG + cg), v);
#pragma unroll
                            for (int k = 0; k < 8; ++k) win[k] = win[k] && v[k] != mx[k];
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[k] += win[k] ? g[k] : 0.0f;
                }
            }
        }
        dx[i] = pack8<F16>(acc);
    }
}

// ---- zero insertion: up[b, 2p, 2q, :] = src[b, p, q, :], everything else 0 (data gradient of stride-2 convolutions) ----
__global__ void __launch_bounds__(TB) zero_insert2_kernel(const uint4* __restrict__ src, uint4* __restrict__ up, int B, int H, int W,
                                                          int CG, int P, int Q) {
    const long long total = (long long)B * H * W * CG;
    for (long long i = (long long)blockIdx.x * TB + threadIdx.x; i < total; i += (long long)gridDim.x * TB) {
        const int cg = (int)(i % CG);
        long long t = i / CG;
        const int w = (int)(t % W);
        t /= W;
        const int h = (int)(t % H);
        const int b = (int)(t / H);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (!(h & 1) && !(w & 1) && (h >> 1) < P && (w >> 1) < Q) v = __ldg(src + (((long long)b * P + (h >> 1)) * Q + (w >> 1)) * CG + cg);
        up[i] = v;
    }
}

// ---- per-step weight re-layout: fp32 master (Cout, Cin, R, S) -> 16-bit forward operand (Cout, R, S, Cin) and the
// data-gradient operand (Cin, R, S, Cout) with the taps flipped (dX = conv(dY, flip(W)^T)) ----
struct PackEntry {
    const float* src;
    void* fwd;
    void* dgrad;         // may be null
    int Cout, Cin, R, S;
    long long offset;    // first element of this layer in the concatenated index space
};

template <bool F16>
__global__ void __launch_bounds__(TB) pack_weights_kernel(const PackEntry* __restrict__ table, int n_layers, long long total) {
    for (long long i = (long long)blockIdx.x * TB + threadIdx.x; i < total; i += (long long)gridDim.x * TB) {
        int lo = 0, hi = n_layers - 1;                       // last entry with offset <= i
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (table[mid].offset <= i) lo = mid; else hi = mid - 1;
        }
        const PackEntry e = table[lo];
        long long j = i - e.offset;                          // index in (Cout, R, S, Cin) order
        const int ci = (int)(j % e.Cin);
        j /= e.Cin;
        const int s = (int)(j % e.S);
        j /= e.S;
        const int r = (int)(j % e.R);
        const int co = (int)(j / e.R);
        const float v = __ldg(e.src + (((long long)co * e.Cin + ci) * e.R + r) * e.S + s);
        if (F16) {
            const __half hv = __float2half_rn(v);
            reinterpret_cast<__half*>(e.fwd)[i - e.offset] = hv;
            if (e.dgrad) reinterpret_cast<__half*>(e.dgrad)[(((long long)ci * e.R + (e.R - 1 - r)) * e.S + (e.S - 1 - s)) * e.Cout + co] = hv;
        } else {
            const __nv_bfloat16 hv = __float2bfloat16_rn(v);
            reinterpret_cast<__nv_bfloat16*>(e.fwd)[i - e.offset] = hv;
            if (e.dgrad) reinterpret_cast<__nv_bfloat16*>(e.dgrad)[(((long long)ci * e.R + (e.R - 1 - r)) * e.S + (e.S - 1 - s)) * e.Cout + co] = hv;
        }
    }
}

// fp32 gradient in the kernel's (Cout, R, S, Cin) layout -> added into the parameter's (Cout, Cin, R, S) gradient
__global__ void __launch_bounds__(TB) unpack_wgrad_kernel(const PackEntry* __restrict__ table, int n_layers, long long total) {
    for (long long i = (long long)blockIdx.x * TB + threadIdx.x; i < total; i += (long long)gridDim.x * TB) {
        int lo = 0, hi = n_layers - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (table[mid].offset <= i) lo = mid; else hi = mid - 1;
        }
        const PackEntry e = table[lo];
        long long j = i - e.offset;                          // index in the parameter's (Cout, Cin, R, S) order
        const int s = (int)(j % e.S);
        j /= e.S;
        const int r = (int)(j % e.R);
        j /= e.R;
        const int ci = (int)(j % e.Cin);
        const int co = (int)(j / e.Cin);
        const float g = __ldg(e.src + (((long long)co * e.R + r) * e.S + s) * e.Cin + ci);
        reinterpret_cast<float*>(e.fwd)[i - e.offset] += g;
    }
}


// ---- stem (conv 7x7 / stride 2 / pad 3, Cin = 3 | 6) weight gradient on the CUDA cores ---------------------------------
//   dw[r][s][c][o] += sum_{b,p,q} dy[b,p,q,o] * x[b, 2p-3+r, 2q-3+s, c]          (same [7][7][Cin][64] layout as the forward)
// Persistent CTAs walk 8x32-pixel output tiles: the dy tile (16-bit) and the fp32 input patch are staged in shared memory;
// thread (r, o) keeps the 7*Cin gradients of filter row r / output channel o in registers: per pixel one dy value and one
// contiguous 7*Cin-float input segment (a warp shares r, so the segment loads are broadcasts).  Each CTA writes its
// partial gradient once; stem_wgrad_reduce_kernel adds the rows into dw.  (K = pixels with 3 input channels does not map
// onto the im2col TMA path: the inner dimension is 12 bytes.)
constexpr int SW_TP = 8, SW_TQ = 32, SW_THREADS = 448, SW_ROWS = 2 * SW_TP + 5, SW_COLS = 2 * SW_TQ + 5;
template <int CIN> struct SwLayout {
    static constexpr int kRowF = ((SW_COLS * CIN + 3) / 4) * 4 + (CIN == 3 ? 0 : 0);     // 208 (Cin 3) / 416 (Cin 6) floats
    static constexpr int kDyBytes = SW_TP * SW_TQ * 64 * 2;
    static constexpr int kXBytes = SW_ROWS * kRowF * 4;
    static constexpr int kTotal = kDyBytes + kXBytes;
};

template <int CIN, bool F16>
__global__ void __launch_bounds__(SW_THREADS, CIN == 3 ? 2 : 1)
stem_wgrad_kernel(const float* __restrict__ x, const uint16_t* __restrict__ dy, float* __restrict__ partials, int B, int H, int W,
                  int P, int Q) {
    using L = SwLayout<CIN>;
    constexpr int NJ = 7 * CIN;                    // gradients per thread
    constexpr int VEC = CIN == 3 ? 2 : 4;          // floats per shared-memory load of the input segment
    constexpr int NV = (NJ + VEC - 1) / VEC;
    extern __shared__ __align__(16) uint8_t sw_smem[];
    uint16_t* s_dy = reinterpret_cast<uint16_t*>(sw_smem);
    float* s_x = reinterpret_cast<float*>(sw_smem + L::kDyBytes);
    const int tid = threadIdx.x;
    const int r = tid >> 6, o = tid & 63;
    const int tiles_q = (Q + SW_TQ - 1) / SW_TQ, tiles_p = (P + SW_TP - 1) / SW_TP;
    const int n_tiles = B * tiles_p * tiles_q;
    float acc[NV * VEC];
#pragma unroll
    for (int j = 0; j < NV * VEC; ++j) acc[j] = 0.0f;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b = tile / (tiles_p * tiles_q);
        const int rem = tile - b * tiles_p * tiles_q;
        const int p0 = (rem / tiles_q) * SW_TP, q0 = (rem % tiles_q) * SW_TQ;
        __syncthreads();                           // the previous tile has been consumed
        for (int i = tid; i < SW_TP * SW_TQ * 8; i += SW_THREADS) {
            const int px = i >> 3, part = i & 7;
            const int p = p0 + (px >> 5), q = q0 + (px & 31);
            uint4 v = make_uint4(0, 0, 0, 0);
            if (p < P && q < Q) v = ld_stream(dy + (((long long)b * P + p) * Q + q) * 64 + part * 8);
            reinterpret_cast<uint4*>(s_dy)[i] = v;
        }
        const int h0 = 2 * p0 - 3, w0 = 2 * q0 - 3;
        for (int i = tid; i < SW_ROWS * L::kRowF; i += SW_THREADS) {
            const int rr = i / L::kRowF, j = i - rr * L::kRowF;
            const int h = h0 + rr, w = w0 + j / CIN;
            float v = 0.0f;
            if (j < SW_COLS * CIN && h >= 0 && h < H && w >= 0 && w < W) v = __ldg(x + ((long long)b * H + h) * W * CIN + (long long)w0 * CIN + j);
            s_x[i] = v;
        }
        __syncthreads();
#pragma unroll 2
        for (int px = 0; px < SW_TP * SW_TQ; ++px) {
            const int pl = px >> 5, ql = px & 31;
            const uint16_t raw = s_dy[px * 64 + o];
            const float d = F16 ? __half2float(__ushort_as_half(raw)) : __bfloat162float(__ushort_as_bfloat16(raw));
            const float* seg = s_x + (2 * pl + r) * L::kRowF + 2 * ql * CIN;
            if constexpr (VEC == 2) {
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const float2 f = *reinterpret_cast<const float2*>(seg + 2 * v);
                    acc[2 * v] = fmaf(d, f.x, acc[2 * v]);
                    acc[2 * v + 1] = fmaf(d, f.y, acc[2 * v + 1]);
                }
            } else {
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const float4 f = *reinterpret_cast<const float4*>(seg + 4 * v);
                    acc[4 * v] = fmaf(d, f.x, acc[4 * v]);
                    acc[4 * v + 1] = fmaf(d, f.y, acc[4 * v + 1]);
                    acc[4 * v + 2] = fmaf(d, f.z, acc[4 * v + 2]);
                    acc[4 * v + 3] = fmaf(d, f.w, acc[4 * v + 3]);
                }
            }
        }
    }
    float* row = partials + (size_t)blockIdx.x * 49 * CIN * 64;
#pragma unroll
    for (int j = 0; j < NJ; ++j) row[(r * NJ + j) * 64 + o] = acc[j];
}

// out[col] += sum over rows of partials[row][col]; blockDim = (32, 8)
__global__ void rows_sum_add_kernel(const float* __restrict__ partials, int nparts, int n, float* __restrict__ out) {
    const int col = blockIdx.x * 32 + threadIdx.x;
    const float tot = column_sum(partials, nparts, n, col);
    if (col < n && threadIdx.y == 0) out[col] += tot;
}

int grid_for(long long n, int sms) {
    long long blocks = (n + TB - 1) / TB;
    const long long cap = (long long)sms * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace
}  // namespace dpft

using namespace dpft;

#define DPFT_REQUIRE_16BIT(what)                                                                    \
    DPFT_REQUIRE(dtype == DPFT_BF16 || dtype == DPFT_F16, what ": dtype must be DPFT_BF16 or DPFT_F16"); \
    const bool is_f16 = dtype == DPFT_F16

static inline bool channels_ok(int C) { return C >= 8 && C % 8 == 0 && C / 8 <= TB && TB % (C / 8) == 0; }

static inline int reduce_grid(long long nvec, int per_thread) {
    long long blocks = (nvec + (long long)RB * per_thread - 1) / ((long long)RB * per_thread);
    const long long cap = DPFT_BN_MAX_PARTS < 2 * sm_count() ? DPFT_BN_MAX_PARTS : 2 * sm_count();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

extern "C" int dpft_bn_forward_stats(const void* y, float* workspace, const float* gamma, const float* beta, float* running_mean,
                                     float* running_var, float momentum, float eps, long long M, int C, float* scale, float* shift,
                                     float* mean, float* invstd, int dtype, void* stream) {
    DPFT_REQUIRE_16BIT("bn_forward_stats");
    DPFT_REQUIRE(y && workspace && scale && shift && mean && invstd && M > 0, "bn_forward_stats: null pointer or empty input");
    DPFT_REQUIRE(channels_ok(C), "bn_forward_stats: C=%d must be a multiple of 8 with 256 %% (C/8) == 0", C);
    const long long nvec = M * (C / 8);
    const int grid = reduce_grid(nvec, 8);
    if (is_f16) bn_stats_kernel<true><<<grid, RB, 0, (cudaStream_t)stream>>>((const uint4*)y, workspace, nvec, C / 8);
    else bn_stats_kernel<false><<<grid, RB, 0, (cudaStream_t)stream>>>((const uint4*)y, workspace, nvec, C / 8);
    DPFT_LAUNCH_CHECK("bn_stats_kernel");
    bn_finalize_kernel<<<(C + 31) / 32, dim3(32, CS_ROWS), 0, (cudaStream_t)stream>>>(workspace, grid, gamma, beta, running_mean, running_var,
                                                                               momentum, eps, (float)M, C, scale, shift, mean, invstd);
    DPFT_LAUNCH_CHECK("bn_finalize_kernel");
    return DPFT_OK;
}

extern "C" int dpft_bn_apply(const void* y, const float* scale, const float* shift, const void* residual, void* z, long long M, int C,
                             int relu, int dtype, void* stream) {
    DPFT_REQUIRE_16BIT("bn_apply");
    DPFT_REQUIRE(y && scale && shift && z && M > 0, "bn_apply: null pointer or empty input");
    DPFT_REQUIRE(channels_ok(C), "bn_apply: C=%d must be a multiple of 8 with 256 %% (C/8) == 0", C);
    const long long nvec = M * (C / 8);
    const int grid = grid_for((nvec + 1) / 2, sm_count());
    if (is_f16) bn_apply_kernel<true><<<grid, TB, 0, (cudaStream_t)stream>>>((const uint4*)y, scale, shift, (const uint4*)residual, (uint4*)z, nvec, C / 8, relu);
    else bn_apply_kernel<false><<<grid, TB, 0, (cudaStream_t)stream>>>((const uint4*)y, scale, shift, (const uint4*)residual, (uint4*)z, nvec, C / 8, relu);
    DPFT_LAUNCH_CHECK("bn_apply_kernel");
    return DPFT_OK;
}

extern "C" int dpft_bn_backward_reduce(const void* dz, const void* z, const void* y, const float* mean, const float* invstd,
                                       float* workspace, float* sum_g, float* sum_gx, float* dgamma, float* dbeta, long long M, int C,
                                       int relu, int dtype, void* stream) {
    DPFT_REQUIRE_16BIT("bn_backward_reduce");
    DPFT_REQUIRE(dz && y && mean && invstd && workspace && sum_g && sum_gx && M > 0 && (!relu || z),
                 "bn_backward_reduce: null pointer or empty input");
    DPFT_REQUIRE(channels_ok(C), "bn_backward_reduce: C=%d must be a multiple of 8 with 256 %% (C/8) == 0", C);
    const long long nvec = M * (C / 8);
    const int grid = reduce_grid(nvec, 4);
    if (is_f16) bn_bwd_reduce_kernel<true><<<grid, RB, 0, (cudaStream_t)stream>>>((const uint4*)dz, (const uint4*)z, (const uint4*)y, mean, invstd, workspace, nvec, C / 8, relu);
    else bn_bwd_reduce_kernel<false><<<grid, RB, 0, (cudaStream_t)stream>>>((const uint4*)dz, (const uint4*)z, (const uint4*)y, mean, invstd, workspace, nvec, C / 8, relu);
    DPFT_LAUNCH_CHECK("bn_bwd_reduce_kernel");
    bn_bwd_sums_kernel<<<(C + 31) / 32, dim3(32, CS_ROWS), 0, (cudaStream_t)stream>>>(workspace, grid, C, sum_g, sum_gx, dgamma, dbeta);
    DPFT_LAUNCH_CHECK("bn_bwd_sums_kernel");
    return DPFT_OK;
}

extern "C" int dpft_bn_backward_apply(const void* dz, const void* z, const void* y, const float* mean, const float* invstd,
                                      const float* gamma, const float* sum_g, const float* sum_gx, void* dy, void* g_out,
                                      long long M, int C, int relu, int dtype, void* stream) {
    DPFT_REQUIRE_16BIT("bn_backward_apply");
    DPFT_REQUIRE(dz && y && mean && invstd && sum_g && sum_gx && dy && M > 0 && (!relu || z), "bn_backward_apply: null pointer or empty input");
    DPFT_REQUIRE(channels_ok(C), "bn_backward_apply: C=%d must be a multiple of 8 with 256 %% (C/8) == 0", C);
    const long long nvec = M * (C / 8);
    const int grid = grid_for((nvec + 1) / 2, sm_count());
    const float inv = 1.0f / (float)M;
    if (is_f16) bn_bwd_apply_kernel<true><<<grid, TB, 0, (cudaStream_t)stream>>>((const uint4*)dz, (const uint4*)z, (const uint4*)y, mean, invstd, gamma, sum_g, sum_gx, (uint4*)dy, (uint4*)g_out, inv, nvec, C / 8, relu);
    else bn_bwd_apply_kernel<false><<<grid, TB, 0, (cudaStream_t)stream>>>((const uint4*)dz, (const uint4*)z, (const uint4*)y, mean, invstd, gamma, sum_g, sum_gx, (uint4*)dy, (uint4*)g_out, inv, nvec, C / 8, relu);
    DPFT_LAUNCH_CHECK("bn_bwd_apply_kernel");
    return DPFT_OK;
}

extern "C" int dpft_maxpool3x3s2_backward(const void* x, const void* pooled, const void* dy, void* dx, int B, int H, int W, int C,
                                          int dtype, void* stream) {
    DPFT_REQUIRE_16BIT("maxpool_backward");
    DPFT_REQUIRE(x && pooled && dy && dx && B > 0 && H > 0 && W > 0 && C % 8 == 0, "maxpool_backward: bad arguments");
    const int P = (H - 1) / 2 + 1, Q = (W - 1) / 2 + 1;
    const long long total = (long long)B * H * W * (C / 8);
    const int grid = grid_for(total, sm_count());
    if (is_f16) maxpool_bwd_kernel<true><<<grid, TB, 0, (cudaStream_t)stream>>>((const uint4*)x, (const uint4*)pooled, (const uint4*)dy, (uint4*)dx, B, H, W, C / 8, P, Q);
    else maxpool_bwd_kernel<false><<<grid, TB, 0, (cudaStream_t)stream>>>((const uint4*)x, (const uint4*)pooled, (const uint4*)dy, (uint4*)dx, B, H, W, C / 8, P, Q);
    DPFT_LAUNCH_CHECK("maxpool_bwd_kernel");
    return DPFT_OK;
}

extern "C" int dpft_zero_insert2_nhwc(const void* src, void* up, int B, int H, int W, int C, int P, int Q, void* stream) {
    DPFT_REQUIRE(src && up && B > 0 && H > 0 && W > 0 && C % 8 == 0 && P > 0 && Q > 0, "zero_insert2: bad arguments");
    DPFT_REQUIRE(2 * (P - 1) < H && 2 * (Q - 1) < W, "zero_insert2: (P, Q) = (%d, %d) does not fit (H, W) = (%d, %d)", P, Q, H, W);
    const long long total = (long long)B * H * W * (C / 8);
    zero_insert2_kernel<<<grid_for(total, sm_count()), TB, 0, (cudaStream_t)stream>>>((const uint4*)src, (uint4*)up, B, H, W, C / 8, P, Q);
    DPFT_LAUNCH_CHECK("zero_insert2_kernel");
    return DPFT_OK;
}

extern "C" int dpft_pack_conv_weights(const void* table, int n_layers, long long total, int dtype, void* stream) {
    DPFT_REQUIRE_16BIT("pack_conv_weights");
    DPFT_REQUIRE(table && n_layers > 0 && total > 0, "pack_conv_weights: bad arguments");
    static_assert(sizeof(PackEntry) == 48, "PackEntry layout is part of the ABI (dpft_pack_entry)");
    const int grid = grid_for(total, sm_count());
    if (is_f16) pack_weights_kernel<true><<<grid, TB, 0, (cudaStream_t)stream>>>((const PackEntry*)table, n_layers, total);
    else pack_weights_kernel<false><<<grid, TB, 0, (cudaStream_t)stream>>>((const PackEntry*)table, n_layers, total);
    DPFT_LAUNCH_CHECK("pack_weights_kernel");
    return DPFT_OK;
}

extern "C" int dpft_unpack_conv_wgrads(const void* table, int n_layers, long long total, void* stream) {
    DPFT_REQUIRE(table && n_layers > 0 && total > 0, "unpack_conv_wgrads: bad arguments");
    unpack_wgrad_kernel<<<grid_for(total, sm_count()), TB, 0, (cudaStream_t)stream>>>((const PackEntry*)table, n_layers, total);
    DPFT_LAUNCH_CHECK("unpack_wgrad_kernel");
    return DPFT_OK;
}

extern "C" int dpft_stem_conv7x7_wgrad(const float* x, const void* dy, float* workspace, long long workspace_floats, float* dw, int B,
                                       int H, int W, int Cin, int dtype, void* stream) {
    DPFT_REQUIRE_16BIT("stem_wgrad");
    DPFT_REQUIRE(x && dy && workspace && dw && B > 0 && H > 0 && W > 0, "stem_wgrad: bad arguments");
    DPFT_REQUIRE(Cin == 3 || Cin == 6, "stem_wgrad: Cin=%d (3 or 6 supported)", Cin);
    const int P = (H - 1) / 2 + 1, Q = (W - 1) / 2 + 1;
    const int n = 49 * Cin * 64;
    const int tiles = B * ((P + SW_TP - 1) / SW_TP) * ((Q + SW_TQ - 1) / SW_TQ);
    int grid = sm_count() * (Cin == 3 ? 2 : 1);
    if (grid > tiles) grid = tiles;
    DPFT_REQUIRE(workspace_floats >= (long long)grid * n, "stem_wgrad: workspace of %lld floats, need %lld", workspace_floats, (long long)grid * n);
    cudaStream_t s = (cudaStream_t)stream;
#define DPFT_SW_LAUNCH(CIN, F16)                                                                                              \
    do {                                                                                                                      \
        auto kern = stem_wgrad_kernel<CIN, F16>;                                                                              \
        static bool configured = false;                                                                                       \
        if (!configured) {                                                                                                    \
            int st = cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SwLayout<CIN>::kTotal), \
                                 "cudaFuncSetAttribute(stem_wgrad_kernel)");                                                  \
            if (st) return st;                                                                                                \
            configured = true;                                                                                                \
        }                                                                                                                     \
        kern<<<grid, SW_THREADS, SwLayout<CIN>::kTotal, s>>>(x, (const uint16_t*)dy, workspace, B, H, W, P, Q);               \
    } while (0)
    if (Cin == 3) { if (is_f16) DPFT_SW_LAUNCH(3, true); else DPFT_SW_LAUNCH(3, false); }
    else { if (is_f16) DPFT_SW_LAUNCH(6, true); else DPFT_SW_LAUNCH(6, false); }
#undef DPFT_SW_LAUNCH
    DPFT_LAUNCH_CHECK("stem_wgrad_kernel");
    rows_sum_add_kernel<<<(n + 31) / 32, dim3(32, CS_ROWS), 0, s>>>(workspace, grid, n, dw);
    DPFT_LAUNCH_CHECK("rows_sum_add_kernel");
    return DPFT_OK;
}
