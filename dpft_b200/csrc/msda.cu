// Multi-scale deformable attention for sm_100a: the bilinear gather across L feature levels with the
// head-weighted reduce (forward) and the scatter / location / weight gradients (backward).
//
// Replaces the external op DPFT calls at src/dprt/models/layers/ms_deform_attn.py:32-39 and :58-66.
//
// Mapping.  A "group" is one (b, q, m) output row of D channels.  Each group is served by
// CH * SPLIT consecutive lanes of one warp:
//   CH    = D / VEC lanes each own VEC contiguous channels, so one corner fetch of a group is one
//           coalesced CH*VEC*sizeof(T)-byte request (16 B per lane where D allows);
//   SPLIT = 1..32/CH lanes stripe the L*P samples (i = split, split+SPLIT, ...) to add parallelism
//           when B*N*M is small (the shipped D=2 configuration is latency-bound), then the partial
//           sums are combined with a xor-shuffle butterfly.
// The op is HBM/L2-gather bound: there is no reuse to stage in shared memory (only the L level shapes),
// loads are predicated rather than branched so the four corner requests of a sample issue back to back,
// and all index arithmetic is 32-bit inside a level.
#include "common.cuh"

namespace dpft {
namespace {

constexpr int kMaxLevels = 16;
constexpr int kThreads = 256;

struct LevelTable {
    int h[kMaxLevels];
    int w[kMaxLevels];
    long long start[kMaxLevels];
};

__device__ __forceinline__ void load_levels(LevelTable& t, const int64_t* __restrict__ shapes,
                                            const int64_t* __restrict__ lsi, int L) {
    if (threadIdx.x < L) {
        t.h[threadIdx.x] = (int)shapes[2 * threadIdx.x];
        t.w[threadIdx.x] = (int)shapes[2 * threadIdx.x + 1];
        t.start[threadIdx.x] = (long long)lsi[threadIdx.x];
    }
    __syncthreads();
}

// Bilinear footprint of one sample: four corner element offsets (relative to the level start, in units
// of one spatial position), weights and validity.  Follows the published op: pixel centres at +0.5,
// zero padding, corners bounds-checked one by one, nothing sampled unless -1 < x < W and -1 < y < H.
template <typename A> struct Footprint {
    int o00, o01, o10, o11;
    bool k00, k01, k10, k11;
    A lh, lw, hh, hw;
};

template <typename A>
__device__ __forceinline__ Footprint<A> footprint(A lx, A ly, int H, int W) {
    Footprint<A> f;
    const A w_im = lx * (A)W - (A)0.5;
    const A h_im = ly * (A)H - (A)0.5;
    const bool inside = (h_im > (A)-1) && (w_im > (A)-1) && (h_im < (A)H) && (w_im < (A)W);
    const A hf = floor(h_im), wf = floor(w_im);
    const int h0 = inside ? (int)hf : 0, w0 = inside ? (int)wf : 0;
    const int h1 = h0 + 1, w1 = w0 + 1;
    // outside the map every weight is zero (also keeps inf/NaN locations from poisoning the sum)
    f.lh = inside ? h_im - hf : (A)0; f.lw = inside ? w_im - wf : (A)0;
    f.hh = inside ? (A)1 - f.lh : (A)0; f.hw = inside ? (A)1 - f.lw : (A)0;
    const bool hok0 = h0 >= 0, hok1 = h1 <= H - 1, wok0 = w0 >= 0, wok1 = w1 <= W - 1;
    f.k00 = inside && hok0 && wok0; f.k01 = inside && hok0 && wok1;
    f.k10 = inside && hok1 && wok0; f.k11 = inside && hok1 && wok1;
    f.o00 = h0 * W + w0; f.o01 = f.o00 + 1; f.o10 = f.o00 + W; f.o11 = f.o10 + 1;
    return f;
}

// ------------------------------------------------------------------------------------------------ forward
template <typename T, int D, int VEC>
__global__ void __launch_bounds__(kThreads)
msda_fwd_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                const T* __restrict__ loc, const T* __restrict__ attn, T* __restrict__ out,
                int S, int M, int N, int L, int P, int split_log2, long long n_groups) {
    using A = typename AccOf<T>::type;
    constexpr int CH = D / VEC;
    __shared__ LevelTable lv;
    load_levels(lv, shapes, lsi, L);

    const int lanes_log2 = ilog2_floor(CH) + split_log2;
    const long long gtid = (long long)blockIdx.x * kThreads + threadIdx.x;
    long long group = gtid >> lanes_log2;
    const int sub = (int)(gtid & ((1 << lanes_log2) - 1));
    const int chunk = sub & (CH - 1);
    const int sp = sub >> ilog2_floor(CH);
    const int split = 1 << split_log2;
    const bool active = group < n_groups;
    if (!active) group = n_groups - 1;  // keep the lane alive for the shuffles; it never stores

    const int m = (int)(group % M);
    const long long b = group / ((long long)M * N);
    const int MD = M * D;
    const T* vb = value + (b * S) * MD + m * D + chunk * VEC;
    const int LP = L * P;
    const T* locg = loc + group * LP * 2;
    const T* attg = attn + group * LP;

    A acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = (A)0;

    // The op is bound by memory-level parallelism, and the best issue pattern differs by element size (measured on B200,
    // BASELINE config 5, D = 32): 32/64-bit values gain from two samples per trip (eight 16-byte corner requests issued back
    // to back: 0.76 -> 0.95 of the HBM roofline), while for 16-bit values the unpack arithmetic that the scheduler
    // interleaves with the loads and the extra registers cost more than they win (0.61 -> 0.48), so they keep one sample per
    // trip with the next sample's (location, weight) words prefetched.
    Pack<T, VEC> z;
#pragma unroll
    for (int k = 0; k < VEC; ++k) z.v[k] = from_acc<T>((A)0);
    if constexpr (sizeof(T) >= 4) {
        for (int i = sp; i < LP; i += 2 * split) {
            const int i2 = i + split;
            const bool two = i2 < LP;
            const int j2 = two ? i2 : i;
            const int l1 = i / P, l2 = j2 / P;
            const Pack<T, 2> xy1 = ldg_pack<T, 2>(locg + 2 * i);
            const Pack<T, 2> xy2 = ldg_pack<T, 2>(locg + 2 * j2);
            const A a1 = to_acc<T>(__ldg(attg + i));
            const A a2 = two ? to_acc<T>(__ldg(attg + j2)) : (A)0;
            const Footprint<A> f1 = footprint<A>(to_acc<T>(xy1.v[0]), to_acc<T>(xy1.v[1]), lv.h[l1], lv.w[l1]);
            Footprint<A> f2 = footprint<A>(to_acc<T>(xy2.v[0]), to_acc<T>(xy2.v[1]), lv.h[l2], lv.w[l2]);
            f2.k00 = f2.k00 && two; f2.k01 = f2.k01 && two; f2.k10 = f2.k10 && two; f2.k11 = f2.k11 && two;
            const T* v1 = vb + lv.start[l1] * MD;
            const T* v2 = vb + lv.start[l2] * MD;
            const Pack<T, VEC> p00 = f1.k00 ? ldg_pack<T, VEC>(v1 + (long long)f1.o00 * MD) : z;
            const Pack<T, VEC> p01 = f1.k01 ? ldg_pack<T, VEC>(v1 + (long long)f1.o01 * MD) : z;
            const Pack<T, VEC> p10 = f1.k10 ? ldg_pack<T, VEC>(v1 + (long long)f1.o10 * MD) : z;
            const Pack<T, VEC> p11 = f1.k11 ? ldg_pack<T, VEC>(v1 + (long long)f1.o11 * MD) : z;
            const Pack<T, VEC> r00 = f2.k00 ? ldg_pack<T, VEC>(v2 + (long long)f2.o00 * MD) : z;
            const Pack<T, VEC> r01 = f2.k01 ? ldg_pack<T, VEC>(v2 + (long long)f2.o01 * MD) : z;
            const Pack<T, VEC> r10 = f2.k10 ? ldg_pack<T, VEC>(v2 + (long long)f2.o10 * MD) : z;
            const Pack<T, VEC> r11 = f2.k11 ? ldg_pack<T, VEC>(v2 + (long long)f2.o11 * MD) : z;
            {
                const A w00 = f1.hh * f1.hw, w01 = f1.hh * f1.lw, w10 = f1.lh * f1.hw, w11 = f1.lh * f1.lw;
#pragma unroll
                for (int k = 0; k < VEC; ++k)
                    acc[k] += a1 * (w00 * to_acc<T>(p00.v[k]) + w01 * to_acc<T>(p01.v[k]) +
                                    w10 * to_acc<T>(p10.v[k]) + w11 * to_acc<T>(p11.v[k]));
            }
            {
                const A w00 = f2.hh * f2.hw, w01 = f2.hh * f2.lw, w10 = f2.lh * f2.hw, w11 = f2.lh * f2.lw;
#pragma unroll
                for (int k = 0; k < VEC; ++k)
                    acc[k] += a2 * (w00 * to_acc<T>(r00.v[k]) + w01 * to_acc<T>(r01.v[k]) +
                                    w10 * to_acc<T>(r10.v[k]) + w11 * to_acc<T>(r11.v[k]));
            }
        }
    } else {
        Pack<T, 2> xy_next = ldg_pack<T, 2>(locg + 2 * (sp < LP ? sp : 0));
        T a_next = __ldg(attg + (sp < LP ? sp : 0));
        for (int i = sp; i < LP; i += split) {
            const int l = i / P;
            const Pack<T, 2> xy = xy_next;
            const A a = to_acc<T>(a_next);
            if (i + split < LP) {
                xy_next = ldg_pack<T, 2>(locg + 2 * (i + split));
                a_next = __ldg(attg + i + split);
            }
            const int H = lv.h[l], W = lv.w[l];
            const Footprint<A> f = footprint<A>(to_acc<T>(xy.v[0]), to_acc<T>(xy.v[1]), H, W);
            const T* vl = vb + lv.start[l] * MD;
            const Pack<T, VEC> v00 = f.k00 ? ldg_pack<T, VEC>(vl + (long long)f.o00 * MD) : z;
            const Pack<T, VEC> v01 = f.k01 ? ldg_pack<T, VEC>(vl + (long long)f.o01 * MD) : z;
            const Pack<T, VEC> v10 = f.k10 ? ldg_pack<T, VEC>(vl + (long long)f.o10 * MD) : z;
            const Pack<T, VEC> v11 = f.k11 ? ldg_pack<T, VEC>(vl + (long long)f.o11 * MD) : z;
            const A w00 = f.hh * f.hw, w01 = f.hh * f.lw, w10 = f.lh * f.hw, w11 = f.lh * f.lw;
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                acc[k] += a * (w00 * to_acc<T>(v00.v[k]) + w01 * to_acc<T>(v01.v[k]) +
                               w10 * to_acc<T>(v10.v[k]) + w11 * to_acc<T>(v11.v[k]));
            }
        }
    }
    // head-weighted reduce across the SPLIT lanes of the group
    for (int o = CH; o < (1 << lanes_log2); o <<= 1) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    }
    if (active && sp == 0) {
        Pack<T, VEC> r;
#pragma unroll
        for (int k = 0; k < VEC; ++k) r.v[k] = from_acc<T>(acc[k]);
        st_pack<T, VEC>(out + group * D + chunk * VEC, r);
    }
}

// Forward variant for 16-byte corner slices of 16-bit values: the four corner rows of each sample are fetched with
// cp.async (LDGSTS, L2 -> shared memory, zero-fill for out-of-bounds corners) into a per-thread ring of RING slots, so
// RING samples x 64 B per thread are in flight without holding them in registers, and the unpack / FMA work of sample n
// overlaps the requests of samples n+1 .. n+RING-1.  No cross-thread communication: each thread only reads what it wrote.
constexpr int kRing = 4;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int bytes = valid ? 16 : 0;               // src-size 0: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T, int D, int VEC>
__global__ void __launch_bounds__(kThreads)
msda_fwd_ring_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                     const T* __restrict__ loc, const T* __restrict__ attn, T* __restrict__ out,
                     int S, int M, int N, int L, int P, int split_log2, long long n_groups) {
    static_assert(sizeof(T) * VEC == 16, "ring variant moves 16-byte corner slices");
    using A = typename AccOf<T>::type;
    constexpr int CH = D / VEC;
    extern __shared__ __align__(16) uint8_t ring_smem[];     // [kRing][4 corners][kThreads] x 16 B
    __shared__ LevelTable lv;
    load_levels(lv, shapes, lsi, L);

    const int lanes_log2 = ilog2_floor(CH) + split_log2;
    const long long gtid = (long long)blockIdx.x * kThreads + threadIdx.x;
    long long group = gtid >> lanes_log2;
    const int sub = (int)(gtid & ((1 << lanes_log2) - 1));
    const int chunk = sub & (CH - 1);
    const int sp = sub >> ilog2_floor(CH);
    const int split = 1 << split_log2;
    const bool active = group < n_groups;
    if (!active) group = n_groups - 1;

    const int m = (int)(group % M);
    const long long b = group / ((long long)M * N);
    const int MD = M * D;
    const T* vb = value + (b * S) * MD + m * D + chunk * VEC;
    const int LP = L * P;
    const T* locg = loc + group * LP * 2;
    const T* attg = attn + group * LP;
    const uint32_t slot0 = (uint32_t)__cvta_generic_to_shared(ring_smem) + threadIdx.x * 16;
    const int count = sp < LP ? (LP - sp + split - 1) >> split_log2 : 0;

    A acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = (A)0;
    A wgt[kRing][4];                                  // attention * bilinear weights of the samples in flight

    auto issue = [&](int n, int slot) {
        if (n < count) {
            const int i = sp + n * split;
            const int l = i / P;
            const Pack<T, 2> xy = ldg_pack<T, 2>(locg + 2 * i);
            const A a = to_acc<T>(__ldg(attg + i));
            const Footprint<A> f = footprint<A>(to_acc<T>(xy.v[0]), to_acc<T>(xy.v[1]), lv.h[l], lv.w[l]);
            const T* vl = vb + lv.start[l] * MD;
            const uint32_t dst = slot0 + slot * (4 * kThreads * 16);
            cp_async16(dst, vl + (long long)(f.k00 ? f.o00 : 0) * MD, f.k00);
            cp_async16(dst + kThreads * 16, vl + (long long)(f.k01 ? f.o01 : 0) * MD, f.k01);
            cp_async16(dst + 2 * kThreads * 16, vl + (long long)(f.k10 ? f.o10 : 0) * MD, f.k10);
            cp_async16(dst + 3 * kThreads * 16, vl + (long long)(f.k11 ? f.o11 : 0) * MD, f.k11);
            wgt[slot][0] = a * (f.hh * f.hw); wgt[slot][1] = a * (f.hh * f.lw);
            wgt[slot][2] = a * (f.lh * f.hw); wgt[slot][3] = a * (f.lh * f.lw);
        }
        cp_async_commit();                            // always commit so the group count stays uniform
    };

#pragma unroll
    for (int s = 0; s < kRing - 1; ++s) issue(s, s);
    for (int base = 0; base < count; base += kRing) {
#pragma unroll
        for (int s = 0; s < kRing; ++s) {
            const int n = base + s;
            issue(n + kRing - 1, (s + kRing - 1) % kRing);
            cp_async_wait<kRing - 1>();
            if (n < count) {
                const uint8_t* src = ring_smem + threadIdx.x * 16 + s * (4 * kThreads * 16);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const Pack<T, VEC> v = *reinterpret_cast<const Pack<T, VEC>*>(src + c * kThreads * 16);
#pragma unroll
                    for (int k = 0; k < VEC; ++k) acc[k] = fma(wgt[s][c], to_acc<T>(v.v[k]), acc[k]);
                }
            }
        }
    }
    for (int o = CH; o < (1 << lanes_log2); o <<= 1) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    }
    if (active && sp == 0) {
        Pack<T, VEC> r;
#pragma unroll
        for (int k = 0; k < VEC; ++k) r.v[k] = from_acc<T>(acc[k]);
        st_pack<T, VEC>(out + group * D + chunk * VEC, r);
    }
}

// Any D (not a power of two, or > 32 lanes' worth): one thread per output element.
template <typename T>
__global__ void __launch_bounds__(kThreads)
msda_fwd_generic_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                        const int64_t* __restrict__ lsi, const T* __restrict__ loc, const T* __restrict__ attn,
                        T* __restrict__ out, int S, int M, int D, int N, int L, int P, long long n_elems) {
    using A = typename AccOf<T>::type;
    __shared__ LevelTable lv;
    load_levels(lv, shapes, lsi, L);
    const long long e = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (e >= n_elems) return;
    const int d = (int)(e % D);
    const long long group = e / D;
    const int m = (int)(group % M);
    const long long b = group / ((long long)M * N);
    const int MD = M * D;
    const T* vb = value + (b * S) * MD + m * D + d;
    const int LP = L * P;
    A acc = (A)0;
    for (int i = 0; i < LP; ++i) {
        const int l = i / P;
        const A lx = to_acc<T>(loc[(group * LP + i) * 2]), ly = to_acc<T>(loc[(group * LP + i) * 2 + 1]);
        const A a = to_acc<T>(attn[group * LP + i]);
        const Footprint<A> f = footprint<A>(lx, ly, lv.h[l], lv.w[l]);
        const T* vl = vb + lv.start[l] * MD;
        const A v00 = f.k00 ? to_acc<T>(vl[(long long)f.o00 * MD]) : (A)0;
        const A v01 = f.k01 ? to_acc<T>(vl[(long long)f.o01 * MD]) : (A)0;
        const A v10 = f.k10 ? to_acc<T>(vl[(long long)f.o10 * MD]) : (A)0;
        const A v11 = f.k11 ? to_acc<T>(vl[(long long)f.o11 * MD]) : (A)0;
        acc += a * (f.hh * f.hw * v00 + f.hh * f.lw * v01 + f.lh * f.hw * v10 + f.lh * f.lw * v11);
    }
    out[e] = from_acc<T>(acc);
}

// ----------------------------------------------------------------------------------------------- backward
template <typename T, int D, int VEC>
__global__ void __launch_bounds__(kThreads)
msda_bwd_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                const T* __restrict__ loc, const T* __restrict__ attn, const T* __restrict__ grad_out,
                typename AccOf<T>::type* __restrict__ grad_value, T* __restrict__ grad_loc,
                T* __restrict__ grad_attn, int S, int M, int N, int L, int P, int split_log2,
                long long n_groups) {
    using A = typename AccOf<T>::type;
    constexpr int CH = D / VEC;
    __shared__ LevelTable lv;
    load_levels(lv, shapes, lsi, L);

    const int lanes_log2 = ilog2_floor(CH) + split_log2;
    const long long gtid = (long long)blockIdx.x * kThreads + threadIdx.x;
    long long group = gtid >> lanes_log2;
    const int sub = (int)(gtid & ((1 << lanes_log2) - 1));
    const int chunk = sub & (CH - 1);
    const int sp = sub >> ilog2_floor(CH);
    const int split = 1 << split_log2;
    const bool active = group < n_groups;
    if (!active) group = n_groups - 1;

    const int m = (int)(group % M);
    const long long b = group / ((long long)M * N);
    const int MD = M * D;
    const long long voff = (b * S) * MD + m * D + chunk * VEC;
    const int LP = L * P;
    const T* locg = loc + group * LP * 2;
    const T* attg = attn + group * LP;

    A go[VEC];
    {
        const Pack<T, VEC> g = ldg_pack<T, VEC>(grad_out + group * D + chunk * VEC);
#pragma unroll
        for (int k = 0; k < VEC; ++k) go[k] = to_acc<T>(g.v[k]);
    }

    // Uniform trip count over the warp: the CH-lane shuffles below need every lane present.
    const int trips = (LP + split - 1) >> split_log2;
    Pack<T, 2> xy_next = ldg_pack<T, 2>(locg + 2 * (sp < LP ? sp : LP - 1));
    T a_next = __ldg(attg + (sp < LP ? sp : LP - 1));
    for (int it = 0; it < trips; ++it) {
        const int i = (it << split_log2) + sp;
        const bool have = active && (i < LP);
        const int ic = (i < LP) ? i : LP - 1;
        const int l = ic / P;
        const Pack<T, 2> xy = xy_next;
        const A a = to_acc<T>(a_next);
        {
            const int in = i + split < LP ? i + split : LP - 1;     // prefetch the next sample's location / weight
            xy_next = ldg_pack<T, 2>(locg + 2 * in);
            a_next = __ldg(attg + in);
        }
        const int H = lv.h[l], W = lv.w[l];
        Footprint<A> f = footprint<A>(to_acc<T>(xy.v[0]), to_acc<T>(xy.v[1]), H, W);
        f.k00 = f.k00 && have; f.k01 = f.k01 && have; f.k10 = f.k10 && have; f.k11 = f.k11 && have;
        const long long lbase = voff + lv.start[l] * MD;
        const T* vl = value + lbase;
        A* gvl = grad_value + lbase;
        Pack<T, VEC> z;
#pragma unroll
        for (int k = 0; k < VEC; ++k) z.v[k] = from_acc<T>((A)0);
        const Pack<T, VEC> v00 = f.k00 ? ldg_pack<T, VEC>(vl + (long long)f.o00 * MD) : z;
        const Pack<T, VEC> v01 = f.k01 ? ldg_pack<T, VEC>(vl + (long long)f.o01 * MD) : z;
        const Pack<T, VEC> v10 = f.k10 ? ldg_pack<T, VEC>(vl + (long long)f.o10 * MD) : z;
        const Pack<T, VEC> v11 = f.k11 ? ldg_pack<T, VEC>(vl + (long long)f.o11 * MD) : z;
        const A w00 = f.hh * f.hw, w01 = f.hh * f.lw, w10 = f.lh * f.hw, w11 = f.lh * f.lw;

        A ga = (A)0, gx = (A)0, gy = (A)0;
        A s00[VEC], s01[VEC], s10[VEC], s11[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const A a00 = to_acc<T>(v00.v[k]), a01 = to_acc<T>(v01.v[k]);
            const A a10 = to_acc<T>(v10.v[k]), a11 = to_acc<T>(v11.v[k]);
            const A ag = a * go[k];
            s00[k] = w00 * ag; s01[k] = w01 * ag; s10[k] = w10 * ag; s11[k] = w11 * ag;
            ga += go[k] * (w00 * a00 + w01 * a01 + w10 * a10 + w11 * a11);
            gx += ag * (f.hh * (a01 - a00) + f.lh * (a11 - a10));
            gy += ag * (f.hw * (a10 - a00) + f.lw * (a11 - a01));
        }
        if (f.k00) red_add<VEC>(gvl + (long long)f.o00 * MD, s00);
        if (f.k01) red_add<VEC>(gvl + (long long)f.o01 * MD, s01);
        if (f.k10) red_add<VEC>(gvl + (long long)f.o10 * MD, s10);
        if (f.k11) red_add<VEC>(gvl + (long long)f.o11 * MD, s11);
        // sum the per-channel partials over the CH lanes of the group (deterministic, no atomics)
#pragma unroll
        for (int o = 1; o < CH; o <<= 1) {
            ga += __shfl_xor_sync(0xffffffffu, ga, o);
            gx += __shfl_xor_sync(0xffffffffu, gx, o);
            gy += __shfl_xor_sync(0xffffffffu, gy, o);
        }
        if (have && chunk == 0) {
            grad_attn[group * LP + i] = from_acc<T>(ga);
            Pack<T, 2> g2;
            g2.v[0] = from_acc<T>(gx * (A)W);
            g2.v[1] = from_acc<T>(gy * (A)H);
            st_pack<T, 2>(grad_loc + (group * LP + i) * 2, g2);
        }
    }
}

// Any D: one thread per sample, looping over the channels.
template <typename T>
__global__ void __launch_bounds__(kThreads)
msda_bwd_generic_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                        const int64_t* __restrict__ lsi, const T* __restrict__ loc, const T* __restrict__ attn,
                        const T* __restrict__ grad_out, typename AccOf<T>::type* __restrict__ grad_value,
                        T* __restrict__ grad_loc, T* __restrict__ grad_attn, int S, int M, int D, int N, int L,
                        int P, long long n_samples) {
    using A = typename AccOf<T>::type;
    __shared__ LevelTable lv;
    load_levels(lv, shapes, lsi, L);
    const long long s = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (s >= n_samples) return;
    const int LP = L * P;
    const int i = (int)(s % LP);
    const long long group = s / LP;
    const int l = i / P;
    const int m = (int)(group % M);
    const long long b = group / ((long long)M * N);
    const int MD = M * D;
    const long long lbase = (b * S + lv.start[l]) * MD + m * D;
    const int H = lv.h[l], W = lv.w[l];
    const A a = to_acc<T>(attn[s]);
    const Footprint<A> f = footprint<A>(to_acc<T>(loc[2 * s]), to_acc<T>(loc[2 * s + 1]), H, W);
    const A w00 = f.hh * f.hw, w01 = f.hh * f.lw, w10 = f.lh * f.hw, w11 = f.lh * f.lw;
    A ga = (A)0, gx = (A)0, gy = (A)0;
    for (int d = 0; d < D; ++d) {
        const long long i00 = lbase + (long long)f.o00 * MD + d, i01 = lbase + (long long)f.o01 * MD + d;
        const long long i10 = lbase + (long long)f.o10 * MD + d, i11 = lbase + (long long)f.o11 * MD + d;
        const A a00 = f.k00 ? to_acc<T>(value[i00]) : (A)0, a01 = f.k01 ? to_acc<T>(value[i01]) : (A)0;
        const A a10 = f.k10 ? to_acc<T>(value[i10]) : (A)0, a11 = f.k11 ? to_acc<T>(value[i11]) : (A)0;
        const A g = to_acc<T>(grad_out[group * D + d]);
        const A ag = a * g;
        if (f.k00) atomicAdd(grad_value + i00, w00 * ag);
        if (f.k01) atomicAdd(grad_value + i01, w01 * ag);
        if (f.k10) atomicAdd(grad_value + i10, w10 * ag);
        if (f.k11) atomicAdd(grad_value + i11, w11 * ag);
        ga += g * (w00 * a00 + w01 * a01 + w10 * a10 + w11 * a11);
        gx += ag * (f.hh * (a01 - a00) + f.lh * (a11 - a10));
        gy += ag * (f.hw * (a10 - a00) + f.lw * (a11 - a01));
    }
    grad_attn[s] = from_acc<T>(ga);
    grad_loc[2 * s] = from_acc<T>(gx * (A)W);
    grad_loc[2 * s + 1] = from_acc<T>(gy * (A)H);
}

// -------------------------------------------------------------------------------------------- host dispatch
template <typename T> constexpr int vec_for(int D) {
    const int cap = 16 / (int)sizeof(T);
    return D < cap ? D : cap;
}

// SPLIT: stripe the L*P samples over more lanes until the grid can fill the 148 SMs a few times over.
inline int pick_split_log2(long long n_groups, int CH, int LP) {
    const long long want = 148LL * 2048 * 2;
    int s = 0;
    while ((n_groups * CH << s) < want && (CH << (s + 1)) <= 32 && (2 << s) <= LP) ++s;
    return s;
}

struct Args {
    const void *value, *loc, *attn, *grad_out;
    const int64_t *shapes, *lsi;
    void *out, *grad_value, *grad_loc, *grad_attn;
    int B, S, M, D, N, L, P;
    cudaStream_t stream;
};

template <typename T, int D> int launch_fwd(const Args& a) {
    constexpr int VEC = vec_for<T>(D), CH = D / VEC;
    const long long n_groups = (long long)a.B * a.N * a.M;
    const int sl = pick_split_log2(n_groups, CH, a.L * a.P);
    const long long threads = n_groups * CH << sl;
    const unsigned grid = (unsigned)((threads + kThreads - 1) / kThreads);
    if constexpr (sizeof(T) == 2 && sizeof(T) * VEC == 16) {
        constexpr int smem = kRing * 4 * kThreads * 16;
        static bool configured = false;
        if (!configured) {
            auto kern = msda_fwd_ring_kernel<T, D, VEC>;
            int st = cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "msda ring attr");
            if (st) return st;
            configured = true;
        }
        msda_fwd_ring_kernel<T, D, VEC><<<grid, kThreads, smem, a.stream>>>(
            (const T*)a.value, a.shapes, a.lsi, (const T*)a.loc, (const T*)a.attn, (T*)a.out, a.S, a.M, a.N, a.L,
            a.P, sl, n_groups);
        DPFT_LAUNCH_CHECK("msda_fwd_ring_kernel");
        return DPFT_OK;
    }
    msda_fwd_kernel<T, D, VEC><<<grid, kThreads, 0, a.stream>>>(
        (const T*)a.value, a.shapes, a.lsi, (const T*)a.loc, (const T*)a.attn, (T*)a.out, a.S, a.M, a.N, a.L,
        a.P, sl, n_groups);
    DPFT_LAUNCH_CHECK("msda_fwd_kernel");
    return DPFT_OK;
}

template <typename T, int D> int launch_bwd(const Args& a) {
    constexpr int VEC = vec_for<T>(D), CH = D / VEC;
    using A = typename AccOf<T>::type;
    const long long n_groups = (long long)a.B * a.N * a.M;
    const int sl = pick_split_log2(n_groups, CH, a.L * a.P);
    const long long threads = n_groups * CH << sl;
    const unsigned grid = (unsigned)((threads + kThreads - 1) / kThreads);
    msda_bwd_kernel<T, D, VEC><<<grid, kThreads, 0, a.stream>>>(
        (const T*)a.value, a.shapes, a.lsi, (const T*)a.loc, (const T*)a.attn, (const T*)a.grad_out,
        (A*)a.grad_value, (T*)a.grad_loc, (T*)a.grad_attn, a.S, a.M, a.N, a.L, a.P, sl, n_groups);
    DPFT_LAUNCH_CHECK("msda_bwd_kernel");
    return DPFT_OK;
}

template <typename T> int dispatch(const Args& a, bool bwd) {
    // D*sizeof(T) must keep 16-byte lanes within 32 lanes: D <= 32 * VEC.
    switch (a.D) {
#define DPFT_CASE(DD)                                                        \
    case DD:                                                                 \
        if constexpr (DD / vec_for<T>(DD) <= 32)                             \
            return bwd ? launch_bwd<T, DD>(a) : launch_fwd<T, DD>(a);        \
        break;
        DPFT_CASE(1) DPFT_CASE(2) DPFT_CASE(4) DPFT_CASE(8) DPFT_CASE(16) DPFT_CASE(32) DPFT_CASE(64) DPFT_CASE(128)
#undef DPFT_CASE
        default: break;
    }
    using A = typename AccOf<T>::type;
    if (!bwd) {
        const long long n = (long long)a.B * a.N * a.M * a.D;
        const unsigned grid = (unsigned)((n + kThreads - 1) / kThreads);
        msda_fwd_generic_kernel<T><<<grid, kThreads, 0, a.stream>>>(
            (const T*)a.value, a.shapes, a.lsi, (const T*)a.loc, (const T*)a.attn, (T*)a.out, a.S, a.M, a.D,
            a.N, a.L, a.P, n);
        DPFT_LAUNCH_CHECK("msda_fwd_generic_kernel");
    } else {
        const long long n = (long long)a.B * a.N * a.M * a.L * a.P;
        const unsigned grid = (unsigned)((n + kThreads - 1) / kThreads);
        msda_bwd_generic_kernel<T><<<grid, kThreads, 0, a.stream>>>(
            (const T*)a.value, a.shapes, a.lsi, (const T*)a.loc, (const T*)a.attn, (const T*)a.grad_out,
            (A*)a.grad_value, (T*)a.grad_loc, (T*)a.grad_attn, a.S, a.M, a.D, a.N, a.L, a.P, n);
        DPFT_LAUNCH_CHECK("msda_bwd_generic_kernel");
    }
    return DPFT_OK;
}

int run(const Args& a, int dtype, bool bwd) {
    DPFT_REQUIRE(a.B >= 0 && a.S >= 0 && a.M > 0 && a.D > 0 && a.N >= 0 && a.L > 0 && a.P > 0,
                 "msda: bad sizes B=%d S=%d M=%d D=%d N=%d L=%d P=%d", a.B, a.S, a.M, a.D, a.N, a.L, a.P);
    DPFT_REQUIRE(a.L <= kMaxLevels, "msda: L=%d exceeds the supported %d levels", a.L, kMaxLevels);
    if (a.B == 0 || a.N == 0) return DPFT_OK;  // empty batch / no queries: nothing to write
    DPFT_REQUIRE(a.value && a.shapes && a.lsi && a.loc && a.attn, "msda: null input pointer");
    if (bwd) DPFT_REQUIRE(a.grad_out && a.grad_value && a.grad_loc && a.grad_attn, "msda: null gradient pointer");
    else DPFT_REQUIRE(a.out, "msda: null output pointer");
    DPFT_REQUIRE(((uintptr_t)a.value & 15) == 0 && ((uintptr_t)a.loc & 15) == 0 &&
                 ((uintptr_t)(bwd ? a.grad_out : a.out) & 15) == 0 &&
                 (!bwd || ((uintptr_t)a.grad_value & 15) == 0),
                 "msda: value/loc/out/grad pointers must be 16-byte aligned");
    switch (dtype) {
        case DPFT_F32: return dispatch<float>(a, bwd);
        case DPFT_F64: return dispatch<double>(a, bwd);
        case DPFT_F16: return dispatch<__half>(a, bwd);
        case DPFT_BF16: return dispatch<__nv_bfloat16>(a, bwd);
        default: set_error("msda: unsupported dtype code %d", dtype); return DPFT_ERR_UNSUPPORTED;
    }
}

}  // namespace
}  // namespace dpft

extern "C" int dpft_msda_forward(const void* value, const int64_t* shapes, const int64_t* lsi, const void* loc,
                                 const void* attn, void* out, int B, int S, int M, int D, int N, int L, int P,
                                 int dtype, void* stream) {
    dpft::Args a{value, loc, attn, nullptr, shapes, lsi, out, nullptr, nullptr, nullptr,
                 B, S, M, D, N, L, P, (cudaStream_t)stream};
    return dpft::run(a, dtype, false);
}

extern "C" int dpft_msda_backward(const void* value, const int64_t* shapes, const int64_t* lsi, const void* loc,
                                  const void* attn, const void* grad_out, void* grad_value, void* grad_loc,
                                  void* grad_attn, int B, int S, int M, int D, int N, int L, int P, int dtype,
                                  void* stream) {
    dpft::Args a{value, loc, attn, grad_out, shapes, lsi, nullptr, grad_value, grad_loc, grad_attn,
                 B, S, M, D, N, L, P, (cudaStream_t)stream};
    return dpft::run(a, dtype, true);
}
