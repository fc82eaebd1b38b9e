/*
 * oracle/msda_oracle.c — CPU restatement of the multi-scale deformable attention op.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under dpft_b200/ may import, link or call this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do,
 * and there only as the checker.
 *
 * What it restates.  DPFT calls a third-party, un-vendored, un-pinned extension
 * (`import MultiScaleDeformableAttention as MSDA`, reference src/dprt/models/layers/ms_deform_attn.py:24;
 * built from fundamentalvision/Deformable-DETR `main`, reference Dockerfile:32-39).  Its source is
 * NOT in /root/reference, so this file restates the *published* algorithm of that op and anchors on
 * the reference's own call sites:
 *   - call contract and argument order: ms_deform_attn.py:27-68 (forward :32-39, backward :58-66,
 *     returned grads :68),
 *   - tensor semantics: ms_deform_attn.py:145-161 (locations normalised to [0,1], (0,0) top-left),
 *     :185-191 (last dim of the locations is (x, y): the normaliser is stacked [W, H]).
 * Published algorithm (Deformable DETR, Zhu et al. 2020, and the op's documented PyTorch equivalent):
 *   out[b,q,m,:] = sum_{l,p} A[b,q,m,l,p] * bilinear(V_l[b,:,m,:], x = loc_x*W_l - 0.5, y = loc_y*H_l - 0.5)
 * with zero padding outside the map  ==  grid_sample(..., 2*loc-1, bilinear, zeros, align_corners=False).
 *
 * PARITY STATUS: the reference holds no tests or golden vectors at this boundary, so
 * parity of this op is pinned only against (a) torch.nn.functional.grid_sample and (b) the independent
 * HuggingFace restatement shipped in this image (tests/test_oracle.py) — "parity unpinned by the
 * reference itself".
 *
 * Layouts (all contiguous, row-major):
 *   value  (B, S, M, D)       S = sum_l H_l*W_l
 *   shapes (L, 2) int64       [H_l, W_l]
 *   lsi    (L,)   int64       level start index into S
 *   loc    (B, N, M, L, P, 2) (x, y) in [0,1]-ish, may leave the range
 *   attn   (B, N, M, L, P)
 *   out    (B, N, M*D)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define DEFINE_MSDA(T, SUF)                                                                          \
                                                                                                     \
void msda_oracle_fwd_##SUF(const T* value, const int64_t* shapes, const int64_t* lsi,               \
                           const T* loc, const T* attn, T* out,                                      \
                           int B, int S, int M, int D, int N, int L, int P)                          \
{                                                                                                    \
    for (int b = 0; b < B; ++b)                                                                      \
    for (int q = 0; q < N; ++q)                                                                      \
    for (int m = 0; m < M; ++m) {                                                                    \
        T* o = out + (((int64_t)b * N + q) * M + m) * D;                                             \
        for (int d = 0; d < D; ++d) o[d] = (T)0;                                                     \
        for (int l = 0; l < L; ++l) {                                                                \
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                            \
            const T* vl = value + ((int64_t)b * S + lsi[l]) * M * D + (int64_t)m * D;                \
            for (int p = 0; p < P; ++p) {                                                            \
                const int64_t s = ((((int64_t)b * N + q) * M + m) * L + l) * P + p;                  \
                const T a = attn[s];                                                                 \
                const T w_im = loc[2 * s] * (T)W - (T)0.5;                                           \
                const T h_im = loc[2 * s + 1] * (T)H - (T)0.5;                                       \
                if (!(h_im > (T)-1 && w_im > (T)-1 && h_im < (T)H && w_im < (T)W)) continue;         \
                const int h0 = (int)floor((double)h_im), w0 = (int)floor((double)w_im);              \
                const int h1 = h0 + 1, w1 = w0 + 1;                                                  \
                const T lh = h_im - (T)h0, lw = w_im - (T)w0, hh = (T)1 - lh, hw = (T)1 - lw;        \
                for (int d = 0; d < D; ++d) {                                                        \
                    T v00 = 0, v01 = 0, v10 = 0, v11 = 0;                                            \
                    if (h0 >= 0 && w0 >= 0)         v00 = vl[((int64_t)h0 * W + w0) * M * D + d];    \
                    if (h0 >= 0 && w1 <= W - 1)     v01 = vl[((int64_t)h0 * W + w1) * M * D + d];    \
                    if (h1 <= H - 1 && w0 >= 0)     v10 = vl[((int64_t)h1 * W + w0) * M * D + d];    \
                    if (h1 <= H - 1 && w1 <= W - 1) v11 = vl[((int64_t)h1 * W + w1) * M * D + d];    \
                    o[d] += a * (hh * hw * v00 + hh * lw * v01 + lh * hw * v10 + lh * lw * v11);     \
                }                                                                                    \
            }                                                                                        \
        }                                                                                            \
    }                                                                                                \
}                                                                                                    \
                                                                                                     \
/* grad_value must be zero-filled by the caller; grad_loc and grad_attn are fully written. */        \
void msda_oracle_bwd_##SUF(const T* value, const int64_t* shapes, const int64_t* lsi,               \
                           const T* loc, const T* attn, const T* grad_out,                           \
                           T* grad_value, T* grad_loc, T* grad_attn,                                 \
                           int B, int S, int M, int D, int N, int L, int P)                          \
{                                                                                                    \
    for (int b = 0; b < B; ++b)                                                                      \
    for (int q = 0; q < N; ++q)                                                                      \
    for (int m = 0; m < M; ++m) {                                                                    \
        const T* go = grad_out + (((int64_t)b * N + q) * M + m) * D;                                 \
        for (int l = 0; l < L; ++l) {                                                                \
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                            \
            const int64_t base = ((int64_t)b * S + lsi[l]) * M * D + (int64_t)m * D;                 \
            const T* vl = value + base;                                                              \
            T* gvl = grad_value + base;                                                              \
            for (int p = 0; p < P; ++p) {                                                            \
                const int64_t s = ((((int64_t)b * N + q) * M + m) * L + l) * P + p;                  \
                const T a = attn[s];                                                                 \
                grad_attn[s] = 0; grad_loc[2 * s] = 0; grad_loc[2 * s + 1] = 0;                      \
                const T w_im = loc[2 * s] * (T)W - (T)0.5;                                           \
                const T h_im = loc[2 * s + 1] * (T)H - (T)0.5;                                       \
                if (!(h_im > (T)-1 && w_im > (T)-1 && h_im < (T)H && w_im < (T)W)) continue;         \
                const int h0 = (int)floor((double)h_im), w0 = (int)floor((double)w_im);              \
                const int h1 = h0 + 1, w1 = w0 + 1;                                                  \
                const T lh = h_im - (T)h0, lw = w_im - (T)w0, hh = (T)1 - lh, hw = (T)1 - lw;        \
                const int ok00 = (h0 >= 0 && w0 >= 0), ok01 = (h0 >= 0 && w1 <= W - 1);              \
                const int ok10 = (h1 <= H - 1 && w0 >= 0), ok11 = (h1 <= H - 1 && w1 <= W - 1);      \
                T ga = 0, gx = 0, gy = 0;                                                            \
                for (int d = 0; d < D; ++d) {                                                        \
                    const int64_t i00 = ((int64_t)h0 * W + w0) * M * D + d;                          \
                    const int64_t i01 = ((int64_t)h0 * W + w1) * M * D + d;                          \
                    const int64_t i10 = ((int64_t)h1 * W + w0) * M * D + d;                          \
                    const int64_t i11 = ((int64_t)h1 * W + w1) * M * D + d;                          \
                    const T v00 = ok00 ? vl[i00] : (T)0, v01 = ok01 ? vl[i01] : (T)0;                \
                    const T v10 = ok10 ? vl[i10] : (T)0, v11 = ok11 ? vl[i11] : (T)0;                \
                    const T g = go[d];                                                               \
                    const T ag = a * g;                                                              \
                    if (ok00) gvl[i00] += hh * hw * ag;                                              \
                    if (ok01) gvl[i01] += hh * lw * ag;                                              \
                    if (ok10) gvl[i10] += lh * hw * ag;                                              \
                    if (ok11) gvl[i11] += lh * lw * ag;                                              \
                    ga += g * (hh * hw * v00 + hh * lw * v01 + lh * hw * v10 + lh * lw * v11);       \
                    /* d sample / d w_im and / d h_im */                                             \
                    gx += ag * (hh * (v01 - v00) + lh * (v11 - v10));                                \
                    gy += ag * (hw * (v10 - v00) + lw * (v11 - v01));                                \
                }                                                                                    \
                grad_attn[s] = ga;                                                                   \
                grad_loc[2 * s] = gx * (T)W;      /* w_im = loc_x * W - 0.5 */                       \
                grad_loc[2 * s + 1] = gy * (T)H;  /* h_im = loc_y * H - 0.5 */                       \
            }                                                                                        \
        }                                                                                            \
    }                                                                                                \
}

DEFINE_MSDA(float, f32)
DEFINE_MSDA(double, f64)
