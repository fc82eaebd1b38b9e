/*
 * dpft_b200.h — C ABI of libdpft_b200.so, the sm_100a implementation of the DPFT (dprt) model hot path.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every tensor pointer is a DEVICE pointer unless the name says host;
 *   - tensors are contiguous, row-major, in the layout the cited reference call site uses;
 *   - the caller owns every buffer; nothing here allocates, frees or synchronises;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - return 0 on success, a negative DPFT_ERR_* for argument errors, a positive cudaError_t value for
 *     CUDA failures; dpft_last_error() gives the message for the calling thread;
 *   - stateless and re-entrant (the only process state is the per-thread error string).
 *
 * Citations are into /root/reference (TUMFTM/DPFT); see INTEGRATION.md for the reference-side bindings.
 */
#ifndef DPFT_B200_H
#define DPFT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPFT_ABI_VERSION 1

#if defined(__GNUC__)
#define DPFT_API __attribute__((visibility("default")))
#else
#define DPFT_API
#endif

/* element types (the `dtype` argument) */
#define DPFT_F32 0
#define DPFT_F64 1
#define DPFT_F16 2
#define DPFT_BF16 3

/* error codes */
#define DPFT_OK 0
#define DPFT_ERR_INVALID_ARGUMENT (-1)
#define DPFT_ERR_UNSUPPORTED (-2)

DPFT_API int dpft_abi_version(void);
DPFT_API const char* dpft_last_error(void);
/* Fills sm_count / cc_major / cc_minor of `device`; returns a cudaError_t value if no usable GPU. */
DPFT_API int dpft_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);

/*
 * Multi-scale deformable attention, forward.
 * Replaces MSDA.ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
 * im2col_step) — called at src/dprt/models/layers/ms_deform_attn.py:32-39 (im2col_step is a batching knob of
 * the upstream kernel and has no effect on results; it is accepted by the Python binding and ignored).
 *   value  (B, S, M, D)          dtype
 *   shapes (L, 2) int64 device   [H_l, W_l]                     (ms_deform_attn.py:153-154)
 *   lsi    (L,)   int64 device   level start index into S       (ms_deform_attn.py:155-156)
 *   loc    (B, N, M, L, P, 2)    dtype, (x, y) normalised       (ms_deform_attn.py:185-191)
 *   attn   (B, N, M, L, P)       dtype
 *   out    (B, N, M*D)           dtype, fully overwritten
 * Arithmetic: accumulate in f32 (f64 for DPFT_F64).  L <= 16.
 */
DPFT_API int dpft_msda_forward(const void* value, const int64_t* shapes, const int64_t* lsi, const void* loc,
                      const void* attn, void* out, int B, int S, int M, int D, int N, int L, int P,
                      int dtype, void* stream);

/*
 * Multi-scale deformable attention, backward.
 * Replaces MSDA.ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
 * grad_output, im2col_step) -> (grad_value, grad_sampling_loc, grad_attn_weight) — called at
 * src/dprt/models/layers/ms_deform_attn.py:58-66; the three results are returned at :68.
 *   grad_out    (B, N, M*D)        dtype
 *   grad_value  (B, S, M, D)       ACCUMULATION type: f32 for DPFT_F32/F16/BF16, f64 for DPFT_F64.
 *                                  The caller zero-fills it; this call adds into it with red.global.add.
 *   grad_loc    (B, N, M, L, P, 2) dtype, fully overwritten
 *   grad_attn   (B, N, M, L, P)    dtype, fully overwritten
 */
DPFT_API int dpft_msda_backward(const void* value, const int64_t* shapes, const int64_t* lsi, const void* loc,
                       const void* attn, const void* grad_out, void* grad_value, void* grad_loc,
                       void* grad_attn, int B, int S, int M, int D, int N, int L, int P, int dtype,
                       void* stream);

/*
 * 2-d convolution + folded BatchNorm (+ residual) (+ ReLU), NHWC 16-bit (dtype = DPFT_BF16 or DPFT_F16; f16 outputs
 * saturate at +-65504), fp32 accumulation, as a tcgen05 implicit GEMM.
 * Replaces the Conv2d/BatchNorm2d/ReLU/add launches of the torchvision Bottleneck blocks the reference runs at
 * src/dprt/models/backbones/resnet.py:101 (built at :54-55), in eval() form:
 *   y[b,p,q,n] = act( sum_{r,s,c} x[b, p*stride-pad+r, q*stride-pad+s, c] * w[n,r,s,c] + bias[n] (+ residual[b,p,q,n]) )
 *   x        (B, H, W, Cin)       bf16, Cin  % 64 == 0
 *   w        (Cout, R, S, Cin)    bf16, Cout % 64 == 0   (BatchNorm scale already folded in)
 *   bias     (Cout,)              f32                    (BatchNorm shift)
 *   residual (B, P, Q, Cout)      bf16 or NULL
 *   y        (B, P, Q, Cout)      bf16, P = (H+2*pad-R)/stride+1, Q likewise
 * block_n: 0 = choose, or 64 / 128 / 256 (output-channel tile; tests sweep it).
 * cluster_mode: 0 = choose, 1 = one CTA per tile, 2 = CTA pairs (tcgen05 cta_group::2, 256-row tiles, B tile split across the pair).
 */
DPFT_API int dpft_conv2d_nhwc(const void* x, const void* w, const float* bias, const void* residual, void* y,
                              int B, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad, int relu,
                              int block_n, int cluster_mode, int dtype, void* stream);
/* Same, with a cap on the number of persistent CTAs (0 = one per SM).  The kernels are persistent with ~200 KB of shared
 * memory per CTA, i.e. a CTA owns its SM; a small layer of a side view (the radar backbones beside the camera's) is
 * latency-bound, so spreading it over every SM buys it nothing and takes the SMs from the view on the critical path. */
DPFT_API int dpft_conv2d_nhwc_ex(const void* x, const void* w, const float* bias, const void* residual, void* y,
                                 int B, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad, int relu,
                                 int block_n, int cluster_mode, int dtype, int max_ctas, void* stream);
/* Two problems of IDENTICAL shape (different tensors and weights) in one launch: the two radar views of the fusion model run
 * the same ResNet-50 on equally sized inputs.  Their layers are latency-bound, so the second problem rides along: half the
 * launches and half the SM-time of the pair.  max_ctas caps the CTAs of both problems together (0 = one per SM).  Returns
 * DPFT_ERR_UNSUPPORTED for a layer that is served by the halo or the weight-stationary kernel (call dpft_conv2d_nhwc_ex twice). */
DPFT_API int dpft_conv2d_nhwc_pair(const void* x0, const void* w0, const float* bias0, const void* residual0, void* y0,
                                   const void* x1, const void* w1, const float* bias1, const void* residual1, void* y1,
                                   int B, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad, int relu,
                                   int dtype, int max_ctas, void* stream);

/* OR-ed into `Cin` of dpft_stem_conv7x7_forward_ex and into `raw_channels` of dpft_fpn_output_forward: the raw input is
 * (B, H, W, C) uint8 — the frames an image decoder produces, before the reference's `.type(float32)`
 * (src/dprt/datasets/kradar/dataset.py read_image) — and is converted on load (0..255 is exact in f16): a quarter of the
 * bytes over PCIe and out of HBM.  Stem: Cin = 3, packed weights, output width >= 64 (the camera path). */
#define DPFT_RAW_U8 0x100

/*
 * ResNet stem: [1x1 adjustment conv (radar, 6 -> 3, resnet.py:47-51) folded into] conv1 7x7 stride 2 pad 3 + BatchNorm
 * (folded) + ReLU, reference src/dprt/models/backbones/resnet.py:98-101 (torchvision conv1/bn1/relu).
 *   x (B, H, W, Cin) f32 NHWC, Cin = 3 or 6, raw 0..255 values;  w [7][7][Cin][64] f32;  bias [64] f32
 *   y (B, P, Q, 64) dtype (DPFT_F16 | DPFT_BF16), P = (H-1)/2+1, Q = (W-1)/2+1
 * w_packed: output of dpft_stem_pack_weights or NULL.
 * impl: 0 = choose (tcgen05 implicit GEMM with f16 operands for Q >= 64 when w_packed is given, else the fp32 CUDA-core
 * kernel), 1 / 2 = force.
 */
DPFT_API int dpft_stem_conv7x7_forward(const float* x, const float* w, const void* w_packed, const float* bias, void* y,
                                       int B, int H, int W, int Cin, int dtype, int impl, void* stream);
/* Same with the ReLU optional (relu = 0: y = conv + bias, the pre-BatchNorm activation the training path stores). */
DPFT_API int dpft_stem_conv7x7_forward_ex(const float* x, const float* w, const void* w_packed, const float* bias, void* y,
                                          int B, int H, int W, int Cin, int dtype, int impl, int relu, void* stream);
/* w [7][7][Cin][64] f32 -> the f16 operand images the tcgen05 stem kernels stage (done once per model): `packed` must hold
 * DPFT_STEM_PACKED_BYTES = 86016 bytes (57344 in the 8-channels-per-tap layout, then, for Cin = 3, 28672 in the
 * 4-channels-per-tap layout of the row-streaming kernel). */
#define DPFT_STEM_PACKED_BYTES 86016
DPFT_API int dpft_stem_pack_weights(const float* w, void* packed, int Cin, void* stream);

/* torchvision ResNet maxpool (kernel 3, stride 2, padding 1), NHWC bf16, C % 8 == 0.  y (B, (H-1)/2+1, (W-1)/2+1, C). */
DPFT_API int dpft_maxpool3x3s2_nhwc(const void* x, void* y, int B, int H, int W, int C, int dtype, void* stream);

/*
 * FPN lateral stage of one backbone level (reference src/dprt/models/necks/fpn.py:77 -> torchvision
 * FeaturePyramidNetwork.forward: inner = conv1x1(x) + bias, + nearest-upsampled inner of the coarser level), as a
 * tcgen05 GEMM with a 16-channel fp32 epilogue.
 *   x (B, H, W, Cin) bf16, Cin % 64 == 0;  w [64][Cin] bf16 (rows 16..63 zero);  bias [64] f32 (16 used)
 *   coarse (B, Hc, Wc, 16) f32 or NULL;  out (B, H, W, 16) f32
 */
DPFT_API int dpft_fpn_lateral_forward(const void* x, const void* w, const float* bias, const float* coarse, int Hc, int Wc,
                                      float* out, int B, int H, int W, int Cin, int dtype, void* stream);

/*
 * FPN output stage of one level fused with the sinusoidal positional embedding, written into the view's pyramid:
 *   pyramid[b, start + p*W + q, :] = conv3x3(inner)[b, p, q, :] + bias + pos_x[q, :] + pos_y[p, :]
 * (fpn.py:77 layer_blocks; embeddings/sinusoidal.py:107-108; the (B, S, 16) layout of mpfusion.py:179); the pyramid is
 * stored as DPFT_F32 or DPFT_F16 (pyramid_dtype; f16 halves the bytes the decoder gathers).
 * Either `inner` (B, H, W, 16) f32 is given (levels fed by dpft_fpn_lateral_forward), or `raw` (B, H, W, raw_channels)
 * f32 with the lateral weights lat_w [16][raw_channels], lat_b [16] and the coarser inner map `coarse` (may be NULL):
 * then inner = lat_w raw + lat_b + nearest-upsampled coarse is formed on the fly (skip-link level, dprt.py:222-225).
 *   w [3][3][16 out][16 in] f32;  bias [16];  pos_y (H, 16), pos_x (W, 16) f32
 * w_packed: output of dpft_fpn_pack_weights or NULL.
 * impl: 0 = choose (tcgen05 row-strip kernel with an f16 inner tile for W >= 96 when w_packed is given, else the fp32
 * CUDA-core kernel), 1 / 2 = force.  3 = variant of 2 for the raw level (needs `coarse`, at most a quarter of the
 * size): column-owning tile builder with the coarse patch staged in shared memory (243 -> 201 us on the 720x1280 camera level);
 * what impl 0 chooses for an eligible raw level unless DPFT_FPN_BUILD=1 is set in the environment.
 */
DPFT_API int dpft_fpn_output_forward(const float* inner, const float* raw, int raw_channels, const float* lat_w,
                                     const float* lat_b, const float* coarse, int Hc, int Wc, const float* w,
                                     const void* w_packed, const float* bias, const float* pos_y, const float* pos_x, void* pyramid,
                                     int pyramid_dtype, long long S, long long start, int B, int H, int W, int impl,
                                     void* stream);
/* Same; `lat_w_host` (optional) = a HOST copy of the raw level's lateral weights [16][raw_channels]: they then travel as kernel
 * parameters and the tile builder's FMAs read them as constant operands (no shared-memory loads); results are identical. */
DPFT_API int dpft_fpn_output_forward_ex(const float* inner, const float* raw, int raw_channels, const float* lat_w,
                                        const float* lat_w_host, const float* lat_b, const float* coarse, int Hc, int Wc,
                                        const float* w, const void* w_packed, const float* bias, const float* pos_y,
                                        const float* pos_x, void* pyramid, int pyramid_dtype, long long S, long long start, int B,
                                        int H, int W, int impl, void* stream);

/* w [3][3][16][16] f32 -> the 4608-byte f16 operand image of the tcgen05 FPN output kernel (done once per model). */
DPFT_API int dpft_fpn_pack_weights(const float* w, void* packed, void* stream);

/*
 * Fused query decoder (inference), d_model = 16, 8 heads.
 *
 * dpft_decoder_layer_forward runs, for every (sample, view), one whole MLFusion layer of the reference
 * (src/dprt/models/fusers/mpfusion.py:231-263): self-attention (:122-148), reference-point projection
 * (IMPFusion.get_reference_points :617-696), multi-scale deformable cross-attention
 * (src/dprt/models/layers/ms_deform_attn.py:138-217) and the feed-forward block (:210-229), each with its
 * residual + LayerNorm.  It replaces MPFusion.forward's per-view loop (:496-509).
 *   views[v]      per-view inputs: the FPN pyramid with positional embedding (B, S, 16) f32, finest level first
 *                 (what mpfusion.py:179 concatenates), its level table, the calibration matrices of the batch
 *                 (label_to_<input>_t / _p, src/dprt/models/dprt.py:188-198), the original input (H, W)
 *                 (dprt.py:216) as f32, the device flag transformation.any() (mpfusion.py:647) and the packed
 *                 layer weights (layout: dpft_b200/decoder.py::pack_layer).
 *   query         (B, N, 16) f32, or (N, 16) with query_batch_stride = 0 (first iteration, mpfusion.py:727)
 *   pos           (N, 16) f32   query_embedding.weight (mpfusion.py:730)
 *   center        (B, N, 3) f32 current box centres, or (N, 3) with center_batch_stride = 0
 *   out           (B, V, N, 16) f32
 *   activation    0 = ReLU, 1 = Mish, 2 = GELU
 */
typedef struct dpft_decoder_view {
    const void* pyramid;        /* (B, S, 16), DPFT_F32 or DPFT_F16 (see pyramid_dtype) */
    const float* weights;
    const float* transform;     /* (B, 4, 4) */
    const float* projection;    /* (B, 4, 4); 3x4 matrices padded with the row [0 0 0 1] */
    const float* shape_hw;      /* (B, 2) */
    const int* use_transform;   /* 1 element */
    long long S;
    int level_h[8];
    int level_w[8];
    long long level_start[8];
} dpft_decoder_view;

DPFT_API int dpft_decoder_layer_forward(const dpft_decoder_view* views, int V, const float* query,
                                        long long query_batch_stride, const float* pos, const float* center,
                                        long long center_batch_stride, float* out, int B, int N, int L, int P,
                                        int d_ffn, int activation, int weight_floats, int pyramid_dtype, void* stream);

/*
 * View reduction + detection head for one iteration: MPFusion.reduce (mpfusion.py:416-470; 0 = 'linear' with the
 * channel-major/view-minor flattening of :438, 1 = 'mean', 2 = 'max') followed by LinearDetectionHead.forward
 * (src/dprt/models/heads/detection.py:252-275; three bias-free Linear layers per branch, ReLU between, centre
 * refinement center += previous centre at :273).  size/angle/class outputs may be NULL on intermediate iterations
 * (only the centre feeds the next iteration, mpfusion.py:732-743).
 *   views (B, V, N, 16); weights packed by dpft_b200/decoder.py::pack_head; outputs (B, N, 16|3|3|2|n_cls) f32.
 * Two kernels with bit-identical results (tests/test_decoder_head16_gpu.py): sixteen lanes per query (the default) and one
 * thread per query; reduction | DPFT_HEAD_LANES16 / reduction | DPFT_HEAD_LANES1 force one of them (DPFT_HEAD_LANES=1 in
 * the environment changes the default for A/B timing).
 */
#define DPFT_HEAD_LANES16 0x100
#define DPFT_HEAD_LANES1 0x200
DPFT_API int dpft_decoder_head_forward(const float* views, const float* weights, const float* center_in,
                                       long long center_batch_stride, float* query_out, float* center_out,
                                       float* size_out, float* angle_out, float* class_out, int B, int V, int N,
                                       int n_cls, int reduction, int weight_floats, void* stream);

/*
 * Decoder self-attention core as a tcgen05 flash-attention kernel (csrc/attention.cu): the scaled-dot-product step of the
 * nn.MultiheadAttention the reference's decoder layer builds at src/dprt/models/fusers/mpfusion.py:56-57 and calls at
 * :139 (q = k = query + pos, v = query, need_weights=False), eval mode (no attention dropout):
 *     out[b, n, h, :] = sum_j softmax_j(scale * <q[b,n,h,:], k[b,j,h,:]>) v[b,j,h,:]
 *   q, k, v: element [b, n, h, d] at  base + b * batch_stride + n * row_stride + h * D + d  (so the q/k/v slices of a packed
 *   in-projection are passed without a copy); out (B, N, H, D) contiguous; all four of `dtype` (DPFT_F32, DPFT_F16, DPFT_BF16).
 *   D <= 64, any N.  S = Q K^T and P V run on tcgen05.mma kind::f16 with fp32 accumulation in TMEM, the online softmax in
 *   registers.  precise != 0 (DPFT_F32 only): operands are split hi + lo and every product is three MMAs, which carries
 *   ~22 mantissa bits (fp32-grade results, |err| ~ 1e-6); precise == 0: single f16 operands (16-bit tier, ~1e-3).
 */
DPFT_API int dpft_self_attention_forward(const void* q, const void* k, const void* v, void* out, int B, int H, int N, int D,
                                         long long q_row_stride, long long k_row_stride, long long v_row_stride,
                                         long long q_batch_stride, long long k_batch_stride, long long v_batch_stride,
                                         float scale, int dtype, int precise, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Training path of the backbone (autograd of the torchvision Bottleneck blocks the reference trains,
 * src/dprt/models/backbones/resnet.py:54-55,101 under src/dprt/training/trainer.py:125-133): BatchNorm with batch
 * statistics is not folded, so a block is conv -> statistics -> normalise(+residual)(+ReLU), and backward is the
 * mirrored chain.  Data gradients reuse dpft_conv2d_nhwc on the flipped/transposed weights produced by
 * dpft_pack_conv_weights (stride-2 layers go through dpft_zero_insert2_nhwc first).
 * ------------------------------------------------------------------------------------------------------------- */

/*
 * Weight gradient: dw[n, r, s, c] += sum_{b,p,q} dy[b,p,q,n] * x[b, p*stride-pad+r, q*stride-pad+s, c]   (fp32, ADDED into dw;
 * the caller zero-fills).  tcgen05 GEMM over the pixel dimension with MN-major operands, split across `splits` work
 * items per tile (0 = choose).
 *   x (B, H, W, Cin) 16-bit, Cin % 64 == 0;  dy (B, P, Q, Cout) 16-bit, Cout % 64 == 0;  dw (Cout, R, S, Cin) f32
 */
DPFT_API int dpft_conv2d_wgrad(const void* x, const void* dy, float* dw, int B, int H, int W, int Cin, int Cout, int R, int S,
                               int stride, int pad, int splits, int dtype, void* stream);

/* BatchNorm2d (train) forward over y (M, C) 16-bit, C % 8 == 0 and 256 % (C/8) == 0:
 *   dpft_bn_forward_stats  per-block partial sums of y and y^2 into `workspace` (>= DPFT_BN_MAX_PARTS * 2 * C floats, no atomics,
 *                          deterministic), then mean, invstd = 1/sqrt(var_biased + eps), scale = gamma*invstd,
 *                          shift = beta - mean*scale, and the running_mean/var momentum update with the unbiased variance
 *                          (running_* may be NULL)
 *   dpft_bn_apply          z = [relu]( y*scale + shift (+ residual) )
 */
#define DPFT_BN_MAX_PARTS 296
DPFT_API int dpft_bn_forward_stats(const void* y, float* workspace, const float* gamma, const float* beta, float* running_mean,
                                   float* running_var, float momentum, float eps, long long M, int C, float* scale, float* shift,
                                   float* mean, float* invstd, int dtype, void* stream);
DPFT_API int dpft_bn_apply(const void* y, const float* scale, const float* shift, const void* residual, void* z, long long M, int C,
                           int relu, int dtype, void* stream);

/* BatchNorm2d (train) backward through z = [relu](bn(y) (+ residual)), two passes:
 *   dpft_bn_backward_reduce  g = dz * [z > 0] (when relu);  sum_g[c] = sum g,  sum_gx[c] = sum g * (y-mean)*invstd (partials in
 *                            `workspace` as above);  dgamma += sum_gx, dbeta += sum_g (fp32, may be NULL)
 *   dpft_bn_backward_apply   dy = gamma*invstd*(g - sum_g/M - xhat*sum_gx/M);  g_out = g (gradient of the residual branch,
 *                            may be NULL)
 */
DPFT_API int dpft_bn_backward_reduce(const void* dz, const void* z, const void* y, const float* mean, const float* invstd,
                                     float* workspace, float* sum_g, float* sum_gx, float* dgamma, float* dbeta, long long M, int C,
                                     int relu, int dtype, void* stream);
DPFT_API int dpft_bn_backward_apply(const void* dz, const void* z, const void* y, const float* mean, const float* invstd,
                                    const float* gamma, const float* sum_g, const float* sum_gx, void* dy, void* g_out,
                                    long long M, int C, int relu, int dtype, void* stream);

/* Weight gradient of the stem (dpft_stem_conv7x7_forward): dw [7][7][Cin][64] f32 += sum dy[b,p,q,o] * x[b,2p-3+r,2q-3+s,c];
 * x (B, H, W, Cin) f32, dy (B, P, Q, 64) 16-bit; workspace >= 2 * sm_count * 49*Cin*64 floats (per-CTA partial gradients). */
DPFT_API int dpft_stem_conv7x7_wgrad(const float* x, const void* dy, float* workspace, long long workspace_floats, float* dw, int B,
                                     int H, int W, int Cin, int dtype, void* stream);

/* Backward of dpft_maxpool3x3s2_nhwc: x (B, H, W, C) is the input of the pooling, pooled (B, P, Q, C) its output, dy (B, P, Q, C);
 * the gradient of a window goes to its first maximum in (row, column) order, as torch.nn.functional.max_pool2d does. */
DPFT_API int dpft_maxpool3x3s2_backward(const void* x, const void* pooled, const void* dy, void* dx, int B, int H, int W, int C,
                                        int dtype, void* stream);

/* up[b, 2p, 2q, :] = src[b, p, q, :], zero elsewhere; src (B, P, Q, C), up (B, H, W, C), 16-bit, C % 8 == 0. */
DPFT_API int dpft_zero_insert2_nhwc(const void* src, void* up, int B, int H, int W, int C, int P, int Q, void* stream);

/* One launch re-lays out every convolution weight of a model after an optimiser step.  `table` is a DEVICE array of: */
typedef struct dpft_pack_entry {
    const float* src;      /* fp32 master weights (Cout, Cin, R, S) (torch layout) */
    void* fwd;             /* 16-bit (Cout, R, S, Cin): operand of dpft_conv2d_nhwc / layout of dpft_conv2d_wgrad */
    void* dgrad;           /* 16-bit (Cin, R, S, Cout), taps flipped: operand of the data-gradient convolution; may be NULL */
    int Cout, Cin, R, S;
    long long offset;      /* first element of this entry in the concatenated index space, ascending */
} dpft_pack_entry;
DPFT_API int dpft_pack_conv_weights(const void* table, int n_layers, long long total, int dtype, void* stream);
/* The reverse for gradients: entry.src = fp32 (Cout, R, S, Cin) gradient from dpft_conv2d_wgrad, entry.fwd = fp32 parameter
 * gradient (Cout, Cin, R, S) which is ADDED to. */
DPFT_API int dpft_unpack_conv_wgrads(const void* table, int n_layers, long long total, void* stream);

/*
 * Hungarian assignment of the training criterion on the device (opt-in: dpft_b200.criterion, lsap_solver="device").  Replaces the
 * per-sample `C.cpu()` + scipy.optimize.linear_sum_assignment of HungarianAnassigner.forward
 * (src/dprt/training/assigner.py:134-141): one warp per sample solves the rectangular linear-sum-assignment problem by
 * shortest augmenting paths (dpft_b200/csrc/lsap_core.h; the same source, built for the host, is checked against scipy).
 *   cost (B, N, Mmax) f32: cost[b][n][m] of matching prediction n to ground-truth box m;  counts (B) int32: valid boxes of
 *   each sample (<= Mmax <= 64 <= N);  col4row (B, Mmax) int64 out: the prediction matched to each box, -1 for padded slots.
 */
DPFT_API int dpft_lsap_forward(const float* cost, const int* counts, long long* col4row, int B, int N, int Mmax, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DPFT_B200_H */
