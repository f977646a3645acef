/*
 * lmnet_b200.h — C ABI of the B200 (sm_100a) hot-path library for LM-Net.
 *
 * This is the drop-in boundary.  The reference has no native
 * code of its own: it borrows this arithmetic from the external `natten` package
 * (/root/reference/core/modules.py:18,509,517) and from cuDNN/ATen library kernels
 * under ReparamConv (/root/reference/core/modules.py:586-600).  Each entry point
 * names the reference interface it replaces.
 *
 * Conventions
 *  - extern "C", plain C types only.  Device pointers + sizes + a CUDA stream
 *    (passed as void* so that this header needs no CUDA include).
 *  - The caller owns every buffer (inputs, outputs, workspaces).  The library never
 *    allocates, frees or retains pointers, never synchronises the device, and
 *    enqueues all work on the given stream.
 *  - Return value: 0 on success, negative lmnet_status on failure.  Nothing throws
 *    or exits.  No CPU fallback exists: without a CUDA device every compute call
 *    returns LMNET_ERR_LAUNCH.
 *  - Thread-safe / re-entrant per stream (no mutable global state).
 *  - dtype is the element type of activations (q,k,v,out,attn,x,...).  Parameters
 *    (rpb, conv weights, BN affine and running statistics) and statistics are fp32.
 */
#ifndef LMNET_B200_H_
#define LMNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LMNET_ABI_VERSION 1

typedef enum lmnet_status {
    LMNET_OK = 0,
    LMNET_ERR_INVALID_ARG = -1, /* null pointer, bad size, kernel_size even, H < K*dilation, ... */
    LMNET_ERR_UNSUPPORTED = -2, /* dtype / shape outside what the kernels implement */
    LMNET_ERR_LAUNCH = -3,      /* cudaGetLastError() after a launch, or no device */
    LMNET_ERR_WORKSPACE = -4    /* workspace too small */
} lmnet_status;

typedef enum lmnet_dtype { LMNET_F32 = 0, LMNET_BF16 = 1, LMNET_F16 = 2 } lmnet_dtype;

/* A logical [B, H, W, heads, D] view: element strides for batch, row, column and
 * head; the head-dim stride is 1.  Covers natten's unfused layout [B,heads,H,W,D],
 * the fused layout [B,H,W,heads,D] and slices of a packed qkv tensor [B,H,W,3,heads,D]. */
typedef struct lmnet_view5 {
    void* ptr;
    int64_t sb, sh, sw, sn;
} lmnet_view5;

typedef struct lmnet_na2d_dims {
    int32_t B, H, W, heads, D;
    int32_t kernel_size; /* odd, 3..13 */
    int32_t dilation;    /* >= 1, H and W must be >= kernel_size*dilation */
} lmnet_na2d_dims;

int lmnet_abi_version(void);
const char* lmnet_status_string(int status);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t lmnet_launch_count(void);

/* Optional per-kernel profile (used by bench.py for the roofline of the dominant kernel).
 * While enabled, every launch is bracketed by CUDA events on its own stream and logged with the
 * ALGORITHMIC bytes of that launch.  collect() synchronises the events and sums, per
 * kernel id, elapsed milliseconds, launch count and algorithmic bytes into arrays of length
 * >= lmnet_profile_num_kernels().  enable() clears the log.  Off by default; never on in a timed run. */
int lmnet_profile_enable(int on);
int lmnet_profile_num_kernels(void);
const char* lmnet_profile_kernel_name(int kernel_id);
int lmnet_profile_collect(double* ms, uint64_t* launches, double* alg_bytes, int n);

/* ---- fused neighbourhood attention -------------------------------------------------
 * Replaces: natten.functional.na2d and the inner sequence of natten's
 * NeighborhoodAttention2D.forward (q*scale -> na2d_qk+rpb -> softmax -> na2d_av),
 * called from /root/reference/core/modules.py:517.  No attention map is materialised.
 *   out[b,i,j,h,:] = sum_n softmax_n(scale*q.k_n + rpb[h,pi,pj]) * v_n
 * rpb: fp32 [heads, 2K-1, 2K-1] or NULL.  lse: fp32 [B,H,W,heads] (natural log) or NULL. */
int lmnet_na2d_fwd(const lmnet_view5* q, const lmnet_view5* k, const lmnet_view5* v,
                   const float* rpb, const lmnet_view5* out, float* lse,
                   const lmnet_na2d_dims* dims, float scale, int dtype, void* stream);

/* Backward of the fused op with recomputation (replaces the autograd of natten's
 * na2d_qk / softmax / na2d_av chain).  drpb: fp32 [heads,2K-1,2K-1]
 * (overwritten) or NULL.  workspace: lmnet_na2d_bwd_workspace_bytes(). */
size_t lmnet_na2d_bwd_workspace_bytes(const lmnet_na2d_dims* dims);
int lmnet_na2d_bwd(const lmnet_view5* q, const lmnet_view5* k, const lmnet_view5* v,
                   const float* rpb, const lmnet_view5* dout,
                   const lmnet_view5* dq, const lmnet_view5* dk, const lmnet_view5* dv,
                   float* drpb, void* workspace, size_t workspace_bytes,
                   const lmnet_na2d_dims* dims, float scale, int dtype, void* stream);

/* ---- unfused ops (natten.functional.na2d_qk / na2d_av, 0.14 names natten2dqkrpb /
 * natten2dav).  attn / dattn are contiguous [B,heads,H,W,K*K]. */
int lmnet_na2d_qk_fwd(const lmnet_view5* q, const lmnet_view5* k, const float* rpb, void* attn,
                      const lmnet_na2d_dims* dims, int dtype, void* stream);
size_t lmnet_na2d_qk_bwd_workspace_bytes(const lmnet_na2d_dims* dims);
int lmnet_na2d_qk_bwd(const lmnet_view5* q, const lmnet_view5* k, const void* dattn,
                      const lmnet_view5* dq, const lmnet_view5* dk, float* drpb,
                      void* workspace, size_t workspace_bytes,
                      const lmnet_na2d_dims* dims, int dtype, void* stream);
int lmnet_na2d_av_fwd(const void* attn, const lmnet_view5* v, const lmnet_view5* out,
                      const lmnet_na2d_dims* dims, int dtype, void* stream);
int lmnet_na2d_av_bwd(const void* attn, const lmnet_view5* v, const lmnet_view5* dout,
                      void* dattn, const lmnet_view5* dv,
                      const lmnet_na2d_dims* dims, int dtype, void* stream);

/* ---- fused depthwise multi-branch conv + BatchNorm + sum + GELU (+ SE pooling) ------
 * Replaces the middle section of ReparamConv.forward,
 * /root/reference/core/modules.py:592-597 (four depthwise branches 5x5, 3x3, 3x1, 1x3
 * each followed by BatchNorm2d, summed, exact-erf GELU; SE's global average pool rides
 * on the writer, /root/reference/core/modules.py:1030-1031).
 * Branch order everywhere: 0 = large (5x5), 1 = square (3x3), 2 = ver (3x1), 3 = hor (1x3).
 * x, u, z: [B,E,H,W] contiguous NCHW of `dtype`.  Weights fp32: w5 [E,25], w3 [E,9],
 * w31 [E,3], w13 [E,3]. */
typedef struct lmnet_dw_params {
    const float* w[4];         /* w5, w3, w31, w13 */
    const float* gamma[4];     /* BN weight  [E] per branch */
    const float* beta[4];      /* BN bias    [E] per branch */
    float* running_mean[4];    /* [E] per branch; updated in training when non-NULL */
    float* running_var[4];     /* [E] per branch */
} lmnet_dw_params;

typedef struct lmnet_dw_grads {
    float* dw[4];              /* same shapes as lmnet_dw_params.w */
    float* dgamma[4];
    float* dbeta[4];
} lmnet_dw_grads;

typedef struct lmnet_dw_dims {
    int32_t B, E, H, W;
} lmnet_dw_dims;

/* One workspace size serves all three entry points below (the backward keeps a [B,E,H,W]
 * scratch tensor of `dtype` in it). */
size_t lmnet_reparam_dw_workspace_bytes(const lmnet_dw_dims* dims, int dtype);

/* Training forward: batch statistics (biased variance for normalisation, unbiased for the
 * running update: running = (1-momentum)*running + momentum*batch), writes
 *   u = sum_br BN_br(dwconv_br(x))  (pre-activation, saved for backward; may be NULL),
 *   z = GELU(u), pool[b,e] = mean_hw z (fp32 [B,E], may be NULL),
 *   save_mean / save_rstd: fp32 [4,E].
 * num_batches_tracked: NULL, or 4 device pointers (NULL entries allowed) to the BatchNorm
 * counters (int64), each incremented by one on the stream. */
int lmnet_reparam_dw_train_fwd(const void* x, const lmnet_dw_params* p, void* u, void* z, float* pool,
                               float* save_mean, float* save_rstd, float eps, float momentum,
                               int64_t* const* num_batches_tracked,
                               void* workspace, size_t workspace_bytes,
                               const lmnet_dw_dims* dims, int dtype, void* stream);

/* Training backward.  dz: gradient w.r.t. z; dpool: fp32 [B,E] gradient w.r.t. pool or NULL.
 * Outputs dx [B,E,H,W] and all parameter gradients (overwritten). */
int lmnet_reparam_dw_train_bwd(const void* x, const void* u, const void* dz, const float* dpool,
                               const lmnet_dw_params* p, const float* save_mean, const float* save_rstd,
                               void* dx, const lmnet_dw_grads* g,
                               void* workspace, size_t workspace_bytes,
                               const lmnet_dw_dims* dims, int dtype, void* stream);

/* Training forward / backward with the forward's Gram by-product (same reference lines as above,
 * /root/reference/core/modules.py:592-597; what autograd keeps between the two passes of the four
 * conv + BatchNorm2d branches).  save_gram: fp32 [E][lmnet_reparam_dw_gram_floats()] lag sums of the
 * branch outputs and of x taken by the statistics pass; with them the backward is ONE composite
 * stencil pass (5x5 on du, 9x9 on x) plus a 2-pixel frame patch and needs no recomputation of
 * the branch outputs.  *gram_saved (host int) is set to 1 when the buffer was filled (16-bit storage,
 * W % 8 == 0, 16-byte aligned tensors), else 0 — pass save_gram to the backward only in the first
 * case, NULL otherwise (the backward then is lmnet_reparam_dw_train_bwd).  save_gram and gram_saved
 * are both NULL or both non-NULL. */
int lmnet_reparam_dw_gram_floats(void);
int lmnet_reparam_dw_train_fwd_gram(const void* x, const lmnet_dw_params* p, void* u, void* z, float* pool,
                                    float* save_mean, float* save_rstd, float eps, float momentum,
                                    int64_t* const* num_batches_tracked, float* save_gram, int* gram_saved,
                                    void* workspace, size_t workspace_bytes,
                                    const lmnet_dw_dims* dims, int dtype, void* stream);
int lmnet_reparam_dw_train_bwd_gram(const void* x, const void* u, const void* dz, const float* dpool,
                                    const lmnet_dw_params* p, const float* save_mean, const float* save_rstd,
                                    const float* save_gram, void* dx, const lmnet_dw_grads* g,
                                    void* workspace, size_t workspace_bytes,
                                    const lmnet_dw_dims* dims, int dtype, void* stream);

/* Inference forward with running statistics folded in (the algebra of
 * ReparamConv.get_equivalent_kernel_bias, /root/reference/core/modules.py:622-642):
 * one 5x5 depthwise conv + bias + GELU (+ pool).  If p->gamma[0] is NULL the weights in
 * p->w[0] are taken as an already-fused 5x5 kernel and `bias` as its bias (deploy mode,
 * /root/reference/core/modules.py:644-657); otherwise `bias` is ignored. */
int lmnet_reparam_dw_eval_fwd(const void* x, const lmnet_dw_params* p, const float* bias, float eps,
                              void* z, float* pool, void* workspace, size_t workspace_bytes,
                              const lmnet_dw_dims* dims, int dtype, void* stream);

/* ---- training loss of the reference loop in one pass per direction (widening step f4) -------------
 * loss = CrossEntropyLoss(weight = ce_weight, label_smoothing)(logits, labels)      /root/reference/train.py:157
 *      + DiceLoss(C)(logits, labels, weight = dice_weight, softmax = True)           /root/reference/utils/loss.py:170-206
 * as summed in /root/reference/utils/train_eval_utils.py:141-142.  logits [B,C,H,W] contiguous of `dtype`, labels int64
 * [B,H,W] with values in [0,C), weights fp32 [C]; fp32 arithmetic, fixed-order reductions, no host synchronisation.
 * stats: fp32 [lmnet_seg_loss_stats_floats(C)], stats[0] = the loss, the rest feeds the backward.
 * dloss: device pointer to the upstream gradient of the scalar (fp32).  C in {2, 3, 4, 8}. */
int lmnet_seg_loss_supported(int C, int dtype);
size_t lmnet_seg_loss_workspace_bytes(int C);
int lmnet_seg_loss_stats_floats(int C);
int lmnet_seg_loss_fwd(const void* logits, const int64_t* labels, const float* ce_weight, const float* dice_weight,
                       float label_smoothing, float* stats, void* workspace, size_t workspace_bytes, int64_t B, int C,
                       int64_t HW, int dtype, void* stream);
int lmnet_seg_loss_bwd(const void* logits, const int64_t* labels, const float* ce_weight, const float* dice_weight,
                       float label_smoothing, const float* stats, const float* dloss, void* dlogits, int64_t B, int C,
                       int64_t HW, int dtype, void* stream);

/* ---- fused BatchNorm2d + activation on NCHW tensors (widening step f1/f3, SURVEY.md §8 f) -------
 * Replaces nn.BatchNorm2d followed by an activation where the reference applies them back to back:
 * expand_conv's BN + Hardswish (/root/reference/core/modules.py:537-539) and the skip blocks'
 * BN + GELU (/root/reference/core/modules.py:97-100, 122-125, 134-137).  y, out, dout, dy: [B,C,HW]
 * contiguous of `dtype`; gamma/beta may be NULL (affine off).  Training: batch statistics, running
 * statistics and the counter updated in place when non-NULL (momentum semantics of nn.BatchNorm2d),
 * save_mean/save_rstd [C] written.  Inference: running statistics are used. */
typedef enum lmnet_act { LMNET_ACT_NONE = 0, LMNET_ACT_HARDSWISH = 1, LMNET_ACT_GELU = 2, LMNET_ACT_RELU = 3 } lmnet_act;
typedef struct lmnet_bn_dims {
    int32_t B, C;
    int64_t HW;
} lmnet_bn_dims;
size_t lmnet_bn_act_workspace_bytes(const lmnet_bn_dims* dims);
int lmnet_bn_act_fwd(const void* y, const float* gamma, const float* beta, float* running_mean,
                     float* running_var, int64_t* num_batches_tracked, void* out, float* save_mean,
                     float* save_rstd, float eps, float momentum, int training, int act,
                     void* workspace, size_t workspace_bytes, const lmnet_bn_dims* dims, int dtype, void* stream);
int lmnet_bn_act_bwd(const void* y, const void* dout, const float* gamma, const float* beta,
                     const float* save_mean, const float* save_rstd, void* dy, float* dgamma, float* dbeta,
                     int act, void* workspace, size_t workspace_bytes, const lmnet_bn_dims* dims, int dtype,
                     void* stream);
/* Channels-last variants: y / out / dout / dy are [B, HW, C] with C contiguous (the layout cuDNN's 16-bit
 * convolutions produce), same semantics, same workspace query.  lmnet_bn_act_cl_supported(): B*HW*C must be a
 * multiple of the 16-byte vector and lcm(C, vector) / vector <= 256. */
int lmnet_bn_act_cl_supported(const lmnet_bn_dims* dims, int dtype);
int lmnet_bn_act_cl_fwd(const void* y, const float* gamma, const float* beta, float* running_mean,
                        float* running_var, int64_t* num_batches_tracked, void* out, float* save_mean,
                        float* save_rstd, float eps, float momentum, int training, int act,
                        void* workspace, size_t workspace_bytes, const lmnet_bn_dims* dims, int dtype, void* stream);
int lmnet_bn_act_cl_bwd(const void* y, const void* dout, const float* gamma, const float* beta,
                        const float* save_mean, const float* save_rstd, void* dy, float* dgamma, float* dbeta,
                        int act, void* workspace, size_t workspace_bytes, const lmnet_bn_dims* dims, int dtype,
                        void* stream);

/* ---- LayerNorm over short channels-last rows (widening step f2, SURVEY.md §8 f) ------------------
 * Replaces nn.LayerNorm(C) as used twice per NeighborhoodTransformer
 * (/root/reference/core/modules.py:507, 510, 515-518) for C in {12, 24, 48, 96}.  x, y, dy, dx: [rows, C]
 * contiguous of `dtype` (16-byte aligned); gamma/beta fp32 [C] or NULL; save_mean/save_rstd fp32 [rows]. */
int lmnet_layer_norm_supported(int C);
size_t lmnet_layer_norm_workspace_bytes(int64_t rows, int C);
int lmnet_layer_norm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* save_mean,
                         float* save_rstd, int64_t rows, int C, float eps, int dtype, void* stream);
int lmnet_layer_norm_bwd(const void* x, const void* dy, const float* gamma, const float* save_mean,
                         const float* save_rstd, void* dx, float* dgamma, float* dbeta,
                         void* workspace, size_t workspace_bytes, int64_t rows, int C, int dtype, void* stream);

/* ---- weight / bias gradients of 1x1 convolutions on NCHW tensors (widening step f1) -------------
 * For the 1x1 convolutions of ReparamConv (/root/reference/core/modules.py:537, 576-584): forward and
 * input gradients are plain GEMMs on the NCHW planes (cuBLAS via torch.bmm on the host side); this entry
 * point is the pixel-reduction GEMM cuBLAS serves badly:
 *     dW[b][m][n] = sum_p A[b][m][p] * Bt[b][n][p],   drow[b][m] = sum_p A[b][m][p]
 * A = grad_output [B,M,P]; Bt = layer input, rows [0,N1) from B1 [B,N1,P] and [N1,N1+N2) from B2 [B,N2,P]
 * (B2 may be NULL when N2 == 0).  16-bit dtypes, P % 8 == 0, 16-byte aligned tensors; outputs fp32, per batch
 * item (dW [B,M,N1+N2], drow [B,M] or NULL).  lmnet_wgrad_1x1_supported() tells whether a shape is covered. */
typedef struct lmnet_wgrad_dims {
    int32_t B, M, N1, N2;
    int64_t P;
} lmnet_wgrad_dims;
int lmnet_wgrad_1x1_supported(const lmnet_wgrad_dims* dims, int dtype);
size_t lmnet_wgrad_1x1_workspace_bytes(const lmnet_wgrad_dims* dims);
int lmnet_wgrad_1x1(const void* A, const void* B1, const void* B2, float* dW, float* drow,
                    void* workspace, size_t workspace_bytes, const lmnet_wgrad_dims* dims, int dtype, void* stream);

/* Mixed-layout variant: an operand flagged channels-last is [B, P, C] (C contiguous: the memory layout of the
 * block's NHWC input / output tensors) instead of [B, C, P]; B2, when present (N2 > 0), is always channels-last.
 * Used by the 1x1 convolutions of ReparamConv when the block runs on channels-last tensors (no transposed copies):
 * expand (A = grad of the expanded planes, B1 = x channels-last) and pointwise + shortcut (A = grad_output
 * channels-last, B1 = z planes, B2 = x channels-last).  /root/reference/core/modules.py:587, 598-599. */
int lmnet_wgrad_1x1_cl_supported(const lmnet_wgrad_dims* dims, int a_channels_last, int b1_channels_last, int dtype);
size_t lmnet_wgrad_1x1_cl_workspace_bytes(const lmnet_wgrad_dims* dims, int a_channels_last, int b1_channels_last);
int lmnet_wgrad_1x1_cl(const void* A, const void* B1, const void* B2, float* dW, float* drow,
                       void* workspace, size_t workspace_bytes, const lmnet_wgrad_dims* dims,
                       int a_channels_last, int b1_channels_last, int dtype, void* stream);
/* Parameter gradients of pointwise(gate * z) + shortcut(x) out of the per-image weight gradient dW [B][M][E + Cin] and
 * drow [B][M] of lmnet_wgrad_1x1_cl (the effective pointwise weight of image b is Wpw * gate_b):
 * dWpw [M][E], dgate [B][E], dWsc [M][Cin], dbias [M]; wpw is read in place with element strides.  fp32. */
int lmnet_pointwise_grads(const float* dW, const float* drow, const float* gate, const float* wpw, int64_t wpw_sm, int64_t wpw_se,
                          float* dwpw, float* dgate, float* dwsc, float* dbias, int B, int M, int E, int Cin, void* stream);
/* same, summed over the batch in the fixed-order reduction: dW [M][N1+N2], drow [M] (layers whose weights are shared
 * by all images need no per-image result) */
int lmnet_wgrad_1x1_cl_sum(const void* A, const void* B1, const void* B2, float* dW, float* drow,
                           void* workspace, size_t workspace_bytes, const lmnet_wgrad_dims* dims,
                           int a_channels_last, int b1_channels_last, int dtype, void* stream);

/* ---- 1x1 convolutions as small-K GEMMs over pixels (widening step f1: forward + input gradients) -----------
 * out[b] = W1[b or shared] . in1[b]  (+ W2 . in2[b])  (+ bias)   for every batch image b and pixel.
 * Replaces expand_conv[0], pointwise_conv[0](gate * z) + shortcut[0](x) and their input gradients
 * (/root/reference/core/modules.py:537, 576-584, 587, 598-599).  Operand layouts: "planes" = [B, C, P] (NCHW),
 * "channels-last" = [B, P, C].  in2 (K2 > 0) is always channels-last and requires a channels-last output; a planes
 * output requires a channels-last in1.  bias fp32 [N] or NULL.  stats_part (planes output only, or NULL): per-CTA partial sums
 * [N][ctas][2] (sum, sum of squares of the stored output) in the layout of lmnet_bn_act_fwd_stats;
 * ctas = lmnet_pixel_gemm_stats_ctas().  16-bit dtypes, P % 8 == 0, channel counts % 4 == 0. */
typedef struct lmnet_pgemm_dims {
    int32_t B;
    int64_t P;
    int32_t N, K1, K2;
} lmnet_pgemm_dims;
/* Weights are the layer's fp32 parameters, read in place with element strides (a transposed view costs nothing) and
 * rounded to `dtype` while they are staged in shared memory — no cast / permute / gate-multiply kernels around the call.
 * W1[n][k] = w1[n * w1_sn + k * w1_sk] * (gate ? gate[b][gate_on_n ? n : k] : 1)   (gate: fp32 [B][K1] or [B][N], the
 * squeeze-excite gate folded into the pointwise weights per image);  W2[n][k] = w2[n * w2_sn + k * w2_sk]. */
typedef struct lmnet_pgemm_weights {
    const float* w1;
    int64_t w1_sn, w1_sk;
    const float* gate;
    int32_t gate_on_n;
    const float* w2;
    int64_t w2_sn, w2_sk;
} lmnet_pgemm_weights;
int lmnet_pixel_gemm_supported(const lmnet_pgemm_dims* dims, int in1_channels_last, int out_channels_last,
                               int want_stats, int dtype);
int lmnet_pixel_gemm_stats_ctas(const lmnet_pgemm_dims* dims, int in1_channels_last, int out_channels_last);
int lmnet_pixel_gemm(const void* in1, int in1_channels_last, const void* in2, const lmnet_pgemm_weights* weights,
                     const float* bias, void* out, int out_channels_last, float* stats_part,
                     const lmnet_pgemm_dims* dims, int dtype, void* stream);
/* BatchNorm + activation forward (training, NCHW planes) whose statistics pass already happened elsewhere:
 * stats_part = [C][nchunks][2] partial (sum, sum of squares) over all B*HW elements of each channel. */
int lmnet_bn_act_fwd_stats(const void* y, const float* stats_part, int nchunks, const float* gamma, const float* beta,
                           float* running_mean, float* running_var, int64_t* num_batches_tracked, void* out,
                           float* save_mean, float* save_rstd, float eps, float momentum, int act,
                           void* workspace, size_t workspace_bytes, const lmnet_bn_dims* dims, int dtype, void* stream);

/* ---- squeeze-excite gate (row a7) -------------------------------------------------------------------------
 * Replaces scale_activation(fc2(activation(fc1(pooled)))) of SE.forward (/root/reference/core/modules.py:1030-1036):
 * gate[b][e] = hardsigmoid(b2[e] + sum_r W2[e][r] * relu(b1[r] + sum_e' W1[r][e'] * pool[b][e'])), everything fp32.
 * W1 [R][E], W2 [E][R] are the 1x1 convolution weights; h1 [B][R] / pre2 [B][E] (pre-activations) are saved for the
 * backward (NULL at inference).  bwd: dpool [B][E], dW1, db1, dW2, db2 from dgate [B][E]; fixed-order sums. */
int lmnet_se_gate_supported(int B, int E, int R);
int lmnet_se_gate_fwd(const float* pool, const float* W1, const float* b1, const float* W2, const float* b2, float* gate,
                      float* h1, float* pre2, int B, int E, int R, void* stream);
int lmnet_se_gate_bwd(const float* dgate, const float* pool, const float* h1, const float* pre2, const float* W1,
                      const float* W2, float* dpool, float* dW1, float* db1, float* dW2, float* db2, int B, int E, int R,
                      void* stream);

/* ---- average pooling by an integer factor, channels-last (widening step f3) --------------------------------
 * Replaces the nn.AdaptiveAvgPool2d calls of PyramidPool (/root/reference/core/modules.py:454-498) when the input size is
 * an exact multiple of the output size.  x / dx: [B, Ho*factor, Wo*factor, C], y / dy: [B, Ho, Wo, C]; 16-bit; C % 4 == 0. */
typedef struct lmnet_pool_dims {
    int32_t B, Ho, Wo, C, factor;
} lmnet_pool_dims;
int lmnet_avgpool_cl_fwd(const void* x, void* y, const lmnet_pool_dims* dims, int dtype, void* stream);
int lmnet_avgpool_cl_bwd(const void* dy, void* dx, const lmnet_pool_dims* dims, int dtype, void* stream);

/* ---- dense 3x3 convolution, padding 1, stride 1 or 2, channels-last (widening step f3) ---------------------
 * Replaces the nn.Conv2d(.., 3, stride, 1) layers of LM-Net outside ReparamConv: down1-4 / up1-4
 * (/root/reference/core/LM_Net.py:14-39, 58-74), M2Skip / M3Skip (/root/reference/core/modules.py:83-143) and the
 * OverlapPatchEmbed of the neighbourhood transformers (/root/reference/core/modules.py:30-39), where the channel counts
 * fit the resident-weights kernel (lmnet_conv3x3_*_supported); cuDNN keeps the wide, low-resolution layers.
 * x: [B, H, W, Cin], y / dy: [B, Ho, Wo, Cout] with Ho = (H - 1) / stride + 1; 16-bit dtypes; Cin, Cout % 4 == 0.
 * weight: the layer's fp32 parameter in torch layout, read in place and rounded to `dtype` while it is staged in shared
 * memory: w_transposed = 0: [Cout][Cin][3][3];  w_transposed = 1 (the stride-1 input gradient: x := dy, Cin := the
 * layer's Cout, Cout := the layer's Cin): the layer's own [Cin][Cout][3][3] tensor, used flipped and transposed.
 * bias fp32 [Cout] or NULL.
 * wgrad: dW fp32 in torch layout [Cout][Cin][3][3], dbias fp32 [Cout] (or NULL); deterministic (per-CTA partials in the
 * workspace + fixed-order reduction). */
typedef struct lmnet_conv3x3_dims {
    int32_t B, H, W, Cin, Cout, stride;
} lmnet_conv3x3_dims;
int lmnet_conv3x3_fwd_supported(const lmnet_conv3x3_dims* dims, int dtype);
int lmnet_conv3x3_fwd(const void* x, const float* weight, int w_transposed, const float* bias, void* y,
                      const lmnet_conv3x3_dims* dims, int dtype, void* stream);
int lmnet_conv3x3_wgrad_supported(const lmnet_conv3x3_dims* dims, int dtype);
size_t lmnet_conv3x3_wgrad_workspace_bytes(const lmnet_conv3x3_dims* dims);
int lmnet_conv3x3_wgrad(const void* x, const void* dy, float* dW, float* dbias, void* workspace, size_t workspace_bytes,
                        const lmnet_conv3x3_dims* dims, int dtype, void* stream);
/* Same kernels with a pixel pitch (elements between consecutive pixels; 0 or the channel count = dense) on the tensor
 * that autograd may hand over as a channel slice of a wider channels-last tensor — the gradient of a convolution whose
 * output went into torch.cat (/root/reference/core/modules.py M2Skip / M3Skip, /root/reference/core/LM_Net.py:58-74):
 * x of the forward kernel (= dy when it computes the stride-1 input gradient) and dy of the weight gradient are read in
 * place instead of being densified first. */
int lmnet_conv3x3_fwd_strided(const void* x, int x_pixel_pitch, const float* weight, int w_transposed, const float* bias, void* y,
                              const lmnet_conv3x3_dims* dims, int dtype, void* stream);
int lmnet_conv3x3_wgrad_strided(const void* x, const void* dy, int dy_pixel_pitch, float* dW, float* dbias, void* workspace,
                                size_t workspace_bytes, const lmnet_conv3x3_dims* dims, int dtype, void* stream);

/* ---- bilinear x2 up-sampling, align_corners=True, NCHW (widening step f3) ----------------------
 * Replaces nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) of the decoder and skip blocks
 * (/root/reference/core/LM_Net.py:58-74, /root/reference/core/modules.py:93-95, 129-131).
 * x: [planes, H, W] -> y: [planes, 2H, 2W] in `dtype`; the backward maps dy -> dx (gather, deterministic). */
typedef struct lmnet_upsample_dims {
    int64_t planes;
    int32_t H, W;
} lmnet_upsample_dims;
int lmnet_upsample2x_fwd(const void* x, void* y, const lmnet_upsample_dims* dims, int dtype, void* stream);
int lmnet_upsample2x_bwd(const void* dy, void* dx, const lmnet_upsample_dims* dims, int dtype, void* stream);
/* channels-last variant: x [B, H, W, C] -> y [B, 2H, 2W, C] (the layout cuDNN's 16-bit convolutions on either side
 * of these call sites compute in, so the tensor never changes layout). */
typedef struct lmnet_upsample_cl_dims {
    int64_t B;
    int32_t H, W, C;
} lmnet_upsample_cl_dims;
int lmnet_upsample2x_cl_fwd(const void* x, void* y, const lmnet_upsample_cl_dims* dims, int dtype, void* stream);
int lmnet_upsample2x_cl_bwd(const void* dy, void* dx, const lmnet_upsample_cl_dims* dims, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LMNET_B200_H_ */
