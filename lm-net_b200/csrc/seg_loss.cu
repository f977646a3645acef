// seg_loss.cu — the training loss of the reference loop in two kernels per direction (widening step f4, SURVEY.md §8 f):
//     loss = CrossEntropyLoss(weight = w, label_smoothing = eps)(logits, y)            (/root/reference/train.py:157)
//          + DiceLoss(C)(logits, y, weight = dw, softmax = True)                       (/root/reference/utils/loss.py:170-206)
// as called in /root/reference/utils/train_eval_utils.py:141-142.  In stock torch ops this is ~45 launches over the
// [B, C, H, W] logits (log-softmax, gather, one-hot, six masked sums, softmax, their backward chain: 0.5 ms of the step
// for 4 M logits); here the forward reads the logits once and leaves 3 + 3C per-CTA partial sums, a one-block finalize
// forms the scalar loss and the coefficients of the backward, and the backward recomputes the softmax and writes
// dlogits once.  fp32 arithmetic on 16-bit or fp32 logits (what autocast does: both losses upcast), fixed-order
// reductions (deterministic), no host synchronisation (CUDA-graph capturable).
//
//   CE   = [(1 - eps) sum_i w[y_i] (-logp_i,y_i) + (eps / C) sum_i sum_c w_c (-logp_ic)] / sum_i w[y_i]
//   Dice = (1 / C) sum_c dw_c (1 - (2 I_c + s) / (Z_c + Y_c + s)),  I_c = sum_i p_ic [y_i = c], Z_c = sum_i p_ic^2,
//          Y_c = sum_i [y_i = c], s = 1e-5
#include <algorithm>

#include "common.cuh"

namespace lmnet {

constexpr int kLossMaxC = 8;
constexpr int kLossThreads = 256;
constexpr int kLossCtas = 148 * 4;

struct LossGeom {
    int64_t n_pix;      // B * H * W
    int64_t hw;         // H * W
    int C;
};

template <typename T, int C>
__global__ void __launch_bounds__(kLossThreads)
seg_loss_fwd_kernel(const T* __restrict__ logits, const int64_t* __restrict__ labels, const float* __restrict__ ce_w,
                    float* __restrict__ part /* [ncta][3 + 3C] */, LossGeom g) {
    constexpr int NP = 3 + 3 * C;
    float w[C];
#pragma unroll
    for (int c = 0; c < C; ++c) w[c] = __ldg(ce_w + c);
    float acc[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) acc[i] = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < g.n_pix; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / g.hw, r = i - b * g.hw;
        const T* px = logits + b * C * g.hw + r;
        float z[C], m = -INFINITY;
#pragma unroll
        for (int c = 0; c < C; ++c) { z[c] = to_f(px[c * g.hw]); m = fmaxf(m, z[c]); }
        float se = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) se += expf(z[c] - m);
        const float lse = m + logf(se), inv = 1.f / se;
        const int y = (int)labels[i];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float logp = z[c] - lse, p = expf(z[c] - m) * inv;
            const float hit = (c == y) ? 1.f : 0.f;
            acc[0] -= hit * w[c] * logp;            // sum w[y] * nll
            acc[1] -= w[c] * logp;                  // smoothing term
            acc[2] += hit * w[c];                   // sum w[y]
            acc[3 + c] += hit * p;                  // I_c
            acc[3 + C + c] = fmaf(p, p, acc[3 + C + c]);   // Z_c
            acc[3 + 2 * C + c] += hit;              // Y_c
        }
    }
    __shared__ float s_red[kLossThreads / 32][NP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const float v = warp_sum(acc[i]);
        if (lane == 0) s_red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < NP) {
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < kLossThreads / 32; ++k) v += s_red[k][threadIdx.x];
        part[(int64_t)blockIdx.x * NP + threadIdx.x] = v;
    }
}

// stats[0] = loss; stats[1] = 1 / sum w[y]; stats[2] = sum_c w_c; stats[3 + c] = D_c; stats[3 + C + c] = 1 / U_c
__global__ void seg_loss_fin_kernel(const float* __restrict__ part, int ncta, int C, const float* __restrict__ dice_w, float eps,
                                    float* __restrict__ stats, const float* __restrict__ ce_w) {
    const int NP = 3 + 3 * C;
    __shared__ double s_tot[3 + 3 * kLossMaxC];
    if ((int)threadIdx.x < NP) {
        double a = 0;
        for (int k = 0; k < ncta; ++k) a += part[(int64_t)k * NP + threadIdx.x];
        s_tot[threadIdx.x] = a;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const double smooth = 1e-5;
    double wsum = 0;
    for (int c = 0; c < C; ++c) wsum += ce_w[c];
    double loss = ((1.0 - eps) * s_tot[0] + (eps / C) * s_tot[1]) / s_tot[2];
    for (int c = 0; c < C; ++c) {
        const double U = s_tot[3 + C + c] + s_tot[3 + 2 * C + c] + smooth;
        const double D = (2.0 * s_tot[3 + c] + smooth) / U;
        loss += (1.0 - D) * dice_w[c] / C;
        stats[3 + c] = (float)D;
        stats[3 + C + c] = (float)(1.0 / U);
    }
    stats[0] = (float)loss;
    stats[1] = (float)(1.0 / s_tot[2]);
    stats[2] = (float)wsum;
}

template <typename T, int C>
__global__ void __launch_bounds__(kLossThreads)
seg_loss_bwd_kernel(const T* __restrict__ logits, const int64_t* __restrict__ labels, const float* __restrict__ ce_w,
                    const float* __restrict__ dice_w, const float* __restrict__ stats, const float* __restrict__ dloss, float eps,
                    T* __restrict__ dlogits, LossGeom g) {
    float w[C], gd[C], Dc[C];
    const float up = __ldg(dloss), inv_wy = __ldg(stats + 1), wsum = __ldg(stats + 2);
#pragma unroll
    for (int c = 0; c < C; ++c) {
        w[c] = __ldg(ce_w + c);
        Dc[c] = __ldg(stats + 3 + c);
        gd[c] = -__ldg(dice_w + c) / C * 2.f * __ldg(stats + 3 + C + c);       // dL/dp_c = gd_c ([y = c] - D_c p_c)
    }
    const float k1 = (1.f - eps) * inv_wy, k2 = eps / C * inv_wy;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < g.n_pix; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / g.hw, r = i - b * g.hw;
        const int64_t off = b * C * g.hw + r;
        float z[C], m = -INFINITY;
#pragma unroll
        for (int c = 0; c < C; ++c) { z[c] = to_f(logits[off + c * g.hw]); m = fmaxf(m, z[c]); }
        float se = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) { z[c] = expf(z[c] - m); se += z[c]; }
        const float inv = 1.f / se;
        const int y = (int)labels[i];
        float p[C], gp[C], dot = 0.f, wy = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            p[c] = z[c] * inv;
            const float hit = (c == y) ? 1.f : 0.f;
            wy += hit * w[c];
            gp[c] = gd[c] * (hit - Dc[c] * p[c]);
            dot = fmaf(p[c], gp[c], dot);
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float hit = (c == y) ? 1.f : 0.f;
            const float ce = k1 * wy * (p[c] - hit) + k2 * (p[c] * wsum - w[c]);
            const float dice = p[c] * (gp[c] - dot);
            dlogits[off + c * g.hw] = from_f<T>(up * (ce + dice));
        }
    }
}

template <typename T>
static int seg_loss_fwd_t(const void* logits, const int64_t* labels, const float* ce_w, const float* dice_w, float eps,
                          float* stats, float* ws, const LossGeom& g, cudaStream_t st) {
    const int ncta = (int)std::min<int64_t>(kLossCtas, (g.n_pix + kLossThreads - 1) / kLossThreads);
    const double bytes = (double)g.n_pix * (g.C * sizeof(T) + 8);
    switch (g.C) {
#define LMNET_LOSS_CASE(CC)                                                                                              \
        case CC:                                                                                                         \
            LMNET_LAUNCH(KID_SEG_LOSS, st, bytes, (seg_loss_fwd_kernel<T, CC><<<ncta, kLossThreads, 0, st>>>((const T*)logits, labels, ce_w, ws, g))); \
            break;
        LMNET_LOSS_CASE(2) LMNET_LOSS_CASE(3) LMNET_LOSS_CASE(4) LMNET_LOSS_CASE(8)
#undef LMNET_LOSS_CASE
        default: return LMNET_ERR_UNSUPPORTED;
    }
    LMNET_LAUNCH(KID_SEG_LOSS, st, 0, (seg_loss_fin_kernel<<<1, 32, 0, st>>>(ws, ncta, g.C, dice_w, eps, stats, ce_w)));
    return LMNET_OK;
}

template <typename T>
static int seg_loss_bwd_t(const void* logits, const int64_t* labels, const float* ce_w, const float* dice_w, float eps,
                          const float* stats, const float* dloss, void* dlogits, const LossGeom& g, cudaStream_t st) {
    const int ncta = (int)std::min<int64_t>(kLossCtas, (g.n_pix + kLossThreads - 1) / kLossThreads);
    const double bytes = (double)g.n_pix * (2 * g.C * sizeof(T) + 8);
    switch (g.C) {
#define LMNET_LOSS_CASE(CC)                                                                                              \
        case CC:                                                                                                         \
            LMNET_LAUNCH(KID_SEG_LOSS, st, bytes, (seg_loss_bwd_kernel<T, CC><<<ncta, kLossThreads, 0, st>>>(              \
                (const T*)logits, labels, ce_w, dice_w, stats, dloss, eps, (T*)dlogits, g)));                            \
            break;
        LMNET_LOSS_CASE(2) LMNET_LOSS_CASE(3) LMNET_LOSS_CASE(4) LMNET_LOSS_CASE(8)
#undef LMNET_LOSS_CASE
        default: return LMNET_ERR_UNSUPPORTED;
    }
    return LMNET_OK;
}

}  // namespace lmnet

using namespace lmnet;

extern "C" int lmnet_seg_loss_supported(int C, int dtype) {
    return (C == 2 || C == 3 || C == 4 || C == 8) && (dtype == LMNET_F32 || dtype == LMNET_BF16 || dtype == LMNET_F16);
}
extern "C" size_t lmnet_seg_loss_workspace_bytes(int C) { return (size_t)kLossCtas * (3 + 3 * (size_t)C) * sizeof(float); }
extern "C" int lmnet_seg_loss_stats_floats(int C) { return 3 + 2 * C; }

extern "C" int lmnet_seg_loss_fwd(const void* logits, const int64_t* labels, const float* ce_weight, const float* dice_weight,
                                  float label_smoothing, float* stats, void* workspace, size_t workspace_bytes, int64_t B, int C,
                                  int64_t HW, int dtype, void* stream) {
    if (!logits || !labels || !ce_weight || !dice_weight || !stats || !workspace || B <= 0 || HW <= 0) return LMNET_ERR_INVALID_ARG;
    if (!lmnet_seg_loss_supported(C, dtype)) return LMNET_ERR_UNSUPPORTED;
    if (workspace_bytes < lmnet_seg_loss_workspace_bytes(C)) return LMNET_ERR_WORKSPACE;
    const LossGeom g{B * HW, HW, C};
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case LMNET_F32: return seg_loss_fwd_t<float>(logits, labels, ce_weight, dice_weight, label_smoothing, stats, (float*)workspace, g, st);
        case LMNET_BF16: return seg_loss_fwd_t<__nv_bfloat16>(logits, labels, ce_weight, dice_weight, label_smoothing, stats, (float*)workspace, g, st);
        default: return seg_loss_fwd_t<__half>(logits, labels, ce_weight, dice_weight, label_smoothing, stats, (float*)workspace, g, st);
    }
}

extern "C" int lmnet_seg_loss_bwd(const void* logits, const int64_t* labels, const float* ce_weight, const float* dice_weight,
                                  float label_smoothing, const float* stats, const float* dloss, void* dlogits, int64_t B, int C,
                                  int64_t HW, int dtype, void* stream) {
    if (!logits || !labels || !ce_weight || !dice_weight || !stats || !dloss || !dlogits || B <= 0 || HW <= 0) return LMNET_ERR_INVALID_ARG;
    if (!lmnet_seg_loss_supported(C, dtype)) return LMNET_ERR_UNSUPPORTED;
    const LossGeom g{B * HW, HW, C};
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case LMNET_F32: return seg_loss_bwd_t<float>(logits, labels, ce_weight, dice_weight, label_smoothing, stats, dloss, dlogits, g, st);
        case LMNET_BF16: return seg_loss_bwd_t<__nv_bfloat16>(logits, labels, ce_weight, dice_weight, label_smoothing, stats, dloss, dlogits, g, st);
        default: return seg_loss_bwd_t<__half>(logits, labels, ce_weight, dice_weight, label_smoothing, stats, dloss, dlogits, g, st);
    }
}
