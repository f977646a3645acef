// layer_norm.cu — LayerNorm over a short channels-last row (C = 12 / 24 / 48 / 96), forward + backward.
//
// Widening step f2 of SURVEY.md §8: the two LayerNorms of NeighborhoodTransformer
// (/root/reference/core/modules.py:507, 510, 515-518) normalise rows of only 12..96 channels at up to
// 16*352*352 = 2 M rows.  ATen assigns a warp-or-more per row, which wastes > 60 % of its lanes on
// such rows and runs in fp32 with separate cast kernels under autocast (round-1 step profile: 12 ms
// of the step).  Here G = 1, 2 or 4 lanes own a row (12 or 24 elements per lane, 8/16-byte vector
// accesses, fp32 math, warp-shuffle row statistics), the output is written in the input's storage
// type, and the backward produces dx plus per-CTA partial sums of dgamma / dbeta that a second small
// kernel reduces in a fixed order.   Bytes: fwd 2*N*es, bwd 3*N*es (N = rows*C).
#include "common.cuh"

namespace lmnet {

constexpr int kLnThreads = 256;

template <typename T, int EPL>
__device__ __forceinline__ void ln_load(const T* __restrict__ p, float (&f)[EPL]) {
    constexpr int VB = (EPL * (int)sizeof(T)) % 16 == 0 ? 16 : 8;
    using V = typename VecOf<VB>::type;
    constexpr int PER = VB / (int)sizeof(T);
#pragma unroll
    for (int i = 0; i < EPL / PER; ++i) {
        V raw = __ldg(reinterpret_cast<const V*>(p) + i);
        const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
        for (int j = 0; j < PER; ++j) f[i * PER + j] = to_f(e[j]);
    }
}
template <typename T, int EPL>
__device__ __forceinline__ void ln_store(T* __restrict__ p, const float (&f)[EPL]) {
    constexpr int VB = (EPL * (int)sizeof(T)) % 16 == 0 ? 16 : 8;
    using V = typename VecOf<VB>::type;
    constexpr int PER = VB / (int)sizeof(T);
#pragma unroll
    for (int i = 0; i < EPL / PER; ++i) {
        V raw;
        T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
        for (int j = 0; j < PER; ++j) e[j] = from_f<T>(f[i * PER + j]);
        reinterpret_cast<V*>(p)[i] = raw;
    }
}
template <int G> __device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// y = (x - mean) * rstd * gamma + beta per row; mean / rstd saved as fp32 [rows] (may be NULL)
template <typename T, int EPL, int G>
__global__ void __launch_bounds__(kLnThreads)
ln_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              T* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int64_t rows, float eps) {
    constexpr int C = EPL * G;
    const int sub = threadIdx.x % G;
    float ga[EPL], be[EPL];
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
        ga[j] = gamma != nullptr ? __ldg(gamma + sub * EPL + j) : 1.f;
        be[j] = beta != nullptr ? __ldg(beta + sub * EPL + j) : 0.f;
    }
    const int64_t rows_per_block = kLnThreads / G;
    for (int64_t r0 = (int64_t)blockIdx.x * rows_per_block; r0 < rows; r0 += (int64_t)gridDim.x * rows_per_block) {
        const int64_t r = r0 + threadIdx.x / G;
        const bool ok = r < rows;                       // lanes of one row agree; shuffles below stay convergent
        float f[EPL];
        if (ok) ln_load<T, EPL>(x + r * C + sub * EPL, f);
        else {
#pragma unroll
            for (int j = 0; j < EPL; ++j) f[j] = 0.f;
        }
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < EPL; ++j) s += f[j];
        const float mean = group_sum<G>(s) * (1.f / C);
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < EPL; ++j) { const float d = f[j] - mean; v = fmaf(d, d, v); }
        const float rstd = rsqrtf(group_sum<G>(v) * (1.f / C) + eps);
        if (ok) {
#pragma unroll
            for (int j = 0; j < EPL; ++j) f[j] = fmaf((f[j] - mean) * rstd, ga[j], be[j]);
            ln_store<T, EPL>(y + r * C + sub * EPL, f);
            if (sub == 0 && mean_out != nullptr) { mean_out[r] = mean; rstd_out[r] = rstd; }
        }
    }
}

// dx = rstd * (dy*gamma - mean_c(dy*gamma) - xhat * mean_c(dy*gamma*xhat));  per-CTA partial sums of
// dgamma = sum_r dy*xhat and dbeta = sum_r dy into part[blockIdx.x][2][C]
template <typename T, int EPL, int G>
__global__ void __launch_bounds__(kLnThreads)
ln_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, const float* __restrict__ gamma,
              const float* __restrict__ mean_in, const float* __restrict__ rstd_in, T* __restrict__ dx,
              float* __restrict__ part, int64_t rows) {
    constexpr int C = EPL * G;
    __shared__ float s_part[2 * C];
    for (int i = threadIdx.x; i < 2 * C; i += kLnThreads) s_part[i] = 0.f;
    __syncthreads();
    const int sub = threadIdx.x % G;
    float ga[EPL], dg[EPL], db[EPL];
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
        ga[j] = gamma != nullptr ? __ldg(gamma + sub * EPL + j) : 1.f;
        dg[j] = db[j] = 0.f;
    }
    const int64_t rows_per_block = kLnThreads / G;
    for (int64_t r0 = (int64_t)blockIdx.x * rows_per_block; r0 < rows; r0 += (int64_t)gridDim.x * rows_per_block) {
        const int64_t r = r0 + threadIdx.x / G;
        const bool ok = r < rows;
        float f[EPL], d[EPL];
        float mean = 0.f, rstd = 0.f;
        if (ok) {
            ln_load<T, EPL>(x + r * C + sub * EPL, f);
            ln_load<T, EPL>(dy + r * C + sub * EPL, d);
            mean = __ldg(mean_in + r);
            rstd = __ldg(rstd_in + r);
        } else {
#pragma unroll
            for (int j = 0; j < EPL; ++j) f[j] = d[j] = 0.f;
        }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
            f[j] = (f[j] - mean) * rstd;                 // xhat
            dg[j] = fmaf(d[j], f[j], dg[j]);
            db[j] += d[j];
            d[j] *= ga[j];                               // dy*gamma
            s1 += d[j];
            s2 = fmaf(d[j], f[j], s2);
        }
        s1 = group_sum<G>(s1) * (1.f / C);
        s2 = group_sum<G>(s2) * (1.f / C);
        if (ok) {
#pragma unroll
            for (int j = 0; j < EPL; ++j) f[j] = rstd * (d[j] - s1 - f[j] * s2);
            ln_store<T, EPL>(dx + r * C + sub * EPL, f);
        }
    }
    // reduce the per-thread parameter sums: lanes with equal `sub` hold the same channels
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
#pragma unroll
        for (int o = 16; o >= G; o >>= 1) {
            dg[j] += __shfl_xor_sync(0xffffffffu, dg[j], o);
            db[j] += __shfl_xor_sync(0xffffffffu, db[j], o);
        }
    }
    if ((threadIdx.x & 31) < G) {
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
            atomicAdd(&s_part[sub * EPL + j], dg[j]);
            atomicAdd(&s_part[C + sub * EPL + j], db[j]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += kLnThreads) part[(int64_t)blockIdx.x * 2 * C + i] = s_part[i];
}

// block = 32 columns x 8 row groups; each thread sums every 8th partial row of its column, then the 8 row groups
// are combined through shared memory (fixed order)
__global__ void __launch_bounds__(256)
ln_bwd_params_kernel(const float* __restrict__ part, int nblocks, int C, float* __restrict__ dgamma,
                     float* __restrict__ dbeta) {
    __shared__ float s[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + cx;
    float a = 0.f;
    if (i < 2 * C)
        for (int b = ry; b < nblocks; b += 8) a += part[(int64_t)b * 2 * C + i];
    s[ry][cx] = a;
    __syncthreads();
    if (ry == 0 && i < 2 * C) {
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) t += s[r][cx];
        if (i < C) { if (dgamma != nullptr) dgamma[i] = t; }
        else if (dbeta != nullptr) dbeta[i - C] = t;
    }
}

static int ln_blocks(int64_t rows, int G) {
    const int64_t rows_per_block = kLnThreads / G;
    int64_t need = (rows + rows_per_block - 1) / rows_per_block;
    const int64_t cap = 148 * 8;
    return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}

template <typename T, int EPL, int G>
static int ln_fwd_launch(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                         int64_t rows, float eps, cudaStream_t st) {
    const double bytes = 2.0 * (double)rows * EPL * G * sizeof(T);
    LMNET_LAUNCH(KID_LN_FWD, st, bytes, (ln_fwd_kernel<T, EPL, G><<<ln_blocks(rows, G), kLnThreads, 0, st>>>(
        (const T*)x, gamma, beta, (T*)y, mean, rstd, rows, eps)));
    return LMNET_OK;
}
template <typename T, int EPL, int G>
static int ln_bwd_launch(const void* x, const void* dy, const float* gamma, const float* mean, const float* rstd,
                         void* dx, float* dgamma, float* dbeta, float* part, int64_t rows, cudaStream_t st) {
    constexpr int C = EPL * G;
    const int nb = ln_blocks(rows, G);
    const double bytes = 3.0 * (double)rows * C * sizeof(T);
    LMNET_LAUNCH(KID_LN_BWD, st, bytes, (ln_bwd_kernel<T, EPL, G><<<nb, kLnThreads, 0, st>>>(
        (const T*)x, (const T*)dy, gamma, mean, rstd, (T*)dx, part, rows)));
    LMNET_LAUNCH(KID_LN_BWD_PARAMS, st, 0, (ln_bwd_params_kernel<<<(2 * C + 31) / 32, 256, 0, st>>>(part, nb, C, dgamma, dbeta)));
    return LMNET_OK;
}

static bool ln_supported(int C) { return C == 12 || C == 24 || C == 48 || C == 96; }

}  // namespace lmnet

using namespace lmnet;

extern "C" int lmnet_layer_norm_supported(int C) { return ln_supported(C) ? 1 : 0; }

extern "C" size_t lmnet_layer_norm_workspace_bytes(int64_t rows, int C) {
    if (rows <= 0 || !ln_supported(C)) return 0;
    return (size_t)148 * 8 * 2 * C * sizeof(float);
}

#define LN_DISPATCH(T, FN, ...)                                              \
    switch (C) {                                                             \
        case 12: return FN<T, 12, 1>(__VA_ARGS__);                           \
        case 24: return FN<T, 24, 1>(__VA_ARGS__);                           \
        case 48: return FN<T, 24, 2>(__VA_ARGS__);                           \
        case 96: return FN<T, 24, 4>(__VA_ARGS__);                           \
        default: return LMNET_ERR_UNSUPPORTED;                               \
    }

template <typename T>
static int ln_fwd_t(int C, const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                    int64_t rows, float eps, cudaStream_t st) {
    LN_DISPATCH(T, ln_fwd_launch, x, gamma, beta, y, mean, rstd, rows, eps, st)
}
template <typename T>
static int ln_bwd_t(int C, const void* x, const void* dy, const float* gamma, const float* mean, const float* rstd,
                    void* dx, float* dgamma, float* dbeta, float* part, int64_t rows, cudaStream_t st) {
    LN_DISPATCH(T, ln_bwd_launch, x, dy, gamma, mean, rstd, dx, dgamma, dbeta, part, rows, st)
}

extern "C" int lmnet_layer_norm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* save_mean,
                                    float* save_rstd, int64_t rows, int C, float eps, int dtype, void* stream) {
    if (!x || !y || rows <= 0) return LMNET_ERR_INVALID_ARG;
    if ((save_mean == nullptr) != (save_rstd == nullptr)) return LMNET_ERR_INVALID_ARG;
    if (!ln_supported(C)) return LMNET_ERR_UNSUPPORTED;
    if ((uintptr_t)x % 16 != 0 || (uintptr_t)y % 16 != 0) return LMNET_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case LMNET_F32: return ln_fwd_t<float>(C, x, gamma, beta, y, save_mean, save_rstd, rows, eps, st);
        case LMNET_BF16: return ln_fwd_t<__nv_bfloat16>(C, x, gamma, beta, y, save_mean, save_rstd, rows, eps, st);
        case LMNET_F16: return ln_fwd_t<__half>(C, x, gamma, beta, y, save_mean, save_rstd, rows, eps, st);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}

extern "C" int lmnet_layer_norm_bwd(const void* x, const void* dy, const float* gamma, const float* save_mean,
                                    const float* save_rstd, void* dx, float* dgamma, float* dbeta,
                                    void* workspace, size_t workspace_bytes, int64_t rows, int C, int dtype,
                                    void* stream) {
    if (!x || !dy || !dx || !save_mean || !save_rstd || !workspace || rows <= 0) return LMNET_ERR_INVALID_ARG;
    if (!ln_supported(C)) return LMNET_ERR_UNSUPPORTED;
    if (workspace_bytes < lmnet_layer_norm_workspace_bytes(rows, C)) return LMNET_ERR_WORKSPACE;
    if ((uintptr_t)x % 16 != 0 || (uintptr_t)dy % 16 != 0 || (uintptr_t)dx % 16 != 0) return LMNET_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    float* part = (float*)workspace;
    switch (dtype) {
        case LMNET_F32: return ln_bwd_t<float>(C, x, dy, gamma, save_mean, save_rstd, dx, dgamma, dbeta, part, rows, st);
        case LMNET_BF16: return ln_bwd_t<__nv_bfloat16>(C, x, dy, gamma, save_mean, save_rstd, dx, dgamma, dbeta, part, rows, st);
        case LMNET_F16: return ln_bwd_t<__half>(C, x, dy, gamma, save_mean, save_rstd, dx, dgamma, dbeta, part, rows, st);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}
