// bn_act.cu — fused train/eval BatchNorm2d + activation on NCHW tensors (sm_100a), forward and backward.
//
// Widening step f1/f3 of SURVEY.md §8: the BatchNorm + Hardswish that follows ReparamConv's expand
// 1x1 convolution (/root/reference/core/modules.py:537-539, 587) and the BatchNorm + GELU that closes
// the skip-fusion blocks (/root/reference/core/modules.py:97-100, 122-125, 134-137).  In the step
// profile of round 1 ATen's bf16 batch_norm_backward alone was 21 % of the training step
// (profiles/r01_step_profile_a_*.txt).  These are pure bandwidth kernels:
//   train fwd  = stats (read y) -> finalize -> apply (read y, write out)            3*T*es
//   train bwd  = reduce (read y, dout) -> finalize -> apply (read y, dout, write dy) 5*T*es
//   eval  fwd  = apply                                                              2*T*es
// A CTA owns one channel and one chunk of its B*HW elements (16-byte vector loads, fp32 math);
// per-CTA partial sums are reduced by the finalize kernels in a fixed order (deterministic).
// The activation input is rounded to the storage type first, as the reference's separate
// BatchNorm -> activation kernels would see it.
#include "common.cuh"

namespace lmnet {

constexpr int kBnThreads = 256;

struct BnGeom {
    int B, C;
    int64_t HW;
    int chunks;            // CTAs per channel
    int64_t per_chunk;     // elements of one channel handled by one CTA (multiple of the vector width)
};

template <int ACT, typename T> __device__ __forceinline__ float act_fwd(float x) {
    if constexpr (ACT == LMNET_ACT_HARDSWISH) return x * fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f);
    else if constexpr (ACT == LMNET_ACT_GELU) return gelu_t<T>(x);
    else if constexpr (ACT == LMNET_ACT_RELU) return fmaxf(x, 0.f);
    else return x;
}
template <int ACT, typename T> __device__ __forceinline__ float act_bwd(float x) {
    if constexpr (ACT == LMNET_ACT_HARDSWISH) return x < -3.f ? 0.f : (x <= 3.f ? (2.f * x + 3.f) * (1.f / 6.f) : 1.f);
    else if constexpr (ACT == LMNET_ACT_GELU) return gelu_grad_t<T>(x);
    else if constexpr (ACT == LMNET_ACT_RELU) return x > 0.f ? 1.f : 0.f;
    else return 1.f;
}

// elements per lane per iteration: a 16-byte vector on the aligned path, one element otherwise
template <typename T, bool VEC> struct Lane { static constexpr int N = VEC ? 16 / (int)sizeof(T) : 1; };

template <typename T, bool VEC>
__device__ __forceinline__ void ld_elems(const T* __restrict__ p, float (&f)[Lane<T, VEC>::N], int n_valid) {
    constexpr int N = Lane<T, VEC>::N;
    if constexpr (VEC) {
        uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
        const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
        for (int j = 0; j < N; ++j) f[j] = to_f(e[j]);
    } else {
#pragma unroll
        for (int j = 0; j < N; ++j) f[j] = j < n_valid ? to_f(p[j]) : 0.f;
    }
}
template <typename T, bool VEC>
__device__ __forceinline__ void st_elems(T* __restrict__ p, const float (&f)[Lane<T, VEC>::N], int n_valid) {
    constexpr int N = Lane<T, VEC>::N;
    if constexpr (VEC) {
        uint4 raw;
        T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
        for (int j = 0; j < N; ++j) e[j] = from_f<T>(f[j]);
        *reinterpret_cast<uint4*>(p) = raw;
    } else {
#pragma unroll
        for (int j = 0; j < N; ++j)
            if (j < n_valid) p[j] = from_f<T>(f[j]);
    }
}

__device__ __forceinline__ void block_sum2(float& a, float& b, float* s_red /* [2*8] */) {
    a = warp_sum(a);
    b = warp_sum(b);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_red[warp] = a; s_red[8 + warp] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float x = 0.f, y = 0.f;
        for (int w = 0; w < kBnThreads / 32; ++w) { x += s_red[w]; y += s_red[8 + w]; }
        a = x; b = y;
    }
}

// Iterates the vectors of channel c that belong to chunk k: fn(pointer offset, n_valid)
template <typename T, bool VEC, typename Fn>
__device__ __forceinline__ void for_chunk(const BnGeom& g, int c, int k, Fn&& fn) {
    constexpr int N = Lane<T, VEC>::N;
    const int64_t total = (int64_t)g.B * g.HW;                 // elements of this channel
    const int64_t lo = (int64_t)k * g.per_chunk, hi = min(lo + g.per_chunk, total);
    for (int64_t n = lo + (int64_t)threadIdx.x * N; n < hi; n += (int64_t)kBnThreads * N) {
        const int64_t b = n / g.HW, i = n - b * g.HW;          // HW % N == 0 on the vector path => no straddle
        const int64_t off = (b * g.C + c) * g.HW + i;
        const int n_valid = (int)min((int64_t)N, min(hi - n, g.HW - i));
        fn(off, n_valid);
    }
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(kBnThreads)
bnact_stats_kernel(const T* __restrict__ y, float* __restrict__ part /* [C][chunks][2] */, BnGeom g) {
    __shared__ float s_red[16];
    const int c = blockIdx.y, k = blockIdx.x;
    float s = 0.f, ss = 0.f;
    for_chunk<T, VEC>(g, c, k, [&](int64_t off, int nv) {
        float f[Lane<T, VEC>::N];
        ld_elems<T, VEC>(y + off, f, nv);
#pragma unroll
        for (int j = 0; j < Lane<T, VEC>::N; ++j) { s += f[j]; ss = fmaf(f[j], f[j], ss); }
    });
    block_sum2(s, ss, s_red);
    if (threadIdx.x == 0) {
        part[((int64_t)c * g.chunks + k) * 2] = s;
        part[((int64_t)c * g.chunks + k) * 2 + 1] = ss;
    }
}

// coef[c] = (a, b) with out = act(a*y + b)
// one warp per channel: lanes stride over the per-CTA partials, then a fixed-order butterfly (deterministic)
__device__ __forceinline__ double bn_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
constexpr int kBnFinThreads = 128;
static inline int bn_fin_blocks(int C) { return (C + kBnFinThreads / 32 - 1) / (kBnFinThreads / 32); }

__global__ void __launch_bounds__(kBnFinThreads)
bnact_fin_fwd_kernel(const float* __restrict__ part, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float* running_mean, float* running_var,
                     int64_t* nbt, float* __restrict__ save_mean, float* __restrict__ save_rstd,
                     float2* __restrict__ coef, float eps, float momentum, BnGeom g) {
    const int c = blockIdx.x * (kBnFinThreads / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= g.C) return;
    const double n = (double)g.B * (double)g.HW;
    double s = 0, ss = 0;
    const float2* pc = reinterpret_cast<const float2*>(part) + (int64_t)c * g.chunks;
    for (int k = lane; k < g.chunks; k += 32) {
        const float2 v = __ldg(pc + k);
        s += v.x;
        ss += v.y;
    }
    s = bn_warp_sum(s);
    ss = bn_warp_sum(ss);
    if (lane != 0) return;
    const double mean = s / n;
    double var = ss / n - mean * mean;
    if (var < 0) var = 0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    save_mean[c] = (float)mean;
    save_rstd[c] = rstd;
    if (running_mean != nullptr) {
        const double unbiased = n > 1 ? var * n / (n - 1) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
    const float ga = gamma != nullptr ? gamma[c] : 1.f, be = beta != nullptr ? beta[c] : 0.f;
    const float a = ga * rstd;
    coef[c] = make_float2(a, be - a * (float)mean);
    if (c == 0 && nbt != nullptr) *nbt += 1;
}

__global__ void bnact_coef_eval_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                       const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                       float2* __restrict__ coef, float eps, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float ga = gamma != nullptr ? gamma[c] : 1.f, be = beta != nullptr ? beta[c] : 0.f;
    const float a = ga / sqrtf(running_var[c] + eps);
    coef[c] = make_float2(a, be - a * running_mean[c]);
}

template <typename T, bool VEC, int ACT>
__global__ void __launch_bounds__(kBnThreads)
bnact_apply_kernel(const T* __restrict__ y, const float2* __restrict__ coef, T* __restrict__ out, BnGeom g) {
    const int c = blockIdx.y, k = blockIdx.x;
    const float2 ab = __ldg(coef + c);
    for_chunk<T, VEC>(g, c, k, [&](int64_t off, int nv) {
        float f[Lane<T, VEC>::N];
        ld_elems<T, VEC>(y + off, f, nv);
#pragma unroll
        for (int j = 0; j < Lane<T, VEC>::N; ++j) {
            const float pre = to_f(from_f<T>(fmaf(f[j], ab.x, ab.y)));
            f[j] = act_fwd<ACT, T>(pre);
        }
        st_elems<T, VEC>(out + off, f, nv);
    });
}

// sums of dh = dout * act'(pre) and dh * yhat, yhat = (y - mean) * rstd
template <typename T, bool VEC, int ACT>
__global__ void __launch_bounds__(kBnThreads)
bnact_bwd_reduce_kernel(const T* __restrict__ y, const T* __restrict__ dout, const float* __restrict__ gamma,
                        const float* __restrict__ beta, const float* __restrict__ save_mean,
                        const float* __restrict__ save_rstd, float* __restrict__ part, BnGeom g) {
    __shared__ float s_red[16];
    const int c = blockIdx.y, k = blockIdx.x;
    const float mean = save_mean[c], rstd = save_rstd[c];
    const float ga = gamma != nullptr ? gamma[c] : 1.f, be = beta != nullptr ? beta[c] : 0.f;
    float s = 0.f, sy = 0.f;
    for_chunk<T, VEC>(g, c, k, [&](int64_t off, int nv) {
        float f[Lane<T, VEC>::N], d[Lane<T, VEC>::N];
        ld_elems<T, VEC>(y + off, f, nv);
        ld_elems<T, VEC>(dout + off, d, nv);
#pragma unroll
        for (int j = 0; j < Lane<T, VEC>::N; ++j) {
            const float yh = (f[j] - mean) * rstd;
            const float pre = to_f(from_f<T>(fmaf(yh, ga, be)));
            const float dh = d[j] * act_bwd<ACT, T>(pre);
            s += dh;
            sy = fmaf(dh, yh, sy);
        }
    });
    block_sum2(s, sy, s_red);
    if (threadIdx.x == 0) {
        part[((int64_t)c * g.chunks + k) * 2] = s;
        part[((int64_t)c * g.chunks + k) * 2 + 1] = sy;
    }
}

// cb[c] = (m1, m2) = (sum dh / n, sum dh*yhat / n); also writes dgamma, dbeta
__global__ void __launch_bounds__(kBnFinThreads)
bnact_fin_bwd_kernel(const float* __restrict__ part, float* __restrict__ dgamma, float* __restrict__ dbeta,
                     float2* __restrict__ cb, BnGeom g) {
    const int c = blockIdx.x * (kBnFinThreads / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= g.C) return;
    const double n = (double)g.B * (double)g.HW;
    double s = 0, sy = 0;
    const float2* pc = reinterpret_cast<const float2*>(part) + (int64_t)c * g.chunks;
    for (int k = lane; k < g.chunks; k += 32) {
        const float2 v = __ldg(pc + k);
        s += v.x;
        sy += v.y;
    }
    s = bn_warp_sum(s);
    sy = bn_warp_sum(sy);
    if (lane != 0) return;
    if (dgamma != nullptr) dgamma[c] = (float)sy;
    if (dbeta != nullptr) dbeta[c] = (float)s;
    cb[c] = make_float2((float)(s / n), (float)(sy / n));
}

template <typename T, bool VEC, int ACT>
__global__ void __launch_bounds__(kBnThreads)
bnact_bwd_apply_kernel(const T* __restrict__ y, const T* __restrict__ dout, const float* __restrict__ gamma,
                       const float* __restrict__ beta, const float* __restrict__ save_mean,
                       const float* __restrict__ save_rstd, const float2* __restrict__ cb, T* __restrict__ dy, BnGeom g) {
    const int c = blockIdx.y, k = blockIdx.x;
    const float mean = save_mean[c], rstd = save_rstd[c];
    const float ga = gamma != nullptr ? gamma[c] : 1.f, be = beta != nullptr ? beta[c] : 0.f;
    const float2 m = __ldg(cb + c);
    const float gr = ga * rstd;
    for_chunk<T, VEC>(g, c, k, [&](int64_t off, int nv) {
        float f[Lane<T, VEC>::N], d[Lane<T, VEC>::N];
        ld_elems<T, VEC>(y + off, f, nv);
        ld_elems<T, VEC>(dout + off, d, nv);
#pragma unroll
        for (int j = 0; j < Lane<T, VEC>::N; ++j) {
            const float yh = (f[j] - mean) * rstd;
            const float pre = to_f(from_f<T>(fmaf(yh, ga, be)));
            const float dh = d[j] * act_bwd<ACT, T>(pre);
            f[j] = gr * (dh - m.x - yh * m.y);
        }
        st_elems<T, VEC>(dy + off, f, nv);
    });
}

// ================================================================================================
// channels-last ([pixels, C], C contiguous) variants — the layout of cuDNN's 16-bit convolution outputs, so that
// conv -> BatchNorm -> activation -> conv never changes layout.  A thread moves 16-byte vectors of the flat array;
// its stride is a multiple of lcm(C, V) elements, so the V channels of its vector never change: per-channel
// coefficients live in registers and the statistics are V running sums per thread.  Partials are combined in a
// fixed order (thread table in shared memory, then one thread per channel): deterministic.
// ================================================================================================
struct BnGeomCl {
    int C, VP, S;            // channels, vectors per channel period, active threads per CTA (multiple of VP)
    int chunks;
    int64_t nvec, vec_per_chunk;   // total vectors, vectors per CTA (multiple of VP)
};

template <typename T> struct ClVec { static constexpr int V = 16 / (int)sizeof(T); };

template <typename T, typename Fn>
__device__ __forceinline__ void for_chunk_cl(const BnGeomCl& g, Fn&& fn) {
    if ((int)threadIdx.x >= g.S) return;
    const int64_t lo = (int64_t)blockIdx.x * g.vec_per_chunk, hi = min(lo + g.vec_per_chunk, g.nvec);
    for (int64_t v = lo + threadIdx.x; v < hi; v += g.S) fn(v * ClVec<T>::V);
}
// channel of slot j of this thread's vectors
template <typename T> __device__ __forceinline__ int cl_channel(const BnGeomCl& g, int tid, int j) {
    return (int)(((int64_t)(tid % g.VP) * ClVec<T>::V + j) % g.C);
}

// sums the per-thread slot values (two quantities) into per-channel totals; part[c][chunk][2]
template <typename T>
__device__ __forceinline__ void cl_block_reduce(const BnGeomCl& g, const float (&a)[ClVec<T>::V], const float (&b)[ClVec<T>::V],
                                                float* s_tab /* [256][2V] */, float* __restrict__ part) {
    constexpr int V = ClVec<T>::V;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        s_tab[threadIdx.x * 2 * V + j] = a[j];
        s_tab[threadIdx.x * 2 * V + V + j] = b[j];
    }
    __syncthreads();
    // tree over the thread rows, pairing rows a multiple of VP apart (same channel phase); fixed order => deterministic
    constexpr int W2 = 2 * V;
    int n = g.S;                                    // live rows, always a multiple of VP
    while (n > g.VP) {
        const int m = ((n / g.VP + 1) / 2) * g.VP;  // rows [m, n) fold onto rows [0, n - m)
        for (int i = threadIdx.x; i < (n - m) * W2; i += kBnThreads) s_tab[i] += s_tab[i + m * W2];
        __syncthreads();
        n = m;
    }
    for (int c = threadIdx.x; c < g.C; c += kBnThreads) {
        float sa = 0.f, sb = 0.f;
        for (int ph = 0; ph < g.VP; ++ph)
            for (int j = 0; j < V; ++j)
                if ((ph * V + j) % g.C == c) {
                    sa += s_tab[ph * W2 + j];
                    sb += s_tab[ph * W2 + V + j];
                }
        part[((int64_t)c * g.chunks + blockIdx.x) * 2] = sa;
        part[((int64_t)c * g.chunks + blockIdx.x) * 2 + 1] = sb;
    }
}

template <typename T>
__global__ void __launch_bounds__(kBnThreads)
bnact_stats_cl_kernel(const T* __restrict__ y, float* __restrict__ part, BnGeomCl g) {
    constexpr int V = ClVec<T>::V;
    extern __shared__ float s_tab[];
    float s[V], ss[V];
#pragma unroll
    for (int j = 0; j < V; ++j) s[j] = ss[j] = 0.f;
    for_chunk_cl<T>(g, [&](int64_t off) {
        float f[V];
        ld_elems<T, true>(y + off, f, V);
#pragma unroll
        for (int j = 0; j < V; ++j) { s[j] += f[j]; ss[j] = fmaf(f[j], f[j], ss[j]); }
    });
    cl_block_reduce<T>(g, s, ss, s_tab, part);
}

template <typename T, int ACT>
__global__ void __launch_bounds__(kBnThreads)
bnact_apply_cl_kernel(const T* __restrict__ y, const float2* __restrict__ coef, T* __restrict__ out, BnGeomCl g) {
    constexpr int V = ClVec<T>::V;
    float2 ab[V];
#pragma unroll
    for (int j = 0; j < V; ++j) ab[j] = __ldg(coef + cl_channel<T>(g, threadIdx.x, j));
    for_chunk_cl<T>(g, [&](int64_t off) {
        float f[V];
        ld_elems<T, true>(y + off, f, V);
#pragma unroll
        for (int j = 0; j < V; ++j) f[j] = act_fwd<ACT, T>(to_f(from_f<T>(fmaf(f[j], ab[j].x, ab[j].y))));
        st_elems<T, true>(out + off, f, V);
    });
}

template <typename T, int ACT>
__global__ void __launch_bounds__(kBnThreads)
bnact_bwd_reduce_cl_kernel(const T* __restrict__ y, const T* __restrict__ dout, const float* __restrict__ gamma,
                           const float* __restrict__ beta, const float* __restrict__ save_mean,
                           const float* __restrict__ save_rstd, float* __restrict__ part, BnGeomCl g) {
    constexpr int V = ClVec<T>::V;
    extern __shared__ float s_tab[];
    float mean[V], rstd[V], ga[V], be[V], s[V], sy[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const int c = cl_channel<T>(g, threadIdx.x, j);
        mean[j] = save_mean[c]; rstd[j] = save_rstd[c];
        ga[j] = gamma != nullptr ? gamma[c] : 1.f; be[j] = beta != nullptr ? beta[c] : 0.f;
        s[j] = sy[j] = 0.f;
    }
    for_chunk_cl<T>(g, [&](int64_t off) {
        float f[V], d[V];
        ld_elems<T, true>(y + off, f, V);
        ld_elems<T, true>(dout + off, d, V);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float yh = (f[j] - mean[j]) * rstd[j];
            const float pre = to_f(from_f<T>(fmaf(yh, ga[j], be[j])));
            const float dh = d[j] * act_bwd<ACT, T>(pre);
            s[j] += dh;
            sy[j] = fmaf(dh, yh, sy[j]);
        }
    });
    cl_block_reduce<T>(g, s, sy, s_tab, part);
}

template <typename T, int ACT>
__global__ void __launch_bounds__(kBnThreads)
bnact_bwd_apply_cl_kernel(const T* __restrict__ y, const T* __restrict__ dout, const float* __restrict__ gamma,
                          const float* __restrict__ beta, const float* __restrict__ save_mean,
                          const float* __restrict__ save_rstd, const float2* __restrict__ cb, T* __restrict__ dy, BnGeomCl g) {
    constexpr int V = ClVec<T>::V;
    float mean[V], rstd[V], ga[V], be[V];
    float2 m[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const int c = cl_channel<T>(g, threadIdx.x, j);
        mean[j] = save_mean[c]; rstd[j] = save_rstd[c];
        ga[j] = gamma != nullptr ? gamma[c] : 1.f; be[j] = beta != nullptr ? beta[c] : 0.f;
        m[j] = __ldg(cb + c);
    }
    for_chunk_cl<T>(g, [&](int64_t off) {
        float f[V], d[V];
        ld_elems<T, true>(y + off, f, V);
        ld_elems<T, true>(dout + off, d, V);
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float yh = (f[j] - mean[j]) * rstd[j];
            const float pre = to_f(from_f<T>(fmaf(yh, ga[j], be[j])));
            const float dh = d[j] * act_bwd<ACT, T>(pre);
            f[j] = ga[j] * rstd[j] * (dh - m[j].x - yh * m[j].y);
        }
        st_elems<T, true>(dy + off, f, V);
    });
}

static int64_t bn_gcd(int64_t a, int64_t b) { return b == 0 ? a : bn_gcd(b, a % b); }
// chunks is bounded by bn_geom's (the shared workspace layout), so the existing finalize kernels and workspace apply
static bool bn_geom_cl(const lmnet_bn_dims* d, size_t es, BnGeomCl& g) {
    const int64_t V = 16 / (int64_t)es;
    const int64_t total = (int64_t)d->B * d->HW * d->C;
    if (total % V != 0) return false;
    const int64_t L = (int64_t)d->C / bn_gcd(d->C, V) * V;           // lcm(C, V) elements
    g.C = d->C;
    g.VP = (int)(L / V);
    if (g.VP > kBnThreads) return false;
    g.S = kBnThreads / g.VP * g.VP;
    g.nvec = total / V;
    int64_t chunks = 4 * 148;
    const int64_t min_per = (int64_t)g.S * 4;
    if (chunks > (g.nvec + min_per - 1) / min_per) chunks = (g.nvec + min_per - 1) / min_per;
    if (chunks < 1) chunks = 1;
    int64_t per = (g.nvec + chunks - 1) / chunks;
    per = (per + g.VP - 1) / g.VP * g.VP;
    g.vec_per_chunk = per;
    g.chunks = (int)((g.nvec + per - 1) / per);
    return true;
}

// ------------------------------------------------------------------------------------------------
static int bn_validate(const lmnet_bn_dims* d) {
    if (d == nullptr || d->B <= 0 || d->C <= 0 || d->HW <= 0) return LMNET_ERR_INVALID_ARG;
    return LMNET_OK;
}

static BnGeom bn_geom(const lmnet_bn_dims* d, size_t es) {
    BnGeom g;
    g.B = d->B; g.C = d->C; g.HW = d->HW;
    const int64_t total = (int64_t)d->B * d->HW;
    const int64_t vec = 16 / (int64_t)es;
    int chunks = (4 * 148 + d->C - 1) / d->C;                      // ~4 CTAs per SM over the whole grid
    const int64_t min_chunk = (int64_t)kBnThreads * vec * 4;       // at least 4 vectors per thread
    const int64_t max_chunks = (total + min_chunk - 1) / min_chunk;
    if (chunks > max_chunks) chunks = (int)max_chunks;
    if (chunks < 1) chunks = 1;
    int64_t per = (total + chunks - 1) / chunks;
    per = (per + vec - 1) / vec * vec;
    g.per_chunk = per;
    g.chunks = (int)((total + per - 1) / per);
    return g;
}

static bool bn_vec_ok(const lmnet_bn_dims* d, size_t es, std::initializer_list<const void*> ptrs) {
    if ((d->HW * (int64_t)es) % 16 != 0) return false;
    for (auto p : ptrs)
        if (p != nullptr && (uintptr_t)p % 16 != 0) return false;
    return true;
}

struct BnWs {
    size_t part, coef, total;
};
static BnWs bn_ws(const lmnet_bn_dims* d) {
    BnGeom g = bn_geom(d, 2);   // the smaller element size gives the larger chunk count
    BnGeom g4 = bn_geom(d, 4);
    int chunks = g.chunks > g4.chunks ? g.chunks : g4.chunks;
    if (chunks < 4 * 148) chunks = 4 * 148;      // the channels-last kernels use up to 4 x 148 CTAs
    BnWs w;
    w.part = 0;
    w.coef = ((size_t)d->C * chunks * 2 * sizeof(float) + 255) / 256 * 256;
    w.total = w.coef + ((size_t)d->C * sizeof(float2) + 255) / 256 * 256;
    return w;
}

template <typename T, bool VEC, int ACT>
static int bn_launch_fwd(bool train, const void* y, const float* gamma, const float* beta, float* rm, float* rv,
                         int64_t* nbt, void* out, float* save_mean, float* save_rstd, float eps, float momentum,
                         char* ws, const lmnet_bn_dims* d, cudaStream_t st) {
    BnGeom g = bn_geom(d, sizeof(T));
    BnWs L = bn_ws(d);
    float* part = (float*)(ws + L.part);
    float2* coef = (float2*)(ws + L.coef);
    dim3 grid(g.chunks, g.C);
    const double t_bytes = (double)d->B * d->C * d->HW * sizeof(T);
    if (train) {
        LMNET_LAUNCH(KID_BN_STATS, st, t_bytes, (bnact_stats_kernel<T, VEC><<<grid, kBnThreads, 0, st>>>((const T*)y, part, g)));
        LMNET_LAUNCH(KID_BN_FIN_FWD, st, 0, (bnact_fin_fwd_kernel<<<bn_fin_blocks(g.C), kBnFinThreads, 0, st>>>(
            part, gamma, beta, rm, rv, nbt, save_mean, save_rstd, coef, eps, momentum, g)));
    } else {
        LMNET_LAUNCH(KID_BN_FIN_FWD, st, 0, (bnact_coef_eval_kernel<<<(g.C + 127) / 128, 128, 0, st>>>(gamma, beta, rm, rv, coef, eps, g.C)));
    }
    LMNET_LAUNCH(KID_BN_APPLY, st, 2 * t_bytes, (bnact_apply_kernel<T, VEC, ACT><<<grid, kBnThreads, 0, st>>>((const T*)y, coef, (T*)out, g)));
    return LMNET_OK;
}

template <typename T, bool VEC, int ACT>
static int bn_launch_bwd(const void* y, const void* dout, const float* gamma, const float* beta, const float* save_mean,
                         const float* save_rstd, void* dy, float* dgamma, float* dbeta, char* ws,
                         const lmnet_bn_dims* d, cudaStream_t st) {
    BnGeom g = bn_geom(d, sizeof(T));
    BnWs L = bn_ws(d);
    float* part = (float*)(ws + L.part);
    float2* cb = (float2*)(ws + L.coef);
    dim3 grid(g.chunks, g.C);
    const double t_bytes = (double)d->B * d->C * d->HW * sizeof(T);
    LMNET_LAUNCH(KID_BN_BWD_REDUCE, st, 2 * t_bytes, (bnact_bwd_reduce_kernel<T, VEC, ACT><<<grid, kBnThreads, 0, st>>>(
        (const T*)y, (const T*)dout, gamma, beta, save_mean, save_rstd, part, g)));
    LMNET_LAUNCH(KID_BN_FIN_BWD, st, 0, (bnact_fin_bwd_kernel<<<bn_fin_blocks(g.C), kBnFinThreads, 0, st>>>(part, dgamma, dbeta, cb, g)));
    LMNET_LAUNCH(KID_BN_BWD_APPLY, st, 3 * t_bytes, (bnact_bwd_apply_kernel<T, VEC, ACT><<<grid, kBnThreads, 0, st>>>(
        (const T*)y, (const T*)dout, gamma, beta, save_mean, save_rstd, cb, (T*)dy, g)));
    return LMNET_OK;
}

template <typename T, int ACT>
static int bn_launch_fwd_cl(bool train, const void* y, const float* gamma, const float* beta, float* rm, float* rv,
                            int64_t* nbt, void* out, float* save_mean, float* save_rstd, float eps, float momentum,
                            char* ws, const lmnet_bn_dims* d, cudaStream_t st) {
    BnGeomCl gc;
    if (!bn_geom_cl(d, sizeof(T), gc)) return LMNET_ERR_UNSUPPORTED;
    BnGeom g = bn_geom(d, sizeof(T));
    g.chunks = gc.chunks;                         // the finalize kernels sum `chunks` partials per channel
    BnWs L = bn_ws(d);
    float* part = (float*)(ws + L.part);
    float2* coef = (float2*)(ws + L.coef);
    const double t_bytes = (double)d->B * d->C * d->HW * sizeof(T);
    const size_t tab = (size_t)kBnThreads * 2 * ClVec<T>::V * sizeof(float);
    if (train) {
        LMNET_LAUNCH(KID_BN_STATS, st, t_bytes, (bnact_stats_cl_kernel<T><<<gc.chunks, kBnThreads, tab, st>>>((const T*)y, part, gc)));
        LMNET_LAUNCH(KID_BN_FIN_FWD, st, 0, (bnact_fin_fwd_kernel<<<bn_fin_blocks(g.C), kBnFinThreads, 0, st>>>(
            part, gamma, beta, rm, rv, nbt, save_mean, save_rstd, coef, eps, momentum, g)));
    } else {
        LMNET_LAUNCH(KID_BN_FIN_FWD, st, 0, (bnact_coef_eval_kernel<<<(g.C + 127) / 128, 128, 0, st>>>(gamma, beta, rm, rv, coef, eps, g.C)));
    }
    LMNET_LAUNCH(KID_BN_APPLY, st, 2 * t_bytes, (bnact_apply_cl_kernel<T, ACT><<<gc.chunks, kBnThreads, 0, st>>>((const T*)y, coef, (T*)out, gc)));
    return LMNET_OK;
}

template <typename T, int ACT>
static int bn_launch_bwd_cl(const void* y, const void* dout, const float* gamma, const float* beta, const float* save_mean,
                            const float* save_rstd, void* dy, float* dgamma, float* dbeta, char* ws,
                            const lmnet_bn_dims* d, cudaStream_t st) {
    BnGeomCl gc;
    if (!bn_geom_cl(d, sizeof(T), gc)) return LMNET_ERR_UNSUPPORTED;
    BnGeom g = bn_geom(d, sizeof(T));
    g.chunks = gc.chunks;
    BnWs L = bn_ws(d);
    float* part = (float*)(ws + L.part);
    float2* cb = (float2*)(ws + L.coef);
    const double t_bytes = (double)d->B * d->C * d->HW * sizeof(T);
    const size_t tab = (size_t)kBnThreads * 2 * ClVec<T>::V * sizeof(float);
    LMNET_LAUNCH(KID_BN_BWD_REDUCE, st, 2 * t_bytes, (bnact_bwd_reduce_cl_kernel<T, ACT><<<gc.chunks, kBnThreads, tab, st>>>(
        (const T*)y, (const T*)dout, gamma, beta, save_mean, save_rstd, part, gc)));
    LMNET_LAUNCH(KID_BN_FIN_BWD, st, 0, (bnact_fin_bwd_kernel<<<bn_fin_blocks(g.C), kBnFinThreads, 0, st>>>(part, dgamma, dbeta, cb, g)));
    LMNET_LAUNCH(KID_BN_BWD_APPLY, st, 3 * t_bytes, (bnact_bwd_apply_cl_kernel<T, ACT><<<gc.chunks, kBnThreads, 0, st>>>(
        (const T*)y, (const T*)dout, gamma, beta, save_mean, save_rstd, cb, (T*)dy, gc)));
    return LMNET_OK;
}

template <typename T>
static int bn_fwd_cl_t(int act, bool train, const void* y, const float* gamma, const float* beta, float* rm, float* rv,
                       int64_t* nbt, void* out, float* sm, float* sr, float eps, float mom, char* ws,
                       const lmnet_bn_dims* d, cudaStream_t st) {
    switch (act) {
        case LMNET_ACT_NONE: return bn_launch_fwd_cl<T, LMNET_ACT_NONE>(train, y, gamma, beta, rm, rv, nbt, out, sm, sr, eps, mom, ws, d, st);
        case LMNET_ACT_HARDSWISH: return bn_launch_fwd_cl<T, LMNET_ACT_HARDSWISH>(train, y, gamma, beta, rm, rv, nbt, out, sm, sr, eps, mom, ws, d, st);
        case LMNET_ACT_GELU: return bn_launch_fwd_cl<T, LMNET_ACT_GELU>(train, y, gamma, beta, rm, rv, nbt, out, sm, sr, eps, mom, ws, d, st);
        case LMNET_ACT_RELU: return bn_launch_fwd_cl<T, LMNET_ACT_RELU>(train, y, gamma, beta, rm, rv, nbt, out, sm, sr, eps, mom, ws, d, st);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}
template <typename T>
static int bn_bwd_cl_t(int act, const void* y, const void* dout, const float* gamma, const float* beta, const float* sm,
                       const float* sr, void* dy, float* dgamma, float* dbeta, char* ws, const lmnet_bn_dims* d, cudaStream_t st) {
    switch (act) {
        case LMNET_ACT_NONE: return bn_launch_bwd_cl<T, LMNET_ACT_NONE>(y, dout, gamma, beta, sm, sr, dy, dgamma, dbeta, ws, d, st);
        case LMNET_ACT_HARDSWISH: return bn_launch_bwd_cl<T, LMNET_ACT_HARDSWISH>(y, dout, gamma, beta, sm, sr, dy, dgamma, dbeta, ws, d, st);
        case LMNET_ACT_GELU: return bn_launch_bwd_cl<T, LMNET_ACT_GELU>(y, dout, gamma, beta, sm, sr, dy, dgamma, dbeta, ws, d, st);
        case LMNET_ACT_RELU: return bn_launch_bwd_cl<T, LMNET_ACT_RELU>(y, dout, gamma, beta, sm, sr, dy, dgamma, dbeta, ws, d, st);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}

#define BN_DISPATCH_ACT(T, VEC, CALL)                                                      \
    switch (act) {                                                                        \
        case LMNET_ACT_NONE: return CALL<T, VEC, LMNET_ACT_NONE>;                         \
        case LMNET_ACT_HARDSWISH: return CALL<T, VEC, LMNET_ACT_HARDSWISH>;               \
        case LMNET_ACT_GELU: return CALL<T, VEC, LMNET_ACT_GELU>;                         \
        case LMNET_ACT_RELU: return CALL<T, VEC, LMNET_ACT_RELU>;                         \
        default: return nullptr;                                                          \
    }

template <typename T, bool VEC> static auto pick_fwd(int act) -> decltype(&bn_launch_fwd<T, VEC, 0>) { BN_DISPATCH_ACT(T, VEC, &bn_launch_fwd) }
template <typename T, bool VEC> static auto pick_bwd(int act) -> decltype(&bn_launch_bwd<T, VEC, 0>) { BN_DISPATCH_ACT(T, VEC, &bn_launch_bwd) }

template <typename T>
static int bn_fwd_t(bool vec, int act, bool train, const void* y, const float* gamma, const float* beta, float* rm,
                    float* rv, int64_t* nbt, void* out, float* sm, float* sr, float eps, float mom, char* ws,
                    const lmnet_bn_dims* d, cudaStream_t st) {
    auto fn = vec ? pick_fwd<T, true>(act) : pick_fwd<T, false>(act);
    if (fn == nullptr) return LMNET_ERR_UNSUPPORTED;
    return fn(train, y, gamma, beta, rm, rv, nbt, out, sm, sr, eps, mom, ws, d, st);
}
template <typename T>
static int bn_bwd_t(bool vec, int act, const void* y, const void* dout, const float* gamma, const float* beta,
                    const float* sm, const float* sr, void* dy, float* dgamma, float* dbeta, char* ws,
                    const lmnet_bn_dims* d, cudaStream_t st) {
    auto fn = vec ? pick_bwd<T, true>(act) : pick_bwd<T, false>(act);
    if (fn == nullptr) return LMNET_ERR_UNSUPPORTED;
    return fn(y, dout, gamma, beta, sm, sr, dy, dgamma, dbeta, ws, d, st);
}

}  // namespace lmnet

using namespace lmnet;

extern "C" size_t lmnet_bn_act_workspace_bytes(const lmnet_bn_dims* dims) {
    if (bn_validate(dims) != LMNET_OK) return 0;
    return bn_ws(dims).total;
}

extern "C" int lmnet_bn_act_fwd(const void* y, const float* gamma, const float* beta, float* running_mean,
                                float* running_var, int64_t* num_batches_tracked, void* out, float* save_mean,
                                float* save_rstd, float eps, float momentum, int training, int act,
                                void* workspace, size_t workspace_bytes, const lmnet_bn_dims* dims, int dtype,
                                void* stream) {
    int rc = bn_validate(dims);
    if (rc != LMNET_OK) return rc;
    if (!y || !out || !workspace) return LMNET_ERR_INVALID_ARG;
    if (training && (!save_mean || !save_rstd)) return LMNET_ERR_INVALID_ARG;
    if (!training && (!running_mean || !running_var)) return LMNET_ERR_INVALID_ARG;
    if ((running_mean == nullptr) != (running_var == nullptr)) return LMNET_ERR_INVALID_ARG;
    if (workspace_bytes < bn_ws(dims).total) return LMNET_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t es = dtype == LMNET_F32 ? 4 : 2;
    const bool vec = bn_vec_ok(dims, es, {y, out});
    switch (dtype) {
        case LMNET_F32: return bn_fwd_t<float>(vec, act, training != 0, y, gamma, beta, running_mean, running_var, num_batches_tracked, out, save_mean, save_rstd, eps, momentum, (char*)workspace, dims, st);
        case LMNET_BF16: return bn_fwd_t<__nv_bfloat16>(vec, act, training != 0, y, gamma, beta, running_mean, running_var, num_batches_tracked, out, save_mean, save_rstd, eps, momentum, (char*)workspace, dims, st);
        case LMNET_F16: return bn_fwd_t<__half>(vec, act, training != 0, y, gamma, beta, running_mean, running_var, num_batches_tracked, out, save_mean, save_rstd, eps, momentum, (char*)workspace, dims, st);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}

extern "C" int lmnet_bn_act_bwd(const void* y, const void* dout, const float* gamma, const float* beta,
                                const float* save_mean, const float* save_rstd, void* dy, float* dgamma,
                                float* dbeta, int act, void* workspace, size_t workspace_bytes,
                                const lmnet_bn_dims* dims, int dtype, void* stream) {
    int rc = bn_validate(dims);
    if (rc != LMNET_OK) return rc;
    if (!y || !dout || !dy || !save_mean || !save_rstd || !workspace) return LMNET_ERR_INVALID_ARG;
    if (workspace_bytes < bn_ws(dims).total) return LMNET_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t es = dtype == LMNET_F32 ? 4 : 2;
    const bool vec = bn_vec_ok(dims, es, {y, dout, dy});
    switch (dtype) {
        case LMNET_F32: return bn_bwd_t<float>(vec, act, y, dout, gamma, beta, save_mean, save_rstd, dy, dgamma, dbeta, (char*)workspace, dims, st);
        case LMNET_BF16: return bn_bwd_t<__nv_bfloat16>(vec, act, y, dout, gamma, beta, save_mean, save_rstd, dy, dgamma, dbeta, (char*)workspace, dims, st);
        case LMNET_F16: return bn_bwd_t<__half>(vec, act, y, dout, gamma, beta, save_mean, save_rstd, dy, dgamma, dbeta, (char*)workspace, dims, st);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}

// training forward with the statistics pass done by the producer (pixel_gemm's epilogue): finalize + apply only
namespace lmnet {
template <typename T, bool VEC, int ACT>
static int bn_launch_fwd_stats(const void* y, const float* part, int nchunks, const float* gamma, const float* beta, float* rm,
                               float* rv, int64_t* nbt, void* out, float* save_mean, float* save_rstd, float eps, float momentum,
                               char* ws, const lmnet_bn_dims* d, cudaStream_t st) {
    BnGeom g = bn_geom(d, sizeof(T));
    BnWs L = bn_ws(d);
    float2* coef = (float2*)(ws + L.coef);
    dim3 grid(g.chunks, g.C);
    const double t_bytes = (double)d->B * d->C * d->HW * sizeof(T);
    BnGeom gs = g;
    gs.chunks = nchunks;
    LMNET_LAUNCH(KID_BN_FIN_FWD, st, 0, (bnact_fin_fwd_kernel<<<bn_fin_blocks(g.C), kBnFinThreads, 0, st>>>(
        part, gamma, beta, rm, rv, nbt, save_mean, save_rstd, coef, eps, momentum, gs)));
    LMNET_LAUNCH(KID_BN_APPLY, st, 2 * t_bytes, (bnact_apply_kernel<T, VEC, ACT><<<grid, kBnThreads, 0, st>>>((const T*)y, coef, (T*)out, g)));
    return LMNET_OK;
}
template <typename T, bool VEC>
static int bn_fwd_stats_act(int act, const void* y, const float* part, int nchunks, const float* gamma, const float* beta, float* rm,
                            float* rv, int64_t* nbt, void* out, float* sm, float* sr, float eps, float mom, char* ws,
                            const lmnet_bn_dims* d, cudaStream_t st) {
    switch (act) {
        case LMNET_ACT_NONE: return bn_launch_fwd_stats<T, VEC, LMNET_ACT_NONE>(y, part, nchunks, gamma, beta, rm, rv, nbt, out, sm, sr, eps, mom, ws, d, st);
        case LMNET_ACT_HARDSWISH: return bn_launch_fwd_stats<T, VEC, LMNET_ACT_HARDSWISH>(y, part, nchunks, gamma, beta, rm, rv, nbt, out, sm, sr, eps, mom, ws, d, st);
        case LMNET_ACT_GELU: return bn_launch_fwd_stats<T, VEC, LMNET_ACT_GELU>(y, part, nchunks, gamma, beta, rm, rv, nbt, out, sm, sr, eps, mom, ws, d, st);
        case LMNET_ACT_RELU: return bn_launch_fwd_stats<T, VEC, LMNET_ACT_RELU>(y, part, nchunks, gamma, beta, rm, rv, nbt, out, sm, sr, eps, mom, ws, d, st);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}
template <typename T>
static int bn_fwd_stats_t(bool vec, int act, const void* y, const float* part, int nchunks, const float* gamma, const float* beta,
                          float* rm, float* rv, int64_t* nbt, void* out, float* sm, float* sr, float eps, float mom, char* ws,
                          const lmnet_bn_dims* d, cudaStream_t st) {
    return vec ? bn_fwd_stats_act<T, true>(act, y, part, nchunks, gamma, beta, rm, rv, nbt, out, sm, sr, eps, mom, ws, d, st)
               : bn_fwd_stats_act<T, false>(act, y, part, nchunks, gamma, beta, rm, rv, nbt, out, sm, sr, eps, mom, ws, d, st);
}
}  // namespace lmnet

extern "C" int lmnet_bn_act_fwd_stats(const void* y, const float* stats_part, int nchunks, const float* gamma,
                                      const float* beta, float* running_mean, float* running_var,
                                      int64_t* num_batches_tracked, void* out, float* save_mean, float* save_rstd,
                                      float eps, float momentum, int act, void* workspace, size_t workspace_bytes,
                                      const lmnet_bn_dims* dims, int dtype, void* stream) {
    int rc = bn_validate(dims);
    if (rc != LMNET_OK) return rc;
    if (!y || !out || !workspace || !stats_part || nchunks <= 0 || !save_mean || !save_rstd) return LMNET_ERR_INVALID_ARG;
    if ((running_mean == nullptr) != (running_var == nullptr)) return LMNET_ERR_INVALID_ARG;
    if (workspace_bytes < bn_ws(dims).total) return LMNET_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t es = dtype == LMNET_F32 ? 4 : 2;
    const bool vec = bn_vec_ok(dims, es, {y, out});
    switch (dtype) {
        case LMNET_F32: return bn_fwd_stats_t<float>(vec, act, y, stats_part, nchunks, gamma, beta, running_mean, running_var, num_batches_tracked, out, save_mean, save_rstd, eps, momentum, (char*)workspace, dims, st);
        case LMNET_BF16: return bn_fwd_stats_t<__nv_bfloat16>(vec, act, y, stats_part, nchunks, gamma, beta, running_mean, running_var, num_batches_tracked, out, save_mean, save_rstd, eps, momentum, (char*)workspace, dims, st);
        case LMNET_F16: return bn_fwd_stats_t<__half>(vec, act, y, stats_part, nchunks, gamma, beta, running_mean, running_var, num_batches_tracked, out, save_mean, save_rstd, eps, momentum, (char*)workspace, dims, st);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}

// ---- channels-last entry points: y / out / dout / dy are [B, HW, C] (C contiguous); same workspace query ----
extern "C" int lmnet_bn_act_cl_supported(const lmnet_bn_dims* dims, int dtype) {
    if (bn_validate(dims) != LMNET_OK) return 0;
    BnGeomCl g;
    return bn_geom_cl(dims, dtype == LMNET_F32 ? 4 : 2, g) ? 1 : 0;
}

extern "C" int lmnet_bn_act_cl_fwd(const void* y, const float* gamma, const float* beta, float* running_mean,
                                   float* running_var, int64_t* num_batches_tracked, void* out, float* save_mean,
                                   float* save_rstd, float eps, float momentum, int training, int act,
                                   void* workspace, size_t workspace_bytes, const lmnet_bn_dims* dims, int dtype,
                                   void* stream) {
    int rc = bn_validate(dims);
    if (rc != LMNET_OK) return rc;
    if (!y || !out || !workspace) return LMNET_ERR_INVALID_ARG;
    if (training && (!save_mean || !save_rstd)) return LMNET_ERR_INVALID_ARG;
    if (!training && (!running_mean || !running_var)) return LMNET_ERR_INVALID_ARG;
    if ((running_mean == nullptr) != (running_var == nullptr)) return LMNET_ERR_INVALID_ARG;
    if (workspace_bytes < bn_ws(dims).total) return LMNET_ERR_WORKSPACE;
    if ((uintptr_t)y % 16 || (uintptr_t)out % 16) return LMNET_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case LMNET_F32: return bn_fwd_cl_t<float>(act, training != 0, y, gamma, beta, running_mean, running_var, num_batches_tracked, out, save_mean, save_rstd, eps, momentum, (char*)workspace, dims, st);
        case LMNET_BF16: return bn_fwd_cl_t<__nv_bfloat16>(act, training != 0, y, gamma, beta, running_mean, running_var, num_batches_tracked, out, save_mean, save_rstd, eps, momentum, (char*)workspace, dims, st);
        case LMNET_F16: return bn_fwd_cl_t<__half>(act, training != 0, y, gamma, beta, running_mean, running_var, num_batches_tracked, out, save_mean, save_rstd, eps, momentum, (char*)workspace, dims, st);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}

extern "C" int lmnet_bn_act_cl_bwd(const void* y, const void* dout, const float* gamma, const float* beta,
                                   const float* save_mean, const float* save_rstd, void* dy, float* dgamma,
                                   float* dbeta, int act, void* workspace, size_t workspace_bytes,
                                   const lmnet_bn_dims* dims, int dtype, void* stream) {
    int rc = bn_validate(dims);
    if (rc != LMNET_OK) return rc;
    if (!y || !dout || !dy || !save_mean || !save_rstd || !workspace) return LMNET_ERR_INVALID_ARG;
    if (workspace_bytes < bn_ws(dims).total) return LMNET_ERR_WORKSPACE;
    if ((uintptr_t)y % 16 || (uintptr_t)dout % 16 || (uintptr_t)dy % 16) return LMNET_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case LMNET_F32: return bn_bwd_cl_t<float>(act, y, dout, gamma, beta, save_mean, save_rstd, dy, dgamma, dbeta, (char*)workspace, dims, st);
        case LMNET_BF16: return bn_bwd_cl_t<__nv_bfloat16>(act, y, dout, gamma, beta, save_mean, save_rstd, dy, dgamma, dbeta, (char*)workspace, dims, st);
        case LMNET_F16: return bn_bwd_cl_t<__half>(act, y, dout, gamma, beta, save_mean, save_rstd, dy, dgamma, dbeta, (char*)workspace, dims, st);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}
