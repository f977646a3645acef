// common.cuh — device helpers shared by the sm_100a kernels of liblmnet_b200.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/lmnet_b200.h"

namespace lmnet {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// process-wide launch counter (bench.py reports it as gpu_launches)
extern unsigned long long g_launch_count;
inline void count_launch(int n = 1) { __atomic_fetch_add(&g_launch_count, (unsigned long long)n, __ATOMIC_RELAXED); }

// Kernel ids for the optional per-kernel profile (lmnet_profile_*; bench.py's roofline leg).
enum KernelId : int {
    KID_NA_FWD = 0, KID_NA_BWD_QUERY, KID_NA_BWD_KEY, KID_NA_DRPB_REDUCE,
    KID_NA_PN, KID_NA_NN, KID_NA_IN, KID_NA_RPBGRAD, KID_NA_RPBGRAD_REDUCE,
    KID_DW_STATS, KID_DW_FIN_FWD, KID_DW_APPLY, KID_DW_POOL_FIN, KID_DW_COEF_EVAL,
    KID_DW_BWD_REDUCE, KID_DW_FIN_BWD, KID_DW_BWD_DX, KID_DW_BWD_DW, KID_DW_FIN_DW,
    KID_BN_STATS, KID_BN_FIN_FWD, KID_BN_APPLY, KID_BN_BWD_REDUCE, KID_BN_FIN_BWD, KID_BN_BWD_APPLY,
    KID_LN_FWD, KID_LN_BWD, KID_LN_BWD_PARAMS,
    KID_WGRAD_1X1, KID_WGRAD_REDUCE, KID_UPSAMPLE_FWD, KID_UPSAMPLE_BWD,
    KID_NA_STREAM_FWD, KID_NA_STREAM_BWD, KID_PIXEL_GEMM, KID_CONV3X3, KID_CONV3X3_WGRAD, KID_CONV3X3_REDUCE, KID_SE_GATE_FWD, KID_SE_GATE_BWD, KID_AVGPOOL_FWD, KID_AVGPOOL_BWD, KID_DW_BWD_FRAME, KID_SEG_LOSS,
    KID_COUNT
};
extern bool g_profile_on;
void profile_record(int kid, cudaStream_t st, double alg_bytes, bool begin, void** slot);

// Brackets one kernel launch: counts it and, when profiling is on, records CUDA events around it on
// the launching stream (algorithmic bytes of the launch are logged next to the events).
struct LaunchScope {
    int kid; cudaStream_t st; void* slot = nullptr;
    LaunchScope(int kid_, cudaStream_t st_, double alg_bytes) : kid(kid_), st(st_) {
        count_launch();
        if (g_profile_on) profile_record(kid, st, alg_bytes, true, &slot);
    }
    ~LaunchScope() { if (slot != nullptr) profile_record(kid, st, 0.0, false, &slot); }
};

#define LMNET_LAUNCH(kid, st, alg_bytes, ...)                            \
    do {                                                                 \
        { lmnet::LaunchScope _scope((kid), (st), (double)(alg_bytes)); __VA_ARGS__; } \
        if (cudaGetLastError() != cudaSuccess) return LMNET_ERR_LAUNCH;  \
    } while (0)

// Opt-in to > 48 KB of dynamic shared memory, once per kernel, device and size (the attribute is per device;
// granted[dev] = largest size set so far).
constexpr int kMaxDevices = 64;
template <typename Kern> bool ensure_smem(Kern kern, size_t bytes, std::atomic<size_t>* granted_per_device) {
    if (bytes > 227 * 1024) return false;
    if (bytes <= 48 * 1024) return true;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return false;
    std::atomic<size_t>& granted = granted_per_device[dev];
    if (bytes <= granted.load(std::memory_order_relaxed)) return true;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    size_t prev = granted.load(std::memory_order_relaxed);
    while (prev < bytes && !granted.compare_exchange_weak(prev, bytes, std::memory_order_relaxed)) {}
    return true;
}

// ---------------------------------------------------------------------------------------
// element <-> float conversion
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
__device__ __forceinline__ float to_f(__half x) { return __half2float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }
template <> __device__ __forceinline__ __half from_f<__half>(float x) { return __float2half_rn(x); }

// Load / store N contiguous elements as floats, using the widest vector access (<= 16 B) that
// N*sizeof(T) allows when `aligned` (pointer known to be aligned to that width).
template <int BYTES> struct VecOf;
template <> struct VecOf<16> { using type = uint4; };
template <> struct VecOf<8> { using type = uint2; };
template <> struct VecOf<4> { using type = uint32_t; };
template <> struct VecOf<2> { using type = uint16_t; };

template <int N, typename T> __host__ __device__ constexpr int vec_bytes() {
    constexpr int total = N * (int)sizeof(T);
    return (total % 16 == 0) ? 16 : (total % 8 == 0) ? 8 : (total % 4 == 0) ? 4 : (int)sizeof(T) == 2 ? 2 : 4;
}

template <int N, typename T, bool ALIGNED>
__device__ __forceinline__ void load_f(const T* __restrict__ p, float (&dst)[N]) {
    if constexpr (ALIGNED && (vec_bytes<N, T>() > (int)sizeof(T))) {
        constexpr int VB = vec_bytes<N, T>();
        using V = typename VecOf<VB>::type;
        constexpr int PER = VB / (int)sizeof(T);
        constexpr int NV = N / PER;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            V raw = __ldg(reinterpret_cast<const V*>(p) + i);
            const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
            for (int j = 0; j < PER; ++j) dst[i * PER + j] = to_f(e[j]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) dst[i] = to_f(p[i]);
    }
}

template <int N, typename T, bool ALIGNED>
__device__ __forceinline__ void store_f(T* __restrict__ p, const float (&src)[N]) {
    if constexpr (ALIGNED && (vec_bytes<N, T>() > (int)sizeof(T))) {
        constexpr int VB = vec_bytes<N, T>();
        using V = typename VecOf<VB>::type;
        constexpr int PER = VB / (int)sizeof(T);
        constexpr int NV = N / PER;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            V raw;
            T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
            for (int j = 0; j < PER; ++j) e[j] = from_f<T>(src[i * PER + j]);
            reinterpret_cast<V*>(p)[i] = raw;
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) p[i] = from_f<T>(src[i]);
    }
}

// ---------------------------------------------------------------------------------------
// neighbourhood window along one axis, in sub-sequence coordinates (dilation handled by the
// caller as d*d interleaved sub-grids): start = clamp(t - K/2, 0, L - K); rpb index of the
// first neighbour = start - t + K - 1.   
// ---------------------------------------------------------------------------------------
struct AxisWin {
    int start;
    int pb;
};
__device__ __forceinline__ AxisWin axis_window(int t, int L, int K) {
    int st = t - (K >> 1);
    st = max(st, 0);
    st = min(st, L - K);
    return {st, st - t + K - 1};
}
// Inverse neighbourhood: the queries t' whose window contains key t form the contiguous range
// [lo, hi]:  lo = t <= K-1 ? 0 : t - K/2,  hi = t >= L-K ? L-1 : t + K/2.
__device__ __forceinline__ void inverse_window(int t, int L, int K, int& lo, int& hi) {
    lo = (t <= K - 1) ? 0 : t - (K >> 1);
    hi = (t >= L - K) ? L - 1 : t + (K >> 1);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7): one ex2, one rcp and a 5-term polynomial instead
// of erff()'s ~40 instructions.  Returns erf(x) and, as a by-product, exp(-x*x) (the Gaussian that GELU'
// needs as well).  The fused GELU / GELU' evaluate ~360 M of these per training step.
__device__ __forceinline__ float erf_as(float x, float& exp_neg_x2) {
    const float ax = fabsf(x);
    const float t = __fdividef(1.f, fmaf(0.3275911f, ax, 1.f));
    const float e = __expf(-ax * ax);
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    exp_neg_x2 = e;
    return copysignf(fmaf(-p * t, e, 1.f), x);
}
// exact-erf GELU and its derivative (nn.GELU() default, /root/reference/core/modules.py:574)
__device__ __forceinline__ float gelu_fast(float u) {
    float e;
    return 0.5f * u * (1.f + erf_as(u * 0.70710678118654752f, e));
}
__device__ __forceinline__ float gelu_grad_fast(float u) {
    float e;                                   // = exp(-u*u/2)
    const float cdf = 0.5f * (1.f + erf_as(u * 0.70710678118654752f, e));
    return fmaf(u * 0.3989422804014327f, e, cdf);
}

// ---------------------------------------------------------------------------------------
// Packed fp32 pairs (sm_100 FFMA2 / FMUL2: two IEEE fp32 lanes per instruction) and the pair versions of the fused
// GELU / GELU'.  The elementwise epilogues of the depthwise and BatchNorm kernels are bound by instruction issue
// (ncu: 45-65 % issue-slot utilisation at < 25 % of HBM peak); evaluating two elements per FMA and replacing
// __fdividef / __expf (range-checked sequences) by the bare MUFU.RCP / MUFU.EX2 approximations cuts the GELU cost
// from ~31 to ~20 instructions per pair.  Same formula (A&S 7.1.26) and coefficients as erf_as above.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t pk2(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t pk2c(float c) { return pk2(c, c); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// cdf = Phi(u) for two elements, gauss = exp(-u*u/2) (by-product, needed by GELU')
__device__ __forceinline__ void gelu_cdf2(float u0, float u1, uint64_t& cdf, uint64_t& gauss) {
    const float s0 = u0 * 0.70710678118654752f, s1 = u1 * 0.70710678118654752f;
    const uint64_t ax = pk2(fabsf(s0), fabsf(s1));
    const uint64_t den = fma2(ax, pk2c(0.3275911f), pk2c(1.f));
    float d0, d1;
    upk2(den, d0, d1);
    const uint64_t t = pk2(rcp_approx(d0), rcp_approx(d1));
    const uint64_t arg = mul2(mul2(ax, pk2c(-kLog2e)), ax);          // -x*x*log2(e)
    float a0, a1;
    upk2(arg, a0, a1);
    gauss = pk2(ex2_approx(a0), ex2_approx(a1));
    // negated polynomial: q = -(((( a5 t + a4) t + a3) t + a2) t + a1) t, so that erf = 1 + q * gauss
    uint64_t q = fma2(t, pk2c(-1.061405429f), pk2c(1.453152027f));
    q = fma2(q, t, pk2c(-1.421413741f));
    q = fma2(q, t, pk2c(0.284496736f));
    q = fma2(q, t, pk2c(-0.254829592f));
    q = mul2(q, t);
    const uint64_t y = fma2(q, gauss, pk2c(1.f));                   // erf(|x|)
    float y0, y1;
    upk2(y, y0, y1);
    cdf = fma2(pk2(copysignf(y0, s0), copysignf(y1, s1)), pk2c(0.5f), pk2c(0.5f));
}
__device__ __forceinline__ void gelu_fast2(float u0, float u1, float& z0, float& z1) {
    uint64_t cdf, gauss;
    gelu_cdf2(u0, u1, cdf, gauss);
    upk2(mul2(pk2(u0, u1), cdf), z0, z1);
}
__device__ __forceinline__ void gelu_grad_fast2(float u0, float u1, float& g0, float& g1) {
    uint64_t cdf, gauss;
    gelu_cdf2(u0, u1, cdf, gauss);
    upk2(fma2(mul2(pk2(u0, u1), pk2c(0.3989422804014327f)), gauss, cdf), g0, g1);
}

// fp32 storage is the parity configuration (1e-4 relative, masks bit-identical): libdevice erff / expf there
__device__ __forceinline__ float gelu_exact(float u) { return 0.5f * u * (1.f + erff(u * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_exact(float u) {
    const float cdf = 0.5f * (1.f + erff(u * 0.70710678118654752f));
    return fmaf(u * 0.3989422804014327f, expf(-0.5f * u * u), cdf);
}
template <typename T> __device__ __forceinline__ float gelu_t(float u) {
    if constexpr (sizeof(T) == 4) return gelu_exact(u); else return gelu_fast(u);
}
template <typename T> __device__ __forceinline__ float gelu_grad_t(float u) {
    if constexpr (sizeof(T) == 4) return gelu_grad_exact(u); else return gelu_grad_fast(u);
}

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

}  // namespace lmnet
