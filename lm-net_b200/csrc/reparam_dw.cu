// reparam_dw.cu — fused depthwise multi-branch conv + BatchNorm + sum + GELU (+ SE pooling), sm_100a.
//
// Replaces the middle of ReparamConv.forward, /root/reference/core/modules.py:592-597:
//     out  = BN(dw5x5(x)) + BN(dw3x3(x)) + BN(dw3x1(x)) + BN(dw1x3(x));  z = GELU(out)
// plus the global average pool that opens SE.forward (/root/reference/core/modules.py:1030-1031),
// forward and backward, training (batch statistics) and inference (running statistics, i.e. the
// algebra of get_equivalent_kernel_bias, /root/reference/core/modules.py:622-642).
//
// Design.  The four branches are linear in x, so once the batch statistics are
// known the whole sum collapses to ONE 5x5 depthwise kernel + bias per channel:
//     u = (sum_br a_br * embed5x5(w_br)) (*) x + sum_br (beta_br - a_br*mean_br),  a_br = gamma_br*rstd_br
// Training forward = stats pass (4 branch outputs in registers, only sum / sum-of-squares leave the
// SM) -> per-channel finalize (mean, rstd, running update, merged kernel) -> apply pass (25-tap
// stencil + exact-erf GELU, writes u and z once, pool partial sums ride on the writer).
// Backward: R pass (du = dz*gelu'(u); P[t] = sum du(p) x(p+t), 25 lags shared by all branches)
// -> finalize (dgamma, dbeta, per-branch coefficients) -> A1 pass (dx) -> A2 pass (weight grads).
//
// Arithmetic: fp32 throughout, packed as FFMA2 (fma.rn.f32x2, new on sm_100): a thread owns two
// adjacent output columns as one f32x2 lane pair; the scalar tap weight is the broadcast operand.
// The input tile is staged in shared memory twice (natural and shifted by one element) so that
// every (col, col+1) pair is an aligned 64-bit shared load.  A CTA owns one channel, one column
// stripe and one band of rows, and loops over the batch, so per-channel reductions need no atomics:
// per-CTA partials are reduced by the finalize kernels in a fixed order (deterministic).
#include <algorithm>
#include <initializer_list>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace lmnet {

struct f2 {
    float x, y;
};
__device__ __forceinline__ f2 mk2(float a, float b) { return f2{a, b}; }
__device__ __forceinline__ f2 ffma2(f2 a, f2 b, f2 c) {
    f2 d;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ f2 ffma2(f2 a, float w, f2 c) { return ffma2(a, mk2(w, w), c); }
__device__ __forceinline__ f2 ld2(const float* p) {
    float2 v = *reinterpret_cast<const float2*>(p);
    return f2{v.x, v.y};
}

constexpr int kDwThreads = 128;
constexpr int kDwWarps = 4;
constexpr int kPad = 8;                  // halo columns staged on each side (>= 4, multiple of the vector width)
constexpr int kPitch = 64 + 2 * kPad;    // floats per shared-memory tile row

struct DwGeom {
    int B, E, H, W;
    int stripes, bands, rows_per_band;  // CTA grid = (stripes, bands, E)
};

// Stages ROWS x kPitch elements of one plane into shared memory as floats:
//   sA[r][i] = x(r0 + r, c0 + i)  (0 outside the image).
// Split in two halves so that the global loads of tile t+1 are in flight while tile t is computed:
// fetch() issues all (vectorised, VEC <= 4 elements) loads into registers, commit() converts and
// writes one VEC-float vector per lane (consecutive lanes -> consecutive vectors: conflict-free).
// VEC > 1 requires W % VEC == 0, c0 % VEC == 0 and a plane pointer aligned to VEC elements.
template <typename T, int VEC, int ROWS>
struct TileLoader {
    static constexpr int NV = kPitch / VEC;                                  // vectors per row
    static constexpr int N = (ROWS * NV + kDwThreads - 1) / kDwThreads;      // vectors per thread
    using V = typename VecOf<VEC * (int)sizeof(T)>::type;
    V raw[N];

    __device__ __forceinline__ void fetch(const T* __restrict__ plane, int H, int W, int r0, int c0) {
#pragma unroll
        for (int u = 0; u < N; ++u) {
            const int idx = threadIdx.x + u * kDwThreads;
            const int r = idx / NV, v = idx - r * NV;
            const int gr = r0 + r, gc = c0 + v * VEC;
            V val{};
            if (idx < ROWS * NV && gr >= 0 && gr < H && gc >= 0 && gc + VEC <= W)
                val = __ldg(reinterpret_cast<const V*>(plane + (int64_t)gr * W + gc));
            raw[u] = val;
        }
    }
    __device__ __forceinline__ void commit(float* sA) const {
#pragma unroll
        for (int u = 0; u < N; ++u) {
            const int idx = threadIdx.x + u * kDwThreads;
            if (idx < ROWS * NV) {
                const T* e = reinterpret_cast<const T*>(&raw[u]);
                float f[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j) f[j] = to_f(e[j]);
                using FV = typename VecOf<VEC * 4>::type;
                *reinterpret_cast<FV*>(sA + idx * VEC) = *reinterpret_cast<const FV*>(f);   // idx*VEC == r*kPitch + v*VEC
            }
        }
    }
};

// Same idea for an un-haloed ROWS x 64 tile kept in its storage type (u, dz, du): element (r, c) of the
// tile lands at s[r * kRawPitch + c + shift].  Out-of-image elements are zero.
constexpr int kRawPitch = 72;
template <typename T, int VEC, int ROWS>
struct RawTileLoader {
    static constexpr int NV = 64 / VEC;
    static constexpr int N = (ROWS * NV + kDwThreads - 1) / kDwThreads;
    using V = typename VecOf<VEC * (int)sizeof(T)>::type;
    V raw[N];
    __device__ __forceinline__ void fetch(const T* __restrict__ plane, int H, int W, int r0, int c0) {
#pragma unroll
        for (int u = 0; u < N; ++u) {
            const int idx = threadIdx.x + u * kDwThreads;
            const int r = idx / NV, v = idx - r * NV;
            const int gr = r0 + r, gc = c0 + v * VEC;
            V val{};
            if (idx < ROWS * NV && gr >= 0 && gr < H && gc >= 0 && gc + VEC <= W)
                val = __ldg(reinterpret_cast<const V*>(plane + (int64_t)gr * W + gc));
            raw[u] = val;
        }
    }
    __device__ __forceinline__ void commit(T* s) const {
#pragma unroll
        for (int u = 0; u < N; ++u) {
            const int idx = threadIdx.x + u * kDwThreads;
            if (idx < ROWS * NV) {
                const int r = idx / NV, v = idx - r * NV;
                *reinterpret_cast<V*>(s + r * kRawPitch + v * VEC) = raw[u];
            }
        }
    }
};

// Runs body(b, tr, last_tile_of_b) over every (batch item, row tile) of this CTA's band, with the
// next tile's global loads prefetched into registers while the current tile is being computed.
// The staged x tile starts `halo` rows above tr and kPad columns left of the stripe origin c0.
// `extra` is an object with fetch(b, tr) / commit() for further operands staged the same way.
struct NoExtra {
    __device__ __forceinline__ void fetch(int, int) {}
    __device__ __forceinline__ void commit() {}
};
template <typename T, int VEC, int ROWS, typename PlaneFn, typename Extra, typename Body>
__device__ __forceinline__ void for_each_tile(const DwGeom& g, int band0, int band1, int th, int halo, int c0,
                                              float* sA, PlaneFn&& plane_of, Extra&& extra, Body&& body) {
    const int ntr = (band1 - band0 + th - 1) / th;
    const int total = band1 > band0 ? g.B * ntr : 0;
    TileLoader<T, VEC, ROWS> ld;
    if (total > 0) {
        ld.fetch(plane_of(0), g.H, g.W, band0 - halo, c0 - kPad);
        extra.fetch(0, band0);
    }
    for (int t = 0; t < total; ++t) {
        const int b = t / ntr, k = t - b * ntr;
        __syncthreads();
        ld.commit(sA);
        extra.commit();
        __syncthreads();
        if (t + 1 < total) {
            const int nb = (t + 1) / ntr, nk = (t + 1) - nb * ntr;
            ld.fetch(plane_of(nb), g.H, g.W, band0 + nk * th - halo, c0 - kPad);
            extra.fetch(nb, band0 + nk * th);
        }
        body(b, band0 + k * th, k == ntr - 1);
    }
}

// 5x5 window walk.  The thread owns output columns (2*lane, 2*lane+1) of the tile and RT consecutive
// rows starting at tile row `row0`; shared row `row0 + o + a`, pair offset b holds the inputs for
// tap (a, b) of output row o.  Three aligned 64-bit shared loads fetch the six inputs of a row; the
// two odd-aligned pairs are assembled in registers.  fn(o, win) is called once per output row.
__device__ __forceinline__ void load_row5(const float* p, f2 (&dst)[5]) {
    const f2 a0 = ld2(p), a1 = ld2(p + 2), a2 = ld2(p + 4);
    dst[0] = a0; dst[1] = mk2(a0.y, a1.x); dst[2] = a1; dst[3] = mk2(a1.y, a2.x); dst[4] = a2;
}
template <int RT, typename Fn>
__device__ __forceinline__ void walk5(const float* sA, int row0, int col0, Fn&& fn, int pitch = kPitch) {
    f2 rows[5][5];
#pragma unroll
    for (int r = 0; r < 4; ++r) load_row5(sA + (row0 + r) * pitch + col0, rows[r]);
#pragma unroll
    for (int o = 0; o < RT; ++o) {
        load_row5(sA + (row0 + o + 4) * pitch + col0, rows[(o + 4) % 5]);
        f2 win[5][5];
#pragma unroll
        for (int a = 0; a < 5; ++a)
#pragma unroll
            for (int b = 0; b < 5; ++b) win[a][b] = rows[(o + a) % 5][b];
        fn(o, win);
    }
}

// 3x3 window walk over the centre of the same geometry: win[a][b] = input at (row o+a-1, col +b-1)
// relative to the output pixel; `row0` is the tile row of the first output row MINUS 1 and col0 the
// same even column index as walk5 uses (the 3-wide window starts at col0+1).
__device__ __forceinline__ void load_row3(const float* p, f2 (&dst)[3]) {
    const f2 a0 = ld2(p), a1 = ld2(p + 2), a2 = ld2(p + 4);
    dst[0] = mk2(a0.y, a1.x); dst[1] = a1; dst[2] = mk2(a1.y, a2.x);
}
template <int RT, typename Fn>
__device__ __forceinline__ void walk3(const float* sA, int row0, int col0, Fn&& fn, int pitch = kPitch) {
    f2 rows[3][3];
#pragma unroll
    for (int r = 0; r < 2; ++r) load_row3(sA + (row0 + r) * pitch + col0, rows[r]);
#pragma unroll
    for (int o = 0; o < RT; ++o) {
        load_row3(sA + (row0 + o + 2) * pitch + col0, rows[(o + 2) % 3]);
        f2 win[3][3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) win[a][b] = rows[(o + a) % 3][b];
        fn(o, win);
    }
}

struct BranchW {
    float w5[25], w3[9], w31[3], w13[3];
};
__device__ __forceinline__ void load_branch_weights(const lmnet_dw_params& p, int e, BranchW& w) {
#pragma unroll
    for (int t = 0; t < 25; ++t) w.w5[t] = __ldg(p.w[0] + e * 25 + t);
#pragma unroll
    for (int t = 0; t < 9; ++t) w.w3[t] = __ldg(p.w[1] + e * 9 + t);
#pragma unroll
    for (int t = 0; t < 3; ++t) w.w31[t] = __ldg(p.w[2] + e * 3 + t);
#pragma unroll
    for (int t = 0; t < 3; ++t) w.w13[t] = __ldg(p.w[3] + e * 3 + t);
}

// the four branch outputs of one output pixel pair from its 5x5 window
__device__ __forceinline__ void branches_from_window(const f2 (&win)[5][5], const BranchW& w, f2 (&y)[4]) {
    y[0] = y[1] = y[2] = y[3] = mk2(0.f, 0.f);
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
        for (int b = 0; b < 5; ++b) y[0] = ffma2(win[a][b], w.w5[a * 5 + b], y[0]);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) y[1] = ffma2(win[a + 1][b + 1], w.w3[a * 3 + b], y[1]);
#pragma unroll
    for (int a = 0; a < 3; ++a) y[2] = ffma2(win[a + 1][2], w.w31[a], y[2]);
#pragma unroll
    for (int b = 0; b < 3; ++b) y[3] = ffma2(win[2][b + 1], w.w13[b], y[3]);
}

// Sum N per-thread floats over the CTA; thread k < N ends up with the total in the return slot.
template <int N>
__device__ __forceinline__ void block_sum(float (&v)[N], float* s_red /* [kDwWarps][N] */, float* out /* [N] or null */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) s_red[warp * N + k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < N && out != nullptr) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < kDwWarps; ++w) a += s_red[w * N + threadIdx.x];
        out[threadIdx.x] = a;
    }
}

// 16-bit storage: fast erf (|err| <= 1.5e-7, far below the storage rounding); fp32 storage: libdevice erff
__device__ __forceinline__ float gelu_f(float u) { return gelu_fast(u); }
__device__ __forceinline__ float gelu_grad_f(float u) { return gelu_grad_fast(u); }

template <typename T>
__device__ __forceinline__ f2 load_pair(const T* __restrict__ p, bool v0, bool v1, bool vec) {
    if (vec && v1) {
        if constexpr (sizeof(T) == 4) {
            float2 r = *reinterpret_cast<const float2*>(p);
            return f2{r.x, r.y};
        } else {
            uint32_t raw = *reinterpret_cast<const uint32_t*>(p);
            const T* e = reinterpret_cast<const T*>(&raw);
            return f2{to_f(e[0]), to_f(e[1])};
        }
    }
    return f2{v0 ? to_f(p[0]) : 0.f, v1 ? to_f(p[1]) : 0.f};
}
// packed (unconverted) pair: keeps prefetched operands at one register per bf16/fp16 pair
template <typename T> struct RawPair { using type = uint32_t; };
template <> struct RawPair<float> { using type = float2; };
template <typename T>
__device__ __forceinline__ typename RawPair<T>::type load_pair_raw(const T* __restrict__ p, bool v0, bool v1, bool vec) {
    typename RawPair<T>::type raw{};
    if (vec && v1) return *reinterpret_cast<const typename RawPair<T>::type*>(p);
    T* e = reinterpret_cast<T*>(&raw);
    if (v0) e[0] = p[0];
    if (v1) e[1] = p[1];
    return raw;
}
template <typename T>
__device__ __forceinline__ f2 cvt_pair(const typename RawPair<T>::type& raw) {
    const T* e = reinterpret_cast<const T*>(&raw);
    return f2{to_f(e[0]), to_f(e[1])};
}

template <typename T>
__device__ __forceinline__ void store_pair(T* __restrict__ p, f2 v, bool v0, bool v1, bool vec) {
    if (vec && v1) {
        if constexpr (sizeof(T) == 4) {
            *reinterpret_cast<float2*>(p) = make_float2(v.x, v.y);
        } else {
            uint32_t raw;
            T* e = reinterpret_cast<T*>(&raw);
            e[0] = from_f<T>(v.x);
            e[1] = from_f<T>(v.y);
            *reinterpret_cast<uint32_t*>(p) = raw;
        }
        return;
    }
    if (v0) p[0] = from_f<T>(v.x);
    if (v1) p[1] = from_f<T>(v.y);
}

// =================================================================================================
// forward: statistics pass
// =================================================================================================
constexpr int kFwdRT = 8;                       // output rows per thread
constexpr int kFwdTH = kFwdRT * kDwWarps;       // 32 output rows per tile
constexpr int kFwdTW = 64;                      // output columns per tile
constexpr int kFwdTileRows = kFwdTH + 4;

template <typename T, int VEC>
__global__ void __launch_bounds__(kDwThreads)
dw_stats_kernel(const T* __restrict__ x, lmnet_dw_params p, float* __restrict__ part /* [E][ncta][8] */, DwGeom g) {
    __shared__ __align__(16) float sA[kFwdTileRows * kPitch];
    __shared__ float s_red[kDwWarps * 8];
    const int e = blockIdx.z, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * kFwdTW;
    const int band0 = blockIdx.y * g.rows_per_band, band1 = min(band0 + g.rows_per_band, g.H);
    BranchW w;
    load_branch_weights(p, e, w);
    f2 s[4], ss[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) s[k] = ss[k] = mk2(0.f, 0.f);
    const int col = c0 + 2 * lane;
    const f2 cmask = mk2(col < g.W ? 1.f : 0.f, col + 1 < g.W ? 1.f : 0.f);
    for_each_tile<T, VEC, kFwdTileRows>(
        g, band0, band1, kFwdTH, 2, c0, sA,
        [&](int b) { return x + ((int64_t)b * g.E + e) * g.H * g.W; }, NoExtra{},
        [&](int, int tr, bool) {
            walk5<kFwdRT>(sA, warp * kFwdRT, 2 * lane + kPad - 2, [&](int o, const f2(&win)[5][5]) {
                const int row = tr + warp * kFwdRT + o;
                f2 y[4];
                branches_from_window(win, w, y);
                const float rm = row < band1 ? 1.f : 0.f;
                const f2 m = mk2(cmask.x * rm, cmask.y * rm);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    s[k] = ffma2(y[k], m, s[k]);
                    ss[k] = ffma2(mk2(y[k].x * m.x, y[k].y * m.y), y[k], ss[k]);
                }
            });
        });
    float v[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = s[k].x + s[k].y; v[4 + k] = ss[k].x + ss[k].y; }
    const int ncta = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
    block_sum<8>(v, s_red, part + ((int64_t)e * ncta + cta) * 8);
}

// per-channel finalize of the forward statistics: mean / rstd, running-stat update, merged 5x5 kernel
// coef[e][0..24] = merged taps, coef[e][25] = bias
// Sum of n floats, `stride` apart, by ONE warp: lane l takes c = l, l + 32, ... (independent loads), then a fixed-order
// shuffle tree in double => deterministic.  A single thread walking the per-CTA partials is a chain of dependent loads:
// 36 partials x ~0.4 us made the finalize kernels 10-15 us each (48 launches per step, tools/bench_dw.py).
__device__ __forceinline__ double warp_sum_strided(const float* __restrict__ p, int n, int64_t stride, int lane) {
    double a = 0;
    for (int c = lane; c < n; c += 32) a += (double)p[(int64_t)c * stride];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    return a;
}

constexpr int kFinThreads = 256;

__global__ void __launch_bounds__(kFinThreads)
dw_fin_fwd_kernel(const float* __restrict__ part, int ncta, int part_stride, lmnet_dw_params p,
                  float* __restrict__ save_mean, float* __restrict__ save_rstd, float* __restrict__ coef,
                  float* __restrict__ gram /* [E][40] or null */, float eps, float momentum,
                  int64_t* nbt0, int64_t* nbt1, int64_t* nbt2, int64_t* nbt3, DwGeom g) {
    // one CTA per channel; every warp reduces some of the 8 (+ 40 with `gram`: the lag sums of the statistics + Gram
    // pass, reparam_dw_tma2.cuh) per-CTA partial sums, thread 0 finishes
    const int e = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ double s_tot[8];
    const int nq = gram != nullptr ? 48 : 8;
    for (int q = warp; q < nq; q += kFinThreads / 32) {
        const double a = warp_sum_strided(part + (int64_t)e * ncta * part_stride + q, ncta, part_stride, lane);
        if (lane == 0) {
            if (q < 8) s_tot[q] = a;
            else gram[e * 40 + (q - 8)] = (float)a;
        }
    }
    __syncthreads();
    // four threads finish one branch each; 25 threads merge one tap each (a single thread walking 4 branches x 40
    // dependent parameter loads made this launch 10-15 us)
    __shared__ float s_a[4], s_b[4];
    if (threadIdx.x < 4) {
        const int k = threadIdx.x;
        const double n = (double)g.B * g.H * g.W;
        const double mean = s_tot[k] / n;
        double var = s_tot[4 + k] / n - mean * mean;
        if (var < 0) var = 0;
        const float rstd = (float)(1.0 / sqrt(var + (double)eps));
        save_mean[k * g.E + e] = (float)mean;
        save_rstd[k * g.E + e] = rstd;
        if (p.running_mean[k] != nullptr) {
            const double unbiased = n > 1 ? var * n / (n - 1) : var;
            p.running_mean[k][e] = (1.f - momentum) * p.running_mean[k][e] + momentum * (float)mean;
            p.running_var[k][e] = (1.f - momentum) * p.running_var[k][e] + momentum * (float)unbiased;
        }
        const float a = p.gamma[k][e] * rstd;
        s_a[k] = a;
        s_b[k] = p.beta[k][e] - a * (float)mean;
    }
    __syncthreads();
    if (threadIdx.x < 25) {
        const int t = threadIdx.x, i = t / 5, j = t % 5;
        float m = s_a[0] * p.w[0][e * 25 + t];                 // same accumulation order as the single-thread version
        if (i >= 1 && i <= 3 && j >= 1 && j <= 3) m += s_a[1] * p.w[1][e * 9 + (i - 1) * 3 + (j - 1)];
        if (i >= 1 && i <= 3 && j == 2) m += s_a[2] * p.w[2][e * 3 + (i - 1)];
        if (i == 2 && j >= 1 && j <= 3) m += s_a[3] * p.w[3][e * 3 + (j - 1)];
        coef[e * 26 + t] = m;
    }
    if (threadIdx.x == 32) {
        coef[e * 26 + 25] = ((s_b[0] + s_b[1]) + s_b[2]) + s_b[3];
        if (e == 0) {
            if (nbt0) *nbt0 += 1;
            if (nbt1) *nbt1 += 1;
            if (nbt2) *nbt2 += 1;
            if (nbt3) *nbt3 += 1;
        }
    }
}

// inference coefficients: running statistics folded into one 5x5 kernel + bias (or deploy weights)
__global__ void dw_coef_eval_kernel(lmnet_dw_params p, const float* __restrict__ deploy_bias, float eps,
                                    float* __restrict__ coef, int E) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    if (p.gamma[0] == nullptr) {  // deploy mode: p.w[0] is the fused kernel
        for (int t = 0; t < 25; ++t) coef[e * 26 + t] = p.w[0][e * 25 + t];
        coef[e * 26 + 25] = deploy_bias != nullptr ? deploy_bias[e] : 0.f;
        return;
    }
    float m5[25];
    for (int t = 0; t < 25; ++t) m5[t] = 0.f;
    float bias = 0.f;
    for (int k = 0; k < 4; ++k) {
        const float a = p.gamma[k][e] / sqrtf(p.running_var[k][e] + eps);
        bias += p.beta[k][e] - a * p.running_mean[k][e];
        if (k == 0) for (int t = 0; t < 25; ++t) m5[t] += a * p.w[0][e * 25 + t];
        if (k == 1) for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m5[(i + 1) * 5 + j + 1] += a * p.w[1][e * 9 + i * 3 + j];
        if (k == 2) for (int i = 0; i < 3; ++i) m5[(i + 1) * 5 + 2] += a * p.w[2][e * 3 + i];
        if (k == 3) for (int j = 0; j < 3; ++j) m5[2 * 5 + j + 1] += a * p.w[3][e * 3 + j];
    }
    for (int t = 0; t < 25; ++t) coef[e * 26 + t] = m5[t];
    coef[e * 26 + 25] = bias;
}

// =================================================================================================
// forward: apply pass  u = merged5x5(x) + bias;  z = GELU(u);  pool partial sums
// =================================================================================================
template <typename T, int VEC>
__global__ void __launch_bounds__(kDwThreads)
dw_apply_kernel(const T* __restrict__ x, const float* __restrict__ coef, T* __restrict__ u_out, T* __restrict__ z_out,
                float* __restrict__ pool_part /* [B*E][ncta] or null */, DwGeom g) {
    __shared__ __align__(16) float sA[kFwdTileRows * kPitch];
    __shared__ float s_red[kDwWarps];
    const int e = blockIdx.z, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * kFwdTW;
    const int band0 = blockIdx.y * g.rows_per_band, band1 = min(band0 + g.rows_per_band, g.H);
    float wm[25];
#pragma unroll
    for (int t = 0; t < 25; ++t) wm[t] = __ldg(coef + e * 26 + t);
    const float bias = __ldg(coef + e * 26 + 25);
    const int col = c0 + 2 * lane;
    const bool v0 = col < g.W, v1 = col + 1 < g.W;
    const bool vec = (g.W & 1) == 0;
    const int ncta = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
    float psum = 0.f;
    for_each_tile<T, VEC, kFwdTileRows>(
        g, band0, band1, kFwdTH, 2, c0, sA,
        [&](int b) { return x + ((int64_t)b * g.E + e) * g.H * g.W; }, NoExtra{},
        [&](int b, int tr, bool last) {
            const int64_t poff = ((int64_t)b * g.E + e) * g.H * g.W;
            walk5<kFwdRT>(sA, warp * kFwdRT, 2 * lane + kPad - 2, [&](int o, const f2(&win)[5][5]) {
                const int row = tr + warp * kFwdRT + o;
                f2 acc = mk2(bias, bias);
#pragma unroll
                for (int a = 0; a < 5; ++a)
#pragma unroll
                    for (int bb = 0; bb < 5; ++bb) acc = ffma2(win[a][bb], wm[a * 5 + bb], acc);
                if (row < band1 && v0) {
                    const int64_t off = poff + (int64_t)row * g.W + col;
                    if (u_out != nullptr) store_pair(u_out + off, acc, v0, v1, vec);
                    // GELU of the value as stored (what the reference's next op would read)
                    f2 ur = mk2(to_f(from_f<T>(acc.x)), to_f(from_f<T>(acc.y)));
                    f2 z = mk2(gelu_t<T>(ur.x), gelu_t<T>(ur.y));
                    store_pair(z_out + off, z, v0, v1, vec);
                    psum += to_f(from_f<T>(z.x)) + (v1 ? to_f(from_f<T>(z.y)) : 0.f);
                }
            });
            if (last && pool_part != nullptr) {   // uniform across the CTA
                float t = warp_sum(psum);
                psum = 0.f;
                __syncthreads();
                if (lane == 0) s_red[warp] = t;
                __syncthreads();
                if (threadIdx.x == 0)
                    pool_part[((int64_t)b * g.E + e) * ncta + cta] = s_red[0] + s_red[1] + s_red[2] + s_red[3];
            }
        });
}

__global__ void dw_pool_fin_kernel(const float* __restrict__ pool_part, int ncta, float inv_hw, float* __restrict__ pool, int n) {
    // one warp per (image, channel)
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n) return;
    const double a = warp_sum_strided(pool_part + (int64_t)i * ncta, ncta, 1, lane);
    if (lane == 0) pool[i] = (float)a * inv_hw;
}

// =================================================================================================
// backward R pass: du = (dz + dpool/HW) * gelu'(u);  P[t] = sum_p du(p) x(p+t) (25 lags);  sum du
// =================================================================================================
template <typename T, int VEC>
__global__ void __launch_bounds__(kDwThreads)
dw_bwd_reduce_kernel(const T* __restrict__ x, const T* __restrict__ u, const T* __restrict__ dz,
                     const float* __restrict__ dpool, T* __restrict__ du_out, float* __restrict__ part /* [E][ncta][26] */,
                     DwGeom g) {
    __shared__ __align__(16) float sA[kFwdTileRows * kPitch];
    __shared__ __align__(16) T s_u[kFwdTH * kRawPitch];
    __shared__ __align__(16) T s_dz[kFwdTH * kRawPitch];
    __shared__ float s_red[kDwWarps * 26];
    const int e = blockIdx.z, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * kFwdTW;
    const int band0 = blockIdx.y * g.rows_per_band, band1 = min(band0 + g.rows_per_band, g.H);
    const int col = c0 + 2 * lane;
    const bool v0 = col < g.W, v1 = col + 1 < g.W;
    const bool vec = (g.W & 1) == 0;
    const float inv_hw = 1.f / ((float)g.H * (float)g.W);
    f2 P[25];
#pragma unroll
    for (int t = 0; t < 25; ++t) P[t] = mk2(0.f, 0.f);
    f2 sdu = mk2(0.f, 0.f);
    // u and dz tiles (storage type, no halo) ride the same register-prefetch pipeline as x
    struct UDz {
        RawTileLoader<T, VEC, kFwdTH> lu, lz;
        const T *u, *dz;
        T *su, *sz;
        int e, c0;
        const DwGeom& g;
        __device__ __forceinline__ void fetch(int b, int tr) {
            const int64_t poff = ((int64_t)b * g.E + e) * g.H * g.W;
            lu.fetch(u + poff, g.H, g.W, tr, c0);
            lz.fetch(dz + poff, g.H, g.W, tr, c0);
        }
        __device__ __forceinline__ void commit() { lu.commit(su); lz.commit(sz); }
    } udz{{}, {}, u, dz, s_u, s_dz, e, c0, g};
    for_each_tile<T, VEC, kFwdTileRows>(
        g, band0, band1, kFwdTH, 2, c0, sA,
        [&](int b) { return x + ((int64_t)b * g.E + e) * g.H * g.W; }, udz,
        [&](int b, int tr, bool) {
            const int64_t poff = ((int64_t)b * g.E + e) * g.H * g.W;
            const float dp = dpool != nullptr ? __ldg(dpool + b * g.E + e) * inv_hw : 0.f;
            walk5<kFwdRT>(sA, warp * kFwdRT, 2 * lane + kPad - 2, [&](int o, const f2(&win)[5][5]) {
                const int trow = warp * kFwdRT + o, row = tr + trow;
                f2 du = mk2(0.f, 0.f);
                if (row < band1 && v0) {
                    const f2 uf = cvt_pair<T>(*reinterpret_cast<const typename RawPair<T>::type*>(s_u + trow * kRawPitch + 2 * lane));
                    const f2 gf = cvt_pair<T>(*reinterpret_cast<const typename RawPair<T>::type*>(s_dz + trow * kRawPitch + 2 * lane));
                    du = mk2((gf.x + dp) * gelu_grad_t<T>(uf.x), v1 ? (gf.y + dp) * gelu_grad_t<T>(uf.y) : 0.f);
                    store_pair(du_out + poff + (int64_t)row * g.W + col, du, v0, v1, vec);
                    // keep exactly what later passes will read back
                    du = mk2(to_f(from_f<T>(du.x)), to_f(from_f<T>(du.y)));
                }
                sdu.x += du.x; sdu.y += du.y;
#pragma unroll
                for (int a = 0; a < 5; ++a)
#pragma unroll
                    for (int bb = 0; bb < 5; ++bb) P[a * 5 + bb] = ffma2(win[a][bb], du, P[a * 5 + bb]);
            });
        });
    float v[26];
#pragma unroll
    for (int t = 0; t < 25; ++t) v[t] = P[t].x + P[t].y;
    v[25] = sdu.x + sdu.y;
    const int ncta = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
    block_sum<26>(v, s_red, part + ((int64_t)e * ncta + cta) * 26);
}

// Per-channel finalize of the backward reductions.
//   Pfin[e][26]   : P[t] summed over CTAs, [25] = sum du
//   dgamma_br = rstd_br * (sum_t w_br[t] P[t] - mean_br * sum du),  dbeta_br = sum du
//   cb[e][br*3 + {0,1,2}] = c1, c2, c0 with dy_br = c1*du - c2*y_br - c0
__global__ void dw_fin_bwd_kernel(const float* __restrict__ part, int ncta, lmnet_dw_params p,
                                  const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
                                  lmnet_dw_grads gr, float* __restrict__ Pfin, float* __restrict__ cb, DwGeom g) {
    // one warp per channel: lanes 0..25 each reduce one partial sum over the CTAs, lane 0 finishes
    const int e = blockIdx.x;
    __shared__ double s_P[26];
    if (threadIdx.x < 26) {
        double a = 0;
        for (int c = 0; c < ncta; ++c) a += part[((int64_t)e * ncta + c) * 26 + threadIdx.x];
        s_P[threadIdx.x] = a;
        Pfin[e * 26 + threadIdx.x] = (float)a;
    }
    __syncwarp();
    if (threadIdx.x != 0) return;
    const double n = (double)g.B * g.H * g.W;
    double P[26];
    for (int t = 0; t < 26; ++t) P[t] = s_P[t];
    const double sdu = P[25];
    for (int k = 0; k < 4; ++k) {
        double sduy = 0;
        if (k == 0) for (int t = 0; t < 25; ++t) sduy += (double)p.w[0][e * 25 + t] * P[t];
        if (k == 1) for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) sduy += (double)p.w[1][e * 9 + i * 3 + j] * P[(i + 1) * 5 + j + 1];
        if (k == 2) for (int i = 0; i < 3; ++i) sduy += (double)p.w[2][e * 3 + i] * P[(i + 1) * 5 + 2];
        if (k == 3) for (int j = 0; j < 3; ++j) sduy += (double)p.w[3][e * 3 + j] * P[2 * 5 + j + 1];
        const double mean = save_mean[k * g.E + e], rstd = save_rstd[k * g.E + e], gamma = p.gamma[k][e];
        const double dgamma = rstd * (sduy - mean * sdu);
        if (gr.dgamma[k] != nullptr) gr.dgamma[k][e] = (float)dgamma;
        if (gr.dbeta[k] != nullptr) gr.dbeta[k][e] = (float)sdu;
        const double c1 = gamma * rstd;
        const double c2 = gamma * rstd * rstd * dgamma / n;
        const double c0 = c1 * sdu / n - c2 * mean;
        cb[e * 12 + k * 3 + 0] = (float)c1;
        cb[e * 12 + k * 3 + 1] = (float)c2;
        cb[e * 12 + k * 3 + 2] = (float)c0;
    }
}

// =================================================================================================
// backward A1 pass: dx = sum_br w_br (*)^T dy_br,  dy_br = c1*du - c2*y_br - c0 inside the image
// =================================================================================================
constexpr int kA1RT = 4;                      // dx rows per thread
constexpr int kA1TH = kA1RT * kDwWarps;       // 16 dx rows per tile
constexpr int kA1TW = 56;                     // dx columns per tile (multiple of 8; dy region = 60 columns)
constexpr int kA1RegPairs = (kA1TW + 4) / 2;  // 30 column pairs of the dy region
constexpr int kA1RegRows = kA1TH + 4;         // dy region rows (20)
constexpr int kA1RegRT = kA1RegRows / kDwWarps;  // 5 region rows per warp
constexpr int kA1XRows = kA1TH + 8;           // x tile rows (halo 4)
constexpr int kDyPitch = 72;                  // floats per row of the dy tiles (region col k at index k)

template <typename T, int VEC>
__global__ void __launch_bounds__(kDwThreads)
dw_bwd_dx_kernel(const T* __restrict__ x, const T* __restrict__ du, lmnet_dw_params p, const float* __restrict__ cb,
                 T* __restrict__ dx, DwGeom g) {
    extern __shared__ __align__(16) float smem[];
    float* xA = smem;                                   // [kA1XRows][kPitch]     x, halo 4
    float* dyA = xA + kA1XRows * kPitch;                // [4][kA1RegRows][kDyPitch]
    T* s_du = reinterpret_cast<T*>(dyA + 4 * kA1RegRows * kDyPitch);   // [kA1RegRows][kRawPitch]
    const int e = blockIdx.z, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * kA1TW;
    const int band0 = blockIdx.y * g.rows_per_band, band1 = min(band0 + g.rows_per_band, g.H);
    BranchW w;
    load_branch_weights(p, e, w);
    float c1[4], c2[4], c0c[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        c1[k] = __ldg(cb + e * 12 + k * 3);
        c2[k] = __ldg(cb + e * 12 + k * 3 + 1);
        c0c[k] = __ldg(cb + e * 12 + k * 3 + 2);
    }
    const bool vec = (g.W & 1) == 0;
    // dy tiles: make the never-written tail columns finite (they are read by idle lanes only)
    for (int i = threadIdx.x; i < 4 * kA1RegRows * kDyPitch; i += kDwThreads) dyA[i] = 0.f;
    // du region tile: rows tr-2 .. tr+TH+2, staged from column c0-4 (vector aligned); region col k -> index k+2
    struct DuStage {
        RawTileLoader<T, VEC, kA1RegRows> ld;
        const T* du;
        T* s;
        int e, c0;
        const DwGeom& g;
        __device__ __forceinline__ void fetch(int b, int tr) {
            ld.fetch(du + ((int64_t)b * g.E + e) * g.H * g.W, g.H, g.W, tr - 2, c0 - 4);
        }
        __device__ __forceinline__ void commit() { ld.commit(s); }
    } dus{{}, du, s_du, e, c0, g};
    for_each_tile<T, VEC, kA1XRows>(
        g, band0, band1, kA1TH, 4, c0, xA,
        [&](int b) { return x + ((int64_t)b * g.E + e) * g.H * g.W; }, dus,
        [&](int b, int tr, bool) {
            const int64_t poff = ((int64_t)b * g.E + e) * g.H * g.W;
            const int rcol = c0 - 2 + 2 * lane;           // global column of the region pair
            // phase 2: y_br and dy_br on the (TH+4) x 60 region whose origin is (tr-2, c0-2);
            // the x tile origin is (tr-4, c0-kPad), so region column k sits at tile index k + kPad - 4
            if (lane < kA1RegPairs) {
                walk5<kA1RegRT>(xA, warp * kA1RegRT, 2 * lane + kPad - 4, [&](int o, const f2(&win)[5][5]) {
                    const int rr = warp * kA1RegRT + o;       // region row
                    const int row = tr - 2 + rr;
                    f2 y[4];
                    branches_from_window(win, w, y);
                    const bool rin = row >= 0 && row < g.H;
                    const bool i0 = rin && rcol >= 0 && rcol < g.W, i1 = rin && rcol + 1 >= 0 && rcol + 1 < g.W;
                    const f2 d = cvt_pair<T>(*reinterpret_cast<const typename RawPair<T>::type*>(s_du + rr * kRawPitch + 2 * lane + 2));
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        f2 dy = mk2(i0 ? c1[k] * d.x - c2[k] * y[k].x - c0c[k] : 0.f, i1 ? c1[k] * d.y - c2[k] * y[k].y - c0c[k] : 0.f);
                        *reinterpret_cast<float2*>(dyA + k * kA1RegRows * kDyPitch + rr * kDyPitch + 2 * lane) = make_float2(dy.x, dy.y);
                    }
                });
            }
            __syncthreads();
            // phase 3: dx(r,c) = sum_br sum_{a,b} w_br[a][b] * dy_br(r-(a-2), c-(b-2))  (flipped taps)
            if (2 * lane < kA1TW) {
                f2 acc[kA1RT];
#pragma unroll
                for (int o = 0; o < kA1RT; ++o) acc[o] = mk2(0.f, 0.f);
                walk5<kA1RT>(dyA, warp * kA1RT, 2 * lane, [&](int o, const f2(&win)[5][5]) {
#pragma unroll
                    for (int a = 0; a < 5; ++a)
#pragma unroll
                        for (int bb = 0; bb < 5; ++bb) acc[o] = ffma2(win[a][bb], w.w5[(4 - a) * 5 + (4 - bb)], acc[o]);
                }, kDyPitch);
                walk3<kA1RT>(dyA + kA1RegRows * kDyPitch, warp * kA1RT + 1, 2 * lane,
                             [&](int o, const f2(&win)[3][3]) {
#pragma unroll
                                 for (int a = 0; a < 3; ++a)
#pragma unroll
                                     for (int bb = 0; bb < 3; ++bb) acc[o] = ffma2(win[a][bb], w.w3[(2 - a) * 3 + (2 - bb)], acc[o]);
                             }, kDyPitch);
                walk3<kA1RT>(dyA + 2 * kA1RegRows * kDyPitch, warp * kA1RT + 1, 2 * lane,
                             [&](int o, const f2(&win)[3][3]) {
#pragma unroll
                                 for (int a = 0; a < 3; ++a) acc[o] = ffma2(win[a][1], w.w31[2 - a], acc[o]);
                             }, kDyPitch);
                walk3<kA1RT>(dyA + 3 * kA1RegRows * kDyPitch, warp * kA1RT + 1, 2 * lane,
                             [&](int o, const f2(&win)[3][3]) {
#pragma unroll
                                 for (int bb = 0; bb < 3; ++bb) acc[o] = ffma2(win[1][bb], w.w13[2 - bb], acc[o]);
                             }, kDyPitch);
                const int col = c0 + 2 * lane;
                if (col < g.W) {
#pragma unroll
                    for (int o = 0; o < kA1RT; ++o) {
                        const int row = tr + warp * kA1RT + o;
                        if (row < band1) store_pair(dx + poff + (int64_t)row * g.W + col, acc[o], true, col + 1 < g.W, vec);
                    }
                }
            }
        });
}

// =================================================================================================
// backward A2 pass: Rbr[t] = sum_p (c2_br*y_br(p) + c0_br) * x(p+t);  dw_br[t] = c1_br*P[t] - Rbr[t]
// =================================================================================================
template <typename T, int VEC>
__global__ void __launch_bounds__(kDwThreads)
dw_bwd_dw_kernel(const T* __restrict__ x, lmnet_dw_params p, const float* __restrict__ cb,
                 float* __restrict__ part /* [E][ncta][40] */, DwGeom g) {
    __shared__ __align__(16) float sA[kFwdTileRows * kPitch];
    __shared__ float s_red[kDwWarps * 40];
    const int e = blockIdx.z, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * kFwdTW;
    const int band0 = blockIdx.y * g.rows_per_band, band1 = min(band0 + g.rows_per_band, g.H);
    BranchW w;
    load_branch_weights(p, e, w);
    float c2[4], c0c[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        c2[k] = __ldg(cb + e * 12 + k * 3 + 1);
        c0c[k] = __ldg(cb + e * 12 + k * 3 + 2);
    }
    const int col = c0 + 2 * lane;
    const f2 cmask = mk2(col < g.W ? 1.f : 0.f, col + 1 < g.W ? 1.f : 0.f);
    f2 A5[25], A3[9], A31[3], A13[3];
#pragma unroll
    for (int t = 0; t < 25; ++t) A5[t] = mk2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 9; ++t) A3[t] = mk2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 3; ++t) A31[t] = A13[t] = mk2(0.f, 0.f);
    for_each_tile<T, VEC, kFwdTileRows>(
        g, band0, band1, kFwdTH, 2, c0, sA,
        [&](int b) { return x + ((int64_t)b * g.E + e) * g.H * g.W; }, NoExtra{},
        [&](int, int tr, bool) {
            walk5<kFwdRT>(sA, warp * kFwdRT, 2 * lane + kPad - 2, [&](int o, const f2(&win)[5][5]) {
                const int row = tr + warp * kFwdRT + o;
                const float rm = row < band1 ? 1.f : 0.f;
                const f2 m = mk2(cmask.x * rm, cmask.y * rm);
                f2 y[4];
                branches_from_window(win, w, y);
                f2 gk[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) gk[k] = mk2((c2[k] * y[k].x + c0c[k]) * m.x, (c2[k] * y[k].y + c0c[k]) * m.y);
#pragma unroll
                for (int a = 0; a < 5; ++a)
#pragma unroll
                    for (int bb = 0; bb < 5; ++bb) A5[a * 5 + bb] = ffma2(win[a][bb], gk[0], A5[a * 5 + bb]);
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int bb = 0; bb < 3; ++bb) A3[a * 3 + bb] = ffma2(win[a + 1][bb + 1], gk[1], A3[a * 3 + bb]);
#pragma unroll
                for (int a = 0; a < 3; ++a) A31[a] = ffma2(win[a + 1][2], gk[2], A31[a]);
#pragma unroll
                for (int bb = 0; bb < 3; ++bb) A13[bb] = ffma2(win[2][bb + 1], gk[3], A13[bb]);
            });
        });
    float v[40];
#pragma unroll
    for (int t = 0; t < 25; ++t) v[t] = A5[t].x + A5[t].y;
#pragma unroll
    for (int t = 0; t < 9; ++t) v[25 + t] = A3[t].x + A3[t].y;
#pragma unroll
    for (int t = 0; t < 3; ++t) { v[34 + t] = A31[t].x + A31[t].y; v[37 + t] = A13[t].x + A13[t].y; }
    const int ncta = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
    block_sum<40>(v, s_red, part + ((int64_t)e * ncta + cta) * 40);
}

__global__ void dw_fin_dw_kernel(const float* __restrict__ part, int ncta, const float* __restrict__ Pfin,
                                 const float* __restrict__ cb, lmnet_dw_grads gr, int E) {
    const int e = blockIdx.x;
    const int t = threadIdx.x;  // 0..39
    if (e >= E || t >= 40) return;
    double r = 0;
    for (int c = 0; c < ncta; ++c) r += part[((int64_t)e * ncta + c) * 40 + t];
    int k, lag, idx;
    if (t < 25) { k = 0; idx = t; lag = t; }
    else if (t < 34) { k = 1; idx = t - 25; lag = (idx / 3 + 1) * 5 + idx % 3 + 1; }
    else if (t < 37) { k = 2; idx = t - 34; lag = (idx + 1) * 5 + 2; }
    else { k = 3; idx = t - 37; lag = 2 * 5 + idx + 1; }
    // Pfin == nullptr: the partials already are dw (the TMA dx kernel correlates dy_br with x directly)
    const float val = Pfin == nullptr ? (float)r : (float)((double)cb[e * 12 + k * 3] * (double)Pfin[e * 26 + lag] - r);
    const int per = k == 0 ? 25 : k == 1 ? 9 : 3;
    if (gr.dw[k] != nullptr) gr.dw[k][e * per + idx] = val;
}

}  // namespace lmnet
#include "reparam_dw_mma.cuh"
#include "reparam_dw_tma.cuh"
#include "reparam_dw_tma2.cuh"
namespace lmnet {

// =================================================================================================
// host side
// =================================================================================================
static size_t dw_align(size_t x) { return (x + 255) / 256 * 256; }

static int dw_validate(const lmnet_dw_dims* d) {
    if (d == nullptr || d->B <= 0 || d->E <= 0 || d->H <= 0 || d->W <= 0) return LMNET_ERR_INVALID_ARG;
    if ((int64_t)d->B * d->E * d->H * d->W >= (int64_t)1 << 40) return LMNET_ERR_INVALID_ARG;
    return LMNET_OK;
}

// CTA grid for a kernel with tile (th x tw).  A CTA owns one channel, one column stripe and one band of row tiles (and
// loops over the batch); `ctas_per_sm` x 148 CTAs are resident at once.  The band count minimises the makespan
// waves(bands) x tiles_per_band(bands): rounding the band count UP to fill the machine once (the first rule) put 576 CTAs
// on 444 slots at level 1 for the 3-CTA kernels — a second, 30 % full wave of whole bands (ncu: 190 us where 4 row
// tiles per CTA in ONE wave need ~130).
static DwGeom dw_geom(const lmnet_dw_dims* d, int th, int tw, int ctas_per_sm = 4, int shift = 0) {
    DwGeom g;
    g.B = d->B; g.E = d->E; g.H = d->H; g.W = d->W;
    g.stripes = (d->W + shift + tw - 1) / tw;          // TMA kernels: stripe s starts at s*tw - shift
    const int row_tiles = (d->H + th - 1) / th;
    static const int kCtaOverride = getenv("LMNET_DW_CTAS") ? atoi(getenv("LMNET_DW_CTAS")) : 0;
    const int64_t capacity = kCtaOverride > 0 ? kCtaOverride : ctas_per_sm * 148;     // resident CTAs
    const int64_t per_band = (int64_t)g.E * g.stripes;
    // Ties (B200, tools/bench_dw.py with LMNET_DW_TPB=1..6): a grid that fills < 90 % of the first wave loses 15 % at
    // level 3 (384 CTAs x 2 tiles against 576 x 1 on 444 slots), beyond that the smaller grid wins or is within 5 %.
    int best_tpb = row_tiles;
    int64_t best_cost = -1, best_ctas = 0;
    bool best_full = false;
    for (int bands = 1; bands <= row_tiles; ++bands) {
        const int tpb = (row_tiles + bands - 1) / bands;
        const int nb = (row_tiles + tpb - 1) / tpb;
        const int64_t ctas = per_band * nb, waves = (ctas + capacity - 1) / capacity;
        const int64_t cost = waves * tpb;
        const bool full = 10 * ctas >= 9 * capacity;
        const bool better = best_cost < 0 || cost < best_cost ||
                            (cost == best_cost && (full != best_full ? full : (full ? ctas < best_ctas : ctas > best_ctas)));
        if (better) { best_cost = cost; best_ctas = ctas; best_tpb = tpb; best_full = full; }
    }
    static const int kTpbOverride = getenv("LMNET_DW_TPB") ? atoi(getenv("LMNET_DW_TPB")) : 0;      // experiments
    if (kTpbOverride > 0) best_tpb = std::min(kTpbOverride, row_tiles);
    g.rows_per_band = best_tpb * th;
    g.bands = (d->H + g.rows_per_band - 1) / g.rows_per_band;
    return g;
}

struct DwWs {
    size_t part, coef, pool_part, pfin, cb, coef2, wrec, du, total;
};
// CTAs per SM the TMA kernels are built for (__launch_bounds__ / shared memory): their grids are ONE wave
constexpr int kStatsOcc = 4, kApplyOcc = 4, kReduceOcc = 3, kDxOcc = LMNET_DX_OCC, kStatsGramOcc = 3, kDx2Occ = 4;
static DwWs dw_ws_layout(const lmnet_dw_dims* d, size_t esize) {
    size_t ncta = 0;
    const DwGeom gs[6] = {dw_geom(d, kFwdTH, kFwdTW), dw_geom(d, kMmaTH, kMmaTW, kStatsOcc, kFwdShift), dw_geom(d, kMmaTH, kMmaTW, kReduceOcc, kFwdShift),
                          dw_geom(d, kDxTH, kDxTW), dw_geom(d, kDxTH, kDxTW, kDxOcc, kDxShift), dw_geom(d, kMmaTH, kMmaTW, kDx2Occ, kDx2Shift)};
    for (const DwGeom& gg : gs) ncta = std::max(ncta, (size_t)gg.stripes * gg.bands);
    DwWs w;
    size_t off = 0;
    w.part = off; off = dw_align(off + (size_t)d->E * ncta * kReducePartS * sizeof(float));
    w.coef = off; off = dw_align(off + (size_t)d->E * 26 * sizeof(float));
    w.pool_part = off; off = dw_align(off + (size_t)d->B * d->E * ncta * kDwWarps * sizeof(float));
    w.pfin = off; off = dw_align(off + (size_t)d->E * 26 * sizeof(float));
    w.cb = off; off = dw_align(off + (size_t)d->E * 12 * sizeof(float));
    w.coef2 = off; off = dw_align(off + (size_t)d->E * kCoef2Stride * sizeof(float));
    w.wrec = off; off = dw_align(off + (size_t)d->E * 27 * sizeof(float4));
    w.du = off; off = dw_align(off + (size_t)d->B * d->E * d->H * d->W * esize);
    w.total = off;
    return w;
}

// widest vector (16 or 8 bytes, else scalar) the tile loader may use for this tensor
static int dw_vec_bytes(const void* const* ptrs, int n, const lmnet_dw_dims* d, size_t es) {
    for (int vb = 16; vb >= 4; vb >>= 1) {
        bool ok = ((size_t)d->W * es) % vb == 0;
        for (int i = 0; i < n; ++i) ok = ok && ((uintptr_t)ptrs[i] % vb == 0);
        if (ok) return vb;
    }
    return 0;
}
// the tensor-core kernels need 16-bit storage, an even row length and 4-byte aligned planes
template <typename T>
static bool dw_mma_ok(const lmnet_dw_dims* d, const void* x) {
    if (sizeof(T) != 2 || getenv("LMNET_DW_NO_MMA") != nullptr) return false;
    return (d->W % 2 == 0) && ((uintptr_t)x % 4 == 0);
}
// the TMA pipeline kernels additionally need 16-byte global strides (W % 8 == 0) and 16-byte aligned tensors
template <typename T>
static bool dw_tma_ok(const lmnet_dw_dims* d, std::initializer_list<const void*> ptrs) {
    if (sizeof(T) != 2 || getenv("LMNET_DW_NO_MMA") != nullptr || getenv("LMNET_DW_NO_TMA") != nullptr) return false;
    for (const void* q : ptrs)
        if (!tma_planes_ok(q, (int64_t)d->B * d->E, d->H, d->W, sizeof(T))) return false;
    return true;
}
template <typename Kern>
static bool dw_tma_smem(Kern kern, size_t bytes, std::atomic<size_t>* granted) { return ensure_smem(kern, bytes, granted); }

template <typename T, typename F>
static int with_vec(int vec_bytes, F&& f) {
    // the loaders move 4, 2 or 1 elements per lane (one float4 / float2 / float shared store each)
    const int elems = vec_bytes / (int)sizeof(T);
    if (elems >= 4) return f(std::integral_constant<int, 4>{});
    if (elems >= 2) return f(std::integral_constant<int, 2>{});
    return f(std::integral_constant<int, 1>{});
}

template <typename T>
static int dw_train_fwd(const void* x, const lmnet_dw_params* p, void* u, void* z, float* pool, float* save_mean,
                        float* save_rstd, float eps, float momentum, int64_t* const* nbt, char* ws,
                        const lmnet_dw_dims* d, cudaStream_t st, float* save_gram = nullptr, int* gram_saved = nullptr) {
    if (gram_saved != nullptr) *gram_saved = 0;
    DwGeom g = dw_geom(d, kFwdTH, kFwdTW);
    DwWs L = dw_ws_layout(d, sizeof(T));
    const double t_bytes = (double)d->B * d->E * d->H * d->W * sizeof(T);  // one [B,E,H,W] tensor
    (void)t_bytes;
    const int ncta = g.stripes * g.bands;
    dim3 grid(g.stripes, g.bands, g.E);
    float* part = (float*)(ws + L.part);
    float* coef = (float*)(ws + L.coef);
    float* pool_part = (float*)(ws + L.pool_part);
    const void* vp[1] = {x};
    const int vb = dw_vec_bytes(vp, 1, d, sizeof(T));
    const bool use_mma = dw_mma_ok<T>(d, x);
    int rc = LMNET_OK;
    if constexpr (sizeof(T) == 2) {
        if (dw_tma_ok<T>(d, {x, z, u ? u : z})) {
            CUtensorMap tm_x;
            if (!tma_make_planes_map(&tm_x, x, (int64_t)d->B * d->E, d->H, d->W, kMmaTileRows, kMmaPitch)) return LMNET_ERR_LAUNCH;
            static std::atomic<size_t> granted_s[kMaxDevices], granted_a[kMaxDevices];
            if (!dw_tma_smem(dw_stats_tma_kernel<T>, kStatsSmem, granted_s) || !dw_tma_smem(dw_apply_tma_kernel<T>, kApplySmem, granted_a))
                return LMNET_ERR_LAUNCH;
            const DwGeom gs = dw_geom(d, kMmaTH, kMmaTW, kStatsOcc, kFwdShift);
            const int ncta_s = gs.stripes * gs.bands;
            const dim3 grid_s(gs.stripes, gs.bands, gs.E);
            const bool kNoGram = getenv("LMNET_DW_NO_GRAM") != nullptr;     // A/B switch (tests)
            if (save_gram != nullptr && gram_saved != nullptr && !kNoGram) {
                // statistics + Gram pass: the lag sums Q_br, S that turn the backward into one composite stencil
                static std::atomic<size_t> granted_g[kMaxDevices];
                if (!dw_tma_smem(dw_stats_gram_tma_kernel<T>, kStatsGramSmem, granted_g)) return LMNET_ERR_LAUNCH;
                const DwGeom gg = dw_geom(d, kMmaTH, kMmaTW, kStatsGramOcc, kFwdShift);
                const dim3 grid_g(gg.stripes, gg.bands, gg.E);
                LMNET_LAUNCH(KID_DW_STATS, st, 1 * t_bytes, (dw_stats_gram_tma_kernel<T><<<grid_g, kTmaThreads, kStatsGramSmem, st>>>(tm_x, *p, part, gg)));
                LMNET_LAUNCH(KID_DW_FIN_FWD, st, 0, (dw_fin_fwd_kernel<<<g.E, kFinThreads, 0, st>>>(part, gg.stripes * gg.bands, kGramPartStride, *p, save_mean, save_rstd, coef, save_gram, eps, momentum,
                                                                 nbt ? nbt[0] : nullptr, nbt ? nbt[1] : nullptr,
                                                                 nbt ? nbt[2] : nullptr, nbt ? nbt[3] : nullptr, g)));
                *gram_saved = 1;
            } else {
                LMNET_LAUNCH(KID_DW_STATS, st, 1 * t_bytes, (dw_stats_tma_kernel<T><<<grid_s, kTmaThreads, kStatsSmem, st>>>(tm_x, *p, part, gs)));
                LMNET_LAUNCH(KID_DW_FIN_FWD, st, 0, (dw_fin_fwd_kernel<<<g.E, kFinThreads, 0, st>>>(part, ncta_s, 8, *p, save_mean, save_rstd, coef, nullptr, eps, momentum,
                                                                 nbt ? nbt[0] : nullptr, nbt ? nbt[1] : nullptr,
                                                                 nbt ? nbt[2] : nullptr, nbt ? nbt[3] : nullptr, g)));
            }
            CUtensorMap tm_u, tm_z;
            if (!tma_make_planes_map(&tm_u, u ? u : z, (int64_t)d->B * d->E, d->H, d->W, kStageRows, kStageCols) ||
                !tma_make_planes_map(&tm_z, z, (int64_t)d->B * d->E, d->H, d->W, kStageRows, kStageCols))
                return LMNET_ERR_LAUNCH;
            LMNET_LAUNCH(KID_DW_APPLY, st, 2 * t_bytes, (dw_apply_tma_kernel<T><<<grid_s, kTmaThreads, kApplySmem, st>>>(tm_x, tm_u, tm_z, coef, (T*)u, (T*)z, pool ? pool_part : nullptr, gs)));
            if (pool) {
                const int n = g.B * g.E;
                LMNET_LAUNCH(KID_DW_POOL_FIN, st, 0, (dw_pool_fin_kernel<<<(n + 3) / 4, 128, 0, st>>>(pool_part, ncta_s * kDwWarps, 1.f / ((float)g.H * g.W), pool, n)));
            }
            return LMNET_OK;
        }
        if (use_mma) LMNET_LAUNCH(KID_DW_STATS, st, 1 * t_bytes, (dw_stats_mma_kernel<T><<<grid, kDwThreads, 0, st>>>((const T*)x, *p, part, g)));
    }
    if (!use_mma) rc = with_vec<T>(vb, [&](auto v) -> int {
        constexpr int VEC = decltype(v)::value;
        LMNET_LAUNCH(KID_DW_STATS, st, 1 * t_bytes, (dw_stats_kernel<T, VEC><<<grid, kDwThreads, 0, st>>>((const T*)x, *p, part, g)));
        return LMNET_OK;
    });
    if (rc != LMNET_OK) return rc;
    LMNET_LAUNCH(KID_DW_FIN_FWD, st, 0, (dw_fin_fwd_kernel<<<g.E, kFinThreads, 0, st>>>(part, ncta, 8, *p, save_mean, save_rstd, coef, nullptr, eps, momentum,
                                                     nbt ? nbt[0] : nullptr, nbt ? nbt[1] : nullptr,
                                                     nbt ? nbt[2] : nullptr, nbt ? nbt[3] : nullptr, g)));
    if constexpr (sizeof(T) == 2) {
        if (use_mma) LMNET_LAUNCH(KID_DW_APPLY, st, 2 * t_bytes, (dw_apply_mma_kernel<T><<<grid, kDwThreads, 0, st>>>((const T*)x, coef, (T*)u, (T*)z, pool ? pool_part : nullptr, g)));
    }
    if (!use_mma) rc = with_vec<T>(vb, [&](auto v) -> int {
        constexpr int VEC = decltype(v)::value;
        LMNET_LAUNCH(KID_DW_APPLY, st, 2 * t_bytes, (dw_apply_kernel<T, VEC><<<grid, kDwThreads, 0, st>>>((const T*)x, coef, (T*)u, (T*)z, pool ? pool_part : nullptr, g)));
        return LMNET_OK;
    });
    if (rc != LMNET_OK) return rc;
    if (pool) {
        const int n = g.B * g.E;
        LMNET_LAUNCH(KID_DW_POOL_FIN, st, 0, (dw_pool_fin_kernel<<<(n + 3) / 4, 128, 0, st>>>(pool_part, ncta, 1.f / ((float)g.H * g.W), pool, n)));
    }
    return LMNET_OK;
}

template <typename T>
static int dw_eval_fwd(const void* x, const lmnet_dw_params* p, const float* bias, float eps, void* z, float* pool,
                       char* ws, const lmnet_dw_dims* d, cudaStream_t st) {
    DwGeom g = dw_geom(d, kFwdTH, kFwdTW);
    DwWs L = dw_ws_layout(d, sizeof(T));
    const double t_bytes = (double)d->B * d->E * d->H * d->W * sizeof(T);  // one [B,E,H,W] tensor
    (void)t_bytes;
    const int ncta = g.stripes * g.bands;
    dim3 grid(g.stripes, g.bands, g.E);
    float* coef = (float*)(ws + L.coef);
    float* pool_part = (float*)(ws + L.pool_part);
    LMNET_LAUNCH(KID_DW_COEF_EVAL, st, 0, (dw_coef_eval_kernel<<<(g.E + 63) / 64, 64, 0, st>>>(*p, bias, eps, coef, g.E)));
    const void* vp[1] = {x};
    const int vb = dw_vec_bytes(vp, 1, d, sizeof(T));
    const bool use_mma = dw_mma_ok<T>(d, x);
    int rc = LMNET_OK;
    if constexpr (sizeof(T) == 2) {
        if (dw_tma_ok<T>(d, {x, z})) {
            CUtensorMap tm_x, tm_z;
            if (!tma_make_planes_map(&tm_x, x, (int64_t)d->B * d->E, d->H, d->W, kMmaTileRows, kMmaPitch) ||
                !tma_make_planes_map(&tm_z, z, (int64_t)d->B * d->E, d->H, d->W, kStageRows, kStageCols))
                return LMNET_ERR_LAUNCH;
            static std::atomic<size_t> granted_a[kMaxDevices];
            if (!dw_tma_smem(dw_apply_tma_kernel<T>, kApplySmem, granted_a)) return LMNET_ERR_LAUNCH;
            const DwGeom gs = dw_geom(d, kMmaTH, kMmaTW, kApplyOcc, kFwdShift);
            const int ncta_s = gs.stripes * gs.bands;
            const dim3 grid_s(gs.stripes, gs.bands, gs.E);
            LMNET_LAUNCH(KID_DW_APPLY, st, 2 * t_bytes, (dw_apply_tma_kernel<T><<<grid_s, kTmaThreads, kApplySmem, st>>>(tm_x, tm_z, tm_z, coef, (T*)nullptr, (T*)z, pool ? pool_part : nullptr, gs)));
            if (pool) {
                const int n = g.B * g.E;
                LMNET_LAUNCH(KID_DW_POOL_FIN, st, 0, (dw_pool_fin_kernel<<<(n + 3) / 4, 128, 0, st>>>(pool_part, ncta_s * kDwWarps, 1.f / ((float)g.H * g.W), pool, n)));
            }
            return LMNET_OK;
        }
        if (use_mma) LMNET_LAUNCH(KID_DW_APPLY, st, 2 * t_bytes, (dw_apply_mma_kernel<T><<<grid, kDwThreads, 0, st>>>((const T*)x, coef, (T*)nullptr, (T*)z, pool ? pool_part : nullptr, g)));
    }
    if (!use_mma) rc = with_vec<T>(vb, [&](auto v) -> int {
        constexpr int VEC = decltype(v)::value;
        LMNET_LAUNCH(KID_DW_APPLY, st, 2 * t_bytes, (dw_apply_kernel<T, VEC><<<grid, kDwThreads, 0, st>>>((const T*)x, coef, (T*)nullptr, (T*)z, pool ? pool_part : nullptr, g)));
        return LMNET_OK;
    });
    if (rc != LMNET_OK) return rc;
    if (pool) {
        const int n = g.B * g.E;
        LMNET_LAUNCH(KID_DW_POOL_FIN, st, 0, (dw_pool_fin_kernel<<<(n + 3) / 4, 128, 0, st>>>(pool_part, ncta, 1.f / ((float)g.H * g.W), pool, n)));
    }
    return LMNET_OK;
}

constexpr size_t kA1SmemBytes = (size_t)(kA1XRows * kPitch + 4 * kA1RegRows * kDyPitch + kA1RegRows * kRawPitch) * sizeof(float);

template <typename T>
static int dw_train_bwd(const void* x, const void* u, const void* dz, const float* dpool, const lmnet_dw_params* p,
                        const float* save_mean, const float* save_rstd, void* dx, const lmnet_dw_grads* gr, char* ws,
                        const lmnet_dw_dims* d, cudaStream_t st, const float* gram = nullptr) {
    DwGeom g = dw_geom(d, kFwdTH, kFwdTW);
    DwWs L = dw_ws_layout(d, sizeof(T));
    const double t_bytes = (double)d->B * d->E * d->H * d->W * sizeof(T);  // one [B,E,H,W] tensor
    (void)t_bytes;
    const int ncta = g.stripes * g.bands;
    dim3 grid(g.stripes, g.bands, g.E);
    float* part = (float*)(ws + L.part);
    float* pfin = (float*)(ws + L.pfin);
    float* cb = (float*)(ws + L.cb);
    T* du = (T*)(ws + L.du);
    const void* vp[4] = {x, u, dz, du};
    const int vb = dw_vec_bytes(vp, 4, d, sizeof(T));
    const bool use_mma = dw_mma_ok<T>(d, x) && ((uintptr_t)u % 4 == 0) && ((uintptr_t)dz % 4 == 0);
    int rc = LMNET_OK;
    if constexpr (sizeof(T) == 2) {
        if (dw_tma_ok<T>(d, {x, u, dz, (const void*)du, (const void*)dx})) {
            const int64_t planes = (int64_t)d->B * d->E;
            CUtensorMap tm_x, tm_u, tm_dz, tm_du;
            if (!tma_make_planes_map(&tm_x, x, planes, d->H, d->W, kMmaTileRows, kMmaPitch) ||
                !tma_make_planes_map(&tm_u, u, planes, d->H, d->W, kMmaTH, kMmaPitch) ||
                !tma_make_planes_map(&tm_dz, dz, planes, d->H, d->W, kMmaTH, kMmaPitch) ||
                !tma_make_planes_map(&tm_du, du, planes, d->H, d->W, kMmaTH, kMmaPitch))
                return LMNET_ERR_LAUNCH;
            static std::atomic<size_t> granted_r[kMaxDevices], granted_r2[kMaxDevices], granted_x[kMaxDevices];
            if (!dw_tma_smem(dw_bwd_reduce_tma_kernel<T, false>, kReduceSmem, granted_r) || !dw_tma_smem(dw_bwd_reduce_tma_kernel<T, true>, kReduceSmem, granted_r2) || !dw_tma_smem(dw_bwd_dx_tma_kernel<T>, kDxTmaSmem, granted_x))
                return LMNET_ERR_LAUNCH;
            const DwGeom gr_ = dw_geom(d, kMmaTH, kMmaTW, kReduceOcc, kFwdShift);
            const dim3 grid_r(gr_.stripes, gr_.bands, gr_.E);
            if (gram == nullptr) LMNET_LAUNCH(KID_DW_BWD_REDUCE, st, 2 * t_bytes, (dw_bwd_reduce_tma_kernel<T, false><<<grid_r, kTmaThreads, kReduceSmem, st>>>(tm_x, tm_u, tm_dz, dpool, du, part, gr_)));
            if (gram != nullptr) {
                LMNET_LAUNCH(KID_DW_BWD_REDUCE, st, 2 * t_bytes, (dw_bwd_reduce_tma_kernel<T, true><<<grid_r, kTmaThreads, kReduceSmem, st>>>(tm_x, tm_u, tm_dz, dpool, du, part, gr_)));
                // composite-stencil backward (reparam_dw_tma2.cuh): finalize (coefficients + all parameter gradients),
                // one tile kernel, frame patch
                CUtensorMap tm_x2, tm_dx;
                if (!tma_make_planes_map(&tm_x2, x, planes, d->H, d->W, kDx2XRows, kMmaPitch) ||
                    !tma_make_planes_map(&tm_du, du, planes, d->H, d->W, kMmaTileRows, kMmaPitch) ||
                    !tma_make_planes_map(&tm_dx, dx, planes, d->H, d->W, kStageRows, kStageCols))
                    return LMNET_ERR_LAUNCH;
                static std::atomic<size_t> granted_x2[kMaxDevices];
                if (!dw_tma_smem(dw_bwd_dx2_tma_kernel<T>, kDx2Smem, granted_x2)) return LMNET_ERR_LAUNCH;
                float* coef2 = (float*)(ws + L.coef2);
                float4* wrec = (float4*)(ws + L.wrec);
                LMNET_LAUNCH(KID_DW_FIN_BWD, st, 0, (dw_fin_bwd2_kernel<<<g.E, kFinThreads, 0, st>>>(part, gr_.stripes * gr_.bands, *p, save_mean, save_rstd, gram, *gr, cb, coef2, wrec, g)));
                const DwGeom gx = dw_geom(d, kMmaTH, kMmaTW, kDx2Occ, kDx2Shift);
                const dim3 grid_x(gx.stripes, gx.bands, gx.E);
                LMNET_LAUNCH(KID_DW_BWD_DX, st, 3 * t_bytes, (dw_bwd_dx2_tma_kernel<T><<<grid_x, kTmaThreads, kDx2Smem, st>>>(tm_x2, tm_du, tm_dx, coef2, (T*)dx, gx)));
                const int nseg = (std::max(d->H, d->W) + kFrameSeg - 1) / kFrameSeg;
                const dim3 grid_f((unsigned)planes, (unsigned)(4 * nseg));
                LMNET_LAUNCH(KID_DW_BWD_FRAME, st, 0, (dw_bwd_frame_kernel<T><<<grid_f, kFrameThreads, 0, st>>>((const T*)x, (T*)dx, wrec, g)));
                return LMNET_OK;
            }
            LMNET_LAUNCH(KID_DW_FIN_BWD, st, 0, (dw_fin_bwd_kernel<<<g.E, 32, 0, st>>>(part, gr_.stripes * gr_.bands, *p, save_mean, save_rstd, *gr, pfin, cb, g)));
            const DwGeom gx = dw_geom(d, kDxTH, kDxTW, kDxOcc, kDxShift);
            const dim3 grid_x(gx.stripes, gx.bands, gx.E);
            LMNET_LAUNCH(KID_DW_BWD_DX, st, 3 * t_bytes, (dw_bwd_dx_tma_kernel<T><<<grid_x, kTmaThreads, kDxTmaSmem, st>>>(tm_x, tm_du, *p, cb, (T*)dx, part, gx)));
            LMNET_LAUNCH(KID_DW_FIN_DW, st, 0, (dw_fin_dw_kernel<<<g.E, 64, 0, st>>>(part, gx.stripes * gx.bands, nullptr, cb, *gr, g.E)));
            return LMNET_OK;
        }
        if (use_mma) LMNET_LAUNCH(KID_DW_BWD_REDUCE, st, 2 * t_bytes, (dw_bwd_reduce_mma_kernel<T><<<grid, kDwThreads, 0, st>>>((const T*)x, (const T*)u, (const T*)dz, dpool, du, part, g)));
    }
    if (!use_mma) rc = with_vec<T>(vb, [&](auto v) -> int {
        constexpr int VEC = decltype(v)::value;
        LMNET_LAUNCH(KID_DW_BWD_REDUCE, st, 2 * t_bytes, (dw_bwd_reduce_kernel<T, VEC><<<grid, kDwThreads, 0, st>>>((const T*)x, (const T*)u, (const T*)dz, dpool, du, part, g)));
        return LMNET_OK;
    });
    if (rc != LMNET_OK) return rc;
    LMNET_LAUNCH(KID_DW_FIN_BWD, st, 0, (dw_fin_bwd_kernel<<<g.E, 32, 0, st>>>(part, ncta, *p, save_mean, save_rstd, *gr, pfin, cb, g)));
    if constexpr (sizeof(T) == 2) {
        if (use_mma) {
            DwGeom gm = dw_geom(d, kDxTH, kDxTW);
            dim3 gm_grid(gm.stripes, gm.bands, gm.E);
            LMNET_LAUNCH(KID_DW_BWD_DX, st, 3 * t_bytes, (dw_bwd_dx_mma_kernel<T><<<gm_grid, kDwThreads, kDxMmaSmemBytes, st>>>((const T*)x, du, *p, cb, (T*)dx, gm)));
        }
    }
    rc = with_vec<T>(vb, [&](auto v) -> int {
        constexpr int VEC = decltype(v)::value;
        if (use_mma) return LMNET_OK;
        DwGeom ga = dw_geom(d, kA1TH, kA1TW);
        static std::atomic<size_t> granted[kMaxDevices];
        if (!ensure_smem(dw_bwd_dx_kernel<T, VEC>, kA1SmemBytes, granted)) return LMNET_ERR_LAUNCH;
        dim3 ga_grid(ga.stripes, ga.bands, ga.E);
        LMNET_LAUNCH(KID_DW_BWD_DX, st, 3 * t_bytes, (dw_bwd_dx_kernel<T, VEC><<<ga_grid, kDwThreads, kA1SmemBytes, st>>>((const T*)x, du, *p, cb, (T*)dx, ga)));
        if (!use_mma) LMNET_LAUNCH(KID_DW_BWD_DW, st, 0, (dw_bwd_dw_kernel<T, VEC><<<grid, kDwThreads, 0, st>>>((const T*)x, *p, cb, part, g)));
        return LMNET_OK;
    });
    if (rc != LMNET_OK) return rc;
    if constexpr (sizeof(T) == 2) {
        if (use_mma) LMNET_LAUNCH(KID_DW_BWD_DW, st, 0, (dw_bwd_dw_mma_kernel<T><<<grid, kDwThreads, 0, st>>>((const T*)x, *p, cb, part, g)));
    }
    LMNET_LAUNCH(KID_DW_FIN_DW, st, 0, (dw_fin_dw_kernel<<<g.E, 64, 0, st>>>(part, ncta, pfin, cb, *gr, g.E)));
    return LMNET_OK;
}

}  // namespace lmnet

using namespace lmnet;

extern "C" size_t lmnet_reparam_dw_workspace_bytes(const lmnet_dw_dims* dims, int dtype) {
    if (dw_validate(dims) != LMNET_OK) return 0;
    return dw_ws_layout(dims, dtype == LMNET_F32 ? 4 : 2).total;
}

static bool dw_params_ok(const lmnet_dw_params* p, bool need_bn) {
    if (p == nullptr || p->w[0] == nullptr) return false;
    if (!need_bn) return true;
    for (int k = 0; k < 4; ++k)
        if (p->w[k] == nullptr || p->gamma[k] == nullptr || p->beta[k] == nullptr) return false;
    return true;
}

extern "C" int lmnet_reparam_dw_train_fwd(const void* x, const lmnet_dw_params* p, void* u, void* z, float* pool,
                                          float* save_mean, float* save_rstd, float eps, float momentum,
                                          int64_t* const* num_batches_tracked,
                                          void* workspace, size_t workspace_bytes,
                                          const lmnet_dw_dims* dims, int dtype, void* stream) {
    int rc = dw_validate(dims);
    if (rc != LMNET_OK) return rc;
    if (!x || !z || !save_mean || !save_rstd || !workspace || !dw_params_ok(p, true)) return LMNET_ERR_INVALID_ARG;
    for (int k = 0; k < 4; ++k)
        if ((p->running_mean[k] == nullptr) != (p->running_var[k] == nullptr)) return LMNET_ERR_INVALID_ARG;
    if (workspace_bytes < lmnet_reparam_dw_workspace_bytes(dims, dtype)) return LMNET_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case LMNET_F32: return dw_train_fwd<float>(x, p, u, z, pool, save_mean, save_rstd, eps, momentum, num_batches_tracked, (char*)workspace, dims, st);
        case LMNET_BF16: return dw_train_fwd<__nv_bfloat16>(x, p, u, z, pool, save_mean, save_rstd, eps, momentum, num_batches_tracked, (char*)workspace, dims, st);
        case LMNET_F16: return dw_train_fwd<__half>(x, p, u, z, pool, save_mean, save_rstd, eps, momentum, num_batches_tracked, (char*)workspace, dims, st);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}

extern "C" int lmnet_reparam_dw_train_fwd_gram(const void* x, const lmnet_dw_params* p, void* u, void* z, float* pool,
                                               float* save_mean, float* save_rstd, float eps, float momentum,
                                               int64_t* const* num_batches_tracked, float* save_gram, int* gram_saved,
                                               void* workspace, size_t workspace_bytes,
                                               const lmnet_dw_dims* dims, int dtype, void* stream) {
    if (gram_saved != nullptr) *gram_saved = 0;
    int rc = dw_validate(dims);
    if (rc != LMNET_OK) return rc;
    if (!x || !z || !save_mean || !save_rstd || !workspace || !dw_params_ok(p, true)) return LMNET_ERR_INVALID_ARG;
    if ((save_gram == nullptr) != (gram_saved == nullptr)) return LMNET_ERR_INVALID_ARG;
    for (int k = 0; k < 4; ++k)
        if ((p->running_mean[k] == nullptr) != (p->running_var[k] == nullptr)) return LMNET_ERR_INVALID_ARG;
    if (workspace_bytes < lmnet_reparam_dw_workspace_bytes(dims, dtype)) return LMNET_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case LMNET_F32: return dw_train_fwd<float>(x, p, u, z, pool, save_mean, save_rstd, eps, momentum, num_batches_tracked, (char*)workspace, dims, st, save_gram, gram_saved);
        case LMNET_BF16: return dw_train_fwd<__nv_bfloat16>(x, p, u, z, pool, save_mean, save_rstd, eps, momentum, num_batches_tracked, (char*)workspace, dims, st, save_gram, gram_saved);
        case LMNET_F16: return dw_train_fwd<__half>(x, p, u, z, pool, save_mean, save_rstd, eps, momentum, num_batches_tracked, (char*)workspace, dims, st, save_gram, gram_saved);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}

extern "C" int lmnet_reparam_dw_gram_floats(void) { return kGramFloats; }

extern "C" int lmnet_reparam_dw_train_bwd_gram(const void* x, const void* u, const void* dz, const float* dpool,
                                               const lmnet_dw_params* p, const float* save_mean, const float* save_rstd,
                                               const float* save_gram, void* dx, const lmnet_dw_grads* g,
                                               void* workspace, size_t workspace_bytes,
                                               const lmnet_dw_dims* dims, int dtype, void* stream) {
    int rc = dw_validate(dims);
    if (rc != LMNET_OK) return rc;
    if (!x || !u || !dz || !dx || !g || !save_mean || !save_rstd || !workspace || !dw_params_ok(p, true)) return LMNET_ERR_INVALID_ARG;
    if (workspace_bytes < lmnet_reparam_dw_workspace_bytes(dims, dtype)) return LMNET_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case LMNET_F32: return dw_train_bwd<float>(x, u, dz, dpool, p, save_mean, save_rstd, dx, g, (char*)workspace, dims, st, save_gram);
        case LMNET_BF16: return dw_train_bwd<__nv_bfloat16>(x, u, dz, dpool, p, save_mean, save_rstd, dx, g, (char*)workspace, dims, st, save_gram);
        case LMNET_F16: return dw_train_bwd<__half>(x, u, dz, dpool, p, save_mean, save_rstd, dx, g, (char*)workspace, dims, st, save_gram);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}

extern "C" int lmnet_reparam_dw_train_bwd(const void* x, const void* u, const void* dz, const float* dpool,
                                          const lmnet_dw_params* p, const float* save_mean, const float* save_rstd,
                                          void* dx, const lmnet_dw_grads* g,
                                          void* workspace, size_t workspace_bytes,
                                          const lmnet_dw_dims* dims, int dtype, void* stream) {
    int rc = dw_validate(dims);
    if (rc != LMNET_OK) return rc;
    if (!x || !u || !dz || !dx || !g || !save_mean || !save_rstd || !workspace || !dw_params_ok(p, true)) return LMNET_ERR_INVALID_ARG;
    if (workspace_bytes < lmnet_reparam_dw_workspace_bytes(dims, dtype)) return LMNET_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case LMNET_F32: return dw_train_bwd<float>(x, u, dz, dpool, p, save_mean, save_rstd, dx, g, (char*)workspace, dims, st);
        case LMNET_BF16: return dw_train_bwd<__nv_bfloat16>(x, u, dz, dpool, p, save_mean, save_rstd, dx, g, (char*)workspace, dims, st);
        case LMNET_F16: return dw_train_bwd<__half>(x, u, dz, dpool, p, save_mean, save_rstd, dx, g, (char*)workspace, dims, st);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}

extern "C" int lmnet_reparam_dw_eval_fwd(const void* x, const lmnet_dw_params* p, const float* bias, float eps,
                                         void* z, float* pool, void* workspace, size_t workspace_bytes,
                                         const lmnet_dw_dims* dims, int dtype, void* stream) {
    int rc = dw_validate(dims);
    if (rc != LMNET_OK) return rc;
    if (!x || !z || !workspace || p == nullptr || p->w[0] == nullptr) return LMNET_ERR_INVALID_ARG;
    const bool deploy = p->gamma[0] == nullptr;
    if (!deploy) {
        if (!dw_params_ok(p, true)) return LMNET_ERR_INVALID_ARG;
        for (int k = 0; k < 4; ++k)
            if (p->running_mean[k] == nullptr || p->running_var[k] == nullptr) return LMNET_ERR_INVALID_ARG;
    }
    if (workspace_bytes < lmnet_reparam_dw_workspace_bytes(dims, dtype)) return LMNET_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case LMNET_F32: return dw_eval_fwd<float>(x, p, bias, eps, z, pool, (char*)workspace, dims, st);
        case LMNET_BF16: return dw_eval_fwd<__nv_bfloat16>(x, p, bias, eps, z, pool, (char*)workspace, dims, st);
        case LMNET_F16: return dw_eval_fwd<__half>(x, p, bias, eps, z, pool, (char*)workspace, dims, st);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}
