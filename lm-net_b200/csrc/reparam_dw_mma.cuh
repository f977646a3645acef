// reparam_dw_mma.cuh — tensor-core (HMMA) versions of the depthwise branch kernels for 16-bit storage.
//
// Why: per 2-byte element the branch section needs 40 taps (statistics) / 25 taps + erf (apply); on the
// fp32 pipe that is >= 40 FLOP/B, far above its ~11 FLOP/B ridge, so the FFMA2 kernels of
// reparam_dw.cu are FP32-issue bound at < 10 % of HBM peak (profiles/r01_ncu_dw_*).  A depthwise
// stencil is a product with a banded Toeplitz matrix: for each row offset a,
//     Y[r][c] += sum_k X[r+a][k] * T_a[k][c],   T_a[k][c] = w[a][k-c] for 0 <= k-c <= 4, else 0,
// which is an m16n8k16 MMA with M = 16 image rows, K = 16 input columns, N = 8 output columns.
// 5 MMAs give the 5x5 branch of a 16x8 output block (12 MMAs give all four branches), the
// accumulators stay fp32, the operands are exactly what the reference's bf16 cuDNN path uses (bf16
// activations, bf16-rounded weights, fp32 accumulate).  The A fragments come straight out of the bf16
// image tile in shared memory via ldmatrix (no fp32 staging, half the shared-memory traffic); the B
// fragments (Toeplitz bands) are built once per CTA in registers.
//
// Scope note (tier rules): this is NOT a reshaping of a bandwidth-bound kernel into a GEMM to "reach
// tensor cores" — it moves a compute-bound stencil off the saturated fp32 pipe so that it can become
// bandwidth-bound at all.  mma.sync is used (per-warp 16x8 blocks with halo reuse fit it naturally;
// tcgen05's 128-row TMEM tiles do not help a 5-tap band).
#pragma once
#include "common.cuh"

namespace lmnet {

constexpr int kMmaPitch = 72;                 // bf16 elements per shared row (144 B: conflict-free ldmatrix)
constexpr int kMmaTH = 32, kMmaTW = 64;       // output tile
constexpr int kMmaTileRows = kMmaTH + 4;
constexpr int kMmaPairs = kMmaPitch / 2;      // 32-bit pairs per row
#ifndef LMNET_DW_MINBLOCKS
#define LMNET_DW_MINBLOCKS 1
#endif

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
template <typename T> struct MmaOp;
template <> struct MmaOp<__nv_bfloat16> {
    __device__ __forceinline__ static void run(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    __device__ __forceinline__ static uint32_t pack(float lo, float hi) {
        __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&v);
    }
    __device__ __forceinline__ static float round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
};
template <> struct MmaOp<__half> {
    __device__ __forceinline__ static void run(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    __device__ __forceinline__ static uint32_t pack(float lo, float hi) {
        __half2 v = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&v);
    }
    __device__ __forceinline__ static float round(float x) { return __half2float(__float2half_rn(x)); }
};

// B fragment of the Toeplitz band of one kernel row: taps[j] multiplies input column (c + j + shift).
// For lane (g = lane/4, t = lane%4): b0 = T[2t][g], b1 = T[2t+1][g], b2 = T[2t+8][g], b3 = T[2t+9][g],
// T[k][c] = taps[k - c - shift] when 0 <= k-c-shift < ntaps.
template <typename T>
__device__ __forceinline__ void toeplitz_frag(const float* taps, int ntaps, int shift, int lane, uint32_t (&b)[2]) {
    const int g = lane >> 2, t = lane & 3;
    auto tap = [&](int k) -> float {
        const int j = k - g - shift;
        return (j >= 0 && j < ntaps) ? taps[j] : 0.f;
    };
    b[0] = MmaOp<T>::pack(tap(2 * t), tap(2 * t + 1));
    b[1] = MmaOp<T>::pack(tap(2 * t + 8), tap(2 * t + 9));
}

// Raw 16-bit tile (rows x kMmaPitch) staged with 32-bit accesses: index i of a row <-> image column c0 + i.
// Thread layout: 36 column lanes (one 32-bit pair each) x 3 row lanes; a thread walks rows rl, rl+3, ... so that
// its global pointer and shared index advance by constants (no per-slot division, one unsigned row check).
template <typename T, int ROWS>
struct MmaTileLoader {
    static constexpr int RL = 3;
    static constexpr int N = (ROWS + RL - 1) / RL;
    uint32_t raw[N];
    __device__ __forceinline__ void fetch(const T* __restrict__ plane, int H, int W, int r0, int c0) {
        const int rl = threadIdx.x / kMmaPairs, v = threadIdx.x - rl * kMmaPairs;
        const int gc = c0 + 2 * v;
        const bool cok = rl < RL && gc >= 0 && gc + 2 <= W;
        const T* ptr = plane + (int64_t)(r0 + rl) * W + gc;
#pragma unroll
        for (int u = 0; u < N; ++u) {
            const int r = rl + RL * u;
            uint32_t val = 0u;
            if (cok && r < ROWS && (unsigned)(r0 + r) < (unsigned)H) val = __ldg(reinterpret_cast<const uint32_t*>(ptr));
            raw[u] = val;
            ptr += RL * W;
        }
    }
    __device__ __forceinline__ void commit(T* s) const {
        const int rl = threadIdx.x / kMmaPairs, v = threadIdx.x - rl * kMmaPairs;
        if (rl >= RL) return;
        uint32_t* dst = reinterpret_cast<uint32_t*>(s) + rl * kMmaPairs + v;
#pragma unroll
        for (int u = 0; u < N; ++u)
            if (rl + RL * u < ROWS) dst[u * RL * kMmaPairs] = raw[u];
    }
};

template <typename T, int ROWS, typename PlaneFn, typename Body>
__device__ __forceinline__ void mma_for_each_tile(const DwGeom& g, int band0, int band1, int th, int halo, int c0,
                                                  T* s_tile, PlaneFn&& plane_of, Body&& body) {
    const int ntr = (band1 - band0 + th - 1) / th;
    const int total = band1 > band0 ? g.B * ntr : 0;
    MmaTileLoader<T, ROWS> ld;
    if (total > 0) ld.fetch(plane_of(0), g.H, g.W, band0 - halo, c0 - halo);
    for (int t = 0; t < total; ++t) {
        const int b = t / ntr, k = t - b * ntr;
        __syncthreads();
        ld.commit(s_tile);
        __syncthreads();
        if (t + 1 < total) {
            const int nb = (t + 1) / ntr, nk = (t + 1) - nb * ntr;
            ld.fetch(plane_of(nb), g.H, g.W, band0 + nk * th - halo, c0 - halo);
        }
        body(b, band0 + k * th, k == ntr - 1);
    }
}

// A fragment: 16 rows starting at tile row `row`, 16 input columns starting at tile index `col`
template <typename T>
__device__ __forceinline__ void load_a(const T* s_tile, int row, int col, int lane, uint32_t (&a)[4]) {
    const int m = lane >> 3, rr = lane & 7;
    ldmatrix_x4(a, s_tile + (row + (m & 1) * 8 + rr) * kMmaPitch + col + (m >> 1) * 8);
}

// ---------------------------------------------------------------------------------------------------
// statistics pass: per-channel sum / sum of squares of the four branch outputs
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kDwThreads, LMNET_DW_MINBLOCKS)
dw_stats_mma_kernel(const T* __restrict__ x, lmnet_dw_params p, float* __restrict__ part /* [E][ncta][8] */, DwGeom g) {
    __shared__ __align__(16) T s_tile[kMmaTileRows * kMmaPitch];
    __shared__ float s_red[kDwWarps * 8];
    const int e = blockIdx.z, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wr = warp >> 1, wc = warp & 1;               // warp owns rows 16*wr.., columns 32*wc..
    const int c0 = blockIdx.x * kMmaTW;
    const int band0 = blockIdx.y * g.rows_per_band, band1 = min(band0 + g.rows_per_band, g.H);
    // Toeplitz fragments: 5x5 rows a=0..4; 3x3 / 3x1 rows a=1..3; 1x3 row a=2
    uint32_t B5[5][2], B3[3][2], B31[3][2], B13[2];
    {
        float w5[25], w3[9], w31[3], w13[3];
#pragma unroll
        for (int t = 0; t < 25; ++t) w5[t] = __ldg(p.w[0] + e * 25 + t);
#pragma unroll
        for (int t = 0; t < 9; ++t) w3[t] = __ldg(p.w[1] + e * 9 + t);
#pragma unroll
        for (int t = 0; t < 3; ++t) { w31[t] = __ldg(p.w[2] + e * 3 + t); w13[t] = __ldg(p.w[3] + e * 3 + t); }
#pragma unroll
        for (int a = 0; a < 5; ++a) toeplitz_frag<T>(w5 + a * 5, 5, 0, lane, B5[a]);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            toeplitz_frag<T>(w3 + a * 3, 3, 1, lane, B3[a]);
            toeplitz_frag<T>(w31 + a, 1, 2, lane, B31[a]);
        }
        toeplitz_frag<T>(w13, 3, 1, lane, B13);
    }
    float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
    const int gq = lane >> 2, tq = lane & 3;
    mma_for_each_tile<T, kMmaTileRows>(
        g, band0, band1, kMmaTH, 2, c0, s_tile,
        [&](int b) { return x + ((int64_t)b * g.E + e) * g.H * g.W; },
        [&](int, int tr, bool) {
            const int row_lo = tr + 16 * wr + gq;            // image rows of c0/c1 and (+8) c2/c3
            const float rm0 = row_lo < band1 ? 1.f : 0.f, rm1 = row_lo + 8 < band1 ? 1.f : 0.f;
            const bool full = tr + kMmaTH <= band1 && c0 + kMmaTW <= g.W;
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) {
                float acc[4][4];
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f;
                const int tcol = 32 * wc + 8 * cb;           // tile index of input column (cb_col - 2)
#pragma unroll
                for (int a = 0; a < 5; ++a) {
                    uint32_t A[4];
                    load_a(s_tile, 16 * wr + a, tcol, lane, A);
                    MmaOp<T>::run(acc[0], A, B5[a]);
                    if (a >= 1 && a <= 3) {
                        MmaOp<T>::run(acc[1], A, B3[a - 1]);
                        MmaOp<T>::run(acc[2], A, B31[a - 1]);
                    }
                    if (a == 2) MmaOp<T>::run(acc[3], A, B13);
                }
                if (full) {                                   // tile inside the band and the image: no masks
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            s[k] += acc[k][i];
                            ss[k] = fmaf(acc[k][i], acc[k][i], ss[k]);
                        }
                } else {
                    const int col = c0 + tcol + 2 * tq;
                    const float cm0 = col < g.W ? 1.f : 0.f, cm1 = col + 1 < g.W ? 1.f : 0.f;
                    const float m[4] = {rm0 * cm0, rm0 * cm1, rm1 * cm0, rm1 * cm1};
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float y = acc[k][i] * m[i];
                            s[k] += y;
                            ss[k] = fmaf(y, acc[k][i], ss[k]);
                        }
                }
            }
        });
    float v[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = s[k]; v[4 + k] = ss[k]; }
    const int ncta = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
    block_sum<8>(v, s_red, part + ((int64_t)e * ncta + cta) * 8);
}

// ---------------------------------------------------------------------------------------------------
// apply pass: u = merged5x5(x) + bias; z = GELU(u); pool partial sums.  The merged fp32 taps are split
// into hi + lo 16-bit parts (two MMAs per kernel row) so that folding the BatchNorm scales adds no
// rounding beyond fp32.
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kDwThreads, LMNET_DW_MINBLOCKS)
dw_apply_mma_kernel(const T* __restrict__ x, const float* __restrict__ coef, T* __restrict__ u_out, T* __restrict__ z_out,
                    float* __restrict__ pool_part, DwGeom g) {
    __shared__ __align__(16) T s_tile[kMmaTileRows * kMmaPitch];
    __shared__ float s_red[kDwWarps];
    const int e = blockIdx.z, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wr = warp >> 1, wc = warp & 1;
    const int c0 = blockIdx.x * kMmaTW;
    const int band0 = blockIdx.y * g.rows_per_band, band1 = min(band0 + g.rows_per_band, g.H);
    uint32_t Bhi[5][2], Blo[5][2];
    {
        float wm[25], hi[5], lo[5];
#pragma unroll
        for (int t = 0; t < 25; ++t) wm[t] = __ldg(coef + e * 26 + t);
#pragma unroll
        for (int a = 0; a < 5; ++a) {
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                hi[j] = MmaOp<T>::round(wm[a * 5 + j]);
                lo[j] = wm[a * 5 + j] - hi[j];
            }
            toeplitz_frag<T>(hi, 5, 0, lane, Bhi[a]);
            toeplitz_frag<T>(lo, 5, 0, lane, Blo[a]);
        }
    }
    const float bias = __ldg(coef + e * 26 + 25);
    const int gq = lane >> 2, tq = lane & 3;
    const int ncta = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
    float psum = 0.f;
    mma_for_each_tile<T, kMmaTileRows>(
        g, band0, band1, kMmaTH, 2, c0, s_tile,
        [&](int b) { return x + ((int64_t)b * g.E + e) * g.H * g.W; },
        [&](int b, int tr, bool last) {
            const int row_lo = tr + 16 * wr + gq;
            const int col_lo = c0 + 32 * wc + 2 * tq;
            // element offset of this lane's first output pair; the other seven pairs are +8*cb columns / +8 rows away
            const int64_t base = ((int64_t)b * g.E + e) * g.H * g.W + (int64_t)row_lo * g.W + col_lo;
            const int64_t row8 = (int64_t)8 * g.W;
            const bool full = tr + kMmaTH <= band1 && c0 + kMmaTW <= g.W;
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) {
                float acc[4] = {bias, bias, bias, bias};
                const int tcol = 32 * wc + 8 * cb;
#pragma unroll
                for (int a = 0; a < 5; ++a) {
                    uint32_t A[4];
                    load_a(s_tile, 16 * wr + a, tcol, lane, A);
                    MmaOp<T>::run(acc, A, Bhi[a]);
                    MmaOp<T>::run(acc, A, Blo[a]);
                }
                const bool cok = full || (col_lo + 8 * cb < g.W);         // W even: the pair is all in or all out
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (cok && (full || row_lo + 8 * h < band1)) {
                        const int64_t off = base + 8 * cb + h * row8;
                        const uint32_t up = MmaOp<T>::pack(acc[2 * h], acc[2 * h + 1]);
                        if (u_out != nullptr) *reinterpret_cast<uint32_t*>(u_out + off) = up;
                        const T* ur = reinterpret_cast<const T*>(&up);      // GELU of the value as stored
                        const float z0 = gelu_f(to_f(ur[0])), z1 = gelu_f(to_f(ur[1]));
                        *reinterpret_cast<uint32_t*>(z_out + off) = MmaOp<T>::pack(z0, z1);
                        psum += z0 + z1;
                    }
                }
            }
            if (last && pool_part != nullptr) {
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

// ---------------------------------------------------------------------------------------------------
// backward R pass on tensor cores: du = (dz + dpool/HW) * gelu'(u) (written out), sum du, and the 25 lag
// sums P[a][b] = sum_p du(p) * x(p + (a-2, b-2)) as Gram products: for each row offset a
//     G_a[i][j] += sum_r X[r + a][i] * du[r][j]      (M = 16 input columns, N = 8 output columns, K = 16 rows)
// and P[a][b] = sum_j G_a[j + b][j].  Both operands are read transposed out of shared memory
// (ldmatrix.trans); the five 16x8 accumulators persist over all tiles of the CTA.
// ---------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void load_a_trans(const T* s_tile, int row, int col, int lane, uint32_t (&a)[4]) {
    const int m = lane >> 3, rr = lane & 7;
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(s_tile + (row + (m >> 1) * 8 + rr) * kMmaPitch + col + (m & 1) * 8);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
}
template <typename T>
__device__ __forceinline__ void load_b_trans(const T* s_tile, int row, int col, int lane, uint32_t (&b)[2]) {
    const int l = lane & 15;
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(s_tile + (row + l) * kMmaPitch + col);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(b[0]), "=r"(b[1]) : "r"(addr));
}

template <typename T>
__global__ void __launch_bounds__(kDwThreads, LMNET_DW_MINBLOCKS)
dw_bwd_reduce_mma_kernel(const T* __restrict__ x, const T* __restrict__ u, const T* __restrict__ dz,
                         const float* __restrict__ dpool, T* __restrict__ du_out, float* __restrict__ part /* [E][ncta][26] */,
                         DwGeom g) {
    __shared__ __align__(16) T s_x[kMmaTileRows * kMmaPitch];
    __shared__ __align__(16) T s_u[kMmaTH * kMmaPitch];
    __shared__ __align__(16) T s_dz[kMmaTH * kMmaPitch];
    __shared__ __align__(16) T s_du[kMmaTH * kMmaPitch];
    __shared__ float s_P[26];
    const int e = blockIdx.z, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wr = warp >> 1, wc = warp & 1;
    const int c0 = blockIdx.x * kMmaTW;
    const int band0 = blockIdx.y * g.rows_per_band, band1 = min(band0 + g.rows_per_band, g.H);
    const float inv_hw = 1.f / ((float)g.H * (float)g.W);
    if (threadIdx.x < 26) s_P[threadIdx.x] = 0.f;
    float G[5][4];
#pragma unroll
    for (int a = 0; a < 5; ++a) G[a][0] = G[a][1] = G[a][2] = G[a][3] = 0.f;
    float sdu = 0.f;

    const int ntr = (band1 - band0 + kMmaTH - 1) / kMmaTH;
    const int total = band1 > band0 ? g.B * ntr : 0;
    MmaTileLoader<T, kMmaTileRows> lx;
    MmaTileLoader<T, kMmaTH> lu, lz;
    auto fetch = [&](int t) {
        const int b = t / ntr, tr = band0 + (t - b * ntr) * kMmaTH;
        const int64_t poff = ((int64_t)b * g.E + e) * g.H * g.W;
        lx.fetch(x + poff, g.H, g.W, tr - 2, c0 - 2);
        lu.fetch(u + poff, g.H, g.W, tr, c0);
        lz.fetch(dz + poff, g.H, g.W, tr, c0);
    };
    if (total > 0) fetch(0);
    for (int t = 0; t < total; ++t) {
        const int b = t / ntr, tr = band0 + (t - b * ntr) * kMmaTH;
        const int64_t poff = ((int64_t)b * g.E + e) * g.H * g.W;
        __syncthreads();
        lx.commit(s_x);
        lu.commit(s_u);
        lz.commit(s_dz);
        __syncthreads();
        if (t + 1 < total) fetch(t + 1);
        // phase 1: du for the 32 x 64 tile (thread: row = warp*8 + o, column pair = lane)
        const float dp = dpool != nullptr ? __ldg(dpool + b * g.E + e) * inv_hw : 0.f;
        const int col = c0 + 2 * lane;
        const bool cok = col < g.W;                           // W even: the pair is all in or all out
        const bool full = tr + kMmaTH <= band1;
        T* dptr = du_out + poff + (int64_t)(tr + warp * 8) * g.W + col;
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            const int trow = warp * 8 + o;
            uint32_t packed = 0u;
            if (cok && (full || tr + trow < band1)) {
                const uint32_t ur = *reinterpret_cast<const uint32_t*>(s_u + trow * kMmaPitch + 2 * lane);
                const uint32_t zr = *reinterpret_cast<const uint32_t*>(s_dz + trow * kMmaPitch + 2 * lane);
                const T* ue = reinterpret_cast<const T*>(&ur);
                const T* ze = reinterpret_cast<const T*>(&zr);
                const float d0 = (to_f(ze[0]) + dp) * gelu_grad_f(to_f(ue[0]));
                const float d1 = (to_f(ze[1]) + dp) * gelu_grad_f(to_f(ue[1]));
                packed = MmaOp<T>::pack(d0, d1);
                *reinterpret_cast<uint32_t*>(dptr) = packed;
                const T* de = reinterpret_cast<const T*>(&packed);
                sdu += to_f(de[0]) + to_f(de[1]);             // exactly what later passes read back
            }
            *reinterpret_cast<uint32_t*>(s_du + trow * kMmaPitch + 2 * lane) = packed;
            dptr += g.W;
        }
        __syncthreads();
        // phase 2: Gram MMAs of this warp's 16 rows x 32 columns
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
            const int tcol = 32 * wc + 8 * cb;
            uint32_t Bf[2];
            load_b_trans(s_du, 16 * wr, tcol, lane, Bf);
#pragma unroll
            for (int a = 0; a < 5; ++a) {
                uint32_t A[4];
                load_a_trans(s_x, 16 * wr + a, tcol, lane, A);
                MmaOp<T>::run(G[a], A, Bf);
            }
        }
    }
    // P[a][b] = sum_j G_a[j + b][j]: this lane holds (i = gq, gq+8 ; j = 2tq, 2tq+1)
    __syncthreads();
    const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = gq + (k >> 1) * 8, j = 2 * tq + (k & 1);
            const int bb = i - j;
            if (bb >= 0 && bb < 5) atomicAdd(&s_P[a * 5 + bb], G[a][k]);
        }
    sdu = warp_sum(sdu);
    if (lane == 0) atomicAdd(&s_P[25], sdu);
    __syncthreads();
    const int ncta = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
    if (threadIdx.x < 26) part[((int64_t)e * ncta + cta) * 26 + threadIdx.x] = s_P[threadIdx.x];
}

// ---------------------------------------------------------------------------------------------------
// backward A2 pass on tensor cores: R_br[t] = sum_p (c2_br*y_br(p) + c0_br) * x(p+t).
// y_br from the Toeplitz MMAs (as in the statistics pass), g_br = c2*y + c0 rounded to the storage type
// into shared memory, then Gram MMAs  G_{br,a}[i][j] += sum_r X[r+a][i] * g_br[r][j]  with the twelve
// (branch, kernel-row) accumulators kept in registers for the whole CTA; R_br[a][b] = sum_j G[j+b][j].
// (g_br is the small BatchNorm-variance correction of the weight gradient: 16-bit rounding of it is far
// below the tolerance; the main term c1*P[t] comes from the reduce pass in fp32.)
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kDwThreads)
dw_bwd_dw_mma_kernel(const T* __restrict__ x, lmnet_dw_params p, const float* __restrict__ cb,
                     float* __restrict__ part /* [E][ncta][40] */, DwGeom g) {
    __shared__ __align__(16) T s_x[kMmaTileRows * kMmaPitch];
    __shared__ __align__(16) T s_g[4][kMmaTH * kMmaPitch];
    __shared__ float s_R[40];
    const int e = blockIdx.z, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wr = warp >> 1, wc = warp & 1;
    const int c0 = blockIdx.x * kMmaTW;
    const int band0 = blockIdx.y * g.rows_per_band, band1 = min(band0 + g.rows_per_band, g.H);
    if (threadIdx.x < 40) s_R[threadIdx.x] = 0.f;
    uint32_t B5[5][2], B3[3][2], B31[3][2], B13[2];
    {
        float w5[25], w3[9], w31[3], w13[3];
#pragma unroll
        for (int t = 0; t < 25; ++t) w5[t] = __ldg(p.w[0] + e * 25 + t);
#pragma unroll
        for (int t = 0; t < 9; ++t) w3[t] = __ldg(p.w[1] + e * 9 + t);
#pragma unroll
        for (int t = 0; t < 3; ++t) { w31[t] = __ldg(p.w[2] + e * 3 + t); w13[t] = __ldg(p.w[3] + e * 3 + t); }
#pragma unroll
        for (int a = 0; a < 5; ++a) toeplitz_frag<T>(w5 + a * 5, 5, 0, lane, B5[a]);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            toeplitz_frag<T>(w3 + a * 3, 3, 1, lane, B3[a]);
            toeplitz_frag<T>(w31 + a, 1, 2, lane, B31[a]);
        }
        toeplitz_frag<T>(w13, 3, 1, lane, B13);
    }
    float c2[4], c0c[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        c2[k] = __ldg(cb + e * 12 + k * 3 + 1);
        c0c[k] = __ldg(cb + e * 12 + k * 3 + 2);
    }
    // Gram accumulators: 5x5 rows a=0..4 | 3x3 rows a=1..3 | 3x1 rows a=1..3 | 1x3 row a=2
    float G5[5][4], G3[3][4], G31[3][4], G13[4];
#pragma unroll
    for (int a = 0; a < 5; ++a) G5[a][0] = G5[a][1] = G5[a][2] = G5[a][3] = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        G3[a][0] = G3[a][1] = G3[a][2] = G3[a][3] = 0.f;
        G31[a][0] = G31[a][1] = G31[a][2] = G31[a][3] = 0.f;
    }
    G13[0] = G13[1] = G13[2] = G13[3] = 0.f;
    const int gq = lane >> 2, tq = lane & 3;
    mma_for_each_tile<T, kMmaTileRows>(
        g, band0, band1, kMmaTH, 2, c0, s_x,
        [&](int b) { return x + ((int64_t)b * g.E + e) * g.H * g.W; },
        [&](int, int tr, bool) {
            const int row_lo = tr + 16 * wr + gq;
            const float rm0 = row_lo < band1 ? 1.f : 0.f, rm1 = row_lo + 8 < band1 ? 1.f : 0.f;
            const bool full = tr + kMmaTH <= band1 && c0 + kMmaTW <= g.W;
            // phase 1: y_br of this warp's 16 x 32 block -> g_br tiles
#pragma unroll
            for (int cbk = 0; cbk < 4; ++cbk) {
                float acc[4][4];
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f;
                const int tcol = 32 * wc + 8 * cbk;
#pragma unroll
                for (int a = 0; a < 5; ++a) {
                    uint32_t A[4];
                    load_a(s_x, 16 * wr + a, tcol, lane, A);
                    MmaOp<T>::run(acc[0], A, B5[a]);
                    if (a >= 1 && a <= 3) {
                        MmaOp<T>::run(acc[1], A, B3[a - 1]);
                        MmaOp<T>::run(acc[2], A, B31[a - 1]);
                    }
                    if (a == 2) MmaOp<T>::run(acc[3], A, B13);
                }
                const int col = c0 + tcol + 2 * tq;
                const float cm = (full || col < g.W) ? 1.f : 0.f;     // W even: pair in or out together
                const float m0 = full ? 1.f : rm0 * cm, m1 = full ? 1.f : rm1 * cm;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t lo = MmaOp<T>::pack(fmaf(c2[k], acc[k][0], c0c[k]) * m0, fmaf(c2[k], acc[k][1], c0c[k]) * m0);
                    const uint32_t hi = MmaOp<T>::pack(fmaf(c2[k], acc[k][2], c0c[k]) * m1, fmaf(c2[k], acc[k][3], c0c[k]) * m1);
                    *reinterpret_cast<uint32_t*>(&s_g[k][(16 * wr + gq) * kMmaPitch + tcol + 2 * tq]) = lo;
                    *reinterpret_cast<uint32_t*>(&s_g[k][(16 * wr + gq + 8) * kMmaPitch + tcol + 2 * tq]) = hi;
                }
            }
            __syncwarp();      // each warp reads back only the 16 x 32 block it wrote
            // phase 2: Gram products
#pragma unroll
            for (int cbk = 0; cbk < 4; ++cbk) {
                const int tcol = 32 * wc + 8 * cbk;
                uint32_t Bg[4][2];
#pragma unroll
                for (int k = 0; k < 4; ++k) load_b_trans(s_g[k], 16 * wr, tcol, lane, Bg[k]);
#pragma unroll
                for (int a = 0; a < 5; ++a) {
                    uint32_t A[4];
                    load_a_trans(s_x, 16 * wr + a, tcol, lane, A);
                    MmaOp<T>::run(G5[a], A, Bg[0]);
                    if (a >= 1 && a <= 3) {
                        MmaOp<T>::run(G3[a - 1], A, Bg[1]);
                        MmaOp<T>::run(G31[a - 1], A, Bg[2]);
                    }
                    if (a == 2) MmaOp<T>::run(G13, A, Bg[3]);
                }
            }
        });
    // R_br[a][b] = sum_j G[j + b][j]; this lane holds (i = gq, gq+8 ; j = 2tq, 2tq+1); b = i - j is the 5-wide offset
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = gq + (k >> 1) * 8, j = 2 * tq + (k & 1);
        const int bb = i - j;
        if (bb >= 0 && bb < 5) {
#pragma unroll
            for (int a = 0; a < 5; ++a) atomicAdd(&s_R[a * 5 + bb], G5[a][k]);
            if (bb >= 1 && bb <= 3) {
#pragma unroll
                for (int a = 0; a < 3; ++a) atomicAdd(&s_R[25 + a * 3 + (bb - 1)], G3[a][k]);
                atomicAdd(&s_R[37 + (bb - 1)], G13[k]);
            }
            if (bb == 2) {
#pragma unroll
                for (int a = 0; a < 3; ++a) atomicAdd(&s_R[34 + a], G31[a][k]);
            }
        }
    }
    __syncthreads();
    const int ncta = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
    if (threadIdx.x < 40) part[((int64_t)e * ncta + cta) * 40 + threadIdx.x] = s_R[threadIdx.x];
}

// ---------------------------------------------------------------------------------------------------
// backward A1 pass on tensor cores: dx = sum_br w_br (*)^T dy_br,  dy_br = c1*du - c2*y_br - c0 in the image.
// Per CTA tile: dx block of 28 x 56 pixels; dy region 32 x 64 with origin (tr-2, c0-2); x tile with origin
// (tr-4, c0-4).  Phase 1: y_br on the region by Toeplitz MMAs -> dy_br (storage type) into shared memory.
// Phase 2: dx by Toeplitz MMAs with the FLIPPED taps over the four dy tiles, all branches into one accumulator.
// ---------------------------------------------------------------------------------------------------
constexpr int kDxTH = 28, kDxTW = 56;

template <typename T>
__global__ void __launch_bounds__(kDwThreads, LMNET_DW_MINBLOCKS)
dw_bwd_dx_mma_kernel(const T* __restrict__ x, const T* __restrict__ du, lmnet_dw_params p, const float* __restrict__ cb,
                     T* __restrict__ dx, DwGeom g) {
    extern __shared__ __align__(16) unsigned char dx_smem_raw[];
    T* s_x = reinterpret_cast<T*>(dx_smem_raw);                 // [36][72]
    T* s_du = s_x + kMmaTileRows * kMmaPitch;                   // [32][72]
    T* s_dy = s_du + kMmaTH * kMmaPitch;                        // [4][36][72]
    const int e = blockIdx.z, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wr = warp >> 1, wc = warp & 1;
    const int c0 = blockIdx.x * kDxTW;
    const int band0 = blockIdx.y * g.rows_per_band, band1 = min(band0 + g.rows_per_band, g.H);
    uint32_t B5[5][2], B3[3][2], B31[3][2], B13[2];           // forward taps (phase 1)
    uint32_t F5[5][2], F3[3][2], F31[3][2], F13[2];           // flipped taps (phase 2), indexed by a' = row offset
    {
        float w5[25], w3[9], w31[3], w13[3], f[5];
#pragma unroll
        for (int t = 0; t < 25; ++t) w5[t] = __ldg(p.w[0] + e * 25 + t);
#pragma unroll
        for (int t = 0; t < 9; ++t) w3[t] = __ldg(p.w[1] + e * 9 + t);
#pragma unroll
        for (int t = 0; t < 3; ++t) { w31[t] = __ldg(p.w[2] + e * 3 + t); w13[t] = __ldg(p.w[3] + e * 3 + t); }
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            toeplitz_frag<T>(w5 + a * 5, 5, 0, lane, B5[a]);
#pragma unroll
            for (int j = 0; j < 5; ++j) f[j] = w5[(4 - a) * 5 + (4 - j)];
            toeplitz_frag<T>(f, 5, 0, lane, F5[a]);
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {                          // a <-> a' = a + 1
            toeplitz_frag<T>(w3 + a * 3, 3, 1, lane, B3[a]);
            toeplitz_frag<T>(w31 + a, 1, 2, lane, B31[a]);
#pragma unroll
            for (int j = 0; j < 3; ++j) f[j] = w3[(2 - a) * 3 + (2 - j)];
            toeplitz_frag<T>(f, 3, 1, lane, F3[a]);
            f[0] = w31[2 - a];
            toeplitz_frag<T>(f, 1, 2, lane, F31[a]);
        }
        toeplitz_frag<T>(w13, 3, 1, lane, B13);
#pragma unroll
        for (int j = 0; j < 3; ++j) f[j] = w13[2 - j];
        toeplitz_frag<T>(f, 3, 1, lane, F13);
    }
    float c1[4], c2[4], c0c[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        c1[k] = __ldg(cb + e * 12 + k * 3);
        c2[k] = __ldg(cb + e * 12 + k * 3 + 1);
        c0c[k] = __ldg(cb + e * 12 + k * 3 + 2);
    }
    // rows 32..35 of the dy tiles are read (for discarded outputs only) but never written: keep them finite
    for (int i = threadIdx.x; i < 4 * 4 * kMmaPitch; i += kDwThreads) {
        const int k = i / (4 * kMmaPitch), r = i - k * 4 * kMmaPitch;
        s_dy[k * kMmaTileRows * kMmaPitch + kMmaTH * kMmaPitch + r] = from_f<T>(0.f);
    }
    const int gq = lane >> 2, tq = lane & 3;
    const int ntr = (band1 - band0 + kDxTH - 1) / kDxTH;
    const int total = band1 > band0 ? g.B * ntr : 0;
    MmaTileLoader<T, kMmaTileRows> lx;
    MmaTileLoader<T, kMmaTH> ld;
    auto fetch = [&](int t) {
        const int b = t / ntr, tr = band0 + (t - b * ntr) * kDxTH;
        const int64_t poff = ((int64_t)b * g.E + e) * g.H * g.W;
        lx.fetch(x + poff, g.H, g.W, tr - 4, c0 - 4);
        ld.fetch(du + poff, g.H, g.W, tr - 2, c0 - 2);
    };
    if (total > 0) fetch(0);
    for (int t = 0; t < total; ++t) {
        const int b = t / ntr, tr = band0 + (t - b * ntr) * kDxTH;
        const int64_t poff = ((int64_t)b * g.E + e) * g.H * g.W;
        __syncthreads();
        lx.commit(s_x);
        ld.commit(s_du);
        __syncthreads();
        if (t + 1 < total) fetch(t + 1);
        // phase 1: dy_br on this warp's 16 x 32 block of the region
        {
            const int rho0 = 16 * wr + gq;                         // region rows rho0, rho0 + 8
            const int row0 = tr - 2 + rho0, row1 = row0 + 8;
            const bool rin0 = row0 >= 0 && row0 < g.H, rin1 = row1 >= 0 && row1 < g.H;
            const bool inside = tr - 2 >= 0 && tr - 2 + kMmaTH <= g.H && c0 - 2 >= 0 && c0 - 2 + kMmaTW <= g.W;
#pragma unroll
            for (int cbk = 0; cbk < 4; ++cbk) {
                float acc[4][4];
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f;
                const int tcol = 32 * wc + 8 * cbk;                // region column of the block = x-tile index
#pragma unroll
                for (int a = 0; a < 5; ++a) {
                    uint32_t A[4];
                    load_a(s_x, 16 * wr + a, tcol, lane, A);
                    MmaOp<T>::run(acc[0], A, B5[a]);
                    if (a >= 1 && a <= 3) {
                        MmaOp<T>::run(acc[1], A, B3[a - 1]);
                        MmaOp<T>::run(acc[2], A, B31[a - 1]);
                    }
                    if (a == 2) MmaOp<T>::run(acc[3], A, B13);
                }
                const int kap = tcol + 2 * tq;                     // region column of this lane's pair
                const int col = c0 - 2 + kap;
                const bool cin = col >= 0 && col < g.W;            // W even, col even: pair in or out together
                const uint32_t d0r = *reinterpret_cast<const uint32_t*>(s_du + rho0 * kMmaPitch + kap);
                const uint32_t d1r = *reinterpret_cast<const uint32_t*>(s_du + (rho0 + 8) * kMmaPitch + kap);
                const T* d0 = reinterpret_cast<const T*>(&d0r);
                const T* d1 = reinterpret_cast<const T*>(&d1r);
                const float m0 = (inside || (rin0 && cin)) ? 1.f : 0.f, m1 = (inside || (rin1 && cin)) ? 1.f : 0.f;
                const float da = to_f(d0[0]), db = to_f(d0[1]), dc = to_f(d1[0]), dd = to_f(d1[1]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t lo = MmaOp<T>::pack(fmaf(c1[k], da, fmaf(-c2[k], acc[k][0], -c0c[k])) * m0,
                                                       fmaf(c1[k], db, fmaf(-c2[k], acc[k][1], -c0c[k])) * m0);
                    const uint32_t hi = MmaOp<T>::pack(fmaf(c1[k], dc, fmaf(-c2[k], acc[k][2], -c0c[k])) * m1,
                                                       fmaf(c1[k], dd, fmaf(-c2[k], acc[k][3], -c0c[k])) * m1);
                    T* tile = s_dy + k * kMmaTileRows * kMmaPitch;
                    *reinterpret_cast<uint32_t*>(tile + rho0 * kMmaPitch + kap) = lo;
                    *reinterpret_cast<uint32_t*>(tile + (rho0 + 8) * kMmaPitch + kap) = hi;
                }
            }
        }
        __syncthreads();
        // phase 2: dx blocks (2 row tiles x 7 column blocks; warp (wr, wc) takes row tile wr, column blocks 4wc..)
#pragma unroll
        for (int cbk = 0; cbk < 4; ++cbk) {
            const int blk = 4 * wc + cbk;
            if (blk < kDxTW / 8) {
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                const int tcol = 8 * blk;                          // dx column c <-> region index c + b'
#pragma unroll
                for (int a = 0; a < 5; ++a) {
                    uint32_t A[4];
                    load_a(s_dy, 16 * wr + a, tcol, lane, A);
                    MmaOp<T>::run(acc, A, F5[a]);
                    if (a >= 1 && a <= 3) {
                        load_a(s_dy + kMmaTileRows * kMmaPitch, 16 * wr + a, tcol, lane, A);
                        MmaOp<T>::run(acc, A, F3[a - 1]);
                        load_a(s_dy + 2 * kMmaTileRows * kMmaPitch, 16 * wr + a, tcol, lane, A);
                        MmaOp<T>::run(acc, A, F31[a - 1]);
                    }
                    if (a == 2) {
                        load_a(s_dy + 3 * kMmaTileRows * kMmaPitch, 16 * wr + a, tcol, lane, A);
                        MmaOp<T>::run(acc, A, F13);
                    }
                }
                const int col = c0 + tcol + 2 * tq;
                if (col < g.W) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int r = 16 * wr + gq + 8 * h, row = tr + r;
                        if (r < kDxTH && row < band1)
                            *reinterpret_cast<uint32_t*>(dx + poff + (int64_t)row * g.W + col) = MmaOp<T>::pack(acc[2 * h], acc[2 * h + 1]);
                    }
                }
            }
        }
    }
}
constexpr size_t kDxMmaSmemBytes = (size_t)(kMmaTileRows + kMmaTH + 4 * kMmaTileRows) * kMmaPitch * 2;

}  // namespace lmnet
