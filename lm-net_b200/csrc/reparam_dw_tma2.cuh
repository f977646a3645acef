// reparam_dw_tma2.cuh — composite-stencil backward of the depthwise branch section (TMA pipeline, 16-bit storage).
//
// The first TMA backward (reparam_dw_tma.cuh::dw_bwd_dx_tma_kernel) recomputes the four branch outputs y_br on tile +
// halo, forms dy_br = c1 du - c2 y_br - c0 in shared memory, walks the four transposed stencils over those tiles and
// correlates them with x for the weight gradients: 36 MMAs per 16 x 8 block in three dependent phases, 168 registers,
// 2 CTAs per SM — issue/latency bound at 0.08 of HBM peak (profiles/r02_ncu_reparam_l1.txt).  Everything in it is
// LINEAR in (du, x) once the per-channel coefficients are known, so the phases collapse algebraically
// (tools/debug/dx_composite_proto.py checks the identities in fp64 against autograd):
//
//   dx(q) = sum_s Wc1[s] du~(q-s)  -  sum_d C[d] x~(q+d)  -  sum_s Wc0[s]  +  frame(q)
//       Wc1 = sum_br c1_br w_br, Wc0 = sum_br c0_br w_br              (merged 5 x 5 kernels)
//       C[d] = sum_br c2_br sum_s w_br[s] w_br[s+d]                   (ONE 9 x 9 kernel: conv_br^T o conv_br)
//       frame(q) = sum_br sum_{s : q-s outside the image} w_br[s] (c2_br y~_br(q-s) + c0_br)
//     (~ = zero-extended.)  The composite kernel counts branch outputs at positions outside the image, which the
//     reference never forms; `frame` removes exactly those terms and is non-zero only on the 2-pixel frame of each
//     plane (2 % of the pixels at 352 x 352) — a tiny scalar kernel patches it after the tile kernel.
//   dw_br[t] = c1_br P[t] - c2_br Q_br[t] - c0_br S[t]
//       P[t] = sum_p du(p) x~(p+t),  S[t] = sum_p x~(p+t)    (backward reduce pass: S is an all-ones Gram product on
//                                                              the fragments the P products load anyway)
//       Q_br[t] = sum_p y_br(p) x~(p+t)                        (a function of x and w only: taken in the FORWARD
//                                                              statistics pass, where y_br exists anyway)
//
// So the backward tile kernel is ONE phase of Toeplitz MMAs (5 rows of du, 9 rows of x; taps rounded to the storage
// type like the reference's own 16-bit convolution weights, -DLMNET_DX2_SPLIT_TAPS adds the rounding remainder) into one
// accumulator set — no dy tiles, no CTA barrier, no Gram accumulators, the apply kernel's register budget (4 CTAs per
// SM) — and the forward statistics kernel gains the Gram products of its own y_br fragments (through a 1 KB per-warp
// shared-memory transpose).
#pragma once
#include "reparam_dw_tma.cuh"

namespace lmnet {

constexpr int kGramFloats = 40;        // per channel: Q5 [25] | Q3 [9] | Q31 [3] | Q13 [3]
constexpr int kGramPartStride = 48;    // per CTA partial record: stats [8] | Q [40]
constexpr int kCoef2Stride = 108;      // per channel: F5 [25] (flipped Wc1) | -C [81] | -sum Wc0 | pad

// ---------------------------------------------------------------------------------------------------
// forward statistics + Gram pass
// ---------------------------------------------------------------------------------------------------
constexpr int kStatsGramStages = 4;
constexpr int kYStageBytes = 4 * 16 * 8 * 2;          // per warp: 4 branches x 16 rows x 8 columns (16-byte rows)
constexpr size_t kStatsGramSmem = 128 + (size_t)kStatsGramStages * kXSlotBytes + kDwWarps * kYStageBytes +
                                  2 * kStatsGramStages * 8 + kDwWarps * 48 * 4;

template <typename T>
__global__ void __launch_bounds__(kTmaThreads, 3)
dw_stats_gram_tma_kernel(const __grid_constant__ CUtensorMap tm_x, lmnet_dw_params p,
                         float* __restrict__ part /* [E][ncta][kGramPartStride] */, DwGeom g) {
    constexpr int S = kStatsGramStages;
    extern __shared__ unsigned char dw_smem_raw[];
    unsigned char* smem = align128(dw_smem_raw);
    T* s_y_all = reinterpret_cast<T*>(smem + S * kXSlotBytes);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S * kXSlotBytes + kDwWarps * kYStageBytes);
    uint64_t* empty = full + S;
    float* s_red = reinterpret_cast<float*>(empty + S);                       // [4 warps][48]
    const int e = blockIdx.z, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * kMmaTW - kFwdShift;
    const int band0 = blockIdx.y * g.rows_per_band, band1 = min(band0 + g.rows_per_band, g.H);
    const int ntr = (band1 - band0 + kMmaTH - 1) / kMmaTH;
    const int total = band1 > band0 ? g.B * ntr : 0;
    ring_init<S>(full, empty);
    if (warp == kDwWarps) {
        if (lane == 0) {
            tma_prefetch_desc(&tm_x);
            RingPos<S> pos;
            TileWalk tw;
            for (int t = 0; t < total; ++t) {
                if (t >= S) mbar_wait(empty + pos.slot, pos.phase ^ 1u);
                mbar_arrive_expect_tx(full + pos.slot, kXTileBytes);
                tma_load_3d(smem + pos.slot * kXSlotBytes, &tm_x, full + pos.slot, c0 - 2, band0 + tw.k * kMmaTH - 2, tw.b * g.E + e);
                pos.advance();
                tw.next(ntr);
            }
        }
    } else {
        const int wr = warp >> 1, wc = warp & 1;
        BranchFrags f;
        load_branch_frags<T>(p, e, lane, f, nullptr);
        float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
        // Gram accumulators: 5x5 rows a=0..4 | 3x3 rows a=1..3 | 3x1 rows a=1..3 | 1x3 row a=2
        float G5[5][4], G3[3][4], G31[3][4], G13[4];
#pragma unroll
        for (int a = 0; a < 5; ++a)
#pragma unroll
            for (int i = 0; i < 4; ++i) G5[a][i] = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int i = 0; i < 4; ++i) G3[a][i] = G31[a][i] = 0.f;
        G13[0] = G13[1] = G13[2] = G13[3] = 0.f;
        const int gq = lane >> 2, tq = lane & 3;
        T* s_y = s_y_all + warp * (kYStageBytes / 2);                        // [4][16][8]
        RingPos<S> pos;
        TileWalk tw;
        for (int t = 0; t < total; ++t) {
            const int tr = band0 + tw.k * kMmaTH;
            mbar_wait(full + pos.slot, pos.phase);
            const T* s_x = reinterpret_cast<const T*>(smem + pos.slot * kXSlotBytes);
            const int row_lo = tr + 16 * wr + gq;
            const float rm0 = row_lo < band1 ? 1.f : 0.f, rm1 = row_lo + 8 < band1 ? 1.f : 0.f;
            const bool full_tile = tr + kMmaTH <= band1 && c0 >= 0 && c0 + kMmaTW <= g.W;
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) {
                float acc[4][4];
                const int tcol = 32 * wc + 8 * cb;
                branch_block<T>(s_x, wr, tcol, lane, f, acc);
                if (!full_tile) {
                    const int col = c0 + tcol + 2 * tq;                   // even; W even: the pair is all in or all out
                    const float cm = (col >= 0 && col < g.W) ? 1.f : 0.f;
                    const float m[4] = {rm0 * cm, rm0 * cm, rm1 * cm, rm1 * cm};
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[k][i] *= m[i];
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        s[k] += acc[k][i];
                        ss[k] = fmaf(acc[k][i], acc[k][i], ss[k]);
                    }
                // y_br fragments -> the warp's transpose buffer (row = block row, 16-byte rows: conflict-free both ways)
                __syncwarp();                                             // the previous block's ldmatrix reads are done
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    *reinterpret_cast<uint32_t*>(s_y + (k * 16 + gq) * 8 + 2 * tq) = MmaOp<T>::pack(acc[k][0], acc[k][1]);
                    *reinterpret_cast<uint32_t*>(s_y + (k * 16 + gq + 8) * 8 + 2 * tq) = MmaOp<T>::pack(acc[k][2], acc[k][3]);
                }
                __syncwarp();
                uint32_t Bg[4][2];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t addr = smem_u32(s_y + (k * 16 + (lane & 15)) * 8);
                    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
                                 : "=r"(Bg[k][0]), "=r"(Bg[k][1]) : "r"(addr));
                }
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
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + pos.slot);
            pos.advance();
            tw.next(ntr);
        }
        float* red = s_red + warp * 48;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float a = warp_sum(s[k]), b = warp_sum(ss[k]);
            if (lane == 0) { red[k] = a; red[4 + k] = b; }
        }
        // lag sums: entry (i, j) of a Gram fragment belongs to column lag bb = i - j (this lane: i = gq, gq + 8;
        // j = 2tq, 2tq + 1); fixed-order shuffle reduction per (accumulator, lag)
        auto diag = [&](const float (&G)[4], int bb) -> float {
            float v = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = gq + (k >> 1) * 8, j = 2 * tq + (k & 1);
                if (i - j == bb) v += G[k];
            }
            return warp_sum(v);
        };
#pragma unroll
        for (int a = 0; a < 5; ++a)
#pragma unroll
            for (int bb = 0; bb < 5; ++bb) {
                const float q = diag(G5[a], bb);
                if (lane == 0) red[8 + a * 5 + bb] = q;
            }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int bb = 1; bb < 4; ++bb) {
                const float q = diag(G3[a], bb);
                if (lane == 0) red[33 + a * 3 + (bb - 1)] = q;
            }
            const float q = diag(G31[a], 2);
            if (lane == 0) red[42 + a] = q;
        }
#pragma unroll
        for (int bb = 1; bb < 4; ++bb) {
            const float q = diag(G13, bb);
            if (lane == 0) red[45 + (bb - 1)] = q;
        }
    }
    __syncthreads();
    if (threadIdx.x < 48) {
        const int ncta = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < kDwWarps; ++w) a += s_red[w * 48 + threadIdx.x];
        part[((int64_t)e * ncta + cta) * kGramPartStride + threadIdx.x] = a;
    }
}

// ---------------------------------------------------------------------------------------------------
// backward: per-channel finalize for the composite path.  One CTA of 128 threads per channel.
//   Pfin-free: dgamma, dbeta, cb (c1, c2, c0 per branch), the weight gradients and the coefficient record of the tile
//   kernel all come out of this launch (the first generation needed dw_fin_bwd + dw_fin_dw around the dx kernel).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float dw_tap5(const lmnet_dw_params& p, int k, int e, int a, int b) {
    // branch k's kernel embedded in the 5 x 5 window, tap (a, b) in 0..4
    if (k == 0) return p.w[0][e * 25 + a * 5 + b];
    if (k == 1) return (a >= 1 && a <= 3 && b >= 1 && b <= 3) ? p.w[1][e * 9 + (a - 1) * 3 + (b - 1)] : 0.f;
    if (k == 2) return (a >= 1 && a <= 3 && b == 2) ? p.w[2][e * 3 + (a - 1)] : 0.f;
    return (a == 2 && b >= 1 && b <= 3) ? p.w[3][e * 3 + (b - 1)] : 0.f;
}

__global__ void __launch_bounds__(kFinThreads)
dw_fin_bwd2_kernel(const float* __restrict__ part, int ncta, lmnet_dw_params p, const float* __restrict__ save_mean,
                   const float* __restrict__ save_rstd, const float* __restrict__ gram /* [E][kGramFloats] */, lmnet_dw_grads gr,
                   float* __restrict__ cb /* [E][12] */, float* __restrict__ coef2 /* [E][kCoef2Stride] */,
                   float4* __restrict__ wrec /* [E][27]: embedded taps of the four branches | c2 | c0 (frame kernel) */, DwGeom g) {
    const int e = blockIdx.x, t = threadIdx.x;
    __shared__ double s_P[kReducePartS];                     // P [25] | sum du | S [25]
    __shared__ float s_w[4][25];
    __shared__ double s_c[4][3];
    for (int q = t >> 5; q < 51; q += kFinThreads / 32) {          // one warp per partial sum (see warp_sum_strided)
        const double a = warp_sum_strided(part + (int64_t)e * ncta * kReducePartS + q, ncta, kReducePartS, t & 31);
        if ((t & 31) == 0) s_P[q] = a;
    }
    if (t >= 64 && t < 89) {
        const int i = t - 64;
#pragma unroll
        for (int k = 0; k < 4; ++k) s_w[k][i] = dw_tap5(p, k, e, i / 5, i % 5);
    }
    __syncthreads();
    if (t < 4) {
        const int k = t;
        const double n = (double)g.B * g.H * g.W;
        const double sdu = s_P[25];
        double sduy = 0;
        for (int i = 0; i < 25; ++i) sduy += (double)s_w[k][i] * s_P[i];
        const double mean = save_mean[k * g.E + e], rstd = save_rstd[k * g.E + e], gamma = p.gamma[k][e];
        const double dgamma = rstd * (sduy - mean * sdu);
        if (gr.dgamma[k] != nullptr) gr.dgamma[k][e] = (float)dgamma;
        if (gr.dbeta[k] != nullptr) gr.dbeta[k][e] = (float)sdu;
        const double c1 = gamma * rstd;
        const double c2 = gamma * rstd * rstd * dgamma / n;
        const double c0 = c1 * sdu / n - c2 * mean;
        s_c[k][0] = c1; s_c[k][1] = c2; s_c[k][2] = c0;
        cb[e * 12 + k * 3 + 0] = (float)c1;
        cb[e * 12 + k * 3 + 1] = (float)c2;
        cb[e * 12 + k * 3 + 2] = (float)c0;
    }
    __syncthreads();
    // weight gradients: dw_br[t] = c1 P[lag] - c2 Q_br[t] - c0 S[lag]
    if (t < 40) {
        int k, idx, lag;
        if (t < 25) { k = 0; idx = t; lag = t; }
        else if (t < 34) { k = 1; idx = t - 25; lag = (idx / 3 + 1) * 5 + idx % 3 + 1; }
        else if (t < 37) { k = 2; idx = t - 34; lag = (idx + 1) * 5 + 2; }
        else { k = 3; idx = t - 37; lag = 2 * 5 + idx + 1; }
        const double v = s_c[k][0] * s_P[lag] - s_c[k][1] * (double)gram[e * kGramFloats + t] - s_c[k][2] * s_P[26 + lag];
        const int per = k == 0 ? 25 : k == 1 ? 9 : 3;
        if (gr.dw[k] != nullptr) gr.dw[k][e * per + idx] = (float)v;
    }
    if (t >= 96 && t < 121) {
        const int i = t - 96;
        wrec[e * 27 + i] = make_float4(s_w[0][i], s_w[1][i], s_w[2][i], s_w[3][i]);
    }
    if (t == 121) wrec[e * 27 + 25] = make_float4((float)s_c[0][1], (float)s_c[1][1], (float)s_c[2][1], (float)s_c[3][1]);
    if (t == 122) wrec[e * 27 + 26] = make_float4((float)s_c[0][2], (float)s_c[1][2], (float)s_c[2][2], (float)s_c[3][2]);
    // coefficient record of the tile kernel
    float* out = coef2 + (int64_t)e * kCoef2Stride;
    if (t < 25) {
        const int a = t / 5, b = t % 5;
        double v = 0, v0 = 0;
        for (int k = 0; k < 4; ++k) v += s_c[k][0] * (double)s_w[k][(4 - a) * 5 + (4 - b)];
        out[t] = (float)v;                                   // F5[a][b] = Wc1[4-a][4-b]
        if (t == 0) {
            for (int k = 0; k < 4; ++k)
                for (int i = 0; i < 25; ++i) v0 += s_c[k][2] * (double)s_w[k][i];
            out[106] = (float)(-v0);
            out[107] = 0.f;
        }
    }
    if (t >= 32 && t < 32 + 81) {
        // fp32 is plenty for taps that are rounded to the storage type (and FP64 runs at 1/64 rate on this part: the
        // double-precision version of this loop was most of the launch's 10 us)
        const int d = t - 32, da = d / 9 - 4, db = d % 9 - 4;
        float v = 0.f;
        for (int k = 0; k < 4; ++k) {
            float acc = 0.f;
            for (int sa = max(0, -da); sa <= min(4, 4 - da); ++sa)
#pragma unroll
                for (int sb = 0; sb < 5; ++sb) {
                    const int tb = sb + db;
                    if (tb >= 0 && tb <= 4) acc = fmaf(s_w[k][sa * 5 + sb], s_w[k][(sa + da) * 5 + tb], acc);
                }
            v = fmaf((float)s_c[k][1], acc, v);
        }
        out[25 + d] = -v;
    }
}

// ---------------------------------------------------------------------------------------------------
// backward tile kernel: dx0 = K5[F5](du) + K9[-C](x) - sum Wc0
// Tile = 32 x 64 outputs; x box 40 x 72 at (tr - 4, c0 - 4); du box 36 x 72 at (tr - 2, c0 - 4) (aligned column;
// the taps are shifted by 2 instead).  Stripe s covers output columns [64 s - 4, 64 s + 60).  Output path as in the
// apply kernel: the three fully owned 16-byte chunks of every row of a warp's 16 x 32 block (columns 4..27) leave
// through a per-warp TMA store box, the 4 + 4 boundary columns through narrow stores.
// ---------------------------------------------------------------------------------------------------
constexpr int kDx2Shift = 4;
constexpr int kDx2Stages = 3;
constexpr int kDx2XRows = kMmaTH + 8;                                    // 40
constexpr int kDx2XBytes = kDx2XRows * kMmaPitch * 2;                    // 5760 (a multiple of 128)
constexpr int kDx2SlotBytes = kDx2XBytes + kXSlotBytes;                  // x | du (36 x 72, padded to 5248)
constexpr int kDx2StagingBytes = kDwWarps * 2 * kStageBoxBytes;
constexpr size_t kDx2Smem = 128 + (size_t)kDx2Stages * kDx2SlotBytes + kDx2StagingBytes + 2 * kDx2Stages * 8;
static_assert(kDx2XBytes % 128 == 0, "TMA destinations are 128-byte aligned");

template <typename T>
__global__ void __launch_bounds__(kTmaThreads, 4)
dw_bwd_dx2_tma_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_du,
                      const __grid_constant__ CUtensorMap tm_dx, const float* __restrict__ coef2, T* __restrict__ dx, DwGeom g) {
    constexpr int S = kDx2Stages;
    extern __shared__ unsigned char dw_smem_raw[];
    unsigned char* smem = align128(dw_smem_raw);
    unsigned char* s_stage = smem + S * kDx2SlotBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(s_stage + kDx2StagingBytes);
    uint64_t* empty = full + S;
    const int e = blockIdx.z, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * kMmaTW - kDx2Shift;
    const int band0 = blockIdx.y * g.rows_per_band, band1 = min(band0 + g.rows_per_band, g.H);
    const int ntr = (band1 - band0 + kMmaTH - 1) / kMmaTH;
    const int total = band1 > band0 ? g.B * ntr : 0;
    ring_init<S>(full, empty);
    if (warp == kDwWarps) {
        if (lane == 0) {
            tma_prefetch_desc(&tm_x);
            tma_prefetch_desc(&tm_du);
            RingPos<S> pos;
            TileWalk tw;
            for (int t = 0; t < total; ++t) {
                if (t >= S) mbar_wait(empty + pos.slot, pos.phase ^ 1u);
                unsigned char* slot = smem + pos.slot * kDx2SlotBytes;
                const int tr = band0 + tw.k * kMmaTH, plane = tw.b * g.E + e;
                mbar_arrive_expect_tx(full + pos.slot, kDx2XBytes + kXTileBytes);
                tma_load_3d(slot, &tm_x, full + pos.slot, c0 - 4, tr - 4, plane);
                tma_load_3d(slot + kDx2XBytes, &tm_du, full + pos.slot, c0 - 4, tr - 2, plane);
                pos.advance();
                tw.next(ntr);
            }
        }
        return;
    }
    const int wr = warp >> 1, wc = warp & 1;
    uint32_t Fhi[5][2], Flo[5][2], C9[9][2];
    const float* cf = coef2 + (int64_t)e * kCoef2Stride;
    {
        float hi[9], lo[5];
#pragma unroll
        for (int a = 0; a < 5; ++a) {
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const float w = __ldg(cf + a * 5 + j);
                hi[j] = MmaOp<T>::round(w);
                lo[j] = w - hi[j];
            }
            toeplitz_frag<T>(hi, 5, 2, lane, Fhi[a]);
            toeplitz_frag<T>(lo, 5, 2, lane, Flo[a]);
        }
#pragma unroll
        for (int a = 0; a < 9; ++a) {
#pragma unroll
            for (int j = 0; j < 9; ++j) hi[j] = __ldg(cf + 25 + a * 9 + j);
            toeplitz_frag<T>(hi, 9, 0, lane, C9[a]);
        }
    }
    const float bias = __ldg(cf + 106);
    const int gq = lane >> 2, tq = lane & 3;
    if (lane == 0) tma_prefetch_desc(&tm_dx);
    T* stage_w = reinterpret_cast<T*>(s_stage + warp * 2 * kStageBoxBytes);
    RingPos<S> pos;
    TileWalk tw;
    for (int t = 0; t < total; ++t) {
        const int b = tw.b, tr = band0 + tw.k * kMmaTH;
        mbar_wait(full + pos.slot, pos.phase);
        const T* s_x = reinterpret_cast<const T*>(smem + pos.slot * kDx2SlotBytes);
        const T* s_du = reinterpret_cast<const T*>(smem + pos.slot * kDx2SlotBytes + kDx2XBytes);
        const int row_lo = tr + 16 * wr + gq;
        const int col_lo = c0 + 32 * wc + 2 * tq;
        const int64_t base = ((int64_t)b * g.E + e) * g.H * g.W + (int64_t)row_lo * g.W + col_lo;
        const int64_t row8 = (int64_t)8 * g.W;
        float acc[4][4];
#pragma unroll
        for (int cbk = 0; cbk < 4; ++cbk) acc[cbk][0] = acc[cbk][1] = acc[cbk][2] = acc[cbk][3] = bias;
#pragma unroll
        for (int a = 0; a < 9; ++a)
#pragma unroll
            for (int cbk = 0; cbk < 4; ++cbk) {
                uint32_t A[4];
                load_a(s_x, 16 * wr + a, 32 * wc + 8 * cbk, lane, A);
                MmaOp<T>::run(acc[cbk], A, C9[a]);
            }
#pragma unroll
        for (int a = 0; a < 5; ++a)
#pragma unroll
            for (int cbk = 0; cbk < 4; ++cbk) {
                uint32_t A[4];
                load_a(s_du, 16 * wr + a, 32 * wc + 8 * cbk, lane, A);
                MmaOp<T>::run(acc[cbk], A, Fhi[a]);
#if defined(LMNET_DX2_SPLIT_TAPS)
                MmaOp<T>::run(acc[cbk], A, Flo[a]);
#endif
            }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(empty + pos.slot);
            tma_store_wait_read<1>();           // the store issued two tiles ago no longer reads this staging buffer
        }
        __syncwarp();
        T* st = stage_w + (t & 1) * kStageRows * kStageCols;
        // columns 4..27 of the warp's block go through the staging box; columns 0..3 (block 0, tq < 2) and 28..31
        // (block 3, tq >= 2) go straight to global memory
        uint32_t edge[2];
#pragma unroll
        for (int cbk = 0; cbk < 4; ++cbk) {
            const bool staged = (cbk == 1 || cbk == 2) || (cbk == 0 && tq >= 2) || (cbk == 3 && tq < 2);
            const int scol = 8 * cbk + 2 * tq - 4;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t v = MmaOp<T>::pack(acc[cbk][2 * h], acc[cbk][2 * h + 1]);
                if (staged) *reinterpret_cast<uint32_t*>(st + (gq + 8 * h) * kStageCols + scol) = v;
                else edge[h] = v;
            }
        }
        {
            const int eoff = tq >= 2 ? 24 : 0;                               // block 3 for tq >= 2, block 0 otherwise
            const int ecol = col_lo + eoff;
            if (ecol >= 0 && ecol < g.W) {
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    if (row_lo + 8 * h < band1) *reinterpret_cast<uint32_t*>(dx + base + eoff + h * row8) = edge[h];
            }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            tma_store_3d(&tm_dx, st, c0 + 32 * wc + 4, tr + 16 * wr, b * g.E + e);
            tma_store_commit();
        }
        pos.advance();
        tw.next(ntr);
    }
    if (lane == 0) tma_store_wait_all<0>();
}

// ---------------------------------------------------------------------------------------------------
// frame patch: dx(q) += sum_br sum_{s : q-s outside} w_br[s] (c2_br y~_br(q-s) + c0_br) on the 2-pixel frame of every
// plane (2 % of the pixels at 352 x 352).  A CTA owns one segment (<= kFrameSeg pixels long) of one side of one plane:
// a rectangle R of frame pixels (top / bottom: 2 rows x segment; left / right: segment x 2 columns, between the top and
// bottom rows).  Two phases through shared memory:
//   1. v_br(p) = c2_br y~_br(p) + c0_br for every position p OUTSIDE the image within 2 pixels of R (one float4 = four
//      branches per position), ONE position per thread: the two outside lines next to R (2 x (segment + 4) = 128
//      positions), plus — only in the CTAs that touch a corner of the plane — the <= 16 positions beside R's rows;
//   2. every pixel q of R (one per thread) gathers w_br[s] v_br(q - s) over the s whose q - s lies outside the image:
//      whole window rows / columns selected by two 5-bit masks, so a non-corner pixel visits 5 or 10 taps.
// ncu on the first version (window scan, 25-tap loops per pixel): 36.6 M warp instructions per level-1 launch, issue-
// bound at 64 us; the per-channel record wrec (25 x float4 embedded branch taps | c2 | c0, written by
// dw_fin_bwd2_kernel) is read through the read-only path, no prologue barrier.
// ---------------------------------------------------------------------------------------------------
constexpr int kFrameThreads = 128;
constexpr int kFrameSeg = 60;                           // 2 x (60 + 4) outside positions = one per thread
constexpr int kFrameLocal = (kFrameSeg + 4) * 6;        // positions of the local window: (2 + 4) x (segment + 4)
constexpr int kWrecStride = 27;                         // float4 per channel

// v(p) for one outside position: `rows` = the taps that reach into the image form rows of the 5 x 5 window (p above /
// below the image), else columns (p left / right of it).  Two lines (10 predicated loads) are fetched together before
// any FMA consumes them: one memory latency per position except at corners and on tiny planes.
template <typename T>
__device__ __forceinline__ float4 dw_frame_value(const T* __restrict__ xp, const float4* __restrict__ wr, int pr, int pc, int H, int W,
                                                 bool rows, const float4& c2, const float4& c0v) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const int l0 = rows ? max(0, 2 - pr) : max(0, 2 - pc), l1 = rows ? min(4, H + 1 - pr) : min(4, W + 1 - pc);
#pragma unroll 1
    for (int l = l0; l <= l1; l += 2) {
        const bool two = l + 1 <= l1;
        const int lb = min(l + 1, 4);
        float xa[5], xb[5];
        if (rows) {
            const T* row = xp + (int64_t)(pr + l - 2) * W + pc - 2;
#pragma unroll
            for (int t = 0; t < 5; ++t) {
                const bool ok = (unsigned)(pc + t - 2) < (unsigned)W;
                xa[t] = ok ? to_f(row[t]) : 0.f;
                xb[t] = (ok && two) ? to_f(row[W + t]) : 0.f;
            }
        } else {
            const T* col = xp + (int64_t)(pr - 2) * W + pc + l - 2;
#pragma unroll
            for (int t = 0; t < 5; ++t) {
                const bool ok = (unsigned)(pr + t - 2) < (unsigned)H;
                xa[t] = ok ? to_f(col[(int64_t)t * W]) : 0.f;
                xb[t] = (ok && two) ? to_f(col[(int64_t)t * W + 1]) : 0.f;
            }
        }
#pragma unroll
        for (int t = 0; t < 5; ++t) {
            const float4 wa = __ldg(wr + (rows ? l * 5 + t : t * 5 + l)), wb = __ldg(wr + (rows ? lb * 5 + t : t * 5 + lb));
            v.x = fmaf(wa.x, xa[t], fmaf(wb.x, xb[t], v.x)); v.y = fmaf(wa.y, xa[t], fmaf(wb.y, xb[t], v.y));
            v.z = fmaf(wa.z, xa[t], fmaf(wb.z, xb[t], v.z)); v.w = fmaf(wa.w, xa[t], fmaf(wb.w, xb[t], v.w));
        }
    }
    v.x = fmaf(c2.x, v.x, c0v.x); v.y = fmaf(c2.y, v.y, c0v.y);
    v.z = fmaf(c2.z, v.z, c0v.z); v.w = fmaf(c2.w, v.w, c0v.w);
    return v;
}

template <typename T>
__global__ void __launch_bounds__(kFrameThreads, 8)
dw_bwd_frame_kernel(const T* __restrict__ x, T* __restrict__ dx, const float4* __restrict__ wrec, DwGeom g) {
    __shared__ float4 s_v[kFrameLocal];
    const int plane = blockIdx.x, e = plane % g.E;
    const int H = g.H, W = g.W;
    const int side = blockIdx.y & 3, a0 = (blockIdx.y >> 2) * kFrameSeg;
    const int nTop = min(2, H), nBot = min(2, H - nTop);
    const int nL = min(2, W), nR = min(2, W - nL);
    // rectangle of frame pixels owned by this CTA
    int r0, r1, c0, c1;
    if (side == 0) { r0 = 0; r1 = nTop; c0 = a0; c1 = min(W, a0 + kFrameSeg); }
    else if (side == 1) { r0 = H - nBot; r1 = H; c0 = a0; c1 = min(W, a0 + kFrameSeg); }
    else {
        r0 = nTop + a0; r1 = min(H - nBot, r0 + kFrameSeg);
        if (side == 2) { c0 = 0; c1 = nL; } else { c0 = W - nR; c1 = W; }
    }
    if (r1 <= r0 || c1 <= c0) return;
    const float4* wr = wrec + (int64_t)e * kWrecStride;
    const T* xp = x + (int64_t)plane * H * W;
    T* dxp = dx + (int64_t)plane * H * W;
    const int t = threadIdx.x;
    // the pixel this thread patches in phase 2: fetch its dx now
    const int rw = c1 - c0, count = (r1 - r0) * rw;       // <= 2 * kFrameSeg <= kFrameThreads
    const int qlr = t / rw, qlc = t - qlr * rw;
    const int qr = r0 + qlr, qc = c0 + qlc;
    float dx_old = 0.f;
    if (t < count) dx_old = to_f(dxp[(int64_t)qr * W + qc]);
    // ---- phase 1: outside positions of the window [r0 - 2, r1 + 2) x [c0 - 2, c1 + 2), index (lr, lc) -> lr * lw + lc
    const float4 c2 = __ldg(wr + 25), c0v = __ldg(wr + 26);
    const int lw = c1 - c0 + 4, lh = r1 - r0 + 4;
    if (side < 2) {
        // rows above (top side) or below (bottom side) the image: window rows 0, 1 resp. lh - 2, lh - 1
        const int k = t / lw, lc = t - k * lw;
        if (k < 2) {
            const int lr = side == 0 ? k : lh - 2 + k;
            s_v[lr * lw + lc] = dw_frame_value(xp, wr, r0 - 2 + lr, c0 - 2 + lc, H, W, true, c2, c0v);
        }
        // on planes shorter than the window the other vertical side is within reach as well
        if (side == 0 ? r1 + 2 > H : r0 - 2 < 0) {
            for (int i = t; i < lw * lh; i += kFrameThreads) {
                const int lr = i / lw, pr = r0 - 2 + lr;
                if (side == 0 ? pr >= H : pr < 0) s_v[i] = dw_frame_value(xp, wr, pr, c0 - 2 + (i - lr * lw), H, W, true, c2, c0v);
            }
        }
        // positions beside the rows of the window that lie inside the image: only where R touches a corner of the plane
        if (c0 < 2 || c1 + 2 > W) {
            for (int i = t; i < lh * 4; i += kFrameThreads) {
                const int lr = i >> 2, j = i & 3;
                const int lc = j < 2 ? j : lw - 4 + j;                      // window columns 0, 1, lw - 2, lw - 1
                const int pr = r0 - 2 + lr, pc = c0 - 2 + lc;
                if ((unsigned)pr < (unsigned)H && (unsigned)pc >= (unsigned)W)
                    s_v[lr * lw + lc] = dw_frame_value(xp, wr, pr, pc, H, W, false, c2, c0v);
            }
        }
    } else {
        // columns left (side 2) or right (side 3) of the image: window columns 0, 1 resp. lw - 2, lw - 1; every window
        // row lies inside the image (R excludes the top and bottom frame rows)
        const int k = t / lh, lr = t - k * lh;
        if (k < 2) {
            const int lc = side == 2 ? k : lw - 2 + k;
            s_v[lr * lw + lc] = dw_frame_value(xp, wr, r0 - 2 + lr, c0 - 2 + lc, H, W, false, c2, c0v);
        }
        // planes narrower than the window: the other horizontal side is within reach as well
        if (side == 2 ? c1 + 2 > W : c0 - 2 < 0) {
            for (int i = t; i < lw * lh; i += kFrameThreads) {
                const int lr2 = i / lw, lc2 = i - lr2 * lw, pc = c0 - 2 + lc2;
                if (side == 2 ? pc >= W : pc < 0) s_v[i] = dw_frame_value(xp, wr, r0 - 2 + lr2, pc, H, W, false, c2, c0v);
            }
        }
    }
    __syncthreads();
    // ---- phase 2: one pixel of R per thread
    if (t >= count) return;
    uint32_t rmask = 0u, cmask = 0u;                      // bit sa / sb: window row / column of q - s outside the image
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        rmask |= ((unsigned)(qr - (k - 2)) >= (unsigned)H ? 1u : 0u) << k;
        cmask |= ((unsigned)(qc - (k - 2)) >= (unsigned)W ? 1u : 0u) << k;
    }
    float acc = 0.f;
#pragma unroll
    for (int sa = 0; sa < 5; ++sa) {
        const float4* vrow = s_v + (qlr + 4 - sa) * lw + qlc + 4;
        if ((rmask >> sa) & 1u) {
#pragma unroll
            for (int sb = 0; sb < 5; ++sb) {
                const float4 v = vrow[-sb], w = __ldg(wr + sa * 5 + sb);
                acc = fmaf(w.x, v.x, fmaf(w.y, v.y, fmaf(w.z, v.z, fmaf(w.w, v.w, acc))));
            }
        } else if (cmask != 0u) {
#pragma unroll
            for (int sb = 0; sb < 5; ++sb)
                if ((cmask >> sb) & 1u) {
                    const float4 v = vrow[-sb], w = __ldg(wr + sa * 5 + sb);
                    acc = fmaf(w.x, v.x, fmaf(w.y, v.y, fmaf(w.z, v.z, fmaf(w.w, v.w, acc))));
                }
        }
    }
    dxp[(int64_t)qr * W + qc] = from_f<T>(dx_old + acc);
}

}  // namespace lmnet
