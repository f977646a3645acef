// reparam_dw_tma.cuh — the depthwise branch kernels of reparam_dw_mma.cuh rebuilt on a TMA + mbarrier tile pipeline.
//
// Same arithmetic as the first-generation tensor-core kernels (banded-Toeplitz / Gram products on mma.sync, fp32
// accumulate, identical operand rounding); what changes is how tiles reach the SM and how warps synchronise:
//
//   * a PRODUCER warp (warp 4; one elected lane) walks the CTA's tile list and issues one `cp.async.bulk.tensor`
//     (3-D tensor map over [B*E, H, W], box = tile + halo) per operand and tile into a ring of kStages shared-memory
//     slots.  Out-of-image elements — the stencil halo, ragged right / bottom tiles, negative coordinates — arrive
//     as zeros from the copy engine, so no thread computes an address or tests a bound for loads;
//   * each slot has a FULL mbarrier (armed with the tile's byte count, completed by the TMA unit) and an EMPTY
//     mbarrier (one arrival per compute warp).  The four compute warps never meet at a CTA-wide barrier inside the
//     tile loop: a warp waits for the slot, computes its 16 x 32 block, releases the slot and moves on, so up to
//     kStages - 1 tiles are in flight per CTA while the tensor pipe works (the first generation had one tile in
//     registers and two __syncthreads per 4 KB tile);
//   * the backward dx kernel also forms the depthwise weight gradients (Gram products of the dy_br tiles it already
//     holds in shared memory with the x tile) — the separate `dw_bwd_dw` pass over x and the second recomputation
//     of y_br are gone, and dw_br[t] = sum_p dy_br(p) x(p+t) needs no correction term;
//   * per-CTA partial sums are reduced in a fixed order (per-warp fragments -> shared memory -> one thread per bin):
//     no shared-memory atomics, results are bit-reproducible.
//
// Requirements (checked on the host, otherwise the first-generation kernels run): 16-bit storage, W % 8 == 0
// (TMA global strides are multiples of 16 bytes), 16-byte aligned base pointers.
//
// Measured TMA rule that shapes the tiling (tools/debug/tma_probe.cu on B200): the INNERMOST box coordinate must be a
// multiple of 16 bytes (8 elements) — negative is fine (-8, -16), but -2 or +6 raise "illegal instruction" — while
// outer coordinates are free.  The stencil wants its x tile to start 2 (forward) or 4 (dx) columns left of the output
// stripe with ldmatrix reading at 8-column steps from that origin, so the STRIPE GRID is shifted instead: stripe s of
// the forward / reduce kernels covers output columns [64 s - 6, 64 s + 58) and its x box starts at column 64 s - 8;
// stripe s of the dx kernel covers [56 s - 4, 56 s + 52) with the x box at 56 s - 8.  Boxes whose natural origin is
// the stripe origin itself (u, dz) or 2 left of it (du in the dx pass) are fetched from the aligned column below and
// read at a +2 column offset in shared memory.
#pragma once
#include "tma.cuh"

namespace lmnet {

constexpr int kFwdShift = 6;            // forward / reduce stripes start at 64 s - 6  (x box at c0 - 2 = 64 s - 8)
constexpr int kDxShift = 4;             // dx stripes start at 56 s - 4                 (x box at c0 - 4 = 56 s - 8)
constexpr int kTmaThreads = kDwThreads + 32;                          // 4 compute warps + the producer warp
constexpr int kXTileBytes = kMmaTileRows * kMmaPitch * 2;             // 36 x 72 x 2 = 5184
constexpr int kXSlotBytes = (kXTileBytes + 127) / 128 * 128;          // TMA destinations are 128-byte aligned
constexpr int kRTileBytes = kMmaTH * kMmaPitch * 2;                   // 32 x 72 x 2 = 4608 (u / dz / du boxes)

__device__ __forceinline__ unsigned char* align128(unsigned char* p) {
    return reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(p) + 127) & ~(uintptr_t)127);
}

// ring bookkeeping shared by producer and consumers: slot index and phase parity advance together
template <int S> struct RingPos {
    int slot = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance() {
        if (++slot == S) { slot = 0; phase ^= 1u; }
    }
};

template <int S>
__device__ __forceinline__ void ring_init(uint64_t* full, uint64_t* empty) {
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, kDwWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();
}

// tile list of one CTA: all batch images x the row tiles of its band
struct TileWalk {
    int b = 0, k = 0;
    __device__ __forceinline__ void next(int ntr) {
        if (++k == ntr) { k = 0; ++b; }
    }
};

struct BranchFrags {
    uint32_t B5[5][2], B3[3][2], B31[3][2], B13[2];
};
template <typename T>
__device__ __forceinline__ void load_branch_frags(const lmnet_dw_params& p, int e, int lane, BranchFrags& f, BranchFrags* flipped) {
    float w5[25], w3[9], w31[3], w13[3], t[5];
#pragma unroll
    for (int i = 0; i < 25; ++i) w5[i] = __ldg(p.w[0] + e * 25 + i);
#pragma unroll
    for (int i = 0; i < 9; ++i) w3[i] = __ldg(p.w[1] + e * 9 + i);
#pragma unroll
    for (int i = 0; i < 3; ++i) { w31[i] = __ldg(p.w[2] + e * 3 + i); w13[i] = __ldg(p.w[3] + e * 3 + i); }
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        toeplitz_frag<T>(w5 + a * 5, 5, 0, lane, f.B5[a]);
        if (flipped != nullptr) {
#pragma unroll
            for (int j = 0; j < 5; ++j) t[j] = w5[(4 - a) * 5 + (4 - j)];
            toeplitz_frag<T>(t, 5, 0, lane, flipped->B5[a]);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        toeplitz_frag<T>(w3 + a * 3, 3, 1, lane, f.B3[a]);
        toeplitz_frag<T>(w31 + a, 1, 2, lane, f.B31[a]);
        if (flipped != nullptr) {
#pragma unroll
            for (int j = 0; j < 3; ++j) t[j] = w3[(2 - a) * 3 + (2 - j)];
            toeplitz_frag<T>(t, 3, 1, lane, flipped->B3[a]);
            t[0] = w31[2 - a];
            toeplitz_frag<T>(t, 1, 2, lane, flipped->B31[a]);
        }
    }
    toeplitz_frag<T>(w13, 3, 1, lane, f.B13);
    if (flipped != nullptr) {
#pragma unroll
        for (int j = 0; j < 3; ++j) t[j] = w13[2 - j];
        toeplitz_frag<T>(t, 3, 1, lane, flipped->B13);
    }
}

// the four branch outputs of one 16 x 8 block (rows 16*wr.., tile columns tcol..) from the x tile
template <typename T>
__device__ __forceinline__ void branch_block(const T* s_x, int wr, int tcol, int lane, const BranchFrags& f, float (&acc)[4][4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f;
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        uint32_t A[4];
        load_a(s_x, 16 * wr + a, tcol, lane, A);
        MmaOp<T>::run(acc[0], A, f.B5[a]);
        if (a >= 1 && a <= 3) {
            MmaOp<T>::run(acc[1], A, f.B3[a - 1]);
            MmaOp<T>::run(acc[2], A, f.B31[a - 1]);
        }
        if (a == 2) MmaOp<T>::run(acc[3], A, f.B13);
    }
}

// ---------------------------------------------------------------------------------------------------
// statistics pass
// ---------------------------------------------------------------------------------------------------
constexpr int kStatsStages = 4;
constexpr size_t kStatsSmem = 128 + (size_t)kStatsStages * kXSlotBytes + 2 * kStatsStages * 8 + kDwWarps * 8 * 4;

template <typename T>
__global__ void __launch_bounds__(kTmaThreads, 4)
dw_stats_tma_kernel(const __grid_constant__ CUtensorMap tm_x, lmnet_dw_params p, float* __restrict__ part /* [E][ncta][8] */,
                    DwGeom g) {
    constexpr int S = kStatsStages;
    extern __shared__ unsigned char dw_smem_raw[];
    unsigned char* smem = align128(dw_smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S * kXSlotBytes);
    uint64_t* empty = full + S;
    float* s_red = reinterpret_cast<float*>(empty + S);
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
        const int gq = lane >> 2, tq = lane & 3;
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
                if (full_tile) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            s[k] += acc[k][i];
                            ss[k] = fmaf(acc[k][i], acc[k][i], ss[k]);
                        }
                } else {
                    const int col = c0 + tcol + 2 * tq;                   // even; W even: the pair is all in or all out
                    const float cm0 = (col >= 0 && col < g.W) ? 1.f : 0.f, cm1 = cm0;
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
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + pos.slot);
            pos.advance();
            tw.next(ntr);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float a = warp_sum(s[k]), b = warp_sum(ss[k]);
            if (lane == 0) { s_red[warp * 8 + k] = a; s_red[warp * 8 + 4 + k] = b; }
        }
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        const int ncta = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < kDwWarps; ++w) a += s_red[w * 8 + threadIdx.x];
        part[((int64_t)e * ncta + cta) * 8 + threadIdx.x] = a;
    }
}

// ---------------------------------------------------------------------------------------------------
// apply pass: u = merged5x5(x) + bias, z = GELU(u), per-warp pool partial sums.
//
// Output path (ablation on B200: with fragment stores straight to global memory — 4 bytes per lane, 8 rows x 16 B per
// instruction — each of the two output streams cost ~90 us of a 196 us kernel; the LSU store path, not HBM, was the
// limit).  A warp's 16 x 32 output block starts 2 columns past a 16-byte boundary (the shifted stripe grid), so each
// warp stages the three fully owned 16-byte chunks of every row (24 of its 32 columns) in shared memory and ONE lane
// hands the 16 x 24 box to the TMA unit (cp.async.bulk.tensor store: full-line writes, no LSU work, clipped at the
// image edge by the tensor map); the 6 + 2 boundary columns shared with the neighbouring warps / stripes keep the
// narrow stores.  The staging boxes are double-buffered per warp; `cp.async.bulk.wait_group.read` guards their reuse.
// ---------------------------------------------------------------------------------------------------
constexpr int kApplyStages = 4;
constexpr int kStageCols = 24, kStageRows = 16;                                    // per-warp TMA-store box
constexpr int kStageBoxBytes = kStageRows * kStageCols * 2;                        // 768 (a multiple of 128)
constexpr int kApplyStagingBytes = kDwWarps * 2 /*buffers*/ * 2 /*u, z*/ * kStageBoxBytes;
constexpr size_t kApplySmem = 128 + (size_t)kApplyStages * kXSlotBytes + kApplyStagingBytes + 2 * kApplyStages * 8;

template <typename T>
__global__ void __launch_bounds__(kTmaThreads, 4)
dw_apply_tma_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_u,
                    const __grid_constant__ CUtensorMap tm_z, const float* __restrict__ coef, T* __restrict__ u_out,
                    T* __restrict__ z_out, float* __restrict__ pool_part /* [B*E][ncta*4] or null */, DwGeom g) {
    constexpr int S = kApplyStages;
    extern __shared__ unsigned char dw_smem_raw[];
    unsigned char* smem = align128(dw_smem_raw);
    unsigned char* s_stage = smem + S * kXSlotBytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(s_stage + kApplyStagingBytes);
    uint64_t* empty = full + S;
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
        return;
    }
    const int wr = warp >> 1, wc = warp & 1;
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
    const bool want_u = u_out != nullptr;
    if (lane == 0) {
        tma_prefetch_desc(&tm_z);
        if (want_u) tma_prefetch_desc(&tm_u);
    }
    // this warp's staging boxes: [buffer][u | z][16][24]
    T* stage_w = reinterpret_cast<T*>(s_stage + warp * 4 * kStageBoxBytes);
    float psum = 0.f;
    RingPos<S> pos;
    TileWalk tw;
    for (int t = 0; t < total; ++t) {
        const int b = tw.b, tr = band0 + tw.k * kMmaTH;
        const bool last = tw.k == ntr - 1;
        mbar_wait(full + pos.slot, pos.phase);
        const T* s_x = reinterpret_cast<const T*>(smem + pos.slot * kXSlotBytes);
        const int row_lo = tr + 16 * wr + gq;
        const int col_lo = c0 + 32 * wc + 2 * tq;
        const int64_t base = ((int64_t)b * g.E + e) * g.H * g.W + (int64_t)row_lo * g.W + col_lo;
        const int64_t row8 = (int64_t)8 * g.W;
        float acc[4][4];
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) acc[cb][0] = acc[cb][1] = acc[cb][2] = acc[cb][3] = bias;
        // kernel-row outer, column-block inner: four independent accumulator chains per warp
#pragma unroll
        for (int a = 0; a < 5; ++a)
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) {
                uint32_t A[4];
                load_a(s_x, 16 * wr + a, 32 * wc + 8 * cb, lane, A);
                MmaOp<T>::run(acc[cb], A, Bhi[a]);
                MmaOp<T>::run(acc[cb], A, Blo[a]);
            }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(empty + pos.slot);      // the tile is consumed: the epilogue overlaps the next copy
            tma_store_wait_read<1>();           // the store issued two tiles ago no longer reads this staging buffer
        }
        __syncwarp();
        T* st_u = stage_w + (t & 1) * 2 * kStageRows * kStageCols;
        T* st_z = st_u + kStageRows * kStageCols;
        // columns 6..29 of the warp's block go through the staging box; the boundary columns 0..5 (block 0, tq < 3) and
        // 30..31 (block 3, tq == 3) go straight to global memory, merged into ONE store per row half and tensor
        uint32_t edge_u[2], edge_z[2];
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
            const bool staged = (cb == 1 || cb == 2) || (cb == 0 && tq == 3) || (cb == 3 && tq < 3);
            const int scol = 8 * cb + 2 * tq - 6;
            const bool cok = col_lo + 8 * cb >= 0 && col_lo + 8 * cb < g.W;     // W even: pair in or out together
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t up = MmaOp<T>::pack(acc[cb][2 * h], acc[cb][2 * h + 1]);
                const T* ur = reinterpret_cast<const T*>(&up);              // GELU of the value as stored
                float z0, z1;
                gelu_fast2(to_f(ur[0]), to_f(ur[1]), z0, z1);
                const uint32_t zp = MmaOp<T>::pack(z0, z1);
                if (staged) {
                    const int so = (gq + 8 * h) * kStageCols + scol;
                    *reinterpret_cast<uint32_t*>(st_u + so) = up;
                    *reinterpret_cast<uint32_t*>(st_z + so) = zp;
                } else {
                    edge_u[h] = up;
                    edge_z[h] = zp;
                }
                if (cok && row_lo + 8 * h < band1) {
                    const T* zr = reinterpret_cast<const T*>(&zp);          // pool of the values as stored
                    psum += to_f(zr[0]) + to_f(zr[1]);
                }
            }
        }
        {
            const int ecol = col_lo + (tq == 3 ? 24 : 0);                    // block 3 for tq == 3, block 0 otherwise
            if (ecol >= 0 && ecol < g.W) {
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    if (row_lo + 8 * h < band1) {
                        const int64_t off = base + (tq == 3 ? 24 : 0) + h * row8;
                        if (want_u) *reinterpret_cast<uint32_t*>(u_out + off) = edge_u[h];
                        *reinterpret_cast<uint32_t*>(z_out + off) = edge_z[h];
                    }
            }
        }
        fence_proxy_async();                    // generic-proxy writes -> visible to the TMA store
        __syncwarp();
        if (lane == 0) {
            // rows beyond the image and columns beyond W are clipped by the tensor map; bands are multiples of the tile
            // height, so rows >= band1 only occur at the bottom of the image
            const int gc = c0 + 32 * wc + 6, gr = tr + 16 * wr, plane = b * g.E + e;
            if (want_u) tma_store_3d(&tm_u, st_u, gc, gr, plane);
            tma_store_3d(&tm_z, st_z, gc, gr, plane);
            tma_store_commit();
        }
        if (last && pool_part != nullptr) {
            const float tsum = warp_sum(psum);
            psum = 0.f;
            if (lane == 0) pool_part[(((int64_t)b * g.E + e) * ncta + cta) * kDwWarps + warp] = tsum;
        }
        pos.advance();
        tw.next(ntr);
    }
    if (lane == 0) tma_store_wait_all<0>();     // shared memory must stay alive until the last stores have read it
}

// ---------------------------------------------------------------------------------------------------
// backward reduce pass: du = (dz + dpool/HW) * gelu'(u) (written out), sum du, and the 25 lag sums
// P[a][b] = sum_p du(p) x(p + (a-2, b-2)) as Gram products.  Each warp forms du for ITS 16 x 32 block and consumes
// only that block, so the du tile needs no cross-warp synchronisation.
// ---------------------------------------------------------------------------------------------------
constexpr int kReduceStages = 3;
constexpr int kReduceSlotBytes = kXSlotBytes + 2 * kRTileBytes;          // x | u | dz
constexpr size_t kReduceSmem = 128 + (size_t)kReduceStages * kReduceSlotBytes + kRTileBytes /* du */ +
                               kDwWarps * 5 * 128 * 4 /* G fragments */ + 2 * kReduceStages * 8 + 64 + kDwWarps * 25 * 4;
constexpr int kReducePartS = 52;       // partial record with the x lag sums: P [25] | sum du | S [25] | pad

template <typename T> __device__ __forceinline__ uint32_t ones_pair();
template <> __device__ __forceinline__ uint32_t ones_pair<__nv_bfloat16>() { return 0x3F803F80u; }
template <> __device__ __forceinline__ uint32_t ones_pair<__half>() { return 0x3C003C00u; }

// WITH_S: additionally S[a][b] = sum_p x(p + (a-2, b-2)) over the pixels p of the image (an all-ones Gram product on
// the A fragments the P products load anyway) for the composite backward of reparam_dw_tma2.cuh; record stride
// kReducePartS instead of 26.
template <typename T, bool WITH_S>
__global__ void __launch_bounds__(kTmaThreads, 3)
dw_bwd_reduce_tma_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_u,
                         const __grid_constant__ CUtensorMap tm_dz, const float* __restrict__ dpool, T* __restrict__ du_out,
                         float* __restrict__ part /* [E][ncta][26 | kReducePartS] */, DwGeom g) {
    constexpr int S = kReduceStages;
    extern __shared__ unsigned char dw_smem_raw[];
    unsigned char* smem = align128(dw_smem_raw);
    T* s_du = reinterpret_cast<T*>(smem + S * kReduceSlotBytes);
    float* s_G = reinterpret_cast<float*>(smem + S * kReduceSlotBytes + kRTileBytes);       // [4 warps][5][128]
    uint64_t* full = reinterpret_cast<uint64_t*>(s_G + kDwWarps * 5 * 128);
    uint64_t* empty = full + S;
    float* s_sdu = reinterpret_cast<float*>(empty + S);                                      // [4]
    float* s_S = s_sdu + 16;                                                                 // [4 warps][25]
    constexpr int STRIDE = WITH_S ? kReducePartS : 26;
    const int e = blockIdx.z, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * kMmaTW - kFwdShift;
    const int band0 = blockIdx.y * g.rows_per_band, band1 = min(band0 + g.rows_per_band, g.H);
    const int ntr = (band1 - band0 + kMmaTH - 1) / kMmaTH;
    const int total = band1 > band0 ? g.B * ntr : 0;
    ring_init<S>(full, empty);
    if (warp == kDwWarps) {
        if (lane == 0) {
            tma_prefetch_desc(&tm_x);
            tma_prefetch_desc(&tm_u);
            tma_prefetch_desc(&tm_dz);
            RingPos<S> pos;
            TileWalk tw;
            for (int t = 0; t < total; ++t) {
                if (t >= S) mbar_wait(empty + pos.slot, pos.phase ^ 1u);
                unsigned char* slot = smem + pos.slot * kReduceSlotBytes;
                const int tr = band0 + tw.k * kMmaTH, plane = tw.b * g.E + e;
                mbar_arrive_expect_tx(full + pos.slot, kXTileBytes + 2 * kRTileBytes);
                tma_load_3d(slot, &tm_x, full + pos.slot, c0 - 2, tr - 2, plane);
                tma_load_3d(slot + kXSlotBytes, &tm_u, full + pos.slot, c0 - 2, tr, plane);       // aligned column; read at +2
                tma_load_3d(slot + kXSlotBytes + kRTileBytes, &tm_dz, full + pos.slot, c0 - 2, tr, plane);
                pos.advance();
                tw.next(ntr);
            }
        }
    } else {
        const int wr = warp >> 1, wc = warp & 1;
        const float inv_hw = 1.f / ((float)g.H * (float)g.W);
        float G[5][4], G1[WITH_S ? 5 : 1][4];
#pragma unroll
        for (int a = 0; a < 5; ++a) G[a][0] = G[a][1] = G[a][2] = G[a][3] = 0.f;
#pragma unroll
        for (int a = 0; a < (WITH_S ? 5 : 1); ++a) G1[a][0] = G1[a][1] = G1[a][2] = G1[a][3] = 0.f;
        float sdu = 0.f;
        // du mapping inside the warp's block: half-warps take rows 4 apart (conflict-free at the 72-element pitch),
        // 16 lanes x one 32-bit pair = 64 contiguous bytes of one image row
        const int half = lane >> 4, cp = lane & 15;
        const int bcol = 32 * wc + 2 * cp;                       // tile column of this lane's pair
        const bool cok = c0 + bcol >= 0 && c0 + bcol < g.W;      // W even: the pair is all in or all out
        RingPos<S> pos;
        TileWalk tw;
        for (int t = 0; t < total; ++t) {
            const int b = tw.b, tr = band0 + tw.k * kMmaTH;
            const int64_t poff = ((int64_t)b * g.E + e) * g.H * g.W;
            const float dp = dpool != nullptr ? __ldg(dpool + b * g.E + e) * inv_hw : 0.f;
            mbar_wait(full + pos.slot, pos.phase);
            const unsigned char* slot = smem + pos.slot * kReduceSlotBytes;
            const T* s_x = reinterpret_cast<const T*>(slot);
            const T* s_u = reinterpret_cast<const T*>(slot + kXSlotBytes);
            const T* s_dz = reinterpret_cast<const T*>(slot + kXSlotBytes + kRTileBytes);
            const bool rows_full = tr + kMmaTH <= band1;
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                const int trow = 16 * wr + 4 * half + (o & 3) + 8 * (o >> 2);
                uint32_t packed = 0u;
                if (cok && (rows_full || tr + trow < band1)) {
                    const uint32_t ur = *reinterpret_cast<const uint32_t*>(s_u + trow * kMmaPitch + bcol + 2);
                    const uint32_t zr = *reinterpret_cast<const uint32_t*>(s_dz + trow * kMmaPitch + bcol + 2);
                    const T* ue = reinterpret_cast<const T*>(&ur);
                    const T* ze = reinterpret_cast<const T*>(&zr);
                    float g0, g1;
                    gelu_grad_fast2(to_f(ue[0]), to_f(ue[1]), g0, g1);
                    packed = MmaOp<T>::pack((to_f(ze[0]) + dp) * g0, (to_f(ze[1]) + dp) * g1);
#if !defined(LMNET_DBG_NO_DU)
                    *reinterpret_cast<uint32_t*>(du_out + poff + (int64_t)(tr + trow) * g.W + c0 + bcol) = packed;
#endif
                    const T* de = reinterpret_cast<const T*>(&packed);
                    sdu += to_f(de[0]) + to_f(de[1]);             // exactly what the dx pass reads back
                }
                *reinterpret_cast<uint32_t*>(s_du + trow * kMmaPitch + bcol) = packed;
            }
            __syncwarp();
            // ones operand in B-fragment layout: block rows 2tq, 2tq+1 (+8) at column gq, masked to the image / band
            uint32_t onesr[2] = {0u, 0u};
            if constexpr (WITH_S) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int r = tr + 16 * wr + 2 * (lane & 3) + 8 * h;
                    onesr[h] = ones_pair<T>() & ((r < band1 ? 0x0000ffffu : 0u) | (r + 1 < band1 ? 0xffff0000u : 0u));
                }
            }
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) {
                const int tcol = 32 * wc + 8 * cb;
                uint32_t Bf[2], B1[2] = {0u, 0u};
                load_b_trans(s_du, 16 * wr, tcol, lane, Bf);
                if constexpr (WITH_S) {
                    const int col = c0 + tcol + (lane >> 2);
                    const uint32_t cmask = (col >= 0 && col < g.W) ? 0xffffffffu : 0u;
                    B1[0] = onesr[0] & cmask;
                    B1[1] = onesr[1] & cmask;
                }
#pragma unroll
                for (int a = 0; a < 5; ++a) {
                    uint32_t A[4];
                    load_a_trans(s_x, 16 * wr + a, tcol, lane, A);
                    MmaOp<T>::run(G[a], A, Bf);
                    if constexpr (WITH_S) MmaOp<T>::run(G1[a], A, B1);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + pos.slot);
            pos.advance();
            tw.next(ntr);
        }
        // fragments -> shared memory; P[a][bb] = sum over (i, j) with i - j == bb, summed in a fixed order below
#pragma unroll
        for (int a = 0; a < 5; ++a)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = (lane >> 2) + (k >> 1) * 8, j = 2 * (lane & 3) + (k & 1);
                s_G[(warp * 5 + a) * 128 + i * 8 + j] = G[a][k];
            }
        sdu = warp_sum(sdu);
        if (lane == 0) s_sdu[warp] = sdu;
        if constexpr (WITH_S) {
            // entry (i, j) of a fragment belongs to column lag bb = i - j (this lane: i = gq, gq + 8; j = 2tq, 2tq + 1)
            const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
            for (int a = 0; a < 5; ++a)
#pragma unroll
                for (int bb = 0; bb < 5; ++bb) {
                    float v = 0.f;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (gq + (k >> 1) * 8 - (2 * tq + (k & 1)) == bb) v += G1[a][k];
                    v = warp_sum(v);
                    if (lane == 0) s_S[warp * 25 + a * 5 + bb] = v;
                }
        }
    }
    __syncthreads();
    if (WITH_S && threadIdx.x >= 32 && threadIdx.x < 57) {
        const int ncta = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x, t = threadIdx.x - 32;
        part[((int64_t)e * ncta + cta) * STRIDE + 26 + t] = (s_S[t] + s_S[25 + t]) + (s_S[50 + t] + s_S[75 + t]);
    }
    if (threadIdx.x < 26) {
        const int ncta = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
        float acc = 0.f;
        if (threadIdx.x == 25) {
            acc = (s_sdu[0] + s_sdu[1]) + (s_sdu[2] + s_sdu[3]);
        } else {
            const int a = threadIdx.x / 5, bb = threadIdx.x % 5;
            for (int w = 0; w < kDwWarps; ++w)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc += s_G[(w * 5 + a) * 128 + (j + bb) * 8 + j];
        }
        part[((int64_t)e * ncta + cta) * STRIDE + threadIdx.x] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------
// backward dx + dw pass.  Per CTA tile: dx block of 28 x 56 pixels; dy region 32 x 64 with origin (tr-2, c0-2);
// x tile with origin (tr-4, c0-4).
//   phase 1 (per warp, 16 x 32 block of the region): y_br by Toeplitz MMAs -> dy_br = c1*du - c2*y_br - c0 (storage
//            type, zero outside the image) into the shared dy tiles;
//   phase G (same warp, same block, only __syncwarp): weight gradients dw_br[t] += sum_p dy_br(p) x(p+t) as Gram
//            products over the block's INTERIOR pixels (the 2-pixel halo of the region belongs to neighbouring tiles);
//   named barrier over the 128 compute threads (the dy tiles are double-buffered: one barrier per tile);
//   phase 2: dx by Toeplitz MMAs with the flipped taps over the four dy tiles.
// ---------------------------------------------------------------------------------------------------
#ifndef LMNET_DX_OCC
#define LMNET_DX_OCC 2
#endif
constexpr int kDxStages = 3;
constexpr int kDxSlotBytes = kXSlotBytes + kRTileBytes;                 // x | du
constexpr int kDyTileElems = kMmaTileRows * kMmaPitch;                  // one branch: 36 x 72
constexpr int kDySetBytes = 4 * kDyTileElems * 2;                       // four branches: 20736
constexpr size_t kDxTmaSmem = 128 + (size_t)kDxStages * kDxSlotBytes + 2 * kDySetBytes + 2 * kDxStages * 8;
static_assert(2 * kDySetBytes >= kDwWarps * 12 * 128 * 4, "the Gram fragments are staged in the dy tiles after the loop");

template <typename T>
__global__ void __launch_bounds__(kTmaThreads, LMNET_DX_OCC)
dw_bwd_dx_tma_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_du, lmnet_dw_params p,
                     const float* __restrict__ cb, T* __restrict__ dx, float* __restrict__ part /* [E][ncta][40] */, DwGeom g) {
    constexpr int S = kDxStages;
    extern __shared__ unsigned char dw_smem_raw[];
    unsigned char* smem = align128(dw_smem_raw);
    T* s_dy_all = reinterpret_cast<T*>(smem + S * kDxSlotBytes);                            // [2][4][36][72]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S * kDxSlotBytes + 2 * kDySetBytes);
    uint64_t* empty = full + S;
    const int e = blockIdx.z, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * kDxTW - kDxShift;
    const int band0 = blockIdx.y * g.rows_per_band, band1 = min(band0 + g.rows_per_band, g.H);
    const int ntr = (band1 - band0 + kDxTH - 1) / kDxTH;
    const int total = band1 > band0 ? g.B * ntr : 0;
    // rows 32..35 of the dy tiles are read (for discarded outputs only) but never written: keep them finite
    for (int i = threadIdx.x; i < 2 * 4 * 4 * kMmaPitch; i += kTmaThreads) {
        const int k = i / (4 * kMmaPitch), r = i - k * 4 * kMmaPitch;
        s_dy_all[k * kDyTileElems + kMmaTH * kMmaPitch + r] = from_f<T>(0.f);
    }
    ring_init<S>(full, empty);
    if (warp == kDwWarps) {
        if (lane == 0) {
            tma_prefetch_desc(&tm_x);
            tma_prefetch_desc(&tm_du);
            RingPos<S> pos;
            TileWalk tw;
            for (int t = 0; t < total; ++t) {
                if (t >= S) mbar_wait(empty + pos.slot, pos.phase ^ 1u);
                unsigned char* slot = smem + pos.slot * kDxSlotBytes;
                const int tr = band0 + tw.k * kDxTH, plane = tw.b * g.E + e;
                mbar_arrive_expect_tx(full + pos.slot, kXTileBytes + kRTileBytes);
                tma_load_3d(slot, &tm_x, full + pos.slot, c0 - 4, tr - 4, plane);
                tma_load_3d(slot + kXSlotBytes, &tm_du, full + pos.slot, c0 - 4, tr - 2, plane);    // aligned column; read at +2
                pos.advance();
                tw.next(ntr);
            }
        }
    } else {
        const int wr = warp >> 1, wc = warp & 1;
        BranchFrags f, fl;
        load_branch_frags<T>(p, e, lane, f, &fl);
        float c1[4], c2[4], c0c[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            c1[k] = __ldg(cb + e * 12 + k * 3);
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
        RingPos<S> pos;
        TileWalk tw;
        for (int t = 0; t < total; ++t) {
            const int b = tw.b, tr = band0 + tw.k * kDxTH;
            const int64_t poff = ((int64_t)b * g.E + e) * g.H * g.W;
            T* s_dy = s_dy_all + (t & 1) * 4 * kDyTileElems;
            mbar_wait(full + pos.slot, pos.phase);
            const unsigned char* slot = smem + pos.slot * kDxSlotBytes;
            const T* s_x = reinterpret_cast<const T*>(slot);
            const T* s_du = reinterpret_cast<const T*>(slot + kXSlotBytes);
            // ---- phase 1: dy_br on this warp's 16 x 32 block of the region
            {
                const int rho0 = 16 * wr + gq;                         // region rows rho0, rho0 + 8
                const int row0 = tr - 2 + rho0, row1 = row0 + 8;
                const bool rin0 = row0 >= 0 && row0 < g.H, rin1 = row1 >= 0 && row1 < g.H;
                const bool inside = tr - 2 >= 0 && tr - 2 + kMmaTH <= g.H && c0 - 2 >= 0 && c0 - 2 + kMmaTW <= g.W;
#pragma unroll
                for (int cbk = 0; cbk < 4; ++cbk) {
                    float acc[4][4];
                    const int tcol = 32 * wc + 8 * cbk;                // region column of the block = x-tile index
                    branch_block<T>(s_x, wr, tcol, lane, f, acc);
                    const int kap = tcol + 2 * tq;                     // region column of this lane's pair
                    const int col = c0 - 2 + kap;
                    const bool cin = col >= 0 && col < g.W;            // W even, col even: pair in or out together
                    const uint32_t d0r = *reinterpret_cast<const uint32_t*>(s_du + rho0 * kMmaPitch + kap + 2);
                    const uint32_t d1r = *reinterpret_cast<const uint32_t*>(s_du + (rho0 + 8) * kMmaPitch + kap + 2);
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
                        T* tile = s_dy + k * kDyTileElems;
                        *reinterpret_cast<uint32_t*>(tile + rho0 * kMmaPitch + kap) = lo;
                        *reinterpret_cast<uint32_t*>(tile + (rho0 + 8) * kMmaPitch + kap) = hi;
                    }
                }
            }
            __syncwarp();
            // ---- phase G: Gram products of the block's interior with the x tile.
            // B fragment (ldmatrix.x2.trans): b[0] holds region rows 16wr + 2tq, +1, b[1] rows 16wr + 2tq + 8, +9, all at
            // region column tcol + gq.  Interior = rows [2, min(30, band1 - tr + 2)), columns [2, 58).
            {
                const int rhi = min(kMmaTH - 2, band1 - tr + 2);
                uint32_t rmask[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int r = 16 * wr + 2 * tq + 8 * h;
                    rmask[h] = ((r >= 2 && r < rhi) ? 0x0000ffffu : 0u) | ((r + 1 >= 2 && r + 1 < rhi) ? 0xffff0000u : 0u);
                }
#pragma unroll
                for (int cbk = 0; cbk < 4; ++cbk) {
                    const int tcol = 32 * wc + 8 * cbk;
                    const int kap = tcol + gq;
                    const uint32_t cmask = (kap >= 2 && kap < kMmaTW - 6) ? 0xffffffffu : 0u;
                    uint32_t Bg[4][2];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        load_b_trans(s_dy + k * kDyTileElems, 16 * wr, tcol, lane, Bg[k]);
                        Bg[k][0] &= rmask[0] & cmask;
                        Bg[k][1] &= rmask[1] & cmask;
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
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + pos.slot);              // x and du tiles are consumed
            named_bar_sync(1, kDwThreads);                             // all four blocks of the dy region are written
            // ---- phase 2: dx blocks (2 row tiles x 7 column blocks; warp (wr, wc) takes row tile wr, blocks 4wc..)
            {
                float acc[4][4];
#pragma unroll
                for (int cbk = 0; cbk < 4; ++cbk) acc[cbk][0] = acc[cbk][1] = acc[cbk][2] = acc[cbk][3] = 0.f;
#pragma unroll
                for (int a = 0; a < 5; ++a)
#pragma unroll
                    for (int cbk = 0; cbk < 4; ++cbk) {
                        const int blk = 4 * wc + cbk;
                        if (blk < kDxTW / 8) {
                            const int tcol = 8 * blk;                  // dx column c <-> region index c + b'
                            uint32_t A[4];
                            load_a(s_dy, 16 * wr + a, tcol, lane, A);
                            MmaOp<T>::run(acc[cbk], A, fl.B5[a]);
                            if (a >= 1 && a <= 3) {
                                load_a(s_dy + kDyTileElems, 16 * wr + a, tcol, lane, A);
                                MmaOp<T>::run(acc[cbk], A, fl.B3[a - 1]);
                                load_a(s_dy + 2 * kDyTileElems, 16 * wr + a, tcol, lane, A);
                                MmaOp<T>::run(acc[cbk], A, fl.B31[a - 1]);
                            }
                            if (a == 2) {
                                load_a(s_dy + 3 * kDyTileElems, 16 * wr + a, tcol, lane, A);
                                MmaOp<T>::run(acc[cbk], A, fl.B13);
                            }
                        }
                    }
#pragma unroll
                for (int cbk = 0; cbk < 4; ++cbk) {
                    const int blk = 4 * wc + cbk;
                    const int col = c0 + 8 * blk + 2 * tq;
                    if (blk < kDxTW / 8 && col >= 0 && col < g.W) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int r = 16 * wr + gq + 8 * h, row = tr + r;
#if !defined(LMNET_DBG_NO_DX)
                            if (r < kDxTH && row < band1)
                                *reinterpret_cast<uint32_t*>(dx + poff + (int64_t)row * g.W + col) =
                                    MmaOp<T>::pack(acc[cbk][2 * h], acc[cbk][2 * h + 1]);
#else
                            if (r < kDxTH && row < band1 && acc[cbk][2 * h] == 123.456f) dx[0] = from_f<T>(1.f);
#endif
                        }
                    }
                }
            }
            pos.advance();
            tw.next(ntr);
        }
        // Gram fragments -> shared memory (the dy tiles are free once every warp is past its last phase 2)
        named_bar_sync(1, kDwThreads);
        float* s_G = reinterpret_cast<float*>(s_dy_all);                                    // [4 warps][12][128]
        auto put = [&](int slot, const float (&G)[4]) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = gq + (k >> 1) * 8, j = 2 * tq + (k & 1);
                s_G[(warp * 12 + slot) * 128 + i * 8 + j] = G[k];
            }
        };
#pragma unroll
        for (int a = 0; a < 5; ++a) put(a, G5[a]);
#pragma unroll
        for (int a = 0; a < 3; ++a) { put(5 + a, G3[a]); put(8 + a, G31[a]); }
        put(11, G13);
    }
    __syncthreads();
    if (threadIdx.x < 40) {
        // dw_br[a][bb] = sum_j G[j + bb'][j] with bb' the 5-wide column offset of the tap
        const float* s_G = reinterpret_cast<const float*>(s_dy_all);
        const int t = threadIdx.x;
        int slot, off;
        if (t < 25) { slot = t / 5; off = t % 5; }
        else if (t < 34) { slot = 5 + (t - 25) / 3; off = (t - 25) % 3 + 1; }
        else if (t < 37) { slot = 8 + (t - 34); off = 2; }
        else { slot = 11; off = (t - 37) + 1; }
        float acc = 0.f;
        for (int w = 0; w < kDwWarps; ++w)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += s_G[(w * 12 + slot) * 128 + (j + off) * 8 + j];
        const int ncta = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
        part[((int64_t)e * ncta + cta) * 40 + t] = acc;
    }
}

}  // namespace lmnet
