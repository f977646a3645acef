// wgrad_stream.cuh — warp-streaming variant of the mixed-layout 1x1 weight-gradient kernel (included by wgrad_1x1.cu).
//
// ncu on wgrad_1x1_cl_kernel at LM-Net's two largest resolutions (profiles/r02_ncu_pixel_gemm_wgrad_l1.txt): 44 warp
// instructions per pixel, issue slots 52-75 % busy, DRAM at 16-20 % — the CTA-wide 64-pixel stages leave two
// __syncthreads per 4 k-steps, N tiles are dealt to warps (with N = 12 + the ones column only 2 of 8 warps own MMAs) and
// the staging loops divide per vector.  For skinny shapes (M x N up to 20 MMA tiles) this variant turns the kernel inside
// out:
//   * every WARP streams its own pixel chunks (32 pixels per stage) through a private cp.async ring — no CTA barrier in
//     the main loop, only cp.async.wait_group + __syncwarp;
//   * every warp accumulates the WHOLE M x N result in registers (all warps issue MMAs); the bias gradient comes from an
//     all-ones B fragment held in registers;
//   * the eight warp results are summed in shared memory in a fixed order at the end, one partial per CTA, same partial
//     layout as the CTA-staged kernel (so the same fixed-order reduction kernel finishes the job): deterministic.
// Layout handling is unchanged: plane operands ([C][P]) are read with ldmatrix, channels-last operands ([P][C]) with
// ldmatrix.trans; no transposed copy of an activation exists.
#pragma once

namespace lmnet {

constexpr int kWsKW = 32;                 // pixels per warp stage (two k-steps)
constexpr int kWsPlanePitch = kWsKW + 8;  // 80-byte rows: conflict-free ldmatrix

struct WsGeom {
    int B, M, N1, N2;
    int pitchA, pitchB1, pitchB2;      // element pitches of the staged tiles
    int a_elems, b1_elems, b2_elems;   // elements per stage and warp
    int stages;
    int64_t P;
    int splits, chunks_per_split;      // chunks of kWsKW pixels per CTA
};

__device__ __forceinline__ uint32_t ws_saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ws_cp16(uint32_t s, const void* gmem, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void ws_cp8(uint32_t s, const void* gmem, bool valid) {
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void ws_ldsm_x4(uint32_t (&r)[4], uint32_t a) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ws_ldsm_x4_t(uint32_t (&r)[4], uint32_t a) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ws_ldsm_x2(uint32_t (&r)[2], uint32_t a) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
__device__ __forceinline__ void ws_ldsm_x2_t(uint32_t (&r)[2], uint32_t a) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
template <typename T> __device__ __forceinline__ uint32_t ws_ones();
template <> __device__ __forceinline__ uint32_t ws_ones<__nv_bfloat16>() { return 0x3f803f80u; }
template <> __device__ __forceinline__ uint32_t ws_ones<__half>() { return 0x3c003c00u; }

// one warp stages a plane tile: C channel rows x 32 pixels starting at pixel p0 (channel stride P) -> [C][40]
template <typename T>
__device__ __forceinline__ void ws_issue_planes(uint32_t s, const T* __restrict__ src, int C, int64_t P, int64_t p0, int lane) {
    for (int i = lane; i < C * (kWsKW / 8); i += 32) {
        const int r = i >> 2, v = i & 3;
        const bool ok = p0 + v * 8 < P;                 // P % 8 == 0: a vector is all in or all out
        ws_cp16(s + (uint32_t)(r * kWsPlanePitch + v * 8) * 2, ok ? src + (int64_t)r * P + p0 + v * 8 : src, ok);
    }
}
// one warp stages a channels-last tile: 32 pixels x C channels, contiguous in global memory -> [32][pitch]
template <typename T>
__device__ __forceinline__ void ws_issue_cl(uint32_t s, int pitch, const T* __restrict__ src, int C, int64_t P, int64_t p0, int lane) {
    const int vrows = (int)min((int64_t)kWsKW, P - p0);
    const T* base = src + p0 * C;
    if ((C & 7) == 0) {
        const int vpr = C >> 3, total = kWsKW * vpr, nvalid = vrows * vpr;
        const int dq = 32 / vpr, dr = 32 - dq * vpr;
        int q = lane / vpr, r = lane - q * vpr;
        for (int i = lane; i < total; i += 32) {
            const bool ok = i < nvalid;
            ws_cp16(s + (uint32_t)(q * pitch + r * 8) * 2, ok ? base + (int64_t)i * 8 : src, ok);
            q += dq; r += dr;
            if (r >= vpr) { r -= vpr; ++q; }
        }
    } else {
        const int vpr = C >> 2, total = kWsKW * vpr, nvalid = vrows * vpr;
        const int dq = 32 / vpr, dr = 32 - dq * vpr;
        int q = lane / vpr, r = lane - q * vpr;
        for (int i = lane; i < total; i += 32) {
            const bool ok = i < nvalid;
            ws_cp8(s + (uint32_t)(q * pitch + r * 4) * 2, ok ? base + (int64_t)i * 4 : src, ok);
            q += dq; r += dr;
            if (r >= vpr) { r -= vpr; ++q; }
        }
    }
}

// MT = 16-row tiles of M; NT1 / NT2 = 8-column tiles of B1 / B2 (no ones column: the bias gradient uses a register fragment)
template <typename T, int MT, int NT1, int NT2, bool ACL, bool B1CL>
__global__ void __launch_bounds__(kWgThreads)
wgrad_stream_kernel(const T* __restrict__ A, const T* __restrict__ B1, const T* __restrict__ B2, float* __restrict__ part,
                    WsGeom g, int ldn, int n2_off, int ones_col) {
    constexpr int NT = NT1 + NT2;
    extern __shared__ __align__(16) unsigned char ws_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y, split = blockIdx.x;
    const int stage_elems = g.a_elems + g.b1_elems + g.b2_elems;
    const int warp_elems = g.stages * stage_elems;
    T* s_warp = reinterpret_cast<T*>(ws_smem) + warp * warp_elems;
    {   // padding rows / columns that no copy ever touches must be zero
        uint4* z = reinterpret_cast<uint4*>(s_warp);
        for (int i = lane; i < warp_elems / 8; i += 32) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncwarp();
    const uint32_t s_base = ws_saddr(s_warp);
    const T* Ab = A + (int64_t)b * g.P * g.M;
    const T* B1b = B1 + (int64_t)b * g.P * g.N1;
    const T* B2b = NT2 > 0 ? B2 + (int64_t)b * g.P * g.N2 : nullptr;
    const int64_t nchunks_total = (g.P + kWsKW - 1) / kWsKW;
    const int64_t chunk0 = (int64_t)split * g.chunks_per_split;
    const int nchunks = (int)max((int64_t)0, min((int64_t)g.chunks_per_split, nchunks_total - chunk0));
    const int mine = nchunks > warp ? (nchunks - warp + kWgWarps - 1) / kWgWarps : 0;     // chunks warp, warp + 8, ...

    auto issue = [&](int k, int st) {
        const int64_t p0 = (chunk0 + warp + (int64_t)k * kWgWarps) * kWsKW;
        const uint32_t sa = s_base + (uint32_t)(st * stage_elems) * 2;
        if (ACL) ws_issue_cl<T>(sa, g.pitchA, Ab, g.M, g.P, p0, lane);
        else ws_issue_planes<T>(sa, Ab, g.M, g.P, p0, lane);
        const uint32_t sb1 = sa + (uint32_t)g.a_elems * 2;
        if (B1CL) ws_issue_cl<T>(sb1, g.pitchB1, B1b, g.N1, g.P, p0, lane);
        else ws_issue_planes<T>(sb1, B1b, g.N1, g.P, p0, lane);
        if constexpr (NT2 > 0) ws_issue_cl<T>(sb1 + (uint32_t)g.b1_elems * 2, g.pitchB2, B2b, g.N2, g.P, p0, lane);
    };

    float acc[MT][NT][4], accb[MT][4];
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        accb[i][0] = accb[i][1] = accb[i][2] = accb[i][3] = 0.f;
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
    }
    const uint32_t ones[2] = {ws_ones<T>(), ws_ones<T>()};
    const int mm = lane >> 3, rr = lane & 7, l16 = lane & 15;
    // per-lane fragment offsets (bytes) inside a stage
    const uint32_t a_lane = ACL ? (uint32_t)(((mm >> 1) * 8 + rr) * g.pitchA + (mm & 1) * 8) * 2
                                : (uint32_t)(((mm & 1) * 8 + rr) * kWsPlanePitch + (mm >> 1) * 8) * 2;
    const uint32_t a_tile = ACL ? 32u : (uint32_t)(16 * kWsPlanePitch) * 2;                 // next 16 rows of M
    const uint32_t a_kstep = ACL ? (uint32_t)(16 * g.pitchA) * 2 : 32u;
    const uint32_t b1_lane = (uint32_t)g.a_elems * 2 +
                             (B1CL ? (uint32_t)(l16 * g.pitchB1) * 2 : (uint32_t)((l16 & 7) * kWsPlanePitch + (l16 >> 3) * 8) * 2);
    const uint32_t b1_tile = B1CL ? 16u : (uint32_t)(8 * kWsPlanePitch) * 2;                // next 8 columns of N
    const uint32_t b1_kstep = B1CL ? (uint32_t)(16 * g.pitchB1) * 2 : 32u;
    const uint32_t b2_lane = (uint32_t)(g.a_elems + g.b1_elems) * 2 + (uint32_t)(l16 * g.pitchB2) * 2;
    const uint32_t b2_kstep = (uint32_t)(16 * g.pitchB2) * 2;

    // prologue: stages - 1 chunks in flight (empty commit groups keep the wait_group arithmetic uniform)
    for (int k = 0; k < g.stages - 1; ++k) {
        if (k < mine) issue(k, k);
        cp_async_commit();
    }
    int st = 0;
    for (int k = 0; k < mine; ++k) {
        const int kn = k + g.stages - 1;
        int stn = st + g.stages - 1;
        if (stn >= g.stages) stn -= g.stages;
        if (kn < mine) issue(kn, stn);                 // the slot of chunk k - 1: every lane finished it (syncwarp below)
        cp_async_commit();
        if (g.stages == 3) cp_async_wait<2>(); else cp_async_wait<1>();
        __syncwarp();
        const uint32_t s0 = s_base + (uint32_t)(st * stage_elems) * 2;
#pragma unroll
        for (int ks = 0; ks < kWsKW / 16; ++ks) {
            uint32_t af[MT][4];
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                if (ACL) ws_ldsm_x4_t(af[i], s0 + a_lane + ks * a_kstep + i * a_tile);
                else ws_ldsm_x4(af[i], s0 + a_lane + ks * a_kstep + i * a_tile);
            }
#pragma unroll
            for (int j = 0; j < NT1; ++j) {
                uint32_t bf[2];
                if (B1CL) ws_ldsm_x2_t(bf, s0 + b1_lane + ks * b1_kstep + j * b1_tile);
                else ws_ldsm_x2(bf, s0 + b1_lane + ks * b1_kstep + j * b1_tile);
#pragma unroll
                for (int i = 0; i < MT; ++i) wg_mma<T>(acc[i][j], af[i], bf);
            }
#pragma unroll
            for (int j = 0; j < NT2; ++j) {
                uint32_t bf[2];
                ws_ldsm_x2_t(bf, s0 + b2_lane + ks * b2_kstep + j * 16u);
#pragma unroll
                for (int i = 0; i < MT; ++i) wg_mma<T>(acc[i][NT1 + j], af[i], bf);
            }
#pragma unroll
            for (int i = 0; i < MT; ++i) wg_mma<T>(accb[i], af[i], ones);
        }
        __syncwarp();                                  // all lanes are done reading this stage before it is refilled
        if (++st == g.stages) st = 0;
    }
    cp_async_wait<0>();
    __syncthreads();                                   // every warp is past its ring: reuse the memory for the reduction
    // warp results -> shared slabs [warp][MT*16][ldn] (ldn >= NT*8 covers the ones column), then a fixed-order sum
    float* red = reinterpret_cast<float*>(ws_smem);
    const int gq = lane >> 2, tq = lane & 3;
    const int slab = MT * 16 * ldn;
    float* mine_red = red + warp * slab;
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int m0 = i * 16 + gq;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int col = (j < NT1 ? j * 8 : n2_off + (j - NT1) * 8) + 2 * tq;
            *reinterpret_cast<float2*>(mine_red + m0 * ldn + col) = make_float2(acc[i][j][0], acc[i][j][1]);
            *reinterpret_cast<float2*>(mine_red + (m0 + 8) * ldn + col) = make_float2(acc[i][j][2], acc[i][j][3]);
        }
    }
    __syncthreads();
    if (tq == 0) {                                     // the ones column overlaps a padding column of the last tile: write it last
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            mine_red[(i * 16 + gq) * ldn + ones_col] = accb[i][0];
            mine_red[(i * 16 + gq + 8) * ldn + ones_col] = accb[i][2];
        }
    }
    __syncthreads();
    float* out = part + ((int64_t)b * g.splits + split) * g.M * ldn;
    for (int i = threadIdx.x; i < g.M * ldn; i += kWgThreads) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kWgWarps; ++w) s += red[w * slab + i];
        out[i] = s;
    }
}

}  // namespace lmnet
