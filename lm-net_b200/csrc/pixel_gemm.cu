// pixel_gemm.cu — the 1x1 convolutions of ReparamConv as small-K GEMMs over pixels (sm_100a, 16-bit storage).
//
// Widening step f1 of SURVEY.md §8: expand 1x1 (/root/reference/core/modules.py:537, used :587), SE-gated pointwise 1x1
// + shortcut 1x1 + add (:576-584, used :598-599) and their input gradients.  Per pixel these are K <= 288, N <= 192
// contractions — 24..576 B of traffic per pixel against 0.6..110 kFLOP: HBM-bound by a wide margin, so the kernel is
// built around streaming: the (tiny) weight matrix lives in shared memory for the whole CTA, pixel tiles flow through a
// two-stage cp.async ring, outputs are staged in shared memory and leave as full 16-byte rows.
//
//     out[b] = W1[b] . in1[b]  (+ W2 . in2[b])  (+ bias)            per batch image b, all pixels
//
// Every operand may be "planes" ([C][P], the NCHW layout the depthwise / BatchNorm kernels use) or "channels-last"
// ([P][C], the layout of the block's input and output and of cuDNN's 3x3 convolutions).  The layout change between the
// two rides on the MMA operand order — ldmatrix vs ldmatrix.trans — so no transposed copy of an activation exists:
//     expand forward        in1 = x  (channels-last)                       -> planes      [+ BatchNorm sum / sum^2]
//     pointwise + shortcut  in1 = z  (planes, per-image gated weights), in2 = x (cl)  -> channels-last
//     expand input grad     in1 = dy (planes)                               -> channels-last
//     pointwise input grad  in1 = dout (channels-last, per-image weights)   -> planes
//     shortcut input grad   in1 = dout (channels-last)                      -> channels-last
// Output orientation decides the MMA roles: channels-last output = pixels on M (16 per MMA), channels on N;
// planes output = channels on M, pixels on N — either way the accumulator fragment is written as 32-bit pairs along the
// output's contiguous axis.  mma.sync m16n8k16 (bf16/fp16 in, fp32 accumulate): at < 110 FLOP/B the tensor pipe is
// nowhere near the limit and per-warp 16 x 8 fragments fit the skinny shapes; tcgen05's 128-row tiles do not.
//
// Expand forward can also emit the per-channel sum / sum of squares of its (rounded) output as per-CTA partials in the
// layout of bn_act's finalize kernel, which removes BatchNorm's statistics pass over the expanded tensor.
#include <initializer_list>

#include "common.cuh"

namespace lmnet {

constexpr int kPgThreads = 256;
constexpr int kPgWarps = 8;
constexpr int kPgMaxAcc = 12;          // 16x8 accumulators per warp

struct PgGeom {
    int B, N, K1, K2;                  // real sizes
    int NC;                            // output channels per CTA (grid.z chunks of a channels-last output; N otherwise)
    int K1p, K2p;                      // padded to 16
    int64_t P;                         // pixels per image
    int PT;                            // pixels per CTA tile
    int pt_shift;                      // PT / 8 == 1 << pt_shift (PT is 64, 128 or 256)
    int tiles, ctas_x, tiles_per_cta;  // per image
    int w_pitch;                       // element pitch of the weight rows [N][K1p + K2p + 8]
    int in1_pitch, in2_pitch, out_pitch;
};

__device__ __forceinline__ void pg_cp16(void* smem, const void* gmem, bool valid) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void pg_cp8(void* smem, const void* gmem, bool valid) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void pg_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void pg_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ void pg_ldsm_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void pg_ldsm_x4_t(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void pg_ldsm_x2(uint32_t (&r)[2], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];"
                 : "=r"(r[0]), "=r"(r[1]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void pg_ldsm_x2_t(uint32_t (&r)[2], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
                 : "=r"(r[0]), "=r"(r[1]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
template <typename T> __device__ __forceinline__ void pg_mma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]);
template <> __device__ __forceinline__ void pg_mma<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <> __device__ __forceinline__ void pg_mma<__half>(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <typename T> __device__ __forceinline__ uint32_t pg_pack(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pg_pack<__nv_bfloat16>(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
template <> __device__ __forceinline__ uint32_t pg_pack<__half>(float lo, float hi) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// ---- shared-memory addressing in the 32-bit shared window (no generic-pointer arithmetic in the hot loops)
__device__ __forceinline__ uint32_t pg_saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pg_cp16a(uint32_t s, const void* gmem, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void pg_cp8a(uint32_t s, const void* gmem, bool valid) {
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void pg_ldsm_x4a(uint32_t (&r)[4], uint32_t a) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void pg_ldsm_x4ta(uint32_t (&r)[4], uint32_t a) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void pg_ldsm_x2a(uint32_t (&r)[2], uint32_t a) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}

// (i / d, i % d) for i = i0, i0 + step, ... without a division per iteration
struct PgDiv {
    int q, r;
    __device__ __forceinline__ PgDiv(int i0, int d) { q = i0 / d; r = i0 - q * d; }
    __device__ __forceinline__ void step(int dq, int dr, int d) {
        q += dq;
        r += dr;
        if (r >= d) { r -= d; ++q; }
    }
};

// stage a channels-last tile: `rows` pixel rows x C channels, contiguous in global memory -> smem [rows][pitch]
template <typename T>
__device__ __forceinline__ void pg_issue_cl(uint32_t s, int pitch, const T* src, int C, int rows, int64_t valid_rows) {
    const int vrows = (int)min((int64_t)rows, valid_rows);
    if ((C & 7) == 0) {
        const int vpr = C >> 3, total = rows * vpr, nvalid = vrows * vpr;
        const int dq = kPgThreads / vpr, dr = kPgThreads - dq * vpr;
        PgDiv d(threadIdx.x, vpr);
        for (int i = threadIdx.x; i < total; i += kPgThreads) {
            const bool ok = i < nvalid;
            pg_cp16a(s + (uint32_t)(d.q * pitch + d.r * 8) * 2, ok ? src + (int64_t)i * 8 : src, ok);
            d.step(dq, dr, vpr);
        }
    } else {
        const int vpr = C >> 2, total = rows * vpr, nvalid = vrows * vpr;
        const int dq = kPgThreads / vpr, dr = kPgThreads - dq * vpr;
        PgDiv d(threadIdx.x, vpr);
        for (int i = threadIdx.x; i < total; i += kPgThreads) {
            const bool ok = i < nvalid;
            pg_cp8a(s + (uint32_t)(d.q * pitch + d.r * 4) * 2, ok ? src + (int64_t)i * 4 : src, ok);
            d.step(dq, dr, vpr);
        }
    }
}
// stage a planes tile: C channel rows x `cols` pixels (cols = 8 << sh) at src (channel stride P) -> smem [C][pitch]
template <typename T>
__device__ __forceinline__ void pg_issue_planes(uint32_t s, int pitch, const T* src, int C, int64_t P, int sh, int64_t valid_cols) {
    const int mask = (1 << sh) - 1;
    for (int i = threadIdx.x; i < (C << sh); i += kPgThreads) {
        const int r = i >> sh, v = i & mask;
        const bool ok = v * 8 < valid_cols;             // P % 8 == 0: a vector is all in or all out
        pg_cp16a(s + (uint32_t)(r * pitch + v * 8) * 2, ok ? src + (int64_t)r * P + v * 8 : src, ok);
    }
}

// ---------------------------------------------------------------------------------------------------
// OUT_CL = true : pixels on M.  A warp owns MW m-tiles (16 pixels each) and all NT n-tiles (8 channels each).
// OUT_CL = false: channels on M. A warp owns NW pixel n-tiles (8 pixels each) and all MT m-tiles (16 channels each).
// TA = tiles along the warp's private axis (MW or NW), TB = tiles along the shared channel axis (NT or MT).
// ---------------------------------------------------------------------------------------------------
template <typename T, bool OUT_CL, bool IN1_CL, bool HAS_IN2, int TA, int TB>
__global__ void __launch_bounds__(kPgThreads)
pixel_gemm_kernel(const T* __restrict__ in1, const T* __restrict__ in2, const lmnet_pgemm_weights w,
                  const float* __restrict__ bias, T* __restrict__ out, float* __restrict__ stats_part, PgGeom g) {
    static_assert(TA * TB <= kPgMaxAcc, "accumulator budget");
    static_assert(OUT_CL || IN1_CL, "planes -> planes is not needed by the block");
    static_assert(OUT_CL || !HAS_IN2, "a second operand is only used by channels-last outputs");
    extern __shared__ __align__(16) unsigned char pg_smem[];
    T* s_w = reinterpret_cast<T*>(pg_smem);                                  // [Npad][w_pitch]
    constexpr int n_pad = OUT_CL ? TB * 8 : TB * 16;
    const int in1_elems = IN1_CL ? g.PT * g.in1_pitch : g.K1p * g.in1_pitch;
    const int in2_elems = HAS_IN2 ? g.PT * g.in2_pitch : 0;
    const int stage_elems = in1_elems + in2_elems;
    T* s_in = s_w + n_pad * g.w_pitch;                                       // 2 stages
    T* s_out = s_in + 2 * stage_elems;                                       // OUT_CL: [PT][out_pitch]; else [Npad][out_pitch]
    float* s_bias = reinterpret_cast<float*>(s_out + (OUT_CL ? g.PT : n_pad) * g.out_pitch);   // [n_pad]
    float* s_stat = s_bias + n_pad;                                          // [8 warps][n_pad][2] (planes output only)
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    const int n0 = blockIdx.z * g.NC;                                        // first output channel of this CTA
    const T zero = from_f<T>(0.f);

    // ---- one-time: weights (zero padded), bias, zero the padding of the input stages
    {
        uint4* z = reinterpret_cast<uint4*>(s_w);                            // w_pitch is a multiple of 8 elements
        for (int i = threadIdx.x; i < (n_pad * g.w_pitch) >> 3; i += kPgThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    {   // fp32 parameters, read in memory order (coalesced whichever way the view is transposed), rounded here (RN, as a
        // host-side cast would) and scattered to [n][k]
        const bool k_fast = w.w1_sk <= w.w1_sn;
        const int inner = k_fast ? g.K1 : g.NC;
        PgDiv d(threadIdx.x, inner);
        const int dq = kPgThreads / inner, dr = kPgThreads - dq * inner;
        for (int i = threadIdx.x; i < g.NC * g.K1; i += kPgThreads) {
            const int n = k_fast ? d.q : d.r, k = k_fast ? d.r : d.q;
            float f = __ldg(w.w1 + (int64_t)(n0 + n) * w.w1_sn + (int64_t)k * w.w1_sk);
            if (w.gate != nullptr) f *= __ldg(w.gate + (w.gate_on_n ? (int64_t)b * g.N + n0 + n : (int64_t)b * g.K1 + k));
            s_w[n * g.w_pitch + k] = from_f<T>(f);
            d.step(dq, dr, inner);
        }
        if constexpr (HAS_IN2) {
            const bool k_fast2 = w.w2_sk <= w.w2_sn;
            const int inner2 = k_fast2 ? g.K2 : g.NC;
            PgDiv d2(threadIdx.x, inner2);
            const int dq2 = kPgThreads / inner2, dr2 = kPgThreads - dq2 * inner2;
            for (int i = threadIdx.x; i < g.NC * g.K2; i += kPgThreads) {
                const int n = k_fast2 ? d2.q : d2.r, k = k_fast2 ? d2.r : d2.q;
                s_w[n * g.w_pitch + g.K1p + k] = from_f<T>(__ldg(w.w2 + (int64_t)(n0 + n) * w.w2_sn + (int64_t)k * w.w2_sk));
                d2.step(dq2, dr2, inner2);
            }
        }
    }
    for (int i = threadIdx.x; i < n_pad; i += kPgThreads) s_bias[i] = (bias != nullptr && i < g.NC) ? bias[n0 + i] : 0.f;
    {
        uint4* z = reinterpret_cast<uint4*>(s_in);                           // stage sizes are multiples of 8 elements
        for (int i = threadIdx.x; i < (2 * stage_elems) >> 3; i += kPgThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();

    const T* in1_b = in1 + (int64_t)b * g.P * g.K1;
    const T* in2_b = HAS_IN2 ? in2 + (int64_t)b * g.P * g.K2 : nullptr;
    T* out_b = out + (int64_t)b * g.P * g.N;
    const int tile0 = blockIdx.x * g.tiles_per_cta;
    const int ntiles = max(0, min(g.tiles_per_cta, g.tiles - tile0));
    const uint32_t sin_a = pg_saddr(s_in), sw_a = pg_saddr(s_w);

    auto issue = [&](int t, int st) {
        const uint32_t s1 = sin_a + (uint32_t)(st * stage_elems) * 2;
        const int64_t p0 = (int64_t)(tile0 + t) * g.PT;
        const int64_t valid = g.P - p0;
        if (IN1_CL) pg_issue_cl<T>(s1, g.in1_pitch, in1_b + p0 * g.K1, g.K1, g.PT, valid);
        else pg_issue_planes<T>(s1, g.in1_pitch, in1_b + p0, g.K1, g.P, g.pt_shift, valid);
        if (HAS_IN2) pg_issue_cl<T>(s1 + (uint32_t)in1_elems * 2, g.in2_pitch, in2_b + p0 * g.K2, g.K2, g.PT, valid);
        pg_commit();
    };

    float st_s[TB][2], st_q[TB][2];                     // BatchNorm partial sums (planes output): rows gq, gq + 8 of each m-tile
#pragma unroll
    for (int i = 0; i < TB; ++i) st_s[i][0] = st_s[i][1] = st_q[i][0] = st_q[i][1] = 0.f;

    // per-lane fragment offsets (bytes) that do not change from tile to tile
    const int m = lane >> 3, rr = lane & 7, l16 = lane & 15;
    uint32_t a_lane, a_lane2 = 0, b_lane;
    if constexpr (OUT_CL) {
        const int px0 = warp * TA * 16;
        // A = input pixels
        if (IN1_CL) a_lane = (uint32_t)((px0 + (m & 1) * 8 + rr) * g.in1_pitch + (m >> 1) * 8) * 2;
        else a_lane = (uint32_t)(((m >> 1) * 8 + rr) * g.in1_pitch + px0 + (m & 1) * 8) * 2;
        if (HAS_IN2) a_lane2 = (uint32_t)(in1_elems + (px0 + (m & 1) * 8 + rr) * g.in2_pitch + (m >> 1) * 8) * 2;
        // B = weights [n][k], x2: rows l & 7, k half l >> 3
        b_lane = sw_a + (uint32_t)((l16 & 7) * g.w_pitch + (l16 >> 3) * 8) * 2;
    } else {
        const int px0 = warp * TA * 8;
        // A = weights [n][k], x4
        a_lane = sw_a + (uint32_t)(((m & 1) * 8 + rr) * g.w_pitch + (m >> 1) * 8) * 2;
        // B = input pixels (channels-last), x2
        b_lane = (uint32_t)((px0 + (l16 & 7)) * g.in1_pitch + (l16 >> 3) * 8) * 2;
    }
    const uint32_t a_tile = OUT_CL ? (uint32_t)(IN1_CL ? 16 * g.in1_pitch : 16) * 2 : (uint32_t)(16 * g.w_pitch) * 2;
    const uint32_t a_tile2 = (uint32_t)(16 * g.in2_pitch) * 2;
    const uint32_t a_kstep = (OUT_CL && !IN1_CL) ? (uint32_t)(16 * g.in1_pitch) * 2 : 32u;
    const uint32_t b_tile = OUT_CL ? (uint32_t)(8 * g.w_pitch) * 2 : (uint32_t)(8 * g.in1_pitch) * 2;
    const int ksteps1 = g.K1p >> 4, ksteps2 = HAS_IN2 ? g.K2p >> 4 : 0;

    if (ntiles > 0) issue(0, 0);
    for (int t = 0; t < ntiles; ++t) {
        const int st = t & 1;
        if (t + 1 < ntiles) {
            issue(t + 1, st ^ 1);
            pg_wait<1>();
        } else {
            pg_wait<0>();
        }
        __syncthreads();                                // tile t landed; everybody is done with s_out of tile t-1
        const uint32_t s1 = sin_a + (uint32_t)(st * stage_elems) * 2;
        const int64_t p0 = (int64_t)(tile0 + t) * g.PT;
        float acc[TA][TB][4];
#pragma unroll
        for (int i = 0; i < TA; ++i)
#pragma unroll
            for (int j = 0; j < TB; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;

        if constexpr (OUT_CL) {
            // pixels on M: A = input (this warp's TA x 16 pixels), B = weights [n][k]
            uint32_t pa = s1 + a_lane, pb = b_lane;
            for (int ks = 0; ks < ksteps1; ++ks, pa += a_kstep, pb += 32u) {
                uint32_t bf[TB][2];
#pragma unroll
                for (int j = 0; j < TB; ++j) pg_ldsm_x2a(bf[j], pb + j * b_tile);
#pragma unroll
                for (int i = 0; i < TA; ++i) {
                    uint32_t af[4];
                    if (IN1_CL) pg_ldsm_x4a(af, pa + i * a_tile);
                    else pg_ldsm_x4ta(af, pa + i * a_tile);
#pragma unroll
                    for (int j = 0; j < TB; ++j) pg_mma<T>(acc[i][j], af, bf[j]);
                }
            }
            if constexpr (HAS_IN2) {
                uint32_t pa2 = s1 + a_lane2;
                pb = b_lane + (uint32_t)g.K1p * 2;
                for (int ks = 0; ks < ksteps2; ++ks, pa2 += 32u, pb += 32u) {
                    uint32_t bf[TB][2];
#pragma unroll
                    for (int j = 0; j < TB; ++j) pg_ldsm_x2a(bf[j], pb + j * b_tile);
#pragma unroll
                    for (int i = 0; i < TA; ++i) {
                        uint32_t af[4];
                        pg_ldsm_x4a(af, pa2 + i * a_tile2);
#pragma unroll
                        for (int j = 0; j < TB; ++j) pg_mma<T>(acc[i][j], af, bf[j]);
                    }
                }
            }
            // fragments -> staging [pixel][channel]
            const int px0 = warp * TA * 16;
#pragma unroll
            for (int j = 0; j < TB; ++j) {
                const int n = j * 8 + 2 * tq;
                const float b0 = s_bias[n], b1 = s_bias[n + 1];
#pragma unroll
                for (int i = 0; i < TA; ++i) {
                    const int r0 = px0 + i * 16 + gq;
                    *reinterpret_cast<uint32_t*>(s_out + r0 * g.out_pitch + n) = pg_pack<T>(acc[i][j][0] + b0, acc[i][j][1] + b1);
                    *reinterpret_cast<uint32_t*>(s_out + (r0 + 8) * g.out_pitch + n) = pg_pack<T>(acc[i][j][2] + b0, acc[i][j][3] + b1);
                }
            }
            __syncthreads();
            // staging -> global: PT pixel rows of NC channels (one contiguous chunk when NC == N)
            const int valid = (int)min((int64_t)g.PT, g.P - p0);
            T* dst = out_b + p0 * g.N + n0;
            if ((g.NC & 7) == 0 && (g.N & 7) == 0) {
                const int vpr = g.NC >> 3, total = valid * vpr;
                const int dq = kPgThreads / vpr, dr = kPgThreads - dq * vpr;
                PgDiv d(threadIdx.x, vpr);
                for (int i = threadIdx.x; i < total; i += kPgThreads) {
                    *reinterpret_cast<uint4*>(dst + (int64_t)d.q * g.N + d.r * 8) = *reinterpret_cast<const uint4*>(s_out + d.q * g.out_pitch + d.r * 8);
                    d.step(dq, dr, vpr);
                }
            } else {
                const int vpr = g.NC >> 2, total = valid * vpr;
                const int dq = kPgThreads / vpr, dr = kPgThreads - dq * vpr;
                PgDiv d(threadIdx.x, vpr);
                for (int i = threadIdx.x; i < total; i += kPgThreads) {
                    *reinterpret_cast<uint2*>(dst + (int64_t)d.q * g.N + d.r * 4) = *reinterpret_cast<const uint2*>(s_out + d.q * g.out_pitch + d.r * 4);
                    d.step(dq, dr, vpr);
                }
            }
        } else {
            // channels on M: A = weights [n][k], B = input pixels (this warp's TA x 8 pixels), channels-last input
            uint32_t pa = a_lane, pb = s1 + b_lane;
            for (int ks = 0; ks < ksteps1; ++ks, pa += 32u, pb += 32u) {
                uint32_t bf[TA][2];
#pragma unroll
                for (int i = 0; i < TA; ++i) pg_ldsm_x2a(bf[i], pb + i * b_tile);
#pragma unroll
                for (int j = 0; j < TB; ++j) {
                    uint32_t af[4];
                    pg_ldsm_x4a(af, pa + j * a_tile);
#pragma unroll
                    for (int i = 0; i < TA; ++i) pg_mma<T>(acc[i][j], af, bf[i]);
                }
            }
            // fragments -> staging [channel][pixel]; BatchNorm partial sums of the values as stored
            const int px0 = warp * TA * 8;
            const int valid = (int)min((int64_t)g.PT, g.P - p0);
#pragma unroll
            for (int j = 0; j < TB; ++j) {
                const int nn = j * 16 + gq;
                const float b0 = s_bias[nn], b1 = s_bias[nn + 8];
#pragma unroll
                for (int i = 0; i < TA; ++i) {
                    const int px = px0 + i * 8 + 2 * tq;
                    const uint32_t lo = pg_pack<T>(acc[i][j][0] + b0, acc[i][j][1] + b0);
                    const uint32_t hi = pg_pack<T>(acc[i][j][2] + b1, acc[i][j][3] + b1);
                    *reinterpret_cast<uint32_t*>(s_out + nn * g.out_pitch + px) = lo;
                    *reinterpret_cast<uint32_t*>(s_out + (nn + 8) * g.out_pitch + px) = hi;
                    if (stats_part != nullptr && px < valid) {           // P even: the pair is in or out together
                        const T* le = reinterpret_cast<const T*>(&lo);
                        const T* he = reinterpret_cast<const T*>(&hi);
                        const float a0 = to_f(le[0]), a1 = to_f(le[1]), c0 = to_f(he[0]), c1 = to_f(he[1]);
                        st_s[j][0] += a0 + a1; st_q[j][0] = fmaf(a0, a0, fmaf(a1, a1, st_q[j][0]));
                        st_s[j][1] += c0 + c1; st_q[j][1] = fmaf(c0, c0, fmaf(c1, c1, st_q[j][1]));
                    }
                }
            }
            __syncthreads();
            // staging -> global: each channel row is PT contiguous pixels (PT / 8 = 1 << pt_shift vectors)
            T* dst = out_b + p0;
            const int sh = g.pt_shift, mask = (1 << sh) - 1;
            for (int i = threadIdx.x; i < (g.N << sh); i += kPgThreads) {
                const int r = i >> sh, v = i & mask;
                if (v * 8 < valid)
                    *reinterpret_cast<uint4*>(dst + (int64_t)r * g.P + v * 8) = *reinterpret_cast<const uint4*>(s_out + r * g.out_pitch + v * 8);
            }
        }
    }
    if constexpr (!OUT_CL) {
        if (stats_part != nullptr) {
            // per lane: rows gq / gq+8 of every m-tile; reduce over the 4 lanes of a quad, then over the 8 warps (fixed order)
#pragma unroll
            for (int j = 0; j < TB; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float s = st_s[j][h], q = st_q[j][h];
                    s += __shfl_xor_sync(0xffffffffu, s, 1); q += __shfl_xor_sync(0xffffffffu, q, 1);
                    s += __shfl_xor_sync(0xffffffffu, s, 2); q += __shfl_xor_sync(0xffffffffu, q, 2);
                    if (tq == 0) {
                        const int n = j * 16 + gq + 8 * h;
                        s_stat[(warp * n_pad + n) * 2] = s;
                        s_stat[(warp * n_pad + n) * 2 + 1] = q;
                    }
                }
            __syncthreads();
            const int ncta = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
            for (int n = threadIdx.x; n < g.N; n += kPgThreads) {
                float s = 0.f, q = 0.f;
                for (int w = 0; w < kPgWarps; ++w) { s += s_stat[(w * n_pad + n) * 2]; q += s_stat[(w * n_pad + n) * 2 + 1]; }
                stats_part[((int64_t)n * ncta + cta) * 2] = s;
                stats_part[((int64_t)n * ncta + cta) * 2 + 1] = q;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static int pg_pitch(int cols) {            // multiple of 8 elements, == 8 (mod 16): conflict-free ldmatrix rows and pair stores
    int p = (cols + 7) / 8 * 8;
    if (p % 16 != 8) p += 8;
    return p;
}

struct PgPlan {
    PgGeom g;
    int TA, TB;
    size_t smem;
    dim3 grid;
};

static bool pg_plan(const lmnet_pgemm_dims* d, bool in1_cl, bool out_cl, bool stats, PgPlan& pl) {
    if (d == nullptr || d->B <= 0 || d->P <= 0 || d->N <= 0 || d->K1 <= 0 || d->K2 < 0) return false;
    if (d->P % 8 != 0) return false;
    if (!out_cl && (!in1_cl || d->K2 > 0)) return false;
    if (in1_cl && d->K1 % 4 != 0) return false;
    if (d->K2 > 0 && d->K2 % 4 != 0) return false;
    if (out_cl && d->N % 4 != 0) return false;
    if (stats && out_cl) return false;
    PgGeom& g = pl.g;
    g.B = d->B; g.N = d->N; g.K1 = d->K1; g.K2 = d->K2; g.P = d->P;
    g.K1p = (d->K1 + 15) / 16 * 16;
    g.K2p = d->K2 > 0 ? (d->K2 + 15) / 16 * 16 : 0;
    g.w_pitch = pg_pitch(g.K1p + g.K2p);
    g.NC = d->N;
    int nchunks = 1;
    if (out_cl) {
        nchunks = (d->N + 8 * kPgMaxAcc - 1) / (8 * kPgMaxAcc);  // wide outputs: grid.z chunks of <= 96 channels
        if (d->N % nchunks != 0 || (d->N / nchunks) % 4 != 0) return false;
        g.NC = d->N / nchunks;
        pl.TB = (g.NC + 7) / 8;                                  // n-tiles of 8 channels
    } else {
        pl.TB = (d->N + 15) / 16;                                // m-tiles of 16 channels
    }
    if (pl.TB > kPgMaxAcc) return false;
    pl.TA = kPgMaxAcc / pl.TB;
    if (pl.TA > 4) pl.TA = 4;
    if (pl.TA == 3) pl.TA = 2;
    // The pixel tile (PT = warps x TA x 16 | 8 pixels) shrinks until the double-buffered operand tiles fit: the SE-gated
    // pointwise + shortcut GEMM of level 2 (48 planes + 24 channels-last inputs) needs 230 KB at TA = 4 and went to
    // cuBLAS baddbmm with a broadcast copy of the bias per call.  Instantiated (TA, TB): see pg_dispatch_tiles.
    for (;; pl.TA /= 2) {
        if (out_cl) {
            g.PT = kPgWarps * pl.TA * 16;
            g.in1_pitch = in1_cl ? pg_pitch(g.K1p) : pg_pitch(g.PT);
            g.in2_pitch = d->K2 > 0 ? pg_pitch(g.K2p) : 0;
            g.out_pitch = pg_pitch(pl.TB * 8);
        } else {
            g.PT = kPgWarps * pl.TA * 8;
            g.in1_pitch = pg_pitch(g.K1p);
            g.in2_pitch = 0;
            g.out_pitch = pg_pitch(g.PT);
        }
        g.pt_shift = 0;
        while ((8 << g.pt_shift) < g.PT) ++g.pt_shift;
        if ((8 << g.pt_shift) != g.PT) return false;
        const int n_pad = out_cl ? pl.TB * 8 : pl.TB * 16;
        const size_t in1_elems = in1_cl ? (size_t)g.PT * g.in1_pitch : (size_t)g.K1p * g.in1_pitch;
        const size_t in2_elems = d->K2 > 0 ? (size_t)g.PT * g.in2_pitch : 0;
        const size_t out_elems = (size_t)(out_cl ? g.PT : n_pad) * g.out_pitch;
        pl.smem = ((size_t)n_pad * g.w_pitch + 2 * (in1_elems + in2_elems) + out_elems) * 2 + (size_t)n_pad * 4 +
                  (stats ? (size_t)kPgWarps * n_pad * 2 * 4 : 0) + 16;
        if (pl.smem <= 200 * 1024) break;
        const bool smaller_exists = (pl.TA == 4 && pl.TB == 3);              // (2, 3) is the one extra instantiation
        if (!smaller_exists) {
            if (pl.smem <= 226 * 1024) break;                                // one CTA per SM still beats the cuBLAS detour
            return false;
        }
    }
    g.tiles = (int)((d->P + g.PT - 1) / g.PT);
    int ctas_x = (2 * 148 + d->B * nchunks - 1) / (d->B * nchunks);   // ~2 CTAs per SM over the whole grid
    if (ctas_x > g.tiles) ctas_x = g.tiles;
    if (ctas_x < 1) ctas_x = 1;
    g.tiles_per_cta = (g.tiles + ctas_x - 1) / ctas_x;
    g.ctas_x = (g.tiles + g.tiles_per_cta - 1) / g.tiles_per_cta;
    pl.grid = dim3((unsigned)g.ctas_x, (unsigned)d->B, (unsigned)nchunks);
    return true;
}

template <typename T, bool OUT_CL, bool IN1_CL, bool HAS_IN2, int TA, int TB>
static int pg_launch(const void* in1, const void* in2, const lmnet_pgemm_weights& w, const float* bias, void* out,
                     float* stats_part, const PgPlan& pl, cudaStream_t st) {
    auto kern = pixel_gemm_kernel<T, OUT_CL, IN1_CL, HAS_IN2, TA, TB>;
    static std::atomic<size_t> granted[kMaxDevices];
    if (!ensure_smem(kern, pl.smem, granted)) return LMNET_ERR_LAUNCH;
    const PgGeom& g = pl.g;
    const double bytes = (double)g.B * g.P * (g.K1 + g.K2 + g.N) * sizeof(T);
    LMNET_LAUNCH(KID_PIXEL_GEMM, st, bytes, (kern<<<pl.grid, kPgThreads, pl.smem, st>>>(
        (const T*)in1, (const T*)in2, w, bias, (T*)out, stats_part, g)));
    return LMNET_OK;
}

template <typename T, bool OUT_CL, bool IN1_CL, bool HAS_IN2>
static int pg_dispatch_tiles(const void* in1, const void* in2, const lmnet_pgemm_weights& w, const float* bias, void* out,
                             float* stats_part, const PgPlan& pl, cudaStream_t st) {
#define PG_CASE(A, B) \
    if (pl.TA == A && pl.TB == B) return pg_launch<T, OUT_CL, IN1_CL, HAS_IN2, A, B>(in1, in2, w, bias, out, stats_part, pl, st);
    PG_CASE(4, 1) PG_CASE(4, 2) PG_CASE(4, 3) PG_CASE(2, 3) PG_CASE(2, 4) PG_CASE(2, 5) PG_CASE(2, 6) PG_CASE(1, 7) PG_CASE(1, 8)
    PG_CASE(1, 9) PG_CASE(1, 10) PG_CASE(1, 11) PG_CASE(1, 12)
#undef PG_CASE
    return LMNET_ERR_UNSUPPORTED;
}

template <typename T>
static int pg_dispatch(bool in1_cl, bool out_cl, bool has_in2, const void* in1, const void* in2, const lmnet_pgemm_weights& w,
                       const float* bias, void* out, float* stats_part, const PgPlan& pl, cudaStream_t st) {
    if (out_cl) {
        if (in1_cl) {
            if (has_in2) return LMNET_ERR_UNSUPPORTED;
            return pg_dispatch_tiles<T, true, true, false>(in1, in2, w, bias, out, stats_part, pl, st);
        }
        if (has_in2) return pg_dispatch_tiles<T, true, false, true>(in1, in2, w, bias, out, stats_part, pl, st);
        return pg_dispatch_tiles<T, true, false, false>(in1, in2, w, bias, out, stats_part, pl, st);
    }
    return pg_dispatch_tiles<T, false, true, false>(in1, in2, w, bias, out, stats_part, pl, st);
}

}  // namespace lmnet

using namespace lmnet;

extern "C" int lmnet_pixel_gemm_supported(const lmnet_pgemm_dims* d, int in1_cl, int out_cl, int want_stats, int dtype) {
    if (dtype != LMNET_BF16 && dtype != LMNET_F16) return 0;
    PgPlan pl;
    return pg_plan(d, in1_cl != 0, out_cl != 0, want_stats != 0, pl) ? 1 : 0;
}

extern "C" int lmnet_pixel_gemm_stats_ctas(const lmnet_pgemm_dims* d, int in1_cl, int out_cl) {
    PgPlan pl;
    if (!pg_plan(d, in1_cl != 0, out_cl != 0, true, pl)) return 0;
    return (int)(pl.grid.x * pl.grid.y);
}

extern "C" int lmnet_pixel_gemm(const void* in1, int in1_cl, const void* in2, const lmnet_pgemm_weights* w, const float* bias,
                                void* out, int out_cl, float* stats_part, const lmnet_pgemm_dims* d, int dtype, void* stream) {
    if (!lmnet_pixel_gemm_supported(d, in1_cl, out_cl, stats_part != nullptr, dtype)) return LMNET_ERR_UNSUPPORTED;
    if (!in1 || !w || !w->w1 || !out || (d->K2 > 0 && (!in2 || !w->w2))) return LMNET_ERR_INVALID_ARG;
    for (const void* q : {in1, in2, (const void*)out})
        if ((uintptr_t)q % 16 != 0) return LMNET_ERR_UNSUPPORTED;
    PgPlan pl;
    pg_plan(d, in1_cl != 0, out_cl != 0, stats_part != nullptr, pl);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == LMNET_BF16)
        return pg_dispatch<__nv_bfloat16>(in1_cl != 0, out_cl != 0, d->K2 > 0, in1, in2, *w, bias, out, stats_part, pl, st);
    return pg_dispatch<__half>(in1_cl != 0, out_cl != 0, d->K2 > 0, in1, in2, *w, bias, out, stats_part, pl, st);
}
