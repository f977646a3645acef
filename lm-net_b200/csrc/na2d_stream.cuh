// na2d_stream.cuh — row-streaming fused neighbourhood attention for 16-bit storage (sm_100a).
//
// Second generation of the fused op of na2d_fused.cuh (same semantics: q*scale -> na2d_qk(+rpb) -> softmax ->
// na2d_av of natten's NeighborhoodAttention2D, /root/reference/core/modules.py:517, and its autograd; SURVEY.md
// §8 a2-a5, d4).  na2d_fused.cuh stays as the general path (fp32 storage, any K <= 13, strided heads, tiny maps).
//
// What changed, and why (profiles/r01_ncu_na_*): the first generation is issue-bound, ~250 thread-instructions
// per (pixel, head): 64-bit address arithmetic for 18 global loads, bf16->fp32 conversions of every neighbour,
// scalar rpb loads; its backward round-trips (lse, delta) = 8 B per (pixel, head) through HBM, which at head
// dim 1 is 4x the size of q itself.  Here:
//   * a CTA owns a column stripe x row band of one (batch, dilation sub-grid) and streams down the rows; the
//     key/value rows (and, in the backward, query/dout rows and the per-query softmax statistics) live in
//     shared-memory rings filled by cp.async one row ahead of the compute, one __syncthreads per row.
//     (TMA tensor maps need 16-byte global strides; q/k/v are slices of the packed qkv tensor with a
//     72-byte pixel stride at head dim 1, so the rings are filled with 8-byte LDGSTS instead.)
//   * every neighbour address is ring-row base (uniform) + thread constant + immediate;
//   * dot products use the sm_100 mixed-precision FMA (fma.rn.f32.bf16 / .f16 -> FHFMA: fp32 accumulate,
//     16-bit operands picked as .H0/.H1 of the packed registers, exact products), so nothing is converted;
//     probabilities / dS are rounded to the storage type for the second contraction exactly like the
//     reference's autocast path, which materialises attn and dattn in 16 bits;
//   * the backward is ONE kernel: phase A (per query, incl. a one-pixel halo of queries) recomputes P, dP,
//     delta, dS, writes dq and leaves (lse, delta) in a shared-memory ring; phase B (per key, one row behind)
//     gathers dk, dv over the inverse neighbourhood from the rings.  No atomics on dq/dk/dv; drpb goes through
//     per-CTA partial tables and the fixed-order reduce of na2d_fused.cuh  => deterministic.
#pragma once
#include "na2d_fused.cuh"

namespace lmnet {

constexpr int kStreamThreads = 192;  // = QW * (heads / HG); 12 heads: 64x3, 32x6, 16x12

struct StreamCfg {
    int NG;        // head groups per pixel (heads / HG)
    int QW;        // query columns a CTA works on per row
    int RB;        // rows owned per band
};

// ---- mixed-precision primitives ------------------------------------------------------------------------
template <typename T> struct Mixed;
template <> struct Mixed<__nv_bfloat16> {
    static __device__ __forceinline__ float fma(uint16_t a, uint16_t b, float c) {
        float d;
        asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
        return d;
    }
    static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
        uint32_t r;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
        return r;
    }
};
template <> struct Mixed<__half> {
    static __device__ __forceinline__ float fma(uint16_t a, uint16_t b, float c) {
        float d;
        asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
        return d;
    }
    static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
        uint32_t r;
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
        return r;
    }
};

__device__ __forceinline__ uint16_t half_of(uint32_t w, int h) {
    return h ? (uint16_t)(w >> 16) : (uint16_t)(w & 0xffffu);
}
template <int VW> __device__ __forceinline__ uint16_t elem_of(const uint32_t (&w)[VW], int e) {
    return half_of(w[e >> 1], e & 1);
}

// VB-byte vector <-> packed 32-bit words (VB in {2,4,8,16}); works for shared and global pointers
template <int VB, int VW> __device__ __forceinline__ void load_words(const void* p, uint32_t (&w)[VW]) {
    if constexpr (VB == 16) { uint4 t = *reinterpret_cast<const uint4*>(p); w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w; }
    else if constexpr (VB == 8) { uint2 t = *reinterpret_cast<const uint2*>(p); w[0] = t.x; w[1] = t.y; }
    else if constexpr (VB == 4) { w[0] = *reinterpret_cast<const uint32_t*>(p); }
    else { w[0] = *reinterpret_cast<const uint16_t*>(p); }
}
template <int VB, int VW> __device__ __forceinline__ void store_words(void* p, const uint32_t (&w)[VW]) {
    if constexpr (VB == 16) { *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]); }
    else if constexpr (VB == 8) { *reinterpret_cast<uint2*>(p) = make_uint2(w[0], w[1]); }
    else if constexpr (VB == 4) { *reinterpret_cast<uint32_t*>(p) = w[0]; }
    else { *reinterpret_cast<uint16_t*>(p) = (uint16_t)(w[0] & 0xffffu); }
}
template <int N> __device__ __forceinline__ void load_floats(const float* p, float (&r)[N]) {
    if constexpr (N == 4) { float4 t = *reinterpret_cast<const float4*>(p); r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w; }
    else if constexpr (N == 2) { float2 t = *reinterpret_cast<const float2*>(p); r[0] = t.x; r[1] = t.y; }
    else {
#pragma unroll
        for (int i = 0; i < N; ++i) r[i] = p[i];
    }
}
template <int N> __device__ __forceinline__ void load_stats(const float2* p, float2 (&r)[N]) {
    if constexpr (N % 2 == 0) {
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            float4 t = *reinterpret_cast<const float4*>(p + i);
            r[i] = make_float2(t.x, t.y); r[i + 1] = make_float2(t.z, t.w);
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) r[i] = p[i];
    }
}
template <int N> __device__ __forceinline__ void store_stats(float2* p, const float2 (&r)[N]) {
    if constexpr (N % 2 == 0) {
#pragma unroll
        for (int i = 0; i < N; i += 2) *reinterpret_cast<float4*>(p + i) = make_float4(r[i].x, r[i].y, r[i + 1].x, r[i + 1].y);
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) p[i] = r[i];
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int VB> __device__ __forceinline__ void cp_async_vec(uint32_t smem_dst, const void* gsrc) {
    static_assert(VB == 4 || VB == 8 || VB == 16, "cp.async copies 4, 8 or 16 bytes");
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_dst), "l"(gsrc), "n"(VB) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Stages consecutive rows of one tensor into ring rows.  Thread (colq, head group) copies its own VB-byte
// vector of pixel col_lo + colq and, for the halo (ncols > QW), of pixel col_lo + QW + colq: per row that is
// one LDGSTS (two for the few halo threads) and a pointer bump, no index arithmetic.
template <typename T, int VB, bool HALO> struct RowStager {
    const T* gmain;   // this thread's vector in the next row to stage
    const T* ghalo;
    bool has_main, has_halo;
    __device__ __forceinline__ void init(const T* first_row, int64_t sw, int d, int rj, int col_lo, int ncols, int colq,
                                         int QW, int hg_el) {
        has_main = colq < ncols;
        has_halo = HALO && QW + colq < ncols;
        gmain = first_row + (int64_t)(rj + d * (col_lo + (has_main ? colq : 0))) * sw + hg_el;
        ghalo = first_row + (int64_t)(rj + d * (col_lo + (has_halo ? QW + colq : 0))) * sw + hg_el;
    }
    __device__ __forceinline__ void stage(uint32_t sdst_main, int halo_bytes, int64_t row_stride) {
        if (has_main) cp_async_vec<VB>(sdst_main, gmain);
        gmain += row_stride;
        if (HALO) {
            if (has_halo) cp_async_vec<VB>(sdst_main + halo_bytes, ghalo);
            ghalo += row_stride;
        }
    }
};

// CTA-local mbarrier: arrive (release) / wait on a phase parity (acquire).  Used instead of __syncthreads where a
// thread has independent work between publishing its data and needing everybody else's.
__device__ __forceinline__ void mbar_init(uint32_t mbar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(mbar), "r"(parity) : "memory");
}

// slot of the ring row `delta` rows away from the row held in `slot` (|delta| < RING)
template <int RING> __device__ __forceinline__ int ring_rel(int slot, int delta) {
    int x = slot + delta;
    x -= x >= RING ? RING : 0;
    x += x < 0 ? RING : 0;
    return x;
}

__host__ __device__ __forceinline__ size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// CTAs per SM the forward's register budget is sized for (no spills at these: 80 / 96 / 168 registers)
__host__ __device__ constexpr int stream_fwd_minblocks(int K, int HG) { return HG * K * K > 60 ? 2 : K == 3 ? 4 : 3; }

// ------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------
// Query rows are computed in groups of RS.  The rings hold three groups: the K-1+RS rows being read, the RS rows of
// the next group (landed or landing) and the RS rows being staged for the group after that.  Hand-off is an
// mbarrier: a thread arrives once ITS copies for the next group have landed (after the first row of the current
// group) and waits at the top of the next group, so nobody waits for the slowest warp's compute; the third group of
// ring rows is what makes the early staging safe (every thread has left group g-1 when staging for g+2 starts).
__host__ __device__ constexpr int stream_fwd_rs(int K, int D) { return (K == 3 && D < 8) ? 4 : 2; }
template <int KT, int D> struct StreamFwdCfg {
    static constexpr int RS = stream_fwd_rs(KT, D);
    static constexpr int RING = KT - 1 + 3 * RS;
};
inline size_t stream_fwd_smem_bytes(int K, int D, int heads, int QW) {
    const int R = 2 * K - 1, Cb = heads * D * 2, KW = QW + 2 * (K / 2), RING = K - 1 + 3 * stream_fwd_rs(K, D);
    return align16((size_t)heads * R * R * 4) + align16(2 * (size_t)RING * KW * Cb) + 16;
}
template <int KT, int D, int HG> struct StreamFwdSmem {
    static size_t bytes(int heads, int QW) { return stream_fwd_smem_bytes(KT, D, heads, QW); }
};

template <typename T, int KT, int D, int HG>
__global__ void __launch_bounds__(kStreamThreads, stream_fwd_minblocks(KT, HG))
na2d_stream_fwd_kernel(V5<const T> q, V5<const T> k, V5<const T> v, const float* __restrict__ rpb,
                       V5<T> out, float* __restrict__ lse, NAGeom g, StreamCfg cfg, float scale) {
    constexpr int K = KT, NS = K / 2, KK = K * K, R = 2 * K - 1;
    constexpr int RS = StreamFwdCfg<KT, D>::RS, RING = StreamFwdCfg<KT, D>::RING;
    constexpr int VEC = HG * D, VB = VEC * 2, VW = (VB + 3) / 4, PW = (KK + 1) / 2;
    using M = Mixed<T>;
    extern __shared__ __align__(16) unsigned char stream_smem[];
    unsigned char* const smem = stream_smem;
    const int heads = g.heads, NG = cfg.NG, QW = cfg.QW, Cb = heads * D * 2, KW = QW + 2 * NS;
    const int nb = heads * R * R;
    float* s_rpb = reinterpret_cast<float*>(smem);  // [R][R][heads], divided by scale
    unsigned char* kring = smem + align16((size_t)nb * 4);
    const int ring_row = KW * Cb, ring_bytes = RING * ring_row;
    unsigned char* vring = kring + ring_bytes;
    const int tid = threadIdx.x, nthr = blockDim.x;

    const SubGrid sg = decode_subgrid(g, blockIdx.z);
    const int c0 = blockIdx.x * QW, r0 = blockIdx.y * cfg.RB;
    if (c0 >= sg.Wr || r0 >= sg.Hr) return;
    const int c1 = min(c0 + QW, sg.Wr), r1 = min(r0 + cfg.RB, sg.Hr);
    const float inv_scale = 1.f / scale, c = scale * kLog2e;
    for (int x = tid; x < nb; x += nthr) {
        const int h = x % heads, pp = x / heads;  // pp = pi*R + pj
        s_rpb[x] = rpb != nullptr ? rpb[h * R * R + pp] * inv_scale : 0.f;
    }
    const int kv_lo = axis_window(c0, sg.Wr, K).start;
    const int kv_n = axis_window(c1 - 1, sg.Wr, K).start + K - kv_lo;
    const int colq = tid / NG, hgi = tid - colq * NG, h0 = hgi * HG;
    const int sthr = colq * Cb + hgi * VB;   // this thread's vector inside a ring row (staging)

    // staging state: rows are staged in order, kv_next is the next one, kv_soff its ring-row byte offset
    int kv_next = axis_window(r0, sg.Hr, K).start;
    int kv_soff = (kv_next % RING) * ring_row;
    const int64_t kdh = g.d * k.sh, vdh = g.d * v.sh;
    RowStager<T, VB, true> ks, vs;
    ks.init(k.ptr + sg.b * k.sb + (sg.ri + (int64_t)g.d * kv_next) * k.sh, k.sw, g.d, sg.rj, kv_lo, kv_n, colq, QW, h0 * D);
    vs.init(v.ptr + sg.b * v.sb + (sg.ri + (int64_t)g.d * kv_next) * v.sh, v.sw, g.d, sg.rj, kv_lo, kv_n, colq, QW, h0 * D);
    const uint32_t k_s32 = smem_u32(kring) + sthr, v_s32 = smem_u32(vring) + sthr;
    auto stage_until = [&](int last_query_row) {   // everything query rows <= last_query_row read
        const int need = axis_window(last_query_row, sg.Hr, K).start + K;
#pragma unroll 1
        while (kv_next < need) {
            ks.stage(k_s32 + kv_soff, QW * Cb, kdh);
            vs.stage(v_s32 + kv_soff, QW * Cb, vdh);
            kv_soff += ring_row;
            kv_soff = kv_soff == ring_bytes ? 0 : kv_soff;
            ++kv_next;
        }
    };
    const uint32_t mbar = smem_u32(vring + align16((size_t)ring_bytes * 2) - ring_bytes);   // after both rings
    if (tid == 0) mbar_init(mbar, nthr);
    __syncthreads();                          // barrier initialised, rpb table visible
    stage_until(min(r0 + RS, r1) - 1);
    cp_async_commit();
    cp_async_wait_all();
    mbar_arrive(mbar);                        // phase 0: the first group's rows
    uint32_t parity = 0;

    const int cq = c0 + colq;
    const bool active = cq < c1;
    const AxisWin wj = axis_window(active ? cq : c0, sg.Wr, K);
    const int kthr = (wj.start - kv_lo) * Cb + hgi * VB;
    const int j = sg.rj + g.d * cq;
    const T* qp = q.ptr + sg.b * q.sb + (sg.ri + (int64_t)g.d * r0) * q.sh + (int64_t)j * q.sw + h0 * D;
    T* op = out.ptr + sg.b * out.sb + (sg.ri + (int64_t)g.d * r0) * out.sh + (int64_t)j * out.sw + h0 * D;
    const int64_t qdh = g.d * q.sh, odh = g.d * out.sh;
    const float* rp_col = s_rpb + (wj.pb * heads + h0);
    uint32_t qnext[VW];
    if (active) load_words<VB, VW>(qp, qnext);

#pragma unroll 1
    for (int t0 = r0; t0 < r1; t0 += RS) {
        const int t1 = min(t0 + RS, r1);
        mbar_wait(mbar, parity);                          // every thread's copies for this group have landed
        parity ^= 1u;
        if (t1 < r1) stage_until(min(t1 + RS, r1) - 1);   // next group's rows land while this group computes
        cp_async_commit();
        if (!active) {
            cp_async_wait_all();
            mbar_arrive(mbar);
            continue;
        }
#pragma unroll 1
        for (int t = t0; t < t1; ++t, qp += qdh, op += odh) {
            uint32_t qw[VW];
#pragma unroll
            for (int x = 0; x < VW; ++x) qw[x] = qnext[x];
            if (t + 1 < r1) load_words<VB, VW>(qp + qdh, qnext);   // next row's query: latency hidden by this row
            const AxisWin wi = axis_window(t, sg.Hr, K);
            const float* rp = rp_col + wi.pb * R * heads;
            const int slot0 = wi.start % RING;

            float s[HG][KK], mx[HG];
#pragma unroll
            for (int hg = 0; hg < HG; ++hg) mx[hg] = -INFINITY;
#pragma unroll
            for (int mi = 0; mi < K; ++mi) {
                const unsigned char* krow = kring + ring_rel<RING>(slot0, mi) * ring_row + kthr;
#pragma unroll
                for (int mj = 0; mj < K; ++mj) {
                    uint32_t kw[VW];
                    load_words<VB, VW>(krow + mj * Cb, kw);
                    float rb[HG];
                    load_floats<HG>(rp + (mi * R + mj) * heads, rb);
#pragma unroll
                    for (int hg = 0; hg < HG; ++hg) {
                        float a = rb[hg];
#pragma unroll
                        for (int e = 0; e < D; ++e) a = M::fma(elem_of<VW>(qw, hg * D + e), elem_of<VW>(kw, hg * D + e), a);
                        s[hg][mi * K + mj] = a;
                        mx[hg] = fmaxf(mx[hg], a);
                    }
                }
            }
            uint32_t pw[HG][PW];
            float inv[HG], lse2[HG];
#pragma unroll
            for (int hg = 0; hg < HG; ++hg) {
                const float mc = mx[hg] * c;
                float den = 0.f;
#pragma unroll
                for (int n = 0; n < KK; ++n) {
                    const float p = fast_exp2(fmaf(s[hg][n], c, -mc));
                    s[hg][n] = p;
                    den += p;
                }
                inv[hg] = __fdividef(1.f, den);
                lse2[hg] = mc + __log2f(den);
#pragma unroll
                for (int n2 = 0; n2 < PW; ++n2) pw[hg][n2] = M::pack(s[hg][2 * n2], 2 * n2 + 1 < KK ? s[hg][2 * n2 + 1] : 0.f);
            }
            float acc[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
#pragma unroll
            for (int mi = 0; mi < K; ++mi) {
                const unsigned char* vrow = vring + ring_rel<RING>(slot0, mi) * ring_row + kthr;
#pragma unroll
                for (int mj = 0; mj < K; ++mj) {
                    uint32_t vw[VW];
                    load_words<VB, VW>(vrow + mj * Cb, vw);
                    const int n = mi * K + mj;
#pragma unroll
                    for (int hg = 0; hg < HG; ++hg)
#pragma unroll
                        for (int e = 0; e < D; ++e)
                            acc[hg * D + e] = M::fma(half_of(pw[hg][n >> 1], n & 1), elem_of<VW>(vw, hg * D + e), acc[hg * D + e]);
                }
            }
            uint32_t ow[VW];
#pragma unroll
            for (int x = 0; x < VW; ++x) {
                const int e0 = 2 * x, e1 = 2 * x + 1;
                ow[x] = M::pack(acc[e0] * inv[e0 / D], e1 < VEC ? acc[e1] * inv[e1 / D] : 0.f);
            }
            store_words<VB, VW>(op, ow);
            if (lse != nullptr) {
                const int i = sg.ri + g.d * t;
#pragma unroll
                for (int hg = 0; hg < HG; ++hg)
                    lse[(((int64_t)sg.b * g.H + i) * g.W + j) * heads + h0 + hg] = lse2[hg] * kLn2;
            }
            if (t == t0) {                    // one row of compute later this thread's copies have landed
                cp_async_wait_all();
                mbar_arrive(mbar);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// backward (one kernel)
// ------------------------------------------------------------------------------------------------------
// Stripe / band boundaries: s*T, except that no boundary may fall inside the last K-1 positions of the axis
// (there the inverse neighbourhood reaches 2*(K/2) instead of K/2 past the boundary); such a boundary moves to L-K.
__host__ __device__ __forceinline__ int stream_boundary(int s, int T, int L, int K) {
    const int64_t b = (int64_t)s * T;
    if (b >= L) return L;
    if (L - b < K) return L - K;
    return (int)b;
}

inline size_t stream_bwd_smem_bytes(int K, int D, int HG, int heads, int QW) {
    const int R = 2 * K - 1, Cb = heads * D * 2, KW = QW + 2 * (K / 2), nthr = QW * (heads / HG), RING = K + K / 2 + 3;
    return 2 * align16((size_t)heads * R * R * 4) + 2 * (size_t)RING * KW * Cb + 2 * (size_t)RING * QW * Cb +
           (size_t)RING * QW * heads * sizeof(float2) + align16((size_t)nthr * K * K * 4) + align16((size_t)QW * 4) + 16;
}
template <int KT, int D, int HG> struct StreamBwdSmem {
    // rows per ring: K + K/2 live rows for phase B, the row of phase A, the row being prefetched, and one row of
    // slack because warps may run one phase apart (mbarrier instead of a CTA-wide barrier, see the kernel)
    static constexpr int RING = KT + KT / 2 + 3;
    // CTAs per SM the register budget is sized for: scores + dP + drpb sums of 2-4 heads need ~165 registers
    // (2 CTAs); one head per thread fits 96 registers without spills (3 CTAs) unless shared memory (head dim 8)
    // allows only 2
    static constexpr int MINB = (HG * KT * KT > 9 || D >= 8) ? 2 : 3;
    static size_t bytes(int heads, int QW) { return stream_bwd_smem_bytes(KT, D, HG, heads, QW); }
};

template <typename T, int KT, int D, int HG>
__global__ void __launch_bounds__(kStreamThreads, StreamBwdSmem<KT, D, HG>::MINB)
na2d_stream_bwd_kernel(V5<const T> q, V5<const T> k, V5<const T> v, V5<const T> dout, const float* __restrict__ rpb,
                       V5<T> dq, V5<T> dk, V5<T> dv, float* __restrict__ drpb_part, NAGeom g, StreamCfg cfg,
                       float scale) {
    constexpr int K = KT, NS = K / 2, KK = K * K, R = 2 * K - 1, RING = StreamBwdSmem<KT, D, HG>::RING;
    constexpr int VEC = HG * D, VB = VEC * 2, VW = (VB + 3) / 4, PW = (KK + 1) / 2;
    using M = Mixed<T>;
    extern __shared__ __align__(16) unsigned char stream_smem[];
    unsigned char* const smem = stream_smem;
    const int heads = g.heads, NG = cfg.NG, QW = cfg.QW, Cb = heads * D * 2, KW = QW + 2 * NS, TW = QW - 2 * NS;
    const int nb = heads * R * R;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int kv_row = KW * Cb, q_row = QW * Cb, kv_bytes = RING * kv_row;
    float* s_rpb = reinterpret_cast<float*>(smem);                               // [R][R][heads], divided by scale
    float* s_acc = reinterpret_cast<float*>(smem + align16((size_t)nb * 4));     // [R][R][heads] drpb partial sums
    unsigned char* kring = smem + 2 * align16((size_t)nb * 4);
    unsigned char* vring = kring + kv_bytes;
    unsigned char* qring = vring + kv_bytes;
    unsigned char* gring = qring + RING * q_row;
    float2* sring = reinterpret_cast<float2*>(gring + RING * q_row);             // [RING][QW][heads] (lse2, delta)
    float* s_tbl = reinterpret_cast<float*>(sring + (size_t)RING * QW * heads);  // [nthr][KK] drpb flush table (one head per pass)
    int* s_pbj = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(s_tbl) + align16((size_t)nthr * KK * 4));  // [QW] column rpb offset
    const uint32_t mbar = smem_u32(reinterpret_cast<unsigned char*>(s_pbj) + align16((size_t)QW * 4));              // row hand-off barrier
    const int64_t cta = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;

    const SubGrid sg = decode_subgrid(g, blockIdx.z);
    const int c0 = stream_boundary(blockIdx.x, TW, sg.Wr, K), c1 = stream_boundary(blockIdx.x + 1, TW, sg.Wr, K);
    const int r0 = stream_boundary(blockIdx.y, cfg.RB, sg.Hr, K), r1 = stream_boundary(blockIdx.y + 1, cfg.RB, sg.Hr, K);
    if (c0 >= c1 || r0 >= r1) {
        if (drpb_part != nullptr)
            for (int x = tid; x < nb; x += nthr) drpb_part[cta * nb + x] = 0.f;
        return;
    }
    const float inv_scale = 1.f / scale, c = scale * kLog2e;
    for (int x = tid; x < nb; x += nthr) {
        const int h = x % heads, pp = x / heads;
        s_rpb[x] = rpb != nullptr ? rpb[h * R * R + pp] * inv_scale : 0.f;
        s_acc[x] = 0.f;
    }
    // queries this CTA recomputes: every query attending a key of the stripe / band
    int ql, qh, tq_lo, tq_hi, dummy;
    inverse_window(c0, sg.Wr, K, ql, dummy);
    inverse_window(c1 - 1, sg.Wr, K, dummy, qh);
    inverse_window(r0, sg.Hr, K, tq_lo, dummy);
    inverse_window(r1 - 1, sg.Hr, K, dummy, tq_hi);
    const int qn = qh - ql + 1;                                     // <= QW
    const int kv_lo = axis_window(ql, sg.Wr, K).start;
    const int kv_n = axis_window(qh, sg.Wr, K).start + K - kv_lo;   // <= KW

    const int colq = tid / NG, hgi = tid - colq * NG, h0 = hgi * HG;
    const int qthr = colq * Cb + hgi * VB;   // this thread's vector inside a ring row
    int kv_next = axis_window(tq_lo, sg.Hr, K).start;
    int kv_soff = (kv_next % RING) * kv_row;
    int slot_t = tq_lo % RING;               // ring slot of query row t (q, dout and statistics rings)
    const int64_t kdh = g.d * k.sh, vdh = g.d * v.sh, qdh = g.d * q.sh, gdh = g.d * dout.sh;
    RowStager<T, VB, true> ks, vs;
    RowStager<T, VB, false> qs, gs;
    ks.init(k.ptr + sg.b * k.sb + (sg.ri + (int64_t)g.d * kv_next) * k.sh, k.sw, g.d, sg.rj, kv_lo, kv_n, colq, QW, h0 * D);
    vs.init(v.ptr + sg.b * v.sb + (sg.ri + (int64_t)g.d * kv_next) * v.sh, v.sw, g.d, sg.rj, kv_lo, kv_n, colq, QW, h0 * D);
    qs.init(q.ptr + sg.b * q.sb + (sg.ri + (int64_t)g.d * tq_lo) * q.sh, q.sw, g.d, sg.rj, ql, qn, colq, QW, h0 * D);
    gs.init(dout.ptr + sg.b * dout.sb + (sg.ri + (int64_t)g.d * tq_lo) * dout.sh, dout.sw, g.d, sg.rj, ql, qn, colq, QW, h0 * D);
    const uint32_t k_s32 = smem_u32(kring) + qthr, v_s32 = smem_u32(vring) + qthr;
    const uint32_t q_s32 = smem_u32(qring) + qthr, g_s32 = smem_u32(gring) + qthr;
    auto stage_kv = [&]() {
        ks.stage(k_s32 + kv_soff, QW * Cb, kdh);
        vs.stage(v_s32 + kv_soff, QW * Cb, vdh);
        kv_soff += kv_row;
        kv_soff = kv_soff == kv_bytes ? 0 : kv_soff;
        ++kv_next;
    };
    auto stage_q = [&](int slot) {
        qs.stage(q_s32 + slot * q_row, 0, qdh);
        gs.stage(g_s32 + slot * q_row, 0, gdh);
    };
    stage_q(slot_t);
#pragma unroll 1
    for (int x = 0; x < K; ++x) stage_kv();
    cp_async_commit();

    // phase A role: one query column (own stripe + halo)
    const int cq = ql + colq;
    const bool a_active = cq <= qh;
    const bool a_own_col = cq >= c0 && cq < c1;
    const AxisWin wjA = axis_window(a_active ? cq : ql, sg.Wr, K);
    const int kthrA = (wjA.start - kv_lo) * Cb + hgi * VB;
    const int sthr = colq * heads + h0;
    const float* rp_col = s_rpb + (wjA.pb * heads + h0);
    T* dqcol = dq.ptr + sg.b * dq.sb + (int64_t)(sg.rj + g.d * cq) * dq.sw + h0 * D;
    // phase B role: the same column as a key column, when it lies inside the stripe
    const int ck = cq;
    const bool b_active = a_own_col;
    int lo_j, hi_j;
    inverse_window(b_active ? ck : c0, sg.Wr, K, lo_j, hi_j);
    const bool b_fast_col = (lo_j == ck - NS) && (hi_j == ck + NS);
    const int kthrB = (ck - kv_lo) * Cb + hgi * VB;
    T* dkcol = dk.ptr + sg.b * dk.sb + (int64_t)(sg.rj + g.d * ck) * dk.sw + h0 * D;
    T* dvcol = dv.ptr + sg.b * dv.sb + (int64_t)(sg.rj + g.d * ck) * dv.sw + h0 * D;

    // drpb: per-thread register sums over the rows that share a row offset (cur_pb), flushed through a
    // shared-memory table and summed per bin in a fixed order (no atomics)
    float racc[HG][KK];
#pragma unroll
    for (int hg = 0; hg < HG; ++hg)
#pragma unroll
        for (int n = 0; n < KK; ++n) racc[hg][n] = 0.f;
    int cur_pb = -1;   // uniform across the CTA
    if (hgi == 0) s_pbj[colq] = wjA.pb;
    auto flush_racc = [&]() {   // called by every thread of the CTA
#pragma unroll
        for (int hg = 0; hg < HG; ++hg) {
#pragma unroll
            for (int n = 0; n < KK; ++n) {
                s_tbl[tid * KK + n] = racc[hg][n];
                racc[hg][n] = 0.f;
            }
            __syncthreads();
            for (int x = tid; x < nb; x += nthr) {
                const int h = x % heads, pp = x / heads, pi = pp / R, pj = pp - pi * R;
                const int mi = pi - cur_pb, hgx = h / HG;
                if (mi < 0 || mi >= K || h - hgx * HG != hg) continue;
                float sum = 0.f;
                for (int cc = 0; cc < QW; ++cc) {
                    const int mj = pj - s_pbj[cc];
                    if (mj >= 0 && mj < K) sum += s_tbl[(cc * NG + hgx) * KK + mi * K + mj];
                }
                s_acc[x] += sum;
            }
            __syncthreads();
        }
    };

    // Row hand-off.  Instead of a CTA-wide barrier per row (30 % of the stall samples when it was one), every
    // thread ARRIVES on an mbarrier once its phase A of row t is done and its copies for row t+1 have landed,
    // runs phase B of row t-1 (which only reads what earlier phases published) and WAITS just before phase A of
    // row t+1: warps may drift one phase apart, which the extra ring row makes safe.
    if (tid == 0) mbar_init(mbar, nthr);
    __syncthreads();
    auto row_arrive = [&]() { mbar_arrive(mbar); };
    auto row_wait = [&](uint32_t parity) { mbar_wait(mbar, parity); };
    cp_async_wait_all();
    row_arrive();
    uint32_t parity = 0;

#pragma unroll 1
    for (int t = tq_lo; t <= tq_hi + 1; ++t) {
        row_wait(parity);
        parity ^= 1u;
        if (t + 1 <= tq_hi) {
            stage_q(ring_rel<RING>(slot_t, 1));
            if (kv_next < axis_window(t + 1, sg.Hr, K).start + K) stage_kv();
        }
        cp_async_commit();

        // ---------------- phase A: query row t ----------------
        if (t <= tq_hi) {
            const AxisWin wi = axis_window(t, sg.Hr, K);
            const bool own_row = t >= r0 && t < r1;
            if (drpb_part != nullptr && own_row && wi.pb != cur_pb) {   // uniform
                if (cur_pb >= 0) flush_racc();
                cur_pb = wi.pb;
            }
            if (a_active) {
                uint32_t qw[VW], gw[VW];
                load_words<VB, VW>(qring + slot_t * q_row + qthr, qw);
                load_words<VB, VW>(gring + slot_t * q_row + qthr, gw);
                const float* rp = rp_col + wi.pb * R * heads;
                const int slot0 = ring_rel<RING>(slot_t, wi.start - t);
                float s[HG][KK], dp[HG][KK], mx[HG];
#pragma unroll
                for (int hg = 0; hg < HG; ++hg) mx[hg] = -INFINITY;
#pragma unroll
                for (int mi = 0; mi < K; ++mi) {
                    const int ro = ring_rel<RING>(slot0, mi) * kv_row + kthrA;
#pragma unroll
                    for (int mj = 0; mj < K; ++mj) {
                        uint32_t kw[VW], vw[VW];
                        load_words<VB, VW>(kring + ro + mj * Cb, kw);
                        load_words<VB, VW>(vring + ro + mj * Cb, vw);
                        float rb[HG];
                        load_floats<HG>(rp + (mi * R + mj) * heads, rb);
#pragma unroll
                        for (int hg = 0; hg < HG; ++hg) {
                            float a = rb[hg], b = 0.f;
#pragma unroll
                            for (int e = 0; e < D; ++e) {
                                a = M::fma(elem_of<VW>(qw, hg * D + e), elem_of<VW>(kw, hg * D + e), a);
                                b = M::fma(elem_of<VW>(gw, hg * D + e), elem_of<VW>(vw, hg * D + e), b);
                            }
                            s[hg][mi * K + mj] = a;
                            dp[hg][mi * K + mj] = b;
                            mx[hg] = fmaxf(mx[hg], a);
                        }
                    }
                }
                float2 st[HG];
                float inv[HG];
#pragma unroll
                for (int hg = 0; hg < HG; ++hg) {
                    const float mc = mx[hg] * c;
                    float den = 0.f, dot = 0.f;
#pragma unroll
                    for (int n = 0; n < KK; ++n) {
                        const float p = fast_exp2(fmaf(s[hg][n], c, -mc));
                        s[hg][n] = p;
                        den += p;
                        dot = fmaf(p, dp[hg][n], dot);
                    }
                    inv[hg] = __fdividef(1.f, den);
                    const float delta = dot * inv[hg];
#pragma unroll
                    for (int n = 0; n < KK; ++n) s[hg][n] *= dp[hg][n] - delta;   // dS * den (normalised below)
                    st[hg] = make_float2(mc + __log2f(den), delta);
                }
                store_stats<HG>(sring + (size_t)slot_t * QW * heads + sthr, st);

                if (a_own_col && own_row) {
                    if (drpb_part != nullptr) {
#pragma unroll
                        for (int hg = 0; hg < HG; ++hg)
#pragma unroll
                            for (int n = 0; n < KK; ++n) racc[hg][n] = fmaf(s[hg][n], inv[hg], racc[hg][n]);
                    }
                    uint32_t dsw[HG][PW];
#pragma unroll
                    for (int hg = 0; hg < HG; ++hg)
#pragma unroll
                        for (int n2 = 0; n2 < PW; ++n2)
                            dsw[hg][n2] = M::pack(s[hg][2 * n2], 2 * n2 + 1 < KK ? s[hg][2 * n2 + 1] : 0.f);
                    float acc[VEC];
#pragma unroll
                    for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
#pragma unroll
                    for (int mi = 0; mi < K; ++mi) {
                        const int ro = ring_rel<RING>(slot0, mi) * kv_row + kthrA;
#pragma unroll
                        for (int mj = 0; mj < K; ++mj) {
                            uint32_t kw[VW];
                            load_words<VB, VW>(kring + ro + mj * Cb, kw);
                            const int n = mi * K + mj;
#pragma unroll
                            for (int hg = 0; hg < HG; ++hg)
#pragma unroll
                                for (int e = 0; e < D; ++e)
                                    acc[hg * D + e] = M::fma(half_of(dsw[hg][n >> 1], n & 1), elem_of<VW>(kw, hg * D + e), acc[hg * D + e]);
                        }
                    }
                    uint32_t ow[VW];
#pragma unroll
                    for (int x = 0; x < VW; ++x) {
                        const int e0 = 2 * x, e1 = 2 * x + 1;
                        ow[x] = M::pack(acc[e0] * (scale * inv[e0 / D]), e1 < VEC ? acc[e1] * (scale * inv[e1 / D]) : 0.f);
                    }
                    store_words<VB, VW>(dqcol + (int64_t)(sg.ri + g.d * t) * dq.sh, ow);
                }
            }
        }

        cp_async_wait_all();
        row_arrive();

        // ---------------- phase B: key rows whose last attending query row is t-1 ----------------
        const int tb = t - 1;
        if (tb >= tq_lo) {
            int rlo, rhi;
            if (tb == sg.Hr - 1) { rlo = sg.Hr - K; rhi = sg.Hr - 1; }
            else { rlo = rhi = tb - NS; if (rlo < 0 || rlo >= sg.Hr - K) { rlo = 0; rhi = -1; } }
            rlo = max(rlo, r0);
            rhi = min(rhi, r1 - 1);
#pragma unroll 1
            for (int r = rlo; r <= rhi; ++r) {
                if (!b_active) continue;
                int lo_i, hi_i;
                inverse_window(r, sg.Hr, K, lo_i, hi_i);
                const int slot_r = ring_rel<RING>(slot_t, r - t);   // ring slot of row r
                uint32_t kw[VW], vw[VW];
                {
                    const int ro = slot_r * kv_row + kthrB;   // all rings hold row r in slot r % RING
                    load_words<VB, VW>(kring + ro, kw);
                    load_words<VB, VW>(vring + ro, vw);
                }
                float ak[VEC], av[VEC];
#pragma unroll
                for (int e = 0; e < VEC; ++e) ak[e] = av[e] = 0.f;
                auto contrib = [&](int slot, int ucol, int pbi, int pbj) {
                    uint32_t qw[VW], gw[VW];
                    const int qo = slot * q_row + ucol * Cb + hgi * VB;
                    load_words<VB, VW>(qring + qo, qw);
                    load_words<VB, VW>(gring + qo, gw);
                    float2 st[HG];
                    load_stats<HG>(sring + ((size_t)slot * QW + ucol) * heads + h0, st);
                    float rb[HG];
                    load_floats<HG>(s_rpb + (pbi * R + pbj) * heads + h0, rb);
#pragma unroll
                    for (int hg = 0; hg < HG; ++hg) {
                        float a = rb[hg], b = 0.f;
#pragma unroll
                        for (int e = 0; e < D; ++e) {
                            a = M::fma(elem_of<VW>(qw, hg * D + e), elem_of<VW>(kw, hg * D + e), a);
                            b = M::fma(elem_of<VW>(gw, hg * D + e), elem_of<VW>(vw, hg * D + e), b);
                        }
                        const float p = fast_exp2(fmaf(a, c, -st[hg].x));
                        const float ds = p * (b - st[hg].y);
                        const uint32_t w = M::pack(p, ds);
#pragma unroll
                        for (int e = 0; e < D; ++e) {
                            ak[hg * D + e] = M::fma(half_of(w, 1), elem_of<VW>(qw, hg * D + e), ak[hg * D + e]);
                            av[hg * D + e] = M::fma(half_of(w, 0), elem_of<VW>(gw, hg * D + e), av[hg * D + e]);
                        }
                    }
                };
                if (b_fast_col && lo_i == r - NS && hi_i == r + NS) {
#pragma unroll
                    for (int a = -NS; a <= NS; ++a) {
                        const int slot = ring_rel<RING>(slot_r, a);
#pragma unroll
                        for (int b = -NS; b <= NS; ++b) contrib(slot, colq + b, K - 1 - a, K - 1 - b);
                    }
                } else {
#pragma unroll 1
                    for (int tq = lo_i; tq <= hi_i; ++tq) {
                        const int slot = ring_rel<RING>(slot_r, tq - r);
#pragma unroll 1
                        for (int u = lo_j; u <= hi_j; ++u) contrib(slot, u - ql, r - tq + K - 1, ck - u + K - 1);
                    }
                }
                uint32_t kwo[VW], vwo[VW];
#pragma unroll
                for (int x = 0; x < VW; ++x) {
                    kwo[x] = M::pack(ak[2 * x] * scale, 2 * x + 1 < VEC ? ak[2 * x + 1] * scale : 0.f);
                    vwo[x] = M::pack(av[2 * x], 2 * x + 1 < VEC ? av[2 * x + 1] : 0.f);
                }
                const int64_t i = sg.ri + g.d * r;
                store_words<VB, VW>(dkcol + i * dk.sh, kwo);
                store_words<VB, VW>(dvcol + i * dv.sh, vwo);
            }
        }
        slot_t = ring_rel<RING>(slot_t, 1);
    }
    if (drpb_part != nullptr) {
        if (cur_pb >= 0) flush_racc();
        __syncthreads();
        for (int x = tid; x < nb; x += nthr) {
            const int h = x % heads, pp = x / heads;
            drpb_part[cta * nb + h * R * R + pp] = s_acc[x];
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
// entry points implemented in na2d_stream.cu (one per 16-bit type); LMNET_ERR_UNSUPPORTED = "use na2d_fused"
int stream_fwd(const FusedArgs& a, int dtype);
int stream_bwd(const FusedArgs& a, int dtype, int64_t max_parts, int64_t* n_parts);

}  // namespace lmnet
