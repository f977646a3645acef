// na2d_fused.cu — fused 2-D neighbourhood attention, forward and backward (sm_100a).
//
// Replaces the q*scale -> na2d_qk(+rpb) -> softmax -> na2d_av chain that natten's
// NeighborhoodAttention2D runs for /root/reference/core/modules.py:517, and its autograd.
// The attention map [B,heads,H,W,K*K] is never written: the forward keeps the K*K scores of a
// (pixel, head) in registers, the backward recomputes them.
//
// Mapping.  LM-Net's head dims are 1..8 with 12 heads, so a pixel's channels are 24..192
// contiguous bytes; the kernels are bandwidth/issue bound, not tensor-core shaped.  One thread
// owns HG consecutive heads of one pixel (a 4..16 byte vector), consecutive lanes walk
// (head-group, column) in memory order so every warp access is contiguous.  Dilation d is
// handled as d*d independent undilated problems on interleaved sub-grids (blockIdx.z).
//
// Backward = two kernels:
//   A (per query): recompute P, dP; delta = sum P*dP; dS = P*(dP-delta); dq; per-CTA drpb
//                  partial sums; writes (lse, delta) per (pixel, head) to the workspace.
//   B (per key):   gathers over the inverse neighbourhood with the stored (lse, delta):
//                  dk = sum dS*q, dv = sum P*dout.   No atomics on dq/dk/dv: deterministic.
//   drpb partials are reduced by a third tiny kernel in a fixed order.
#pragma once
#include "common.cuh"

namespace lmnet {

template <typename T>
struct V5 {
    T* ptr;
    int64_t sb, sh, sw, sn;
};

struct NAGeom {
    int B, H, W, heads, D, K, d;
    int Hmax, Wmax;  // ceil(H/d), ceil(W/d): extent of the largest sub-grid
};

struct SubGrid {
    int b, ri, rj, Hr, Wr;
};
__device__ __forceinline__ SubGrid decode_subgrid(const NAGeom& g, int z) {
    SubGrid s;
    int dd = g.d * g.d;
    s.b = z / dd;
    int r = z - s.b * dd;
    s.ri = r / g.d;
    s.rj = r - s.ri * g.d;
    s.Hr = (g.H - s.ri + g.d - 1) / g.d;
    s.Wr = (g.W - s.rj + g.d - 1) / g.d;
    return s;
}

constexpr int kThreads = 128;
constexpr int kRowChunk = 16;  // rows walked by one thread of backward kernel A
constexpr int kMaxKK = 169;    // K <= 13

// KT > 0: compile-time kernel size; KT == 0: runtime kernel size (scores live in local memory).
template <int KT> struct KSize {
    __device__ __forceinline__ static int get(int) { return KT; }
    static constexpr int kk_cap = KT * KT;
};
template <> struct KSize<0> {
    __device__ __forceinline__ static int get(int k) { return k; }
    static constexpr int kk_cap = kMaxKK;
};

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <typename T, int KT, int D, int HG>
__global__ void __launch_bounds__(kThreads)
na2d_fwd_kernel(V5<const T> q, V5<const T> k, V5<const T> v, const float* __restrict__ rpb,
                V5<T> out, float* __restrict__ lse, NAGeom g, float scale_log2e) {
    constexpr int VEC = HG * D;
    constexpr int CAP = KSize<KT>::kk_cap;
    const int K = KSize<KT>::get(g.K);
    const int R = 2 * K - 1;
    extern __shared__ float s_rpb[];
    if (rpb != nullptr) {
        for (int i = threadIdx.x; i < g.heads * R * R; i += kThreads) s_rpb[i] = rpb[i] * kLog2e;
        __syncthreads();
    }
    const SubGrid sg = decode_subgrid(g, blockIdx.z);
    const int NG = g.heads / HG;
    const int64_t gid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (gid >= (int64_t)sg.Hr * sg.Wr * NG) return;
    const int grp = (int)(gid % NG);
    const int pix = (int)(gid / NG);
    const int tj = pix % sg.Wr;
    const int ti = pix / sg.Wr;
    const int h0 = grp * HG;
    const AxisWin wi = axis_window(ti, sg.Hr, K);
    const AxisWin wj = axis_window(tj, sg.Wr, K);
    const int i = sg.ri + g.d * ti, j = sg.rj + g.d * tj;

    float qv[VEC];
    load_f<VEC, T, true>(q.ptr + sg.b * q.sb + i * q.sh + j * q.sw + h0 * q.sn, qv);
#pragma unroll
    for (int e = 0; e < VEC; ++e) qv[e] *= scale_log2e;

    const T* kbase = k.ptr + sg.b * k.sb + (sg.ri + g.d * wi.start) * k.sh + (sg.rj + g.d * wj.start) * k.sw + h0 * k.sn;
    const T* vbase = v.ptr + sg.b * v.sb + (sg.ri + g.d * wi.start) * v.sh + (sg.rj + g.d * wj.start) * v.sw + h0 * v.sn;
    const int64_t kdh = g.d * k.sh, kdw = g.d * k.sw, vdh = g.d * v.sh, vdw = g.d * v.sw;

    float s[HG][CAP];
    float mx[HG];
#pragma unroll
    for (int hg = 0; hg < HG; ++hg) mx[hg] = -INFINITY;
#pragma unroll
    for (int mi = 0; mi < K; ++mi) {
#pragma unroll
        for (int mj = 0; mj < K; ++mj) {
            float kv[VEC];
            load_f<VEC, T, true>(kbase + mi * kdh + mj * kdw, kv);
#pragma unroll
            for (int hg = 0; hg < HG; ++hg) {
                float a = 0.f;
#pragma unroll
                for (int e = 0; e < D; ++e) a = fmaf(qv[hg * D + e], kv[hg * D + e], a);
                if (rpb != nullptr) a += s_rpb[((h0 + hg) * R + wi.pb + mi) * R + wj.pb + mj];
                s[hg][mi * K + mj] = a;
                mx[hg] = fmaxf(mx[hg], a);
            }
        }
    }
    float den[HG];
#pragma unroll
    for (int hg = 0; hg < HG; ++hg) {
        den[hg] = 0.f;
#pragma unroll
        for (int n = 0; n < K * K; ++n) {
            float p = fast_exp2(s[hg][n] - mx[hg]);
            s[hg][n] = p;
            den[hg] += p;
        }
    }
    float acc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
#pragma unroll
    for (int mi = 0; mi < K; ++mi) {
#pragma unroll
        for (int mj = 0; mj < K; ++mj) {
            float vv[VEC];
            load_f<VEC, T, true>(vbase + mi * vdh + mj * vdw, vv);
#pragma unroll
            for (int hg = 0; hg < HG; ++hg)
#pragma unroll
                for (int e = 0; e < D; ++e) acc[hg * D + e] = fmaf(s[hg][mi * K + mj], vv[hg * D + e], acc[hg * D + e]);
        }
    }
#pragma unroll
    for (int hg = 0; hg < HG; ++hg) {
        float inv = 1.f / den[hg];
#pragma unroll
        for (int e = 0; e < D; ++e) acc[hg * D + e] *= inv;
    }
    store_f<VEC, T, true>(out.ptr + sg.b * out.sb + i * out.sh + j * out.sw + h0 * out.sn, acc);
    if (lse != nullptr) {
#pragma unroll
        for (int hg = 0; hg < HG; ++hg)
            lse[(((int64_t)sg.b * g.H + i) * g.W + j) * g.heads + h0 + hg] = (mx[hg] + log2f(den[hg])) * kLn2;
    }
}

// ------------------------------------------------------------------------------------------------
// backward kernel A: per query
// ------------------------------------------------------------------------------------------------
template <typename T, int KT, int D, int HG>
__global__ void __launch_bounds__(kThreads)
na2d_bwd_query_kernel(V5<const T> q, V5<const T> k, V5<const T> v, V5<const T> dout,
                      const float* __restrict__ rpb, V5<T> dq, float2* __restrict__ stats,
                      float* __restrict__ drpb_part, NAGeom g, float scale) {
    constexpr int VEC = HG * D;
    constexpr int CAP = KSize<KT>::kk_cap;
    const int K = KSize<KT>::get(g.K);
    const int R = 2 * K - 1;
    const int nbins = g.heads * R * R;
    extern __shared__ float smem[];
    float* s_rpb = smem;          // [heads*R*R], pre-multiplied by log2e
    float* s_acc = smem + nbins;  // [heads*R*R], drpb partial sums of this CTA
    for (int x = threadIdx.x; x < nbins; x += kThreads) {
        s_rpb[x] = rpb != nullptr ? rpb[x] * kLog2e : 0.f;
        s_acc[x] = 0.f;
    }
    __syncthreads();

    const SubGrid sg = decode_subgrid(g, blockIdx.z);
    const int NG = g.heads / HG;
    const int gid = blockIdx.x * kThreads + threadIdx.x;
    const int grp = gid % NG;
    const int tj = gid / NG;
    const int h0 = grp * HG;
    const bool valid = tj < sg.Wr;
    const int row0 = blockIdx.y * kRowChunk;
    const int row1 = min(row0 + kRowChunk, sg.Hr);
    const float scale_log2e = scale * kLog2e;

    if (valid && row0 < row1) {
        const AxisWin wj = axis_window(tj, sg.Wr, K);
        const int j = sg.rj + g.d * tj;
        float racc[HG][CAP];
        int cur_pb = -1;
        for (int ti = row0; ti < row1; ++ti) {
            const AxisWin wi = axis_window(ti, sg.Hr, K);
            const int i = sg.ri + g.d * ti;
            if (drpb_part != nullptr && wi.pb != cur_pb) {
                if (cur_pb >= 0) {
#pragma unroll
                    for (int hg = 0; hg < HG; ++hg)
#pragma unroll
                        for (int mi = 0; mi < K; ++mi)
#pragma unroll
                            for (int mj = 0; mj < K; ++mj)
                                atomicAdd(&s_acc[((h0 + hg) * R + cur_pb + mi) * R + wj.pb + mj], racc[hg][mi * K + mj]);
                }
#pragma unroll
                for (int hg = 0; hg < HG; ++hg)
#pragma unroll
                    for (int n = 0; n < K * K; ++n) racc[hg][n] = 0.f;
                cur_pb = wi.pb;
            }
            float qv[VEC], gv[VEC];
            load_f<VEC, T, true>(q.ptr + sg.b * q.sb + i * q.sh + j * q.sw + h0 * q.sn, qv);
            load_f<VEC, T, true>(dout.ptr + sg.b * dout.sb + i * dout.sh + j * dout.sw + h0 * dout.sn, gv);
            const T* kbase = k.ptr + sg.b * k.sb + (sg.ri + g.d * wi.start) * k.sh + (sg.rj + g.d * wj.start) * k.sw + h0 * k.sn;
            const T* vbase = v.ptr + sg.b * v.sb + (sg.ri + g.d * wi.start) * v.sh + (sg.rj + g.d * wj.start) * v.sw + h0 * v.sn;
            const int64_t kdh = g.d * k.sh, kdw = g.d * k.sw, vdh = g.d * v.sh, vdw = g.d * v.sw;

            float s[HG][CAP], dp[HG][CAP], mx[HG];
#pragma unroll
            for (int hg = 0; hg < HG; ++hg) mx[hg] = -INFINITY;
#pragma unroll
            for (int mi = 0; mi < K; ++mi) {
#pragma unroll
                for (int mj = 0; mj < K; ++mj) {
                    float kv[VEC], vv[VEC];
                    load_f<VEC, T, true>(kbase + mi * kdh + mj * kdw, kv);
                    load_f<VEC, T, true>(vbase + mi * vdh + mj * vdw, vv);
#pragma unroll
                    for (int hg = 0; hg < HG; ++hg) {
                        float a = 0.f, b = 0.f;
#pragma unroll
                        for (int e = 0; e < D; ++e) {
                            a = fmaf(qv[hg * D + e], kv[hg * D + e], a);
                            b = fmaf(gv[hg * D + e], vv[hg * D + e], b);
                        }
                        a = a * scale_log2e + s_rpb[((h0 + hg) * R + wi.pb + mi) * R + wj.pb + mj];
                        s[hg][mi * K + mj] = a;
                        dp[hg][mi * K + mj] = b;
                        mx[hg] = fmaxf(mx[hg], a);
                    }
                }
            }
            float2 st[HG];
#pragma unroll
            for (int hg = 0; hg < HG; ++hg) {
                float den = 0.f;
#pragma unroll
                for (int n = 0; n < K * K; ++n) {
                    float p = fast_exp2(s[hg][n] - mx[hg]);
                    s[hg][n] = p;
                    den += p;
                }
                float inv = 1.f / den, delta = 0.f;
#pragma unroll
                for (int n = 0; n < K * K; ++n) {
                    s[hg][n] *= inv;
                    delta = fmaf(s[hg][n], dp[hg][n], delta);
                }
#pragma unroll
                for (int n = 0; n < K * K; ++n) {
                    float ds = s[hg][n] * (dp[hg][n] - delta);
                    s[hg][n] = ds;
                    if (drpb_part != nullptr) racc[hg][n] += ds;
                }
                st[hg] = make_float2(mx[hg] + log2f(den), delta);  // lse in the log2 domain
            }
            float acc[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
#pragma unroll
            for (int mi = 0; mi < K; ++mi) {
#pragma unroll
                for (int mj = 0; mj < K; ++mj) {
                    float kv[VEC];
                    load_f<VEC, T, true>(kbase + mi * kdh + mj * kdw, kv);
#pragma unroll
                    for (int hg = 0; hg < HG; ++hg)
#pragma unroll
                        for (int e = 0; e < D; ++e) acc[hg * D + e] = fmaf(s[hg][mi * K + mj], kv[hg * D + e], acc[hg * D + e]);
                }
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] *= scale;
            store_f<VEC, T, true>(dq.ptr + sg.b * dq.sb + i * dq.sh + j * dq.sw + h0 * dq.sn, acc);
            float2* sp = stats + (((int64_t)sg.b * g.H + i) * g.W + j) * g.heads + h0;
#pragma unroll
            for (int hg = 0; hg < HG; ++hg) sp[hg] = st[hg];
        }
        if (drpb_part != nullptr && cur_pb >= 0) {
#pragma unroll
            for (int hg = 0; hg < HG; ++hg)
#pragma unroll
                for (int mi = 0; mi < K; ++mi)
#pragma unroll
                    for (int mj = 0; mj < K; ++mj)
                        atomicAdd(&s_acc[((h0 + hg) * R + cur_pb + mi) * R + wj.pb + mj], racc[hg][mi * K + mj]);
        }
    }
    if (drpb_part != nullptr) {
        __syncthreads();
        const int64_t cta = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        for (int x = threadIdx.x; x < nbins; x += kThreads) drpb_part[cta * nbins + x] = s_acc[x];
    }
}

// ------------------------------------------------------------------------------------------------
// backward kernel B: per key, gather over the inverse neighbourhood
// ------------------------------------------------------------------------------------------------
template <typename T, int KT, int D, int HG>
__global__ void __launch_bounds__(kThreads)
na2d_bwd_key_kernel(V5<const T> q, V5<const T> k, V5<const T> v, V5<const T> dout,
                    const float* __restrict__ rpb, const float2* __restrict__ stats,
                    V5<T> dk, V5<T> dv, NAGeom g, float scale) {
    constexpr int VEC = HG * D;
    const int K = KSize<KT>::get(g.K);
    const int R = 2 * K - 1;
    extern __shared__ float s_rpb[];
    if (rpb != nullptr) {
        for (int x = threadIdx.x; x < g.heads * R * R; x += kThreads) s_rpb[x] = rpb[x] * kLog2e;
        __syncthreads();
    }
    const SubGrid sg = decode_subgrid(g, blockIdx.z);
    const int NG = g.heads / HG;
    const int64_t gid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (gid >= (int64_t)sg.Hr * sg.Wr * NG) return;
    const int grp = (int)(gid % NG);
    const int pix = (int)(gid / NG);
    const int tj = pix % sg.Wr;
    const int ti = pix / sg.Wr;
    const int h0 = grp * HG;
    const int i = sg.ri + g.d * ti, j = sg.rj + g.d * tj;
    const float scale_log2e = scale * kLog2e;

    float kv[VEC], vv[VEC], ak[VEC], av[VEC];
    load_f<VEC, T, true>(k.ptr + sg.b * k.sb + i * k.sh + j * k.sw + h0 * k.sn, kv);
    load_f<VEC, T, true>(v.ptr + sg.b * v.sb + i * v.sh + j * v.sw + h0 * v.sn, vv);
#pragma unroll
    for (int e = 0; e < VEC; ++e) ak[e] = av[e] = 0.f;

    int lo_i, hi_i, lo_j, hi_j;
    inverse_window(ti, sg.Hr, K, lo_i, hi_i);
    inverse_window(tj, sg.Wr, K, lo_j, hi_j);
    for (int qi = lo_i; qi <= hi_i; ++qi) {
        const AxisWin wi = axis_window(qi, sg.Hr, K);
        const int pbi = wi.pb + (ti - wi.start);
        const int ii = sg.ri + g.d * qi;
        for (int qj = lo_j; qj <= hi_j; ++qj) {
            const AxisWin wj = axis_window(qj, sg.Wr, K);
            const int pbj = wj.pb + (tj - wj.start);
            const int jj = sg.rj + g.d * qj;
            float qv[VEC], gv[VEC];
            load_f<VEC, T, true>(q.ptr + sg.b * q.sb + ii * q.sh + jj * q.sw + h0 * q.sn, qv);
            load_f<VEC, T, true>(dout.ptr + sg.b * dout.sb + ii * dout.sh + jj * dout.sw + h0 * dout.sn, gv);
            const float2* sp = stats + (((int64_t)sg.b * g.H + ii) * g.W + jj) * g.heads + h0;
#pragma unroll
            for (int hg = 0; hg < HG; ++hg) {
                const float2 st = __ldg(sp + hg);
                float a = 0.f, b = 0.f;
#pragma unroll
                for (int e = 0; e < D; ++e) {
                    a = fmaf(qv[hg * D + e], kv[hg * D + e], a);
                    b = fmaf(gv[hg * D + e], vv[hg * D + e], b);
                }
                a *= scale_log2e;
                if (rpb != nullptr) a += s_rpb[((h0 + hg) * R + pbi) * R + pbj];
                const float p = fast_exp2(a - st.x);
                const float ds = p * (b - st.y);
#pragma unroll
                for (int e = 0; e < D; ++e) {
                    ak[hg * D + e] = fmaf(ds, qv[hg * D + e], ak[hg * D + e]);
                    av[hg * D + e] = fmaf(p, gv[hg * D + e], av[hg * D + e]);
                }
            }
        }
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) ak[e] *= scale;
    store_f<VEC, T, true>(dk.ptr + sg.b * dk.sb + i * dk.sh + j * dk.sw + h0 * dk.sn, ak);
    store_f<VEC, T, true>(dv.ptr + sg.b * dv.sb + i * dv.sh + j * dv.sw + h0 * dv.sn, av);
}

// Sums the per-CTA drpb partial tables in a fixed order: one CTA per bin, fixed-shape tree.
static __global__ void __launch_bounds__(256)
drpb_reduce_kernel(const float* __restrict__ part, int64_t n_part, int nbins, float* __restrict__ drpb) {
    __shared__ float sh[256];
    const int bin = blockIdx.x;
    float a = 0.f;
    for (int64_t c = threadIdx.x; c < n_part; c += 256) a += part[c * nbins + bin];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) drpb[bin] = sh[0];
}

// ------------------------------------------------------------------------------------------------
// host side: validation + dispatch
// ------------------------------------------------------------------------------------------------
inline int validate_dims(const lmnet_na2d_dims* d) {
    if (d == nullptr) return LMNET_ERR_INVALID_ARG;
    if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->heads <= 0 || d->D <= 0) return LMNET_ERR_INVALID_ARG;
    if (d->kernel_size < 3 || d->kernel_size > 13 || (d->kernel_size & 1) == 0) return LMNET_ERR_INVALID_ARG;
    if (d->dilation < 1) return LMNET_ERR_INVALID_ARG;
    if ((int64_t)d->kernel_size * d->dilation > d->H || (int64_t)d->kernel_size * d->dilation > d->W)
        return LMNET_ERR_INVALID_ARG;
    return LMNET_OK;
}

inline NAGeom make_geom(const lmnet_na2d_dims* d) {
    NAGeom g;
    g.B = d->B; g.H = d->H; g.W = d->W; g.heads = d->heads; g.D = d->D;
    g.K = d->kernel_size; g.d = d->dilation;
    g.Hmax = (d->H + d->dilation - 1) / d->dilation;
    g.Wmax = (d->W + d->dilation - 1) / d->dilation;
    return g;
}

template <typename T> static V5<T> as_v5(const lmnet_view5* v) {
    return V5<T>{reinterpret_cast<T*>(v->ptr), v->sb, v->sh, v->sw, v->sn};
}

// heads per thread: keep a thread's vector at 4..16 bytes and its score registers bounded
inline int pick_hg(int K, int D, int heads, size_t esize, bool heads_contig) {
    if (!heads_contig) return 1;
    int want = 1;
    if (K == 3) want = D == 1 ? 4 : D == 2 ? 2 : 1;
    else if (K == 5) want = D == 1 ? 2 : 1;
    (void)esize;
    while (want > 1 && heads % want != 0) want >>= 1;
    return want;
}

inline bool view_ok(const lmnet_view5* v, int vec_elems, size_t esize) {
    if (v == nullptr || v->ptr == nullptr) return false;
    size_t total = (size_t)vec_elems * esize;
    size_t vb = total % 16 == 0 ? 16 : total % 8 == 0 ? 8 : total % 4 == 0 ? 4 : esize;
    if (vb <= esize) return true;
    auto al = [&](int64_t s) { return ((size_t)(s < 0 ? -s : s) * esize) % vb == 0; };
    // the head stride only enters as (head-group index) * vec_elems when heads are contiguous (callers pass
    // vec_elems = HG*D with sn == D); for HG == 1 it is the stride between single heads and must be aligned
    const bool head_ok = (vec_elems > 0 && v->sn > 0 && vec_elems % v->sn == 0 && vec_elems != v->sn) ? true : al(v->sn);
    return ((uintptr_t)v->ptr % vb == 0) && al(v->sb) && al(v->sh) && al(v->sw) && head_ok;
}

enum class Op { Fwd, BwdQ, BwdK };

struct FusedArgs {
    const lmnet_view5 *q, *k, *v, *out, *dout, *dq, *dk, *dv;
    const float* rpb;
    float* lse;
    float2* stats;
    float* drpb_part;
    NAGeom g;
    float scale;
    cudaStream_t stream;
};

template <typename T, int KT, int D, int HG>
static int launch(Op op, const FusedArgs& a) {
    const NAGeom& g = a.g;
    const int NG = g.heads / HG;
    const int R = 2 * g.K - 1;
    const size_t rpb_bytes = (size_t)g.heads * R * R * sizeof(float);
    const unsigned z = (unsigned)(g.B * g.d * g.d);
    const double n_bytes = (double)g.B * g.H * g.W * g.heads * g.D * sizeof(T);  // one q-sized tensor
    {   // rpb tables above 48 KB (e.g. 12 heads x kernel 13) need the per-device opt-in
        static std::atomic<size_t> granted_f[kMaxDevices], granted_q[kMaxDevices], granted_k[kMaxDevices];
        const bool ok = op == Op::Fwd    ? ensure_smem(na2d_fwd_kernel<T, KT, D, HG>, a.rpb ? rpb_bytes : 0, granted_f)
                        : op == Op::BwdQ ? ensure_smem(na2d_bwd_query_kernel<T, KT, D, HG>, 2 * rpb_bytes, granted_q)
                                         : ensure_smem(na2d_bwd_key_kernel<T, KT, D, HG>, a.rpb ? rpb_bytes : 0, granted_k);
        if (!ok) return LMNET_ERR_UNSUPPORTED;
    }
    if (op == Op::Fwd) {
        int64_t total = (int64_t)g.Hmax * g.Wmax * NG;
        dim3 grid((unsigned)((total + kThreads - 1) / kThreads), 1, z);
        LMNET_LAUNCH(KID_NA_FWD, a.stream, 4 * n_bytes,
            (na2d_fwd_kernel<T, KT, D, HG><<<grid, kThreads, a.rpb ? rpb_bytes : 0, a.stream>>>(
                as_v5<const T>(a.q), as_v5<const T>(a.k), as_v5<const T>(a.v), a.rpb, as_v5<T>(a.out), a.lse, g,
                a.scale * kLog2e)));
    } else if (op == Op::BwdQ) {
        dim3 grid((unsigned)((g.Wmax * NG + kThreads - 1) / kThreads), (unsigned)((g.Hmax + kRowChunk - 1) / kRowChunk), z);
        LMNET_LAUNCH(KID_NA_BWD_QUERY, a.stream, 5 * n_bytes,
            (na2d_bwd_query_kernel<T, KT, D, HG><<<grid, kThreads, 2 * rpb_bytes, a.stream>>>(
                as_v5<const T>(a.q), as_v5<const T>(a.k), as_v5<const T>(a.v), as_v5<const T>(a.dout), a.rpb,
                as_v5<T>(a.dq), a.stats, a.drpb_part, g, a.scale)));
    } else {
        int64_t total = (int64_t)g.Hmax * g.Wmax * NG;
        dim3 grid((unsigned)((total + kThreads - 1) / kThreads), 1, z);
        LMNET_LAUNCH(KID_NA_BWD_KEY, a.stream, 2 * n_bytes,
            (na2d_bwd_key_kernel<T, KT, D, HG><<<grid, kThreads, a.rpb ? rpb_bytes : 0, a.stream>>>(
                as_v5<const T>(a.q), as_v5<const T>(a.k), as_v5<const T>(a.v), as_v5<const T>(a.dout), a.rpb, a.stats,
                as_v5<T>(a.dk), as_v5<T>(a.dv), g, a.scale)));
    }
    return LMNET_OK;
}

template <typename T, int KT, int D>
static int dispatch_hg(Op op, const FusedArgs& a, int hg) {
    if constexpr (KT == 3 && D == 1) { if (hg == 4) return launch<T, KT, D, 4>(op, a); }
    if constexpr ((KT == 3 && D <= 2) || (KT == 5 && D == 1)) { if (hg >= 2) return launch<T, KT, D, 2>(op, a); }
    return launch<T, KT, D, 1>(op, a);
}

template <typename T, int KT>
static int dispatch_d(Op op, const FusedArgs& a, int hg) {
    switch (a.g.D) {
        case 1: return dispatch_hg<T, KT, 1>(op, a, hg);
        case 2: return dispatch_hg<T, KT, 2>(op, a, hg);
        case 4: return dispatch_hg<T, KT, 4>(op, a, hg);
        case 8: return dispatch_hg<T, KT, 8>(op, a, hg);
        case 16: return dispatch_hg<T, KT, 16>(op, a, hg);
        case 32: return dispatch_hg<T, KT, 32>(op, a, hg);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}

template <typename T>
static int dispatch_k(Op op, const FusedArgs& a, int hg) {
    switch (a.g.K) {
        case 3: return dispatch_d<T, 3>(op, a, hg);
        case 5: return dispatch_d<T, 5>(op, a, hg);
        case 7: return dispatch_d<T, 7>(op, a, hg);
        default: return dispatch_d<T, 0>(op, a, hg);
    }
}

}  // namespace lmnet
