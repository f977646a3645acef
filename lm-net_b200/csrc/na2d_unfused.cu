// na2d_unfused.cu — the unfused neighbourhood-attention operators (sm_100a).
//
// Replaces natten.functional.na2d_qk / na2d_av (0.14 names natten2dqkrpb / natten2dav) and
// their autograd (§2b K1, K3, K4).  These materialise the attention map
// [B,heads,H,W,K*K] by definition, so they are the microbenchmark / compatibility surface; the
// LM-Net module itself runs the fused kernels of na2d_fused.cuh.
//
// Three stencil shapes cover all six passes (names follow the NAT papers):
//   PN  pointwise-neighbourhood      o[n] = sum_e a[e] * b_n[e] (+ rpb)   : qk fwd, av dattn
//   NN  neighbourhood-neighbourhood  o[e] = sum_n w[n] * b_n[e]           : av fwd, qk dq
//   IN  inverse-neighbourhood        o_t[e] = sum_{i : t in N(i)} w_i[n(i,t)] * a_i[e] : dk, dv
// plus the rpb gradient (binned sum of dattn, per-CTA partials reduced in a fixed order).
// One thread per (pixel, head); any head dim (runtime loop in chunks of 8 registers).
#include "na2d_fused.cuh"

namespace lmnet {

constexpr int kDC = 8;  // head-dim chunk held in registers

template <typename T>
__device__ __forceinline__ const T* px(const V5<const T>& t, int b, int i, int j, int h) {
    return t.ptr + b * t.sb + i * t.sh + j * t.sw + h * t.sn;
}
template <typename T>
__device__ __forceinline__ T* px(const V5<T>& t, int b, int i, int j, int h) {
    return t.ptr + b * t.sb + i * t.sh + j * t.sw + h * t.sn;
}

struct PixelId {
    SubGrid sg;
    int h, ti, tj, i, j;
    bool valid;
};
// grid: x over pixels of the sub-grid, y = head, z = (batch, sub-grid)
__device__ __forceinline__ PixelId decode_pixel(const NAGeom& g) {
    PixelId p;
    p.sg = decode_subgrid(g, blockIdx.z);
    p.h = blockIdx.y;
    int64_t gid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    p.valid = gid < (int64_t)p.sg.Hr * p.sg.Wr;
    int pix = p.valid ? (int)gid : 0;
    p.tj = pix % p.sg.Wr;
    p.ti = pix / p.sg.Wr;
    p.i = p.sg.ri + g.d * p.ti;
    p.j = p.sg.rj + g.d * p.tj;
    return p;
}

__device__ __forceinline__ int64_t attn_index(const NAGeom& g, int b, int h, int i, int j, int KK) {
    return ((((int64_t)b * g.heads + h) * g.H + i) * g.W + j) * KK;
}

// PN: o[b,h,i,j,n] = sum_e a[b,i,j,h,e] * bb[b,key_n,h,e] (+ rpb[h,pi,pj])
template <typename T, int KT>
__global__ void __launch_bounds__(kThreads)
na2d_pn_kernel(V5<const T> a, V5<const T> bb, const float* __restrict__ rpb, T* __restrict__ o, NAGeom g) {
    const int K = KSize<KT>::get(g.K);
    const int R = 2 * K - 1, KK = K * K;
    const PixelId p = decode_pixel(g);
    if (!p.valid) return;
    const AxisWin wi = axis_window(p.ti, p.sg.Hr, K), wj = axis_window(p.tj, p.sg.Wr, K);
    const T* ap = px(a, p.sg.b, p.i, p.j, p.h);
    T* op = o + attn_index(g, p.sg.b, p.h, p.i, p.j, KK);
#pragma unroll
    for (int mi = 0; mi < K; ++mi) {
#pragma unroll
        for (int mj = 0; mj < K; ++mj) {
            const T* bp = px(bb, p.sg.b, p.sg.ri + g.d * (wi.start + mi), p.sg.rj + g.d * (wj.start + mj), p.h);
            float s = 0.f;
            for (int e = 0; e < g.D; ++e) s = fmaf(to_f(ap[e]), to_f(bp[e]), s);
            if (rpb != nullptr) s += __ldg(rpb + ((int64_t)p.h * R + wi.pb + mi) * R + wj.pb + mj);
            op[mi * K + mj] = from_f<T>(s);
        }
    }
}

// NN: o[b,i,j,h,e] = sum_n w[b,h,i,j,n] * bb[b,key_n,h,e]
template <typename T, int KT>
__global__ void __launch_bounds__(kThreads)
na2d_nn_kernel(const T* __restrict__ w, V5<const T> bb, V5<T> o, NAGeom g) {
    const int K = KSize<KT>::get(g.K);
    const int KK = K * K;
    const PixelId p = decode_pixel(g);
    if (!p.valid) return;
    const AxisWin wi = axis_window(p.ti, p.sg.Hr, K), wj = axis_window(p.tj, p.sg.Wr, K);
    const T* wp = w + attn_index(g, p.sg.b, p.h, p.i, p.j, KK);
    T* op = px(o, p.sg.b, p.i, p.j, p.h);
    for (int e0 = 0; e0 < g.D; e0 += kDC) {
        float acc[kDC];
#pragma unroll
        for (int e = 0; e < kDC; ++e) acc[e] = 0.f;
#pragma unroll
        for (int mi = 0; mi < K; ++mi) {
#pragma unroll
            for (int mj = 0; mj < K; ++mj) {
                const T* bp = px(bb, p.sg.b, p.sg.ri + g.d * (wi.start + mi), p.sg.rj + g.d * (wj.start + mj), p.h);
                const float ww = to_f(wp[mi * K + mj]);
#pragma unroll
                for (int e = 0; e < kDC; ++e)
                    if (e0 + e < g.D) acc[e] = fmaf(ww, to_f(bp[e0 + e]), acc[e]);
            }
        }
#pragma unroll
        for (int e = 0; e < kDC; ++e)
            if (e0 + e < g.D) op[e0 + e] = from_f<T>(acc[e]);
    }
}

// IN: o[b,t,h,e] = sum over queries i whose window contains key t of w[b,h,i,n(i,t)] * a[b,i,h,e]
template <typename T, int KT>
__global__ void __launch_bounds__(kThreads)
na2d_in_kernel(const T* __restrict__ w, V5<const T> a, V5<T> o, NAGeom g) {
    const int K = KSize<KT>::get(g.K);
    const int KK = K * K;
    const PixelId p = decode_pixel(g);
    if (!p.valid) return;
    int lo_i, hi_i, lo_j, hi_j;
    inverse_window(p.ti, p.sg.Hr, K, lo_i, hi_i);
    inverse_window(p.tj, p.sg.Wr, K, lo_j, hi_j);
    T* op = px(o, p.sg.b, p.i, p.j, p.h);
    for (int e0 = 0; e0 < g.D; e0 += kDC) {
        float acc[kDC];
#pragma unroll
        for (int e = 0; e < kDC; ++e) acc[e] = 0.f;
        for (int qi = lo_i; qi <= hi_i; ++qi) {
            const int mi = p.ti - axis_window(qi, p.sg.Hr, K).start;
            const int ii = p.sg.ri + g.d * qi;
            for (int qj = lo_j; qj <= hi_j; ++qj) {
                const int mj = p.tj - axis_window(qj, p.sg.Wr, K).start;
                const int jj = p.sg.rj + g.d * qj;
                const float ww = to_f(w[attn_index(g, p.sg.b, p.h, ii, jj, KK) + mi * K + mj]);
                const T* ap = px(a, p.sg.b, ii, jj, p.h);
#pragma unroll
                for (int e = 0; e < kDC; ++e)
                    if (e0 + e < g.D) acc[e] = fmaf(ww, to_f(ap[e0 + e]), acc[e]);
            }
        }
#pragma unroll
        for (int e = 0; e < kDC; ++e)
            if (e0 + e < g.D) op[e0 + e] = from_f<T>(acc[e]);
    }
}

// rpb gradient: drpb[h,pi,pj] = sum_{b,i,j} dattn[b,h,i,j,n] over the neighbours that map to the bin.
// grid: x over columns, y over row chunks, z = (batch, sub-grid, head).  Per-CTA partial table.
template <typename T, int KT>
__global__ void __launch_bounds__(kThreads)
na2d_rpbgrad_kernel(const T* __restrict__ dattn, float* __restrict__ part, NAGeom g) {
    constexpr int CAP = KSize<KT>::kk_cap;
    const int K = KSize<KT>::get(g.K);
    const int R = 2 * K - 1, KK = K * K;
    extern __shared__ float s_acc[];  // [R*R]
    for (int x = threadIdx.x; x < R * R; x += kThreads) s_acc[x] = 0.f;
    __syncthreads();
    const int h = blockIdx.z % g.heads;
    const SubGrid sg = decode_subgrid(g, blockIdx.z / g.heads);
    const int tj = blockIdx.x * kThreads + threadIdx.x;
    const int row0 = blockIdx.y * kRowChunk, row1 = min(row0 + kRowChunk, sg.Hr);
    if (tj < sg.Wr && row0 < row1) {
        const AxisWin wj = axis_window(tj, sg.Wr, K);
        const int j = sg.rj + g.d * tj;
        float racc[CAP];
        int cur_pb = -1;
        for (int ti = row0; ti < row1; ++ti) {
            const AxisWin wi = axis_window(ti, sg.Hr, K);
            if (wi.pb != cur_pb) {
                if (cur_pb >= 0)
                    for (int mi = 0; mi < K; ++mi)
                        for (int mj = 0; mj < K; ++mj) atomicAdd(&s_acc[(cur_pb + mi) * R + wj.pb + mj], racc[mi * K + mj]);
#pragma unroll
                for (int n = 0; n < KK; ++n) racc[n] = 0.f;
                cur_pb = wi.pb;
            }
            const T* dp = dattn + attn_index(g, sg.b, h, sg.ri + g.d * ti, j, KK);
#pragma unroll
            for (int n = 0; n < KK; ++n) racc[n] += to_f(dp[n]);
        }
        for (int mi = 0; mi < K; ++mi)
            for (int mj = 0; mj < K; ++mj) atomicAdd(&s_acc[(cur_pb + mi) * R + wj.pb + mj], racc[mi * K + mj]);
    }
    __syncthreads();
    // partial tables are grouped by head: part[h][cta_in_head][R*R]
    const int64_t n_per_head = (int64_t)gridDim.x * gridDim.y * (gridDim.z / g.heads);
    const int64_t cta = ((int64_t)(blockIdx.z / g.heads) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    for (int x = threadIdx.x; x < R * R; x += kThreads) part[((int64_t)h * n_per_head + cta) * R * R + x] = s_acc[x];
}

static __global__ void __launch_bounds__(256)
rpbgrad_reduce_kernel(const float* __restrict__ part, int64_t n_per_head, int RR, float* __restrict__ drpb) {
    __shared__ float sh[256];
    const int h = blockIdx.x / RR, x = blockIdx.x % RR;
    float a = 0.f;
    for (int64_t c = threadIdx.x; c < n_per_head; c += 256) a += part[((int64_t)h * n_per_head + c) * RR + x];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) drpb[blockIdx.x] = sh[0];
}

// ---------------------------------------------------------------------------------------------
enum class UOp { PN, NN, IN, RPB };

struct UArgs {
    V5<const void> a, b;  // strided operands (typed later)
    const void* w;        // attention-shaped operand
    const float* rpb;
    void* o_attn;         // attention-shaped output
    V5<void> o;           // strided output
    float* part;
    NAGeom g;
    cudaStream_t stream;
};

template <typename T> static V5<const T> cv(const V5<const void>& v) {
    return V5<const T>{reinterpret_cast<const T*>(v.ptr), v.sb, v.sh, v.sw, v.sn};
}
template <typename T> static V5<T> mv(const V5<void>& v) {
    return V5<T>{reinterpret_cast<T*>(v.ptr), v.sb, v.sh, v.sw, v.sn};
}

template <typename T, int KT>
static int ulaunch(UOp op, const UArgs& u) {
    const NAGeom& g = u.g;
    const unsigned gx = (unsigned)(((int64_t)g.Hmax * g.Wmax + kThreads - 1) / kThreads);
    const unsigned z = (unsigned)(g.B * g.d * g.d);
    dim3 grid(gx, (unsigned)g.heads, z);
    const double n_bytes = (double)g.B * g.H * g.W * g.heads * g.D * sizeof(T);
    const double a_bytes = (double)g.B * g.H * g.W * g.heads * g.K * g.K * sizeof(T);
    switch (op) {
        case UOp::PN:
            LMNET_LAUNCH(KID_NA_PN, u.stream, 2 * n_bytes + a_bytes,
                (na2d_pn_kernel<T, KT><<<grid, kThreads, 0, u.stream>>>(cv<T>(u.a), cv<T>(u.b), u.rpb, (T*)u.o_attn, g)));
            break;
        case UOp::NN:
            LMNET_LAUNCH(KID_NA_NN, u.stream, 2 * n_bytes + a_bytes,
                (na2d_nn_kernel<T, KT><<<grid, kThreads, 0, u.stream>>>((const T*)u.w, cv<T>(u.b), mv<T>(u.o), g)));
            break;
        case UOp::IN:
            LMNET_LAUNCH(KID_NA_IN, u.stream, 2 * n_bytes + a_bytes,
                (na2d_in_kernel<T, KT><<<grid, kThreads, 0, u.stream>>>((const T*)u.w, cv<T>(u.a), mv<T>(u.o), g)));
            break;
        case UOp::RPB: {
            const int R = 2 * g.K - 1;
            dim3 rg((unsigned)((g.Wmax + kThreads - 1) / kThreads), (unsigned)((g.Hmax + kRowChunk - 1) / kRowChunk),
                    z * (unsigned)g.heads);
            LMNET_LAUNCH(KID_NA_RPBGRAD, u.stream, a_bytes,
                (na2d_rpbgrad_kernel<T, KT><<<rg, kThreads, R * R * sizeof(float), u.stream>>>((const T*)u.w, u.part, g)));
            break;
        }
    }
    return LMNET_OK;
}

template <typename T>
static int udispatch_k(UOp op, const UArgs& u) {
    switch (u.g.K) {
        case 3: return ulaunch<T, 3>(op, u);
        case 5: return ulaunch<T, 5>(op, u);
        case 7: return ulaunch<T, 7>(op, u);
        default: return ulaunch<T, 0>(op, u);
    }
}

static int udispatch(UOp op, const UArgs& u, int dtype) {
    switch (dtype) {
        case LMNET_F32: return udispatch_k<float>(op, u);
        case LMNET_BF16: return udispatch_k<__nv_bfloat16>(op, u);
        case LMNET_F16: return udispatch_k<__half>(op, u);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}

static V5<const void> cview(const lmnet_view5* v) { return V5<const void>{v->ptr, v->sb, v->sh, v->sw, v->sn}; }
static V5<void> mview(const lmnet_view5* v) { return V5<void>{v->ptr, v->sb, v->sh, v->sw, v->sn}; }

static int64_t rpb_parts_per_head(const NAGeom& g) {
    int64_t gx = (g.Wmax + kThreads - 1) / kThreads, gy = (g.Hmax + kRowChunk - 1) / kRowChunk;
    return gx * gy * g.B * g.d * g.d;
}

}  // namespace lmnet

using namespace lmnet;

extern "C" int lmnet_na2d_qk_fwd(const lmnet_view5* q, const lmnet_view5* k, const float* rpb, void* attn,
                                 const lmnet_na2d_dims* dims, int dtype, void* stream) {
    int rc = validate_dims(dims);
    if (rc != LMNET_OK) return rc;
    if (!q || !k || !q->ptr || !k->ptr || !attn) return LMNET_ERR_INVALID_ARG;
    UArgs u{};
    u.a = cview(q); u.b = cview(k); u.rpb = rpb; u.o_attn = attn;
    u.g = make_geom(dims); u.stream = (cudaStream_t)stream;
    return udispatch(UOp::PN, u, dtype);
}

extern "C" size_t lmnet_na2d_qk_bwd_workspace_bytes(const lmnet_na2d_dims* dims) {
    if (validate_dims(dims) != LMNET_OK) return 0;
    NAGeom g = make_geom(dims);
    int R = 2 * g.K - 1;
    return (size_t)rpb_parts_per_head(g) * g.heads * R * R * sizeof(float);
}

extern "C" int lmnet_na2d_qk_bwd(const lmnet_view5* q, const lmnet_view5* k, const void* dattn,
                                 const lmnet_view5* dq, const lmnet_view5* dk, float* drpb,
                                 void* workspace, size_t workspace_bytes,
                                 const lmnet_na2d_dims* dims, int dtype, void* stream) {
    int rc = validate_dims(dims);
    if (rc != LMNET_OK) return rc;
    if (!q || !k || !dattn || !dq || !dk || !q->ptr || !k->ptr || !dq->ptr || !dk->ptr) return LMNET_ERR_INVALID_ARG;
    UArgs u{};
    u.g = make_geom(dims); u.stream = (cudaStream_t)stream;
    u.w = dattn;
    // dq = NN(dattn, k)
    u.b = cview(k); u.o = mview(dq);
    rc = udispatch(UOp::NN, u, dtype);
    if (rc != LMNET_OK) return rc;
    // dk = IN(dattn, q)
    u.a = cview(q); u.o = mview(dk);
    rc = udispatch(UOp::IN, u, dtype);
    if (rc != LMNET_OK) return rc;
    if (drpb != nullptr) {
        if (workspace == nullptr || workspace_bytes < lmnet_na2d_qk_bwd_workspace_bytes(dims)) return LMNET_ERR_WORKSPACE;
        u.part = (float*)workspace;
        rc = udispatch(UOp::RPB, u, dtype);
        if (rc != LMNET_OK) return rc;
        int R = 2 * u.g.K - 1;
        LMNET_LAUNCH(KID_NA_RPBGRAD_REDUCE, u.stream, 0,
            (rpbgrad_reduce_kernel<<<u.g.heads * R * R, 256, 0, u.stream>>>(u.part, rpb_parts_per_head(u.g), R * R, drpb)));
    }
    return LMNET_OK;
}

extern "C" int lmnet_na2d_av_fwd(const void* attn, const lmnet_view5* v, const lmnet_view5* out,
                                 const lmnet_na2d_dims* dims, int dtype, void* stream) {
    int rc = validate_dims(dims);
    if (rc != LMNET_OK) return rc;
    if (!attn || !v || !out || !v->ptr || !out->ptr) return LMNET_ERR_INVALID_ARG;
    UArgs u{};
    u.g = make_geom(dims); u.stream = (cudaStream_t)stream;
    u.w = attn; u.b = cview(v); u.o = mview(out);
    return udispatch(UOp::NN, u, dtype);
}

extern "C" int lmnet_na2d_av_bwd(const void* attn, const lmnet_view5* v, const lmnet_view5* dout,
                                 void* dattn, const lmnet_view5* dv,
                                 const lmnet_na2d_dims* dims, int dtype, void* stream) {
    int rc = validate_dims(dims);
    if (rc != LMNET_OK) return rc;
    if (!attn || !v || !dout || !dattn || !dv || !v->ptr || !dout->ptr || !dv->ptr) return LMNET_ERR_INVALID_ARG;
    UArgs u{};
    u.g = make_geom(dims); u.stream = (cudaStream_t)stream;
    // dattn = PN(dout, v)
    u.a = cview(dout); u.b = cview(v); u.rpb = nullptr; u.o_attn = dattn;
    rc = udispatch(UOp::PN, u, dtype);
    if (rc != LMNET_OK) return rc;
    // dv = IN(attn, dout)
    u.w = attn; u.a = cview(dout); u.o = mview(dv);
    return udispatch(UOp::IN, u, dtype);
}
