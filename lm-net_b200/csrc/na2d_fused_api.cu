// na2d_fused_api.cu — extern "C" entry points of the fused neighbourhood attention
// (validation, workspace layout, per-dtype dispatch).  Kernels: na2d_fused.cuh.
#include "na2d_stream.cuh"

namespace lmnet {

// one translation unit per dtype (na2d_fused_inst.cu compiled three times)
int fused_dispatch_f32(Op op, const FusedArgs& a, int hg);
int fused_dispatch_bf16(Op op, const FusedArgs& a, int hg);
int fused_dispatch_f16(Op op, const FusedArgs& a, int hg);

static int dispatch(Op op, const FusedArgs& a, int dtype, int hg) {
    switch (dtype) {
        case LMNET_F32: return fused_dispatch_f32(op, a, hg);
        case LMNET_BF16: return fused_dispatch_bf16(op, a, hg);
        case LMNET_F16: return fused_dispatch_f16(op, a, hg);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}

static size_t esize_of(int dtype) { return dtype == LMNET_F32 ? 4 : 2; }

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct BwdLayout {
    size_t stats_off, part_off, total;
    int64_t n_part;
};
static BwdLayout bwd_layout(const lmnet_na2d_dims* d) {
    // worst case HG == 1 (most CTAs along x)
    NAGeom g = make_geom(d);
    int R = 2 * g.K - 1;
    BwdLayout L;
    L.stats_off = 0;
    size_t stats_bytes = (size_t)g.B * g.H * g.W * g.heads * sizeof(float2);
    L.part_off = align_up(stats_bytes, 256);
    int64_t gx = ((int64_t)g.Wmax * g.heads + kThreads - 1) / kThreads;
    int64_t gy = (g.Hmax + kRowChunk - 1) / kRowChunk;
    L.n_part = gx * gy * g.B * g.d * g.d;
    L.total = L.part_off + (size_t)L.n_part * g.heads * R * R * sizeof(float);
    return L;
}

}  // namespace lmnet

using namespace lmnet;

extern "C" int lmnet_na2d_fwd(const lmnet_view5* q, const lmnet_view5* k, const lmnet_view5* v,
                              const float* rpb, const lmnet_view5* out, float* lse,
                              const lmnet_na2d_dims* dims, float scale, int dtype, void* stream) {
    int rc = validate_dims(dims);
    if (rc != LMNET_OK) return rc;
    if (!q || !k || !v || !out) return LMNET_ERR_INVALID_ARG;
    const size_t es = esize_of(dtype);
    bool contig = q->sn == dims->D && k->sn == dims->D && v->sn == dims->D && out->sn == dims->D;
    int hg = pick_hg(dims->kernel_size, dims->D, dims->heads, es, contig);
    for (; hg >= 1; hg >>= 1) {
        int vec = hg * dims->D;
        if (view_ok(q, vec, es) && view_ok(k, vec, es) && view_ok(v, vec, es) && view_ok(out, vec, es)) break;
    }
    if (hg < 1) return LMNET_ERR_UNSUPPORTED;
    FusedArgs a{};
    a.q = q; a.k = k; a.v = v; a.out = out; a.rpb = rpb; a.lse = lse;
    a.g = make_geom(dims); a.scale = scale; a.stream = (cudaStream_t)stream;
    rc = stream_fwd(a, dtype);   // row-streaming kernel for 16-bit storage; UNSUPPORTED = not eligible
    if (rc != LMNET_ERR_UNSUPPORTED) return rc;
    return dispatch(Op::Fwd, a, dtype, hg);
}

extern "C" size_t lmnet_na2d_bwd_workspace_bytes(const lmnet_na2d_dims* dims) {
    if (validate_dims(dims) != LMNET_OK) return 0;
    return bwd_layout(dims).total;
}

extern "C" int lmnet_na2d_bwd(const lmnet_view5* q, const lmnet_view5* k, const lmnet_view5* v,
                              const float* rpb, const lmnet_view5* dout,
                              const lmnet_view5* dq, const lmnet_view5* dk, const lmnet_view5* dv,
                              float* drpb, void* workspace, size_t workspace_bytes,
                              const lmnet_na2d_dims* dims, float scale, int dtype, void* stream) {
    int rc = validate_dims(dims);
    if (rc != LMNET_OK) return rc;
    if (!q || !k || !v || !dout || !dq || !dk || !dv || !workspace) return LMNET_ERR_INVALID_ARG;
    BwdLayout L = bwd_layout(dims);
    if (workspace_bytes < L.total) return LMNET_ERR_WORKSPACE;
    const size_t es = esize_of(dtype);
    const lmnet_view5* views[] = {q, k, v, dout, dq, dk, dv};
    bool contig = true;
    for (auto* x : views) contig = contig && x->sn == dims->D;
    int hg = pick_hg(dims->kernel_size, dims->D, dims->heads, es, contig);
    for (; hg >= 1; hg >>= 1) {
        bool ok = true;
        for (auto* x : views) ok = ok && view_ok(x, hg * dims->D, es);
        if (ok) break;
    }
    if (hg < 1) return LMNET_ERR_UNSUPPORTED;
    FusedArgs a{};
    a.q = q; a.k = k; a.v = v; a.dout = dout; a.dq = dq; a.dk = dk; a.dv = dv; a.rpb = rpb;
    a.stats = reinterpret_cast<float2*>((char*)workspace + L.stats_off);
    a.drpb_part = drpb ? reinterpret_cast<float*>((char*)workspace + L.part_off) : nullptr;
    a.g = make_geom(dims); a.scale = scale; a.stream = (cudaStream_t)stream;
    const NAGeom& g = a.g;
    int64_t n_part = 0;
    rc = stream_bwd(a, dtype, L.n_part, &n_part);   // one-kernel streaming backward when eligible
    if (rc != LMNET_OK && rc != LMNET_ERR_UNSUPPORTED) return rc;
    if (rc == LMNET_ERR_UNSUPPORTED) {
        rc = dispatch(Op::BwdQ, a, dtype, hg);
        if (rc != LMNET_OK) return rc;
        rc = dispatch(Op::BwdK, a, dtype, hg);
        if (rc != LMNET_OK) return rc;
        int NG = g.heads / hg;
        int64_t gx = ((int64_t)g.Wmax * NG + kThreads - 1) / kThreads;
        int64_t gy = (g.Hmax + kRowChunk - 1) / kRowChunk;
        n_part = gx * gy * g.B * g.d * g.d;
    }
    if (drpb) {
        int R = 2 * g.K - 1;
        LMNET_LAUNCH(KID_NA_DRPB_REDUCE, a.stream, 0,
            (drpb_reduce_kernel<<<g.heads * R * R, 256, 0, a.stream>>>(a.drpb_part, n_part, g.heads * R * R, drpb)));
    }
    return LMNET_OK;
}
