// na2d_stream.cu — host side of the row-streaming fused neighbourhood attention (na2d_stream.cuh):
// eligibility, stripe / band sizing, shared-memory opt-in and the template fan-out for bf16 / fp16.
#include "na2d_stream.cuh"

#include <atomic>
#include <cstdlib>

namespace lmnet {
namespace {

bool stream_disabled() {
    static const bool off = getenv("LMNET_NA_V1") != nullptr && atoi(getenv("LMNET_NA_V1")) != 0;
    return off;
}

// heads per thread: 8-byte vectors at K = 3 (16 at head dim 8), narrower for larger K (score registers)
int stream_hg(int K, int D) {
    if (K == 3) return D == 1 ? 4 : D == 2 ? 2 : 1;
    return D == 1 ? 2 : 1;   // >= 4-byte vectors (cp.async granularity)
}

bool strides_ok(const lmnet_view5* v, int D, int align_bytes) {
    if (v == nullptr || v->ptr == nullptr || v->sn != D) return false;
    if ((uintptr_t)v->ptr % align_bytes != 0) return false;
    const int64_t ss[3] = {v->sb, v->sh, v->sw};
    for (int64_t s : ss)
        if (s < 0 || (s * 2) % align_bytes != 0) return false;
    return true;
}

struct Plan {
    StreamCfg cfg;
    int HG, stripes, bands;
    unsigned z;
};

bool make_plan(const NAGeom& g, bool bwd, int64_t max_parts, int ctas_per_sm, Plan& p) {
    const int K = g.K, NS = K / 2;
    if (g.D != 1 && g.D != 2 && g.D != 4 && g.D != 8) return false;
    p.HG = stream_hg(K, g.D);
    if (g.heads % p.HG != 0 || (g.heads * g.D) % 2 != 0) return false;
    p.cfg.NG = g.heads / p.HG;
    if (p.cfg.NG > kStreamThreads) return false;
    p.cfg.QW = kStreamThreads / p.cfg.NG;
    const int halo = bwd ? 2 * NS : 0;
    if (p.cfg.QW - halo < K) return false;
    if (2 * NS > p.cfg.QW) return false;   // the halo columns are staged by the first 2*(K/2) column threads
    if (bwd && (g.H / g.d < 2 * K || g.W / g.d < 2 * K)) return false;   // stream_boundary needs L >= 2K
    p.z = (unsigned)(g.B * g.d * g.d);
    // equal-width stripes: as few as the widest CTA allows, then shrink the CTA to the stripe
    p.stripes = (g.Wmax + (p.cfg.QW - halo) - 1) / (p.cfg.QW - halo);
    int TW = (g.Wmax + p.stripes - 1) / p.stripes;
    if (TW < K) TW = K;
    p.cfg.QW = TW + halo;
    // bands: fill the GPU once (148 SMs x resident CTAs) rather than leave a mostly empty last wave
    const int min_rb = 8 > K ? 8 : K;
    const size_t smem = bwd ? stream_bwd_smem_bytes(K, g.D, p.HG, g.heads, p.cfg.QW) : stream_fwd_smem_bytes(K, g.D, g.heads, p.cfg.QW);
    const int by_smem = (int)((size_t)(227 * 1024) / (smem + 1024));   // 1 KB per CTA reserved by the driver
    if (by_smem < 1) return false;
    const int64_t resident = 148 * (int64_t)(ctas_per_sm < by_smem ? ctas_per_sm : by_smem);
    // whole waves: with b bands the grid is b*stripes*z CTAs; pick the wave count (1..6) whose last wave is fullest
    const int64_t sz = (int64_t)p.stripes * p.z;
    const int64_t max_bands = g.Hmax / min_rb > 0 ? g.Hmax / min_rb : 1;
    int64_t bands = 1;
    double best = -1.0;
    for (int w = 1; w <= 6; ++w) {
        int64_t b = w * resident / sz;
        b = b < 1 ? 1 : b > max_bands ? max_bands : b;
        const int64_t waves = (b * sz + resident - 1) / resident;
        const double eff = (double)(b * sz) / (double)(waves * resident);
        if (eff > best + 0.03) { best = eff; bands = b; }
    }
    if (bwd && max_parts > 0) {
        const int64_t cap = max_parts / ((int64_t)p.stripes * p.z);
        if (cap < 1) return false;
        if (bands > cap) bands = cap;
    }
    p.cfg.RB = (int)((g.Hmax + bands - 1) / bands);
    if (p.cfg.RB < K) p.cfg.RB = K;
    p.bands = (g.Hmax + p.cfg.RB - 1) / p.cfg.RB;
    return true;
}

template <typename T, int KT, int D, int HG>
int launch_fwd(const FusedArgs& a, const Plan& p) {
    const NAGeom& g = a.g;
    const size_t smem = StreamFwdSmem<KT, D, HG>::bytes(g.heads, p.cfg.QW);
    auto kern = na2d_stream_fwd_kernel<T, KT, D, HG>;
    static std::atomic<size_t> granted[kMaxDevices];
    if (!ensure_smem(kern, smem, granted)) return LMNET_ERR_UNSUPPORTED;
    const dim3 grid((unsigned)p.stripes, (unsigned)p.bands, p.z);
    const double n_bytes = (double)g.B * g.H * g.W * g.heads * g.D * sizeof(T);
    LMNET_LAUNCH(KID_NA_STREAM_FWD, a.stream, 4 * n_bytes,
        (kern<<<grid, p.cfg.QW * p.cfg.NG, smem, a.stream>>>(
            as_v5<const T>(a.q), as_v5<const T>(a.k), as_v5<const T>(a.v), a.rpb, as_v5<T>(a.out), a.lse, g, p.cfg, a.scale)));
    return LMNET_OK;
}

template <typename T, int KT, int D, int HG>
int launch_bwd(const FusedArgs& a, const Plan& p) {
    const NAGeom& g = a.g;
    const size_t smem = StreamBwdSmem<KT, D, HG>::bytes(g.heads, p.cfg.QW);
    auto kern = na2d_stream_bwd_kernel<T, KT, D, HG>;
    static std::atomic<size_t> granted[kMaxDevices];
    if (!ensure_smem(kern, smem, granted)) return LMNET_ERR_UNSUPPORTED;
    const dim3 grid((unsigned)p.stripes, (unsigned)p.bands, p.z);
    const double n_bytes = (double)g.B * g.H * g.W * g.heads * g.D * sizeof(T);
    LMNET_LAUNCH(KID_NA_STREAM_BWD, a.stream, 7 * n_bytes,
        (kern<<<grid, p.cfg.QW * p.cfg.NG, smem, a.stream>>>(
            as_v5<const T>(a.q), as_v5<const T>(a.k), as_v5<const T>(a.v), as_v5<const T>(a.dout), a.rpb,
            as_v5<T>(a.dq), as_v5<T>(a.dk), as_v5<T>(a.dv), a.drpb_part, g, p.cfg, a.scale)));
    return LMNET_OK;
}

template <typename T, int KT>
int fwd_by_d(const FusedArgs& a, const Plan& p) {
    constexpr int HG1 = KT == 3 ? 4 : 2;
    constexpr int HG2 = KT == 3 ? 2 : 1;
    switch (a.g.D) {
        case 1: return launch_fwd<T, KT, 1, HG1>(a, p);
        case 2: return launch_fwd<T, KT, 2, HG2>(a, p);
        case 4: return launch_fwd<T, KT, 4, 1>(a, p);
        case 8: return launch_fwd<T, KT, 8, 1>(a, p);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}

template <typename T>
int fwd_typed(const FusedArgs& a, const Plan& p) {
    switch (a.g.K) {
        case 3: return fwd_by_d<T, 3>(a, p);
        case 5: return fwd_by_d<T, 5>(a, p);
        case 7: return fwd_by_d<T, 7>(a, p);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}

template <typename T>
int bwd_typed(const FusedArgs& a, const Plan& p) {
    if (a.g.K != 3) return LMNET_ERR_UNSUPPORTED;
    switch (a.g.D) {
        case 1: return launch_bwd<T, 3, 1, 4>(a, p);
        case 2: return launch_bwd<T, 3, 2, 2>(a, p);
        case 4: return launch_bwd<T, 3, 4, 1>(a, p);
        case 8: return launch_bwd<T, 3, 8, 1>(a, p);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}

}  // namespace

int stream_fwd(const FusedArgs& a, int dtype) {
    if (stream_disabled() || (dtype != LMNET_BF16 && dtype != LMNET_F16)) return LMNET_ERR_UNSUPPORTED;
    const NAGeom& g = a.g;
    if (g.K != 3 && g.K != 5 && g.K != 7) return LMNET_ERR_UNSUPPORTED;
    Plan p;
    if (!make_plan(g, false, 0, stream_fwd_minblocks(g.K, stream_hg(g.K, g.D)), p)) return LMNET_ERR_UNSUPPORTED;
    const int vb = p.HG * g.D * 2;
    if (!strides_ok(a.k, g.D, vb) || !strides_ok(a.v, g.D, vb) || !strides_ok(a.q, g.D, vb) ||
        !strides_ok(a.out, g.D, vb))
        return LMNET_ERR_UNSUPPORTED;
    return dtype == LMNET_BF16 ? fwd_typed<__nv_bfloat16>(a, p) : fwd_typed<__half>(a, p);
}

int stream_bwd(const FusedArgs& a, int dtype, int64_t max_parts, int64_t* n_parts) {
    if (stream_disabled() || (dtype != LMNET_BF16 && dtype != LMNET_F16)) return LMNET_ERR_UNSUPPORTED;
    const NAGeom& g = a.g;
    if (g.K != 3) return LMNET_ERR_UNSUPPORTED;
    Plan p;
    const int resident = g.D == 4 ? 3 : 2;   // StreamBwdSmem::MINB
    if (!make_plan(g, true, max_parts, resident, p)) return LMNET_ERR_UNSUPPORTED;
    const int vb = p.HG * g.D * 2;
    const lmnet_view5* staged[] = {a.q, a.k, a.v, a.dout};
    for (auto* x : staged)
        if (!strides_ok(x, g.D, vb)) return LMNET_ERR_UNSUPPORTED;
    const lmnet_view5* written[] = {a.dq, a.dk, a.dv};
    for (auto* x : written)
        if (!strides_ok(x, g.D, vb)) return LMNET_ERR_UNSUPPORTED;
    *n_parts = (int64_t)p.stripes * p.bands * p.z;
    return dtype == LMNET_BF16 ? bwd_typed<__nv_bfloat16>(a, p) : bwd_typed<__half>(a, p);
}

}  // namespace lmnet
