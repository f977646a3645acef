// conv3x3_fast.cu — the 3x3 convolution kernels of conv3x3.cu specialised on (stride, Cin, Cout) at compile time.
//
// Same algorithm, tiles and results as the generic kernels (see conv3x3.cu); what the specialisation buys is issue slots.
// ncu on the generic forward (profiles/r02_ncu_conv_generic.txt): 24 -> 12 channels at 352 x 352 ran 1 360 instructions
// per thread and 16 x 16 tile, of which 8 in 65 inside the tap loop were HMMA, and the staging / copy-out loops spent
// ~25 instructions per 16-byte vector on runtime divisions — issue-bound at 9 % of HBM peak.  With Cin / Cout / stride as
// template parameters every pitch, tile extent and vector count is a constant: the 9-tap x k-step contraction is fully
// unrolled with immediate shared-memory offsets (two ldmatrix.x4 feed four HMMAs), index arithmetic in the staging and
// copy-out loops reduces to multiply-shift, weights are staged with 8-byte loads.
// The list of instantiated shapes is what LM-Net runs on its three largest resolutions (forward and the stride-1 input
// gradients, whose channel counts are swapped); other shapes use the generic kernels.
#include "conv3x3.cuh"

namespace lmnet {

template <int S, int CIN, int COUT>
struct CvCfg {
    static constexpr int KS = (CIN + 15) / 16, NT = (COUT + 7) / 8;
    static constexpr int MW = (S == 1 && NT <= 6) ? 2 : 1;
    static constexpr int TH = kCvWarps * MW, IH = (TH - 1) * S + 3, IW = (kCvTW - 1) * S + 3;
    static constexpr int PX = cv_pitch_c(KS * 16), PW = PX, PO = cv_pitch_c(NT * 8);
    static constexpr int VIN = CIN % 8 == 0 ? 8 : 4, VPP_IN = CIN / VIN;
    static constexpr int VOUT = COUT % 8 == 0 ? 8 : 4, VPP_OUT = COUT / VOUT;
    static constexpr int X_ELEMS = IH * IW * PX, W_ELEMS = 9 * NT * 8 * PW, O_ELEMS = TH * kCvTW * PO;
    static constexpr size_t SMEM = (size_t)(W_ELEMS + 2 * X_ELEMS + O_ELEMS) * 2 + NT * 8 * 4 + 16;
    // weight gradient
    static constexpr int MT = (COUT + 15) / 16, NTC = (CIN + 7) / 8;
    static constexpr int WTH = S == 1 ? 16 : 8, WIH = (WTH - 1) * S + 3;
    static constexpr int WPX = cv_pitch_c(NTC * 8), WPD = cv_pitch_c(MT * 16);
    static constexpr int WX_ELEMS = WIH * IW * WPX, WD_ELEMS = WTH * kCvTW * WPD;
    static constexpr size_t WSMEM = (size_t)2 * (WX_ELEMS + WD_ELEMS) * 2 + 16;
};

__device__ __forceinline__ uint32_t cv_saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cv_ldsm_x4_a(uint32_t (&r)[4], uint32_t a) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void cv_ldsm_x2_a(uint32_t (&r)[2], uint32_t a) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
__device__ __forceinline__ void cv_ldsm_x4_ta(uint32_t (&r)[4], uint32_t a) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void cv_ldsm_x2_ta(uint32_t (&r)[2], uint32_t a) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
__device__ __forceinline__ void cv_cp_a(uint32_t s, const void* gmem, bool valid, int bytes16) {
    const int sz = valid ? (bytes16 ? 16 : 8) : 0;
    if (bytes16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz));
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(sz));
}

// stage IHT x IW input pixels of CIN channels (pixel pitch PX elements) around output tile `tl`; zero outside the image
template <typename T, int NTHREADS, int S, int CIN, int IHT, int IW, int PX>
__device__ __forceinline__ void cvf_issue_x(uint32_t s, const T* __restrict__ x, const CvGeom& g, const CvTile& tl) {
    constexpr int VIN = CIN % 8 == 0 ? 8 : 4, VPP = CIN / VIN, TOTAL = IHT * IW * VPP;
    const int iy0 = tl.oy0 * S - 1, ix0 = tl.ox0 * S - 1;
    const T* xb = x + (int64_t)tl.b * g.H * g.W * g.gx_pitch;
#pragma unroll 4
    for (int i = threadIdx.x; i < TOTAL; i += NTHREADS) {
        const int pix = i / VPP, v = i - pix * VPP;
        const int r = pix / IW, c = pix - r * IW;
        const int iy = iy0 + r, ix = ix0 + c;
        const bool ok = (unsigned)iy < (unsigned)g.H && (unsigned)ix < (unsigned)g.W;
        cv_cp_a(s + (pix * PX + v * VIN) * 2, ok ? xb + ((int64_t)iy * g.W + ix) * g.gx_pitch + v * VIN : x, ok, VIN == 8);
    }
}

template <typename T, int S, int CIN, int COUT>
__global__ void __launch_bounds__(kCvThreads)
conv3x3_fwd_fast_kernel(const T* __restrict__ x, const float* __restrict__ w /* fp32 parameter, see cv_weight */, int w_t,
                        const float* __restrict__ bias, T* __restrict__ y, CvGeom g) {
    using C = CvCfg<S, CIN, COUT>;
    constexpr int KS = C::KS, NT = C::NT, MW = C::MW, PX = C::PX, PW = C::PW, PO = C::PO, IW = C::IW, n_pad = NT * 8;
    extern __shared__ __align__(16) unsigned char cv_smem[];
    T* s_w = reinterpret_cast<T*>(cv_smem);
    T* s_x = s_w + C::W_ELEMS;
    T* s_o = s_x + 2 * C::X_ELEMS;
    float* s_bias = reinterpret_cast<float*>(s_o + C::O_ELEMS);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gq = lane >> 2, tq = lane & 3;

    {   // weights [9][n_pad][PW] from the fp32 parameter (coalesced reads, rounded here), zero padding
        uint4* zw = reinterpret_cast<uint4*>(s_w);
        for (int i = threadIdx.x; i < C::W_ELEMS / 8; i += kCvThreads) zw[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        cv_stage_weights<T, kCvThreads>(s_w, w, w_t, CIN, COUT, n_pad, PW);
        for (int i = threadIdx.x; i < n_pad; i += kCvThreads) s_bias[i] = (bias != nullptr && i < COUT) ? bias[i] : 0.f;
        uint4* z = reinterpret_cast<uint4*>(s_x);
        for (int i = threadIdx.x; i < 2 * C::X_ELEMS / 8; i += kCvThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();

    const uint32_t sx_base = cv_saddr(s_x), sw_base = cv_saddr(s_w);
    const int m = lane >> 3, rr = lane & 7;
    // A: row = pixel (m & 1) * 8 + rr of the warp's output row, k half (m >> 1)
    const uint32_t a_lane = (uint32_t)(((warp * MW * S) * IW + ((m & 1) * 8 + rr) * S) * PX + (m >> 1) * 8) * 2;
    // B (x4 = two channel tiles): row = channel (lane >> 4) * 8 + (lane & 7), k half (lane >> 3) & 1
    const uint32_t b_lane = (uint32_t)((((lane >> 4) * 8 + (lane & 7)) * PW) + ((lane >> 3) & 1) * 8) * 2;
    const uint32_t b_lane2 = (uint32_t)(((lane & 7) * PW) + ((lane >> 3) & 1) * 8) * 2;       // x2: lanes 0..15 matter

    const int first = blockIdx.x, step = gridDim.x;
    if (first < g.tiles) {
        cvf_issue_x<T, kCvThreads, S, CIN, C::IH, IW, PX>(sx_base, x, g, cv_tile(g, first));
        cv_commit();
    }
    int st = 0;
    for (int t = first; t < g.tiles; t += step, st ^= 1) {
        const CvTile tl = cv_tile(g, t);
        if (t + step < g.tiles) {
            cvf_issue_x<T, kCvThreads, S, CIN, C::IH, IW, PX>(sx_base + (st ^ 1) * C::X_ELEMS * 2, x, g, cv_tile(g, t + step));
            cv_commit();
            cv_wait<1>();
        } else {
            cv_wait<0>();
        }
        __syncthreads();
        const uint32_t sa = sx_base + st * C::X_ELEMS * 2 + a_lane;
        float acc[MW][NT][4];
#pragma unroll
        for (int i = 0; i < MW; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int ky = tap / 3, kx = tap - ky * 3;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                uint32_t bf[NT][2];
#pragma unroll
                for (int j = 0; j + 1 < NT; j += 2) {
                    uint32_t q[4];
                    cv_ldsm_x4_a(q, sw_base + b_lane + (uint32_t)(((tap * n_pad + j * 8) * PW + ks * 16) * 2));
                    bf[j][0] = q[0]; bf[j][1] = q[1]; bf[j + 1][0] = q[2]; bf[j + 1][1] = q[3];
                }
                if constexpr (NT & 1)
                    cv_ldsm_x2_a(bf[NT - 1], sw_base + b_lane2 + (uint32_t)(((tap * n_pad + (NT - 1) * 8) * PW + ks * 16) * 2));
#pragma unroll
                for (int i = 0; i < MW; ++i) {
                    uint32_t af[4];
                    cv_ldsm_x4_a(af, sa + (uint32_t)((((i * S + ky) * IW + kx) * PX + ks * 16) * 2));
#pragma unroll
                    for (int j = 0; j < NT; ++j) cv_mma<T>(acc[i][j], af, bf[j]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < MW; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const int n = j * 8 + 2 * tq;
                const float b0 = s_bias[n], b1 = s_bias[n + 1];
                const int p0 = (warp * MW + i) * kCvTW + gq;
                *reinterpret_cast<uint32_t*>(s_o + p0 * PO + n) = cv_pack<T>(acc[i][j][0] + b0, acc[i][j][1] + b1);
                *reinterpret_cast<uint32_t*>(s_o + (p0 + 8) * PO + n) = cv_pack<T>(acc[i][j][2] + b0, acc[i][j][3] + b1);
            }
        __syncthreads();
        const int vw = min(kCvTW, g.Wo - tl.ox0);
        T* yb = y + ((int64_t)tl.b * g.Ho * g.Wo) * COUT;
        constexpr int VOUT = C::VOUT, VPP = C::VPP_OUT, TOTAL = C::TH * kCvTW * VPP;
#pragma unroll 2
        for (int i = threadIdx.x; i < TOTAL; i += kCvThreads) {
            const int pix = i / VPP, v = i - pix * VPP;
            const int r = pix >> 4, c = pix & 15;
            const int oy = tl.oy0 + r;
            if (oy < g.Ho && c < vw) {
                T* dst = yb + ((int64_t)oy * g.Wo + tl.ox0 + c) * COUT + v * VOUT;
                const T* src = s_o + pix * PO + v * VOUT;
                if constexpr (VOUT == 8) *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
                else *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(src);
            }
        }
    }
}

// weight gradient, warp = tap (see conv3x3.cu); partial layout [cta][tap][MT*16][NTC*8], part_b [cta][MT*16]
template <typename T, int S, int CIN, int COUT>
__global__ void __launch_bounds__(kCvWgThreads)
conv3x3_wgrad_fast_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ part, float* __restrict__ part_b,
                          CvGeom g) {
    using C = CvCfg<S, CIN, COUT>;
    constexpr int MT = C::MT, NTC = C::NTC, TH = C::WTH, IW = C::IW, PX = C::WPX, PD = C::WPD;
    constexpr int STAGE = C::WX_ELEMS + C::WD_ELEMS;
    extern __shared__ __align__(16) unsigned char cv_smem[];
    T* s_all = reinterpret_cast<T*>(cv_smem);
    const int lane = threadIdx.x & 31, tap = threadIdx.x >> 5;
    const int ky = tap / 3, kx = tap - ky * 3;
    {
        uint4* z = reinterpret_cast<uint4*>(s_all);
        for (int i = threadIdx.x; i < 2 * STAGE / 8; i += kCvWgThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    const uint32_t s_base = cv_saddr(s_all);

    auto issue = [&](int t, int st) {
        const CvTile tl = cv_tile(g, t);
        const uint32_t sx = s_base + st * STAGE * 2;
        cvf_issue_x<T, kCvWgThreads, S, CIN, C::WIH, IW, PX>(sx, x, g, tl);
        const uint32_t sd = sx + C::WX_ELEMS * 2;
        const T* db = dy + (int64_t)tl.b * g.Ho * g.Wo * g.gd_pitch;
        constexpr int VOUT = C::VOUT, VPP = C::VPP_OUT, TOTAL = TH * kCvTW * VPP;
#pragma unroll 2
        for (int i = threadIdx.x; i < TOTAL; i += kCvWgThreads) {
            const int pix = i / VPP, v = i - pix * VPP;
            const int oy = tl.oy0 + (pix >> 4), ox = tl.ox0 + (pix & 15);
            const bool ok = oy < g.Ho && ox < g.Wo;
            cv_cp_a(sd + (pix * PD + v * VOUT) * 2, ok ? db + ((int64_t)oy * g.Wo + ox) * g.gd_pitch + v * VOUT : dy, ok, VOUT == 8);
        }
        cv_commit();
    };

    float acc[MT][NTC][4];
    float accb[MT][4];
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        accb[i][0] = accb[i][1] = accb[i][2] = accb[i][3] = 0.f;
#pragma unroll
        for (int j = 0; j < NTC; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
    }
    const uint32_t ones[2] = {cv_ones<T>(), cv_ones<T>()};
    const int m = lane >> 3, rr = lane & 7;
    const uint32_t a_lane = (uint32_t)(((m >> 1) * 8 + rr) * PD + (m & 1) * 8) * 2;                       // dy: rows = pixels, cols = Cout
    const uint32_t b_lane = (uint32_t)((ky * IW + kx + (lane & 15) * S) * PX + (lane >> 4) * 8) * 2;      // x4: two Cin tiles
    const uint32_t b_lane2 = (uint32_t)((ky * IW + kx + (lane & 15) * S) * PX) * 2;

    const int first = blockIdx.x, step = gridDim.x;
    if (first < g.tiles) issue(first, 0);
    int st = 0;
    for (int t = first; t < g.tiles; t += step, st ^= 1) {
        if (t + step < g.tiles) {
            issue(t + step, st ^ 1);
            cv_wait<1>();
        } else {
            cv_wait<0>();
        }
        __syncthreads();
        const uint32_t sx = s_base + st * STAGE * 2;
        const uint32_t sd = sx + C::WX_ELEMS * 2;
#pragma unroll
        for (int r = 0; r < TH; ++r) {
            uint32_t af[MT][4];
#pragma unroll
            for (int i = 0; i < MT; ++i) cv_ldsm_x4_ta(af[i], sd + a_lane + (uint32_t)((r * kCvTW * PD + i * 16) * 2));
#pragma unroll
            for (int j = 0; j + 1 < NTC; j += 2) {
                uint32_t q[4];
                cv_ldsm_x4_ta(q, sx + b_lane + (uint32_t)((r * S * IW * PX + j * 8) * 2));
                const uint32_t b0[2] = {q[0], q[1]}, b1[2] = {q[2], q[3]};
#pragma unroll
                for (int i = 0; i < MT; ++i) {
                    cv_mma<T>(acc[i][j], af[i], b0);
                    cv_mma<T>(acc[i][j + 1], af[i], b1);
                }
            }
            if constexpr (NTC & 1) {
                uint32_t b[2];
                cv_ldsm_x2_ta(b, sx + b_lane2 + (uint32_t)((r * S * IW * PX + (NTC - 1) * 8) * 2));
#pragma unroll
                for (int i = 0; i < MT; ++i) cv_mma<T>(acc[i][NTC - 1], af[i], b);
            }
            if (tap == 4) {
#pragma unroll
                for (int i = 0; i < MT; ++i) cv_mma<T>(accb[i], af[i], ones);
            }
        }
        __syncthreads();
    }
    const int gq = lane >> 2, tq = lane & 3;
    constexpr int ldn = NTC * 8;
    float* out = part + ((int64_t)blockIdx.x * 9 + tap) * (MT * 16) * ldn;
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NTC; ++j) {
            const int n = j * 8 + 2 * tq, m0 = i * 16 + gq;
            *reinterpret_cast<float2*>(out + m0 * ldn + n) = make_float2(acc[i][j][0], acc[i][j][1]);
            *reinterpret_cast<float2*>(out + (m0 + 8) * ldn + n) = make_float2(acc[i][j][2], acc[i][j][3]);
        }
    if (tap == 4 && tq == 0) {
#pragma unroll
        for (int i = 0; i < MT; ++i) {
            part_b[(int64_t)blockIdx.x * (MT * 16) + i * 16 + gq] = accb[i][0];
            part_b[(int64_t)blockIdx.x * (MT * 16) + i * 16 + gq + 8] = accb[i][2];
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// host side: shape list and dispatch
// ---------------------------------------------------------------------------------------------------
// (stride, Cin, Cout): forward shapes of LM-Net at 352 / 176 / 88 pixels and their stride-1 input gradients (swapped)
#define CV_FAST_FWD_LIST(X) \
    X(1, 12, 12) X(1, 24, 12) X(1, 12, 24) X(1, 24, 24) X(1, 48, 24) X(1, 24, 48) X(1, 72, 24) X(1, 24, 72) X(1, 48, 48) \
    X(2, 12, 24) X(2, 24, 48)
#define CV_FAST_WG_LIST(X) \
    X(1, 12, 12) X(1, 24, 12) X(1, 24, 24) X(1, 48, 24) X(1, 72, 24) X(1, 48, 48) X(2, 12, 24) X(2, 24, 48)

bool cv_fast_fwd_has(int S, int Cin, int Cout) {
#define X(SS, CI, CO) if (S == SS && Cin == CI && Cout == CO) return true;
    CV_FAST_FWD_LIST(X)
#undef X
    return false;
}
bool cv_fast_wgrad_has(int S, int Cin, int Cout) {
#define X(SS, CI, CO) if (S == SS && Cin == CI && Cout == CO) return true;
    CV_FAST_WG_LIST(X)
#undef X
    return false;
}

static void cvf_geom(const lmnet_conv3x3_dims* d, int TH, CvGeom& g) {
    g.B = d->B; g.H = d->H; g.W = d->W; g.Cin = d->Cin; g.Cout = d->Cout; g.S = d->stride;
    g.Ho = (d->H - 1) / d->stride + 1;
    g.Wo = (d->W - 1) / d->stride + 1;
    g.TH = TH;
    g.IH = (TH - 1) * g.S + 3;
    g.IW = (kCvTW - 1) * g.S + 3;
    g.tiles_x = (g.Wo + kCvTW - 1) / kCvTW;
    g.tiles_y = (g.Ho + TH - 1) / TH;
    g.tiles = g.B * g.tiles_x * g.tiles_y;
    g.ksteps = (d->Cin + 15) / 16;
    g.pitch_x = g.pitch_w = g.pitch_o = 0;
    g.gx_pitch = d->Cin; g.gd_pitch = d->Cout;
}

static int cvf_grid(int tiles, size_t smem, int cap) {
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm > cap) per_sm = cap;
    if (per_sm < 1) per_sm = 1;
    return tiles < 148 * per_sm ? tiles : 148 * per_sm;
}

template <typename T, int S, int CIN, int COUT>
static int cvf_fwd_launch(const void* x, const float* w, int w_t, const float* bias, void* y, const lmnet_conv3x3_dims* d, cudaStream_t st,
                          int x_pitch) {
    using C = CvCfg<S, CIN, COUT>;
    auto kern = conv3x3_fwd_fast_kernel<T, S, CIN, COUT>;
    static std::atomic<size_t> granted[kMaxDevices];
    if (!ensure_smem(kern, C::SMEM, granted)) return LMNET_ERR_LAUNCH;
    CvGeom g;
    cvf_geom(d, C::TH, g);
    if (x_pitch > 0) g.gx_pitch = x_pitch;
    g.ncta = cvf_grid(g.tiles, C::SMEM, 4);
    const double bytes = ((double)g.B * g.H * g.W * g.Cin + (double)g.B * g.Ho * g.Wo * g.Cout) * sizeof(T);
    LMNET_LAUNCH(KID_CONV3X3, st, bytes, (kern<<<g.ncta, kCvThreads, C::SMEM, st>>>((const T*)x, w, w_t, bias, (T*)y, g)));
    return LMNET_OK;
}

int cv_fast_fwd(const void* x, const float* w, int w_t, const float* bias, void* y, const lmnet_conv3x3_dims* d, int dtype, cudaStream_t st,
                int x_pitch) {
#define X(SS, CI, CO)                                                                                              \
    if (d->stride == SS && d->Cin == CI && d->Cout == CO)                                                          \
        return dtype == LMNET_BF16 ? cvf_fwd_launch<__nv_bfloat16, SS, CI, CO>(x, w, w_t, bias, y, d, st, x_pitch) \
                                   : cvf_fwd_launch<__half, SS, CI, CO>(x, w, w_t, bias, y, d, st, x_pitch);
    CV_FAST_FWD_LIST(X)
#undef X
    return LMNET_ERR_UNSUPPORTED;
}

template <int S, int CIN, int COUT>
static int cvf_wg_grid(const lmnet_conv3x3_dims* d, int* mp, int* ldn) {
    using C = CvCfg<S, CIN, COUT>;
    CvGeom g;
    cvf_geom(d, C::WTH, g);
    *mp = C::MT * 16;
    *ldn = C::NTC * 8;
    return cvf_grid(g.tiles, C::WSMEM, 2);
}

int cv_fast_wgrad_grid(const lmnet_conv3x3_dims* d, int* mp, int* ldn) {
#define X(SS, CI, CO) if (d->stride == SS && d->Cin == CI && d->Cout == CO) return cvf_wg_grid<SS, CI, CO>(d, mp, ldn);
    CV_FAST_WG_LIST(X)
#undef X
    return 0;
}

template <typename T, int S, int CIN, int COUT>
static int cvf_wg_launch(const void* x, const void* dy, float* part, float* part_b, const lmnet_conv3x3_dims* d, cudaStream_t st,
                         int dy_pitch) {
    using C = CvCfg<S, CIN, COUT>;
    auto kern = conv3x3_wgrad_fast_kernel<T, S, CIN, COUT>;
    static std::atomic<size_t> granted[kMaxDevices];
    if (!ensure_smem(kern, C::WSMEM, granted)) return LMNET_ERR_LAUNCH;
    CvGeom g;
    cvf_geom(d, C::WTH, g);
    if (dy_pitch > 0) g.gd_pitch = dy_pitch;
    g.ncta = cvf_grid(g.tiles, C::WSMEM, 2);
    const double bytes = ((double)g.B * g.H * g.W * g.Cin + (double)g.B * g.Ho * g.Wo * g.Cout) * sizeof(T);
    LMNET_LAUNCH(KID_CONV3X3_WGRAD, st, bytes, (kern<<<g.ncta, kCvWgThreads, C::WSMEM, st>>>((const T*)x, (const T*)dy, part, part_b, g)));
    return LMNET_OK;
}

int cv_fast_wgrad(const void* x, const void* dy, float* part, float* part_b, const lmnet_conv3x3_dims* d, int dtype, cudaStream_t st,
                  int dy_pitch) {
#define X(SS, CI, CO)                                                                                              \
    if (d->stride == SS && d->Cin == CI && d->Cout == CO)                                                          \
        return dtype == LMNET_BF16 ? cvf_wg_launch<__nv_bfloat16, SS, CI, CO>(x, dy, part, part_b, d, st, dy_pitch) \
                                   : cvf_wg_launch<__half, SS, CI, CO>(x, dy, part, part_b, d, st, dy_pitch);
    CV_FAST_WG_LIST(X)
#undef X
    return LMNET_ERR_UNSUPPORTED;
}

}  // namespace lmnet
