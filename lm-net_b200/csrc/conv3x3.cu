// conv3x3.cu — dense 3x3 convolutions (padding 1, stride 1 or 2) on channels-last 16-bit tensors, sm_100a.
//
// Widening step f3 of SURVEY.md §8: the stride-2 down-sampling convolutions, the skip-fusion convolutions, the decoder
// convolutions after the bilinear up-sampling and the patch embeddings of the neighbourhood transformers
// (/root/reference/core/LM_Net.py:14-39, 58-74; /root/reference/core/modules.py:30-39, 83-143).  At LM-Net's widths
// (12 .. 72 input channels on the two largest resolutions) these are streaming kernels: 24 .. 144 B of input per pixel
// against <= 62 kFLOP, i.e. below the tensor ridge, and cuDNN pays for its 8-channel alignment with a padded copy of
// every 12-channel tensor (nhwcAddPadding, 1.8 ms of the stock step) plus separate bias-gradient reductions.
//
// Forward (also the stride-1 input gradient, called with flipped / transposed weights):
//   * persistent CTAs walk output tiles of TH x 16 pixels; the input tile + 1-pixel halo is staged by cp.async into a
//     two-stage shared-memory ring (zero-fill for out-of-image pixels = the convolution's padding), pixel-major with a
//     conflict-free pitch; 12-channel tensors are staged with 8-byte copies, so no padded copy exists anywhere;
//   * the whole weight tensor [9][Cout][Cin] lives in shared memory for the lifetime of the CTA;
//   * implicit GEMM on mma.sync m16n8k16: M = 16 pixels of an output row, N = output channels, K = input channels of one
//     tap; the A fragment of tap (ky, kx) is the same shared tile read at a shifted pixel address (ldmatrix takes one
//     address per row, so stride 2 is just a different row pitch);
//   * bias in the epilogue, output staged through shared memory and written as whole pixel rows.
// Weight gradient:
//   * same tiles (dy tile + x halo tile); nine warps, one per tap: dW_tap[co][ci] += sum_pixels dy[p][co] x[p + tap][ci],
//     both operands read with ldmatrix.trans straight from the pixel-major tiles; fp32 accumulators stay in registers
//     across all tiles of the CTA; bias gradient from an all-ones B fragment on the centre-tap warp;
//   * per-CTA partials + a fixed-order reduction kernel: deterministic, no atomics.
// mma.sync rather than tcgen05: N = 12 .. 48 output channels and K = 12 .. 72 per tap are far below a 128 x N x 16 UMMA
// tile, and the kernels are bounded by HBM / issue, not by the tensor pipe.
#include "conv3x3.cuh"

namespace lmnet {

// ---------------------------------------------------------------------------------------------------
// forward: y[b, oy, ox, :] = bias + sum_tap W[tap] . x[b, oy*S + ky - 1, ox*S + kx - 1, :]
// warp w owns output rows w*MW .. w*MW + MW - 1 of the tile (16 pixels each) and all NT channel tiles
// ---------------------------------------------------------------------------------------------------
template <typename T, int S, int MW, int NT>
__global__ void __launch_bounds__(kCvThreads)
conv3x3_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w /* fp32 parameter, see cv_weight */, int w_t,
                   const float* __restrict__ bias, T* __restrict__ y, CvGeom g) {
    extern __shared__ __align__(16) unsigned char cv_smem[];
    constexpr int n_pad = NT * 8;
    T* s_w = reinterpret_cast<T*>(cv_smem);                       // [9][n_pad][pitch_w]
    T* s_x = s_w + 9 * n_pad * g.pitch_w;                         // 2 stages of [IH * IW][pitch_x]
    const int x_elems = g.IH * g.IW * g.pitch_x;
    T* s_o = s_x + 2 * x_elems;                                   // [TH * 16][pitch_o]
    float* s_bias = reinterpret_cast<float*>(s_o + g.TH * kCvTW * g.pitch_o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    const T zero = from_f<T>(0.f);

    for (int i = threadIdx.x; i < 9 * n_pad * g.pitch_w; i += kCvThreads) s_w[i] = zero;
    __syncthreads();
    cv_stage_weights<T, kCvThreads>(s_w, w, w_t, g.Cin, g.Cout, n_pad, g.pitch_w);
    for (int i = threadIdx.x; i < n_pad; i += kCvThreads) s_bias[i] = (bias != nullptr && i < g.Cout) ? bias[i] : 0.f;
    for (int i = threadIdx.x; i < 2 * x_elems; i += kCvThreads) s_x[i] = zero;       // channel padding stays zero
    __syncthreads();

    const int first = blockIdx.x, step = gridDim.x;
    if (first < g.tiles) {
        cv_issue_x<T, kCvThreads>(s_x, x, g, cv_tile(g, first));
        cv_commit();
    }
    int st = 0;
    for (int t = first; t < g.tiles; t += step, st ^= 1) {
        const CvTile tl = cv_tile(g, t);
        if (t + step < g.tiles) {
            cv_issue_x<T, kCvThreads>(s_x + (st ^ 1) * x_elems, x, g, cv_tile(g, t + step));
            cv_commit();
            cv_wait<1>();
        } else {
            cv_wait<0>();
        }
        __syncthreads();                                          // tile t landed; s_o of the previous tile is drained
        const T* sx = s_x + st * x_elems;
        float acc[MW][NT][4];
#pragma unroll
        for (int i = 0; i < MW; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
        const int m = lane >> 3, rr = lane & 7, l16 = lane & 15;
        // A row of this lane: pixel (m & 1) * 8 + rr of the output row, k half (m >> 1)
        const int a_off = ((m & 1) * 8 + rr) * S * g.pitch_x + (m >> 1) * 8;
        const int b_off = (l16 & 7) * g.pitch_w + (l16 >> 3) * 8;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
            const int ky = tap / 3, kx = tap - ky * 3;
            const T* wt = s_w + tap * n_pad * g.pitch_w + b_off;
            for (int ks = 0; ks < g.ksteps; ++ks) {
                uint32_t bf[NT][2];
#pragma unroll
                for (int j = 0; j < NT; ++j) cv_ldsm_x2(bf[j], wt + j * 8 * g.pitch_w + ks * 16);
#pragma unroll
                for (int i = 0; i < MW; ++i) {
                    uint32_t af[4];
                    const int row = (warp * MW + i) * S + ky;
                    cv_ldsm_x4(af, sx + (row * g.IW + kx) * g.pitch_x + a_off + ks * 16);
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
                *reinterpret_cast<uint32_t*>(s_o + p0 * g.pitch_o + n) = cv_pack<T>(acc[i][j][0] + b0, acc[i][j][1] + b1);
                *reinterpret_cast<uint32_t*>(s_o + (p0 + 8) * g.pitch_o + n) = cv_pack<T>(acc[i][j][2] + b0, acc[i][j][3] + b1);
            }
        __syncthreads();
        // staged tile -> global: each tile row is a contiguous run of valid pixels x Cout elements
        const int vw = min(kCvTW, g.Wo - tl.ox0);
        T* yb = y + ((int64_t)tl.b * g.Ho * g.Wo) * g.Cout;
        if ((g.Cout & 7) == 0) {
            const int vpp = g.Cout >> 3;
            for (int i = threadIdx.x; i < g.TH * kCvTW * vpp; i += kCvThreads) {
                const int pix = i / vpp, v = i - pix * vpp;
                const int r = pix >> 4, c = pix & 15;
                const int oy = tl.oy0 + r;
                if (oy < g.Ho && c < vw)
                    *reinterpret_cast<uint4*>(yb + ((int64_t)oy * g.Wo + tl.ox0 + c) * g.Cout + v * 8) =
                        *reinterpret_cast<const uint4*>(s_o + pix * g.pitch_o + v * 8);
            }
        } else {
            const int vpp = g.Cout >> 2;
            for (int i = threadIdx.x; i < g.TH * kCvTW * vpp; i += kCvThreads) {
                const int pix = i / vpp, v = i - pix * vpp;
                const int r = pix >> 4, c = pix & 15;
                const int oy = tl.oy0 + r;
                if (oy < g.Ho && c < vw)
                    *reinterpret_cast<uint2*>(yb + ((int64_t)oy * g.Wo + tl.ox0 + c) * g.Cout + v * 4) =
                        *reinterpret_cast<const uint2*>(s_o + pix * g.pitch_o + v * 4);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// weight gradient: part[cta][tap][co][ci] = sum over the CTA's tiles of dy[p][co] * x[p*S + tap - 1][ci];
// dbias partials part_b[cta][co].  Warp = tap.  MT = ceil(Cout / 16) M tiles, NTC = ceil(Cin / 8) N tiles.
// ---------------------------------------------------------------------------------------------------
template <typename T, int S, int MT, int NTC>
__global__ void __launch_bounds__(kCvWgThreads)
conv3x3_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ part, float* __restrict__ part_b,
                     CvGeom g) {
    extern __shared__ __align__(16) unsigned char cv_smem[];
    T* s_x = reinterpret_cast<T*>(cv_smem);                       // 2 stages of [IH*IW][pitch_x] + [TH*16][pitch_o]
    const int x_elems = g.IH * g.IW * g.pitch_x;
    const int d_elems = g.TH * kCvTW * g.pitch_o;
    const int stage = x_elems + d_elems;
    const int lane = threadIdx.x & 31, tap = threadIdx.x >> 5;
    const int ky = tap / 3, kx = tap - ky * 3;
    const T zero = from_f<T>(0.f);
    for (int i = threadIdx.x; i < 2 * stage; i += kCvWgThreads) s_x[i] = zero;
    __syncthreads();

    auto issue = [&](int t, int st) {
        const CvTile tl = cv_tile(g, t);
        T* sx = s_x + st * stage;
        cv_issue_x<T, kCvWgThreads>(sx, x, g, tl);
        T* sd = sx + x_elems;
        const T* db = dy + (int64_t)tl.b * g.Ho * g.Wo * g.gd_pitch;
        if ((g.Cout & 7) == 0) {
            const int vpp = g.Cout >> 3;
            for (int i = threadIdx.x; i < g.TH * kCvTW * vpp; i += kCvWgThreads) {
                const int pix = i / vpp, v = i - pix * vpp;
                const int oy = tl.oy0 + (pix >> 4), ox = tl.ox0 + (pix & 15);
                const bool ok = oy < g.Ho && ox < g.Wo;
                cv_cp16(sd + pix * g.pitch_o + v * 8, ok ? db + ((int64_t)oy * g.Wo + ox) * g.gd_pitch + v * 8 : dy, ok);
            }
        } else {
            const int vpp = g.Cout >> 2;
            for (int i = threadIdx.x; i < g.TH * kCvTW * vpp; i += kCvWgThreads) {
                const int pix = i / vpp, v = i - pix * vpp;
                const int oy = tl.oy0 + (pix >> 4), ox = tl.ox0 + (pix & 15);
                const bool ok = oy < g.Ho && ox < g.Wo;
                cv_cp8(sd + pix * g.pitch_o + v * 4, ok ? db + ((int64_t)oy * g.Wo + ox) * g.gd_pitch + v * 4 : dy, ok);
            }
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
    const int m = lane >> 3, rr = lane & 7, l16 = lane & 15;
    const int a_off = ((m >> 1) * 8 + rr) * g.pitch_o + (m & 1) * 8;      // dy: pixel rows (k), channel columns (m)
    const int b_off = l16 * S * g.pitch_x;                                // x: pixel rows (k), channel columns (n)

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
        const T* sx = s_x + st * stage;
        const T* sd = sx + x_elems;
#pragma unroll 2
        for (int r = 0; r < g.TH; ++r) {
            uint32_t af[MT][4];
#pragma unroll
            for (int i = 0; i < MT; ++i) cv_ldsm_x4_t(af[i], sd + r * kCvTW * g.pitch_o + a_off + i * 16);
            const T* xr = sx + ((r * S + ky) * g.IW + kx) * g.pitch_x + b_off;
#pragma unroll
            for (int j = 0; j < NTC; ++j) {
                uint32_t bf[2];
                cv_ldsm_x2_t(bf, xr + j * 8);
#pragma unroll
                for (int i = 0; i < MT; ++i) cv_mma<T>(acc[i][j], af[i], bf);
            }
            if (tap == 4) {
#pragma unroll
                for (int i = 0; i < MT; ++i) cv_mma<T>(accb[i], af[i], ones);
            }
        }
        __syncthreads();                                          // everybody is done with this stage before it is refilled
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

// dW[co][ci][ky][kx] = sum_cta part[cta][tap][co][ci];  db[co] = sum_cta part_b[cta][co].  64 outputs per block, four
// threads per output each summing every fourth partial, combined in a fixed order (deterministic)
__global__ void __launch_bounds__(256)
conv3x3_wgrad_reduce_kernel(const float* __restrict__ part, const float* __restrict__ part_b, float* __restrict__ dW,
                            float* __restrict__ db, int ncta, int Cout, int Cin, int Mp, int ldn) {
    __shared__ float s_red[4][64];
    const int o = threadIdx.x & 63, lane = threadIdx.x >> 6;
    const int i = blockIdx.x * 64 + o;
    const int nW = 9 * Cout * Cin;
    float s = 0.f;
    int tap = 0, co = 0, ci = 0;
    if (i < nW) {
        tap = i / (Cout * Cin);
        const int rem = i - tap * Cout * Cin;
        co = rem / Cin;
        ci = rem - co * Cin;
        const float* p = part + ((int64_t)tap * Mp + co) * ldn + ci;
        const int64_t stride = (int64_t)9 * Mp * ldn;
        float s0 = 0.f, s1 = 0.f;
        int c = lane;
        for (; c + 4 < ncta; c += 8) { s0 += p[c * stride]; s1 += p[(c + 4) * stride]; }
        if (c < ncta) s0 += p[c * stride];
        s = s0 + s1;
    } else if (i < nW + Cout && db != nullptr) {
        for (int c = lane; c < ncta; c += 4) s += part_b[(int64_t)c * Mp + (i - nW)];
    }
    s_red[lane][o] = s;
    __syncthreads();
    if (lane == 0) {
        const float t = (s_red[0][o] + s_red[1][o]) + (s_red[2][o] + s_red[3][o]);
        if (i < nW) dW[((int64_t)co * Cin + ci) * 9 + tap] = t;
        else if (i < nW + Cout && db != nullptr) db[i - nW] = t;
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static int cv_pitch(int cols) {            // multiple of 8 elements, == 8 (mod 16): conflict-free ldmatrix rows
    int p = (cols + 7) / 8 * 8;
    if (p % 16 != 8) p += 8;
    return p;
}

static bool cv_valid(const lmnet_conv3x3_dims* d) {
    if (d == nullptr || d->B <= 0 || d->H <= 0 || d->W <= 0 || d->Cin <= 0 || d->Cout <= 0) return false;
    if (d->stride != 1 && d->stride != 2) return false;
    if (d->Cin % 4 != 0 || d->Cout % 4 != 0) return false;
    if ((int64_t)d->B * d->H * d->W * (d->Cin > d->Cout ? d->Cin : d->Cout) >= ((int64_t)1 << 40)) return false;
    return true;
}

static void cv_base_geom(const lmnet_conv3x3_dims* d, int TH, CvGeom& g) {
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
    g.gx_pitch = d->Cin; g.gd_pitch = d->Cout;
}

struct CvFwdPlan {
    CvGeom g;
    int MW, NT;
    size_t smem;
    int grid;
};
static const int kCvFwdNT[] = {2, 3, 6, 9, 12};

static bool cv_fwd_plan(const lmnet_conv3x3_dims* d, CvFwdPlan& pl) {
    if (!cv_valid(d)) return false;
    const int nt_need = (d->Cout + 7) / 8;
    pl.NT = 0;
    for (int nt : kCvFwdNT)
        if (nt >= nt_need) { pl.NT = nt; break; }
    if (pl.NT == 0) return false;
    pl.MW = (d->stride == 1 && pl.NT <= 6) ? 2 : 1;
    CvGeom& g = pl.g;
    cv_base_geom(d, kCvWarps * pl.MW, g);
    g.pitch_x = cv_pitch(g.ksteps * 16);
    g.pitch_w = cv_pitch(g.ksteps * 16);
    g.pitch_o = cv_pitch(pl.NT * 8);
    const size_t elems = (size_t)9 * pl.NT * 8 * g.pitch_w + 2 * (size_t)g.IH * g.IW * g.pitch_x + (size_t)g.TH * kCvTW * g.pitch_o;
    pl.smem = elems * 2 + (size_t)pl.NT * 8 * 4 + 16;
    if (pl.smem > 200 * 1024) return false;
    int per_sm = (int)((220 * 1024) / (pl.smem + 1024));
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    pl.grid = g.tiles < 148 * per_sm ? g.tiles : 148 * per_sm;
    g.ncta = pl.grid;
    return true;
}

struct CvWgPlan {
    CvGeom g;
    int MT, NTC;
    size_t smem;
    int grid;
};

static bool cv_wg_combo(int MT, int NTC) {
    return (MT == 1 && (NTC == 2 || NTC == 3)) || (MT == 2 && (NTC == 2 || NTC == 3 || NTC == 6 || NTC == 9)) ||
           (MT == 3 && (NTC == 3 || NTC == 6));
}

static bool cv_wg_plan(const lmnet_conv3x3_dims* d, CvWgPlan& pl) {
    if (!cv_valid(d)) return false;
    pl.MT = (d->Cout + 15) / 16;
    pl.NTC = (d->Cin + 7) / 8;
    if (pl.NTC == 4 || pl.NTC == 5) pl.NTC = 6;
    if (pl.NTC == 7 || pl.NTC == 8) pl.NTC = 9;
    if (pl.NTC == 1) pl.NTC = 2;
    if (!cv_wg_combo(pl.MT, pl.NTC)) return false;
    CvGeom& g = pl.g;
    cv_base_geom(d, d->stride == 1 ? 16 : 8, g);
    g.pitch_x = cv_pitch(pl.NTC * 8);
    g.pitch_o = cv_pitch(pl.MT * 16);
    g.pitch_w = 0;
    pl.smem = 2 * ((size_t)g.IH * g.IW * g.pitch_x + (size_t)g.TH * kCvTW * g.pitch_o) * 2 + 16;
    if (pl.smem > 200 * 1024) return false;
    int per_sm = (int)((220 * 1024) / (pl.smem + 1024));
    if (per_sm > 2) per_sm = 2;
    if (per_sm < 1) per_sm = 1;
    pl.grid = g.tiles < 148 * per_sm ? g.tiles : 148 * per_sm;
    g.ncta = pl.grid;
    return true;
}

template <typename T, int S, int MW, int NT>
static int cv_fwd_launch(const void* x, const float* w, int w_t, const float* bias, void* y, const CvFwdPlan& pl, cudaStream_t st) {
    auto kern = conv3x3_fwd_kernel<T, S, MW, NT>;
    static std::atomic<size_t> granted[kMaxDevices];
    if (!ensure_smem(kern, pl.smem, granted)) return LMNET_ERR_LAUNCH;
    const CvGeom& g = pl.g;
    const double bytes = ((double)g.B * g.H * g.W * g.Cin + (double)g.B * g.Ho * g.Wo * g.Cout) * sizeof(T);
    LMNET_LAUNCH(KID_CONV3X3, st, bytes, (kern<<<pl.grid, kCvThreads, pl.smem, st>>>((const T*)x, w, w_t, bias, (T*)y, g)));
    return LMNET_OK;
}

template <typename T>
static int cv_fwd_dispatch(const void* x, const float* w, int w_t, const float* bias, void* y, const CvFwdPlan& pl, cudaStream_t st) {
#define CV_CASE(SS, MWW, NTT) \
    if (pl.g.S == SS && pl.MW == MWW && pl.NT == NTT) return cv_fwd_launch<T, SS, MWW, NTT>(x, w, w_t, bias, y, pl, st);
    CV_CASE(1, 2, 2) CV_CASE(1, 2, 3) CV_CASE(1, 2, 6) CV_CASE(1, 1, 9) CV_CASE(1, 1, 12)
    CV_CASE(2, 1, 2) CV_CASE(2, 1, 3) CV_CASE(2, 1, 6) CV_CASE(2, 1, 9) CV_CASE(2, 1, 12)
#undef CV_CASE
    return LMNET_ERR_UNSUPPORTED;
}

template <typename T, int S, int MT, int NTC>
static int cv_wg_launch(const void* x, const void* dy, float* part, float* part_b, const CvWgPlan& pl, cudaStream_t st) {
    auto kern = conv3x3_wgrad_kernel<T, S, MT, NTC>;
    static std::atomic<size_t> granted[kMaxDevices];
    if (!ensure_smem(kern, pl.smem, granted)) return LMNET_ERR_LAUNCH;
    const CvGeom& g = pl.g;
    const double bytes = ((double)g.B * g.H * g.W * g.Cin + (double)g.B * g.Ho * g.Wo * g.Cout) * sizeof(T);
    LMNET_LAUNCH(KID_CONV3X3_WGRAD, st, bytes, (kern<<<pl.grid, kCvWgThreads, pl.smem, st>>>((const T*)x, (const T*)dy, part, part_b, g)));
    return LMNET_OK;
}

template <typename T>
static int cv_wg_dispatch(const void* x, const void* dy, float* part, float* part_b, const CvWgPlan& pl, cudaStream_t st) {
#define CV_CASE(SS, MTT, NTT) \
    if (pl.g.S == SS && pl.MT == MTT && pl.NTC == NTT) return cv_wg_launch<T, SS, MTT, NTT>(x, dy, part, part_b, pl, st);
    CV_CASE(1, 1, 2) CV_CASE(1, 1, 3) CV_CASE(1, 2, 2) CV_CASE(1, 2, 3) CV_CASE(1, 2, 6) CV_CASE(1, 2, 9) CV_CASE(1, 3, 3) CV_CASE(1, 3, 6)
    CV_CASE(2, 1, 2) CV_CASE(2, 1, 3) CV_CASE(2, 2, 2) CV_CASE(2, 2, 3) CV_CASE(2, 2, 6) CV_CASE(2, 2, 9) CV_CASE(2, 3, 3) CV_CASE(2, 3, 6)
#undef CV_CASE
    return LMNET_ERR_UNSUPPORTED;
}

}  // namespace lmnet

using namespace lmnet;

extern "C" int lmnet_conv3x3_fwd_supported(const lmnet_conv3x3_dims* d, int dtype) {
    if (dtype != LMNET_BF16 && dtype != LMNET_F16) return 0;
    if (cv_valid(d) && cv_fast_fwd_has(d->stride, d->Cin, d->Cout)) return 1;
    CvFwdPlan pl{};
    return cv_fwd_plan(d, pl) ? 1 : 0;
}

// x_pixel_pitch: elements between consecutive pixels of x (0 or Cin = dense; larger = a channel slice of a wider
// channels-last tensor); vectors of 8 (Cin % 8 == 0) or 4 elements must stay aligned
static bool cv_pitch_ok(const void* base, int pitch, int channels) {
    if (pitch == 0 || pitch == channels) return (uintptr_t)base % 16 == 0;
    const int vec_bytes = channels % 8 == 0 ? 16 : 8;
    return pitch > channels && (pitch * 2) % vec_bytes == 0 && (uintptr_t)base % vec_bytes == 0;
}

extern "C" int lmnet_conv3x3_fwd_strided(const void* x, int x_pixel_pitch, const float* weight, int w_transposed, const float* bias,
                                         void* y, const lmnet_conv3x3_dims* d, int dtype, void* stream) {
    if (!lmnet_conv3x3_fwd_supported(d, dtype)) return LMNET_ERR_UNSUPPORTED;
    if (!x || !weight || !y) return LMNET_ERR_INVALID_ARG;
    if (!cv_pitch_ok(x, x_pixel_pitch, d->Cin) || (uintptr_t)y % 16 != 0) return LMNET_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int w_t = w_transposed ? 1 : 0;
    if (cv_fast_fwd_has(d->stride, d->Cin, d->Cout)) return cv_fast_fwd(x, weight, w_t, bias, y, d, dtype, st, x_pixel_pitch);
    CvFwdPlan pl{};
    cv_fwd_plan(d, pl);
    if (x_pixel_pitch > 0) pl.g.gx_pitch = x_pixel_pitch;
    return dtype == LMNET_BF16 ? cv_fwd_dispatch<__nv_bfloat16>(x, weight, w_t, bias, y, pl, st)
                               : cv_fwd_dispatch<__half>(x, weight, w_t, bias, y, pl, st);
}

extern "C" int lmnet_conv3x3_fwd(const void* x, const float* weight, int w_transposed, const float* bias, void* y,
                                 const lmnet_conv3x3_dims* d, int dtype, void* stream) {
    return lmnet_conv3x3_fwd_strided(x, 0, weight, w_transposed, bias, y, d, dtype, stream);
}

extern "C" int lmnet_conv3x3_wgrad_supported(const lmnet_conv3x3_dims* d, int dtype) {
    if (dtype != LMNET_BF16 && dtype != LMNET_F16) return 0;
    if (cv_valid(d) && cv_fast_wgrad_has(d->stride, d->Cin, d->Cout)) return 1;
    CvWgPlan pl{};
    return cv_wg_plan(d, pl) ? 1 : 0;
}

extern "C" size_t lmnet_conv3x3_wgrad_workspace_bytes(const lmnet_conv3x3_dims* d) {
    if (cv_valid(d) && cv_fast_wgrad_has(d->stride, d->Cin, d->Cout)) {
        int mp = 0, ldn = 0;
        const int grid = cv_fast_wgrad_grid(d, &mp, &ldn);
        return ((size_t)grid * 9 * mp * ldn + (size_t)grid * mp) * sizeof(float);
    }
    CvWgPlan pl{};
    if (!cv_wg_plan(d, pl)) return 0;
    return ((size_t)pl.grid * 9 * pl.MT * 16 * pl.NTC * 8 + (size_t)pl.grid * pl.MT * 16) * sizeof(float);
}

extern "C" int lmnet_conv3x3_wgrad_strided(const void* x, const void* dy, int dy_pixel_pitch, float* dW, float* dbias, void* workspace,
                                           size_t workspace_bytes, const lmnet_conv3x3_dims* d, int dtype, void* stream) {
    if (!lmnet_conv3x3_wgrad_supported(d, dtype)) return LMNET_ERR_UNSUPPORTED;
    if (!x || !dy || !dW || !workspace) return LMNET_ERR_INVALID_ARG;
    if ((uintptr_t)x % 16 != 0 || !cv_pitch_ok(dy, dy_pixel_pitch, d->Cout)) return LMNET_ERR_UNSUPPORTED;
    if (workspace_bytes < lmnet_conv3x3_wgrad_workspace_bytes(d)) return LMNET_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int n_out = 9 * d->Cout * d->Cin + d->Cout;
    if (cv_fast_wgrad_has(d->stride, d->Cin, d->Cout)) {
        int mp = 0, ldn = 0;
        const int grid = cv_fast_wgrad_grid(d, &mp, &ldn);
        float* fpart = (float*)workspace;
        float* fpart_b = fpart + (size_t)grid * 9 * mp * ldn;
        const int frc = cv_fast_wgrad(x, dy, fpart, fpart_b, d, dtype, st, dy_pixel_pitch);
        if (frc != LMNET_OK) return frc;
        LMNET_LAUNCH(KID_CONV3X3_REDUCE, st, 0, (conv3x3_wgrad_reduce_kernel<<<(n_out + 63) / 64, 256, 0, st>>>(
            fpart, fpart_b, dW, dbias, grid, d->Cout, d->Cin, mp, ldn)));
        return LMNET_OK;
    }
    CvWgPlan pl{};
    cv_wg_plan(d, pl);
    if (dy_pixel_pitch > 0) pl.g.gd_pitch = dy_pixel_pitch;
    float* part = (float*)workspace;
    float* part_b = part + (size_t)pl.grid * 9 * pl.MT * 16 * pl.NTC * 8;
    int rc = dtype == LMNET_BF16 ? cv_wg_dispatch<__nv_bfloat16>(x, dy, part, part_b, pl, st)
                                 : cv_wg_dispatch<__half>(x, dy, part, part_b, pl, st);
    if (rc != LMNET_OK) return rc;
    const int n = 9 * d->Cout * d->Cin + d->Cout;
    LMNET_LAUNCH(KID_CONV3X3_REDUCE, st, 0, (conv3x3_wgrad_reduce_kernel<<<(n + 63) / 64, 256, 0, st>>>(
        part, part_b, dW, dbias, pl.grid, d->Cout, d->Cin, pl.MT * 16, pl.NTC * 8)));
    return LMNET_OK;
}

extern "C" int lmnet_conv3x3_wgrad(const void* x, const void* dy, float* dW, float* dbias, void* workspace, size_t workspace_bytes,
                                   const lmnet_conv3x3_dims* d, int dtype, void* stream) {
    return lmnet_conv3x3_wgrad_strided(x, dy, 0, dW, dbias, workspace, workspace_bytes, d, dtype, stream);
}
