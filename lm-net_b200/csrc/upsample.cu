// upsample.cu — bilinear x2 up-sampling with align_corners=True on NCHW tensors, forward and backward.
//
// Widening step f3 of SURVEY.md §8: the nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)
// in front of the decoder's up-convolutions and inside the skip-fusion blocks
// (/root/reference/core/LM_Net.py:58-74, /root/reference/core/modules.py:93-95, 129-131).  Under autocast
// ATen runs this op in fp32 with a cast on either side (3.5 ms + casts of the round-1 step).  Here it is one
// bandwidth kernel per direction in the storage type, fp32 interpolation arithmetic with ATen's index rule:
//   src = dst * (in-1)/(out-1);  i0 = floor(src);  i1 = i0 + (i0 < in-1);  l1 = src - i0;  l0 = 1 - l1.
// The backward is a gather (no atomics, deterministic): every input pixel sums the <= 6x6 output pixels whose
// interpolation footprint contains it.
#include "common.cuh"

namespace lmnet {

struct UpGeom {
    int H, W, OH, OW;
    int64_t planes;       // B*C
    float rh, rw;         // (H-1)/(OH-1), (W-1)/(OW-1)
};

__device__ __forceinline__ void src_index(int o, float r, int in, int& i0, int& i1, float& l0, float& l1) {
    const float s = r * (float)o;
    i0 = (int)s;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = s - (float)i0;
    l0 = 1.f - l1;
}

template <typename T>
__global__ void __launch_bounds__(256)
upsample2x_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, UpGeom g) {
    // one thread: kOut adjacent output columns of one output row of one plane (one 16-byte store for 16-bit types)
    constexpr int kOut = 8;
    const int64_t groups_per_row = (g.OW + kOut - 1) / kOut;
    const int64_t total = g.planes * g.OH * groups_per_row;
    const bool vec_ok = (g.OW % kOut) == 0 && sizeof(T) == 2;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t gr = idx % groups_per_row;
        const int64_t t = idx / groups_per_row;
        const int oy = (int)(t % g.OH);
        const int64_t plane = t / g.OH;
        int y0, y1;
        float ly0, ly1;
        src_index(oy, g.rh, g.H, y0, y1, ly0, ly1);
        const T* r0 = x + (plane * g.H + y0) * g.W;
        const T* r1 = x + (plane * g.H + y1) * g.W;
        const int ox0 = (int)gr * kOut;
        // the kOut outputs read input columns [xa, xa + 5] at most (scale just under 1/2): fetch them once
        int xa, xdummy;
        float l0, l1;
        src_index(ox0, g.rw, g.W, xa, xdummy, l0, l1);
        float c0v[6], c1v[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int xi = min(xa + k, g.W - 1);
            c0v[k] = to_f(r0[xi]);
            c1v[k] = to_f(r1[xi]);
        }
        float out[kOut];
#pragma unroll
        for (int k = 0; k < kOut; ++k) {
            const int ox = ox0 + k;
            out[k] = 0.f;
            if (ox < g.OW) {
                int x0, x1;
                float lx0, lx1;
                src_index(ox, g.rw, g.W, x0, x1, lx0, lx1);
                const int i0 = x0 - xa, i1 = x1 - xa;        // 0..5
                float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    a0 = i0 == q ? c0v[q] : a0; a1 = i1 == q ? c0v[q] : a1;
                    b0 = i0 == q ? c1v[q] : b0; b1 = i1 == q ? c1v[q] : b1;
                }
                out[k] = ly0 * (lx0 * a0 + lx1 * a1) + ly1 * (lx0 * b0 + lx1 * b1);
            }
        }
        T* o = y + (plane * g.OH + oy) * g.OW + ox0;
        if (vec_ok) {
            uint4 raw;
            T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
            for (int k = 0; k < kOut; ++k) e[k] = from_f<T>(out[k]);
            *reinterpret_cast<uint4*>(o) = raw;
        } else {
#pragma unroll
            for (int k = 0; k < kOut; ++k)
                if (ox0 + k < g.OW) o[k] = from_f<T>(out[k]);
        }
    }
}

// weight with which output index o contributes to input index i along one axis
__device__ __forceinline__ float axis_weight(int o, int i, float r, int in) {
    int i0, i1;
    float l0, l1;
    src_index(o, r, in, i0, i1, l0, l1);
    float w = 0.f;
    if (i0 == i) w += l0;
    if (i1 == i) w += l1;
    return w;
}

template <typename T>
__global__ void __launch_bounds__(256)
upsample2x_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, UpGeom g) {
    const int64_t total = g.planes * g.H * g.W;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int ix = (int)(idx % g.W);
        const int64_t t = idx / g.W;
        const int iy = (int)(t % g.H);
        const int64_t plane = t / g.H;
        // candidates: output rows/cols in [2*i - 2, 2*i + 3] (scale is just under 1/2)
        float wy[6], wx[6];
        int oy0 = 2 * iy - 2, ox0 = 2 * ix - 2;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int oy = oy0 + k, ox = ox0 + k;
            wy[k] = (oy >= 0 && oy < g.OH) ? axis_weight(oy, iy, g.rh, g.H) : 0.f;
            wx[k] = (ox >= 0 && ox < g.OW) ? axis_weight(ox, ix, g.rw, g.W) : 0.f;
        }
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            if (wy[a] != 0.f) {
                const T* row = dy + (plane * g.OH + oy0 + a) * g.OW;
                float racc = 0.f;
#pragma unroll
                for (int b = 0; b < 6; ++b)
                    if (wx[b] != 0.f) racc = fmaf(wx[b], to_f(row[ox0 + b]), racc);
                acc = fmaf(wy[a], racc, acc);
            }
        }
        dx[idx] = from_f<T>(acc);
    }
}

// ---------------------------------------------------------------------------------------------------
// channels-last ([B, H, W, C]) variants: what cuDNN's 16-bit convolutions on either side compute in, so the tensor
// never changes layout.  One thread = VEC channels of one pixel (one 16-byte access for 16-bit types when C % 8 == 0).
// ---------------------------------------------------------------------------------------------------
struct UpGeomCl {
    int H, W, OH, OW, C;
    int64_t B;
    float rh, rw;
};

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
upsample2x_cl_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, UpGeomCl g) {
    const int cv = g.C / VEC;
    const int64_t total = g.B * g.OH * g.OW * cv;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % cv) * VEC;
        int64_t t = idx / cv;
        const int ox = (int)(t % g.OW);
        t /= g.OW;
        const int oy = (int)(t % g.OH);
        const int64_t b = t / g.OH;
        int y0, y1, x0, x1;
        float ly0, ly1, lx0, lx1;
        src_index(oy, g.rh, g.H, y0, y1, ly0, ly1);
        src_index(ox, g.rw, g.W, x0, x1, lx0, lx1);
        const T* base = x + b * g.H * g.W * g.C + c;
        float a00[VEC], a01[VEC], a10[VEC], a11[VEC], out[VEC];
        load_f<VEC, T, true>(base + ((int64_t)y0 * g.W + x0) * g.C, a00);
        load_f<VEC, T, true>(base + ((int64_t)y0 * g.W + x1) * g.C, a01);
        load_f<VEC, T, true>(base + ((int64_t)y1 * g.W + x0) * g.C, a10);
        load_f<VEC, T, true>(base + ((int64_t)y1 * g.W + x1) * g.C, a11);
#pragma unroll
        for (int j = 0; j < VEC; ++j) out[j] = ly0 * (lx0 * a00[j] + lx1 * a01[j]) + ly1 * (lx0 * a10[j] + lx1 * a11[j]);
        store_f<VEC, T, true>(y + ((b * g.OH + oy) * g.OW + ox) * g.C + c, out);
    }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
upsample2x_cl_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, UpGeomCl g) {
    const int cv = g.C / VEC;
    const int64_t total = g.B * g.H * g.W * cv;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % cv) * VEC;
        int64_t t = idx / cv;
        const int ix = (int)(t % g.W);
        t /= g.W;
        const int iy = (int)(t % g.H);
        const int64_t b = t / g.H;
        float wy[6], wx[6];
        const int oy0 = 2 * iy - 2, ox0 = 2 * ix - 2;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int oy = oy0 + k, ox = ox0 + k;
            wy[k] = (oy >= 0 && oy < g.OH) ? axis_weight(oy, iy, g.rh, g.H) : 0.f;
            wx[k] = (ox >= 0 && ox < g.OW) ? axis_weight(ox, ix, g.rw, g.W) : 0.f;
        }
        float acc[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
        const T* base = dy + b * g.OH * g.OW * g.C + c;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            if (wy[a] != 0.f) {
                float racc[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j) racc[j] = 0.f;
#pragma unroll
                for (int bb = 0; bb < 6; ++bb)
                    if (wx[bb] != 0.f) {
                        float v[VEC];
                        load_f<VEC, T, true>(base + ((int64_t)(oy0 + a) * g.OW + ox0 + bb) * g.C, v);
#pragma unroll
                        for (int j = 0; j < VEC; ++j) racc[j] = fmaf(wx[bb], v[j], racc[j]);
                    }
#pragma unroll
                for (int j = 0; j < VEC; ++j) acc[j] = fmaf(wy[a], racc[j], acc[j]);
            }
        }
        store_f<VEC, T, true>(dx + ((b * g.H + iy) * g.W + ix) * g.C + c, acc);
    }
}

static int up_validate(const lmnet_upsample_dims* d) {
    if (d == nullptr || d->planes <= 0 || d->H <= 0 || d->W <= 0) return LMNET_ERR_INVALID_ARG;
    return LMNET_OK;
}
static UpGeom up_geom(const lmnet_upsample_dims* d) {
    UpGeom g;
    g.H = d->H; g.W = d->W; g.OH = 2 * d->H; g.OW = 2 * d->W; g.planes = d->planes;
    g.rh = g.OH > 1 ? (float)(g.H - 1) / (float)(g.OH - 1) : 0.f;
    g.rw = g.OW > 1 ? (float)(g.W - 1) / (float)(g.OW - 1) : 0.f;
    return g;
}
static unsigned up_blocks(int64_t total) {
    int64_t b = (total + 255) / 256;
    const int64_t cap = 148 * 16;
    return (unsigned)(b < cap ? (b < 1 ? 1 : b) : cap);
}

template <typename T>
static int up_fwd(const void* x, void* y, const lmnet_upsample_dims* d, cudaStream_t st) {
    UpGeom g = up_geom(d);
    const int64_t total = g.planes * g.OH * ((g.OW + 7) / 8);
    const double bytes = (double)g.planes * ((double)g.H * g.W + (double)g.OH * g.OW) * sizeof(T);
    LMNET_LAUNCH(KID_UPSAMPLE_FWD, st, bytes, (upsample2x_fwd_kernel<T><<<up_blocks(total), 256, 0, st>>>((const T*)x, (T*)y, g)));
    return LMNET_OK;
}
template <typename T>
static int up_bwd(const void* dy, void* dx, const lmnet_upsample_dims* d, cudaStream_t st) {
    UpGeom g = up_geom(d);
    const int64_t total = g.planes * g.H * g.W;
    const double bytes = (double)g.planes * ((double)g.H * g.W + (double)g.OH * g.OW) * sizeof(T);
    LMNET_LAUNCH(KID_UPSAMPLE_BWD, st, bytes, (upsample2x_bwd_kernel<T><<<up_blocks(total), 256, 0, st>>>((const T*)dy, (T*)dx, g)));
    return LMNET_OK;
}

}  // namespace lmnet

using namespace lmnet;

extern "C" int lmnet_upsample2x_fwd(const void* x, void* y, const lmnet_upsample_dims* dims, int dtype, void* stream) {
    int rc = up_validate(dims);
    if (rc != LMNET_OK) return rc;
    if (!x || !y) return LMNET_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case LMNET_F32: return up_fwd<float>(x, y, dims, st);
        case LMNET_BF16: return up_fwd<__nv_bfloat16>(x, y, dims, st);
        case LMNET_F16: return up_fwd<__half>(x, y, dims, st);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}

extern "C" int lmnet_upsample2x_bwd(const void* dy, void* dx, const lmnet_upsample_dims* dims, int dtype, void* stream) {
    int rc = up_validate(dims);
    if (rc != LMNET_OK) return rc;
    if (!dy || !dx) return LMNET_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case LMNET_F32: return up_bwd<float>(dy, dx, dims, st);
        case LMNET_BF16: return up_bwd<__nv_bfloat16>(dy, dx, dims, st);
        case LMNET_F16: return up_bwd<__half>(dy, dx, dims, st);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}

// channels-last entry points: x [B, H, W, C] -> y [B, 2H, 2W, C]
namespace lmnet {
static int up_cl_vec(const lmnet_upsample_cl_dims* d, size_t es, const void* a, const void* b) {
    for (int v = (int)(16 / es); v > 1; v >>= 1)
        if (d->C % v == 0 && (uintptr_t)a % (v * es) == 0 && (uintptr_t)b % (v * es) == 0) return v;
    return 1;
}
template <typename T, int VEC>
static int up_cl_launch(bool fwd, const void* src, void* dst, const lmnet_upsample_cl_dims* d, cudaStream_t st) {
    UpGeomCl g;
    g.B = d->B; g.H = d->H; g.W = d->W; g.C = d->C; g.OH = 2 * d->H; g.OW = 2 * d->W;
    g.rh = g.OH > 1 ? (float)(g.H - 1) / (float)(g.OH - 1) : 0.f;
    g.rw = g.OW > 1 ? (float)(g.W - 1) / (float)(g.OW - 1) : 0.f;
    const double bytes = (double)g.B * g.C * ((double)g.H * g.W + (double)g.OH * g.OW) * sizeof(T);
    if (fwd) {
        const int64_t total = g.B * g.OH * g.OW * (g.C / VEC);
        LMNET_LAUNCH(KID_UPSAMPLE_FWD, st, bytes, (upsample2x_cl_fwd_kernel<T, VEC><<<up_blocks(total), 256, 0, st>>>((const T*)src, (T*)dst, g)));
    } else {
        const int64_t total = g.B * g.H * g.W * (g.C / VEC);
        LMNET_LAUNCH(KID_UPSAMPLE_BWD, st, bytes, (upsample2x_cl_bwd_kernel<T, VEC><<<up_blocks(total), 256, 0, st>>>((const T*)src, (T*)dst, g)));
    }
    return LMNET_OK;
}
template <typename T>
static int up_cl(bool fwd, const void* src, void* dst, const lmnet_upsample_cl_dims* d, cudaStream_t st) {
    switch (up_cl_vec(d, sizeof(T), src, dst)) {
        case 8: if constexpr (sizeof(T) == 2) return up_cl_launch<T, 8>(fwd, src, dst, d, st); else return LMNET_ERR_UNSUPPORTED;
        case 4: return up_cl_launch<T, 4>(fwd, src, dst, d, st);
        case 2: return up_cl_launch<T, 2>(fwd, src, dst, d, st);
        default: return up_cl_launch<T, 1>(fwd, src, dst, d, st);
    }
}
static int up_cl_entry(bool fwd, const void* src, void* dst, const lmnet_upsample_cl_dims* d, int dtype, void* stream) {
    if (d == nullptr || d->B <= 0 || d->H <= 0 || d->W <= 0 || d->C <= 0) return LMNET_ERR_INVALID_ARG;
    if (!src || !dst) return LMNET_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case LMNET_F32: return up_cl<float>(fwd, src, dst, d, st);
        case LMNET_BF16: return up_cl<__nv_bfloat16>(fwd, src, dst, d, st);
        case LMNET_F16: return up_cl<__half>(fwd, src, dst, d, st);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}
}  // namespace lmnet

extern "C" int lmnet_upsample2x_cl_fwd(const void* x, void* y, const lmnet_upsample_cl_dims* dims, int dtype, void* stream) {
    return lmnet::up_cl_entry(true, x, y, dims, dtype, stream);
}
extern "C" int lmnet_upsample2x_cl_bwd(const void* dy, void* dx, const lmnet_upsample_cl_dims* dims, int dtype, void* stream) {
    return lmnet::up_cl_entry(false, dy, dx, dims, dtype, stream);
}
