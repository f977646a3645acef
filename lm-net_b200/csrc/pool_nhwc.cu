// pool_nhwc.cu — average pooling by an integer factor on channels-last tensors (sm_100a).
//
// Reference: PyramidPool (/root/reference/core/modules.py:454-498) pools the four encoder outputs to the bottleneck
// resolution with nn.AdaptiveAvgPool2d; at 352 x 352 the factors are exactly 16 / 8 / 4 / 2.  ATen's channels-last
// adaptive kernels take 162 us + 79 us for the 48 MB level-1 tensor (0.3 TB/s); this is a plain streaming reduction:
// one thread per (output pixel, 4-channel vector) walks its f x f window (the f pixels of a window row are contiguous in
// memory), fp32 accumulation; the backward broadcasts dy / f^2 with fully coalesced stores.
#include "common.cuh"

namespace lmnet {

template <typename T>
__global__ void __launch_bounds__(256)
avgpool_cl_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t nout, int Ho, int Wo, int C4, int f, float inv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nout) return;
    const int v = (int)(i % C4);
    const int64_t pix = i / C4;
    const int ox = (int)(pix % Wo);
    const int64_t t = pix / Wo;
    const int oy = (int)(t % Ho);
    const int64_t b = t / Ho;
    const int W = Wo * f, C = C4 * 4;
    const T* p = x + (((b * Ho + oy) * (int64_t)f) * W + (int64_t)ox * f) * C + v * 4;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int wy = 0; wy < f; ++wy) {
        const T* q = p + (int64_t)wy * W * C;
#pragma unroll 4
        for (int wx = 0; wx < f; ++wx) {
            const uint2 raw = __ldg(reinterpret_cast<const uint2*>(q + (int64_t)wx * C));
            const T* e = reinterpret_cast<const T*>(&raw);
            a0 += to_f(e[0]); a1 += to_f(e[1]); a2 += to_f(e[2]); a3 += to_f(e[3]);
        }
    }
    T o[4] = {from_f<T>(a0 * inv), from_f<T>(a1 * inv), from_f<T>(a2 * inv), from_f<T>(a3 * inv)};
    *reinterpret_cast<uint2*>(y + pix * C + v * 4) = *reinterpret_cast<const uint2*>(o);
}

template <typename T>
__global__ void __launch_bounds__(256)
avgpool_cl_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int64_t nin, int Ho, int Wo, int C4, int f, float inv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nin) return;
    const int v = (int)(i % C4);
    const int64_t pix = i / C4;
    const int W = Wo * f, H = Ho * f, C = C4 * 4;
    const int xx = (int)(pix % W);
    const int64_t t = pix / W;
    const int yy = (int)(t % H);
    const int64_t b = t / H;
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(dy + (((b * Ho + yy / f) * Wo) + xx / f) * C + v * 4));
    const T* e = reinterpret_cast<const T*>(&raw);
    T o[4] = {from_f<T>(to_f(e[0]) * inv), from_f<T>(to_f(e[1]) * inv), from_f<T>(to_f(e[2]) * inv), from_f<T>(to_f(e[3]) * inv)};
    *reinterpret_cast<uint2*>(dx + pix * C + v * 4) = *reinterpret_cast<const uint2*>(o);
}

template <typename T>
static int pool_launch(bool fwd, const void* in, void* out, const lmnet_pool_dims* d, cudaStream_t st) {
    const int C4 = d->C / 4;
    const float inv = 1.f / (float)(d->factor * d->factor);
    const int64_t npix_out = (int64_t)d->B * d->Ho * d->Wo, npix_in = npix_out * d->factor * d->factor;
    const double bytes = (double)(npix_out + npix_in) * d->C * sizeof(T);
    if (fwd) {
        const int64_t n = npix_out * C4;
        LMNET_LAUNCH(KID_AVGPOOL_FWD, st, bytes, (avgpool_cl_fwd_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
            (const T*)in, (T*)out, n, d->Ho, d->Wo, C4, d->factor, inv)));
    } else {
        const int64_t n = npix_in * C4;
        LMNET_LAUNCH(KID_AVGPOOL_BWD, st, bytes, (avgpool_cl_bwd_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
            (const T*)in, (T*)out, n, d->Ho, d->Wo, C4, d->factor, inv)));
    }
    return LMNET_OK;
}

}  // namespace lmnet

using namespace lmnet;

static int pool_check(const void* a, const void* b, const lmnet_pool_dims* d, int dtype) {
    if (d == nullptr || d->B <= 0 || d->Ho <= 0 || d->Wo <= 0 || d->C <= 0 || d->factor < 1) return LMNET_ERR_INVALID_ARG;
    if (d->C % 4 != 0 || (dtype != LMNET_BF16 && dtype != LMNET_F16)) return LMNET_ERR_UNSUPPORTED;
    if (!a || !b) return LMNET_ERR_INVALID_ARG;
    if ((uintptr_t)a % 8 != 0 || (uintptr_t)b % 8 != 0) return LMNET_ERR_UNSUPPORTED;
    if ((int64_t)d->B * d->Ho * d->Wo * d->factor * d->factor * (d->C / 4) >= ((int64_t)1 << 39)) return LMNET_ERR_UNSUPPORTED;
    return LMNET_OK;
}

extern "C" int lmnet_avgpool_cl_fwd(const void* x, void* y, const lmnet_pool_dims* d, int dtype, void* stream) {
    const int rc = pool_check(x, y, d, dtype);
    if (rc != LMNET_OK) return rc;
    return dtype == LMNET_BF16 ? pool_launch<__nv_bfloat16>(true, x, y, d, (cudaStream_t)stream)
                               : pool_launch<__half>(true, x, y, d, (cudaStream_t)stream);
}

extern "C" int lmnet_avgpool_cl_bwd(const void* dy, void* dx, const lmnet_pool_dims* d, int dtype, void* stream) {
    const int rc = pool_check(dy, dx, d, dtype);
    if (rc != LMNET_OK) return rc;
    return dtype == LMNET_BF16 ? pool_launch<__nv_bfloat16>(false, dy, dx, d, (cudaStream_t)stream)
                               : pool_launch<__half>(false, dy, dx, d, (cudaStream_t)stream);
}
