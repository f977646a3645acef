// conv3x3.cuh — helpers shared by the generic (conv3x3.cu) and the channel-specialised (conv3x3_fast.cu) 3x3 kernels.
#pragma once
#include "common.cuh"

namespace lmnet {

constexpr int kCvThreads = 256;
constexpr int kCvWarps = 8;
constexpr int kCvWgThreads = 288;      // weight gradient: one warp per tap
constexpr int kCvTW = 16;              // output pixels per tile row = one MMA M tile

struct CvGeom {
    int B, H, W, Ho, Wo, Cin, Cout, S;
    int pitch_x, pitch_w, pitch_o;     // element pitches: staged input pixel, weight row, staged output pixel
    int gx_pitch, gd_pitch;            // GLOBAL pixel pitch (elements) of the input tensor and of dy in the weight gradient:
                                       // Cin / Cout for dense tensors, larger for a channel slice of a wider channels-last
                                       // tensor (the gradient slices autograd hands a convolution whose output was concatenated)
    int TH, IH, IW;                    // output rows per tile; staged input rows / pixels per row
    int tiles_x, tiles_y, tiles;
    int ksteps;                        // ceil(Cin / 16)
    int ncta;
};

__device__ __forceinline__ void cv_cp16(void* smem, const void* gmem, bool valid) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cv_cp8(void* smem, const void* gmem, bool valid) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cv_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cv_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ void cv_ldsm_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void cv_ldsm_x4_t(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void cv_ldsm_x2(uint32_t (&r)[2], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];"
                 : "=r"(r[0]), "=r"(r[1]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void cv_ldsm_x2_t(uint32_t (&r)[2], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
                 : "=r"(r[0]), "=r"(r[1]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
template <typename T> __device__ __forceinline__ void cv_mma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]);
template <> __device__ __forceinline__ void cv_mma<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <> __device__ __forceinline__ void cv_mma<__half>(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <typename T> __device__ __forceinline__ uint32_t cv_pack(float lo, float hi);
template <> __device__ __forceinline__ uint32_t cv_pack<__nv_bfloat16>(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
template <> __device__ __forceinline__ uint32_t cv_pack<__half>(float lo, float hi) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
template <typename T> __device__ __forceinline__ uint32_t cv_ones();
template <> __device__ __forceinline__ uint32_t cv_ones<__nv_bfloat16>() { return 0x3f803f80u; }
template <> __device__ __forceinline__ uint32_t cv_ones<__half>() { return 0x3c003c00u; }

struct CvTile {
    int b, oy0, ox0;
};
__device__ __forceinline__ CvTile cv_tile(const CvGeom& g, int t) {
    CvTile r;
    const int per_img = g.tiles_x * g.tiles_y;
    r.b = t / per_img;
    const int k = t - r.b * per_img;
    const int ty = k / g.tiles_x;
    r.oy0 = ty * g.TH;
    r.ox0 = (k - ty * g.tiles_x) * kCvTW;
    return r;
}

// stage the input tile of output tile `tl`: IH x IW pixels starting at (oy0*S - 1, ox0*S - 1), zero outside the image
template <typename T, int NTHREADS>
__device__ __forceinline__ void cv_issue_x(T* s, const T* __restrict__ x, const CvGeom& g, const CvTile& tl) {
    const int iy0 = tl.oy0 * g.S - 1, ix0 = tl.ox0 * g.S - 1;
    const T* xb = x + (int64_t)tl.b * g.H * g.W * g.gx_pitch;
    const int npix = g.IH * g.IW;
    if ((g.Cin & 7) == 0) {
        const int vpp = g.Cin >> 3;
        for (int i = threadIdx.x; i < npix * vpp; i += NTHREADS) {
            const int pix = i / vpp, v = i - pix * vpp;
            const int r = pix / g.IW, c = pix - r * g.IW;
            const int iy = iy0 + r, ix = ix0 + c;
            const bool ok = iy >= 0 && iy < g.H && ix >= 0 && ix < g.W;
            cv_cp16(s + pix * g.pitch_x + v * 8, ok ? xb + ((int64_t)iy * g.W + ix) * g.gx_pitch + v * 8 : x, ok);
        }
    } else {
        const int vpp = g.Cin >> 2;
        for (int i = threadIdx.x; i < npix * vpp; i += NTHREADS) {
            const int pix = i / vpp, v = i - pix * vpp;
            const int r = pix / g.IW, c = pix - r * g.IW;
            const int iy = iy0 + r, ix = ix0 + c;
            const bool ok = iy >= 0 && iy < g.H && ix >= 0 && ix < g.W;
            cv_cp8(s + pix * g.pitch_x + v * 4, ok ? xb + ((int64_t)iy * g.W + ix) * g.gx_pitch + v * 4 : x, ok);
        }
    }
}


__host__ __device__ constexpr int cv_pitch_c(int cols) {     // multiple of 8 elements, == 8 (mod 16): conflict-free ldmatrix rows
    int p = (cols + 7) / 8 * 8;
    if (p % 16 != 8) p += 8;
    return p;
}

// W[tap][n][k] of the contraction out of the layer's fp32 parameter (torch layout [Co][Ci][3][3]):
//   forward (w_t = 0):            n = co, k = ci:  w[co][ci][tap]
//   stride-1 input gradient (1):  n = ci, k = co:  w[co][ci][8 - tap]   (flipped taps, transposed channels)
__device__ __forceinline__ float cv_weight(const float* __restrict__ w, int w_t, int tap, int n, int k, int Cin, int Cout) {
    return w_t ? __ldg(w + ((int64_t)k * Cout + n) * 9 + (8 - tap)) : __ldg(w + ((int64_t)n * Cin + k) * 9 + tap);
}

// stage the whole fp32 parameter [Co][Ci][3][3] into shared memory as W[tap][n][k] (pitch PW, n_pad rows per tap), read
// in memory order (coalesced) and rounded to T here; the padding must have been zeroed before
template <typename T, int NTHREADS>
__device__ __forceinline__ void cv_stage_weights(T* s_w, const float* __restrict__ w, int w_t, int Cin, int Cout, int n_pad, int PW) {
    // forward: the tensor is [Cout][Cin][9]; transposed: the layer's own tensor is [Cin_op = its Cout ... ] = [K][N][9]
    const int outer = w_t ? Cin : Cout, inner = w_t ? Cout : Cin;           // memory: [outer][inner][9]
    const int total = outer * inner * 9;
    for (int i = threadIdx.x; i < total; i += NTHREADS) {
        const int o = i / (inner * 9), rem = i - o * inner * 9;
        const int in = rem / 9, tap = rem - in * 9;
        const float v = __ldg(w + i);
        if (w_t) s_w[((8 - tap) * n_pad + in) * PW + o] = from_f<T>(v);      // n = ci (inner), k = co (outer), flipped tap
        else s_w[(tap * n_pad + o) * PW + in] = from_f<T>(v);                // n = co (outer), k = ci (inner)
    }
}

// channel-specialised kernels (conv3x3_fast.cu): LMNET_ERR_UNSUPPORTED when (stride, Cin, Cout) is not in their list
bool cv_fast_fwd_has(int S, int Cin, int Cout);
bool cv_fast_wgrad_has(int S, int Cin, int Cout);
int cv_fast_fwd(const void* x, const float* w, int w_t, const float* bias, void* y, const lmnet_conv3x3_dims* d, int dtype, cudaStream_t st,
                int x_pitch = 0);
int cv_fast_wgrad_grid(const lmnet_conv3x3_dims* d, int* mp, int* ldn);     // CTAs (= partials), padded Cout, padded Cin
int cv_fast_wgrad(const void* x, const void* dy, float* part, float* part_b, const lmnet_conv3x3_dims* d, int dtype, cudaStream_t st,
                  int dy_pitch = 0);

}  // namespace lmnet
