// wgrad_1x1.cu — weight / bias gradients of 1x1 convolutions on NCHW tensors (sm_100a, 16-bit storage).
//
// Widening step f1 of SURVEY.md §8 (rest of ReparamConv: expand 1x1, pointwise 1x1, shortcut 1x1,
// /root/reference/core/modules.py:537, 576-584, 587, 598-599).  Forward and input gradients of a 1x1
// convolution on NCHW data are plain GEMMs whose big operand is already laid out right (cuBLAS via
// torch.bmm).  The WEIGHT gradient is the awkward one:
//     dW[b][m][n] = sum_p A[b][m][p] * Bt[b][n][p]          (A = grad_output, Bt = layer input)
// — a tiny M x N result reduced over K = H*W up to 124k pixels.  cuBLAS has no good kernel for it (7 ms
// of a 70 ms step in profiles/r01_step_profile_c_*).  Here both operands are K-contiguous, i.e. exactly
// the row/col fragment layouts of mma.sync m16n8k16, so the kernel streams pixel chunks through a
// 2-stage cp.async pipeline, every warp keeps a block of M-tiles x N-tiles of fp32 accumulators in
// registers, and the pixel range is split over CTAs whose partials are summed in a fixed order.
// An extra all-ones row appended to Bt makes the bias gradient (row sums of A) fall out of the same MMAs.
// Bt may come from two tensors (rows [0,N1) from B1, [N1,N1+N2) from B2) so that pointwise+shortcut,
// which share grad_output, read it once.
#include "common.cuh"

namespace lmnet {

constexpr int kWgThreads = 256;
constexpr int kWgWarps = 8;
constexpr int kWgKP = 64;              // pixels per pipeline stage
constexpr int kWgPitch = kWgKP + 8;    // 144-byte rows: conflict-free ldmatrix

struct WgGeom {
    int B, M, N1, N2, NT;              // NT = number of 8-wide N tiles incl. the ones row
    int64_t P;
    int splits, chunks_per_split;      // pixel chunks of kWgKP handled by one CTA
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ void wg_ldmatrix_x4(uint32_t (&r)[4], const void* p) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void wg_ldmatrix_x2(uint32_t (&r)[2], const void* p) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
template <typename T> __device__ __forceinline__ void wg_mma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]);
template <> __device__ __forceinline__ void wg_mma<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <> __device__ __forceinline__ void wg_mma<__half>(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <typename T> __device__ __forceinline__ T wg_one();
template <> __device__ __forceinline__ __nv_bfloat16 wg_one<__nv_bfloat16>() { return __float2bfloat16_rn(1.f); }
template <> __device__ __forceinline__ __half wg_one<__half>() { return __float2half_rn(1.f); }

// part[b][split][M][NT*8] (fp32).  MT = M tiles per CTA (all of them), NTW = N tiles per warp.
template <typename T, int MT, int NTW>
__global__ void __launch_bounds__(kWgThreads)
wgrad_1x1_kernel(const T* __restrict__ A, const T* __restrict__ B1, const T* __restrict__ B2,
                 float* __restrict__ part, WgGeom g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int a_rows = MT * 16, b_rows = g.NT * 8;
    const int stage_elems = (a_rows + b_rows) * kWgPitch;
    T* smem = reinterpret_cast<T*>(smem_raw);
    const int b = blockIdx.y, split = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Ntot = g.N1 + g.N2;                      // real input channels; row Ntot is the ones row
    const int64_t chunk0 = (int64_t)split * g.chunks_per_split;
    const int64_t nchunks_total = (g.P + kWgKP - 1) / kWgKP;
    const int nchunks = (int)max((int64_t)0, min((int64_t)g.chunks_per_split, nchunks_total - chunk0));

    // rows that no copy ever touches: zero padding rows of A / Bt, and the ones row (both stages)
    for (int st = 0; st < 2; ++st) {
        T* sa = smem + st * stage_elems;
        T* sb = sa + a_rows * kWgPitch;
        for (int i = threadIdx.x; i < (a_rows - g.M) * kWgPitch; i += kWgThreads) sa[g.M * kWgPitch + i] = from_f<T>(0.f);
        for (int i = threadIdx.x; i < (b_rows - Ntot) * kWgPitch; i += kWgThreads)
            sb[Ntot * kWgPitch + i] = (i < kWgPitch) ? wg_one<T>() : from_f<T>(0.f);
    }

    auto issue = [&](int c, int st) {
        T* sa = smem + st * stage_elems;
        T* sb = sa + a_rows * kWgPitch;
        const int64_t p0 = (chunk0 + c) * kWgKP;
        constexpr int VPR = kWgKP / 8;                  // 16-byte vectors per row
        const int rows = g.M + Ntot;
        for (int i = threadIdx.x; i < rows * VPR; i += kWgThreads) {
            const int r = i / VPR, v = i - r * VPR;
            const int64_t p = p0 + v * 8;
            const bool ok = p < g.P;                    // P % 8 == 0: a vector is all in or all out
            const T* src;
            T* dst;
            if (r < g.M) {
                src = A + ((int64_t)b * g.M + r) * g.P + p;
                dst = sa + r * kWgPitch + v * 8;
            } else {
                const int n = r - g.M;
                src = n < g.N1 ? B1 + ((int64_t)b * g.N1 + n) * g.P + p : B2 + ((int64_t)b * g.N2 + (n - g.N1)) * g.P + p;
                dst = sb + n * kWgPitch + v * 8;
            }
            cp_async16(dst, ok ? src : A, ok);
        }
        cp_async_commit();
    };

    float acc[MT][NTW][4];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NTW; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
    const int nt0 = warp * NTW;                        // this warp's first N tile

    if (nchunks > 0) issue(0, 0);
    for (int c = 0; c < nchunks; ++c) {
        const int st = c & 1;
        if (c + 1 < nchunks) {
            issue(c + 1, st ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const T* sa = smem + st * stage_elems;
        const T* sb = sa + a_rows * kWgPitch;
        // the ones row must only count real pixels of the tail chunk
        if ((chunk0 + c + 1) * kWgKP > g.P) {
            const int valid = (int)(g.P - (chunk0 + c) * kWgKP);
            T* ones = const_cast<T*>(sb) + Ntot * kWgPitch;
            for (int i = threadIdx.x; i < kWgKP; i += kWgThreads) ones[i] = i < valid ? wg_one<T>() : from_f<T>(0.f);
            __syncthreads();
        }
#pragma unroll
        for (int kt = 0; kt < kWgKP / 16; ++kt) {
            uint32_t bf[NTW][2];
#pragma unroll
            for (int j = 0; j < NTW; ++j) {
                const int nt = nt0 + j;
                if (nt < g.NT) {
                    // matrix 0: rows n 0..7, k 0..7 ; matrix 1: k 8..15   (lanes 0-15 give the addresses)
                    const int l = lane & 15;
                    wg_ldmatrix_x2(bf[j], sb + (nt * 8 + (l & 7)) * kWgPitch + kt * 16 + (l >> 3) * 8);
                }
            }
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                uint32_t af[4];
                const int m = lane >> 3, rr = lane & 7;
                wg_ldmatrix_x4(af, sa + (i * 16 + (m & 1) * 8 + rr) * kWgPitch + kt * 16 + (m >> 1) * 8);
#pragma unroll
                for (int j = 0; j < NTW; ++j)
                    if (nt0 + j < g.NT) wg_mma<T>(acc[i][j], af, bf[j]);
            }
        }
        __syncthreads();
        if ((chunk0 + c + 1) * kWgKP > g.P) {           // restore the ones row for the next use of this stage
            T* ones = const_cast<T*>(sb) + Ntot * kWgPitch;
            for (int i = threadIdx.x; i < kWgKP; i += kWgThreads) ones[i] = wg_one<T>();
        }
    }
    // write this CTA's partial: part[((b*splits + split)*M + m) * (NT*8) + n]
    const int ldn = g.NT * 8;
    float* out = part + ((int64_t)b * g.splits + split) * g.M * ldn;
    const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NTW; ++j) {
            const int nt = nt0 + j;
            if (nt < g.NT) {
                const int n = nt * 8 + 2 * tq;
                const int m0 = i * 16 + gq, m1 = m0 + 8;
                if (m0 < g.M) *reinterpret_cast<float2*>(out + (int64_t)m0 * ldn + n) = make_float2(acc[i][j][0], acc[i][j][1]);
                if (m1 < g.M) *reinterpret_cast<float2*>(out + (int64_t)m1 * ldn + n) = make_float2(acc[i][j][2], acc[i][j][3]);
            }
        }
}

// dW[b][m][n] (n < N1+N2) and drow[b][m] (the ones column) = sum over splits, fixed order
// (partial column of output n: n < N1 -> n; N1 <= n < Ntot -> n2_off + n - N1; the ones column -> ones_col)
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dW, float* __restrict__ drow,
                                    int B, int splits, int M, int Ntot, int ldn, int N1, int n2_off, int ones_col) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)B * M * (Ntot + 1);
    if (idx >= total) return;
    const int n = (int)(idx % (Ntot + 1));
    const int m = (int)((idx / (Ntot + 1)) % M);
    const int b = (int)(idx / ((int64_t)(Ntot + 1) * M));
    const int col = n < N1 ? n : n < Ntot ? n2_off + (n - N1) : ones_col;
    float a = 0.f;
    for (int s = 0; s < splits; ++s) a += part[(((int64_t)b * splits + s) * M + m) * ldn + col];
    if (n < Ntot) dW[((int64_t)b * M + m) * Ntot + n] = a;
    else if (drow != nullptr) drow[(int64_t)b * M + m] = a;
}

// the same reduction, additionally summed over the batch: dW[m][n], drow[m]
__global__ void __launch_bounds__(256)
wgrad_reduce_bsum_kernel(const float* __restrict__ part, float* __restrict__ dW, float* __restrict__ drow,
                         int B, int splits, int M, int Ntot, int ldn, int N1, int n2_off, int ones_col) {
    // one block per output row m; threads = (column of the partial, partial lane): adjacent threads read adjacent columns
    // (coalesced), every lane sums its share of the B * splits partials, lanes are combined in a fixed order
    extern __shared__ float s_red[];                      // [G][ldn]
    const int m = blockIdx.x;
    const int G = blockDim.x / ldn;
    const int col = threadIdx.x % ldn, g = threadIdx.x / ldn;
    const int total = B * splits;
    if (g < G) {
        const float* p = part + (int64_t)m * ldn + col;
        const int64_t stride = (int64_t)M * ldn;
        float a0 = 0.f, a1 = 0.f;
        int s = g;
        for (; s + G < total; s += 2 * G) { a0 += p[s * stride]; a1 += p[(s + G) * stride]; }
        if (s < total) a0 += p[s * stride];
        s_red[g * ldn + col] = a0 + a1;
    }
    __syncthreads();
    for (int n = threadIdx.x; n <= Ntot; n += blockDim.x) {
        const int c = n < N1 ? n : n < Ntot ? n2_off + (n - N1) : ones_col;
        float a = 0.f;
        for (int k = 0; k < G; ++k) a += s_red[k * ldn + c];
        if (n < Ntot) dW[(int64_t)m * Ntot + n] = a;
        else if (drow != nullptr) drow[m] = a;
    }
}

// ---------------------------------------------------------------------------------------------------
// Mixed-layout variant: operands may be channels-last ([B, P, C], C contiguous — the layout of the block's
// input / output tensors) instead of NCHW planes ([B, C, P]).  A channels-last tile is staged as [64 pixels][C]
// and read with ldmatrix.trans, so no transposed copy of the activation is ever made:
//     expand:              dW[E, Cin]      = sum_p dy[e, p]   * x[p, cin]          A planes, B1 channels-last
//     pointwise+shortcut:  dW[Cout, E+Cin] = sum_p dout[p, co] * (z[e, p] | x[p, cin])   A cl, B1 planes, B2 cl
// B2 (when present) is always channels-last.  The all-ones column that yields the bias gradient lives with the last
// operand.  Partials: part[b][split][M][(NT1 + NT2) * 8], B2's columns start at NT1 * 8.
// ---------------------------------------------------------------------------------------------------
struct WgGeomCl {
    int B, M, N1, N2, NT1, NT2;
    int pitchA, pitchB1, pitchB2;      // element pitch of the channels-last tiles (unused for plane operands)
    int64_t P;
    int splits, chunks_per_split;
};

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, bool valid) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void wg_ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void wg_ldmatrix_x2_trans(uint32_t (&r)[2], const void* p) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}

// stage a channels-last operand tile: 64 pixel rows x C channels (16-byte vectors when C % 8 == 0, else 8-byte)
template <typename T>
__device__ __forceinline__ void wg_issue_cl(T* s, int pitch, const T* src, int C, int64_t p0, int64_t P) {
    if ((C & 7) == 0) {
        const int vpr = C >> 3;
        for (int i = threadIdx.x; i < kWgKP * vpr; i += kWgThreads) {
            const int r = i / vpr, v = i - r * vpr;
            const bool ok = p0 + r < P;
            cp_async16(s + r * pitch + v * 8, ok ? src + (p0 + r) * C + v * 8 : src, ok);
        }
    } else {
        const int vpr = C >> 2;
        for (int i = threadIdx.x; i < kWgKP * vpr; i += kWgThreads) {
            const int r = i / vpr, v = i - r * vpr;
            const bool ok = p0 + r < P;
            cp_async8(s + r * pitch + v * 4, ok ? src + (p0 + r) * C + v * 4 : src, ok);
        }
    }
}
// stage a plane operand tile: `rows` channel rows x 64 pixels
template <typename T>
__device__ __forceinline__ void wg_issue_planes(T* s, const T* src, int rows, int64_t p0, int64_t P) {
    constexpr int VPR = kWgKP / 8;
    for (int i = threadIdx.x; i < rows * VPR; i += kWgThreads) {
        const int r = i / VPR, v = i - r * VPR;
        const int64_t p = p0 + v * 8;
        const bool ok = p < P;                          // P % 8 == 0: a vector is all in or all out
        cp_async16(s + r * kWgPitch + v * 8, ok ? src + (int64_t)r * P + p : src, ok);
    }
}

template <typename T, int MT, int NTW, bool ACL, bool B1CL>
__global__ void __launch_bounds__(kWgThreads)
wgrad_1x1_cl_kernel(const T* __restrict__ A, const T* __restrict__ B1, const T* __restrict__ B2,
                    float* __restrict__ part, WgGeomCl g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int a_elems = ACL ? kWgKP * g.pitchA : MT * 16 * kWgPitch;
    const int b1_elems = B1CL ? kWgKP * g.pitchB1 : g.NT1 * 8 * kWgPitch;
    const int b2_elems = g.N2 > 0 ? kWgKP * g.pitchB2 : 0;
    const int stage_elems = a_elems + b1_elems + b2_elems;
    T* smem = reinterpret_cast<T*>(smem_raw);
    const int b = blockIdx.y, split = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t chunk0 = (int64_t)split * g.chunks_per_split;
    const int64_t nchunks_total = (g.P + kWgKP - 1) / kWgKP;
    const int nchunks = (int)max((int64_t)0, min((int64_t)g.chunks_per_split, nchunks_total - chunk0));
    const bool ones_in_b2 = g.N2 > 0;
    const T one = wg_one<T>(), zero = from_f<T>(0.f);

    // everything no copy ever touches: padding rows / columns are zero, the ones row / column is one (both stages)
    for (int st = 0; st < 2; ++st) {
        T* sa = smem + st * stage_elems;
        T* sb1 = sa + a_elems;
        T* sb2 = sb1 + b1_elems;
        if (ACL) {
            for (int i = threadIdx.x; i < kWgKP * g.pitchA; i += kWgThreads) sa[i] = zero;
        } else {
            for (int i = threadIdx.x; i < (MT * 16 - g.M) * kWgPitch; i += kWgThreads) sa[g.M * kWgPitch + i] = zero;
        }
        if (B1CL) {
            for (int i = threadIdx.x; i < kWgKP * g.pitchB1; i += kWgThreads)
                sb1[i] = (!ones_in_b2 && (i % g.pitchB1) == g.N1) ? one : zero;
        } else {
            for (int i = threadIdx.x; i < (g.NT1 * 8 - g.N1) * kWgPitch; i += kWgThreads)
                sb1[g.N1 * kWgPitch + i] = (!ones_in_b2 && i < kWgPitch) ? one : zero;
        }
        for (int i = threadIdx.x; i < b2_elems; i += kWgThreads) sb2[i] = ((i % g.pitchB2) == g.N2) ? one : zero;
    }
    __syncthreads();       // the cp.async writes below must not race with the fills above

    auto issue = [&](int c, int st) {
        T* sa = smem + st * stage_elems;
        T* sb1 = sa + a_elems;
        T* sb2 = sb1 + b1_elems;
        const int64_t p0 = (chunk0 + c) * kWgKP;
        if (ACL) wg_issue_cl<T>(sa, g.pitchA, A + (int64_t)b * g.P * g.M, g.M, p0, g.P);
        else wg_issue_planes<T>(sa, A + (int64_t)b * g.M * g.P, g.M, p0, g.P);
        if (B1CL) wg_issue_cl<T>(sb1, g.pitchB1, B1 + (int64_t)b * g.P * g.N1, g.N1, p0, g.P);
        else wg_issue_planes<T>(sb1, B1 + (int64_t)b * g.N1 * g.P, g.N1, p0, g.P);
        if (g.N2 > 0) wg_issue_cl<T>(sb2, g.pitchB2, B2 + (int64_t)b * g.P * g.N2, g.N2, p0, g.P);
        cp_async_commit();
    };

    float acc[MT][NTW][4];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NTW; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
    const int nt0 = warp * NTW;
    const int NT = g.NT1 + g.NT2;

    if (nchunks > 0) issue(0, 0);
    for (int c = 0; c < nchunks; ++c) {
        const int st = c & 1;
        if (c + 1 < nchunks) {
            issue(c + 1, st ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        T* sa = smem + st * stage_elems;
        T* sb1 = sa + a_elems;
        T* sb2 = sb1 + b1_elems;
        const bool tail = (chunk0 + c + 1) * kWgKP > g.P;
        if (tail) {            // the ones entries must only count real pixels of the tail chunk
            const int valid = (int)(g.P - (chunk0 + c) * kWgKP);
            for (int i = threadIdx.x; i < kWgKP; i += kWgThreads) {
                const T v = i < valid ? one : zero;
                if (ones_in_b2) sb2[i * g.pitchB2 + g.N2] = v;
                else if (B1CL) sb1[i * g.pitchB1 + g.N1] = v;
                else sb1[g.N1 * kWgPitch + i] = v;
            }
            __syncthreads();
        }
#pragma unroll
        for (int kt = 0; kt < kWgKP / 16; ++kt) {
            uint32_t bf[NTW][2];
#pragma unroll
            for (int j = 0; j < NTW; ++j) {
                const int nt = nt0 + j;
                if (nt < NT) {
                    const int l = lane & 15;
                    if (nt >= g.NT1) wg_ldmatrix_x2_trans(bf[j], sb2 + (kt * 16 + l) * g.pitchB2 + (nt - g.NT1) * 8);
                    else if (B1CL) wg_ldmatrix_x2_trans(bf[j], sb1 + (kt * 16 + l) * g.pitchB1 + nt * 8);
                    else wg_ldmatrix_x2(bf[j], sb1 + (nt * 8 + (l & 7)) * kWgPitch + kt * 16 + (l >> 3) * 8);
                }
            }
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                uint32_t af[4];
                const int m = lane >> 3, rr = lane & 7;
                if (ACL) wg_ldmatrix_x4_trans(af, sa + (kt * 16 + (m >> 1) * 8 + rr) * g.pitchA + i * 16 + (m & 1) * 8);
                else wg_ldmatrix_x4(af, sa + (i * 16 + (m & 1) * 8 + rr) * kWgPitch + kt * 16 + (m >> 1) * 8);
#pragma unroll
                for (int j = 0; j < NTW; ++j)
                    if (nt0 + j < NT) wg_mma<T>(acc[i][j], af, bf[j]);
            }
        }
        __syncthreads();
        if (tail) {            // restore the ones entries for the next use of this stage
            for (int i = threadIdx.x; i < kWgKP; i += kWgThreads) {
                if (ones_in_b2) sb2[i * g.pitchB2 + g.N2] = one;
                else if (B1CL) sb1[i * g.pitchB1 + g.N1] = one;
                else sb1[g.N1 * kWgPitch + i] = one;
            }
        }
    }
    const int ldn = NT * 8;
    float* out = part + ((int64_t)b * g.splits + split) * g.M * ldn;
    const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NTW; ++j) {
            const int nt = nt0 + j;
            if (nt < NT) {
                const int n = nt * 8 + 2 * tq;
                const int m0 = i * 16 + gq, m1 = m0 + 8;
                if (m0 < g.M) *reinterpret_cast<float2*>(out + (int64_t)m0 * ldn + n) = make_float2(acc[i][j][0], acc[i][j][1]);
                if (m1 < g.M) *reinterpret_cast<float2*>(out + (int64_t)m1 * ldn + n) = make_float2(acc[i][j][2], acc[i][j][3]);
            }
        }
}

}  // namespace lmnet
#include "wgrad_stream.cuh"
namespace lmnet {

// element pitch of a channels-last tile that must cover `cols` columns: multiple of 8, == 8 (mod 16) so that the
// eight 16-byte rows of an ldmatrix land in distinct bank groups
static int wg_cl_pitch(int cols) {
    int p = (cols + 7) / 8 * 8;
    if (p % 16 != 8) p += 8;
    return p;
}

static bool wg_shape(int M, int N, int& MT, int& NTW, int& NT) {
    MT = (M + 15) / 16;
    NT = (N + 1 + 7) / 8;
    NTW = (NT + kWgWarps - 1) / kWgWarps;
    const bool mt_ok = MT == 1 || MT == 2 || MT == 3 || MT == 6 || MT == 12;
    const bool nt_ok = NTW == 1 || NTW == 2 || NTW == 3 || NTW == 5;
    return mt_ok && nt_ok && MT * NTW <= 30;
}

static WgGeom wg_geom(const lmnet_wgrad_dims* d, int NT) {
    WgGeom g;
    g.B = d->B; g.M = d->M; g.N1 = d->N1; g.N2 = d->N2; g.NT = NT; g.P = d->P;
    const int64_t nchunks = (d->P + kWgKP - 1) / kWgKP;
    int64_t splits = (4 * 148 + d->B - 1) / d->B;
    if (splits > nchunks) splits = nchunks;
    if (splits < 1) splits = 1;
    g.chunks_per_split = (int)((nchunks + splits - 1) / splits);
    g.splits = (int)((nchunks + g.chunks_per_split - 1) / g.chunks_per_split);
    return g;
}

template <typename T, int MT, int NTW>
static int wg_launch(const void* A, const void* B1, const void* B2, float* part, const WgGeom& g, cudaStream_t st) {
    const size_t smem = 2 * (size_t)(MT * 16 + g.NT * 8) * kWgPitch * sizeof(T);
    static std::atomic<size_t> granted[kMaxDevices];
    if (!ensure_smem(wgrad_1x1_kernel<T, MT, NTW>, smem, granted)) return LMNET_ERR_LAUNCH;
    const double bytes = (double)g.B * (g.M + g.N1 + g.N2) * g.P * sizeof(T);
    dim3 grid(g.splits, g.B);
    LMNET_LAUNCH(KID_WGRAD_1X1, st, bytes, (wgrad_1x1_kernel<T, MT, NTW><<<grid, kWgThreads, smem, st>>>(
        (const T*)A, (const T*)B1, (const T*)B2, part, g)));
    return LMNET_OK;
}

template <typename T, int MT>
static int wg_dispatch_n(int NTW, const void* A, const void* B1, const void* B2, float* part, const WgGeom& g, cudaStream_t st) {
    switch (NTW) {
        case 1: return wg_launch<T, MT, 1>(A, B1, B2, part, g, st);
        case 2: return wg_launch<T, MT, 2>(A, B1, B2, part, g, st);
        case 3: if constexpr (MT <= 6) return wg_launch<T, MT, 3>(A, B1, B2, part, g, st); else return LMNET_ERR_UNSUPPORTED;
        case 5: if constexpr (MT <= 6) return wg_launch<T, MT, 5>(A, B1, B2, part, g, st); else return LMNET_ERR_UNSUPPORTED;
        default: return LMNET_ERR_UNSUPPORTED;
    }
}
template <typename T>
static int wg_dispatch(int MT, int NTW, const void* A, const void* B1, const void* B2, float* part, const WgGeom& g, cudaStream_t st) {
    switch (MT) {
        case 1: return wg_dispatch_n<T, 1>(NTW, A, B1, B2, part, g, st);
        case 2: return wg_dispatch_n<T, 2>(NTW, A, B1, B2, part, g, st);
        case 3: return wg_dispatch_n<T, 3>(NTW, A, B1, B2, part, g, st);
        case 6: return wg_dispatch_n<T, 6>(NTW, A, B1, B2, part, g, st);
        case 12: return wg_dispatch_n<T, 12>(NTW, A, B1, B2, part, g, st);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}

// ---- mixed-layout host side ----
struct WgClPlan {
    int MT, NTW;
    WgGeomCl g;
    size_t smem;
};
static bool wg_cl_plan(const lmnet_wgrad_dims* d, bool a_cl, bool b1_cl, WgClPlan& pl) {
    if (d == nullptr || d->B <= 0 || d->M <= 0 || d->N1 <= 0 || d->N2 < 0 || d->P <= 0) return false;
    if (d->P % 8 != 0) return false;                                   // plane operands: 16-byte pixel vectors
    if (a_cl && d->M % 4 != 0) return false;                           // channels-last rows: 8-byte vectors at least
    if (b1_cl && d->N1 % 4 != 0) return false;
    if (d->N2 > 0 && (d->N2 % 4 != 0 || d->N1 % 8 != 0)) return false; // B2 tiles start at an 8-column boundary
    WgGeomCl& g = pl.g;
    g.B = d->B; g.M = d->M; g.N1 = d->N1; g.N2 = d->N2; g.P = d->P;
    g.NT1 = (d->N1 + (d->N2 > 0 ? 0 : 1) + 7) / 8;
    g.NT2 = d->N2 > 0 ? (d->N2 + 1 + 7) / 8 : 0;
    pl.MT = (d->M + 15) / 16;
    if (pl.MT == 4 || pl.MT == 5) pl.MT = 6;                           // instantiated: 1, 2, 3, 6, 12 (padding rows are zero)
    if (pl.MT > 6 && pl.MT < 12) pl.MT = 12;
    const int NT = g.NT1 + g.NT2;
    pl.NTW = (NT + kWgWarps - 1) / kWgWarps;
    if (pl.NTW == 4) pl.NTW = 5;
    const bool mt_ok = pl.MT == 1 || pl.MT == 2 || pl.MT == 3 || pl.MT == 6 || pl.MT == 12;
    const bool nt_ok = pl.NTW == 1 || pl.NTW == 2 || pl.NTW == 3 || pl.NTW == 5;
    if (!mt_ok || !nt_ok || pl.MT * pl.NTW > 30) return false;
    g.pitchA = wg_cl_pitch(pl.MT * 16);
    g.pitchB1 = wg_cl_pitch(g.NT1 * 8);
    g.pitchB2 = g.N2 > 0 ? wg_cl_pitch(g.NT2 * 8) : 0;
    const int64_t nchunks = (d->P + kWgKP - 1) / kWgKP;
    int64_t splits = (4 * 148 + d->B - 1) / d->B;
    if (splits > nchunks) splits = nchunks;
    if (splits < 1) splits = 1;
    g.chunks_per_split = (int)((nchunks + splits - 1) / splits);
    g.splits = (int)((nchunks + g.chunks_per_split - 1) / g.chunks_per_split);
    const size_t a_elems = a_cl ? (size_t)kWgKP * g.pitchA : (size_t)pl.MT * 16 * kWgPitch;
    const size_t b1_elems = b1_cl ? (size_t)kWgKP * g.pitchB1 : (size_t)g.NT1 * 8 * kWgPitch;
    const size_t b2_elems = g.N2 > 0 ? (size_t)kWgKP * g.pitchB2 : 0;
    pl.smem = 2 * (a_elems + b1_elems + b2_elems) * 2;
    return pl.smem <= 220 * 1024;
}

template <typename T, int MT, int NTW, bool ACL, bool B1CL>
static int wg_cl_launch(const void* A, const void* B1, const void* B2, float* part, const WgClPlan& pl, cudaStream_t st) {
    static std::atomic<size_t> granted[kMaxDevices];
    if (!ensure_smem(wgrad_1x1_cl_kernel<T, MT, NTW, ACL, B1CL>, pl.smem, granted)) return LMNET_ERR_LAUNCH;
    const WgGeomCl& g = pl.g;
    const double bytes = (double)g.B * (g.M + g.N1 + g.N2) * g.P * sizeof(T);
    dim3 grid(g.splits, g.B);
    LMNET_LAUNCH(KID_WGRAD_1X1, st, bytes, (wgrad_1x1_cl_kernel<T, MT, NTW, ACL, B1CL><<<grid, kWgThreads, pl.smem, st>>>(
        (const T*)A, (const T*)B1, (const T*)B2, part, g)));
    return LMNET_OK;
}
template <typename T, int MT, bool ACL, bool B1CL>
static int wg_cl_dispatch_n(const void* A, const void* B1, const void* B2, float* part, const WgClPlan& pl, cudaStream_t st) {
    switch (pl.NTW) {
        case 1: return wg_cl_launch<T, MT, 1, ACL, B1CL>(A, B1, B2, part, pl, st);
        case 2: return wg_cl_launch<T, MT, 2, ACL, B1CL>(A, B1, B2, part, pl, st);
        case 3: if constexpr (MT <= 6) return wg_cl_launch<T, MT, 3, ACL, B1CL>(A, B1, B2, part, pl, st); else return LMNET_ERR_UNSUPPORTED;
        case 5: if constexpr (MT <= 6) return wg_cl_launch<T, MT, 5, ACL, B1CL>(A, B1, B2, part, pl, st); else return LMNET_ERR_UNSUPPORTED;
        default: return LMNET_ERR_UNSUPPORTED;
    }
}
template <typename T, bool ACL, bool B1CL>
static int wg_cl_dispatch(const void* A, const void* B1, const void* B2, float* part, const WgClPlan& pl, cudaStream_t st) {
    switch (pl.MT) {
        case 1: return wg_cl_dispatch_n<T, 1, ACL, B1CL>(A, B1, B2, part, pl, st);
        case 2: return wg_cl_dispatch_n<T, 2, ACL, B1CL>(A, B1, B2, part, pl, st);
        case 3: return wg_cl_dispatch_n<T, 3, ACL, B1CL>(A, B1, B2, part, pl, st);
        case 6: return wg_cl_dispatch_n<T, 6, ACL, B1CL>(A, B1, B2, part, pl, st);
        case 12: return wg_cl_dispatch_n<T, 12, ACL, B1CL>(A, B1, B2, part, pl, st);
        default: return LMNET_ERR_UNSUPPORTED;
    }
}
template <typename T>
static int wg_cl_layouts(bool a_cl, bool b1_cl, const void* A, const void* B1, const void* B2, float* part, const WgClPlan& pl, cudaStream_t st) {
    if (a_cl && !b1_cl) return wg_cl_dispatch<T, true, false>(A, B1, B2, part, pl, st);
    if (!a_cl && b1_cl) return wg_cl_dispatch<T, false, true>(A, B1, B2, part, pl, st);
    if (a_cl && b1_cl) return wg_cl_dispatch<T, true, true>(A, B1, B2, part, pl, st);
    return LMNET_ERR_UNSUPPORTED;      // all-plane operands: lmnet_wgrad_1x1
}

}  // namespace lmnet

using namespace lmnet;

// ---- warp-streaming variant (wgrad_stream.cuh): skinny shapes of the two largest resolutions ----
// (MT, NT1, NT2, A channels-last, B1 channels-last)
#define WS_LIST(X) \
    X(2, 2, 0, false, true) X(2, 1, 0, false, true) X(3, 3, 0, false, true) \
    X(1, 3, 2, true, false) X(1, 3, 1, true, false) X(2, 6, 3, true, false) \
    X(3, 2, 0, true, true) X(1, 2, 0, true, true) X(2, 2, 0, true, true) X(1, 3, 0, true, true) \
    X(5, 3, 0, true, true) X(2, 3, 0, true, true) X(3, 3, 0, true, true) X(2, 6, 0, true, true)

struct WsPlan {
    WsGeom g;
    int MT, NT1, NT2;
    size_t smem;
};
static bool ws_plan(const lmnet_wgrad_dims* d, bool a_cl, bool b1_cl, const WgClPlan& cl, WsPlan& pl) {
    static const bool off = getenv("LMNET_WGRAD_NO_STREAM") != nullptr;
    if (off || d->P < 8 * kWsKW) return false;
    pl.MT = (d->M + 15) / 16;
    pl.NT1 = (d->N1 + 7) / 8;
    pl.NT2 = d->N2 > 0 ? (d->N2 + 7) / 8 : 0;
    bool listed = false;
#define X(mt, n1, n2, acl, bcl) if (pl.MT == mt && pl.NT1 == n1 && pl.NT2 == n2 && a_cl == acl && b1_cl == bcl) listed = true;
    WS_LIST(X)
#undef X
    if (!listed) return false;
    WsGeom& g = pl.g;
    g.B = d->B; g.M = d->M; g.N1 = d->N1; g.N2 = d->N2; g.P = d->P;
    g.pitchA = wg_cl_pitch(pl.MT * 16);
    g.pitchB1 = wg_cl_pitch(pl.NT1 * 8);
    g.pitchB2 = pl.NT2 > 0 ? wg_cl_pitch(pl.NT2 * 8) : 0;
    g.a_elems = a_cl ? kWsKW * g.pitchA : pl.MT * 16 * kWsPlanePitch;
    g.b1_elems = b1_cl ? kWsKW * g.pitchB1 : pl.NT1 * 8 * kWsPlanePitch;
    g.b2_elems = pl.NT2 > 0 ? kWsKW * g.pitchB2 : 0;
    const size_t stage_bytes = (size_t)(g.a_elems + g.b1_elems + g.b2_elems) * 2 * kWgWarps;
    g.stages = 3 * stage_bytes <= 100 * 1024 ? 3 : 2;
    const int ldn = (cl.g.NT1 + cl.g.NT2) * 8;
    const size_t red_bytes = (size_t)kWgWarps * pl.MT * 16 * ldn * sizeof(float);
    pl.smem = g.stages * stage_bytes;
    if (red_bytes > pl.smem) pl.smem = red_bytes;
    if (pl.smem > 200 * 1024) return false;
    const int64_t nchunks = (d->P + kWsKW - 1) / kWsKW;
    int per_sm = (int)((220 * 1024) / (pl.smem + 1024));
    if (per_sm > 3) per_sm = 3;
    if (per_sm < 1) per_sm = 1;
    int64_t splits = (per_sm * 148 + d->B - 1) / d->B;
    if (splits > nchunks / (2 * kWgWarps)) splits = nchunks / (2 * kWgWarps);       // >= 2 chunks per warp
    if (splits < 1) splits = 1;
    g.chunks_per_split = (int)((nchunks + splits - 1) / splits);
    g.splits = (int)((nchunks + g.chunks_per_split - 1) / g.chunks_per_split);
    return true;
}

template <typename T, int MT, int NT1, int NT2, bool ACL, bool B1CL>
static int ws_launch(const void* A, const void* B1, const void* B2, float* part, const WsPlan& pl, int ldn, int n2_off, int ones_col,
                     cudaStream_t st) {
    auto kern = wgrad_stream_kernel<T, MT, NT1, NT2, ACL, B1CL>;
    static std::atomic<size_t> granted[kMaxDevices];
    if (!ensure_smem(kern, pl.smem, granted)) return LMNET_ERR_LAUNCH;
    const WsGeom& g = pl.g;
    const double bytes = (double)g.B * (g.M + g.N1 + g.N2) * g.P * sizeof(T);
    dim3 grid(g.splits, g.B);
    LMNET_LAUNCH(KID_WGRAD_1X1, st, bytes, (kern<<<grid, kWgThreads, pl.smem, st>>>((const T*)A, (const T*)B1, (const T*)B2, part, g,
                                                                                   ldn, n2_off, ones_col)));
    return LMNET_OK;
}
template <typename T>
static int ws_dispatch(bool a_cl, bool b1_cl, const void* A, const void* B1, const void* B2, float* part, const WsPlan& pl, int ldn,
                       int n2_off, int ones_col, cudaStream_t st) {
#define X(mt, n1, n2, acl, bcl) \
    if (pl.MT == mt && pl.NT1 == n1 && pl.NT2 == n2 && a_cl == acl && b1_cl == bcl) \
        return ws_launch<T, mt, n1, n2, acl, bcl>(A, B1, B2, part, pl, ldn, n2_off, ones_col, st);
    WS_LIST(X)
#undef X
    return LMNET_ERR_UNSUPPORTED;
}

extern "C" int lmnet_wgrad_1x1_cl_supported(const lmnet_wgrad_dims* d, int a_cl, int b1_cl, int dtype) {
    if (dtype != LMNET_BF16 && dtype != LMNET_F16) return 0;
    if (!a_cl && !b1_cl) return 0;
    WgClPlan pl{};
    return wg_cl_plan(d, a_cl != 0, b1_cl != 0, pl) ? 1 : 0;
}
extern "C" size_t lmnet_wgrad_1x1_cl_workspace_bytes(const lmnet_wgrad_dims* d, int a_cl, int b1_cl) {
    WgClPlan pl{};
    if (!wg_cl_plan(d, a_cl != 0, b1_cl != 0, pl)) return 0;
    int splits = pl.g.splits;
    WsPlan ws;
    if (ws_plan(d, a_cl != 0, b1_cl != 0, pl, ws) && ws.g.splits > splits) splits = ws.g.splits;
    return (size_t)pl.g.B * splits * pl.g.M * ((pl.g.NT1 + pl.g.NT2) * 8) * sizeof(float);
}
static int wgrad_1x1_cl_impl(const void* A, const void* B1, const void* B2, float* dW, float* drow,
                             void* workspace, size_t workspace_bytes, const lmnet_wgrad_dims* d, int a_cl, int b1_cl,
                             int dtype, void* stream, bool batch_sum) {
    if (!lmnet_wgrad_1x1_cl_supported(d, a_cl, b1_cl, dtype)) return LMNET_ERR_UNSUPPORTED;
    if (!A || !B1 || (d->N2 > 0 && !B2) || !dW || !workspace) return LMNET_ERR_INVALID_ARG;
    if (workspace_bytes < lmnet_wgrad_1x1_cl_workspace_bytes(d, a_cl, b1_cl)) return LMNET_ERR_WORKSPACE;
    if ((uintptr_t)A % 16 || (uintptr_t)B1 % 16 || (uintptr_t)B2 % 16) return LMNET_ERR_UNSUPPORTED;
    WgClPlan pl{};
    wg_cl_plan(d, a_cl != 0, b1_cl != 0, pl);
    cudaStream_t st = (cudaStream_t)stream;
    float* part = (float*)workspace;
    const int Ntot = d->N1 + d->N2;
    const int64_t total = (int64_t)d->B * d->M * (Ntot + 1);
    const int ones_col = d->N2 > 0 ? pl.g.NT1 * 8 + d->N2 : d->N1;
    const int ldn = (pl.g.NT1 + pl.g.NT2) * 8, n2_off = pl.g.NT1 * 8;
    int splits = pl.g.splits;
    WsPlan ws;
    int rc;
    if (ws_plan(d, a_cl != 0, b1_cl != 0, pl, ws)) {
        splits = ws.g.splits;
        rc = dtype == LMNET_BF16 ? ws_dispatch<__nv_bfloat16>(a_cl != 0, b1_cl != 0, A, B1, B2, part, ws, ldn, n2_off, ones_col, st)
                                 : ws_dispatch<__half>(a_cl != 0, b1_cl != 0, A, B1, B2, part, ws, ldn, n2_off, ones_col, st);
    } else {
        rc = dtype == LMNET_BF16 ? wg_cl_layouts<__nv_bfloat16>(a_cl != 0, b1_cl != 0, A, B1, d->N2 > 0 ? B2 : B1, part, pl, st)
                                 : wg_cl_layouts<__half>(a_cl != 0, b1_cl != 0, A, B1, d->N2 > 0 ? B2 : B1, part, pl, st);
    }
    if (rc != LMNET_OK) return rc;
    if (batch_sum) {
        const int n_out = d->M * (Ntot + 1);
        const int threads = ldn <= 256 ? 256 : ((ldn + 31) / 32) * 32;
        const size_t red_smem = (size_t)(threads / ldn) * ldn * sizeof(float);
        (void)n_out;
        LMNET_LAUNCH(KID_WGRAD_REDUCE, st, 0, (wgrad_reduce_bsum_kernel<<<d->M, threads, red_smem, st>>>(
            part, dW, drow, d->B, splits, d->M, Ntot, ldn, d->N1, n2_off, ones_col)));
        return LMNET_OK;
    }
    LMNET_LAUNCH(KID_WGRAD_REDUCE, st, 0, (wgrad_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        part, dW, drow, d->B, splits, d->M, Ntot, ldn, d->N1, n2_off, ones_col)));
    return LMNET_OK;
}

extern "C" int lmnet_wgrad_1x1_cl(const void* A, const void* B1, const void* B2, float* dW, float* drow,
                                  void* workspace, size_t workspace_bytes, const lmnet_wgrad_dims* d, int a_cl, int b1_cl,
                                  int dtype, void* stream) {
    return wgrad_1x1_cl_impl(A, B1, B2, dW, drow, workspace, workspace_bytes, d, a_cl, b1_cl, dtype, stream, false);
}
extern "C" int lmnet_wgrad_1x1_cl_sum(const void* A, const void* B1, const void* B2, float* dW, float* drow,
                                      void* workspace, size_t workspace_bytes, const lmnet_wgrad_dims* d, int a_cl, int b1_cl,
                                      int dtype, void* stream) {
    return wgrad_1x1_cl_impl(A, B1, B2, dW, drow, workspace, workspace_bytes, d, a_cl, b1_cl, dtype, stream, true);
}

// Parameter gradients of the SE-gated pointwise + shortcut pair out of the per-image weight gradient dW [B][M][E + Cin]
// (W_b = Wpw * gate_b): dWpw[m][e] = sum_b dW[b][m][e] * gate[b][e];  dgate[b][e] = sum_m dW[b][m][e] * Wpw[m][e];
// dWsc[m][c] = sum_b dW[b][m][E + c];  dbias[m] = sum_b drow[b][m].  One launch instead of six ATen kernels per block.
namespace lmnet {
__global__ void __launch_bounds__(256)
pointwise_grads_kernel(const float* __restrict__ dW, const float* __restrict__ drow, const float* __restrict__ gate,
                       const float* __restrict__ wpw, int64_t wpw_sm, int64_t wpw_se, float* __restrict__ dwpw,
                       float* __restrict__ dgate, float* __restrict__ dwsc, float* __restrict__ dbias, int B, int M, int E, int Cin) {
    const int N = E + Cin;
    const int n_wpw = M * E, n_gate = B * E, n_wsc = M * Cin;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_wpw) {
        const int m = i / E, e = i - m * E;
        float a = 0.f;
#pragma unroll 8                       // loads of 8 iterations in flight together instead of one dependent load -> FMA chain
        for (int b = 0; b < B; ++b) a = fmaf(dW[((int64_t)b * M + m) * N + e], gate[(int64_t)b * E + e], a);
        dwpw[i] = a;
        return;
    }
    i -= n_wpw;
    if (i < n_gate) {
        const int b = i / E, e = i - b * E;
        float a = 0.f;
#pragma unroll 8
        for (int m = 0; m < M; ++m) a = fmaf(dW[((int64_t)b * M + m) * N + e], __ldg(wpw + m * wpw_sm + e * wpw_se), a);
        dgate[i] = a;
        return;
    }
    i -= n_gate;
    if (i < n_wsc) {
        const int m = i / Cin, c = i - m * Cin;
        float a = 0.f;
#pragma unroll 8
        for (int b = 0; b < B; ++b) a += dW[((int64_t)b * M + m) * N + E + c];
        dwsc[i] = a;
        return;
    }
    i -= n_wsc;
    if (i < M) {
        float a = 0.f;
#pragma unroll 8
        for (int b = 0; b < B; ++b) a += drow[(int64_t)b * M + i];
        dbias[i] = a;
    }
}
}  // namespace lmnet

extern "C" int lmnet_pointwise_grads(const float* dW, const float* drow, const float* gate, const float* wpw, int64_t wpw_sm,
                                     int64_t wpw_se, float* dwpw, float* dgate, float* dwsc, float* dbias, int B, int M, int E,
                                     int Cin, void* stream) {
    if (!dW || !drow || !gate || !wpw || !dwpw || !dgate || !dwsc || !dbias) return LMNET_ERR_INVALID_ARG;
    if (B <= 0 || M <= 0 || E <= 0 || Cin <= 0) return LMNET_ERR_INVALID_ARG;
    const int64_t n = (int64_t)M * E + (int64_t)B * E + (int64_t)M * Cin + M;
    if (n >= ((int64_t)1 << 30)) return LMNET_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    LMNET_LAUNCH(KID_WGRAD_REDUCE, st, 0, (pointwise_grads_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        dW, drow, gate, wpw, wpw_sm, wpw_se, dwpw, dgate, dwsc, dbias, B, M, E, Cin)));
    return LMNET_OK;
}

extern "C" int lmnet_wgrad_1x1_supported(const lmnet_wgrad_dims* d, int dtype) {
    if (d == nullptr || d->B <= 0 || d->M <= 0 || d->N1 <= 0 || d->N2 < 0 || d->P <= 0) return 0;
    if (dtype != LMNET_BF16 && dtype != LMNET_F16) return 0;
    if (d->P % 8 != 0) return 0;
    int MT, NTW, NT;
    return wg_shape(d->M, d->N1 + d->N2, MT, NTW, NT) ? 1 : 0;
}

extern "C" size_t lmnet_wgrad_1x1_workspace_bytes(const lmnet_wgrad_dims* d) {
    int MT, NTW, NT;
    if (d == nullptr || d->B <= 0 || d->M <= 0 || d->P <= 0 || !wg_shape(d->M, d->N1 + d->N2, MT, NTW, NT)) return 0;
    WgGeom g = wg_geom(d, NT);
    return (size_t)g.B * g.splits * g.M * (NT * 8) * sizeof(float);
}

extern "C" int lmnet_wgrad_1x1(const void* A, const void* B1, const void* B2, float* dW, float* drow,
                               void* workspace, size_t workspace_bytes, const lmnet_wgrad_dims* d, int dtype,
                               void* stream) {
    if (!lmnet_wgrad_1x1_supported(d, dtype)) return LMNET_ERR_UNSUPPORTED;
    if (!A || !B1 || (d->N2 > 0 && !B2) || !dW || !workspace) return LMNET_ERR_INVALID_ARG;
    if (workspace_bytes < lmnet_wgrad_1x1_workspace_bytes(d)) return LMNET_ERR_WORKSPACE;
    if ((uintptr_t)A % 16 || (uintptr_t)B1 % 16 || (uintptr_t)B2 % 16) return LMNET_ERR_UNSUPPORTED;
    int MT, NTW, NT;
    wg_shape(d->M, d->N1 + d->N2, MT, NTW, NT);
    WgGeom g = wg_geom(d, NT);
    cudaStream_t st = (cudaStream_t)stream;
    float* part = (float*)workspace;
    int rc = dtype == LMNET_BF16 ? wg_dispatch<__nv_bfloat16>(MT, NTW, A, B1, d->N2 > 0 ? B2 : B1, part, g, st)
                                 : wg_dispatch<__half>(MT, NTW, A, B1, d->N2 > 0 ? B2 : B1, part, g, st);
    if (rc != LMNET_OK) return rc;
    const int Ntot = d->N1 + d->N2;
    const int64_t total = (int64_t)d->B * d->M * (Ntot + 1);
    LMNET_LAUNCH(KID_WGRAD_REDUCE, st, 0, (wgrad_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        part, dW, drow, d->B, g.splits, d->M, Ntot, NT * 8, d->N1, d->N1, Ntot)));
    return LMNET_OK;
}
