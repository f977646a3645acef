// se_gate.cu — the squeeze-excite gate of ReparamConv as one kernel per direction (sm_100a).
//
// Reference: SE.forward, /root/reference/core/modules.py:1030-1036 — scale = Hardsigmoid(fc2(ReLU(fc1(avgpool(x))))) with
// fc1 / fc2 two 1x1 convolutions on a [B, E, 1, 1] tensor (E <= 192, E / 4 hidden units).  The average pool already rides
// on the depthwise kernel's writer and the multiply is folded into the pointwise weights; what is left is arithmetic on
// B x E numbers that stock torch runs as ~27 tiny launches per block and direction (casts, layout transforms around
// cuDNN, bias adds, activations and their backward, bias-gradient sums): 430 launches of a 1 900-launch step
// (profiles/r02_glue_ops_before_se_fusion.txt).  Here: one CTA per image forward, one CTA for the whole backward
// (all sums in a fixed order: deterministic).  fp32 throughout.
#include "common.cuh"

namespace lmnet {

constexpr int kSeThreads = 256;

// gate[b][e] = hardsigmoid(b2[e] + sum_r W2[e][r] * relu(h1[b][r])),  h1[b][r] = b1[r] + sum_e W1[r][e] * pool[b][e]
__global__ void __launch_bounds__(kSeThreads)
se_gate_fwd_kernel(const float* __restrict__ pool, const float* __restrict__ W1, const float* __restrict__ b1,
                   const float* __restrict__ W2, const float* __restrict__ b2, float* __restrict__ gate,
                   float* __restrict__ h1_out, float* __restrict__ pre2_out, int E, int R) {
    extern __shared__ float se_smem[];
    float* s_pool = se_smem;          // [E]
    float* s_h = s_pool + E;          // [R] relu(h1)
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int e = threadIdx.x; e < E; e += kSeThreads) s_pool[e] = pool[(int64_t)b * E + e];
    __syncthreads();
    for (int r = warp; r < R; r += kSeThreads / 32) {
        float a = 0.f;
        for (int e = lane; e < E; e += 32) a = fmaf(__ldg(W1 + (int64_t)r * E + e), s_pool[e], a);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) {
            a += b1 != nullptr ? b1[r] : 0.f;
            if (h1_out != nullptr) h1_out[(int64_t)b * R + r] = a;
            s_h[r] = fmaxf(a, 0.f);
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += kSeThreads) {
        float a = b2 != nullptr ? b2[e] : 0.f;
        for (int r = 0; r < R; ++r) a = fmaf(__ldg(W2 + (int64_t)e * R + r), s_h[r], a);
        if (pre2_out != nullptr) pre2_out[(int64_t)b * E + e] = a;
        gate[(int64_t)b * E + e] = fminf(fmaxf(a * (1.f / 6.f) + 0.5f, 0.f), 1.f);
    }
}

// dpool, dW1, db1, dW2, db2 from dgate (B <= 64 images).  kSeBwdCtas CTAs: each recomputes the two small intermediate
// tables (d pre2, d h1 — B x E and B x R numbers) in shared memory and then owns a slice of every output; all sums run in
// a fixed order.
constexpr int kSeBwdCtas = 16;
__global__ void __launch_bounds__(kSeThreads)
se_gate_bwd_kernel(const float* __restrict__ dgate, const float* __restrict__ pool, const float* __restrict__ h1,
                   const float* __restrict__ pre2, const float* __restrict__ W1, const float* __restrict__ W2,
                   float* __restrict__ dpool, float* __restrict__ dW1, float* __restrict__ db1, float* __restrict__ dW2,
                   float* __restrict__ db2, int B, int E, int R) {
    extern __shared__ float se_smem[];
    float* s_d2 = se_smem;            // [B][E] d pre2
    float* s_a = s_d2 + B * E;        // [B][R] relu(h1)
    float* s_dh = s_a + B * R;        // [B][R] d h1
    float* s_pool = s_dh + B * R;     // [B][E]
    float* s_w2 = s_pool + B * E;     // [E][R] (coalesced copy: the d h1 loop walks W2 with stride R)
    const int tid = threadIdx.x, nthr = blockDim.x, cta = blockIdx.x, ncta = gridDim.x;
    for (int i = tid; i < E * R; i += nthr) s_w2[i] = __ldg(W2 + i);
    for (int i = tid; i < B * E; i += nthr) {
        const float p = pre2[i];
        s_d2[i] = (p > -3.f && p < 3.f) ? dgate[i] * (1.f / 6.f) : 0.f;
        s_pool[i] = pool[i];
    }
    for (int i = tid; i < B * R; i += nthr) s_a[i] = fmaxf(h1[i], 0.f);
    __syncthreads();
    // d h1[b][r] = (h1 > 0) * sum_e d2[b][e] W2[e][r]   (every CTA needs the whole table)
    for (int i = tid; i < B * R; i += nthr) {
        const int b = i / R, r = i - b * R;
        float a = 0.f;
        for (int e = 0; e < E; ++e) a = fmaf(s_d2[b * E + e], s_w2[e * R + r], a);
        s_dh[i] = h1[i] > 0.f ? a : 0.f;
    }
    __syncthreads();
    // this CTA's slice of the outputs (element index i with i % ncta == cta would scatter; use contiguous chunks)
    auto chunk = [&](int n, int& lo, int& hi) {
        const int per = (n + ncta - 1) / ncta;
        lo = min(n, cta * per);
        hi = min(n, lo + per);
    };
    int lo, hi;
    chunk(E * R, lo, hi);             // dW2[e][r] = sum_b d2[b][e] relu(h1[b][r])
    for (int i = lo + tid; i < hi; i += nthr) {
        const int e = i / R, r = i - e * R;
        float a = 0.f;
        for (int b = 0; b < B; ++b) a = fmaf(s_d2[b * E + e], s_a[b * R + r], a);
        dW2[i] = a;
    }
    chunk(R * E, lo, hi);             // dW1[r][e] = sum_b dh[b][r] pool[b][e]
    for (int i = lo + tid; i < hi; i += nthr) {
        const int r = i / E, e = i - r * E;
        float a = 0.f;
        for (int b = 0; b < B; ++b) a = fmaf(s_dh[b * R + r], s_pool[b * E + e], a);
        dW1[i] = a;
    }
    chunk(B * E, lo, hi);             // dpool[b][e] = sum_r dh[b][r] W1[r][e]
    for (int i = lo + tid; i < hi; i += nthr) {
        const int b = i / E, e = i - b * E;
        float a = 0.f;
        for (int r = 0; r < R; ++r) a = fmaf(s_dh[b * R + r], __ldg(W1 + (int64_t)r * E + e), a);
        dpool[i] = a;
    }
    chunk(E, lo, hi);                 // db2[e] = sum_b d2[b][e]
    for (int e = lo + tid; e < hi; e += nthr) {
        float a = 0.f;
        for (int b = 0; b < B; ++b) a += s_d2[b * E + e];
        if (db2 != nullptr) db2[e] = a;
    }
    chunk(R, lo, hi);                 // db1[r] = sum_b dh[b][r]
    for (int r = lo + tid; r < hi; r += nthr) {
        float a = 0.f;
        for (int b = 0; b < B; ++b) a += s_dh[b * R + r];
        if (db1 != nullptr) db1[r] = a;
    }
}

}  // namespace lmnet

using namespace lmnet;

static bool se_ok(int B, int E, int R) { return B > 0 && E > 0 && R > 0 && B <= 64 && E <= 1024 && R <= 256; }

extern "C" int lmnet_se_gate_supported(int B, int E, int R) {
    if (!se_ok(B, E, R)) return 0;
    return (size_t)(2 * B * E + 2 * B * R + E * R) * sizeof(float) <= 200 * 1024 ? 1 : 0;
}

extern "C" int lmnet_se_gate_fwd(const float* pool, const float* W1, const float* b1, const float* W2, const float* b2,
                                 float* gate, float* h1, float* pre2, int B, int E, int R, void* stream) {
    if (!lmnet_se_gate_supported(B, E, R)) return LMNET_ERR_UNSUPPORTED;
    if (!pool || !W1 || !W2 || !gate) return LMNET_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)(E + R) * sizeof(float);
    LMNET_LAUNCH(KID_SE_GATE_FWD, st, 0, (se_gate_fwd_kernel<<<B, kSeThreads, smem, st>>>(pool, W1, b1, W2, b2, gate, h1, pre2, E, R)));
    return LMNET_OK;
}

extern "C" int lmnet_se_gate_bwd(const float* dgate, const float* pool, const float* h1, const float* pre2, const float* W1,
                                 const float* W2, float* dpool, float* dW1, float* db1, float* dW2, float* db2, int B, int E,
                                 int R, void* stream) {
    if (!lmnet_se_gate_supported(B, E, R)) return LMNET_ERR_UNSUPPORTED;
    if (!dgate || !pool || !h1 || !pre2 || !W1 || !W2 || !dpool || !dW1 || !dW2) return LMNET_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)(2 * B * E + 2 * B * R + E * R) * sizeof(float);
    static std::atomic<size_t> granted[kMaxDevices];
    if (!ensure_smem(se_gate_bwd_kernel, smem, granted)) return LMNET_ERR_LAUNCH;
    LMNET_LAUNCH(KID_SE_GATE_BWD, st, 0, (se_gate_bwd_kernel<<<kSeBwdCtas, kSeThreads, smem, st>>>(dgate, pool, h1, pre2, W1, W2, dpool, dW1, db1, dW2,
                                                                                  db2, B, E, R)));
    return LMNET_OK;
}
