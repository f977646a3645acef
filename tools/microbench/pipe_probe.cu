// Issue-rate probe for the instructions the NA kernels are built from (sm_100a):
//   FFMA, FFMA2 (fma.rn.f32x2), FHFMA.BF16 (fma.rn.f32.bf16, mixed precision), MUFU.EX2, F2FP pack.
// Prints thread-instructions per clock per SM with 8 independent chains per thread, 4 warps/scheduler.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096;
constexpr int kChains = 8;

template <int MODE>
__global__ void __launch_bounds__(512) probe(float* out, uint32_t seed, long long* clk) {
    float acc[kChains];
    unsigned long long acc2[kChains / 2];
    uint32_t a = seed + threadIdx.x, b = seed * 3 + threadIdx.x;
    float fa = __uint_as_float(0x3f800000u | (a & 0xffff)), fb = __uint_as_float(0x3f000000u | (b & 0xffff));
#pragma unroll
    for (int c = 0; c < kChains; ++c) acc[c] = (float)c;
#pragma unroll
    for (int c = 0; c < kChains / 2; ++c) acc2[c] = c;
    unsigned long long ab = ((unsigned long long)__float_as_uint(fa) << 32) | __float_as_uint(fb);
    long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) {
            if (MODE == 0) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[c]) : "f"(fa), "f"(fb));
            if (MODE == 1) {
                unsigned short lo = a & 0xffff, hi = b & 0xffff;
                asm volatile("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(acc[c]) : "h"(lo), "h"(hi));
            }
            if (MODE == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(acc[c]));
            if (MODE == 3 && c < kChains / 2)
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[c]) : "l"(ab), "l"(ab));
            if (MODE == 4) {
                uint32_t r;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(acc[c]), "f"(fa));
                acc[c] = __uint_as_float(r);
            }
            if (MODE == 5) {  // shift-convert (ALU pipe) interleaved with an FFMA: do they dual-issue?
                uint32_t r;
                asm volatile("shl.b32 %0, %1, 16;" : "=r"(r) : "r"(a + c));
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[c]) : "f"(__uint_as_float(r)), "f"(fb));
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += acc[c];
#pragma unroll
    for (int c = 0; c < kChains / 2; ++c) s += (float)acc2[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name, double per_iter) {
    float* out; long long* clk;
    const int blocks = 148, threads = 512;
    cudaMalloc(&out, blocks * threads * sizeof(float));
    cudaMalloc(&clk, blocks * sizeof(long long));
    probe<MODE><<<blocks, threads>>>(out, 7, clk);
    probe<MODE><<<blocks, threads>>>(out, 7, clk);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < blocks; ++i) mean += (double)h[i]; mean /= blocks;
    printf("%-28s %8.1f thread-instr/clk/SM  (%.0f clk)\n", name, per_iter * kIters * threads / mean, mean);
    cudaFree(out); cudaFree(clk);
}

int main() {
    run<0>("FFMA", kChains);
    run<1>("FHFMA.BF16 (f32 += bf16*bf16)", kChains);
    run<2>("MUFU.EX2", kChains);
    run<3>("FFMA2 (per 64-bit instr)", kChains / 2);
    run<4>("F2FP.BF16 pack", kChains);
    run<5>("SHL + FFMA pair (per pair)", kChains);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
