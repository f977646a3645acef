// debug probe: one TMA 3-D tile load in several variants (argv[1] = variant), verifies the tile against the host copy
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../lm-net_b200/csrc/tma.cuh"
using namespace lmnet;

__constant__ int BH_d, BW_d;
static int BH = 36, BW = 72;
__global__ void probe_param(const __grid_constant__ CUtensorMap tm, unsigned short* out, int c0, int c1, int c2, int variant) {
    extern __shared__ unsigned char raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 127) & ~(uintptr_t)127);
    uint64_t* bar = (uint64_t*)(smem + 8192);
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 32 * (variant & 1)) {   // variant bit0: issue from warp 1 lane 0 instead of thread 0
        mbar_arrive_expect_tx(bar, BH_d * BW_d * 2);
        tma_load_3d(smem, &tm, bar, c0, c1, c2);
    }
    mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < BH_d * BW_d; i += blockDim.x) out[i] = ((unsigned short*)smem)[i];
}
__global__ void probe_global(const CUtensorMap* tm, unsigned short* out, int c0, int c1, int c2) {
    extern __shared__ unsigned char raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 127) & ~(uintptr_t)127);
    uint64_t* bar = (uint64_t*)(smem + 8192);
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, BH_d * BW_d * 2);
        tma_load_3d(smem, tm, bar, c0, c1, c2);
    }
    mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < BH_d * BW_d; i += blockDim.x) out[i] = ((unsigned short*)smem)[i];
}
int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int planes = 16, H = argc > 2 ? atoi(argv[2]) : 40, W = argc > 3 ? atoi(argv[3]) : 72;
    std::vector<unsigned short> h((size_t)planes * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned short)(i * 7 + 1);
    unsigned short *d, *out;
    cudaMalloc(&d, h.size() * 2);
    cudaMalloc(&out, 256 * 256 * 2);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    const int c0 = argc > 4 ? atoi(argv[4]) : -2, c1 = argc > 5 ? atoi(argv[5]) : 30, c2 = 3;
    if (argc > 7) { BH = atoi(argv[6]); BW = atoi(argv[7]); }
    cudaMemcpyToSymbol(BH_d, &BH, 4); cudaMemcpyToSymbol(BW_d, &BW, 4);
    CUtensorMap tm;
    if (!tma_make_planes_map(&tm, d, planes, H, W, BH, BW)) { printf("encode failed\n"); return 1; }
    if (variant < 2) probe_param<<<1, 160, 8192 + 256>>>(tm, out, c0, c1, c2, variant);
    else {
        CUtensorMap* dtm;
        cudaMalloc(&dtm, sizeof(tm));
        cudaMemcpy(dtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
        probe_global<<<1, 160, 8192 + 256>>>(dtm, out, c0, c1, c2);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d H %d W %d c0 %d c1 %d box %dx%d: %s\n", variant, H, W, c0, c1, BH, BW, cudaGetErrorString(e));
    if (e != cudaSuccess) return 2;
    std::vector<unsigned short> o(BH * BW);
    cudaMemcpy(o.data(), out, o.size() * 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < BH; ++r)
        for (int c = 0; c < BW; ++c) {
            const int gr = c1 + r, gc = c0 + c;
            unsigned short want = (gr >= 0 && gr < H && gc >= 0 && gc < W) ? h[((size_t)c2 * H + gr) * W + gc] : 0;
            if (o[r * BW + c] != want) ++bad;
        }
    printf("mismatches: %d\n", bad);
    return bad != 0;
}
