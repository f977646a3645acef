// launch_count.cu — process-wide counters and status strings of liblmnet_b200.
#include "common.cuh"

#include <mutex>
#include <vector>

namespace lmnet {
unsigned long long g_launch_count = 0;
bool g_profile_on = false;

struct ProfEntry {
    int kid;
    cudaEvent_t a, b;
    double bytes;
};
static std::mutex g_prof_mu;
static std::vector<ProfEntry*> g_prof;

void profile_record(int kid, cudaStream_t st, double alg_bytes, bool begin, void** slot) {
    if (begin) {
        ProfEntry* e = new ProfEntry{kid, nullptr, nullptr, alg_bytes};
        if (cudaEventCreate(&e->a) != cudaSuccess || cudaEventCreate(&e->b) != cudaSuccess) { delete e; return; }
        cudaEventRecord(e->a, st);
        *slot = e;
    } else {
        ProfEntry* e = static_cast<ProfEntry*>(*slot);
        cudaEventRecord(e->b, st);
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof.push_back(e);
    }
}

static const char* kKernelNames[KID_COUNT] = {
    "na2d_fwd", "na2d_bwd_query", "na2d_bwd_key", "na2d_drpb_reduce",
    "na2d_pn", "na2d_nn", "na2d_in", "na2d_rpbgrad", "na2d_rpbgrad_reduce",
    "dw_stats", "dw_fin_fwd", "dw_apply", "dw_pool_fin", "dw_coef_eval",
    "dw_bwd_reduce", "dw_fin_bwd", "dw_bwd_dx", "dw_bwd_dw", "dw_fin_dw",
    "bn_stats", "bn_fin_fwd", "bn_apply", "bn_bwd_reduce", "bn_fin_bwd", "bn_bwd_apply",
    "ln_fwd", "ln_bwd", "ln_bwd_params",
    "wgrad_1x1", "wgrad_reduce", "upsample2x_fwd", "upsample2x_bwd",
    "na2d_stream_fwd", "na2d_stream_bwd", "pixel_gemm", "conv3x3", "conv3x3_wgrad", "conv3x3_wgrad_reduce", "se_gate_fwd", "se_gate_bwd", "avgpool_fwd", "avgpool_bwd", "dw_bwd_frame", "seg_loss"};
}  // namespace lmnet

extern "C" int lmnet_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(lmnet::g_prof_mu);
    for (auto* e : lmnet::g_prof) { cudaEventDestroy(e->a); cudaEventDestroy(e->b); delete e; }
    lmnet::g_prof.clear();
    lmnet::g_profile_on = on != 0;
    return LMNET_OK;
}
extern "C" int lmnet_profile_num_kernels(void) { return lmnet::KID_COUNT; }
extern "C" const char* lmnet_profile_kernel_name(int kid) {
    return (kid >= 0 && kid < lmnet::KID_COUNT) ? lmnet::kKernelNames[kid] : "?";
}
extern "C" int lmnet_profile_collect(double* ms, uint64_t* launches, double* alg_bytes, int n) {
    if (!ms || !launches || !alg_bytes || n < lmnet::KID_COUNT) return LMNET_ERR_INVALID_ARG;
    for (int i = 0; i < n; ++i) { ms[i] = 0; launches[i] = 0; alg_bytes[i] = 0; }
    std::lock_guard<std::mutex> lk(lmnet::g_prof_mu);
    for (auto* e : lmnet::g_prof) {
        if (cudaEventSynchronize(e->b) != cudaSuccess) return LMNET_ERR_LAUNCH;
        float t = 0.f;
        if (cudaEventElapsedTime(&t, e->a, e->b) != cudaSuccess) return LMNET_ERR_LAUNCH;
        ms[e->kid] += t; launches[e->kid] += 1; alg_bytes[e->kid] += e->bytes;
    }
    return LMNET_OK;
}

extern "C" int lmnet_abi_version(void) { return LMNET_ABI_VERSION; }

extern "C" uint64_t lmnet_launch_count(void) { return __atomic_load_n(&lmnet::g_launch_count, __ATOMIC_RELAXED); }

extern "C" const char* lmnet_status_string(int status) {
    switch (status) {
        case LMNET_OK: return "ok";
        case LMNET_ERR_INVALID_ARG: return "invalid argument";
        case LMNET_ERR_UNSUPPORTED: return "unsupported dtype/shape/alignment";
        case LMNET_ERR_LAUNCH: return "CUDA launch error (or no CUDA device)";
        case LMNET_ERR_WORKSPACE: return "workspace too small";
        default: return "unknown status";
    }
}
