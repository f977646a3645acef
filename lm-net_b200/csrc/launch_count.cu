// launch_count.cu — process-wide counters and status strings of liblmnet_b200.
#include "common.cuh"

namespace lmnet {
unsigned long long g_launch_count = 0;
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
