// na2d_fused_inst.cu — instantiates the fused NA kernels for ONE element type.
// Compiled three times (-DLMNET_INST=0/1/2) so the template fan-out builds in parallel.
#include "na2d_fused.cuh"

namespace lmnet {
#if LMNET_INST == 0
int fused_dispatch_f32(Op op, const FusedArgs& a, int hg) { return dispatch_k<float>(op, a, hg); }
#elif LMNET_INST == 1
int fused_dispatch_bf16(Op op, const FusedArgs& a, int hg) { return dispatch_k<__nv_bfloat16>(op, a, hg); }
#elif LMNET_INST == 2
int fused_dispatch_f16(Op op, const FusedArgs& a, int hg) { return dispatch_k<__half>(op, a, hg); }
#else
#error "LMNET_INST must be 0 (f32), 1 (bf16) or 2 (f16)"
#endif
}  // namespace lmnet
