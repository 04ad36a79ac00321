// tcgen05 (TF32) path -- placeholder until the kernel lands; fails loudly.
#include "tl_common.cuh"
namespace tl {
int conv_fwd_tc(const tl_conv_desc& d, cudaStream_t stream) {
    set_error("tl_conv_fwd(tf32): tcgen05 path not built yet");
    return TL_ERR_UNSUPPORTED;
}
}  // namespace tl
