// tcgen05 kind::tf32 path of the weight-sampling layer (placeholder until the kernels land).
#include "bbb_common.cuh"
#include "bbb_kernels.h"

namespace bbb {
bool linear_tc_supported(const LinArgs &) { return false; }
int launch_linear_fwd_tc(const LinArgs &, cudaStream_t) { return fail(BBB_EUNSUPPORTED, "tcgen05 path not built"); }
int launch_linear_bwd_tc(const LinArgs &, cudaStream_t) { return fail(BBB_EUNSUPPORTED, "tcgen05 path not built"); }
}  // namespace bbb
