// Instantiations of conv_tc_kernel for storage planes = 2, MMA passes per k-step = 3 (see conv_tc.cuh).
#include "conv_tc.cuh"

namespace v2x {
int conv_dispatch_m3(V2X_CONV_DISPATCH_ARGS) { return conv_dispatch<2, 3>(d, bn, a0, a1, b, t, smem, stream); }
}  // namespace v2x
