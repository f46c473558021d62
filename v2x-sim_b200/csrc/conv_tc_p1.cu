// Instantiations of conv_tc_kernel for storage planes = 1, MMA passes per k-step = 1 (see conv_tc.cuh).
#include "conv_tc.cuh"

namespace v2x {
int conv_dispatch_p1(V2X_CONV_DISPATCH_ARGS) { return conv_dispatch<1, 1>(d, bn, a0, a1, b, t, smem, stream); }
}  // namespace v2x
