// tcgen05 (5th-gen tensor core) split-TF32 kernels -- placeholder until the FMA path is validated on hardware.
#include "adil_common.cuh"

namespace adil {

bool tc_shape_ok(int, int, int) { return false; }

int launch_synth_tc(float*, float*, const float*, const int64_t*, const float*, const float*, const int64_t*, int, int,
                    int, const ChannelConsts&, float, int, cudaStream_t) {
  return set_error(-4, "tcgen05 path not built");
}

int launch_grad_tc(float*, float*, float*, float*, float*, const float*, const float*, const float*, const int64_t*,
                   int, int, int, const ChannelConsts&, const AdamwDev*, int, float*, size_t, cudaStream_t) {
  return set_error(-4, "tcgen05 path not built");
}

}  // namespace adil
