// tcgen05 (5th-gen tensor core) implicit-GEMM convolution — placeholder until the TMA/TMEM kernel lands.
#include "common.cuh"

bool tt_conv2d_tc_supported(const tt_conv_desc*, const void*, const void*, const void*) { return false; }
int tt_conv2d_tc(const tt_conv_desc*, const float*, const float*, const float*, const float*, const float*, float*,
                 cudaStream_t) {
  tt_set_error("tt_conv2d: tcgen05 path not built");
  return TT_ERR_UNSUPPORTED;
}
