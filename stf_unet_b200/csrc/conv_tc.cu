// tcgen05 / TMEM implicit-GEMM convolution family (placeholder until the kernel lands).
#include "common.cuh"
namespace stfb {
int conv2d_tcgen05_supported(const stfb_conv_params*) { return 0; }
int conv2d_tcgen05(const stfb_conv_params*, cudaStream_t) {
  set_error("conv2d: tcgen05 family not built");
  return STFB_ENOTSUP;
}
}  // namespace stfb
