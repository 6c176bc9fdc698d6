// placeholder until the tcgen05 kernels land
#include "kernels.h"
namespace l3 {
int conv_tc_supported() { return 0; }
int launch_pack_weights_tc(const float*, bf16*, int, int, int, cudaStream_t) { set_error("tcgen05 path not built"); return -1; }
int launch_conv3x3_tc(const bf16*, const bf16*, const float*, bf16*, int, int, int, int, int, cudaStream_t) { set_error("tcgen05 path not built"); return -1; }
int launch_wgrad3x3_tc(const bf16*, const bf16*, float*, float*, int, int, int, int, int, cudaStream_t) { set_error("tcgen05 path not built"); return -1; }
}
