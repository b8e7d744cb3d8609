// inst_bf16.cu -- explicit instantiation of the forward plan for T = __nv_bfloat16
#include "forward_impl.cuh"
namespace lpsr {
LPSR_INSTANTIATE(__nv_bfloat16)
}
