// inst_f16.cu -- explicit instantiation of the forward plan for T = __half
#include "forward_impl.cuh"
namespace lpsr {
LPSR_INSTANTIATE(__half)
}
