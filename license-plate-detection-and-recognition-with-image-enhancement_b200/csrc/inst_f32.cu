// inst_f32.cu -- explicit instantiation of the forward plan for T = float
#include "forward_impl.cuh"
namespace lpsr {
LPSR_INSTANTIATE(float)
}
