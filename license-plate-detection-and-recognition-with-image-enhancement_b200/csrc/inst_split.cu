// inst_split.cu -- the fp32-accuracy forward on tensor cores (split / double-fp16 operands), its own translation unit
#include "forward_split.cuh"
namespace lpsr {
int forward_split_entry(lpsr_handle* h, const float* x, float* y, int B, int H, int W, char* ws, cudaStream_t st, bool dry, int* n_launch,
                        LaunchProfile* prof) {
  return forward_split(h, x, y, B, H, W, ws, st, dry, n_launch, prof);
}
int tap_copy_split(lpsr_handle* h, const void* src, float* dst, int B, int C, int H, int W, cudaStream_t st) {
  const long long need = (long long)B * C * H * W;
  split_nhwc_to_nchw_kernel<<<(int)std::min<long long>(4096, (need + 255) / 256), 256, 0, st>>>(static_cast<const __half*>(src), dst, B, C, H, W);
  CUDA_TRY(h, cudaGetLastError());
  return LPSR_OK;
}
}  // namespace lpsr
