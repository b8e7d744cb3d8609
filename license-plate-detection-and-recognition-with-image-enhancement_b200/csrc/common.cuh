// common.cuh -- shared types and device helpers for the LPSR sm_100a kernels.
//
// Data layout in HBM: every internal activation is NHWC ("pixel-major": all channels of
// one pixel are contiguous), element type T = float (fp32 mode) or a 16-bit float (bf16 / fp16 modes).
// A tensor is addressed as base + ((n*H + y)*W + x)*pitch + ch_off + c, so a layer can read or write a
// channel SLICE of a wider buffer: this is how dense concatenation (lpsr.py:39-40) and the final
// torch.cat of local features (lpsr.py:224) cost no copy.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace lpsr {

constexpr int kTileH = 8;      // CUDA-core kernels: one CTA = 8x32 output pixels, one thread per pixel
constexpr int kTileW = 32;
constexpr int kThreads = 256;
constexpr int kMaxChunks = 16;   // 16-channel chunks of a gather list (split mode: gff.0 reads 4 x 64 = 256 16-bit channels)

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) {
  // saturating convert: |v| > 65504 clamps instead of becoming inf (fp16 mode safety)
  v = fminf(fmaxf(v, -65504.f), 65504.f);
  return __float2half_rn(v);
}

__device__ __forceinline__ float sigmoid_f32(float v) {
  // matches torch.sigmoid to ~1 ulp: 1/(1+exp(-v)); expf (not __expf) keeps fp32-mode parity
  return 1.0f / (1.0f + expf(-v));
}

// load N contiguous elements of T (N*sizeof(T) multiple of 16 bytes, 16-byte aligned) as floats
template <typename T, int N>
__device__ __forceinline__ void load_vec(const T* __restrict__ p, float (&out)[N]) {
  constexpr int kPer = 16 / sizeof(T);
  static_assert(N % kPer == 0, "vector width");
#pragma unroll
  for (int i = 0; i < N / kPer; ++i) {
    uint4 raw = *reinterpret_cast<const uint4*>(p + i * kPer);
    const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
    for (int j = 0; j < kPer; ++j) out[i * kPer + j] = to_f32<T>(e[j]);
  }
}

template <typename T, int N>
__device__ __forceinline__ void store_vec(T* __restrict__ p, const float (&v)[N]) {
  constexpr int kPer = 16 / sizeof(T);
  if constexpr (N % kPer == 0) {
#pragma unroll
    for (int i = 0; i < N / kPer; ++i) {
      uint4 raw;
      T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
      for (int j = 0; j < kPer; ++j) e[j] = from_f32<T>(v[i * kPer + j]);
      *reinterpret_cast<uint4*>(p + i * kPer) = raw;
    }
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = from_f32<T>(v[i]);
  }
}

// ---------------------------------------------------------------------------------------------------
// Pixel (un)shuffle address maps, NHWC.  These two functions ARE the PixelUnshuffle(2)/PixelShuffle(2)
// of the reference (lpsr.py:72,79,88,95); the fused DConv kernels call them from their store path and the
// standalone lpsr_op_pixel_* entry points call the very same functions, so the bit-exact remap tests pin
// the addresses the forward uses.
//   unshuffle: out[n, c*4 + i*2 + j, h, w] = in[n, c, 2h+i, 2w+j]
//   shuffle  : out[n, c, 2h+i, 2w+j]       = in[n, c*4 + i*2 + j, h, w]
// ---------------------------------------------------------------------------------------------------
// destination element index (NHWC, out dims H/2 x W/2, pitch out_pitch) of input element (n,y,x,c)
__device__ __forceinline__ size_t unshuffle2_dst(int n, int y, int x, int c, int H, int W, int out_pitch, int out_off) {
  const int Ho = H >> 1, Wo = W >> 1;
  return ((size_t)(n * Ho + (y >> 1)) * Wo + (x >> 1)) * out_pitch + out_off + (c << 2) + ((y & 1) << 1) + (x & 1);
}
// destination element index (NHWC, out dims 2H x 2W, pitch out_pitch) of input element (n,y,x,c), c = c'*4+i*2+j
__device__ __forceinline__ size_t shuffle2_dst(int n, int y, int x, int c, int H, int W, int out_pitch, int out_off) {
  const int Wo = W << 1, Ho = H << 1;
  const int i = (c >> 1) & 1, j = c & 1;
  return ((size_t)(n * Ho + (2 * y + i)) * Wo + (2 * x + j)) * out_pitch + out_off + (c >> 2);
}

// ---------------------------------------------------------------------------------------------------
// parameter blocks (passed by value to kernels)
// ---------------------------------------------------------------------------------------------------
struct ConvParams {
  const void* in;        // NHWC T (or NCHW fp32 when the kernel is instantiated with IN_NCHW)
  int in_pitch;          // channels per pixel of the input buffer
  int n_chunks;          // input channels = n_chunks * CCH, gathered chunk by chunk (dense concat costs no copy):
  int chunk_off[kMaxChunks];          //   channel offset of logical chunk k inside its tensor
  const void* chunk_ptr[kMaxChunks];  //   base of the NHWC tensor holding chunk k (nullptr: `in`)
  int chunk_pitch[kMaxChunks];        //   channels per pixel of that tensor
  const float* w;        // packed fp32 [KS*KS][Cin][COUT] (logical channel order)
  const float* bias;     // [COUT] or nullptr
  void* out;             // NHWC T (or NCHW fp32 when OUT_SIG)
  int out_pitch, out_off;
  const void* res;       // optional residual, NHWC T, added after activation (nullptr = none)
  int res_pitch, res_off;
  int B, H, W;           // output domain
  int inH, inW;          // valid extent of the input (IN_NCHW only: zero beyond -> pad-to-4, lpsr.py:107-111)
  int relu;
};

struct DConvParams {
  const void* in; int in_pitch, in_off;
  const float* dw_w;     // [CIN][25]
  const float* dw_b;     // [CIN]
  const float* pw_w;     // [CIN][COUT]
  const float* pw_b;     // [COUT]
  void* out; int out_pitch, out_off;
  const void* res; int res_pitch, res_off;   // added after ReLU (AutoEncoder skip, lpsr.py:115)
  int B, H, W;           // input == dconv output domain (before the shuffle)
};

struct TailParams {
  const void* x_in; int xin_pitch, xin_off;     // CSAR conv_in output (lpsr.py:181)
  const void* res; int res_pitch, res_off;      // CSAR input x (residual, lpsr.py:186)
  void* out; int out_pitch, out_off;
  void* out2; int out2_pitch, out2_off;         // optional second copy of the output (nullptr = none)
  const float* sc;                              // [B][32] channel gates computed by channel_gate_kernel (nullptr: from pool_partial here)
  const float* pool_partial;                    // [B][S][32] per-slice channel sums of x_in
  int S;
  const float* ca_w1; const float* ca_b1;       // Linear 32->8  [8][32], [8]      (lpsr.py:126)
  const float* ca_w2; const float* ca_b2;       // Linear 8->32  [32][8], [32]     (lpsr.py:128)
  const float* sa_w1; const float* sa_b1;       // 1x1 32->64 packed [32][64], [64] (lpsr.py:142-144)
  const float* sa_w2; const float* sa_b2;       // 1x1 64->32 packed [64][32], [32] (lpsr.py:146-148)
  const float* co_w;  const float* co_b;        // 1x1 64->32 packed [64][32], [32] (lpsr.py:175-177)
  int B, P;                                      // P = pixels per crop
};

}  // namespace lpsr
