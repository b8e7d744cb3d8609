// preprocess.cuh -- batched pre-processing in front of the LPSR forward (reference inference/run.py:80-96
// `preprocess_for_sr`): BGR uint8 crops of arbitrary sizes -> RGB -> Pillow's antialiased bicubic resize to out_w x out_h ->
// float32 / 255, NCHW.  Bit-exact with Pillow (src/libImaging/Resample.c): the weight tables are built on the host in double
// precision with Pillow's operation order and rounded to 22 fractional bits; the device only does Pillow's int32 accumulation
// (start 2^21, clip(acc >> 22)), horizontal pass first (only the rows the vertical pass reads), uint8 between the passes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <vector>

namespace lpsr {

constexpr int kResamplePrecisionBits = 32 - 8 - 2;

struct ResampleCrop {            // one crop of the batch (all offsets in elements of their buffers)
  long long src;                 // offset of the crop's first byte in the packed HWC BGR buffer
  int H, W;                      // source size
  int y0, rows;                  // source rows [y0, y0 + rows) feed the vertical pass (all rows when there is no vertical pass)
  int need_h, need_v;            // Pillow skips a pass whose size does not change
  int hb, hk, hks;               // horizontal: bounds table [out_w][2], weights [out_w][hks]
  int vb, vk, vks;               // vertical:   bounds table [out_h][2] (relative to y0), weights [out_h][vks]
  long long tmp;                 // offset of this crop's [rows][out_w][3] uint8 scratch
};

// Pillow's bicubic kernel (a = -0.5), Resample.c: bicubic_filter
inline double resample_bicubic(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// Resample.c: precompute_coeffs + normalize_coeffs_8bpc for the full-image box; appends bounds and weights to `tab`
inline void resample_coeffs(int in_size, int out_size, std::vector<int>& tab, int& bounds_off, int& k_off, int& ksize) {
  double scale, filterscale;
  scale = filterscale = (double)in_size / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  ksize = (int)std::ceil(support) * 2 + 1;
  bounds_off = (int)tab.size();
  tab.resize(tab.size() + (size_t)out_size * 2);
  k_off = (int)tab.size();
  tab.resize(tab.size() + (size_t)out_size * ksize, 0);
  std::vector<double> w((size_t)ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    double ww = 0.0;
    const double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      w[x] = resample_bicubic((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    for (int x = 0; x < xmax; ++x) {
      double v = w[x];
      if (ww != 0.0) v /= ww;
      tab[(size_t)k_off + (size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << kResamplePrecisionBits)) : (int)(0.5 + v * (1 << kResamplePrecisionBits));
    }
    tab[(size_t)bounds_off + 2 * xx] = xmin;
    tab[(size_t)bounds_off + 2 * xx + 1] = xmax;
  }
}

__device__ __forceinline__ uint8_t resample_clip8(int acc) {
  const int v = acc >> kResamplePrecisionBits;
  return (uint8_t)min(max(v, 0), 255);
}

// one CTA per crop: horizontal pass into the crop's scratch rows, CTA barrier, vertical pass + /255 into the NCHW output
__global__ void __launch_bounds__(256) preprocess_resize_kernel(const uint8_t* __restrict__ crops, const ResampleCrop* __restrict__ desc,
                                                                const int* __restrict__ tab, uint8_t* __restrict__ scratch,
                                                                float* __restrict__ out, int out_h, int out_w) {
  const ResampleCrop d = desc[blockIdx.x];
  const uint8_t* src = crops + d.src;
  uint8_t* tmp = scratch + d.tmp;
  const int n1 = d.rows * out_w * 3;
  for (int i = threadIdx.x; i < n1; i += blockDim.x) {
    const int c = i % 3, xx = (i / 3) % out_w, r = i / (3 * out_w);
    const uint8_t* row = src + ((size_t)(d.y0 + r) * d.W) * 3 + (2 - c);          // BGR -> RGB (cv2.COLOR_BGR2RGB)
    uint8_t v;
    if (d.need_h) {
      const int xmin = tab[d.hb + 2 * xx], xmax = tab[d.hb + 2 * xx + 1];
      const int* k = tab + d.hk + (size_t)xx * d.hks;
      int acc = 1 << (kResamplePrecisionBits - 1);
      for (int x = 0; x < xmax; ++x) acc += (int)row[(size_t)(xmin + x) * 3] * k[x];
      v = resample_clip8(acc);
    } else {
      v = row[(size_t)xx * 3];
    }
    tmp[i] = v;
  }
  __threadfence_block();
  __syncthreads();
  const int n2 = out_h * out_w * 3;
  float* o = out + (size_t)blockIdx.x * 3 * out_h * out_w;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    const int xx = i % out_w, yy = (i / out_w) % out_h, c = i / (out_w * out_h);   // NCHW order: coalesced stores
    uint8_t v;
    if (d.need_v) {
      const int ymin = tab[d.vb + 2 * yy], ymax = tab[d.vb + 2 * yy + 1];
      const int* k = tab + d.vk + (size_t)yy * d.vks;
      int acc = 1 << (kResamplePrecisionBits - 1);
      for (int y = 0; y < ymax; ++y) acc += (int)tmp[((size_t)(ymin + y) * out_w + xx) * 3 + c] * k[y];
      v = resample_clip8(acc);
    } else {
      v = tmp[((size_t)yy * out_w + xx) * 3 + c];
    }
    o[i] = __fdiv_rn((float)v, 255.0f);                                            // torchvision ToTensor: uint8 -> float32 / 255
  }
}

}  // namespace lpsr
