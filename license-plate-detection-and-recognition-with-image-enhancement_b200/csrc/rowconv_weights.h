// rowconv_weights.h -- host-side weight packing for the row-streaming 3x3 kernel (layout and rationale: rowconv.cuh).
#pragma once
#include <stdint.h>

#include <vector>

#include "umma_weights.h"

namespace lpsr {

struct RowWeights {
  bool packed = false;
  int cin = 0, cout = 0;
  uint16_t* w = nullptr;   // [dx][cin/16][2 K core matrices][7 blocks x cout rows][8]: blocks = ky 2,1,0,zero,2,1,0
  float* bias = nullptr;   // [cout]
  uint32_t bytes = 0;
};

// pw: fp32 [9 taps][cin][cout] (tap = ky*3 + kx)
template <typename PutU16, typename PutF32>
bool rowconv_pack_weights(RowWeights& r, const float* pw, const float* bias, int cin, int cout, bool fp16, PutU16 put16, PutF32 put32) {
  const int n_ks = cin / 16, NB = 7 * cout;
  std::vector<uint16_t> v((size_t)3 * n_ks * 2 * NB * 8, 0);
  static const int ky_of[7] = {2, 1, 0, -1, 2, 1, 0};
  for (int dx = 0; dx < 3; ++dx)
    for (int ks = 0; ks < n_ks; ++ks)
      for (int pl = 0; pl < 2; ++pl)
        for (int blk = 0; blk < 7; ++blk) {
          const int ky = ky_of[blk];
          if (ky < 0) continue;
          for (int n = 0; n < cout; ++n)
            for (int j = 0; j < 8; ++j) {
              const int ci = ks * 16 + pl * 8 + j;
              v[((((size_t)dx * n_ks + ks) * 2 + pl) * NB + blk * cout + n) * 8 + j] = f32_to_bits16(pw[((size_t)(ky * 3 + dx) * cin + ci) * cout + n], fp16);
            }
        }
  std::vector<float> b(cout, 0.f);
  if (bias) b.assign(bias, bias + cout);
  r.w = put16(v);
  r.bias = put32(b);
  r.cin = cin; r.cout = cout;
  r.bytes = (uint32_t)(v.size() * 2);
  r.packed = (r.w != nullptr && r.bias != nullptr);
  return r.packed;
}

}  // namespace lpsr
