// csar_tail_split_umma.cuh -- the CSAR tail of the fp32-ACCURACY mode on tensor cores (split / double-fp16 operands).  sm_100a only.
//
// Same three chained GEMMs per 128-pixel tile as csar_tail_umma.cuh (reference lpsr.py:138-153, 180-186):
//     hid = relu(W3 x_in + b3)   32 -> 64;   s_s = sigmoid(W4 hid + b4)   64 -> 32;   out = x + Wo [x_in^2 * s_c ; x_in * s_s] + bo   64 -> 32
// but every operand is a split value hi + lo (two fp16 numbers, 22 significant bits; store_chunk16_split in umma_conv.cuh) and every GEMM
// accumulates A_hi W_hi + A_lo W_hi + A_hi W_lo in fp32 (TMEM):
//   * x_in, the residual x and the output are split tensors in HBM ([16 hi | 16 lo] per 16-channel chunk, 128 bytes per pixel and 32 channels);
//   * the hidden map and the gated concat are written by the epilogue warps to shared memory as split planar K-major operands (32 KB);
//   * weights are packed by umma_pack_weights_split (K rows: hi(W) for the hi and lo chunk of every 16 channels, then lo(W); scaled by a
//     power of two so that the lo terms stay normal fp16 numbers -- the accumulator is scaled back in the epilogue);
//   * biases, the residual and the logistic are fp32 arithmetic in the epilogue (no bias / identity MMAs: a scaled bias could leave the
//     fp16 range for arbitrary weights).
// Replaces the CUDA-core csar_tail_kernel<__half, true> (2.0 ms per application at B = 256, 35 % of the fp32-mode step).
#pragma once
#include "umma_conv.cuh"

namespace lpsr {

constexpr int kTailSpSlots = 3;                                   // tile slots in flight: 48 KB of operands each
constexpr int kTailSpThreads = (4 * kTailSpSlots + 2) * 32;       // G x 4 epilogue warps, one MMA warp, one TMA producer warp

struct TailSplitParams {
  const void* x_in;                          // split tensor [BP][64] (32 real channels), dense
  const void* res; int res_pitch, res_off;   // split tensor, pitch / offset in 16-bit elements
  void* out; int out_pitch, out_off;         // split tensor
  const uint16_t* w3; const float* b3; float s3;   // 32 -> 64: packed [12][64][8]; s = 1 / (power-of-two weight scale)
  const uint16_t* w4; const float* b4; float s4;   // 64 -> 32: packed [24][32][8]
  const uint16_t* wo; const float* bo; float so;   // 64 -> 32: packed [24][32][8], channel-branch rows x kCsarChanScale
  const float* s_c;                          // [B][32] channel gates / kCsarChanScale
  long long total_px;
  int px_per_crop;
  int n_tiles;
};

struct TailSplitTmap { CUtensorMap m; };     // x_in tiles: box [128 px][64 halfs], 128-byte swizzle

__global__ void __launch_bounds__(kTailSpThreads, 1) csar_tail_split_umma_kernel(const __grid_constant__ TailSplitParams p, const __grid_constant__ TailSplitTmap tm) {
  using T = __half;
  constexpr int G = kTailSpSlots;
  constexpr uint32_t kA1 = 128 * 128, kA2 = 2 * 128 * 128, kSlot = kA1 + kA2;           // x_in tile, split 64-channel planar operand
  constexpr uint32_t kW3 = 12 * 64 * 16, kW4 = 24 * 32 * 16, kWo = kW4;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* slots = smem;
  uint8_t* w3_s = smem + (size_t)G * kSlot;
  uint8_t* w4_s = w3_s + kW3;
  uint8_t* wo_s = w4_s + kW4;
  float* b3_s = reinterpret_cast<float*>(wo_s + kWo);            // [64]
  float* b4_s = b3_s + 64;                                       // [32]
  float* bo_s = b4_s + 32;                                       // [32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bo_s + 32);
  // per slot: 0 a1_full, 1 h_full, 2 a2_ready, 3 s_full, 4 a3_ready, 5 o_full, 6 slot_free
  const uint32_t bar0 = ptx::smem_u32(bars);
  auto bar = [&](int slot, int which) { return bar0 + 8u * (uint32_t)(slot * 7 + which); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7 * G);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < kW3 / 16; i += kTailSpThreads) reinterpret_cast<uint4*>(w3_s)[i] = __ldg(reinterpret_cast<const uint4*>(p.w3) + i);
  for (uint32_t i = threadIdx.x; i < kW4 / 16; i += kTailSpThreads) {
    reinterpret_cast<uint4*>(w4_s)[i] = __ldg(reinterpret_cast<const uint4*>(p.w4) + i);
    reinterpret_cast<uint4*>(wo_s)[i] = __ldg(reinterpret_cast<const uint4*>(p.wo) + i);
  }
  if (threadIdx.x < 64) b3_s[threadIdx.x] = __ldg(p.b3 + threadIdx.x);
  if (threadIdx.x < 32) { b4_s[threadIdx.x] = __ldg(p.b4 + threadIdx.x); bo_s[threadIdx.x] = __ldg(p.bo + threadIdx.x); }
  if (threadIdx.x == 0) {
    for (int s = 0; s < G; ++s) {
      ptx::mbar_init(bar(s, 0), 1);
      ptx::mbar_init(bar(s, 1), 1);
      ptx::mbar_init(bar(s, 2), 4);   // one arrival per epilogue warp
      ptx::mbar_init(bar(s, 3), 1);
      ptx::mbar_init(bar(s, 4), 4);
      ptx::mbar_init(bar(s, 5), 1);
      ptx::mbar_init(bar(s, 6), 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 4 * G) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_my = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA

  if (warp == 4 * G + 1) {
    // =================================== TMA producer ==============================================
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&tm.m);
      for (int t = 0; t < n_my; ++t) {
        const int s = t % G;
        ptx::mbar_wait(bar(s, 6), (((uint32_t)(t / G)) & 1u) ^ 1u);        // slot free
        ptx::mbar_arrive_expect_tx(bar(s, 0), kA1);
        const long long tile = (long long)blockIdx.x + (long long)t * gridDim.x;
        ptx::tma_load_2d(ptx::smem_u32(slots + (size_t)s * kSlot), &tm.m, bar(s, 0), 0, (int)(tile * 128));
      }
    }
  } else if (warp == 4 * G) {
    // =================================== MMA issuer =================================================
    // Event driven over the G slots (as in csar_tail_umma.cuh): each slot walks its tiles through the three phases on its own.
    if (ptx::elect_one()) {
      constexpr uint32_t idesc64 = umma_idesc_f16(false, 64), idesc32 = umma_idesc_f16(false, 32);
      const uint32_t w3_lo = umma_desc_lo(ptx::smem_u32(w3_s), 64 * 16), w4_lo = umma_desc_lo(ptx::smem_u32(w4_s), 32 * 16),
                     wo_lo = umma_desc_lo(ptx::smem_u32(wo_s), 32 * 16);
      const uint32_t a1_hi = umma_desc_hi_swizzled(128);
      int tile_s[G], phase_s[G];
      uint32_t par_s[G];
#pragma unroll
      for (int s = 0; s < G; ++s) { tile_s[s] = s; phase_s[s] = 0; par_s[s] = 0; }
      int remaining = n_my;
      uint32_t idle = 0;
      while (remaining > 0) {
        bool progressed = false;
#pragma unroll
        for (int s = 0; s < G; ++s) {
          if (tile_s[s] >= n_my) continue;
          const int phase = phase_s[s];
          if (!ptx::mbar_test_wait(bar(s, phase * 2), par_s[s])) continue;   // a1_full / a2_ready / a3_ready
          ptx::tc_fence_after();
          const uint32_t slot16 = ptx::smem_u32(slots + (size_t)s * kSlot) >> 4;
          // 128 columns per slot: hi x hi products in [0,64) (H, then S in [0,32) and O in [32,64)), the lo terms A_lo W_hi + A_hi W_lo in the same
          // layout 64 columns further: the tensor core truncates every accumulation step, so the small terms get their own accumulator
          // (umma_conv.cuh, NACCW) and the epilogue adds the two blocks
          const uint32_t d = tmem_base + (uint32_t)(s * 128);
          if (phase == 0) {
            // K rows of W3 (umma_pack_weights_split, 2 real chunks): [hiW0 | hiW0 | hiW1 | hiW1 | loW0 | loW1] against the x_in tile's
            // 32-byte column blocks [hi0 | lo0 | hi1 | lo1] of its 128-byte (swizzled) rows
#pragma unroll
            for (int ks = 0; ks < 6; ++ks) {
              const uint32_t as = ks < 4 ? (uint32_t)ks : (uint32_t)(ks - 4) * 2u;
              const bool lo_term = ks >= 4 || (ks & 1);
              ptx::tc_mma_f16_lohi(d + (lo_term ? 64u : 0u), (slot16 + 2u * as) | (1u << 16), a1_hi, w3_lo + (uint32_t)(ks * 2 * 64), kUmmaDescHi, idesc64, ks > 1 ? 1u : 0u);
            }
          } else {
            // planar split operand [cg][128 rows][16 B], K-slices [hi0 | lo0 | hi1 | lo1 | hi2 | lo2 | hi3 | lo3]; weight rows as above, 4 real chunks
            const uint32_t a16 = slot16 + (kA1 >> 4);
            const uint32_t w_lo = phase == 1 ? w4_lo : wo_lo;
            const uint32_t dd = d + (phase == 1 ? 0u : 32u);
#pragma unroll
            for (int ks = 0; ks < 12; ++ks) {
              const uint32_t as = ks < 8 ? (uint32_t)ks : (uint32_t)(ks - 8) * 2u;
              const bool lo_term = ks >= 8 || (ks & 1);
              ptx::tc_mma_f16_lohi(dd + (lo_term ? 64u : 0u), (a16 + as * 2u * 128u) | (128u << 16), kUmmaDescHi, w_lo + (uint32_t)(ks * 2 * 32), kUmmaDescHi, idesc32, ks > 1 ? 1u : 0u);
            }
          }
          ptx::tc_commit(bar(s, phase * 2 + 1));                         // h_full / s_full / o_full
          progressed = true;
          if (phase == 2) { phase_s[s] = 0; par_s[s] ^= 1u; tile_s[s] += G; --remaining; }
          else phase_s[s] = phase + 1;
        }
        if (progressed) idle = 0;
        else if (++idle > (1u << 27)) __trap();                            // bounded: a protocol bug becomes a CUDA error
      }
    }
  } else {
    // =================================== epilogue groups ==========================================
    const int g = warp >> 2, wq = warp & 3, row = wq * 32 + lane;
    uint8_t* slot = slots + (size_t)g * kSlot;
    uint8_t* a2 = slot + kA1;
    const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(g * 128);
    const T* res = static_cast<const T*>(p.res);
    T* out = static_cast<T*>(p.out);
    const float s3 = p.s3, s4 = p.s4, so = p.so;
    // write 16 real channels (fp32) as a split chunk of the planar operand: K-slice 2q = hi (core matrices 4q, 4q+1), 2q+1 = lo (4q+2, 4q+3)
    auto put_split16 = [&](int q, const float (&v)[16]) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        hi[i] = pack2<T, false>(v[2 * i], v[2 * i + 1]);
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hi[i]));
        lo[i] = pack2<T, false>(v[2 * i] - f.x, v[2 * i + 1] - f.y);
      }
      uint8_t* base = a2 + ((size_t)(4 * q) * 128 + row) * 16;
      *reinterpret_cast<uint4*>(base) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(base + 2048) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
      *reinterpret_cast<uint4*>(base + 4096) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      *reinterpret_cast<uint4*>(base + 6144) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
    };
    uint32_t par = 0;
    for (int t = g; t < n_my; t += G, par ^= 1u) {
      const long long tile = (long long)blockIdx.x + (long long)t * gridDim.x;
      const long long pix = tile * 128 + row;
      const bool valid = pix < p.total_px;
      const float* sc = p.s_c + (size_t)(valid ? pix / p.px_per_crop : 0) * 32;
      // the residual row (split: 64 halfs) is requested now and used in phase 3
      uint4 rraw[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) rraw[j] = valid ? *reinterpret_cast<const uint4*>(res + (size_t)pix * p.res_pitch + p.res_off + j * 8) : make_uint4(0, 0, 0, 0);
      // ---- phase 1: hidden = relu(acc * s3 + b3) -> split operand
      ptx::mbar_wait(bar(g, 1), par);
      ptx::tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        float v[16], vl[16];
        ptx::tc_ld16_nowait(taddr + c0, v);
        ptx::tc_ld16_nowait(taddr + 64 + c0, vl);
        ptx::tc_wait_ld();
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] += vl[c];
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
          const float4 b = *reinterpret_cast<const float4*>(b3_s + c0 + c);
          v[c] = fmaxf(fmaf(v[c], s3, b.x), 0.f); v[c + 1] = fmaxf(fmaf(v[c + 1], s3, b.y), 0.f);
          v[c + 2] = fmaxf(fmaf(v[c + 2], s3, b.z), 0.f); v[c + 3] = fmaxf(fmaf(v[c + 3], s3, b.w), 0.f);
        }
        put_split16(c0 / 16, v);
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar(g, 2));
      // ---- phase 2: s_s = sigmoid(acc * s4 + b4); gated concat [x_in^2 * s_c | x_in * s_s] -> split operand (MMA2 has consumed the hidden map)
      ptx::mbar_wait(bar(g, 3), par);
      ptx::tc_fence_after();
      // x_in row from the TMA-written tile: 128-byte rows, SWIZZLE_128B: 16-byte chunk j sits at j ^ (row & 7); chunks [hi hi lo lo] per 16 channels
      float ga[2][16];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint4 xr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) xr[j] = *reinterpret_cast<const uint4*>(slot + row * 128 + (((q * 4 + j) ^ (row & 7)) << 4));
        const __half2* xh = reinterpret_cast<const __half2*>(xr);       // [0,8): hi pairs, [8,16): lo pairs
        float x[16], v[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 a = __half22float2(xh[i]), b = __half22float2(xh[8 + i]);
          x[2 * i] = a.x + b.x; x[2 * i + 1] = a.y + b.y;
        }
        {
          float vl[16];
          ptx::tc_ld16_nowait(taddr + q * 16, v);
          ptx::tc_ld16_nowait(taddr + 64 + q * 16, vl);
          ptx::tc_wait_ld();
#pragma unroll
          for (int c = 0; c < 16; ++c) v[c] += vl[c];
        }
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
          const float4 b = *reinterpret_cast<const float4*>(b4_s + q * 16 + c);
          const float4 s4v = __ldg(reinterpret_cast<const float4*>(sc + q * 16 + c));
          const float bb[4] = {b.x, b.y, b.z, b.w}, ss[4] = {s4v.x, s4v.y, s4v.z, s4v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float z = fmaf(v[c + j], s4, bb[j]);
            const float sg = __fdividef(1.f, 1.f + __expf(-z));             // fp32 logistic (ex2.approx + rcp.approx: ~2^-21 relative)
            ga[q][c + j] = x[c + j] * (x[c + j] * ss[j]);                     // channel branch x_in^2 * s_c / kCsarChanScale (lpsr.py:133-135)
            v[c + j] = x[c + j] * sg;                                         // spatial branch x_in * s_s (lpsr.py:150-153)
          }
        }
        put_split16(2 + q, v);
      }
      put_split16(0, ga[0]);
      put_split16(1, ga[1]);
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar(g, 4));
      // ---- phase 3: out = acc * so + bo + x -> split tensor
      ptx::mbar_wait(bar(g, 5), par);
      ptx::tc_fence_after();
      float o[32], ol[32];
      ptx::tc_ld16_nowait(taddr + 32, o);
      ptx::tc_ld16_nowait(taddr + 48, o + 16);
      ptx::tc_ld16_nowait(taddr + 96, ol);
      ptx::tc_ld16_nowait(taddr + 112, ol + 16);
      ptx::tc_wait_ld();
#pragma unroll
      for (int c = 0; c < 32; ++c) o[c] += ol[c];
      ptx::tc_fence_before();                                             // accumulators are in registers: free the slot
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar(g, 6));
      const int pix32 = valid ? (int)pix : -1;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const __half2* rh = reinterpret_cast<const __half2*>(rraw + 4 * q);
        float v[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 a = __half22float2(rh[i]), b = __half22float2(rh[8 + i]);
          v[2 * i] = fmaf(o[q * 16 + 2 * i], so, bo_s[q * 16 + 2 * i]) + (a.x + b.x);
          v[2 * i + 1] = fmaf(o[q * 16 + 2 * i + 1], so, bo_s[q * 16 + 2 * i + 1]) + (a.y + b.y);
        }
        store_chunk16_split<T, false>(out, p.out_pitch, p.out_off + q * 32, pix32, v);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4 * G) ptx::tmem_dealloc(tmem_base, 512);
}

inline const char* csar_tail_split_umma_launch(const TailSplitParams& pin, int num_sms, cudaStream_t st) {
  TailSplitParams p = pin;
  if (p.total_px >= (1LL << 31)) return "batch too large for 32-bit pixel indices";
  if (p.res_pitch % 8 || p.res_off % 8 || p.out_pitch % 16 || p.out_off % 16 || reinterpret_cast<uintptr_t>(p.out) % 32 ||
      reinterpret_cast<uintptr_t>(p.res) % 16)
    return "pitch/offset not aligned";
  p.n_tiles = (int)((p.total_px + 127) / 128);
  TailSplitTmap tm;
  if (const char* msg = umma_make_tmap(&tm.m, p.x_in, true, 64, 64, false, 0, 0, 0, 128, 0, p.total_px)) return msg;
  constexpr size_t kSlot = 128 * 128 + 2 * 128 * 128;
  const size_t smem = 1024 + kTailSpSlots * kSlot + 12 * 64 * 16 + 2 * 24 * 32 * 16 + 128 * 4 + (7 * kTailSpSlots + 2) * 8 + 64;
  static bool configured[kMaxDevices] = {};
  bool* flag = func_configured_flag(configured);
  if (!flag || !*flag) {
    cudaError_t e = cudaFuncSetAttribute(csar_tail_split_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    if (flag) *flag = true;
  }
  csar_tail_split_umma_kernel<<<dim3(std::min(p.n_tiles, num_sms)), dim3(kTailSpThreads), smem, st>>>(p, tm);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace lpsr
