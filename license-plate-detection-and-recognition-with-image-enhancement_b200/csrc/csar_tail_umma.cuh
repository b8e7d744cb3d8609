// csar_tail_umma.cuh -- the CSAR tail as ONE tensor-core kernel (16-bit modes).  sm_100a only.
//
// Reference (lpsr.py:138-153, 180-186), per pixel with x_in = conv_in(x):
//     hid = relu(W3 x_in + b3)              32 -> 64     (SpatialAttention conv 0)
//     s_s = sigmoid(W4 hid + b4)            64 -> 32     (SpatialAttention conv 2)
//     out = x + Wo [x_in^2 * s_c ; x_in * s_s] + bo      64 -> 32  (gating, conv_out, residual; s_c per crop)
// Three chained GEMMs on 128-pixel tiles; the 64-channel hidden map and the 64-channel gated concat never leave the SM:
//     TMA(x_in tile) -> MMA1 -> TMEM -> epilogue (bias, ReLU, 16-bit) -> smem A2 -> MMA2 -> TMEM -> epilogue (sigmoid, gates)
//     -> smem A3 -> MMA3 -> TMEM -> epilogue (+bias, +x) -> coalesced store.
// HBM traffic is the algorithmic 96 elements/pixel (read x_in, read x, write out) plus a second (L2-hot) read of x_in for
// the gates; the three-launch version it replaces moves 352 elements/pixel.
// Warp roles: G groups x 4 epilogue warps (group g owns tile slot g: its smem operands and 64 TMEM columns), kTailMmaWarps issuing
// threads (slots s % kTailMmaWarps), one TMA producer thread.  The MMA warp walks the G slots phase by phase (MMA1 for all slots, MMA2, MMA3), so while one
// group is in an epilogue phase the tensor core works for the others.
#pragma once
#include "umma_conv.cuh"

namespace lpsr {

constexpr int kTailGroups = 6;   // tile slots in flight (4 x 128 TMEM columns); one tile's three-phase chain is latency bound
#ifndef LPSR_TAIL_MMA_WARPS
#define LPSR_TAIL_MMA_WARPS 3
#endif
constexpr int kTailMmaWarps = LPSR_TAIL_MMA_WARPS;   // MMA issuers (slots s % kTailMmaWarps): one thread issuing 19 MMAs + 3 commits per tile was the bottleneck (tcgen05.mma blocks the issuing thread,
                                                     // a commit costs it ~250-300 clk: 2 -> 3 issuers 610 -> 547 us per application at B = 1024, 6 issuers (66 registers, spills) 616)
constexpr int kTailThreads = (4 * kTailGroups + kTailMmaWarps + 1) * 32;

struct TailUmmaParams {
  const void* x_in;            // [BP][32] dense
  const void* res; int res_pitch, res_off;
  void* out; int out_pitch, out_off;
  const uint16_t* w3; const float* b3;   // 1x1 32->64, packed [32/8][64][8]
  const uint16_t* w4; const float* b4;   // 1x1 64->32, packed [64/8][32][8]
  const uint16_t* wo; const float* bo;   // 1x1 64->32 as hi + lo weights: packed [128/8][32][8], K-slices 4..7 hold the lo terms
  const float* s_c;            // [B][32]
  long long total_px;
  int px_per_crop;
  int n_tiles;
};

struct TailTmap { CUtensorMap m; CUtensorMap r; };   // x_in tiles, residual (CSAR input) tiles

template <typename T>
__global__ void __launch_bounds__(kTailThreads, 1) csar_tail_umma_kernel(const __grid_constant__ TailUmmaParams p, const __grid_constant__ TailTmap tm) {
  constexpr int G = kTailGroups;
  // bytes: x_in tile, residual tile, and ONE 64-channel planar operand: the hidden map, then (once MMA2 has consumed it) the gated map
  constexpr uint32_t kA1 = 128 * 64, kA2 = 128 * 128, kSlot = 2 * kA1 + kA2;
  constexpr uint32_t kW3 = 32 * 64 * 2, kW4 = 64 * 32 * 2, kWo = 2 * kW4;
  // The additions of the three epilogues run on the tensor core as well (its pipe is mostly idle here, the epilogue warps are the
  // bottleneck): a constant "ones" operand times a K=16 bias operand {hi(b), lo(b), 0...} starts every accumulator at its bias
  // (hi + lo keeps the fp32 bias to 2^-17), and the residual tile times a 32x32 identity adds x exactly (1.0 * 16-bit value).
  constexpr uint32_t kOnes = 2 * 128 * 16, kB3 = 2 * 64 * 16, kB4 = 2 * 32 * 16, kIdent = 4 * 32 * 16;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* slots = smem;                                       // G slots, each 1024-aligned (40 KB)
  uint8_t* w3_s = smem + (size_t)G * kSlot;
  uint8_t* w4_s = w3_s + kW3;
  uint8_t* wo_s = w4_s + kW4;
  uint8_t* ones_s = wo_s + kWo;                                // [2 planes][128 rows][8]: channels 0 and 1 are 1.0
  uint8_t* b3_s = ones_s + kOnes;                              // [2][64][8]: k = 0 -> hi(b3[n]), k = 1 -> lo(b3[n])
  uint8_t* b4_s = b3_s + kB3;
  uint8_t* bo_s = b4_s + kB4;
  uint8_t* id_s = bo_s + kB4;                                  // [4][32][8]: identity
  uint64_t* bars = reinterpret_cast<uint64_t*>(id_s + kIdent);
  // per slot: 0 a1_full (x_in tile landed), 1 h_full, 2 a2_ready, 3 s_full, 4 a3_ready, 5 o_full, 6 slot_free (accumulators read: TMEM and the
  // planar operand may be reused), 7 xin_free (the gates have read the x_in tile), 8 res_full (residual tile landed), 9 res_free (conv_out's MMAs,
  // the last readers of the residual tile, have retired).  The two input tiles are reloaded as soon as THEY are free -- x_in while the tile's third
  // GEMM still runs -- instead of after the whole tile: the TMA latency (~1.5 k clk) leaves the tile's dependent chain.
  constexpr int kBars = 10;
  const uint32_t bar0 = ptx::smem_u32(bars);
  auto bar = [&](int slot, int which) { return bar0 + 8u * (uint32_t)(slot * kBars + which); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kBars * G);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < kW3 / 16; i += kTailThreads) reinterpret_cast<uint4*>(w3_s)[i] = __ldg(reinterpret_cast<const uint4*>(p.w3) + i);
  for (uint32_t i = threadIdx.x; i < kW4 / 16; i += kTailThreads) {
    // W4 and b4 are halved (exact in a binary float format): the logistic is evaluated as 0.5 + 0.5 * tanh(z / 2)
    uint4 w = __ldg(reinterpret_cast<const uint4*>(p.w4) + i);
    T* we = reinterpret_cast<T*>(&w);
#pragma unroll
    for (int j = 0; j < 8; ++j) we[j] = from_f32<T>(0.5f * to_f32<T>(we[j]));
    reinterpret_cast<uint4*>(w4_s)[i] = w;
  }
  for (uint32_t i = threadIdx.x; i < kWo / 16; i += kTailThreads) reinterpret_cast<uint4*>(wo_s)[i] = __ldg(reinterpret_cast<const uint4*>(p.wo) + i);
  {
    T* ones = reinterpret_cast<T*>(ones_s);
    for (uint32_t i = threadIdx.x; i < 2 * 128 * 8; i += kTailThreads) ones[i] = from_f32<T>((i < 128 * 8 && (i & 7) < 2) ? 1.f : 0.f);
    auto put_bias = [&](uint8_t* dst, const float* b, int n, float scale) {
      T* d = reinterpret_cast<T*>(dst);
      for (uint32_t i = threadIdx.x; i < (uint32_t)(2 * n * 8); i += kTailThreads) {
        const uint32_t k = i & 7, col = (i >> 3) % (uint32_t)n, plane = (i >> 3) / (uint32_t)n;
        float v = 0.f;
        if (plane == 0 && k < 2) {
          const float bv = scale * __ldg(b + col);
          const float hi = to_f32<T>(from_f32<T>(bv));
          v = k == 0 ? hi : bv - hi;
        }
        d[i] = from_f32<T>(v);
      }
    };
    put_bias(b3_s, p.b3, 64, 1.f);
    put_bias(b4_s, p.b4, 32, 0.5f);
    put_bias(bo_s, p.bo, 32, 1.f);
    T* idm = reinterpret_cast<T*>(id_s);
    for (uint32_t i = threadIdx.x; i < 4 * 32 * 8; i += kTailThreads) {
      const uint32_t k = (i >> 8) * 8 + (i & 7), col = (i >> 3) & 31;   // [cg][col][8]: K index = cg*8 + j
      idm[i] = from_f32<T>(k == col ? 1.f : 0.f);
    }
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < G; ++s) {
      ptx::mbar_init(bar(s, 0), 1);
      ptx::mbar_init(bar(s, 1), 1);
      ptx::mbar_init(bar(s, 2), 4);   // one arrival per epilogue warp
      ptx::mbar_init(bar(s, 3), 1);
      ptx::mbar_init(bar(s, 4), 4);   // one arrival per epilogue warp
      ptx::mbar_init(bar(s, 5), 1);
      ptx::mbar_init(bar(s, 6), 4);   // one arrival per epilogue warp
      ptx::mbar_init(bar(s, 7), 4);   // one arrival per epilogue warp
      ptx::mbar_init(bar(s, 8), 1);
      ptx::mbar_init(bar(s, 9), 1);   // epilogue warp 0 of the group, once it has seen o_full
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
  ptx::griddep_wait();
  const int n_my = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA

  if (warp == 4 * G + kTailMmaWarps) {
    // =================================== TMA producer ==============================================
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&tm.m);
      ptx::prefetch_tmap(&tm.r);
      // two independent sequences (x_in tiles, residual tiles), each in tile order, polled so that neither blocks the other
      int tx = 0, tr = 0;
      uint32_t idle = 0;
      while (tx < n_my || tr < n_my) {
        bool progressed = false;
        if (tx < n_my) {
          const int s = tx % G;
          if (ptx::mbar_test_wait(bar(s, 7), (((uint32_t)(tx / G)) & 1u) ^ 1u)) {        // the previous tile's gates have read the x_in tile
            ptx::mbar_arrive_expect_tx(bar(s, 0), kA1);
            const long long tile = (long long)blockIdx.x + (long long)tx * gridDim.x;
            ptx::tma_load_2d(ptx::smem_u32(slots + (size_t)s * kSlot), &tm.m, bar(s, 0), 0, (int)(tile * 128));
            ++tx;
            progressed = true;
          }
        }
        if (tr < n_my) {
          const int s = tr % G;
          if (ptx::mbar_test_wait(bar(s, 9), (((uint32_t)(tr / G)) & 1u) ^ 1u)) {        // the previous tile's conv_out MMAs have retired
            ptx::mbar_arrive_expect_tx(bar(s, 8), kA1);
            const long long tile = (long long)blockIdx.x + (long long)tr * gridDim.x;
            ptx::tma_load_2d(ptx::smem_u32(slots + (size_t)s * kSlot + kA1), &tm.r, bar(s, 8), p.res_off, (int)(tile * 128));
            ++tr;
            progressed = true;
          }
        }
        if (progressed) idle = 0;
        else if (++idle > (1u << 27)) __trap();
      }
    }
  } else if (warp >= 4 * G) {
    // =================================== MMA issuers ================================================
    const bool leader = ptx::elect_one();
    constexpr uint32_t idesc64 = umma_idesc_f16(IsBf16<T>::value, 64), idesc32 = umma_idesc_f16(IsBf16<T>::value, 32);
    const uint32_t w3_lo = umma_desc_lo(ptx::smem_u32(w3_s), 64 * 16), w4_lo = umma_desc_lo(ptx::smem_u32(w4_s), 32 * 16),
                   wo_lo = umma_desc_lo(ptx::smem_u32(wo_s), 32 * 16);
    const uint32_t a1_hi = umma_desc_hi_swizzled(64);
    const uint32_t ones_lo = umma_desc_lo(ptx::smem_u32(ones_s), 128 * 16);
    const uint32_t b3_lo = umma_desc_lo(ptx::smem_u32(b3_s), 64 * 16), b4_lo = umma_desc_lo(ptx::smem_u32(b4_s), 32 * 16),
                   bo_lo = umma_desc_lo(ptx::smem_u32(bo_s), 32 * 16), id_lo = umma_desc_lo(ptx::smem_u32(id_s), 32 * 16);
    // Event driven: every slot walks its own tile sequence (slot s: tiles s, s+G, ...) through the three phases; the issuing thread polls
    // the slots round robin and issues whatever is ready.  (A lock-step "phase p for all slots" loop drained the pipeline once per G
    // tiles: ncu showed the epilogue warps 77 % of their time in long-scoreboard stalls on the h_full / s_full / o_full barriers.)
    if (leader) {
      int tile_s[G], phase_s[G];
      uint32_t par_s[G];
#pragma unroll
      for (int s = 0; s < G; ++s) { tile_s[s] = s; phase_s[s] = 0; par_s[s] = 0; }
      const int mw = warp - 4 * G;                                // this warp issues for slots mw, mw + kTailMmaWarps, ...
      int remaining = 0;
#pragma unroll
      for (int s = 0; s < G; ++s)
        if (s % kTailMmaWarps == mw && s < n_my) remaining += (n_my - s + G - 1) / G;
      uint32_t idle = 0;
      while (remaining > 0) {
        bool progressed = false;
#pragma unroll
        for (int s = 0; s < G; ++s) {
          if (s % kTailMmaWarps != mw || tile_s[s] >= n_my) continue;
          const int phase = phase_s[s];
          if (!ptx::mbar_test_wait(bar(s, phase * 2), par_s[s])) continue;   // a1_full / a2_ready / a3_ready
          if (phase == 0 && !ptx::mbar_test_wait(bar(s, 6), par_s[s] ^ 1u)) continue;   // the slot's previous tile has been read out of TMEM
          if (phase == 2 && !ptx::mbar_test_wait(bar(s, 8), par_s[s])) continue;        // residual tile landed
          ptx::tc_fence_after();
          const uint32_t slot16 = ptx::smem_u32(slots + (size_t)s * kSlot) >> 4;
          const uint32_t d = tmem_base + (uint32_t)(s * 64);     // 64 columns per slot: H, then S in [0,32) and O in [32,64)
          if (phase == 0) {          // hid = b3 + x_in (128x32, swizzle-64B rows) * W3^T -> 64 columns
            ptx::tc_mma_f16_lohi(d, ones_lo, kUmmaDescHi, b3_lo, kUmmaDescHi, idesc64, 0u);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              ptx::tc_mma_f16_lohi(d, (slot16 + 2u * ks) | (1u << 16), a1_hi, w3_lo + (uint32_t)(ks * 2 * 64), kUmmaDescHi, idesc64, 1u);
          } else {                   // bias (+ residual * I) + 128x64 planar operand ([8 cg][128 rows][16 B]) * W^T -> 32 columns
            const uint32_t a16 = slot16 + ((2 * kA1) >> 4);
            const uint32_t w_lo = phase == 1 ? w4_lo : wo_lo;
            const uint32_t dd = d + (phase == 1 ? 0u : 32u);
            ptx::tc_mma_f16_lohi(dd, ones_lo, kUmmaDescHi, phase == 1 ? b4_lo : bo_lo, kUmmaDescHi, idesc32, 0u);
            if (phase == 2) {
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)                             // + x: the residual tile (swizzle-64B rows) times the identity
                ptx::tc_mma_f16_lohi(dd, (slot16 + (kA1 >> 4) + 2u * ks) | (1u << 16), a1_hi, id_lo + (uint32_t)(ks * 2 * 32), kUmmaDescHi, idesc32, 1u);
            }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              ptx::tc_mma_f16_lohi(dd, (a16 + (uint32_t)(ks * 2 * 128)) | (128u << 16), kUmmaDescHi, w_lo + (uint32_t)(ks * 2 * 32), kUmmaDescHi, idesc32, 1u);
            if (phase == 2) {                                               // conv_out: the lo terms of the weights over the same gated operand
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                ptx::tc_mma_f16_lohi(dd, (a16 + (uint32_t)(ks * 2 * 128)) | (128u << 16), kUmmaDescHi, w_lo + (uint32_t)((4 + ks) * 2 * 32), kUmmaDescHi, idesc32, 1u);
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
    uint8_t* a2 = slot + 2 * kA1;
    uint8_t* a3 = a2;                                            // s_full (MMA2 complete) precedes the first write of the gated map
    const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(g * 64);
    T* out = static_cast<T*>(p.out);
    uint32_t par = 0;
    for (int t = g; t < n_my; t += G, par ^= 1u) {
      const long long tile = (long long)blockIdx.x + (long long)t * gridDim.x;
      const long long pix = tile * 128 + row;
      const bool valid = pix < p.total_px;
      // operands that do not depend on the tensor core: x_in row (gates), residual row, channel gates of this crop
      // (rows past the end of the batch compute on zeros; their stores are masked)

      const float* sc = p.s_c + (size_t)(valid ? pix / p.px_per_crop : 0) * 32;
      // ---- phase 1: hidden = relu(acc + b3) -> A2
      ptx::mbar_wait(bar(g, 1), par);
      ptx::tc_fence_after();
      float blk[32];
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        float v[16];
        if (c0 % 32 == 0) {                                               // two TMEM loads in flight per wait
          ptx::tc_ld16_nowait(taddr + c0, blk);
          ptx::tc_ld16_nowait(taddr + c0 + 16, blk + 16);
          ptx::tc_wait_ld();
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = blk[c0 % 32 + c];
        *reinterpret_cast<uint4*>(a2 + ((size_t)(c0 / 8) * 128 + row) * 16) =
            make_uint4(pack2<T, true>(v[0], v[1]), pack2<T, true>(v[2], v[3]), pack2<T, true>(v[4], v[5]), pack2<T, true>(v[6], v[7]));
        *reinterpret_cast<uint4*>(a2 + ((size_t)(c0 / 8 + 1) * 128 + row) * 16) =
            make_uint4(pack2<T, true>(v[8], v[9]), pack2<T, true>(v[10], v[11]), pack2<T, true>(v[12], v[13]), pack2<T, true>(v[14], v[15]));
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar(g, 2));
      // ---- phase 2: s_s = sigmoid(acc + b4); gated concat [x_in^2 * s_c | x_in * s_s] -> A3
      ptx::mbar_wait(bar(g, 3), par);
      ptx::tc_fence_after();
      // x_in row of this pixel from the TMA-written tile (64-byte rows, SWIZZLE_64B: 16-byte chunk j sits at j ^ ((row >> 1) & 3))
      uint4 xraw[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) xraw[j] = *reinterpret_cast<const uint4*>(slot + row * 64 + ((j ^ ((row >> 1) & 3)) << 4));
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar(g, 7));                         // x_in tile read (its MMAs retired before h_full): the producer may reload it
#pragma unroll
      for (int c0 = 0; c0 < 32; c0 += 16) {
        float v[16], ga[16];
        ptx::tc_ld16(taddr + c0, v);
        const T* xe = reinterpret_cast<const T*>(xraw) + c0;
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
          const float4 s4 = __ldg(reinterpret_cast<const float4*>(sc + c0 + c));
          const float ss4[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float x = to_f32<T>(xe[c + j]);
            const float zh = v[c + j];                                       // (W4 hid + b4) / 2, bias and halving done by the MMA
            // logistic through one MUFU: 0.5 + 0.5 * tanh(z / 2); tanh.approx's 2^-11 error (2^-12 on the logistic) is at the level of
            // the 16-bit rounding of the product in both modes (bf16 2^-9, fp16 2^-12 relative)
            float th;
            asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(zh));
            const float sg = fmaf(0.5f, th, 0.5f);
            ga[c + j] = x * (x * ss4[j]);                                    // channel branch x_in^2 * s_c / 16 (lpsr.py:133-135; kCsarChanScale)
            v[c + j] = x * sg;                                               // spatial branch x_in * s_s  (lpsr.py:150-153)
          }
        }
        *reinterpret_cast<uint4*>(a3 + ((size_t)(c0 / 8) * 128 + row) * 16) =
            make_uint4(pack2<T, false>(ga[0], ga[1]), pack2<T, false>(ga[2], ga[3]), pack2<T, false>(ga[4], ga[5]), pack2<T, false>(ga[6], ga[7]));
        *reinterpret_cast<uint4*>(a3 + ((size_t)(c0 / 8 + 1) * 128 + row) * 16) =
            make_uint4(pack2<T, false>(ga[8], ga[9]), pack2<T, false>(ga[10], ga[11]), pack2<T, false>(ga[12], ga[13]), pack2<T, false>(ga[14], ga[15]));
        *reinterpret_cast<uint4*>(a3 + ((size_t)(4 + c0 / 8) * 128 + row) * 16) =
            make_uint4(pack2<T, false>(v[0], v[1]), pack2<T, false>(v[2], v[3]), pack2<T, false>(v[4], v[5]), pack2<T, false>(v[6], v[7]));
        *reinterpret_cast<uint4*>(a3 + ((size_t)(4 + c0 / 8 + 1) * 128 + row) * 16) =
            make_uint4(pack2<T, false>(v[8], v[9]), pack2<T, false>(v[10], v[11]), pack2<T, false>(v[12], v[13]), pack2<T, false>(v[14], v[15]));
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar(g, 4));
      // ---- phase 3: out = acc (bias and residual already accumulated by the tensor core)
      ptx::mbar_wait(bar(g, 5), par);
      ptx::tc_fence_after();
      if (wq == 0 && lane == 0) ptx::mbar_arrive(bar(g, 9));              // conv_out's MMAs (the residual tile's last readers) have retired
      const int pix32 = valid ? (int)pix : -1;
      ptx::tc_ld16_nowait(taddr + 32, blk);
      ptx::tc_ld16_nowait(taddr + 48, blk + 16);
      ptx::tc_wait_ld();
      ptx::tc_fence_before();                                             // accumulators are in registers: free the slot
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar(g, 6));
#pragma unroll
      for (int c0 = 0; c0 < 32; c0 += 16) {
        float v[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = blk[c0 + c];
        store_chunk16<T>(out, p.out_pitch, p.out_off + c0, pix32, v);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4 * G) ptx::tmem_dealloc(tmem_base, 512);
}

template <typename T>
inline const char* csar_tail_umma_launch(const TailUmmaParams& pin, bool fp16, int num_sms, cudaStream_t st) {
  TailUmmaParams p = pin;
  if (p.total_px >= (1LL << 31)) return "batch too large for 32-bit pixel indices";
  if (p.res_pitch % 8 || p.res_off % 8 || p.out_pitch % 16 || p.out_off % 16 || reinterpret_cast<uintptr_t>(p.out) % 32) return "pitch/offset not aligned";
  p.n_tiles = (int)((p.total_px + 127) / 128);
  TailTmap tm;
  if (const char* msg = umma_make_tmap(&tm.m, p.x_in, fp16, 32, 32, false, 0, 0, 0, 128, 0, p.total_px)) return msg;
  if (const char* msg = umma_make_tmap(&tm.r, p.res, fp16, p.res_pitch, 32, false, 0, 0, 0, 128, 0, p.total_px)) return msg;
  constexpr size_t kSlot = 2 * 128 * 64 + 128 * 128;
  const size_t smem = 1024 + kTailGroups * kSlot + 32 * 64 * 2 + 3 * 64 * 32 * 2 + (2 * 128 + 2 * 64 + 2 * 32 + 2 * 32 + 4 * 32) * 16 +
                      (10 * kTailGroups + 2) * 8 + 64;
  static bool configured[kMaxDevices] = {};
  bool* flag = func_configured_flag(configured);
  if (!flag || !*flag) {
    cudaError_t e = cudaFuncSetAttribute(csar_tail_umma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    if (flag) *flag = true;
  }
  cudaError_t e = launch_pdl(csar_tail_umma_kernel<T>, dim3(std::min(p.n_tiles, num_sms)), dim3(kTailThreads), smem, st, p, tm);
  if (e == cudaSuccess) e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

template <> inline const char* csar_tail_umma_launch<float>(const TailUmmaParams&, bool, int, cudaStream_t) { return "16-bit only"; }

}  // namespace lpsr
