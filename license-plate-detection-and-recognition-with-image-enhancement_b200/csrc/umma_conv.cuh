// umma_conv.cuh -- TMA-fed tcgen05 / TMEM implicit-GEMM convolution (3x3 and 1x1, stride 1, zero 'same' padding) for
// the 16-bit modes.  sm_100a only.
//
// GEMM view: M = output pixels, K = Cin per tap, N = Cout (or 3*Cout when the dx taps are folded into N).
// A (pixels x channels) is never materialised as im2col:
//   * TMA (cp.async.bulk.tensor, 4-D tiled map over the NHWC activation buffer, hardware swizzle) copies a HALOED block
//     of full strip rows [rows][TW+2 pixels][Cbox channels] into shared memory once per work item and K-chunk; image
//     borders (padding='same') are the TMA's out-of-bounds zero fill, negative coordinates included.
//   * that block IS the tcgen05 K-major swizzled operand layout (row = pixel, 32/64/128 B of channels per row), with
//     rows linearised as slot = r*(TW+2) + x.  A 3x3 tap (dy,dx) is only a different descriptor START ADDRESS
//     (slot + dy*pitch + dx); the hardware swizzle is a function of the absolute smem address, so any row shift is legal
//     with base_offset = 0 (verified on B200: profiles/r1_umma_swizzle_rowshift_probe.txt).  One staged block feeds all
//     9 taps: shared memory is written once, L2->smem traffic is ~1.3x the input instead of 9x.
//   * Cout = 16 layers fold the three dx taps into N (N = 48, 3 MMAs per K-slice instead of 9): an M=128,N=16 MMA costs
//     the same ~45 clk as N=48 (profiles/r1_umma_mma_rate_microbench.txt), so folding is a 3x cut of tensor time there.
//     The epilogue then forms out[q] = D[q-1,dx=0] + D[q,dx=1] + D[q+1,dx=2] with warp shuffles (row == TMEM lane ==
//     thread); tiles overlap by 2 rows (stride 126) so only warp boundaries need a tiny smem exchange.
//
// Work decomposition: the image is cut into column strips of TW pixels (pitch = TW+2 with the halo columns); a work item
// is k tiles of 128 (126 when folded) consecutive linear positions of one strip of one crop (1x1 convs: k*128 consecutive
// pixels of the whole batch, 2-D tensor map).  Persistent CTAs (one per SM) loop over items; item buffers form a ring of
// 2 so 50-100 KB of TMA loads are in flight per SM.  The MMA loop runs tile-outer / K-inner: G tile accumulators
// (G x N fp32 TMEM columns), each with its own MMA warp and epilogue group.
// Warp roles ((5G+1) warps, G = kEpiGroups = 3 epilogue groups / TMEM tile accumulators):
//     warps 0..4G-1    G epilogue groups (group g owns TMEM accumulator g and walks the tiles t % G == g; TMEM lane quadrant =
//                      warp id % 4): tcgen05.ld, dx shifted sum, compile-time epilogue variant, one 256-bit store per row and chunk
//     warps 4G..5G-1   one MMA issuer per accumulator: tcgen05.mma (M=128, N, K=16) from one elected lane, tcgen05.commit to mbarriers
//     warp  5G         TMA producer: one elected lane arms the item's mbarrier (expect_tx) and issues the box copies
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>
#include <utility>
#include <vector>

#include "common.cuh"
#include "umma_weights.h"

namespace lpsr {

#ifndef LPSR_UMMA_EPI_GROUPS
#define LPSR_UMMA_EPI_GROUPS 3
#endif
constexpr int kEpiGroups = LPSR_UMMA_EPI_GROUPS;   // epilogue groups == TMEM tile accumulators in flight
// threads of a launch: G x 4 x CS epilogue warps, G MMA warps (one per accumulator), 1 TMA producer warp.  CS = 2 (column split): TWO
// warps per TMEM lane quadrant and accumulator, each taking half of the output channels -- the epilogue is bound by the latency of its
// own dependent instruction chain with only three warps per scheduler (ncu: 8.9 cycles per issued instruction), so twice the warps with
// half the work each is what shortens it
__host__ __device__ constexpr int umma_threads(int cs) { return (4 * kEpiGroups * cs + kEpiGroups + 1) * 32; }
constexpr int kUmmaThreads = umma_threads(1);
constexpr int kUmmaMaxKChunks = 8;      // TMA boxes (K-chunks of 16/32/64 channels) per item
constexpr int kUmmaMaxSteps = 64;       // K-steps (MMAs per tap) per tile: Cin/16, or 28 (56 with hi + lo weights) pixel-pair steps of the 7x7 conv
constexpr int kUmmaMaxK = 16;           // max M-tiles per item
constexpr int kUmmaMaxBufs = 4;         // item buffers in the shared-memory ring

struct UmmaParams {
  int n_ks;                         // K-slices of 16 channels (weights split into hi + lo terms: twice the activation K-slices)
  int n_ks_real;                    // K-slices staged in shared memory; K-slice ks >= n_ks_real re-reads the tile of (ks - n_ks_real) * wlo_step
  float acc_scale;                  // split mode: accumulator scale 2^-s undoing the power-of-two scaling of the packed weights
  int wlo_step;                     // 1; 2 in the split (hi | lo activation) mode, whose lo-weight K-slices re-read only the hi tiles
  // A operand = n_chunks TMA boxes per item: chunk c has chunk_ch[c] channels (16/32/64 -> swizzle 32/64/128 B) starting at
  // channel coordinate chunk_coff[c] of tensor map chunk_map[c], and lives at byte offset chunk_smem[c] of the item buffer
  int n_chunks;
  int chunk_ch[kUmmaMaxKChunks], chunk_coff[kUmmaMaxKChunks], chunk_map[kUmmaMaxKChunks];
  uint32_t chunk_smem[kUmmaMaxKChunks];
  int halo;                         // strip geometry: halo columns/rows on each side (1 for 3x3, 3 for 7x7)
  int rbox;                         // 3x3 / 7x7: strip rows per TMA box
  uint32_t buf_bytes;               // bytes of one item buffer (all chunks, 1024-aligned each)
  const uint16_t* w; const float* bias;
  void* out; int out_pitch, out_off;
  const void* res; int res_pitch, res_off;
  int B, H, W, relu;
  int k;                            // M-tiles (128 rows) per item
  int tstride;                      // valid rows per tile: 126 (3x3, tiles overlap by the 2 shuffle-halo rows) or 128
  int TW, pitch, n_strips, items_per_strip, n_items;
  int npx;                          // pixel slots per staged item (3x3: rbox*pitch, 1x1: 128*k)
  int n_bufs;                       // item buffers in the ring
  long long total_px;               // B*H*W (1x1 mode)
  // CSAR gate epilogue (mode 1, 1x1 Cout=32 only): v = sigmoid(acc + bias); out[off2 + c] = x_in[c] * v  (spatial branch)
  // and out[out_off + c] = x_in[c]^2 * s_c[n][c] (channel branch, reference lpsr.py:133-135,182-184)
  int mode;
  const void* aux; int aux_pitch, aux_off;   // x_in
  const float* gate;                         // s_c [B][32]
  int out_off2;
  int px_per_crop;
  float* pool;                      // kEpiPool: [n_items * k][4 warps][32] channel sums
  long long* trace;                 // LPSR_UMMA_TRACE (profiling experiments): clock64 stamps of CTA 0 [role][tile][stamp]
  int debug;                        // LPSR_UMMA_DEBUG bitmask (profiling experiments only): 1 skip MMAs, 2 skip stores, 4 skip TMA loads
};

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking phase test (event-driven issue loops poll several barriers)
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must become a trap (reported as a CUDA error), never a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
    if (spins > (1u << 26)) __trap();
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const void* tmap) { asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory"); }
// TMA tiled loads: box -> shared memory, completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::f16 (bf16 or fp16 operands, fp32 accumulate)
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// same MMA with the 64-bit descriptors given as (lo, hi) words: the issue loop only ever adds to the lo words
__device__ __forceinline__ void tc_mma_f16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Programmatic dependent launch (launch attribute programmaticStreamSerialization): the kernel may be scheduled before the previous
// grid of the stream has drained, so it waits for that grid (completion + memory flush) before touching any activation.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 16 consecutive fp32 columns: thread t of the warp gets lane (quadrant base + t)
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  tc_wait_ld();
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// issue only (no wait): several loads can be in flight before one tcgen05.wait::ld
__device__ __forceinline__ void tc_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// 32 lanes x 8 consecutive fp32 columns, issue only
__device__ __forceinline__ void tc_ld8_nowait(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

}  // namespace ptx

// instruction descriptor, kind::f16: fp32 accumulate, A and B K-major, M=128
__host__ __device__ constexpr uint32_t umma_idesc_f16(bool bf16, int N) {
  return (1u << 4)                        // c_format = F32
         | ((bf16 ? 1u : 0u) << 7)        // a_format
         | ((bf16 ? 1u : 0u) << 10)       // b_format
         | (0u << 15) | (0u << 16)        // a_major, b_major = K
         | ((uint32_t)(N >> 3) << 17)     // n_dim
         | ((uint32_t)(128 >> 4) << 24);  // m_dim
}
// shared-memory matrix descriptor (no swizzle, K-major; core matrix = 8 rows x 16 B contiguous):
//   lo = (addr >> 4) | (LBO >> 4) << 16   (LBO = byte stride between the two core matrices along K)
//   hi = (SBO >> 4) | 1 << 14             (SBO = 128 B between 8-row groups; bit 46 = sm_100 descriptor version)
// swizzled K-major (the TMA-written A blocks): row = rowbytes (32/64/128) of channels, SBO = 8 rows, LBO unused (=1),
//   hi |= layout << 29 (bits 61-63: 2 = 128B, 4 = 64B, 6 = 32B swizzle), base_offset = 0 for ANY row-shifted start
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
constexpr uint32_t kUmmaDescHi = (128u >> 4) | (1u << 14);
__device__ __forceinline__ uint32_t umma_desc_hi_swizzled(uint32_t rowbytes) {
  const uint32_t layout = rowbytes == 128 ? 2u : rowbytes == 64 ? 4u : 6u;
  return ((8u * rowbytes) >> 4) | (1u << 14) | (layout << 29);
}

// Profiling experiments only (build with LPSR_NVCC_EXTRA=-DLPSR_UMMA_TRACE_BUILD, run with LPSR_UMMA_TRACE=1): clock64 stamps of CTA 0,
// [tile][8] = {MMA warp: before / after the tmem_empty wait, after commit; epilogue warp 0 of the group: before / after the tmem_full
// wait, after the TMEM loads, tile done; MMA warp: clocks spent waiting for the item's TMA data (first tile of an item)}
#ifdef LPSR_UMMA_TRACE_BUILD
#define LPSR_DBG(bit) ((p.debug & (bit)) != 0)
#else
#define LPSR_DBG(bit) false
#endif
#ifdef LPSR_UMMA_TRACE_BUILD
#define LPSR_TRACE(cond, ti, k, v) do { if (p.trace && blockIdx.x == 0 && (cond) && (ti) < 512) p.trace[(ti) * 8 + (k)] = (v); } while (0)
#else
#define LPSR_TRACE(cond, ti, k, v) do { } while (0)
#endif

template <typename T> struct IsBf16 { static constexpr bool value = false; };
template <> struct IsBf16<__nv_bfloat16> { static constexpr bool value = true; };

// Epilogues.  FinalSigmoid: channel 0 -> sigmoid -> fp32 [pixel] (lpsr.py:273-274).  The last three serve the AutoEncoder, whose
// PixelUnshuffle / PixelShuffle (lpsr.py:72,79,88,95) are folded into the convolutions around them:
//   UnshuffleRelu: 12 real channels of pixel (y,x) -> ReLU -> channel c*4 + (y&1)*2 + (x&1) of pixel (y/2, x/2) of a half-size tensor
//   ReluResidual : ReLU, then + residual (the decoder's last stage: c0 + relu(...), lpsr.py:115)
//   Up2Store     : N = 4 sub-pixels x N/4: column block (I*2 + J) is pixel (2y + I, 2x + J) of a double-size tensor of type TOUT, optional
//                  ReLU (shallowF1's 7x7 evaluated on the half grid, lpsr.py:195-197; decoder.0 + PixelShuffle, lpsr.py:83-89)
//   Pool         : plain (+bias) store, and per (tile, warp) channel sums of the fp32 results -> pool[(tile*4 + warp)*32 + c]: the partial
//                  sums of AdaptiveAvgPool2d(1) over CSAR's x_in (lpsr.py:124,181) without reading x_in back; fixed order, so deterministic
enum { kEpiPlain = 0, kEpiGate = 1, kEpiFinalSigmoid = 2, kEpiRelu = 3, kEpiResidual = 4, kEpiUnshuffleRelu = 5, kEpiReluResidual = 6, kEpiUp2Store = 7,
       kEpiPool = 8 };

// sum over the 32 lanes of 16 values per lane (a 32 x 16 transpose-reduce): 16 shuffles; lane l returns the total of value (l >> 1) & 15
__device__ __forceinline__ float warp_column_sums16(const float (&r)[16], int lane) {
  float a8[8], a4[4], a2[2];
  {
    const bool hi = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) a8[i] = (hi ? r[8 + i] : r[i]) + __shfl_xor_sync(0xffffffffu, hi ? r[i] : r[8 + i], 16);
  }
  {
    const bool hi = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) a4[i] = (hi ? a8[4 + i] : a8[i]) + __shfl_xor_sync(0xffffffffu, hi ? a8[i] : a8[4 + i], 8);
  }
  {
    const bool hi = (lane & 4) != 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) a2[i] = (hi ? a4[2 + i] : a4[i]) + __shfl_xor_sync(0xffffffffu, hi ? a4[i] : a4[2 + i], 4);
  }
  const bool hi = (lane & 2) != 0;
  const float a1 = (hi ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, hi ? a2[0] : a2[1], 2);
  return a1 + __shfl_xor_sync(0xffffffffu, a1, 1);
}

// ---------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------
// Coalesced store of one 16-channel chunk (32 bytes) per accumulator row.  Thread = row, and consecutive rows are
// consecutive pixels of the output tensor, but a per-thread 32-byte store makes every STG.128 touch 32 different sectors
// half-way.  Instead the warp stages its 32 rows x 32 B in shared memory and re-reads them so that each of the two store
// instructions writes one contiguous 512-byte run (lane -> row 16j + lane/2, 16-byte half lane%2).
// two floats -> packed 16-bit pair (low half = a), optionally with ReLU folded into the conversion (cvt.rn.relu, one instruction)
template <typename T, bool RELU>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if constexpr (IsBf16<T>::value && RELU) {
    uint32_t r;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
  } else if constexpr (IsBf16<T>::value) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
  } else if constexpr (sizeof(T) == 2 && RELU) {
    // fp16: saturate to +-65504 instead of inf (same guarantee as from_f32<__half>), ReLU folded in: one instruction per pair
    uint32_t r;
    asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
  } else if constexpr (sizeof(T) == 2) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
  } else {
    if constexpr (RELU) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
    T t[2] = {from_f32<T>(a), from_f32<T>(b)};
    return *reinterpret_cast<uint32_t*>(t);
  }
}

// Store of one 16-channel chunk (32 bytes) per accumulator row: ONE 256-bit store per thread (STG.E.ENL2.256, sm_100).  Thread = row
// and consecutive rows are consecutive pixels, so a warp writes 32 full 32-byte sectors (1 KB contiguous for a 16-channel tensor)
// without the shared-memory staging round trip (two warp syncs, 4 smem ops, 2 shuffles) the 128-bit version needed.
// pix < 0: the row has no pixel.  Requires pitch % 16 == 0, off % 16 == 0 and a 32-byte aligned tensor.
template <typename T, bool RELU = false>
__device__ __forceinline__ void store_chunk16(T* __restrict__ out, int pitch, int off, int pix, const float (&v)[16]) {
  if (pix < 0) return;
  T* dst = out + (size_t)pix * pitch + off;
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(pack2<T, RELU>(v[0], v[1])), "r"(pack2<T, RELU>(v[2], v[3])),
               "r"(pack2<T, RELU>(v[4], v[5])), "r"(pack2<T, RELU>(v[6], v[7])), "r"(pack2<T, RELU>(v[8], v[9])),
               "r"(pack2<T, RELU>(v[10], v[11])), "r"(pack2<T, RELU>(v[12], v[13])), "r"(pack2<T, RELU>(v[14], v[15]))
               : "memory");
}

// Split (double-fp16) tensors -- the fp32-accuracy mode on tensor cores.  A real channel is carried as hi + lo, two 16-bit values with
// hi = rn16(v), lo = rn16(v - hi) (22 significant bits in fp16), stored per 16-channel chunk as [16 hi | 16 lo]: a tensor of C real
// channels is an NHWC 16-bit tensor of 2C channels, its chunks alternate hi and lo, and every convolution multiplies
// A_hi * W_hi + A_lo * W_hi + A_hi * W_lo (fp32 accumulation in TMEM; the dropped lo * lo term is 2^-22 relative).
// One 64-byte store per accumulator row and real 16-channel chunk.  pitch / off are in 16-bit elements of the split tensor.
template <typename T, bool RELU = false>
__device__ __forceinline__ void store_chunk16_split(T* __restrict__ out, int pitch, int off, int pix, const float (&vin)[16]) {
  static_assert(!IsBf16<T>::value, "split tensors are fp16 pairs");
  if (pix < 0) return;
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float a = vin[2 * i], b = vin[2 * i + 1];
    if constexpr (RELU) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
    hi[i] = pack2<T, false>(a, b);
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hi[i]));
    lo[i] = pack2<T, false>(a - f.x, b - f.y);
  }
  T* dst = out + (size_t)pix * pitch + off;
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]),
               "r"(hi[6]), "r"(hi[7]) : "memory");
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + 16), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]),
               "r"(lo[6]), "r"(lo[7]) : "memory");
}
// SP selects the tensor representation of an epilogue store of real channels [cc, cc + 16) (off = the tensor's channel offset in ELEMENTS)
template <typename T, bool RELU, bool SP>
__device__ __forceinline__ void store_out16(T* __restrict__ out, int pitch, int off, int cc, int pix, const float (&v)[16]) {
  if constexpr (SP) store_chunk16_split<T, RELU>(out, pitch, off + 2 * cc, pix, v);
  else store_chunk16<T, RELU>(out, pitch, off + cc, pix, v);
}

// half a chunk (8 channels, 16 bytes) per accumulator row: the column-split epilogue (two warps share a 32-byte pixel chunk)
template <typename T, bool RELU = false>
__device__ __forceinline__ void store_chunk8(T* __restrict__ out, int pitch, int off, int pix, const float (&v)[8]) {
  if (pix < 0) return;
  T* dst = out + (size_t)pix * pitch + off;
  asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(pack2<T, RELU>(v[0], v[1])), "r"(pack2<T, RELU>(v[2], v[3])),
               "r"(pack2<T, RELU>(v[4], v[5])), "r"(pack2<T, RELU>(v[6], v[7]))
               : "memory");
}

// MODE: kConv1x1 | kConv3x3Taps (one MMA per tap, N = Cout: used for Cout >= 32 where the MMA is already efficient)
//       | kConv3x3Fold (dx folded into N = 3*Cout: used for Cout = 16 where per-tap MMAs would be issue/smem bound)
//       | kConv7x7 (Cin = 3 padded to 8: K = 16 is a PAIR of horizontally adjacent pixels x 8 channels; 7 dy x 4 dx-pairs = 28 MMAs,
//         A rows are 16-byte pixels in a no-swizzle layout whose second K core-matrix is simply the next pixel, LBO = 16 B)
//       | kConv5x5Taps (25 taps, halo 2: the AutoEncoder's depthwise 5x5 + pointwise 1x1 pairs composed into one dense conv)
//       | kConv3x3FoldLff (last dense layer of an RDB fused with the block's 1x1 local feature fusion + residual, lpsr.py:52-61:
//         the dy = 1 MMAs carry 32 extra columns = lff over the layer's own 80 input channels (same A rows, no extra traffic); the
//         epilogue turns the folded 48 columns into g3 = relu(conv + bias), writes it to shared memory as a K = 16 operand, one more
//         MMA adds lff's g3 slice, and a second epilogue pass adds bias + x and stores the block output.  g3 never goes to HBM.)
//       | kConv5x5Fold (Cout = 16: the five dx taps folded into N = 80, 5 MMAs per K-slice instead of 25; the epilogue forms
//         out[q] = sum_dx D[q + dx - 2, dx] with shuffles by 1 and 2 rows; tiles overlap by 4 rows (stride 124))
enum { kConv1x1 = 0, kConv3x3Taps = 1, kConv3x3Fold = 2, kConv7x7 = 3, kConv5x5Taps = 4, kConv3x3FoldLff = 5, kConv5x5Fold = 6 };
// fused layer: TMEM columns per accumulator = 3*16 folded + 32 lff (hi weights) + 32 lff (lo weights); the dy = 1 weight block carries all 112
// columns, the dy = 0 / dy = 2 blocks only the 48 folded ones (an MMA has a floor of ~46 clk for N <= 64 whatever it does --
// profiles/r2_umma_variants_microbench.txt -- so the lo terms ride in the dy = 1 MMA (N = 112: 56 clk) instead of a separate N = 32 MMA)
constexpr int kLffN = 32, kLffCols = 48 + 2 * kLffN, kLffWCols = kLffCols + 48 + 48;

struct UmmaTmaps { CUtensorMap m[kUmmaMaxKChunks]; };   // one tiled map per K-chunk (its tensor, its box width)

// EPI (epilogue, compile time so the per-tile instruction stream carries no dead branches):
//   kEpiPlain: +bias | kEpiRelu: +bias, ReLU | kEpiResidual: +bias, +residual | kEpiGate: CSAR gates | kEpiFinalSigmoid
// TOUT: element type of the Up2Store output (the AutoEncoder and shallowF1 run fp16 operands in both 16-bit modes; the stage that
// feeds the trunk writes the trunk's type)
// NS = 2 (per-tap 3x3, Cout = 32): the weights are carried as hi + lo concatenated along GEMM-N (columns [0, NOUT) hi, [NOUT, 2 NOUT) lo).
// An M = 128 MMA costs max(N/2, 32 + N/4) clk, so N = 64 instead of 32 is 48 instead of 45 clk: weights at 22 significant bits for ~7 % of
// tensor time; the epilogue adds the two column blocks.
template <typename T, int NOUT, int MODE, int EPI, typename TOUT = T, int CS = 1, int NS = 1, bool SP = false>
__global__ void __launch_bounds__(umma_threads(CS), 1) umma_conv_kernel(const __grid_constant__ UmmaParams p, const __grid_constant__ UmmaTmaps tm) {
  static_assert(sizeof(T) == 2, "16-bit operands");
  constexpr int kThreadsAll = umma_threads(CS);
  constexpr int kEpiWarps = 4 * kEpiGroups * CS;               // warps [0, kEpiWarps): epilogue; then G MMA warps; then the TMA producer
  constexpr int kMmaWarp0 = kEpiWarps, kLoaderWarp = kEpiWarps + kEpiGroups;
  static_assert(CS == 1 || (CS == 2 && NOUT == 16 && MODE == kConv3x3Fold && EPI == kEpiRelu),
                "column split: the folded Cout = 16 layers");
  constexpr bool LFF = (MODE == kConv3x3FoldLff);
  constexpr bool FOLD = (MODE == kConv3x3Fold) || LFF;
  constexpr bool FOLD5 = (MODE == kConv5x5Fold);
  constexpr int HF = FOLD5 ? 2 : (FOLD ? 1 : 0);               // rows of a tile lost on each side to the dx fold
  static_assert(!LFF || (NOUT == 16 && !SP && CS == 1), "fused layer: growth rate 16, plain 16-bit tensors");
  constexpr bool K3 = (MODE != kConv1x1);
  static_assert(!SP || (CS == 1 && NS == 1 && !IsBf16<T>::value), "split tensors: fp16 pairs, plain epilogue layout");
  static_assert(NS == 1 || (NS == 2 && MODE == kConv3x3Taps && NOUT == 32 && CS == 1), "N-split weights: per-tap 3x3, Cout = 32");
  constexpr int NMMA = LFF ? kLffCols : (FOLD5 ? 5 * NOUT : (FOLD ? 3 * NOUT : NOUT * NS));   // TMEM columns per tile (and weight rows per K core-matrix)
  constexpr int KSZ = (MODE == kConv5x5Taps) ? 5 : 3;          // taps per kernel row (per-tap / folded modes)
  constexpr int NTAP = (MODE == kConv1x1 || MODE == kConv7x7) ? 1 : (FOLD5 ? 5 : (FOLD ? 3 : KSZ * KSZ));   // MMAs per K-step
  // floats of warp-boundary exchange (folded epilogues): [group][parity][warp][side][rows x NOUT]; 5-wide fold: 3 rows per side
  constexpr int XROW = FOLD5 ? 3 : 1;
  constexpr int XCH = (FOLD || FOLD5) ? kEpiGroups * 2 * 4 * 2 * XROW * NOUT : 0;
  constexpr int G = kEpiGroups;
  // TMEM accumulators per epilogue group.  Two (dense folded layers): the group's MMA warp issues the next tile into the second
  // accumulator while the epilogue warps still drain the first.  Clock traces (tools/umma_trace.py) show why: with one accumulator the three
  // MMA warps issue in the same phase (tensor pipe saturated for ~850 clk) and then all wait for their epilogues (pipe idle): the pipe is
  // 39 % busy although a tile's own MMAs take 270 clk.  Only together with the column-split epilogue (CS = 2), which is otherwise the next
  // limiter at ~620 clk per tile.
  #ifndef LPSR_UMMA_NACC_FOLD
#define LPSR_UMMA_NACC_FOLD 0
#endif
  // Split operands: the hi x hi products and the (2^-11 smaller) lo terms A_lo W_hi + A_hi W_lo accumulate in SEPARATE TMEM column blocks and are
  // added once in the epilogue.  The tensor core truncates every fp32 accumulation step; with all three terms in one accumulator the 3 K/16 steps
  // per output left a relative error of ~1.4e-6 per layer (tools/parity_stages.py), which this network amplifies 30-60x; the main block now takes
  // K/16 steps and the lo block's truncation errors are 2^-11 smaller.
  constexpr int NACCW = SP ? 2 * NMMA : NMMA;                  // TMEM columns per accumulator
  constexpr int NACC = (MODE == kConv3x3Fold && (CS == 2 || LPSR_UMMA_NACC_FOLD) && !SP && 2 * G * NMMA <= 512) ? 2 : 1;
  constexpr int GA = G * NACC;                                  // accumulators in TMEM
  constexpr uint32_t kTmemCols = (GA * NACCW <= 32) ? 32 : (GA * NACCW <= 64) ? 64 : (GA * NACCW <= 128) ? 128 : (GA * NACCW <= 256) ? 256 : 512;
  static_assert(GA * NACCW <= 512, "TMEM accumulators");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // TMA swizzle atoms need 1024-byte aligned destinations: align the carve-up by hand
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int CG = p.n_ks * 2;                                  // 8-channel groups of the whole K extent
  const uint32_t w_main_bytes = LFF ? (uint32_t)CG * kLffWCols * 16 : (uint32_t)NTAP * CG * NMMA * 16;
  const uint32_t w_bytes = w_main_bytes + (LFF ? 4u * kLffN * 16u : 0u);   // fused: + lff's g3 slice [2][32 hi | 32 lo][8]
  const uint32_t buf_bytes = p.buf_bytes;                      // multiple of 1024
  uint8_t* a_smem = smem;                                      // item buffers first (1024-aligned)
  uint8_t* w_smem = smem + (size_t)p.n_bufs * buf_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_smem + ((w_bytes + 127) & ~127u));
  // bars[0..R) full, [R..2R) empty, [2R..2R+2) tmem_full, [2R+2..2R+4) tmem_empty, then the TMEM base address,
  // then the warp-boundary exchange buffers of the folded epilogue: [2 groups][2 parities][4 warps][2 sides][NOUT] floats
  const int R = p.n_bufs;
  const uint32_t bar0 = ptx::smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (R + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * R + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * R + GA + a); };
  auto a2full_bar = [&](int a) { return bar0 + 8u * (2 * R + 2 * GA + a); };     // fused layer: g3 operand of group a is in shared memory
  auto tfull2_bar = [&](int a) { return bar0 + 8u * (2 * R + 2 * GA + G + a); }; // fused layer: lff accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * R + 2 * GA + 2 * G);
  float* xchg = reinterpret_cast<float*>(bars + 2 * R + 2 * GA + 2 * G + 2);
  // per-K-slice MMA operand table {A offset in 16-B units inside the item buffer, row bytes/16, descriptor hi word, dy shift/16}
  uint4* steps = reinterpret_cast<uint4*>(xchg + XCH);
  int* slot_base_s = reinterpret_cast<int*>(steps + kUmmaMaxSteps);   // [R] written by the producer, read by the MMA warp
  // fused layer: per group a 4 KB K=16 operand, two planes of [128 rows][8 ch] (no-swizzle K-major: row stride 16 B, LBO = 2048 B)
  uint8_t* a2_all = reinterpret_cast<uint8_t*>(slot_base_s + 8);
  // wide layers (N > 32): the bias vector lives in shared memory (broadcast LDS.128) -- per-chunk __ldg loads left the epilogue's
  // FADDs waiting on L2 latency (ncu: long-scoreboard stalls on every bias add of the N = 128 shallowF1 kernel)
  float* bias_s = reinterpret_cast<float*>(a2_all);            // same place as the fused layer's operands (that layer has N = 16)
  // fused layer: the additions of its two epilogue passes run on the tensor core (its pipe has slack, the epilogue warps do not): a
  // "ones" operand times {hi(bias), lo(bias)} starts the accumulator at the biases (b3 in the dx = 1 block, lff's in columns 48..79);
  // the residual x is part of lff's weights (1 + w on the diagonal of the block-input channels, hi + lo: exact to 22 bits).
  constexpr uint32_t kA2Bytes = SP ? 8192u : 4096u;            // g3 operand per group: [2 planes][128 rows][8] (split: hi, then lo)
  uint8_t* ones_s = a2_all + (size_t)G * kA2Bytes;             // [2 planes][128 rows][8]: k = 0, 1 are 1.0
  uint8_t* lffb_s = ones_s + 2 * 128 * 16;                     // [2][112][8]

  // ---- one-time setup ------------------------------------------------------------------------------
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.w);
    uint4* dst = reinterpret_cast<uint4*>(w_smem);
    for (uint32_t i = threadIdx.x; i < w_bytes / 16; i += kThreadsAll) dst[i] = __ldg(src + i);
  }
  if constexpr (NOUT > 32) {
    for (uint32_t i = threadIdx.x; i < (uint32_t)NOUT; i += kThreadsAll) bias_s[i] = __ldg(p.bias + i);
  }
  if constexpr (LFF) {
    T* ones = reinterpret_cast<T*>(ones_s);
    for (uint32_t i = threadIdx.x; i < 2 * 128 * 8; i += kThreadsAll) ones[i] = from_f32<T>((i < 128 * 8 && (i & 7) < 2) ? 1.f : 0.f);
    T* bb = reinterpret_cast<T*>(lffb_s);
    for (uint32_t i = threadIdx.x; i < 2 * kLffCols * 8; i += kThreadsAll) {
      const uint32_t k = i & 7, col = (i >> 3) % kLffCols, plane = (i >> 3) / kLffCols;
      float v = 0.f;
      if (plane == 0 && k < 2 && ((col >= 16 && col < 32) || (col >= 48 && col < 48 + kLffN))) {
        const float bv = __ldg(p.bias + (col < 32 ? col - 16 : col - 32));   // bias = [b3 (16) | lff bias (32)]
        const float hi = to_f32<T>(from_f32<T>(bv));
        v = k == 0 ? hi : bv - hi;
      }
      bb[i] = from_f32<T>(v);
    }
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < R; ++s) {
      ptx::mbar_init(full_bar(s), 1);                           // the producer's arrive.expect_tx; TMA completes the bytes
      ptx::mbar_init(empty_bar(s), G);                          // one tcgen05.commit per MMA warp and item
    }
    for (int a = 0; a < GA; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);
      ptx::mbar_init(tempty_bar(a), 4 * CS);                     // one arrival per epilogue warp
    }
    for (int a = 0; a < G; ++a) {
      if constexpr (LFF) {
        ptx::mbar_init(a2full_bar(a), 4 * CS);
        ptx::mbar_init(tfull2_bar(a), 1);
      }
    }
    ptx::fence_mbar_init();
  }
  if (warp == kMmaWarp0) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), kTmemCols);
    ptx::tmem_relinquish();
    if (lane == 0 && MODE == kConv7x7) {
      // step (dy, dx-pair): A start shifts by dy strip rows + 2*pair pixels; rows are 16-byte pixels (no swizzle, LBO = next pixel)
      for (int dy = 0; dy < 7; ++dy)
        for (int pr = 0; pr < 4; ++pr) steps[dy * 4 + pr] = make_uint4((uint32_t)(dy * p.pitch + 2 * pr), 1u, kUmmaDescHi, 0u);
      for (int ks = 28; ks < p.n_ks; ++ks) steps[ks] = steps[(ks - p.n_ks_real) * p.wlo_step];   // lo-weight K-steps: the same pixel pairs
    } else if (lane == 0) {
      int ks = 0;
      for (int c = 0; c < p.n_chunks; ++c) {
        const uint32_t rb16 = (uint32_t)p.chunk_ch[c] >> 3;
        for (int kk = 0; kk < (p.chunk_ch[c] >> 4); ++kk, ++ks)
          steps[ks] = make_uint4((p.chunk_smem[c] >> 4) + 2u * (uint32_t)kk, rb16, umma_desc_hi_swizzled(rb16 << 4), (uint32_t)p.pitch * rb16);
      }
      for (; ks < p.n_ks; ++ks) steps[ks] = steps[(ks - p.n_ks_real) * p.wlo_step];   // lo-weight K-slices: same activation tiles (split: the hi tiles)
    }
  }
  ptx::fence_proxy_async();      // weights were written with st.shared: make them visible to the tensor core proxy
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // No explicit griddepcontrol.launch_dependents: measured on B200 (B = 1024), triggering at kernel start costs 4% (10.54 vs 10.12 ms
  // per forward) and triggering before the teardown 8%; the implicit trigger at grid completion still shortens the launch gap.
  ptx::griddep_wait();                                          // everything above touched only weights and on-chip state

  const int n_my_items = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int rows_per_item = p.k * p.tstride;

  if (warp >= kLoaderWarp) {
    // =================================== TMA producer ==============================================
    if (warp == kLoaderWarp && ptx::elect_one()) {
      for (int c = 0; c < p.n_chunks; ++c) ptx::prefetch_tmap(&tm.m[c]);
      uint32_t item_bytes = 0;
      for (int c = 0; c < p.n_chunks; ++c) item_bytes += (uint32_t)p.npx * (uint32_t)p.chunk_ch[c] * 2u;
      int buf = 0;
      uint32_t ph = 0;
      for (int ii = 0; ii < n_my_items; ++ii, (++buf == R ? (buf = 0, ph ^= 1u) : 0u)) {
        const int item = blockIdx.x + ii * gridDim.x;
        ptx::mbar_wait(empty_bar(buf), ph ^ 1u);
        const uint32_t dst0 = ptx::smem_u32(a_smem + (size_t)buf * buf_bytes);
        const uint32_t bar = full_bar(buf);
        if (p.debug & 4) { slot_base_s[buf] = 0; ptx::mbar_arrive(bar); continue; }
        if constexpr (K3) {
          const int per_crop = p.n_strips * p.items_per_strip;
          const int n = item / per_crop;
          const int rem = item % per_crop;
          const int strip = rem / p.items_per_strip, j = rem % p.items_per_strip;
          const int qlo = j * rows_per_item - p.halo * (p.pitch + 1);   // first linear strip position any tap of this item reads
          const int y_lo = (qlo + (p.halo + 1) * p.pitch) / p.pitch - (p.halo + 1);   // floor(qlo / pitch), qlo >= -halo*(pitch+1)
          slot_base_s[buf] = qlo - y_lo * p.pitch;               // slot of accumulator row 0 of tile 0 at tap (0,0); in [0, pitch)
          for (int c = 0; c < p.n_chunks; ++c)                   // box [rbox rows][pitch px][ch]; out-of-image = zero fill
            ptx::tma_load_4d(dst0 + p.chunk_smem[c], &tm.m[p.chunk_map[c]], bar, p.chunk_coff[c], strip * p.TW - p.halo, y_lo, n);
          ptx::mbar_arrive_expect_tx(bar, item_bytes);            // after the slot_base store: release-orders it for the MMA warp
        } else {
          ptx::mbar_arrive_expect_tx(bar, item_bytes);
          const long long base_px = (long long)item * rows_per_item;
          for (int c = 0; c < p.n_chunks; ++c)
            for (int m = 0; m < p.k; ++m)                        // one 128-pixel box per tile; rows past the end are zero filled
              ptx::tma_load_2d(dst0 + p.chunk_smem[c] + (uint32_t)m * 128u * (uint32_t)p.chunk_ch[c] * 2u, &tm.m[p.chunk_map[c]], bar,
                               p.chunk_coff[c], (int)(base_px + (long long)m * 128));
        }
      }
    }
  } else if (warp >= kMmaWarp0) {
    // =================================== MMA issuers ===============================================
    // One MMA warp per accumulator / epilogue group: warp g issues the tiles whose turn is g.  The serial latency of one
    // tile (barrier wake-up, descriptor set-up, MMA issue, commit: ~1000 clk measured) then overlaps G ways instead of
    // bounding the whole CTA.  Each warp runs the (uniform) control flow; one elected lane issues MMAs and commits.
    const int mg = warp - kMmaWarp0;
    const bool leader = ptx::elect_one();
    constexpr uint32_t idesc = umma_idesc_f16(IsBf16<T>::value, NMMA);
    constexpr uint32_t idesc48 = umma_idesc_f16(IsBf16<T>::value, 48), idesc64 = umma_idesc_f16(IsBf16<T>::value, 2 * kLffN);
    const uint32_t w2_lo = umma_desc_lo(ptx::smem_u32(w_smem) + w_main_bytes, 2u * kLffN * 16);
    const uint32_t a2_lo = umma_desc_lo(ptx::smem_u32(a2_all) + (uint32_t)mg * kA2Bytes, 2048u);
    const uint32_t ones_lo = umma_desc_lo(ptx::smem_u32(ones_s), 2048u), lffb_lo = umma_desc_lo(ptx::smem_u32(lffb_s), (uint32_t)kLffCols * 16);
    // fused layer: weight blocks [dy = 1: CG x 112 | dy = 0: CG x 48 | dy = 2: CG x 48] rows of 16 bytes
    const uint32_t wl1_lo = umma_desc_lo(ptx::smem_u32(w_smem), (uint32_t)kLffCols * 16),
                   wl0_lo = umma_desc_lo(ptx::smem_u32(w_smem) + (uint32_t)CG * kLffCols * 16, 48u * 16),
                   wl2_lo = umma_desc_lo(ptx::smem_u32(w_smem) + (uint32_t)CG * (kLffCols + 48) * 16, 48u * 16);
    uint32_t a2_par = 0;
    const uint32_t w_lo = umma_desc_lo(ptx::smem_u32(w_smem), (uint32_t)NMMA * 16);
    const uint32_t cgn = (uint32_t)(CG * NMMA);               // weights: 16-byte units between taps
    const uint32_t a_smem16 = ptx::smem_u32(a_smem) >> 4, buf16_sz = buf_bytes >> 4;
    const int n_ks = p.n_ks, k_tiles = p.k;
    [[maybe_unused]] const int n_ks_real = p.n_ks_real;
    [[maybe_unused]] const bool wlo2 = p.wlo_step == 2;
    const uint32_t tstride = (uint32_t)p.tstride;
    const bool no_mma = (p.debug & 1) != 0;
    // The whole issue loop runs in ONE thread (no warp-wide control flow around an elected lane, no __syncwarp per tile): tcgen05.mma blocks the
    // issuing thread while the pipe's short queue is full, so everything else this thread does per tile adds to its MMAs' time
    if (leader) {
    uint32_t asel = 0;                                        // which of this warp's NACC accumulators the next tile uses
    uint32_t acc_bits = 3u;                                   // parities to wait for on their tmem_empty barriers (bit per accumulator)
    int turn = 0;
    int buf = 0;
    uint32_t buf_par = 0;
    for (int ii = 0; ii < n_my_items; ++ii) {
      [[maybe_unused]] const long long tw0 = clock64();
      ptx::mbar_wait(full_bar(buf), buf_par);
      ptx::tc_fence_after();
      [[maybe_unused]] const long long tw1 = clock64();
      uint32_t slot = K3 ? (uint32_t)slot_base_s[buf] : 0u;
      const uint32_t buf16 = a_smem16 + (uint32_t)buf * buf16_sz;
      for (int m = 0; m < k_tiles; ++m, slot += tstride) {
        const bool mine = (turn == mg);
        if (++turn == G) turn = 0;
        if (!mine) continue;
        LPSR_TRACE(leader, ii * k_tiles + m, 0, clock64());
        LPSR_TRACE(leader, ii * k_tiles + m, 7, m < G ? tw1 - tw0 : 0);
        const uint32_t acc = (uint32_t)mg + (uint32_t)G * asel;
        ptx::mbar_wait(tempty_bar(acc), (acc_bits >> asel) & 1u);   // the epilogue group drained this accumulator
        ptx::tc_fence_after();
        LPSR_TRACE(leader, ii * k_tiles + m, 1, clock64());
        if (leader) {
          const uint32_t d = tmem_base + acc * NACCW;
          uint32_t b_lo = w_lo;
          [[maybe_unused]] uint32_t lo_started = 0;               // split operands: the lo-term accumulator block has been initialised
          if (!no_mma) {
            if constexpr (LFF) ptx::tc_mma_f16_lohi(d, ones_lo, kUmmaDescHi, lffb_lo, kUmmaDescHi, idesc, 0u);   // accumulator := biases
#pragma unroll 1
            for (int ks = 0; ks < n_ks; ++ks) {
              const uint4 e = steps[ks];
              const uint32_t a0 = (buf16 + e.x + slot * e.y) | (1u << 16);
              if constexpr (LFF) {
                // dy = 1: 48 folded columns + lff over this K-slice (hi and lo weights) in one N = 112 MMA; dy = 0, 2: the folded columns
                const uint32_t k2 = 2u * (uint32_t)ks;
                ptx::tc_mma_f16_lohi(d, a0 + e.w, e.z, wl1_lo + k2 * kLffCols, kUmmaDescHi, idesc, 1u);
                ptx::tc_mma_f16_lohi(d, a0, e.z, wl0_lo + k2 * 48u, kUmmaDescHi, idesc48, 1u);
                ptx::tc_mma_f16_lohi(d, a0 + 2u * e.w, e.z, wl2_lo + k2 * 48u, kUmmaDescHi, idesc48, 1u);
              } else if constexpr (SP) {
                // lo term: a lo-weight K-slice (ks >= n_ks_real) or the lo chunk of a split activation tensor (odd K-slices; wlo_step == 2)
                const bool lo_term = ks >= n_ks_real || (wlo2 && (ks & 1));
                const uint32_t dd = lo_term ? d + NMMA : d;
#pragma unroll
                for (int t = 0; t < NTAP; ++t) {
                  const uint32_t shift = (FOLD || FOLD5) ? (uint32_t)t * e.w : (uint32_t)(t / KSZ) * e.w + (uint32_t)(t % KSZ) * e.y;
                  ptx::tc_mma_f16_lohi(dd, a0 + shift, e.z, b_lo + (uint32_t)t * cgn, kUmmaDescHi, idesc, lo_term ? (lo_started | (uint32_t)t) : (uint32_t)(ks | t));
                }
                if (lo_term) lo_started = 1u;
              } else {
#pragma unroll
                for (int t = 0; t < NTAP; ++t) {
                  // tap -> start shift in 16-byte units: folded: dy rows (dx lives in N); per-tap: dy rows + dx pixels
                  const uint32_t shift = (FOLD || FOLD5) ? (uint32_t)t * e.w : (uint32_t)(t / KSZ) * e.w + (uint32_t)(t % KSZ) * e.y;
                  ptx::tc_mma_f16_lohi(d, a0 + shift, e.z, b_lo + (uint32_t)t * cgn, kUmmaDescHi, idesc, (uint32_t)(ks | t));
                }
              }
              b_lo += 2 * NMMA;                               // next K-slice of the weights
            }
          }
          ptx::tc_commit(tfull_bar(acc));                     // this tile's accumulator is complete
          LPSR_TRACE(true, ii * k_tiles + m, 2, clock64());
        }
        if constexpr (LFF) {
          // second stage: the epilogue group has written g3 (K = 16) to shared memory; add lff's g3 slice to columns 48..79
          ptx::mbar_wait(a2full_bar(acc), a2_par);
          a2_par ^= 1u;
          ptx::tc_fence_after();
          if (leader) {
            if (!no_mma) {
              // lff's g3 slice, hi | lo weights along N: one N = 64 MMA into columns 48..111
              ptx::tc_mma_f16_lohi(tmem_base + acc * NACCW + 48, a2_lo, kUmmaDescHi, w2_lo, kUmmaDescHi, idesc64, 1u);
            }
            ptx::tc_commit(tfull2_bar(acc));
          }
        }
        acc_bits ^= 1u << asel;
        asel ^= (uint32_t)(NACC - 1);
      }
      if (leader) ptx::tc_commit(empty_bar(buf));             // this warp's MMAs on the item buffer have retired (count G)
      if (++buf == R) { buf = 0; buf_par ^= 1u; }
    }
    }
  } else {
    // =================================== epilogue groups ============================================
    const int wq = warp & 3;                                    // TMEM lane quadrant
    const int grp = (warp >> 2) % G;                            // group == TMEM accumulator (0..G-1)
    [[maybe_unused]] const int half = (warp >> 2) / G;          // column split: which half of the output channels (0 when CS == 1)
    const int row = wq * 32 + lane;                             // accumulator row
    constexpr int CH = 16;                                      // output channels handled per pass (bounds registers)
    // raw residual registers (uint4) prefetched per tile, before the accumulator wait
    constexpr int NRES = (EPI == kEpiReluResidual && NOUT <= 48) ? NOUT * 2 / 16 : (NOUT <= 32 ? NOUT * 2 / 16 : 1);
    constexpr int NB = NOUT <= 32 ? NOUT : 1;
    T* out = static_cast<T*>(p.out);
    const T* res = static_cast<const T*>(p.res);
    float* xg = xchg + (size_t)grp * (2 * 4 * 2 * XROW * NOUT);
    const uint32_t taddr0 = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(grp * NACCW);
    // everything the per-tile path needs lives in registers
    const int pitch = p.pitch, tstride = p.tstride, k_tiles = p.k, Himg = p.H, Wimg = p.W, halo = p.halo, TWs = p.TW;
    const int ips = p.items_per_strip, per_crop = p.n_strips * p.items_per_strip, crop_px = p.H * p.W;
    const int out_pitch = p.out_pitch, out_off = p.out_off, res_pitch = p.res_pitch, res_off = p.res_off;
    const int total_px32 = (int)p.total_px;
    const bool skip_store = (p.debug & 2) != 0;
    [[maybe_unused]] const float acc_scale = p.acc_scale;
    const int adv_y = K3 ? tstride / pitch : 0, adv_x = K3 ? tstride - adv_y * pitch : 0;   // one tile further along the strip
    const int advg_y = K3 ? (G * tstride) / pitch : 0, advg_x = K3 ? G * tstride - advg_y * pitch : 0;   // G tiles further (the group's next tile)
    float bias_r[NB];
    if constexpr (CS == 2) {
#pragma unroll
      for (int c = 0; c < 8; ++c) bias_r[c] = __ldg(p.bias + half * 8 + c);
    } else if constexpr (NOUT <= 32) {
#pragma unroll
      for (int c = 0; c < NB; ++c) bias_r[c] = __ldg(p.bias + c);
    }
    uint32_t my_par = 0;                                        // flips with every tile of this group (exchange buffers, fused layer's barriers)
    uint32_t esel = 0, full_bits = 0;                           // accumulator of the group's next tile; parities of the tmem_full barriers
    int t0mod = 0;                                              // (index of the item's first tile in the CTA's tile sequence) % G: tiles are dealt
                                                                // round robin to the groups, the same way the MMA warps count them
    for (int ii = 0; ii < n_my_items; ++ii) {
      const int item = blockIdx.x + ii * gridDim.x;
      // per item: one division per thread; per tile the row position advances incrementally (32-bit pixel indices)
      int y = 0, xs = 0, tw = 0, xbase = 0, px = 0;             // K3: strip row / column of this thread's row, crop base; 1x1: pixel index
      int nn = 0;                                               // crop of this item (shuffling epilogues)
      if constexpr (K3) {
        const int n = item / per_crop;
        nn = n;
        const int rem = item - n * per_crop;
        const int strip = rem / ips, j = rem - strip * ips;
        xbase = strip * TWs - halo;                             // image x of strip column xs is xbase + xs
        tw = min(TWs, Wimg - strip * TWs) + halo;               // valid strip columns are [halo, tw)
        const int q = j * rows_per_item - HF + row;             // linear strip position of this thread's row in tile 0 (>= -HF)
        y = (q + 2 * pitch) / pitch - 2;
        xs = q - y * pitch;
        px = n * crop_px + xbase;
      } else {
        px = item * rows_per_item + row;
      }
      int m = grp - t0mod;                                      // the group's first tile of this item
      if (m < 0) m += G;
      t0mod += k_tiles % G;
      if (t0mod >= G) t0mod -= G;
      if constexpr (K3) {
        for (int s = 0; s < m; ++s) {                           // at most G-1 single-tile steps
          xs += adv_x;
          y += adv_y;
          if (xs >= pitch) { xs -= pitch; ++y; }
        }
      } else {
        px += m * 128;
      }
      for (; m < k_tiles; m += G) {
        int pix = -1;
        [[maybe_unused]] int yy = 0, xx = 0;                    // image coordinates of this row's pixel (shuffling epilogues)
        if constexpr (K3) {
          // folded: rows 0 and 127 are the shuffle halo of the tile
          if (row >= HF && row <= 127 - HF && (unsigned)y < (unsigned)Himg && xs >= halo && xs < tw) pix = px + y * Wimg + xs;
          yy = y; xx = xbase + xs;
          xs += advg_x;                                         // advance to the group's next tile
          y += advg_y;
          if (xs >= pitch) { xs -= pitch; ++y; }
        } else {
          if (px < total_px32) pix = px;
          px += 128 * G;
        }
        // operands that do not depend on the accumulator are requested before waiting for it
        uint4 rsd_raw[NRES];                                    // raw 16-bit residual: converted only after the accumulator arrived
        if constexpr (!SP && ((EPI == kEpiResidual && NOUT <= 32 && !LFF) || (EPI == kEpiReluResidual && NOUT <= 48))) {
          if (pix >= 0) {
#pragma unroll
            for (int i = 0; i < NRES; ++i) rsd_raw[i] = *reinterpret_cast<const uint4*>(res + (size_t)pix * res_pitch + res_off + i * 8);
          }
        }
        LPSR_TRACE(wq == 0 && lane == 0, ii * k_tiles + m, 3, clock64());
        const int accI = grp + G * (int)esel;                   // this tile's accumulator
        const uint32_t taddr = taddr0 + esel * (uint32_t)(G * NACCW);
        ptx::mbar_wait(tfull_bar(accI), (full_bits >> esel) & 1u);
        full_bits ^= 1u << esel;
        esel ^= (uint32_t)(NACC - 1);
        ptx::tc_fence_after();
        LPSR_TRACE(wq == 0 && lane == 0, ii * k_tiles + m, 4, clock64());
        float* xb = xg + (size_t)my_par * (4 * 2 * XROW * NOUT);
        [[maybe_unused]] const uint32_t tile_par = my_par;
        my_par ^= 1u;
        const bool valid = pix >= 0 && !skip_store;
        const int pix32 = valid ? pix : -1;
        // plain accumulators are read in blocks of up to 64 columns: all TMEM loads of a block in flight at once, one wait, and after
        // the last block the accumulator goes back to the MMA warp before any math
        if constexpr (CS == 2) {
          // ---- column-split folded epilogue: this warp owns output channels [cb, cb + 8) of its 32 rows
          const int cb = half * 8;
          float lf[8], v[8], rg[8];
          ptx::tc_ld8_nowait(taddr + cb, lf);
          ptx::tc_ld8_nowait(taddr + NOUT + cb, v);
          ptx::tc_ld8_nowait(taddr + 2 * NOUT + cb, rg);
          ptx::tc_wait_ld();
          if constexpr (!LFF) {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty_bar(accI));    // accumulator is in registers
          }
          // out[q] = D[q-1, dx=0] + D[q, dx=1] + D[q+1, dx=2]; warp boundaries through the small exchange (disjoint channels per half)
          if (lane == 31) {
            *reinterpret_cast<float4*>(&xb[(wq * 2 + 0) * NOUT + cb]) = make_float4(lf[0], lf[1], lf[2], lf[3]);
            *reinterpret_cast<float4*>(&xb[(wq * 2 + 0) * NOUT + cb + 4]) = make_float4(lf[4], lf[5], lf[6], lf[7]);
          }
          if (lane == 0) {
            *reinterpret_cast<float4*>(&xb[(wq * 2 + 1) * NOUT + cb]) = make_float4(rg[0], rg[1], rg[2], rg[3]);
            *reinterpret_cast<float4*>(&xb[(wq * 2 + 1) * NOUT + cb + 4]) = make_float4(rg[4], rg[5], rg[6], rg[7]);
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            lf[c] = __shfl_up_sync(0xffffffffu, lf[c], 1);
            rg[c] = __shfl_down_sync(0xffffffffu, rg[c], 1);
          }
          ptx::bar_sync_named(1 + grp * 2 + half, 128);
          if (lane == 0 && wq > 0) {
            const float4 t0 = *reinterpret_cast<const float4*>(&xb[((wq - 1) * 2 + 0) * NOUT + cb]);
            const float4 t1 = *reinterpret_cast<const float4*>(&xb[((wq - 1) * 2 + 0) * NOUT + cb + 4]);
            lf[0] = t0.x; lf[1] = t0.y; lf[2] = t0.z; lf[3] = t0.w; lf[4] = t1.x; lf[5] = t1.y; lf[6] = t1.z; lf[7] = t1.w;
          }
          if (lane == 31 && wq < 3) {
            const float4 t0 = *reinterpret_cast<const float4*>(&xb[((wq + 1) * 2 + 1) * NOUT + cb]);
            const float4 t1 = *reinterpret_cast<const float4*>(&xb[((wq + 1) * 2 + 1) * NOUT + cb + 4]);
            rg[0] = t0.x; rg[1] = t0.y; rg[2] = t0.z; rg[3] = t0.w; rg[4] = t1.x; rg[5] = t1.y; rg[6] = t1.z; rg[7] = t1.w;
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) v[c] += lf[c] + rg[c];
          if constexpr (!LFF) {
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] += bias_r[c];
            store_chunk8<T, true>(out, out_pitch, out_off + cb, pix32, v);
          } else {
            // stage 1: this half's 8 channels of g3 = relu(conv + b3) are one plane of the K = 16 operand (bias came from the tensor core)
            *reinterpret_cast<uint4*>(a2_all + (size_t)grp * 4096 + half * 2048 + row * 16) =
                make_uint4(pack2<T, true>(v[0], v[1]), pack2<T, true>(v[2], v[3]), pack2<T, true>(v[4], v[5]), pack2<T, true>(v[6], v[7]));
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(a2full_bar(grp));
            // stage 2: this half's 16 of the 32 columns x + alpha * lff(cat[x, g0..g3]) -> block output
            ptx::mbar_wait(tfull2_bar(grp), tile_par);
            ptx::tc_fence_after();
            float o[16];
            ptx::tc_ld16_nowait(taddr + 48 + half * 16, o);
            ptx::tc_wait_ld();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty_bar(accI));
            store_chunk16<T>(out, out_pitch, out_off + half * 16, pix32, o);
          }
        } else {
        constexpr bool PRELOAD = !FOLD && !FOLD5;
        constexpr bool DUAL = (NS == 2) || SP;                        // a second column block (lo weights along N / split lo terms) is added to the first
        constexpr int PB = DUAL ? 16 : (NOUT <= 64 ? NOUT : 64);      // dual: block by block, the lo block is added right after its load
        static_assert(NOUT % PB == 0, "block size");
        [[maybe_unused]] float vall[PRELOAD ? PB : 1];
#pragma unroll
        for (int cc = 0; cc < NOUT; cc += CH) {
          float v[CH];
          if constexpr (PRELOAD) {
            if (cc % PB == 0) {
#pragma unroll
              for (int c2 = 0; c2 < PB; c2 += CH) ptx::tc_ld16_nowait(taddr + cc + c2, vall + c2);
              [[maybe_unused]] float vlo[DUAL ? PB : 1];
              if constexpr (DUAL) ptx::tc_ld16_nowait(taddr + NOUT + cc, vlo);
              ptx::tc_wait_ld();
              if constexpr (DUAL) {
#pragma unroll
                for (int c = 0; c < PB; ++c) vall[c] += vlo[c];
              }
              if (cc + PB >= NOUT) {
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(tempty_bar(accI));
                LPSR_TRACE(wq == 0 && lane == 0, ii * k_tiles + m, 5, clock64());
              }
            }
          }
          if constexpr (FOLD5) {
            // out[q] = D[q-2, dx=0] + D[q-1, dx=1] + D[q, dx=2] + D[q+1, dx=3] + D[q+2, dx=4]
            static_assert(!FOLD5 || NOUT == 16, "5-wide fold: Cout = 16");
            float d0[CH], d1[CH], d3[CH], d4[CH];
            ptx::tc_ld16_nowait(taddr, d0);
            ptx::tc_ld16_nowait(taddr + NOUT, d1);
            ptx::tc_ld16_nowait(taddr + 2 * NOUT, v);
            ptx::tc_ld16_nowait(taddr + 3 * NOUT, d3);
            ptx::tc_ld16_nowait(taddr + 4 * NOUT, d4);
            ptx::tc_wait_ld();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty_bar(accI));
            // rows that cross a warp boundary travel through shared memory: up side {D0@30, D0@31, D1@31}, down side {D4@0, D4@1, D3@0}
            float* up = xb + (size_t)(wq * 2 + 0) * (XROW * NOUT);
            float* dn = xb + (size_t)(wq * 2 + 1) * (XROW * NOUT);
            if (lane >= 30) {
#pragma unroll
              for (int c = 0; c < CH; c += 4) *reinterpret_cast<float4*>(&up[(lane - 30) * NOUT + c]) = make_float4(d0[c], d0[c + 1], d0[c + 2], d0[c + 3]);
            }
            if (lane == 31) {
#pragma unroll
              for (int c = 0; c < CH; c += 4) *reinterpret_cast<float4*>(&up[2 * NOUT + c]) = make_float4(d1[c], d1[c + 1], d1[c + 2], d1[c + 3]);
            }
            if (lane <= 1) {
#pragma unroll
              for (int c = 0; c < CH; c += 4) *reinterpret_cast<float4*>(&dn[lane * NOUT + c]) = make_float4(d4[c], d4[c + 1], d4[c + 2], d4[c + 3]);
            }
            if (lane == 0) {
#pragma unroll
              for (int c = 0; c < CH; c += 4) *reinterpret_cast<float4*>(&dn[2 * NOUT + c]) = make_float4(d3[c], d3[c + 1], d3[c + 2], d3[c + 3]);
            }
#pragma unroll
            for (int c = 0; c < CH; ++c) {
              d0[c] = __shfl_up_sync(0xffffffffu, d0[c], 2);
              d1[c] = __shfl_up_sync(0xffffffffu, d1[c], 1);
              d3[c] = __shfl_down_sync(0xffffffffu, d3[c], 1);
              d4[c] = __shfl_down_sync(0xffffffffu, d4[c], 2);
            }
            ptx::bar_sync_named(1 + grp, 128);
            if (wq > 0 && lane <= 1) {                           // from the previous warp's rows 30, 31
              const float* pu = xb + (size_t)((wq - 1) * 2 + 0) * (XROW * NOUT);
#pragma unroll
              for (int c = 0; c < CH; ++c) d0[c] = pu[lane * NOUT + c];
              if (lane == 0) {
#pragma unroll
                for (int c = 0; c < CH; ++c) d1[c] = pu[2 * NOUT + c];
              }
            }
            if (wq < 3 && lane >= 30) {                          // from the next warp's rows 0, 1
              const float* pd = xb + (size_t)((wq + 1) * 2 + 1) * (XROW * NOUT);
#pragma unroll
              for (int c = 0; c < CH; ++c) d4[c] = pd[(lane - 30) * NOUT + c];
              if (lane == 31) {
#pragma unroll
                for (int c = 0; c < CH; ++c) d3[c] = pd[2 * NOUT + c];
              }
            }
#pragma unroll
            for (int c = 0; c < CH; ++c) v[c] += (d0[c] + d1[c]) + (d3[c] + d4[c]);
          } else if constexpr (FOLD) {
            float lf[CH], rg[CH];
            if (!LPSR_DBG(64)) ptx::tc_ld16_nowait(taddr + cc, lf);
            ptx::tc_ld16_nowait(taddr + NOUT + cc, v);
            if (!LPSR_DBG(64)) ptx::tc_ld16_nowait(taddr + 2 * NOUT + cc, rg);
            if constexpr (SP) {   // + the lo-term accumulator block (same three dx sub-blocks, NMMA columns further)
              float lf2[CH], v2[CH], rg2[CH];
              ptx::tc_ld16_nowait(taddr + NMMA + cc, lf2);
              ptx::tc_ld16_nowait(taddr + NMMA + NOUT + cc, v2);
              ptx::tc_ld16_nowait(taddr + NMMA + 2 * NOUT + cc, rg2);
              ptx::tc_wait_ld();
#pragma unroll
              for (int c = 0; c < CH; ++c) { lf[c] += lf2[c]; v[c] += v2[c]; rg[c] += rg2[c]; }
            }
            ptx::tc_wait_ld();
            if (cc + CH >= NOUT && !LFF) {
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(tempty_bar(accI));  // accumulator is in registers: hand TMEM back to the MMA warp
            }
            if (cc + CH >= NOUT) LPSR_TRACE(wq == 0 && lane == 0, ii * k_tiles + m, 5, clock64());
            // out[q] = D[q-1, dx=0] + D[q, dx=1] + D[q+1, dx=2]: neighbours by warp shuffle; across warp boundaries lane 31's
            // dx=0 partial / lane 0's dx=2 partial travel through a small smem exchange
            if (lane == 31) {
#pragma unroll
              for (int c = 0; c < CH; c += 4) *reinterpret_cast<float4*>(&xb[(wq * 2 + 0) * NOUT + cc + c]) = make_float4(lf[c], lf[c + 1], lf[c + 2], lf[c + 3]);
            }
            if (lane == 0) {
#pragma unroll
              for (int c = 0; c < CH; c += 4) *reinterpret_cast<float4*>(&xb[(wq * 2 + 1) * NOUT + cc + c]) = make_float4(rg[c], rg[c + 1], rg[c + 2], rg[c + 3]);
            }
            if (!LPSR_DBG(8)) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
              lf[c] = __shfl_up_sync(0xffffffffu, lf[c], 1);
              rg[c] = __shfl_down_sync(0xffffffffu, rg[c], 1);
            }
            }
            if (!LPSR_DBG(16)) ptx::bar_sync_named(1 + grp, 128);
            if (lane == 0 && wq > 0) {
#pragma unroll
              for (int c = 0; c < CH; c += 4) {
                const float4 t = *reinterpret_cast<const float4*>(&xb[((wq - 1) * 2 + 0) * NOUT + cc + c]);
                lf[c] = t.x; lf[c + 1] = t.y; lf[c + 2] = t.z; lf[c + 3] = t.w;
              }
            }
            if (lane == 31 && wq < 3) {
#pragma unroll
              for (int c = 0; c < CH; c += 4) {
                const float4 t = *reinterpret_cast<const float4*>(&xb[((wq + 1) * 2 + 1) * NOUT + cc + c]);
                rg[c] = t.x; rg[c + 1] = t.y; rg[c + 2] = t.z; rg[c + 3] = t.w;
              }
            }
#pragma unroll
            for (int c = 0; c < CH; ++c) v[c] += lf[c] + rg[c];
          } else {
#pragma unroll
            for (int c = 0; c < CH; ++c) v[c] = vall[cc % PB + c];
          }
          if constexpr (LFF) {
            // biases and the residual were accumulated by the tensor core
          } else if constexpr (NOUT <= 32 && SP) {
#pragma unroll
            for (int c = 0; c < CH; ++c) v[c] = fmaf(v[c], acc_scale, bias_r[cc + c]);
          } else if constexpr (NOUT <= 32) {
#pragma unroll
            for (int c = 0; c < CH; ++c) v[c] += bias_r[cc + c];
          } else {
#pragma unroll
            for (int c = 0; c < CH; c += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(bias_s + cc + c);
              v[c] += b4.x; v[c + 1] += b4.y; v[c + 2] += b4.z; v[c + 3] += b4.w;
            }
          }
          if constexpr (LFF) {
            // ---- stage 1: g3 = relu(conv + bias) becomes the K = 16 operand of lff's last slice (planes of 8 channels, row = 16 B)
            {
              uint8_t* a2 = a2_all + (size_t)grp * kA2Bytes;
              uint32_t hi[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) hi[i] = pack2<T, true>(v[2 * i], v[2 * i + 1]);
              *reinterpret_cast<uint4*>(a2 + row * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(a2 + 2048 + row * 16) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
              if constexpr (SP) {   // g3's lo operand: relu(v) - hi
                uint32_t lo[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hi[i]));
                  lo[i] = pack2<T, false>(fmaxf(v[2 * i], 0.f) - f.x, fmaxf(v[2 * i + 1], 0.f) - f.y);
                }
                *reinterpret_cast<uint4*>(a2 + 4096 + row * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                *reinterpret_cast<uint4*>(a2 + 6144 + row * 16) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
              }
            }
            ptx::fence_proxy_async();                           // generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(a2full_bar(grp));
            // ---- stage 2: columns 48..79 = x + lff(cat[x, g0..g3]) + bias -> block output
            ptx::mbar_wait(tfull2_bar(grp), tile_par);
            ptx::tc_fence_after();
            float o[kLffN], ol[kLffN];                          // products with the hi and with the lo weights
#pragma unroll
            for (int hh = 0; hh < kLffN / CH; ++hh) {
              ptx::tc_ld16_nowait(taddr + 48 + hh * CH, o + hh * CH);
              ptx::tc_ld16_nowait(taddr + 48 + kLffN + hh * CH, ol + hh * CH);
            }
            ptx::tc_wait_ld();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty_bar(accI));
#pragma unroll
            for (int c = 0; c < kLffN; ++c) o[c] += ol[c];
#pragma unroll
            for (int hh = 0; hh < kLffN / CH; ++hh) {
              float oc[CH];
#pragma unroll
              for (int c = 0; c < CH; ++c) oc[c] = o[hh * CH + c];
              store_out16<T, false, SP>(out, out_pitch, out_off, hh * CH, pix32, oc);
            }
          } else if constexpr (EPI == kEpiFinalSigmoid) {
            // final conv (Cout = 1, padded to 16): channel 0 only, fp32; lanes are consecutive pixels -> coalesced 4-byte stores
            if (valid) static_cast<float*>(p.out)[pix] = __fdividef(1.f, 1.f + __expf(-v[0]));
          } else if constexpr (EPI == kEpiUnshuffleRelu) {
            // PixelUnshuffle(2) + ReLU folded into the store: 2-byte scatter, the 2x2 neighbours fill the rest of each 96-byte pixel
            static_assert(EPI != kEpiUnshuffleRelu || NOUT == 16, "12 real channels in one chunk");
            if (valid) {
              T* o = out + unshuffle2_dst(nn, yy, xx, 0, Himg, Wimg, out_pitch, out_off);   // + c*4: the address map of PixelUnshuffle(2)
#pragma unroll
              for (int c = 0; c < 12; ++c) o[c * 4] = from_f32<T>(fmaxf(v[c], 0.f));
            }
          } else if constexpr (EPI == kEpiUp2Store) {
            static_assert(EPI != kEpiUp2Store || NOUT == 128 || NOUT == 64, "four pixels of 32 or 16 channels per row");
            constexpr int CSUB = NOUT / 4;                        // channels per sub-pixel
            const int up = valid ? ((nn * 2 * Himg + 2 * yy + cc / (2 * CSUB)) * 2 * Wimg + 2 * xx + (cc / CSUB) % 2) : -1;
            if (p.relu) store_chunk16<TOUT, true>(reinterpret_cast<TOUT*>(out), out_pitch, out_off + cc % CSUB, up, v);
            else store_chunk16<TOUT>(reinterpret_cast<TOUT*>(out), out_pitch, out_off + cc % CSUB, up, v);
          } else if constexpr (EPI == kEpiGate) {
            // CSAR gates (1x1, NOUT = 32): spatial branch x_in * sigmoid(.), channel branch x_in^2 * s_c
            float g1[CH];
            if (valid) {
              float xi[CH];
              load_vec<T, CH>(static_cast<const T*>(p.aux) + (size_t)pix * p.aux_pitch + p.aux_off + cc, xi);
              const float4* sc4 = reinterpret_cast<const float4*>(p.gate + (size_t)(pix / p.px_per_crop) * NOUT + cc);
#pragma unroll
              for (int c = 0; c < CH; c += 4) {
                const float4 s4 = __ldg(sc4 + c / 4);
                const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  // 16-bit modes: ex2.approx-based logistic (the result is rounded to bf16/fp16 anyway)
                  v[c + j] = xi[c + j] * __fdividef(1.f, 1.f + __expf(-v[c + j]));
                  g1[c + j] = xi[c + j] * (xi[c + j] * sv[j]);
                }
              }
            }
            store_chunk16<T>(out, out_pitch, out_off + cc, pix32, g1);
            store_chunk16<T>(out, out_pitch, p.out_off2 + cc, pix32, v);
          } else {

            if constexpr (EPI == kEpiReluResidual) {
              if (valid) {
                if constexpr (NOUT <= 48) {
                  const T* re = reinterpret_cast<const T*>(rsd_raw) + cc;
#pragma unroll
                  for (int c = 0; c < CH; ++c) v[c] = fmaxf(v[c], 0.f) + to_f32<T>(re[c]);
                } else {
                  float r[CH];
                  load_vec<T, CH>(res + (size_t)pix * res_pitch + res_off + cc, r);
#pragma unroll
                  for (int c = 0; c < CH; ++c) v[c] = fmaxf(v[c], 0.f) + r[c];
                }
              }
            }
            if constexpr (EPI == kEpiResidual && SP) {
              if (valid) {   // residual as a split tensor: [16 hi | 16 lo] per chunk
                const uint4* rp = reinterpret_cast<const uint4*>(res + (size_t)pix * res_pitch + res_off + 2 * cc);
                const uint4 h0 = rp[0], h1 = rp[1], l0 = rp[2], l1 = rp[3];
                const __half2* hh2[2] = {reinterpret_cast<const __half2*>(&h0), reinterpret_cast<const __half2*>(&h1)};
                const __half2* ll2[2] = {reinterpret_cast<const __half2*>(&l0), reinterpret_cast<const __half2*>(&l1)};
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float2 a = __half22float2(hh2[q][i]), b = __half22float2(ll2[q][i]);
                    v[q * 8 + 2 * i] += a.x + b.x;
                    v[q * 8 + 2 * i + 1] += a.y + b.y;
                  }
              }
            }
            if constexpr (EPI == kEpiResidual && !SP) {
              if (valid) {
                float r[CH];
                if constexpr (NOUT > 32) {
                  load_vec<T, CH>(res + (size_t)pix * res_pitch + res_off + cc, r);
                } else {
                  const T* re = reinterpret_cast<const T*>(rsd_raw) + cc;
#pragma unroll
                  for (int c = 0; c < CH; ++c) r[c] = to_f32<T>(re[c]);
                }
#pragma unroll
                for (int c = 0; c < CH; ++c) v[c] += r[c];
              }
            }
            if constexpr (EPI == kEpiPool) {
              float r[CH];
#pragma unroll
              for (int c = 0; c < CH; ++c) r[c] = pix >= 0 ? v[c] : 0.f;
              const float tot = warp_column_sums16(r, lane);
              if (!(lane & 1)) p.pool[((size_t)(item * k_tiles + m) * 4 + wq) * NOUT + cc + ((lane >> 1) & 15)] = tot;
            }
            // all 32 lanes take part in the staged, coalesced store (rows without a pixel are skipped inside)
            if (!LPSR_DBG(32)) store_out16<T, EPI == kEpiRelu, SP>(out, out_pitch, out_off, cc, pix32, v);
            else if (v[0] == 123.456f) out[0] = from_f32<T>(v[1] + v[5] + v[9] + v[13]);
          }
        }
        }   // CS == 1
        LPSR_TRACE(wq == 0 && lane == 0, ii * k_tiles + m, 6, clock64());
      }
    }
  }

  // ---- teardown -----------------------------------------------------------------------------------
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp0) ptx::tmem_dealloc(tmem_base, kTmemCols);   // the warp that allocated
}

// ---------------------------------------------------------------------------------------------------
// host launch
// ---------------------------------------------------------------------------------------------------
struct UmmaPlan {
  UmmaParams p;
  UmmaTmaps tm;
  size_t smem_bytes;
  int grid;
};

typedef CUresult (*PFN_lpsr_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PFN_lpsr_tmapEncodeTiled umma_encode_fn() {
  static PFN_lpsr_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_lpsr_tmapEncodeTiled>(p);
  }
  return fn;
}

// tensor map over an NHWC activation buffer (pitch channels per pixel), box = [rows][px][ch] (4-D) or [px][ch] (2-D)
inline const char* umma_make_tmap(CUtensorMap* out, const void* base, bool fp16, int pitch_ch, int ch_box, bool k3, int B, int H, int W,
                                  int box_px, int box_rows, long long total_px) {
  PFN_lpsr_tmapEncodeTiled enc = umma_encode_fn();
  if (!enc) return "cuTensorMapEncodeTiled entry point not found";
  const CUtensorMapSwizzle sw = ch_box == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : ch_box == 32 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : ch_box == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  const CUtensorMapDataType dt = fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r;
  if (k3) {
    const cuuint64_t gdim[4] = {(cuuint64_t)pitch_ch, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t gstr[3] = {(cuuint64_t)pitch_ch * 2, (cuuint64_t)W * pitch_ch * 2, (cuuint64_t)H * W * pitch_ch * 2};
    const cuuint32_t box[4] = {(cuuint32_t)ch_box, (cuuint32_t)box_px, (cuuint32_t)box_rows, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    r = enc(out, dt, 4, const_cast<void*>(base), gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    const cuuint64_t gdim[2] = {(cuuint64_t)pitch_ch, (cuuint64_t)total_px};
    const cuuint64_t gstr[1] = {(cuuint64_t)pitch_ch * 2};
    const cuuint32_t box[2] = {(cuuint32_t)ch_box, (cuuint32_t)box_px};
    const cuuint32_t es[2] = {1, 1};
    r = enc(out, dt, 2, const_cast<void*>(base), gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  return r == CUDA_SUCCESS ? nullptr : "cuTensorMapEncodeTiled failed";
}

inline long long*& umma_trace_buffer() { static long long* b = nullptr; return b; }

// batch_invariant: the tiling of a crop must not depend on the batch size (the pooled partial sums are added tile by tile)
inline const char* umma_plan(UmmaPlan& plan, const UmmaWeights& w, const ConvParams& cp, int num_sms, bool fp16, bool fp32_out = false,
                             bool batch_invariant = false) {
  UmmaParams& p = plan.p;
  p = UmmaParams{};
  memset(&plan.tm, 0, sizeof plan.tm);
  const bool c7 = (w.ks == 7);
  const bool k3 = (w.ks == 3) || (w.ks == 5) || c7, fold = umma_fold(w.ks, w.cout);
  const bool lff = w.fused_lff;
  const bool fold5 = fold && w.ks == 5;
  const int N = w.cout, NMMA = lff ? kLffCols : (fold ? w.ks * N : (w.nsplit ? 2 * N : N)), ntap = fold ? w.ks : ((w.ks == 3 || w.ks == 5) ? w.ks * w.ks : 1);
  const int halo = w.ks / 2;
  const size_t xch_bytes = fold ? (size_t)kEpiGroups * 2 * 4 * 2 * (fold5 ? 3 : 1) * N * 4 : 0;
  p.halo = halo;
  p.n_ks = w.cin / 16;                                       // 7x7: 28 pixel-pair steps (UmmaWeights::cin = 448 pseudo channels)
  p.n_ks_real = w.split ? w.n_real : (w.wsplit ? p.n_ks / 2 : p.n_ks);
  p.wlo_step = w.wlo_step ? w.wlo_step : (w.split ? 2 : 1);
  p.acc_scale = w.split ? w.acc_scale : 1.f;
  if (p.n_ks > kUmmaMaxSteps) return "too many K-steps";
  if (w.wsplit && !w.split && (lff || c7)) return "hi + lo weights are not supported by this mode";
  if (w.nsplit && (w.ks != 3 || fold || lff || w.wsplit || N != 32)) return "N-split weights: per-tap 3x3 with Cout = 32 only";
  if (!c7 && cp.n_chunks != p.n_ks_real) return "chunk table does not match Cin/16";
  if (!fp32_out && (cp.out_pitch % 16 || cp.out_off % 16 || reinterpret_cast<uintptr_t>(cp.out) % 32)) return "output pitch/offset not 32-byte aligned (256-bit stores)";
  if (cp.res && (cp.res_pitch % 8 || cp.res_off % 8)) return "residual pitch/offset not 16-byte aligned";
  if (kEpiGroups * NMMA * (w.split ? 2 : 1) > 512) return "N too large for the TMEM accumulators";
  // ---- K-chunks: merge runs of 16-channel slices that are contiguous in the SAME tensor into TMA boxes of 64 / 32 / 16 ch
  const void* chunk_base[kUmmaMaxKChunks];
  int chunk_pitch[kUmmaMaxKChunks];
  p.n_chunks = 0;
  auto base_of = [&](int k) { return cp.chunk_ptr[k] ? cp.chunk_ptr[k] : cp.in; };
  auto pitch_of = [&](int k) { return cp.chunk_ptr[k] ? cp.chunk_pitch[k] : cp.in_pitch; };
  if (c7) {   // one box of 8 (3 real + 5 zero) channels per pixel
    if (cp.in_pitch % 8 || reinterpret_cast<uintptr_t>(cp.in) % 16) return "7x7 input must be a tensor padded to a multiple of 8 channels";
    p.n_chunks = 1;
    p.chunk_ch[0] = 8; p.chunk_coff[0] = 0; p.chunk_map[0] = 0;
    chunk_base[0] = cp.in; chunk_pitch[0] = cp.in_pitch;
  }
  for (int k = 0; !c7 && k < p.n_ks_real;) {
    int run = 1;
    while (k + run < p.n_ks_real && base_of(k + run) == base_of(k) && cp.chunk_off[k + run] == cp.chunk_off[k] + 16 * run) ++run;
    int off = cp.chunk_off[k];
    if (off % 8 || pitch_of(k) % 8) return "chunk offset/pitch not 16-byte aligned";
    if (reinterpret_cast<uintptr_t>(base_of(k)) % 16) return "input base not 16-byte aligned";
    for (int left = run; left > 0;) {
      const int take = left >= 4 ? 4 : left >= 2 ? 2 : 1;      // 64, 32 or 16 channels
      if (p.n_chunks == kUmmaMaxKChunks) return "too many K-chunks";
      p.chunk_ch[p.n_chunks] = 16 * take;
      p.chunk_coff[p.n_chunks] = off;
      p.chunk_map[p.n_chunks] = p.n_chunks;
      chunk_base[p.n_chunks] = base_of(k);
      chunk_pitch[p.n_chunks] = pitch_of(k);
      ++p.n_chunks;
      off += 16 * take;
      left -= take;
    }
    k += run;
  }
  p.w = w.w; p.bias = w.bias;
  p.out = cp.out; p.out_pitch = cp.out_pitch; p.out_off = cp.out_off;
  p.res = cp.res; p.res_pitch = cp.res_pitch; p.res_off = cp.res_off;
  p.B = cp.B; p.H = cp.H; p.W = cp.W; p.relu = cp.relu;
  p.total_px = (long long)cp.B * cp.H * cp.W;
  if (p.total_px >= (1LL << 31)) return "batch too large for 32-bit pixel indices (split the batch)";
  p.mode = kEpiPlain;
  p.px_per_crop = cp.H * cp.W;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("LPSR_UMMA_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.debug = dbg;
    p.trace = nullptr;
#ifdef LPSR_UMMA_TRACE_BUILD
    static long long* trace = nullptr;
    static int tr_on = -1;
    if (tr_on < 0) {
      tr_on = getenv("LPSR_UMMA_TRACE") ? 1 : 0;
      if (tr_on) cudaMalloc(&trace, 512 * 8 * sizeof(long long));
    }
    if (tr_on) { cudaMemsetAsync(trace, 0, 512 * 8 * sizeof(long long)); p.trace = trace; umma_trace_buffer() = trace; }
#endif
  }
  const size_t w_bytes = ((lff ? (size_t)(w.cin / 8) * kLffWCols * 16 + 4 * kLffN * 16 : (size_t)ntap * w.cin * NMMA * 2) + 127) & ~(size_t)127;
  const size_t a2_bytes = lff ? (size_t)kEpiGroups * (w.split ? 8192 : 4096) + 2 * 128 * 16 + 2 * kLffCols * 16 + 4 * 32 * 16   // g3 operands, ones, biases, identity
                              : (N > 32 ? (size_t)N * 4 : 0);                                                   // wide layers: bias vector
  const size_t fixed = w_bytes + (2 * kUmmaMaxBufs + 6 * kEpiGroups + 2) * 8 + xch_bytes + a2_bytes + (kUmmaMaxSteps * 16 + 128) /*steps, slot_base*/ + 1024 /*alignment slack*/ + 256;
  const size_t smem_cap = 227 * 1024 - fixed;
  auto item_buf_bytes = [&](size_t npx) {                      // every chunk is 1024-aligned inside the buffer
    size_t b = 0;
    for (int c = 0; c < p.n_chunks; ++c) b += (npx * p.chunk_ch[c] * 2 + 1023) & ~(size_t)1023;
    return b;
  };
  if (k3) {
    const int ts = fold5 ? 124 : (fold ? 126 : 128);
    p.tstride = ts;
    // choose strip width TW (equalised over W) and tiles per item k by a cost model:
    //   MMA/epilogue work ~ computed rows per output pixel; staging traffic ~ staged slots per output pixel
    // Several item loads must be in flight while one item is consumed (TMA latency ~ several thousand clocks): prefer plans
    // whose item buffer fits `want_bufs` times; relax only if nothing fits.
    static int want_bufs_env = -1;
    if (want_bufs_env < 0) { const char* e = getenv("LPSR_UMMA_BUFS"); want_bufs_env = e ? atoi(e) : 2; }
    double best_cost = 1e30;
    int best_k = 0, best_ns = 0;
    for (int want_bufs = std::min(std::max(want_bufs_env, 2), kUmmaMaxBufs); want_bufs >= 2 && !best_k; --want_bufs)
    for (int ns = 1; ns <= std::max(1, (cp.W + 15) / 16); ++ns) {
      const int TW = (cp.W + ns - 1) / ns, pitch = TW + 2 * halo;
      if (pitch > 256) continue;                               // TMA box dimension limit
      if (ns > 1 && TW < 24) break;
      const long long lin = (long long)cp.H * pitch;
      for (int k = 1; k <= kUmmaMaxK; ++k) {
        const int rbox = (k * ts + (2 * halo + 1) * pitch + 2 * halo - 1 + pitch - 1) / pitch;   // rows covering any item's tap footprint
        if (rbox > 256) break;
        const size_t npx = (size_t)rbox * pitch;
        if (item_buf_bytes(npx) * want_bufs > smem_cap) break;
        const long long items_strip = (lin + (long long)ts * k - 1) / ((long long)ts * k);
        const long long items = items_strip * ns * cp.B;
        const double work = (double)items_strip * k * 128 / (double)(cp.H * TW);
        const double stage = (double)items_strip * (double)npx / (double)(cp.H * TW);
        const long long waves = (items + num_sms - 1) / num_sms;
        const double fill = batch_invariant ? 1.0 : (double)(waves * num_sms) / (double)items;
        const double cost = fill * (0.75 * work + 0.25 * stage);
        if (cost < best_cost - 1e-9) { best_cost = cost; best_k = k; best_ns = ns; }
      }
    }
    if (!best_k) return "tile does not fit in shared memory";
    p.k = best_k;
    p.n_strips = best_ns;
    p.TW = (cp.W + best_ns - 1) / best_ns;
    p.pitch = p.TW + 2 * halo;
    const long long lin = (long long)cp.H * p.pitch;
    p.items_per_strip = (int)((lin + (long long)ts * p.k - 1) / ((long long)ts * p.k));
    p.n_items = p.items_per_strip * p.n_strips * cp.B;
    p.rbox = (p.k * ts + (2 * halo + 1) * p.pitch + 2 * halo - 1 + p.pitch - 1) / p.pitch;
    p.npx = p.rbox * p.pitch;
  } else {
    p.tstride = 128;
    int best_k = 1;
    double best_cost = 1e30;
    for (int k = 1; k <= kUmmaMaxK; ++k) {
      if (item_buf_bytes((size_t)128 * k) * 4 > smem_cap && k > 1) break;
      const long long items = (p.total_px + 128LL * k - 1) / (128LL * k);
      const long long waves = (items + num_sms - 1) / num_sms;
      const double cost = (double)(waves * num_sms) * k * 128 / (double)p.total_px + 0.04 / k;
      if (cost < best_cost - 1e-9) { best_cost = cost; best_k = k; }
    }
    p.k = best_k;
    p.TW = p.pitch = p.n_strips = p.items_per_strip = p.rbox = 0;
    p.n_items = (int)((p.total_px + 128LL * p.k - 1) / (128LL * p.k));
    p.npx = 128 * p.k;
  }
  size_t off = 0;
  for (int c = 0; c < p.n_chunks; ++c) {
    p.chunk_smem[c] = (uint32_t)off;
    off += ((size_t)p.npx * p.chunk_ch[c] * 2 + 1023) & ~(size_t)1023;
  }
  p.buf_bytes = (uint32_t)off;
  int bufs = (int)(smem_cap / off);
  if (bufs < 2) return "tile does not fit in shared memory";
  if (bufs > kUmmaMaxBufs) bufs = kUmmaMaxBufs;
  p.n_bufs = bufs;
  plan.smem_bytes = (size_t)bufs * off + w_bytes + (2 * bufs + 6 * kEpiGroups + 2) * 8 + xch_bytes + a2_bytes + (kUmmaMaxSteps * 16 + 128) + 1024 + 64;
  plan.grid = std::min(p.n_items, num_sms);
  // ---- tensor maps, one per K-chunk
  for (int c = 0; c < p.n_chunks; ++c)
    if (const char* msg = umma_make_tmap(&plan.tm.m[c], chunk_base[c], fp16, chunk_pitch[c], p.chunk_ch[c], k3, cp.B, cp.H, cp.W,
                                         k3 ? p.pitch : 128, p.rbox, p.total_px))
      return msg;
  return nullptr;
}

// stream-ordered launch that may overlap the previous kernel's tail (the kernels call griddepcontrol.wait before reading activations)
inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("LPSR_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// cudaFuncSetAttribute is per device (context): one flag per kernel instantiation AND device ordinal.  The flags are only ever set
// to true after the attribute call succeeded, so a race between two host threads costs at most a redundant call.
constexpr int kMaxDevices = 64;
inline bool* func_configured_flag(bool (&flags)[kMaxDevices]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;   // unknown device: configure every time
  return &flags[dev];
}

inline bool umma_column_split() {   // LPSR_EPI_SPLIT=0: one epilogue warp per lane quadrant everywhere (the round-1 layout)
  static int v = -1;
  if (v < 0) { const char* e = getenv("LPSR_EPI_SPLIT"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

template <typename T, int N, int MODE, int EPI, typename TOUT = T, int CS = 1, int NS = 1, bool SP = false>
inline const char* umma_launch_inst(const UmmaPlan& plan, cudaStream_t st) {
  static bool configured[kMaxDevices] = {};
  bool* flag = func_configured_flag(configured);
  if (!flag || !*flag) {
    cudaError_t e = cudaFuncSetAttribute(umma_conv_kernel<T, N, MODE, EPI, TOUT, CS, NS, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    if (flag) *flag = true;
  }
  cudaError_t e = launch_pdl(umma_conv_kernel<T, N, MODE, EPI, TOUT, CS, NS, SP>, dim3(plan.grid), dim3(umma_threads(CS)), plan.smem_bytes, st, plan.p, plan.tm);
  if (e == cudaSuccess) e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

// epilogue selection: plain / ReLU / residual (mutually exclusive in this network) or the special modes
template <typename T, int N, int MODE>
inline const char* umma_launch_epi(const UmmaPlan& plan, cudaStream_t st, bool nsplit = false) {
  const UmmaParams& p = plan.p;
  if constexpr (N == 32 && MODE == kConv3x3Taps) {
    if (nsplit) {   // hi + lo weights along N (the CSAR conv_in pair)
      if (p.mode == kEpiPlain && p.relu && !p.res) return umma_launch_inst<T, N, MODE, kEpiRelu, T, 1, 2>(plan, st);
      if (p.mode == kEpiPlain && !p.relu && !p.res) return umma_launch_inst<T, N, MODE, kEpiPlain, T, 1, 2>(plan, st);
      if (p.mode == kEpiPool) return umma_launch_inst<T, N, MODE, kEpiPool, T, 1, 2>(plan, st);
      return "N-split weights: epilogue not instantiated";
    }
  }
  if (nsplit) return "N-split weights: shape not instantiated";
  if (p.mode == kEpiPlain && p.relu && p.res) return "ReLU + residual epilogue is not instantiated";
  if (p.mode == kEpiPlain && p.res) return umma_launch_inst<T, N, MODE, kEpiResidual>(plan, st);
  if constexpr (N == 16 && MODE == kConv3x3Fold) {
    if (p.mode == kEpiPlain && p.relu && umma_column_split()) return umma_launch_inst<T, N, MODE, kEpiRelu, T, 2>(plan, st);
  }
  if (p.mode == kEpiPlain && p.relu) return umma_launch_inst<T, N, MODE, kEpiRelu>(plan, st);
  if (p.mode == kEpiPlain) return umma_launch_inst<T, N, MODE, kEpiPlain>(plan, st);
  if constexpr (N == 32 && MODE == kConv1x1) {
    if (p.mode == kEpiGate) return umma_launch_inst<T, N, MODE, kEpiGate>(plan, st);
  }
  if constexpr (N == 32 && MODE == kConv3x3Taps) {
    if (p.mode == kEpiPool) return umma_launch_inst<T, N, MODE, kEpiPool>(plan, st);
  }
  if constexpr (N == 16 && MODE == kConv3x3Fold) {
    if (p.mode == kEpiFinalSigmoid) return umma_launch_inst<T, N, MODE, kEpiFinalSigmoid>(plan, st);
  }
  return "epilogue mode not instantiated for this shape";
}

// split (double-fp16) tensors: the dense trunk layers of the fp32-accuracy mode
template <typename T>
inline const char* umma_plan_launch_split(const UmmaPlan& plan, const UmmaWeights& w, cudaStream_t st) {
  if constexpr (IsBf16<T>::value) {
    return "split tensors are fp16 pairs";
  } else {
    const UmmaParams& p = plan.p;
    const bool plain = p.mode == kEpiPlain;
    if (w.fused_lff) return "the fused layer + lff kernel is not used with split tensors (its weights would not fit shared memory)";
    if (w.ks == 3 && w.cout == 16) {
      if (p.mode == kEpiFinalSigmoid) return umma_launch_inst<T, 16, kConv3x3Fold, kEpiFinalSigmoid, T, 1, 1, true>(plan, st);
      if (plain && p.relu && !p.res) return umma_launch_inst<T, 16, kConv3x3Fold, kEpiRelu, T, 1, 1, true>(plan, st);
    }
    if (w.ks == 3 && w.cout == 32) {
      if (p.mode == kEpiPool) return umma_launch_inst<T, 32, kConv3x3Taps, kEpiPool, T, 1, 1, true>(plan, st);
      if (plain && p.res && !p.relu) return umma_launch_inst<T, 32, kConv3x3Taps, kEpiResidual, T, 1, 1, true>(plan, st);
      if (plain && p.relu && !p.res) return umma_launch_inst<T, 32, kConv3x3Taps, kEpiRelu, T, 1, 1, true>(plan, st);
      if (plain && !p.relu && !p.res) return umma_launch_inst<T, 32, kConv3x3Taps, kEpiPlain, T, 1, 1, true>(plan, st);
    }
    if (w.ks == 7 && w.cout == 32 && plain && !p.relu && !p.res) return umma_launch_inst<T, 32, kConv7x7, kEpiPlain, T, 1, 1, true>(plan, st);
    if (w.ks == 1 && w.cout == 32 && plain && !p.relu && !p.res) return umma_launch_inst<T, 32, kConv1x1, kEpiPlain, T, 1, 1, true>(plan, st);
    if (w.ks == 1 && w.cout == 32 && plain && !p.relu && p.res) return umma_launch_inst<T, 32, kConv1x1, kEpiResidual, T, 1, 1, true>(plan, st);
    return "split tensors: shape/epilogue not instantiated";
  }
}

template <typename T, typename TOUT = T>
inline const char* umma_plan_launch(const UmmaPlan& plan, const UmmaWeights& w, cudaStream_t st) {
  if (w.split) return umma_plan_launch_split<T>(plan, w, st);
  const int mode = plan.p.mode;
  const bool plain = (mode == kEpiPlain && !plan.p.res);
  if (w.ks == 7) return w.cout == 32 ? umma_launch_epi<T, 32, kConv7x7>(plan, st) : "unsupported Cout";
  if (w.ks == 5) {   // AutoEncoder stages (composed depthwise + pointwise)
    if (w.cout == 16 && mode == kEpiUnshuffleRelu) return umma_launch_inst<T, 16, kConv5x5Fold, kEpiUnshuffleRelu>(plan, st);
    if (w.cout == 48 && mode == kEpiPlain && plan.p.relu && plan.p.res) return umma_launch_inst<T, 48, kConv5x5Taps, kEpiReluResidual>(plan, st);
    if (w.cout == 64 && mode == kEpiUp2Store) return umma_launch_inst<T, 64, kConv5x5Taps, kEpiUp2Store, T>(plan, st);
    if (w.cout == 128 && mode == kEpiUp2Store) return umma_launch_inst<T, 128, kConv5x5Taps, kEpiUp2Store, TOUT>(plan, st);
    return "5x5 conv: shape/epilogue not instantiated";
  }
  if (w.ks == 3) {
    if (w.fused_lff) {
      if (!(mode == kEpiPlain && plan.p.res)) return "fused dense layer + lff needs the residual";
      return umma_launch_inst<T, 16, kConv3x3FoldLff, kEpiResidual>(plan, st);
    }
    if (w.cout == 16) return umma_launch_epi<T, 16, kConv3x3Fold>(plan, st);
    if (w.cout == 32) return umma_launch_epi<T, 32, kConv3x3Taps>(plan, st, w.nsplit);
    if (w.cout == 48 && plain) return plan.p.relu ? umma_launch_inst<T, 48, kConv3x3Taps, kEpiRelu>(plan, st)
                                                  : umma_launch_inst<T, 48, kConv3x3Taps, kEpiPlain>(plan, st);
  } else {
    if (w.cout == 16) return umma_launch_epi<T, 16, kConv1x1>(plan, st);
    if (w.cout == 32) return umma_launch_epi<T, 32, kConv1x1>(plan, st);
    if (w.cout == 64) return umma_launch_epi<T, 64, kConv1x1>(plan, st);
  }
  return "unsupported Cout";
}

// special epilogue request: CSAR gate (1x1, Cout = 32) or final sigmoid (see UmmaParams::mode)
struct UmmaGate {
  const void* x_in; int xin_pitch, xin_off;
  const float* s_c;        // [B][32] channel gates
  int out_off_spatial;     // channel offset of x_in * sigmoid(.) in the output buffer (x_in^2 * s_c goes to ConvParams::out_off)
  int final_sigmoid;       // 1: kEpiFinalSigmoid instead (ConvParams::out is a float [B*H*W] tensor)
  int epi = 0;             // kEpiUnshuffleRelu / kEpiUp2Store / kEpiPool (else 0)
  float* pool = nullptr;   // kEpiPool: partial-sum buffer, pool_slots_per_crop x 32 floats per crop
  int pool_slots_per_crop = 0;          // capacity of `pool` per crop
  int* out_slots_per_crop = nullptr;    // kEpiPool: receives the slots the launch wrote per crop (tiles per crop x 4)
};

template <typename T, typename TOUT = T>
inline const char* umma_conv_launch(const UmmaWeights& w, const ConvParams& cp, int num_sms, cudaStream_t st, const UmmaGate* gate = nullptr) {
  UmmaPlan plan;
  if (const char* msg = umma_plan(plan, w, cp, num_sms, !IsBf16<T>::value, gate && gate->final_sigmoid, gate && gate->epi == kEpiPool)) return msg;
  if (gate && gate->epi) {
    plan.p.mode = gate->epi;
    if (gate->epi == kEpiPool) {
      const int slots = plan.p.n_strips * plan.p.items_per_strip * plan.p.k * 4;
      if (w.ks != 3 || w.cout != 32 || !gate->pool || !gate->out_slots_per_crop) return "pool epilogue needs a 3x3 conv with Cout = 32 and a partial buffer";
      if (slots > gate->pool_slots_per_crop) return "pool partial buffer too small";
      plan.p.pool = gate->pool;
      *gate->out_slots_per_crop = slots;
    }
  } else if (gate && gate->final_sigmoid) {
    if (w.ks != 3 || w.cout != 16) return "final epilogue needs the folded 3x3 conv with Cout padded to 16";
    plan.p.mode = kEpiFinalSigmoid;
  } else if (gate) {
    if (w.ks != 1 || w.cout != 32) return "gate epilogue needs a 1x1 conv with Cout = 32";
    if (gate->xin_pitch % 8 || gate->xin_off % 8 || gate->out_off_spatial % 8) return "gate operands not 16-byte aligned";
    plan.p.mode = kEpiGate;
    plan.p.aux = gate->x_in; plan.p.aux_pitch = gate->xin_pitch; plan.p.aux_off = gate->xin_off;
    plan.p.gate = gate->s_c;
    plan.p.out_off2 = gate->out_off_spatial;
  }
  return umma_plan_launch<T, TOUT>(plan, w, st);
}

template <> inline const char* umma_conv_launch<float, float>(const UmmaWeights&, const ConvParams&, int, cudaStream_t, const UmmaGate*) {
  return "tensor-core path is 16-bit only";
}

}  // namespace lpsr
