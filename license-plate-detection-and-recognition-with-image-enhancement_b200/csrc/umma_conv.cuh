// umma_conv.cuh -- tcgen05 / TMEM implicit-GEMM convolution (3x3 and 1x1, stride 1, zero 'same' padding)
// for the 16-bit modes.  sm_100a only.
//
// GEMM view: M = output pixels, K = Cin per tap row, N = Cout (1x1) or 3*Cout (3x3: the three dx taps
// are FOLDED into N).  A (pixels x channels) is never materialised as im2col.  Each CTA stages a HALOED, linearised
// pixel range of the NHWC input once per 16-channel K-slice into shared memory in a "channel-group planar" layout
//
//        smem_A[cg (2 per K-slice)][pixel slot][8 channels = 16 bytes]
//
// which is exactly the tcgen05 no-swizzle K-major canonical layout with SBO = 128 B (8 rows x 16 B): row (pixel) r
// lives at start + r*16 B, linearly, so a row shift is just a different descriptor START ADDRESS.
//   * the dy taps are three MMAs whose A start is shifted by dy*pitch slots (same staged data, read 3x not 9x);
//   * the dx taps are folded into N: D[q, dx*Cout+co] = sum_{dy,ci} X[q+(dy-1)*pitch, ci] * W[dy][dx][ci][co], and the
//     epilogue forms out[q] = D[q-1, 0] + D[q, 1] + D[q+1, 2] with warp shuffles (row == TMEM lane == thread).
//     Consecutive M-tiles overlap by 2 rows (stride 126) so only warp boundaries need a tiny smem exchange.
//   This cuts the tensor-core smem reads 3x and the number of MMAs 3x versus one MMA per tap (N=16 MMAs are
//   issue- and smem-bound: measured 120 clk each in the per-tap version, profiles/r1_ncu_v1_summary.md).
//
// Work decomposition: the image is cut into column strips of TW pixels; inside a strip pixels are linearised with
// pitch = TW+2 (left/right halo columns, zero filled at the image border); a work item is k tiles of 126 consecutive
// linear positions of one strip of one crop (1x1 convs: k*128 consecutive pixels of the whole batch).  Persistent
// CTAs (one per SM) loop over items.  An item's input (all K-slices, haloed) is staged as ONE shared-memory buffer
// (ring of 2-4 buffers), so 50-130 KB of loads are in flight per SM, and the MMA loop runs tile-outer / K-inner:
// only two tile accumulators (2 x N fp32 TMEM columns) ping-pong between the MMA warp and two epilogue groups.
// Warp roles (352 threads):
//     warps 0-3, 4-7  two epilogue groups (group g owns TMEM accumulator g; TMEM lane quadrant = warp id % 4):
//                     tcgen05.ld, dx shifted sum, bias/ReLU/residual (or the CSAR gate), 16-byte stores
//     warp  8         MMA issuer: tcgen05.mma (M=128, N, K=16) from one elected lane, tcgen05.commit to mbarriers
//     warps 9-10      loaders: cp.async (zero-fill = padding) into the item ring, proxy fence, mbarrier arrive
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "umma_weights.h"

namespace lpsr {

constexpr int kUmmaThreads = 352;       // 11 warps: 2 x 4 epilogue, 1 MMA, 2 loaders (<= 184 registers per thread)
constexpr int kUmmaLoaderThreads = 64;
constexpr int kUmmaMmaWarp = 8;
constexpr int kUmmaFirstLoaderWarp = 9;
constexpr int kUmmaMaxK = 16;           // max M-tiles per item
constexpr int kUmmaMaxBufs = 4;         // item buffers in the shared-memory ring

struct UmmaParams {
  const void* in; int in_pitch;
  int n_ks;                         // K-slices of 16 channels
  int chunk_off[kMaxChunks];        // physical channel offset of each K-slice
  const uint16_t* w; const float* bias;
  void* out; int out_pitch, out_off;
  const void* res; int res_pitch, res_off;
  int B, H, W, relu;
  int k;                            // M-tiles (128 rows) per item
  int tstride;                      // valid rows per tile: 126 (3x3, tiles overlap by the 2 shuffle-halo rows) or 128
  int TW, pitch, n_strips, items_per_strip, n_items;
  int npx;                          // pixel slots per staged item
  int n_bufs;                       // item buffers in the ring
  long long total_px;               // B*H*W (1x1 mode)
  // CSAR gate epilogue (mode 1, 1x1 Cout=32 only): v = sigmoid(acc + bias); out[off2 + c] = x_in[c] * v  (spatial branch)
  // and out[out_off + c] = x_in[c]^2 * s_c[n][c] (channel branch, reference lpsr.py:133-135,182-184)
  int mode;
  const void* aux; int aux_pitch, aux_off;   // x_in
  const float* gate;                         // s_c [B][32]
  int out_off2;
  int px_per_crop;
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
// Bounded wait: a protocol bug must become a trap (reported as a CUDA error), never a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) {
    if (spins > (1u << 26)) {
      printf("umma_conv: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
constexpr uint32_t kUmmaDescHi = (128u >> 4) | (1u << 14);

template <typename T> struct IsBf16 { static constexpr bool value = false; };
template <> struct IsBf16<__nv_bfloat16> { static constexpr bool value = true; };

enum { kEpiPlain = 0, kEpiGate = 1 };

// ---------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------
// MODE: kConv1x1 | kConv3x3Taps (one MMA per tap, N = Cout: used for Cout >= 32 where the MMA is already efficient)
//       | kConv3x3Fold (dx folded into N = 3*Cout: used for Cout = 16 where per-tap MMAs would be issue/smem bound)
enum { kConv1x1 = 0, kConv3x3Taps = 1, kConv3x3Fold = 2 };

template <typename T, int NOUT, int MODE>
__global__ void __launch_bounds__(kUmmaThreads, 1) umma_conv_kernel(const UmmaParams p) {
  static_assert(sizeof(T) == 2, "16-bit operands");
  constexpr bool FOLD = (MODE == kConv3x3Fold);
  constexpr bool K3 = (MODE != kConv1x1);
  constexpr int NMMA = FOLD ? 3 * NOUT : NOUT;                 // GEMM-N of one MMA = TMEM columns per tile
  constexpr int NTAP = (MODE == kConv1x1) ? 1 : (FOLD ? 3 : 9);   // MMAs per K-slice
  constexpr uint32_t kTmemCols = (2 * NMMA <= 32) ? 32 : (2 * NMMA <= 64) ? 64 : (2 * NMMA <= 128) ? 128 : (2 * NMMA <= 256) ? 256 : 512;
  extern __shared__ __align__(128) uint8_t smem[];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int CG = p.n_ks * 2;                                  // 8-channel groups of the whole K extent
  const uint32_t w_bytes = (uint32_t)NTAP * CG * NMMA * 16;
  const uint32_t buf_bytes = (uint32_t)p.npx * CG * 16;        // [cg][slot][16 B]
  uint8_t* w_smem = smem;
  uint8_t* a_smem = smem + ((w_bytes + 127) & ~127u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_smem + (size_t)p.n_bufs * buf_bytes);
  // bars[0..R) full, [R..2R) empty, [2R..2R+2) tmem_full, [2R+2..2R+4) tmem_empty, then the TMEM base address,
  // then the warp-boundary exchange buffers of the folded epilogue: [2 groups][2 parities][4 warps][2 sides][NOUT] floats
  const int R = p.n_bufs;
  const uint32_t bar0 = ptx::smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (R + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * R + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * R + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * R + 4);
  float* xchg = reinterpret_cast<float*>(bars + 2 * R + 6);

  // ---- one-time setup ------------------------------------------------------------------------------
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.w);
    uint4* dst = reinterpret_cast<uint4*>(w_smem);
    for (uint32_t i = threadIdx.x; i < w_bytes / 16; i += kUmmaThreads) dst[i] = __ldg(src + i);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < R; ++s) {
      ptx::mbar_init(full_bar(s), kUmmaLoaderThreads);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);
      ptx::mbar_init(tempty_bar(a), 128);
    }
    ptx::fence_mbar_init();
  }
  if (warp == kUmmaMmaWarp) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async();      // weights were written with st.shared: make them visible to the tensor core proxy
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_my_items = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int rows_per_item = p.k * p.tstride;

  if (warp >= kUmmaFirstLoaderWarp) {
    // =================================== loaders ===================================================
    const int lt = threadIdx.x - kUmmaFirstLoaderWarp * 32;     // 0..63
    const T* in = static_cast<const T*>(p.in);
    int pending = 0;                                            // committed cp.async groups not yet signalled
    for (int ii = 0; ii < n_my_items; ++ii) {
      const int item = blockIdx.x + ii * gridDim.x;
      const int buf = ii % R;
      const uint32_t ph = (uint32_t)(ii / R) & 1u;
      if (!ptx::mbar_try_wait(empty_bar(buf), ph ^ 1u)) {
        // about to block on the consumer: publish everything already requested first
        ptx::cp_async_wait<0>();
        ptx::fence_proxy_async();
        for (; pending > 0; --pending) ptx::mbar_arrive(full_bar((ii - pending) % R));
        ptx::mbar_wait(empty_bar(buf), ph ^ 1u);
      }
      long long base_px = 0;   // 1x1: first pixel of the item
      int n = 0, x0 = 0, qlo = 0;
      if constexpr (K3) {
        const int per_crop = p.n_strips * p.items_per_strip;
        n = item / per_crop;
        const int rem = item % per_crop;
        const int strip = rem / p.items_per_strip, j = rem % p.items_per_strip;
        x0 = strip * p.TW;
        qlo = j * rows_per_item - 1 - p.pitch;                  // linear strip position held by slot 0
      } else {
        base_px = (long long)item * rows_per_item;
      }
      const uint32_t dst0 = ptx::smem_u32(a_smem + (size_t)buf * buf_bytes);
      for (int i = lt; i < p.npx; i += kUmmaLoaderThreads) {
        const T* src = in;
        uint32_t nbytes = 0;
        if constexpr (K3) {
          const int qq = qlo + i + 2 * p.pitch;                // >= 0
          const int y = qq / p.pitch - 2, xs = qq % p.pitch;
          const int x = x0 + xs - 1;
          if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
            src = in + ((size_t)(n * p.H + y) * p.W + x) * p.in_pitch;
            nbytes = 16;
          }
        } else {
          const long long px = base_px + i;
          if (px < p.total_px) {
            src = in + (size_t)px * p.in_pitch;
            nbytes = 16;
          }
        }
        uint32_t dst = dst0 + (uint32_t)i * 16;
        for (int ks = 0; ks < p.n_ks; ++ks) {
          const T* s2 = nbytes ? src + p.chunk_off[ks] : src;
          ptx::cp_async_16(dst, s2, nbytes);
          ptx::cp_async_16(dst + (uint32_t)p.npx * 16, nbytes ? s2 + 8 : s2, nbytes);
          dst += (uint32_t)p.npx * 32;
        }
      }
      ptx::cp_async_commit();
      ++pending;
      if (pending > 1) {                                        // keep one item in flight, publish the older one
        ptx::cp_async_wait<1>();
        ptx::fence_proxy_async();
        ptx::mbar_arrive(full_bar((ii + 1 - pending) % R));
        --pending;
      }
    }
    ptx::cp_async_wait<0>();
    ptx::fence_proxy_async();
    for (; pending > 0; --pending) ptx::mbar_arrive(full_bar((n_my_items - pending) % R));
  } else if (warp == kUmmaMmaWarp) {
    // =================================== MMA issuer ================================================
    // The whole warp runs the (uniform) control flow and waits; one elected lane issues MMAs and commits.
    const bool leader = ptx::elect_one();
    constexpr uint32_t idesc = umma_idesc_f16(IsBf16<T>::value, NMMA);
    const uint32_t w_lo = umma_desc_lo(ptx::smem_u32(w_smem), (uint32_t)NMMA * 16);
    const uint32_t a_lbo = (uint32_t)p.npx * 16;
    uint32_t tile_ctr = 0;
    for (int ii = 0; ii < n_my_items; ++ii) {
      const int buf = ii % R;
      const uint32_t ph = (uint32_t)(ii / R) & 1u;
      ptx::mbar_wait(full_bar(buf), ph);
      ptx::tc_fence_after();
      const uint32_t a_lo_item = umma_desc_lo(ptx::smem_u32(a_smem + (size_t)buf * buf_bytes), a_lbo);
      for (int m = 0; m < p.k; ++m, ++tile_ctr) {
        const uint32_t acc = tile_ctr & 1u;
        ptx::mbar_wait(tempty_bar(acc), ((tile_ctr >> 1) & 1u) ^ 1u);   // the epilogue group drained this accumulator
        ptx::tc_fence_after();
        if (leader) {
          const uint32_t d = tmem_base + acc * NMMA;
          uint32_t a_lo = a_lo_item + (uint32_t)(m * p.tstride);          // slot units == 16-byte units
          uint32_t b_lo = w_lo;
#pragma unroll 1
          for (int ks = 0; ks < p.n_ks; ++ks) {
#pragma unroll
            for (int t = 0; t < NTAP; ++t) {
              // tap -> slot shift: folded: dy*pitch (dx lives in N); per-tap: dy*pitch + dx
              const uint32_t shift = FOLD ? (uint32_t)(t * p.pitch) : (uint32_t)((t / 3) * p.pitch + (t % 3));
              ptx::tc_mma_f16_lohi(d, a_lo + shift, kUmmaDescHi, b_lo + (uint32_t)(t * CG * NMMA), kUmmaDescHi, idesc, (uint32_t)(ks | t));
            }
            a_lo += (uint32_t)p.npx * 2;                                  // next K-slice: 2 channel-group planes
            b_lo += 2 * NMMA;
          }
          ptx::tc_commit(tfull_bar(acc));                                 // this tile's accumulator is complete
          if (m == p.k - 1) ptx::tc_commit(empty_bar(buf));               // item buffer reusable once all MMAs retire
        }
        __syncwarp();
      }
    }
  } else {
    // =================================== epilogue (two groups) ======================================
    const int grp = warp >> 2, wq = warp & 3;                   // group == TMEM accumulator, wq == TMEM lane quadrant
    const int row = wq * 32 + lane;                             // accumulator row
    constexpr int CH = NOUT < 32 ? NOUT : 32;                   // output channels handled per pass (bounds registers)
    float bias[CH];
    if constexpr (NOUT <= 32) {
#pragma unroll
      for (int c = 0; c < CH; ++c) bias[c] = __ldg(p.bias + c);
    }
    T* out = static_cast<T*>(p.out);
    const T* res = static_cast<const T*>(p.res);
    float* xg = xchg + (size_t)grp * (2 * 4 * 2 * NOUT);
    const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(grp * NMMA);
    uint32_t tile_ctr = 0, my_ctr = 0;
    for (int ii = 0; ii < n_my_items; ++ii) {
      const int item = blockIdx.x + ii * gridDim.x;
      int n = 0, x0 = 0, q0 = 0, tw = 0;
      long long base_px = 0;
      if constexpr (K3) {
        const int per_crop = p.n_strips * p.items_per_strip;
        n = item / per_crop;
        const int rem = item % per_crop;
        const int strip = rem / p.items_per_strip, j = rem % p.items_per_strip;
        x0 = strip * p.TW;
        tw = min(p.TW, p.W - x0);
        q0 = j * rows_per_item - (FOLD ? 1 : 0);                // linear position of accumulator row 0 of tile 0
      } else {
        base_px = (long long)item * rows_per_item;
      }
      for (int m = 0; m < p.k; ++m, ++tile_ctr) {
        if ((int)(tile_ctr & 1u) != grp) continue;
        // pixel of this row
        long long pix = -1;
        if constexpr (K3) {
          const int q = q0 + m * p.tstride + row;               // folded: rows 0 and 127 are the shuffle halo of the tile
          if (!FOLD || (row >= 1 && row <= 126)) {
            const int y = q / p.pitch, xs = q - y * p.pitch;
            if (y < p.H && xs >= 1 && xs <= tw) pix = (long long)(n * p.H + y) * p.W + (x0 + xs - 1);
          }
        } else {
          const long long px = base_px + m * 128 + row;
          if (px < p.total_px) pix = px;
        }
        // operands that do not depend on the accumulator are requested before waiting for it
        float rsd[CH];
        const bool has_res = (p.mode == kEpiPlain && res != nullptr && pix >= 0);
        if constexpr (NOUT <= 32) {
          if (has_res) load_vec<T, CH>(res + (size_t)pix * p.res_pitch + p.res_off, rsd);
        }
        ptx::mbar_wait(tfull_bar(grp), my_ctr & 1u);
        ptx::tc_fence_after();
        float* xb = xg + (size_t)(my_ctr & 1u) * (4 * 2 * NOUT);
        ++my_ctr;
#pragma unroll
        for (int cc = 0; cc < NOUT; cc += CH) {
          float v[CH];
          if constexpr (FOLD) {
            static_assert(!FOLD || NOUT <= 32, "folded epilogue keeps the whole tile row in registers");
            float lf[CH], rg[CH];
#pragma unroll
            for (int c0 = 0; c0 < CH; c0 += 16) {
              ptx::tc_ld16_nowait(taddr + c0, &lf[c0]);
              ptx::tc_ld16_nowait(taddr + NOUT + c0, &v[c0]);
              ptx::tc_ld16_nowait(taddr + 2 * NOUT + c0, &rg[c0]);
            }
            ptx::tc_wait_ld();
            ptx::tc_fence_before();
            ptx::mbar_arrive(tempty_bar(grp));                  // accumulator is in registers: hand TMEM back to the MMA warp
            // out[q] = D[q-1, dx=0] + D[q, dx=1] + D[q+1, dx=2]: neighbours by warp shuffle; across warp boundaries lane 31's
            // dx=0 partial / lane 0's dx=2 partial travel through a small smem exchange
            if (lane == 31) {
#pragma unroll
              for (int c = 0; c < CH; c += 4) *reinterpret_cast<float4*>(&xb[(wq * 2 + 0) * NOUT + c]) = make_float4(lf[c], lf[c + 1], lf[c + 2], lf[c + 3]);
            }
            if (lane == 0) {
#pragma unroll
              for (int c = 0; c < CH; c += 4) *reinterpret_cast<float4*>(&xb[(wq * 2 + 1) * NOUT + c]) = make_float4(rg[c], rg[c + 1], rg[c + 2], rg[c + 3]);
            }
#pragma unroll
            for (int c = 0; c < CH; ++c) {
              lf[c] = __shfl_up_sync(0xffffffffu, lf[c], 1);
              rg[c] = __shfl_down_sync(0xffffffffu, rg[c], 1);
            }
            ptx::bar_sync_named(1 + grp, 128);
            if (lane == 0 && wq > 0) {
#pragma unroll
              for (int c = 0; c < CH; c += 4) {
                const float4 t = *reinterpret_cast<const float4*>(&xb[((wq - 1) * 2 + 0) * NOUT + c]);
                lf[c] = t.x; lf[c + 1] = t.y; lf[c + 2] = t.z; lf[c + 3] = t.w;
              }
            }
            if (lane == 31 && wq < 3) {
#pragma unroll
              for (int c = 0; c < CH; c += 4) {
                const float4 t = *reinterpret_cast<const float4*>(&xb[((wq + 1) * 2 + 1) * NOUT + c]);
                rg[c] = t.x; rg[c + 1] = t.y; rg[c + 2] = t.z; rg[c + 3] = t.w;
              }
            }
#pragma unroll
            for (int c = 0; c < CH; ++c) v[c] += lf[c] + rg[c];
          } else {
#pragma unroll
            for (int c0 = 0; c0 < CH; c0 += 16) ptx::tc_ld16_nowait(taddr + cc + c0, &v[c0]);
            ptx::tc_wait_ld();
            if (cc + CH >= NOUT) {
              ptx::tc_fence_before();
              ptx::mbar_arrive(tempty_bar(grp));
            }
          }
          if (pix >= 0) {
            if constexpr (NOUT <= 32) {
#pragma unroll
              for (int c = 0; c < CH; ++c) v[c] += bias[c];
            } else {
#pragma unroll
              for (int c = 0; c < CH; ++c) v[c] += __ldg(p.bias + cc + c);
            }
            if (p.mode == kEpiGate) {
              // CSAR gates (1x1, NOUT = 32): spatial branch x_in * sigmoid(.), channel branch x_in^2 * s_c
              if constexpr (NOUT == 32 && MODE == kConv1x1) {
                float xi[NOUT], g1[NOUT];
                load_vec<T, NOUT>(static_cast<const T*>(p.aux) + (size_t)pix * p.aux_pitch + p.aux_off, xi);
                const float* sc = p.gate + (size_t)(pix / p.px_per_crop) * NOUT;
#pragma unroll
                for (int c = 0; c < NOUT; ++c) {
                  v[c] = xi[c] * sigmoid_f32(v[c]);
                  g1[c] = xi[c] * (xi[c] * __ldg(sc + c));
                }
                store_vec<T, NOUT>(out + (size_t)pix * p.out_pitch + p.out_off, g1);
                store_vec<T, NOUT>(out + (size_t)pix * p.out_pitch + p.out_off2, v);
              }
            } else {
              if (p.relu) {
#pragma unroll
                for (int c = 0; c < CH; ++c) v[c] = fmaxf(v[c], 0.f);
              }
              if (has_res) {
                if constexpr (NOUT > 32) load_vec<T, CH>(res + (size_t)pix * p.res_pitch + p.res_off + cc, rsd);
#pragma unroll
                for (int c = 0; c < CH; ++c) v[c] += rsd[c];
              }
              store_vec<T, CH>(out + (size_t)pix * p.out_pitch + p.out_off + cc, v);
            }
          }
        }
      }
    }
  }

  // ---- teardown -----------------------------------------------------------------------------------
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kUmmaMmaWarp) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------------------------------
// host launch
// ---------------------------------------------------------------------------------------------------
struct UmmaPlan {
  UmmaParams p;
  size_t smem_bytes;
  int grid;
};

inline const char* umma_plan(UmmaPlan& plan, const UmmaWeights& w, const ConvParams& cp, int num_sms) {
  UmmaParams& p = plan.p;
  p = UmmaParams{};
  const bool k3 = (w.ks == 3), fold = umma_fold(w.ks, w.cout);
  const int N = w.cout, NMMA = fold ? 3 * N : N, ndy = fold ? 3 : (k3 ? 9 : 1);
  p.in = cp.in; p.in_pitch = cp.in_pitch;
  p.n_ks = w.cin / 16;
  if (cp.n_chunks != p.n_ks) return "chunk table does not match Cin/16";
  for (int k = 0; k < p.n_ks; ++k) p.chunk_off[k] = cp.chunk_off[k];
  if (cp.in_pitch % 8 || cp.out_pitch % 8 || cp.out_off % 8 || (cp.res && (cp.res_pitch % 8 || cp.res_off % 8))) return "pitch/offset not 16-byte aligned";
  for (int k = 0; k < p.n_ks; ++k) if (p.chunk_off[k] % 8) return "chunk offset not 16-byte aligned";
  if (2 * NMMA > 512) return "N too large for two TMEM accumulators";
  p.w = w.w; p.bias = w.bias;
  p.out = cp.out; p.out_pitch = cp.out_pitch; p.out_off = cp.out_off;
  p.res = cp.res; p.res_pitch = cp.res_pitch; p.res_off = cp.res_off;
  p.B = cp.B; p.H = cp.H; p.W = cp.W; p.relu = cp.relu;
  p.total_px = (long long)cp.B * cp.H * cp.W;
  p.mode = kEpiPlain;
  p.px_per_crop = cp.H * cp.W;
  const size_t w_bytes = ((size_t)ndy * w.cin * NMMA * 2 + 127) & ~(size_t)127;
  const size_t fixed = w_bytes + (2 * kUmmaMaxBufs + 6) * 8 + 2 * 2 * 4 * 2 * N * 4 + 256;
  const size_t smem_cap = 227 * 1024 - fixed;
  const size_t slot_bytes = (size_t)w.cin * 2;                 // all K-slices of one pixel slot
  if (k3) {
    const int ts = fold ? 126 : 128;
    p.tstride = ts;
    // choose strip width TW (equalised over W) and tiles per item k by a cost model:
    //   MMA/epilogue work ~ computed rows per output pixel; staging traffic ~ staged slots per output pixel;
    //   at least 2 item buffers must fit; prefer >= 48 KB per buffer in flight
    double best_cost = 1e30;
    int best_k = 0, best_ns = 0;
    for (int ns = 1; ns <= std::max(1, (cp.W + 15) / 16); ++ns) {
      const int TW = (cp.W + ns - 1) / ns, pitch = TW + 2;
      if (TW > 254) continue;
      if (ns > 1 && TW < 24) break;
      const long long lin = (long long)cp.H * pitch;
      for (int k = 1; k <= kUmmaMaxK; ++k) {
        const size_t npx = (size_t)((ts * k + 2 * pitch + 2 + 7) & ~7);
        if (npx * slot_bytes * 2 > smem_cap) break;
        if (npx * 16 >= (1u << 18)) break;                     // LBO field: 14 bits of 16-byte units
        const long long items_strip = (lin + (long long)ts * k - 1) / ((long long)ts * k);
        const long long items = items_strip * ns * cp.B;
        const double work = (double)items_strip * k * 128 / (double)(cp.H * TW);
        const double stage = (double)items_strip * (double)npx / (double)(cp.H * TW);
        const long long waves = (items + num_sms - 1) / num_sms;
        const double fill = (double)(waves * num_sms) / (double)items;
        const double cost = fill * (0.6 * work + 0.4 * stage);
        if (cost < best_cost - 1e-9) { best_cost = cost; best_k = k; best_ns = ns; }
      }
    }
    if (!best_k) return "tile does not fit in shared memory";
    p.k = best_k;
    p.n_strips = best_ns;
    p.TW = (cp.W + best_ns - 1) / best_ns;
    p.pitch = p.TW + 2;
    const long long lin = (long long)cp.H * p.pitch;
    p.items_per_strip = (int)((lin + (long long)ts * p.k - 1) / ((long long)ts * p.k));
    p.n_items = p.items_per_strip * p.n_strips * cp.B;
    p.npx = (ts * p.k + 2 * p.pitch + 2 + 7) & ~7;
  } else {
    p.tstride = 128;
    int best_k = 1;
    double best_cost = 1e30;
    for (int k = 1; k <= kUmmaMaxK; ++k) {
      if ((size_t)128 * k * slot_bytes * 2 > smem_cap) break;
      const long long items = (p.total_px + 128LL * k - 1) / (128LL * k);
      const long long waves = (items + num_sms - 1) / num_sms;
      const double cost = (double)(waves * num_sms) * k * 128 / (double)p.total_px + 0.04 / k;
      if (cost < best_cost - 1e-9) { best_cost = cost; best_k = k; }
    }
    p.k = best_k;
    p.TW = p.pitch = p.n_strips = p.items_per_strip = 0;
    p.n_items = (int)((p.total_px + 128LL * p.k - 1) / (128LL * p.k));
    p.npx = 128 * p.k;
  }
  const size_t bb = (size_t)p.npx * slot_bytes;
  int bufs = (int)(smem_cap / bb);
  if (bufs < 2) return "tile does not fit in shared memory";
  if (bufs > kUmmaMaxBufs) bufs = kUmmaMaxBufs;
  p.n_bufs = bufs;
  plan.smem_bytes = w_bytes + (size_t)bufs * bb + (2 * bufs + 6) * 8 + 2 * 2 * 4 * 2 * N * 4 + 16;
  plan.grid = std::min(p.n_items, num_sms);
  return nullptr;
}

template <typename T, int N, int MODE>
inline const char* umma_launch_inst(const UmmaPlan& plan, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(umma_conv_kernel<T, N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    configured = true;
  }
  umma_conv_kernel<T, N, MODE><<<plan.grid, kUmmaThreads, plan.smem_bytes, st>>>(plan.p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

template <typename T>
inline const char* umma_plan_launch(const UmmaPlan& plan, const UmmaWeights& w, cudaStream_t st) {
  if (w.ks == 3) {
    if (w.cout == 16) return umma_launch_inst<T, 16, kConv3x3Fold>(plan, st);
    if (w.cout == 32) return umma_launch_inst<T, 32, kConv3x3Taps>(plan, st);
    if (w.cout == 64) return umma_launch_inst<T, 64, kConv3x3Taps>(plan, st);
  } else {
    if (w.cout == 16) return umma_launch_inst<T, 16, kConv1x1>(plan, st);
    if (w.cout == 32) return umma_launch_inst<T, 32, kConv1x1>(plan, st);
    if (w.cout == 64) return umma_launch_inst<T, 64, kConv1x1>(plan, st);
  }
  return "unsupported Cout";
}

template <typename T>
inline const char* umma_conv_launch(const UmmaWeights& w, const ConvParams& cp, int num_sms, cudaStream_t st) {
  UmmaPlan plan;
  if (const char* msg = umma_plan(plan, w, cp, num_sms)) return msg;
  return umma_plan_launch<T>(plan, w, st);
}

template <> inline const char* umma_conv_launch<float>(const UmmaWeights&, const ConvParams&, int, cudaStream_t) {
  return "tensor-core path is 16-bit only";
}

}  // namespace lpsr
