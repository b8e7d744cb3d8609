// umma_conv.cuh -- tcgen05 / TMEM implicit-GEMM convolution (3x3 and 1x1, stride 1, zero 'same' padding)
// for the 16-bit modes.  sm_100a only.
//
// GEMM view: M = output pixels, N = Cout (16/32/64), K = taps * Cin.  A (pixels x channels) is
// never materialised as im2col.  Instead each CTA stages a HALOED pixel range of the NHWC input once per
// 16-channel K-slice into shared memory in a "channel-group planar" layout
//
//        smem_A[cg (2 per K-slice)][pixel slot][8 channels = 16 bytes]
//
// which is exactly the tcgen05 no-swizzle K-major canonical layout with SBO = 128 B (8 rows x 16 B), so a row
// (pixel) r lives at start + r*16 B, linearly.  A 3x3 tap (dy,dx) is then just a different START ADDRESS of the
// same staged data: slot + dy*pitch + dx.  One staged K-slice feeds 9 taps x k M-tiles of MMAs, so shared
// memory is written once and read 9 times, and L2->smem traffic is ~1.3x the input instead of 9x.
//
// Work decomposition: the image is cut into column strips of TW pixels; inside a strip pixels are linearised
// with pitch = TW+2 (left/right halo columns, zero filled at the image border) and a work item is k*128
// consecutive linear positions of one strip of one crop (for 1x1 convs: k*128 consecutive pixels of the whole
// batch).  Persistent CTAs (one per SM) loop over items.  Warp roles:
//     warps 0-3  epilogue   : tcgen05.ld accumulators (TMEM lane quadrant = warp id), bias/ReLU/residual, store
//     warp  4    MMA issuer : one elected thread issues tcgen05.mma (M=128, N=Cout, K=16), commits to mbarriers
//     warps 5-8  loaders    : cp.async (zero-fill for padding) into the stage ring, proxy fence, mbarrier arrive
// Accumulators: k M-tiles x N fp32 columns in TMEM, double buffered across items so the epilogue of item i
// overlaps the MMAs of item i+1.
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

constexpr int kUmmaThreads = 288;       // 9 warps
constexpr int kUmmaLoaderThreads = 128;
constexpr int kUmmaMaxK = 8;            // max M-tiles per item
constexpr int kUmmaLookahead = 2;       // cp.async groups in flight per loader thread

struct UmmaParams {
  const void* in; int in_pitch;
  int n_ks;                         // K-slices of 16 channels
  int chunk_off[kMaxChunks];        // physical channel offset of each K-slice
  const uint16_t* w; const float* bias;
  void* out; int out_pitch, out_off;
  const void* res; int res_pitch, res_off;
  int B, H, W, relu;
  int k;                            // M-tiles (128 rows) per item
  int TW, pitch, n_strips, items_per_strip, n_items;
  int npx;                          // pixel slots per staged K-slice
  int n_stages;
  long long total_px;               // B*H*W (1x1 mode)
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

}  // namespace ptx

// shared-memory matrix descriptor, no swizzle, K-major: core matrix = 8 rows x 16 bytes (contiguous 128 B);
// SBO = byte stride between 8-row groups (M/N direction), LBO = byte stride between core matrices along K.
__device__ __forceinline__ uint64_t umma_desc_noswizzle(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (sm_100)
  return d;                 // base_offset = 0, layout_type = SWIZZLE_NONE (0)
}

// instruction descriptor, kind::f16: fp32 accumulate, A and B K-major, M=128
__host__ __device__ constexpr uint32_t umma_idesc_f16(bool bf16, int N) {
  return (1u << 4)                        // c_format = F32
         | ((bf16 ? 1u : 0u) << 7)        // a_format
         | ((bf16 ? 1u : 0u) << 10)       // b_format
         | (0u << 15) | (0u << 16)        // a_major, b_major = K
         | ((uint32_t)(N >> 3) << 17)     // n_dim
         | ((uint32_t)(128 >> 4) << 24);  // m_dim
}

template <typename T> struct IsBf16 { static constexpr bool value = false; };
template <> struct IsBf16<__nv_bfloat16> { static constexpr bool value = true; };

// ---------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------
template <typename T, int N, int TAPS>
__global__ void __launch_bounds__(kUmmaThreads, 1) umma_conv_kernel(const UmmaParams p) {
  static_assert(sizeof(T) == 2, "16-bit operands");
  constexpr int KS = (TAPS == 9) ? 3 : 1;
  extern __shared__ __align__(128) uint8_t smem[];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int CG = p.n_ks * 2;                                  // 8-channel groups of the whole K extent
  const uint32_t w_bytes = (uint32_t)TAPS * CG * N * 16;
  const uint32_t stage_bytes = (uint32_t)p.npx * 32;           // 2 channel groups x npx slots x 16 B
  uint8_t* w_smem = smem;
  uint8_t* a_smem = smem + ((w_bytes + 127) & ~127u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_smem + (size_t)p.n_stages * stage_bytes);
  // bars[0..S) full, [S..2S) empty, [2S..2S+2) tmem_full, [2S+2..2S+4) tmem_empty, then the TMEM base address
  const int S = p.n_stages;
  const uint32_t bar0 = ptx::smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * S + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

  const int acc_cols = p.k * N;                                // fp32 columns per accumulator stage
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < 2 * acc_cols) tmem_cols <<= 1;

  // ---- one-time setup ------------------------------------------------------------------------------
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.w);
    uint4* dst = reinterpret_cast<uint4*>(w_smem);
    for (uint32_t i = threadIdx.x; i < w_bytes / 16; i += kUmmaThreads) dst[i] = __ldg(src + i);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(full_bar(s), kUmmaLoaderThreads);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);
      ptx::mbar_init(tempty_bar(a), 128);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 4) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async();      // weights were written with st.shared: make them visible to the tensor core proxy
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_my_items = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp >= 5) {
    // =================================== loaders ===================================================
    const int lt = threadIdx.x - 5 * 32;                       // 0..127
    const T* in = static_cast<const T*>(p.in);
    int it = 0;                                                // running stage-iteration counter
    for (int ii = 0; ii < n_my_items; ++ii) {
      const int item = blockIdx.x + ii * gridDim.x;
      long long base_px = 0;   // 1x1: first pixel of the item
      int n = 0, x0 = 0, qlo = 0;
      if constexpr (TAPS == 9) {
        const int per_crop = p.n_strips * p.items_per_strip;
        n = item / per_crop;
        const int rem = item % per_crop;
        const int strip = rem / p.items_per_strip, j = rem % p.items_per_strip;
        x0 = strip * p.TW;
        qlo = j * p.k * 128 - p.pitch - 1;                      // linear strip position of slot 0
      } else {
        base_px = (long long)item * p.k * 128;
      }
      for (int ks = 0; ks < p.n_ks; ++ks, ++it) {
        const int s = it % S;
        const uint32_t ph = (uint32_t)(it / S) & 1u;
        ptx::mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t dst0 = ptx::smem_u32(a_smem + (size_t)s * stage_bytes);
        const int coff = p.chunk_off[ks];
        for (int i = lt; i < p.npx; i += kUmmaLoaderThreads) {
          const T* src = in;
          uint32_t nbytes = 0;
          if constexpr (TAPS == 9) {
            const int qq = qlo + i + 2 * p.pitch;              // >= 0
            const int y = qq / p.pitch - 2, xs = qq % p.pitch;
            const int x = x0 + xs - 1;
            if (y >= 0 && y < p.H && x >= 0 && x < p.W && xs <= p.TW + 1) {
              src = in + ((size_t)(n * p.H + y) * p.W + x) * p.in_pitch + coff;
              nbytes = 16;
            }
          } else {
            const long long px = base_px + i;
            if (px < p.total_px) {
              src = in + (size_t)px * p.in_pitch + coff;
              nbytes = 16;
            }
          }
          ptx::cp_async_16(dst0 + (uint32_t)i * 16, src, nbytes);
          ptx::cp_async_16(dst0 + (uint32_t)(p.npx + i) * 16, src + (nbytes ? 8 : 0), nbytes);
        }
        ptx::cp_async_commit();
        if (it >= kUmmaLookahead) {
          ptx::cp_async_wait<kUmmaLookahead>();
          ptx::fence_proxy_async();
          ptx::mbar_arrive(full_bar((it - kUmmaLookahead) % S));
        }
      }
    }
    // drain the last groups
    ptx::cp_async_wait<0>();
    ptx::fence_proxy_async();
    for (int d = (it < kUmmaLookahead ? it : kUmmaLookahead); d > 0; --d) ptx::mbar_arrive(full_bar((it - d) % S));
  } else if (warp == 4) {
    // =================================== MMA issuer ================================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(IsBf16<T>::value, N);
      const uint32_t a_lbo = (uint32_t)p.npx * 16, b_lbo = (uint32_t)N * 16;
      const uint32_t w_addr = ptx::smem_u32(w_smem);
      int it = 0;
      for (int ii = 0; ii < n_my_items; ++ii) {
        const int a = ii & 1;
        const uint32_t aph = (uint32_t)(ii >> 1) & 1u;
        ptx::mbar_wait(tempty_bar(a), aph ^ 1u);               // epilogue has drained this accumulator stage
        ptx::tc_fence_after();
        const uint32_t d_base = tmem_base + (uint32_t)(a * acc_cols);
        for (int ks = 0; ks < p.n_ks; ++ks, ++it) {
          const int s = it % S;
          const uint32_t ph = (uint32_t)(it / S) & 1u;
          ptx::mbar_wait(full_bar(s), ph);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(a_smem + (size_t)s * stage_bytes);
#pragma unroll 1
          for (int t = 0; t < TAPS; ++t) {
            const int dy = t / KS, dx = t % KS;                 // slot offset of this tap (halo origin = slot 0)
            const uint32_t tap_slot = (TAPS == 9) ? (uint32_t)(dy * p.pitch + dx) : 0u;
            const uint64_t b_desc = umma_desc_noswizzle(w_addr + (uint32_t)((t * CG + 2 * ks) * N) * 16, b_lbo, 128);
#pragma unroll 1
            for (int m = 0; m < p.k; ++m) {
              const uint64_t a_desc = umma_desc_noswizzle(a_addr + (tap_slot + (uint32_t)m * 128) * 16, a_lbo, 128);
              ptx::tc_mma_f16(d_base + (uint32_t)(m * N), a_desc, b_desc, idesc, (ks | t) != 0);
            }
          }
          ptx::tc_commit(empty_bar(s));                         // smem stage reusable once these MMAs retire
        }
        ptx::tc_commit(tfull_bar(a));                           // accumulators of this item complete
      }
    }
    __syncwarp();
  } else {
    // =================================== epilogue ==================================================
    const int row = warp * 32 + lane;                           // TMEM lane == accumulator row
    float bias[N];
#pragma unroll
    for (int c = 0; c < N; ++c) bias[c] = __ldg(p.bias + c);
    T* out = static_cast<T*>(p.out);
    const T* res = static_cast<const T*>(p.res);
    for (int ii = 0; ii < n_my_items; ++ii) {
      const int item = blockIdx.x + ii * gridDim.x;
      const int a = ii & 1;
      const uint32_t aph = (uint32_t)(ii >> 1) & 1u;
      int n = 0, x0 = 0, q0 = 0, tw = 0;
      long long base_px = 0;
      if constexpr (TAPS == 9) {
        const int per_crop = p.n_strips * p.items_per_strip;
        n = item / per_crop;
        const int rem = item % per_crop;
        const int strip = rem / p.items_per_strip, j = rem % p.items_per_strip;
        x0 = strip * p.TW;
        tw = min(p.TW, p.W - x0);
        q0 = j * p.k * 128;
      } else {
        base_px = (long long)item * p.k * 128;
      }
      ptx::mbar_wait(tfull_bar(a), aph);
      ptx::tc_fence_after();
      for (int m = 0; m < p.k; ++m) {
        // pixel of this row
        long long pix = -1;
        if constexpr (TAPS == 9) {
          const int q = q0 + m * 128 + row;
          const int y = q / p.pitch, xs = q % p.pitch;
          if (y < p.H && xs >= 1 && xs <= tw) pix = (long long)(n * p.H + y) * p.W + (x0 + xs - 1);
        } else {
          const long long px = base_px + m * 128 + row;
          if (px < p.total_px) pix = px;
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * acc_cols + m * N);
#pragma unroll
        for (int c0 = 0; c0 < N; c0 += 16) {
          float v[16];
          ptx::tc_ld16(taddr + c0, v);                          // warp-collective: every lane executes it
          if (pix >= 0) {
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] += bias[c0 + c];
            if (p.relu) {
#pragma unroll
              for (int c = 0; c < 16; ++c) v[c] = fmaxf(v[c], 0.f);
            }
            if (res) {
              float r[16];
              load_vec<T, 16>(res + (size_t)pix * p.res_pitch + p.res_off + c0, r);
#pragma unroll
              for (int c = 0; c < 16; ++c) v[c] += r[c];
            }
            store_vec<T, 16>(out + (size_t)pix * p.out_pitch + p.out_off + c0, v);
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(tempty_bar(a));
    }
  }

  // ---- teardown -----------------------------------------------------------------------------------
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) ptx::tmem_dealloc(tmem_base, tmem_cols);
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
  const int N = w.cout, taps = w.ks * w.ks;
  p.in = cp.in; p.in_pitch = cp.in_pitch;
  p.n_ks = w.cin / 16;
  if (cp.n_chunks != p.n_ks) return "chunk table does not match Cin/16";
  for (int k = 0; k < p.n_ks; ++k) p.chunk_off[k] = cp.chunk_off[k];
  if (cp.in_pitch % 8 || cp.out_pitch % 8 || cp.out_off % 8 || (cp.res && (cp.res_pitch % 8 || cp.res_off % 8))) return "pitch/offset not 16-byte aligned";
  for (int k = 0; k < p.n_ks; ++k) if (p.chunk_off[k] % 8) return "chunk offset not 16-byte aligned";
  p.w = w.w; p.bias = w.bias;
  p.out = cp.out; p.out_pitch = cp.out_pitch; p.out_off = cp.out_off;
  p.res = cp.res; p.res_pitch = cp.res_pitch; p.res_off = cp.res_off;
  p.B = cp.B; p.H = cp.H; p.W = cp.W; p.relu = cp.relu;
  p.total_px = (long long)cp.B * cp.H * cp.W;
  const int kmax_tmem = std::min(kUmmaMaxK, 256 / N);          // 2 accumulator stages x k x N <= 512 columns
  const size_t w_bytes = ((size_t)taps * w.cin * N * 2 + 127) & ~(size_t)127;
  const size_t smem_cap = 227 * 1024 - 1024 - w_bytes;
  auto stage_bytes = [&](int k, int pitch) { return (size_t)(taps == 9 ? 128 * k + 2 * pitch + 2 + 6 : 128 * k) * 32; };
  if (taps == 9) {
    // strip width: <= 96 columns, equalised over the image width
    const int n_strips = (cp.W + 95) / 96;
    p.n_strips = n_strips;
    p.TW = (cp.W + n_strips - 1) / n_strips;
    p.pitch = p.TW + 2;
    const long long lin = (long long)cp.H * p.pitch;           // linear positions per strip
    // pick k: fewest wasted rows, then enough items to fill the machine, subject to TMEM and >= 3 smem stages
    int best_k = 1;
    double best_cost = 1e30;
    for (int k = 1; k <= kmax_tmem; ++k) {
      if (stage_bytes(k, p.pitch) * 3 > smem_cap) break;
      const long long items_strip = (lin + 128LL * k - 1) / (128LL * k);
      const long long items = items_strip * n_strips * cp.B;
      const double work = (double)items_strip * k * 128 / (double)lin;            // >= 1: padded MMA work
      const double halo = 1.0 + (2.0 * p.pitch + 2) / (128.0 * k);                // staged pixels per output pixel
      const long long waves = (items + num_sms - 1) / num_sms;
      const double fill = (double)(waves * num_sms) / (double)items;              // >= 1: tail-wave inefficiency
      const double cost = work * fill * (0.75 + 0.25 * halo);
      if (cost < best_cost - 1e-9) { best_cost = cost; best_k = k; }
    }
    p.k = best_k;
    p.items_per_strip = (int)((lin + 128LL * p.k - 1) / (128LL * p.k));
    p.n_items = p.items_per_strip * n_strips * cp.B;
    p.npx = 128 * p.k + 2 * p.pitch + 2;
    p.npx = (p.npx + 7) & ~7;
  } else {
    int best_k = 1;
    double best_cost = 1e30;
    for (int k = 1; k <= kmax_tmem; ++k) {
      if (stage_bytes(k, 0) * 3 > smem_cap) break;
      const long long items = (p.total_px + 128LL * k - 1) / (128LL * k);
      const long long waves = (items + num_sms - 1) / num_sms;
      const double cost = (double)(waves * num_sms) * k * 128 / (double)p.total_px + 0.02 / k;
      if (cost < best_cost - 1e-9) { best_cost = cost; best_k = k; }
    }
    p.k = best_k;
    p.TW = p.pitch = p.n_strips = p.items_per_strip = 0;
    p.n_items = (int)((p.total_px + 128LL * p.k - 1) / (128LL * p.k));
    p.npx = 128 * p.k;
  }
  const size_t sb = (size_t)p.npx * 32;
  int stages = (int)(smem_cap / sb);
  if (stages < kUmmaLookahead + 1) return "tile does not fit in shared memory";
  if (stages > 8) stages = 8;
  p.n_stages = stages;
  plan.smem_bytes = w_bytes + (size_t)stages * sb + (2 * stages + 4) * 8 + 16;
  plan.grid = std::min(p.n_items, num_sms);
  return nullptr;
}

template <typename T, int N, int TAPS>
inline const char* umma_launch_inst(const UmmaPlan& plan, cudaStream_t st) {
  static size_t configured = 0;
  if (plan.smem_bytes > configured) {
    cudaError_t e = cudaFuncSetAttribute(umma_conv_kernel<T, N, TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    configured = 227 * 1024;
  }
  umma_conv_kernel<T, N, TAPS><<<plan.grid, kUmmaThreads, plan.smem_bytes, st>>>(plan.p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

template <typename T>
inline const char* umma_conv_launch(const UmmaWeights& w, const ConvParams& cp, int num_sms, cudaStream_t st) {
  UmmaPlan plan;
  if (const char* msg = umma_plan(plan, w, cp, num_sms)) return msg;
  const int taps = w.ks * w.ks;
  if (taps == 9) {
    if (w.cout == 16) return umma_launch_inst<T, 16, 9>(plan, st);
    if (w.cout == 32) return umma_launch_inst<T, 32, 9>(plan, st);
    if (w.cout == 64) return umma_launch_inst<T, 64, 9>(plan, st);
  } else {
    if (w.cout == 16) return umma_launch_inst<T, 16, 1>(plan, st);
    if (w.cout == 32) return umma_launch_inst<T, 32, 1>(plan, st);
    if (w.cout == 64) return umma_launch_inst<T, 64, 1>(plan, st);
  }
  return "unsupported Cout";
}

template <> inline const char* umma_conv_launch<float>(const UmmaWeights&, const ConvParams&, int, cudaStream_t) {
  return "tensor-core path is 16-bit only";
}

}  // namespace lpsr
