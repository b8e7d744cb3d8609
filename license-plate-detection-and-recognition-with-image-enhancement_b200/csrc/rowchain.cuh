// rowchain.cuh -- TWO chained 3x3 convolutions in one row-streaming kernel: layer A (32 -> 32 channels, + residual) feeds layer B (32 -> 16)
// through shared memory, so A's output never goes to HBM.  Used for gff.1 (+ sfe1 residual) -> final conv (+ sigmoid), lpsr.py:211,224,273-274:
// the 32-channel tensor between them is written and read back by nothing else.  sm_100a only, 16-bit modes.
//
// Both layers are the row-streaming convolution of rowconv.cuh (128 lanes of an image row as GEMM-M, dy folded into N and landing in a ring of
// four TMEM slots, dx as a start-address shift, reset-to-bias epilogue).  What is new is the hand-over: because a staged row is consumed ONCE and
// a tile row spans the whole crop width, layer A's output row r -- complete after A's input row r + 1 -- is exactly layer B's staged row r, with
// no halo in either direction.  A's epilogue warps write it (residual added, rounded to the 16-bit type exactly as the unfused kernel would store
// it: the fused result is bit-identical) into a ring of row buffers in the TMA's SWIZZLE_64B layout, zeros in the lanes without a pixel (they
// are the 'same' padding of B's horizontal taps), fence the async proxy and arrive on `mid_full`; B's issuers read it like a TMA-staged row.
// One crop per work item (G = 1, two 128-lane tiles: the two layers' rings take 2 x (128 + 64) TMEM columns), whole crops per item so that B
// never needs a row of A that another CTA computes.
// Warps: 0 TMA producer | 1, 2 issuers of layer A (tile 0, 1) | 3, 4 issuers of layer B | 8..15 epilogue of A | 16..23 epilogue of B.
#pragma once
#include "rowconv.cuh"

namespace lpsr {

constexpr int kChainStagesA = 4;    // staged input rows of layer A in flight
constexpr int kChainRing = 4;       // rows of A's output held for layer B
constexpr int kChainThreads = 24 * 32;

struct ChainParams {
  int coff;                               // channel offset of A's 32 input channels in their tensor
  uint32_t entry_bytes;                   // bytes of one row buffer: (kRowLead + 256 + 8) rows x 64 B
  int P8;                                 // lane pitch of the crop (W + >= 1 zero columns, multiple of 8, <= 256)
  const uint16_t* wa; uint32_t wa_bytes; const float* bias_a;   // layer A: 7-block weights (rowconv_weights.h), Cout 32
  const uint16_t* wb; uint32_t wb_bytes; const float* bias_b;   // layer B: Cout 16
  const void* res; int res_pitch, res_off;                     // residual added to A's output (nullptr: none)
  void* out;                              // B's output: fp32 [pixel] (final sigmoid)
  int B, H, W;
};
struct ChainTmap { CUtensorMap m; };

template <typename T>
__global__ void __launch_bounds__(kChainThreads, 1) rowchain_kernel(const __grid_constant__ ChainParams p, const __grid_constant__ ChainTmap tm) {
  static_assert(sizeof(T) == 2, "16-bit operands");
  constexpr int CA = 32, CB = 16, NBA = 7 * CA, NBB = 7 * CB, NT = 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.H, W = p.W;
  uint8_t* in_smem = smem;                                                   // kChainStagesA entries
  uint8_t* mid_smem = smem + (size_t)kChainStagesA * p.entry_bytes;          // kChainRing entries
  uint8_t* wa_smem = mid_smem + (size_t)kChainRing * p.entry_bytes;
  uint8_t* wb_smem = wa_smem + ((p.wa_bytes + 127u) & ~127u);
  float* bias_s = reinterpret_cast<float*>(wb_smem + ((p.wb_bytes + 127u) & ~127u));   // [32 A | 16 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 64);
  const uint32_t bar0 = ptx::smem_u32(bars);
  auto full_in = [&](int e) { return bar0 + 8u * e; };                                   // [4]  TMA bytes landed
  auto mid_full = [&](int e) { return bar0 + 8u * (4 + e); };                            // [4]  A's output row is in shared memory (8 arrivals)
  auto step_a = [&](int t, int k) { return bar0 + 8u * (8 + t * 4 + k); };               // [2][4] A issuer t: MMAs of staged row s retired
  auto empty_a = [&](int t, int k) { return bar0 + 8u * (16 + t * 4 + k); };             // [2][4] A's slot drained and reset (4 arrivals)
  auto step_b = [&](int t, int k) { return bar0 + 8u * (24 + t * 4 + k); };              // [2][4] B issuer t: MMAs of row g retired
  auto empty_b = [&](int t, int k) { return bar0 + 8u * (32 + t * 4 + k); };             // [2][4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 40);

  // ---- one-time setup ------------------------------------------------------------------------------
  {
    for (uint32_t i = threadIdx.x; i < p.wa_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(wa_smem)[i] = __ldg(reinterpret_cast<const uint4*>(p.wa) + i);
    for (uint32_t i = threadIdx.x; i < p.wb_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(wb_smem)[i] = __ldg(reinterpret_cast<const uint4*>(p.wb) + i);
    uint4* z = reinterpret_cast<uint4*>(smem);                  // row buffers start as zeros (lead rows, rows behind the last lane)
    const uint32_t nz = (uint32_t)(kChainStagesA + kChainRing) * (p.entry_bytes / 16);
    for (uint32_t i = threadIdx.x; i < nz; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x < CA) bias_s[threadIdx.x] = __ldg(p.bias_a + threadIdx.x);
    else if (threadIdx.x < CA + CB) bias_s[threadIdx.x] = __ldg(p.bias_b + threadIdx.x - CA);
  }
  if (threadIdx.x == 0) {
    for (int e = 0; e < 4; ++e) { ptx::mbar_init(full_in(e), 1); ptx::mbar_init(mid_full(e), 8); }
    for (int t = 0; t < NT; ++t)
      for (int k = 0; k < 4; ++k) {
        ptx::mbar_init(step_a(t, k), 1); ptx::mbar_init(empty_a(t, k), 4);
        ptx::mbar_init(step_b(t, k), 1); ptx::mbar_init(empty_b(t, k), 4);
      }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: layer A tile t at t * 128 (4 slots x 32), layer B tile t at 256 + t * 64 (4 slots x 16)
  const bool epi_a = warp >= 8 && warp < 16, epi_b = warp >= 16;
  const int et = ((warp - 8) >> 2) & 1, eq = warp & 3;
  if (epi_a || epi_b) {                                         // every accumulator slot starts at its bias vector
    const int C = epi_a ? CA : CB;
    const uint32_t taddr = tmem_base + ((uint32_t)(eq * 32) << 16) + (uint32_t)(epi_a ? et * 128 : 256 + et * 64);
    for (int hb = 0; hb < C / 16; ++hb) {
      float b[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) b[i] = bias_s[(epi_a ? 0 : CA) + hb * 16 + i];
      for (int k = 0; k < 4; ++k) ptx::tc_st16(taddr + (uint32_t)(k * C + hb * 16), b);
    }
    ptx::tc_wait_st();
    ptx::tc_fence_before();
  }
  __syncthreads();
  ptx::tc_fence_after();
  ptx::griddep_wait();

  const int n_my = (p.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // crops of this CTA: blockIdx.x, + gridDim.x, ...
  const long long n_rows = (long long)n_my * H;                // rows of this CTA's sequence; index g (multiple-of-4 aligned per crop: H % 4 == 0)

  if (warp == 0) {
    // =================================== TMA producer (layer A's input rows) ==============================
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&tm.m);
      const uint32_t tx = (uint32_t)p.P8 * 64u;
      long long g = 0;
      for (int ii = 0; ii < n_my; ++ii) {
        const int n = blockIdx.x + ii * gridDim.x;
        for (int y = 0; y < H; ++y, ++g) {
          if (g >= kChainStagesA) {                             // the row that used this entry last has been consumed by both A issuers
            const long long sp = g - kChainStagesA;
            for (int t = 0; t < NT; ++t) ptx::mbar_wait(step_a(t, (int)(sp & 3)), (uint32_t)(sp >> 2) & 1u);
          }
          const int e = (int)(g % kChainStagesA);
          ptx::mbar_arrive_expect_tx(full_in(e), tx);
          ptx::tma_load_4d(ptx::smem_u32(in_smem + (size_t)e * p.entry_bytes) + (uint32_t)(kRowLead * 64), &tm.m, full_in(e), p.coff, 0, y, n);
        }
      }
    }
  } else if (warp >= 1 && warp <= 4) {
    // =================================== MMA issuers: warps 1, 2 layer A; warps 3, 4 layer B =======================
    if (ptx::elect_one()) {
      const bool is_b = warp >= 3;
      const int t = (warp - 1) & 1;
      constexpr bool kBf16 = IsBf16<T>::value;
      const int C = is_b ? CB : CA, NB = is_b ? NBB : NBA;
      const uint32_t w_lo = umma_desc_lo(ptx::smem_u32(is_b ? wb_smem : wa_smem), (uint32_t)NB * 16);
      const uint32_t src16 = ptx::smem_u32(is_b ? mid_smem : in_smem) >> 4, entry16 = p.entry_bytes >> 4;
      const uint32_t idesc1 = umma_idesc_f16(kBf16, C), idesc2 = umma_idesc_f16(kBf16, 2 * C), idesc4 = umma_idesc_f16(kBf16, 4 * C);
      const uint32_t a_hi = umma_desc_hi_swizzled(64);
      const uint32_t trow = (uint32_t)(kRowLead + 128 * t - 1) * 4u;          // 64-byte rows: 4 sixteen-byte units per row
      const uint32_t dtile = tmem_base + (uint32_t)(is_b ? 256 + t * 64 : t * 128);
      for (long long g = 0; g < n_rows; ++g) {
        const int yin = (int)(g % H);
        const long long g0 = g - yin;                           // index of the crop's first row
        const int ra = max(yin - 1, 0), rb = min(yin + 1, H - 1), nr = rb - ra + 1;
        const bool full = (nr == 3);
        const uint32_t idesc = full ? idesc4 : (nr == 2 ? idesc2 : idesc1);
        const uint32_t d = dtile + (full ? 0u : (uint32_t)((ra & 3) * C));
        const uint32_t blk = full ? (uint32_t)((4 - ((yin - 1) & 3)) & 3) : (uint32_t)(1 - yin + ra);
        const long long need = g0 + rb - (full ? 3 : 4);        // earlier occupants of the slots this row touches must have been drained
        const uint32_t wb = w_lo + blk * (uint32_t)C;
        const int e = (int)(g % 4);                             // kChainStagesA == kChainRing == 4
        ptx::mbar_wait(is_b ? mid_full(e) : full_in(e), (uint32_t)(g >> 2) & 1u);
        if (need >= 0) ptx::mbar_wait(is_b ? empty_b(t, (int)(need & 3)) : empty_a(t, (int)(need & 3)), (uint32_t)(need >> 2) & 1u);
        ptx::tc_fence_after();
        const uint32_t a0 = src16 + (uint32_t)e * entry16 + trow;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx)
            ptx::tc_mma_f16_lohi(d, (a0 + 2u * (uint32_t)ks + 4u * (uint32_t)dx) | (1u << 16), a_hi, wb + (uint32_t)((dx * 2 + ks) * 2 * NB), kUmmaDescHi, idesc, 1u);
        ptx::tc_commit(is_b ? step_b(t, (int)(g & 3)) : step_a(t, (int)(g & 3)));
      }
    }
  } else if (epi_a) {
    // =================================== epilogue of layer A: TMEM -> (+ residual) -> shared-memory row for layer B ==============
    const int L = 128 * et + 32 * eq + lane;                     // lane of the row == pixel x (G = 1)
    const bool in_img = L < W;
    const uint32_t taddr = tmem_base + ((uint32_t)(eq * 32) << 16) + (uint32_t)(et * 128);
    const T* res = static_cast<const T*>(p.res);
    const uint32_t row = (uint32_t)(kRowLead + L);
    const uint32_t row_off = row * 64u, sw = (row >> 1) & 3u;     // SWIZZLE_64B: 16-byte chunk j of a row sits at j ^ ((row >> 1) & 3)
    long long g = 0;
    for (int ii = 0; ii < n_my; ++ii) {
      const int n = blockIdx.x + ii * gridDim.x;
      int pix = in_img ? n * H * W + L : -1;
      for (int r = 0; r < H; ++r, ++g) {
        __syncwarp();
        uint4 rq[4];
        if (res && pix >= 0) {
          const uint4* rp = reinterpret_cast<const uint4*>(res + (size_t)pix * p.res_pitch + p.res_off);
#pragma unroll
          for (int i = 0; i < 4; ++i) rq[i] = __ldg(rp + i);
        }
        const int slot = r & 3;
        const long long sc = g - r + min(r + 1, H - 1);          // the staged row whose MMAs complete output row r
        ptx::mbar_wait(step_a(et, (int)(sc & 3)), (uint32_t)(sc >> 2) & 1u);
        ptx::tc_fence_after();
        float v[32];
        ptx::tc_ld16_nowait(taddr + (uint32_t)(slot * CA), v);
        ptx::tc_ld16_nowait(taddr + (uint32_t)(slot * CA + 16), v + 16);
        ptx::tc_wait_ld();
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          float b[16];
#pragma unroll
          for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(b + i) = *reinterpret_cast<const float4*>(bias_s + hb * 16 + i);
          ptx::tc_st16(taddr + (uint32_t)(slot * CA + hb * 16), b);
        }
        ptx::tc_wait_st();
        ptx::tc_fence_before();
        __syncwarp();
        if (ptx::elect_one()) ptx::mbar_arrive(empty_a(et, slot));
        // the ring entry of this row: layer B must have consumed the row that used it last (both B issuers passed row g - kChainRing)
        if (g >= kChainRing) {
          const long long sp = g - kChainRing;
          for (int t = 0; t < NT; ++t) ptx::mbar_wait(step_b(t, (int)(sp & 3)), (uint32_t)(sp >> 2) & 1u);
        }
        uint8_t* dst = mid_smem + (size_t)(g % kChainRing) * p.entry_bytes + row_off;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 q = make_uint4(0u, 0u, 0u, 0u);                  // lanes without a pixel: the zero padding of B's horizontal taps
          if (pix >= 0) {
            float o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = v[j * 8 + k];
            if (res) {
              const T* e = reinterpret_cast<const T*>(&rq[j]);
#pragma unroll
              for (int k = 0; k < 8; ++k) o[k] += to_f32<T>(e[k]);
            }
            q = make_uint4(pack2<T, false>(o[0], o[1]), pack2<T, false>(o[2], o[3]), pack2<T, false>(o[4], o[5]), pack2<T, false>(o[6], o[7]));
          }
          *reinterpret_cast<uint4*>(dst + (((uint32_t)j ^ sw) << 4)) = q;
        }
        ptx::fence_proxy_async();                                 // generic-proxy stores -> visible to the tensor core's operand fetch
        __syncwarp();
        if (ptx::elect_one()) ptx::mbar_arrive(mid_full((int)(g % kChainRing)));
        if (pix >= 0) pix += W;
      }
    }
  } else if (epi_b) {
    // =================================== epilogue of layer B: channel 0 -> logistic -> fp32 [pixel] ==============================
    const int L = 128 * et + 32 * eq + lane;
    const bool in_img = L < W;
    const uint32_t taddr = tmem_base + ((uint32_t)(eq * 32) << 16) + (uint32_t)(256 + et * 64);
    float* out = static_cast<float*>(p.out);
    long long g = 0;
    for (int ii = 0; ii < n_my; ++ii) {
      const int n = blockIdx.x + ii * gridDim.x;
      int pix = in_img ? n * H * W + L : -1;
      for (int r = 0; r < H; ++r, ++g) {
        __syncwarp();
        const int slot = r & 3;
        const long long sc = g - r + min(r + 1, H - 1);
        ptx::mbar_wait(step_b(et, (int)(sc & 3)), (uint32_t)(sc >> 2) & 1u);
        ptx::tc_fence_after();
        float v[16];
        ptx::tc_ld16(taddr + (uint32_t)(slot * CB), v);
        {
          float b[16];
#pragma unroll
          for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(b + i) = *reinterpret_cast<const float4*>(bias_s + CA + i);
          ptx::tc_st16(taddr + (uint32_t)(slot * CB), b);
        }
        ptx::tc_wait_st();
        ptx::tc_fence_before();
        __syncwarp();
        if (ptx::elect_one()) ptx::mbar_arrive(empty_b(et, slot));
        if (pix >= 0) {
          out[pix] = __fdividef(1.f, 1.f + __expf(-v[0]));        // final conv + sigmoid, lpsr.py:273-274
          pix += W;
        }
      }
    }
  }
  // ---- teardown ----------------------------------------------------------------------------------------
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

// OFF by default -- a measured experiment.  The chained kernel is bit-identical to the two launches (tests/test_gpu_parity.py) and moves 132 instead of
// 260 bytes per pixel, but at B = 1024 it takes 846 us against 442 + 180 us: with one crop per item a layer has only TWO tiles, and the per-row
// chain of a tile (MMAs retire -> commit -> epilogue drains and resets the slot -> the issuer may touch it again, one row of slack in a 4-slot ring:
// ~1500 clk) cannot be hidden behind one other tile (3700 clk per row against 1344 clk of tensor time).  More tiles per layer do not fit: two
// layers' rings at G = 2 need 768 TMEM columns.  LPSR_ROWCHAIN=1 switches it on (read per launch).
inline bool rowchain_enabled() {
  const char* e = getenv("LPSR_ROWCHAIN");
  return e && e[0] == '1' && rowconv_enabled();
}

// gff.1 (+ residual) -> final conv (+ sigmoid).  Returns nullptr when launched; "unhandled" conditions set *handled = false (the caller runs the two
// layers one after the other); any other string is a launch error.
template <typename T>
inline const char* rowchain_launch(const RowWeights& wa, const RowWeights& wb, const void* in, int in_pitch, int in_off, const void* res, int res_pitch,
                                   int res_off, float* out, int B, int H, int W, int num_sms, cudaStream_t st, bool* handled, bool dry = false) {
  *handled = false;
  if constexpr (sizeof(T) != 2) {
    return nullptr;
  } else {
    if (!rowchain_enabled() || !wa.packed || !wb.packed || wa.cin != 32 || wa.cout != 32 || wb.cin != 32 || wb.cout != 16) return nullptr;
    if (H % 4 || H < 4 || B < 1) return nullptr;
    const int P8 = (W + 1 + 7) / 8 * 8;
    if (P8 > 256) return nullptr;
    if (in_pitch % 8 || in_off % 8 || reinterpret_cast<uintptr_t>(in) % 16) return nullptr;
    if (res && (res_pitch % 8 || res_off % 8 || reinterpret_cast<uintptr_t>(res) % 16)) return nullptr;
    if ((long long)B * H * W >= (1LL << 31)) return nullptr;
    const char* mu = getenv("LPSR_ROWCONV_MIN_UNITS");
    const int min_units = mu ? atoi(mu) : 4;
    if ((long long)B * (H / 4) < (long long)min_units * num_sms) return nullptr;       // same rule as rowconv.cuh (one crop per group here)
    ChainParams p{};
    p.coff = in_off;
    p.entry_bytes = (uint32_t)((kRowLead + 256 + 8) * 64);
    p.P8 = P8;
    p.wa = wa.w; p.wa_bytes = wa.bytes; p.bias_a = wa.bias;
    p.wb = wb.w; p.wb_bytes = wb.bytes; p.bias_b = wb.bias;
    p.res = res; p.res_pitch = res_pitch; p.res_off = res_off;
    p.out = out;
    p.B = B; p.H = H; p.W = W;
    const size_t smem = 1024 + (size_t)(kChainStagesA + kChainRing) * p.entry_bytes + ((wa.bytes + 127) & ~127u) + ((wb.bytes + 127) & ~127u) + 64 * 4 + 48 * 8;
    if (smem > 227 * 1024) return nullptr;
    if (dry) { *handled = true; return nullptr; }
    ChainTmap tm;
    memset(&tm, 0, sizeof tm);
    PFN_lpsr_tmapEncodeTiled enc = umma_encode_fn();
    if (!enc) return "cuTensorMapEncodeTiled entry point not found";
    const cuuint64_t gdim[4] = {(cuuint64_t)in_pitch, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t gstr[3] = {(cuuint64_t)in_pitch * 2, (cuuint64_t)W * in_pitch * 2, (cuuint64_t)H * W * in_pitch * 2};
    const cuuint32_t box[4] = {32, (cuuint32_t)P8, 1, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    if (enc(&tm.m, IsBf16<T>::value ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(in), gdim, gstr, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return "cuTensorMapEncodeTiled failed";
    *handled = true;
    static bool configured[kMaxDevices] = {};
    bool* flag = func_configured_flag(configured);
    if (!flag || !*flag) {
      cudaError_t e = cudaFuncSetAttribute(rowchain_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return cudaGetErrorString(e);
      if (flag) *flag = true;
    }
    cudaError_t e = launch_pdl(rowchain_kernel<T>, dim3(std::min(B, num_sms)), dim3(kChainThreads), smem, st, p, tm);
    if (e == cudaSuccess) e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
  }
}

}  // namespace lpsr
