// rowconv.cuh -- row-streaming 3x3 convolution on tcgen05 / TMEM: the dy taps are folded into GEMM-N and land in NEIGHBOURING TMEM
// accumulators, so the epilogue is one TMEM load, a convert and one store per output row (no shifted sums).  sm_100a only, 16-bit modes.
//
// Why: the dx-folded kernel of umma_conv.cuh is bound by its epilogue's shifted sum (~286 instructions per 128-row tile and
// warp), the per-tap kernel by 9 small-N MMAs per K-slice (an M=128,K=16 tcgen05.mma costs ~48 clk for any N <= 64).  Here:
//   * GEMM-M = 128 LANES of one image row: lane L of a row is pixel x = L % P8 of crop L / P8 (G crops side by side, each padded to a pitch
//     P8 = W + >= 1 zero columns, a multiple of 8).  The lane <-> pixel map is the same for every row of a crop group, so output row r of
//     tile t always lives in the same TMEM lanes.
//   * One input row y is staged by TMA ([ch][P8 px][1 row][G crops] box per K-chunk; the zero columns between crops and beyond the
//     image are the TMA's out-of-bounds fill) and consumed ONCE: for every K-slice and dx tap one MMA whose weight blocks
//     [ky=2 | ky=1 | ky=0] add the row's contributions to output rows y-1, y, y+1.  The accumulators of a tile form a ring of 4 slots of
//     Cout columns (slot = output row % 4).  Where the three slots are adjacent the MMA has N = 3 * Cout; where they wrap around the ring it
//     has N = 4 * Cout with a zero block for the fourth slot, and the rotation of the blocks against the slots is a different START ADDRESS
//     into a 7-block copy [ky2 ky1 ky0 0 ky2 ky1 ky0] of the weights -- no MMA is ever split at the wrap-around (an N = 64 MMA costs the
//     same 48 clk as N = 48).  dx is a row shift of the A descriptor's start address (same trick as umma_conv.cuh).
//   * 3 * Cin/16 MMAs per tile and row instead of 9 * Cin/16 (Cout = 32) -- and, for Cout = 16, the same 3 * Cin/16 as the dx fold but with
//     an epilogue of ~40 instead of ~286 instructions: after input row y + 1 the slot of output row y is complete; an epilogue warp loads
//     it, RESETS it to the bias vector (tcgen05.st: the next occupant starts at its bias, no add in the epilogue) and hands it back.
//   * Work is cut in units of 4 output rows of one crop group, dealt to the persistent CTAs as contiguous ranges (balance to one unit);
//     where a range starts or ends inside a crop the neighbouring input row is staged once more and contributes through a narrower MMA.
#pragma once
#include "rowconv_weights.h"
#include "umma_conv.cuh"

namespace lpsr {

constexpr int kRowMaxChunks = 8;   // TMA boxes (K-chunks of 16/32/64 channels) per input row
constexpr int kRowMaxTiles = 5;    // 128-lane tiles per row (epilogue warps: 4 per tile)
constexpr int kRowMaxStages = 6;   // ring entries (staged sub-rows) in flight
constexpr int kRowMaxSub = 4;      // sub-rows (groups of K-chunks staged and consumed together) per input row
constexpr int kRowLead = 8;        // zero rows in front of lane 0 (one swizzle atom): the dx = -1 tap of lane 0 reads the last of them

struct RowParams {
  int n_chunks;
  int chunk_ch[kRowMaxChunks], chunk_coff[kRowMaxChunks];
  uint32_t chunk_smem[kRowMaxChunks];     // byte offset of the chunk's row buffer inside its ring entry (1024-aligned)
  // an input row is staged as n_sub SUB-ROWS (consecutive K-chunks of <= 64 bytes per pixel together): uniform ring entries, so a wide
  // layer (80 channels) keeps as many bytes in flight as a narrow one and its MMAs start when the first chunk has landed
  int n_sub, sub_c0[kRowMaxSub + 1], sub_step0[kRowMaxSub + 1];
  uint32_t sub_tx[kRowMaxSub];            // bytes the TMA boxes of one sub-row deliver
  int n_ks;                               // Cin / 16
  uint32_t stage_bytes;                   // bytes of one ring entry
  int n_stages;                           // ring entries
  int G, P8, NT;                          // crops per group, lane pitch per crop, 128-lane tiles per row
  int n_xbox, box_px;                     // a crop row wider than one TMA box (256 pixels) is staged as n_xbox boxes of box_px pixels (then G = 1)
  const uint16_t* w; uint32_t w_bytes; const float* bias;
  void* out; int out_pitch, out_off;
  const void* res; int res_pitch, res_off;
  int B, H, W;
  int n_groups;
  long long* trace;                       // LPSR_ROW_TRACE=1 (profiling experiments): clock64 stamps of CTA 0, [step][8]
  int debug;                              // LPSR_UMMA_DEBUG bitmask (profiling experiments only): 1 skip MMAs, 2 skip stores, 4 skip TMA loads
};
struct RowTmaps { CUtensorMap m[kRowMaxChunks]; };

namespace ptx {
__device__ __forceinline__ void tc_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
               "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
               "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
               : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
}  // namespace ptx

constexpr int kRowEpiWarp0 = 8;    // warps: 0 TMA producer, 1..5 MMA issuers (one thread each, one per tile), 8.. epilogue (4 per tile)
__host__ __device__ constexpr int rowconv_max_tiles(int cout) { return cout == 16 ? kRowMaxTiles : 4; }   // 4 slots x Cout columns x tiles <= 512
__host__ __device__ constexpr int rowconv_max_threads(int cout) { return (kRowEpiWarp0 + 4 * rowconv_max_tiles(cout)) * 32; }

// EPI: kEpiPlain (+bias) | kEpiRelu (+bias, ReLU) | kEpiResidual (+bias, + residual tensor) | kEpiFinalSigmoid (channel 0 -> logistic -> fp32
// [pixel]: the final conv with Cout padded 1 -> 16, lpsr.py:273-274)
//
// Synchronisation (all mbarriers in shared memory; s = index of an input row in this CTA's sequence of staged rows):
//   full_in[e]        TMA bytes of ring entry e have landed                       producer -> issuers
//   step_done[t][s&3] issuer t's MMAs of staged row s have retired (ONE tcgen05.commit per row and issuer: a commit costs the issuing
//                     thread ~300 clk, tools/rowconv_trace.py).  Tells tile t's epilogue warps that the output row completed by row s is
//                     in TMEM, and the producer that row s's ring entries are free once every issuer has passed it
//   acc_empty[t][r&3] the four epilogue warps of tile t have read output row r's slot and reset it to the bias      epilogue -> issuer t
template <typename T, int COUT, int EPI>
__global__ void __launch_bounds__(rowconv_max_threads(COUT), 1) rowconv_kernel(const __grid_constant__ RowParams p, const __grid_constant__ RowTmaps tm) {
  static_assert(sizeof(T) == 2 && (COUT == 16 || COUT == 32), "16-bit operands, Cout 16 or 32");
  constexpr int NB = 7 * COUT;                                  // weight rows per K core matrix
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NT = p.NT, NS = p.n_stages, H = p.H;
  uint8_t* a_smem = smem;
  uint8_t* w_smem = smem + (size_t)NS * p.stage_bytes;
  float* bias_s = reinterpret_cast<float*>(w_smem + ((p.w_bytes + 127u) & ~127u));
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 32);
  const uint32_t bar0 = ptx::smem_u32(bars);
  auto full_in = [&](int e) { return bar0 + 8u * e; };
  auto step_done = [&](int t, int k) { return bar0 + 8u * (kRowMaxStages + t * 4 + k); };
  auto acc_empty = [&](int t, int k) { return bar0 + 8u * (kRowMaxStages + 4 * kRowMaxTiles + t * 4 + k); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kRowMaxStages + 8 * kRowMaxTiles);
  // per-MMA operand table {A offset in 16-B units inside a ring entry (chunk, K-slice, dx shift), row bytes / 16, A descriptor hi word, weight offset}
  uint4* steps = reinterpret_cast<uint4*>(bars + kRowMaxStages + 8 * kRowMaxTiles + 2);
  int* sub_s = reinterpret_cast<int*>(steps + 3 * kMaxChunks);   // [n_sub + 1] first table entry of every sub-row

  // ---- one-time setup ------------------------------------------------------------------------------
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.w);
    uint4* dst = reinterpret_cast<uint4*>(w_smem);
    for (uint32_t i = threadIdx.x; i < p.w_bytes / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    // the ring entries start as zeros: the kRowLead rows in front of lane 0 are never written by the TMA, and the rows behind the last
    // crop only ever feed lanes without a pixel (they must stay finite: 0 * NaN would poison the idle accumulator slot)
    uint4* z = reinterpret_cast<uint4*>(a_smem);
    const uint32_t nz = (uint32_t)NS * (p.stage_bytes / 16);
    for (uint32_t i = threadIdx.x; i < nz; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x < COUT) bias_s[threadIdx.x] = __ldg(p.bias + threadIdx.x);
  }
  if (threadIdx.x == 0) {
    for (int e = 0; e < NS; ++e) ptx::mbar_init(full_in(e), 1);
    for (int sb = 0; sb <= p.n_sub; ++sb) sub_s[sb] = p.sub_step0[sb];
    int i = 0, ks = 0;
    for (int c = 0; c < p.n_chunks; ++c) {
      const uint32_t rb16 = (uint32_t)p.chunk_ch[c] >> 3;
      for (int kk = 0; kk < (p.chunk_ch[c] >> 4); ++kk, ++ks)
        for (int dx = 0; dx < 3; ++dx)
          steps[i++] = make_uint4((p.chunk_smem[c] >> 4) + 2u * (uint32_t)kk + (uint32_t)dx * rb16, rb16, umma_desc_hi_swizzled(rb16 << 4),
                                  (uint32_t)((dx * p.n_ks + ks) * 2 * NB));
    }
    for (int t = 0; t < NT; ++t)
      for (int k = 0; k < 4; ++k) { ptx::mbar_init(step_done(t, k), 1); ptx::mbar_init(acc_empty(t, k), 4); }
    ptx::fence_mbar_init();
  }
  const uint32_t tmem_cols = (NT * 4 * COUT <= 64) ? 64u : (NT * 4 * COUT <= 128) ? 128u : (NT * 4 * COUT <= 256) ? 256u : 512u;
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int et = (warp - kRowEpiWarp0) >> 2, eq = warp & 3;    // epilogue warps: tile, TMEM lane quadrant
  if (warp >= kRowEpiWarp0) {
    // every accumulator slot starts at the bias vector
    const uint32_t taddr = tmem_base + ((uint32_t)(eq * 32) << 16) + (uint32_t)(et * 4 * COUT);
#pragma unroll
    for (int hb = 0; hb < COUT / 16; ++hb) {
      float b[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) b[i] = bias_s[hb * 16 + i];
#pragma unroll
      for (int k = 0; k < 4; ++k) ptx::tc_st16(taddr + (uint32_t)(k * COUT + hb * 16), b);
    }
    ptx::tc_wait_st();
    ptx::tc_fence_before();
  }
  __syncthreads();
  ptx::tc_fence_after();
  ptx::griddep_wait();                                          // everything above touched only weights and on-chip state

  // ---- this CTA's range of work units (4 output rows of one crop group) ---------------------------------
  const int U = H >> 2;
  const long long total_u = (long long)p.n_groups * U;
  const long long per = total_u / gridDim.x, rem = total_u % gridDim.x;
  const long long u0 = (long long)blockIdx.x * per + min((long long)blockIdx.x, rem);
  const long long u1 = u0 + per + ((long long)blockIdx.x < rem ? 1 : 0);

  if (warp == 0) {
    // =================================== TMA producer ==============================================
    if (ptx::elect_one()) {
      for (int c = 0; c < p.n_chunks; ++c) ptx::prefetch_tmap(&tm.m[c]);
      const int n_sub = p.n_sub;
      const int depth = NS / n_sub;                             // staged rows in flight (every row takes n_sub consecutive ring entries)
      int entry = 0;
      int sidx = 0;                                             // staged-row index
      for (long long u = u0; u < u1;) {
        const int grp = (int)(u / U), ua = (int)(u - (long long)grp * U), ub = (int)min((long long)U, ua + (u1 - u));
        const int ya = 4 * ua, yb = 4 * ub;
        const int ylo = max(ya - 1, 0), yhi = min(yb, H - 1);
        for (int yin = ylo; yin <= yhi; ++yin, ++sidx) {
          if (sidx >= depth) {                                  // the row that used these entries last has been consumed by every issuer
            const int sp = sidx - depth;
            for (int t = 0; t < NT; ++t) ptx::mbar_wait(step_done(t, sp & 3), (uint32_t)(sp >> 2) & 1u);
          }
          for (int sb = 0; sb < n_sub; ++sb) {
            const uint32_t bar = full_in(entry);
            if (p.debug & 4) {
              ptx::mbar_arrive(bar);
            } else {
              ptx::mbar_arrive_expect_tx(bar, p.sub_tx[sb]);
              const uint32_t dst0 = ptx::smem_u32(a_smem + (size_t)entry * p.stage_bytes);
              for (int c = p.sub_c0[sb]; c < p.sub_c0[sb + 1]; ++c)
                for (int xb = 0; xb < p.n_xbox; ++xb)
                  ptx::tma_load_4d(dst0 + p.chunk_smem[c] + (uint32_t)((kRowLead + xb * p.box_px) * p.chunk_ch[c] * 2), &tm.m[c], bar, p.chunk_coff[c],
                                   xb * p.box_px, yin, grp * p.G);
            }
            ++entry;
          }
          if (entry == depth * n_sub) entry = 0;
        }
        u += ub - ua;
      }
    }
  } else if (warp >= 1 && warp <= kRowMaxTiles) {
    // =================================== MMA issuers ===============================================
    // One issuing THREAD per tile (warp 1 + t), each running its whole loop alone: tcgen05.mma blocks the issuing thread while the pipe's
    // short queue is full, so an issuer's serial work per row (barrier waits, fence, commit: ~700 clk) adds to ITS MMAs' execution time;
    // with one tile per issuer that sum stays below the row's tensor time and the issuers de-phase against each other.
    const int t = warp - 1;
    if (t < NT && ptx::elect_one()) {
      constexpr bool kBf16 = IsBf16<T>::value;
      const uint32_t w_lo = umma_desc_lo(ptx::smem_u32(w_smem), (uint32_t)NB * 16);
      const uint32_t a_smem16 = ptx::smem_u32(a_smem) >> 4, stage16 = p.stage_bytes >> 4;
      const int n_sub = p.n_sub;
      const int ring = (NS / n_sub) * n_sub;
      const bool no_mma = (p.debug & 1) != 0;
      constexpr uint32_t idesc1 = umma_idesc_f16(kBf16, COUT), idesc2 = umma_idesc_f16(kBf16, 2 * COUT), idesc3 = umma_idesc_f16(kBf16, 3 * COUT),
                         idesc4 = umma_idesc_f16(kBf16, 4 * COUT);
      const uint32_t trow = (uint32_t)(kRowLead + 128 * t - 1);
      const uint32_t dtile = tmem_base + (uint32_t)(t * 4 * COUT);
      int entry = 0;
      uint32_t ph = 0;
      int g0 = 0;                                               // index (in this CTA's output-row sequence) of the band's first row; multiple of 4
      int sidx = 0;
      long long* const trace = (blockIdx.x == 0 && t == 0) ? p.trace : nullptr;
#define ROW_TRACE(k) do { if (trace && sidx < 512) trace[sidx * 8 + (k)] = clock64(); } while (0)
      for (long long u = u0; u < u1;) {
        const int grp = (int)(u / U), ua = (int)(u - (long long)grp * U), ub = (int)min((long long)U, ua + (u1 - u));
        const int ya = 4 * ua, yb = 4 * ub;
        const int ylo = max(ya - 1, 0), yhi = min(yb, H - 1);
        for (int yin = ylo; yin <= yhi; ++yin, ++sidx) {
          // output rows of this band the input row contributes to
          const int ra = max(yin - 1, ya), rb = min(yin + 1, yb - 1), nr = rb - ra + 1;
          // three rows whose slots do not wrap around the ring (slot(ra) <= 1) take an N = 3 * Cout window like the partial ones: the fourth slot is
          // not touched, so its previous row may still be draining (one more row of slack), and N = 96 costs 56 instead of 64 clk
          const bool full = (ra & 3) + nr > 4;                  // only a three-row window can wrap (band ends sit on multiples of 4)
          const uint32_t idesc = full ? idesc4 : (nr == 3 ? idesc3 : (nr == 2 ? idesc2 : idesc1));
          const uint32_t d = dtile + (full ? 0u : (uint32_t)((ra & 3) * COUT));
          // first weight block of the window: full ring -> slot 0 gets the block of ((0 - slot(yin-1)) mod 4); partial -> the block of row ra
          const uint32_t blk = full ? (uint32_t)((4 - ((yin - 1) & 3)) & 3) : (uint32_t)(1 - yin + ra);
          const int need = g0 + (rb - ya) - (full ? 3 : 4);     // every earlier occupant of the slots this row touches must have been drained
          const uint32_t wb = w_lo + blk * (uint32_t)COUT;
          ROW_TRACE(0);
          ptx::mbar_wait(full_in(entry), ph);
          ROW_TRACE(1);
          if (need >= 0) {
            ptx::mbar_wait(acc_empty(t, need & 3), (uint32_t)(need >> 2) & 1u);
            ptx::tc_fence_after();
          }
          ROW_TRACE(2);
          for (int sb = 0; sb < n_sub; ++sb) {
            if (sb) ptx::mbar_wait(full_in(entry), ph);
            const uint32_t sbase16 = a_smem16 + (uint32_t)entry * stage16;
            if (!no_mma) {
#pragma unroll 3
              for (int i = sub_s[sb]; i < sub_s[sb + 1]; ++i) {
                const uint4 e = steps[i];
                ptx::tc_mma_f16_lohi(d, (sbase16 + e.x + trow * e.y) | (1u << 16), e.z, wb + e.w, kUmmaDescHi, idesc, 1u);
              }
            }
            ++entry;
          }
          ROW_TRACE(3);
          ptx::tc_commit(step_done(t, sidx & 3));               // this tile's MMAs of the staged row have retired
          ROW_TRACE(4);
          if (entry == ring) { entry = 0; ph ^= 1u; }
        }
        g0 += yb - ya;
        u += ub - ua;
      }
#undef ROW_TRACE
    }
  } else if (warp >= kRowEpiWarp0 && et < NT) {
    // =================================== epilogue warps ==============================================
    const int L = 128 * et + 32 * eq + lane;
    const int j = L / p.P8, x = L - j * p.P8;
    const uint32_t taddr = tmem_base + ((uint32_t)(eq * 32) << 16) + (uint32_t)(et * 4 * COUT);
    T* out = static_cast<T*>(p.out);
    [[maybe_unused]] const T* res = static_cast<const T*>(p.res);
    const int W = p.W, out_pitch = p.out_pitch, out_off = p.out_off;
    [[maybe_unused]] const int res_pitch = p.res_pitch, res_off = p.res_off;
    const bool skip_store = (p.debug & 2) != 0;
    int g = 0;                                                  // output-row index in this CTA's sequence (slot = g & 3)
    int s0 = 0;                                                 // staged-row index of the band's first input row
    for (long long u = u0; u < u1;) {
      const int grp = (int)(u / U), ua = (int)(u - (long long)grp * U), ub = (int)min((long long)U, ua + (u1 - u));
      const int ya = 4 * ua, yb = 4 * ub;
      const int ylo = max(ya - 1, 0), yhi = min(yb, H - 1);
      const int n = grp * p.G + j;
      const bool valid = (j < p.G) && (x < W) && (n < p.B);
      int pix = valid ? (n * H + ya) * W + x : -1;
      for (int r = ya; r < yb; ++r, ++g) {
        __syncwarp();
        [[maybe_unused]] uint4 rq[COUT / 8];
        if constexpr (EPI == kEpiResidual) {
          if (pix >= 0) {
            const uint4* rp = reinterpret_cast<const uint4*>(res + (size_t)pix * res_pitch + res_off);
#pragma unroll
            for (int i = 0; i < COUT / 8; ++i) rq[i] = __ldg(rp + i);
          }
        }
        const int slot = r & 3;
        const int sc = s0 + min(r + 1, yhi) - ylo;              // the staged row whose MMAs complete output row r
        const bool tr = p.trace && blockIdx.x == 0 && warp == kRowEpiWarp0 && lane == 0 && g < 512;
        ptx::mbar_wait(step_done(et, sc & 3), (uint32_t)(sc >> 2) & 1u);
        ptx::tc_fence_after();
        if (tr) p.trace[g * 8 + 5] = clock64();
#pragma unroll
        for (int hb = 0; hb < COUT / 16; ++hb) {
          if (hb) __syncwarp();
          float v[16];
          ptx::tc_ld16(taddr + (uint32_t)(slot * COUT + hb * 16), v);       // includes tcgen05.wait::ld
          {
            float b[16];
#pragma unroll
            for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(b + i) = *reinterpret_cast<const float4*>(bias_s + hb * 16 + i);
            ptx::tc_st16(taddr + (uint32_t)(slot * COUT + hb * 16), b);      // the slot's next occupant starts at the bias
          }
          if (hb == COUT / 16 - 1) {
            ptx::tc_wait_st();
            ptx::tc_fence_before();
            __syncwarp();
            if (ptx::elect_one()) ptx::mbar_arrive(acc_empty(et, slot));
            if (tr) p.trace[g * 8 + 6] = clock64();
          }
          if (pix >= 0 && !skip_store) {
            if constexpr (EPI == kEpiResidual) {
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const T* e = reinterpret_cast<const T*>(&rq[hb * 2 + i]);
#pragma unroll
                for (int k = 0; k < 8; ++k) v[i * 8 + k] += to_f32<T>(e[k]);
              }
            }
            if constexpr (EPI == kEpiFinalSigmoid) static_cast<float*>(p.out)[pix] = __fdividef(1.f, 1.f + __expf(-v[0]));   // lanes = consecutive pixels
            else store_chunk16<T, EPI == kEpiRelu>(out, out_pitch, out_off + hb * 16, pix, v);
          }
        }
        if (pix >= 0) pix += W;
        if (tr) p.trace[g * 8 + 7] = clock64();
      }
      s0 += yhi - ylo + 1;
      u += ub - ua;
    }
  }
  // ---- teardown ----------------------------------------------------------------------------------------
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, tmem_cols);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
struct RowPlan {
  RowParams p;
  RowTmaps tm;
  size_t smem_bytes;
  int grid, threads;
};

inline bool rowconv_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("LPSR_ROWCONV"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

// nullptr: planned; otherwise why this shape is not handled (the caller falls back to umma_conv)
inline const char* rowconv_plan(RowPlan& plan, const RowWeights& w, const ConvParams& cp, int num_sms, bool fp16, bool fp32_out = false) {
  RowParams& p = plan.p;
  p = RowParams{};
  memset(&plan.tm, 0, sizeof plan.tm);
  if (!w.packed) return "weights not packed";
  if (cp.H % 4 || cp.H < 4) return "image height not a multiple of 4";
  const int n_xbox = (cp.W + 1 + 255) / 256;                    // TMA boxes per crop row (a box dimension is at most 256)
  const int P8 = (cp.W + 1 + 8 * n_xbox - 1) / (8 * n_xbox) * (8 * n_xbox);
  if (n_xbox > 2) return "image wider than two TMA boxes";      // W <= 503: one crop per group then (128 x 384: 4 tiles, 75 % of the lanes)
  if (cp.n_chunks != w.cin / 16) return "chunk table does not match Cin/16";
  if (!fp32_out && (cp.out_pitch % 16 || cp.out_off % 16 || reinterpret_cast<uintptr_t>(cp.out) % 32)) return "output not 32-byte aligned";
  if (cp.res && (cp.res_pitch % 8 || cp.res_off % 8 || reinterpret_cast<uintptr_t>(cp.res) % 16)) return "residual not 16-byte aligned";
  if ((long long)cp.B * cp.H * cp.W >= (1LL << 31)) return "batch too large for 32-bit pixel indices";
  // ---- K-chunks: runs of 16-channel slices contiguous in the same tensor become TMA boxes of 64 / 32 / 16 channels
  const void* chunk_base[kRowMaxChunks];
  int chunk_pitch[kRowMaxChunks];
  auto base_of = [&](int k) { return cp.chunk_ptr[k] ? cp.chunk_ptr[k] : cp.in; };
  auto pitch_of = [&](int k) { return cp.chunk_ptr[k] ? cp.chunk_pitch[k] : cp.in_pitch; };
  p.n_ks = w.cin / 16;
  for (int k = 0; k < p.n_ks;) {
    int run = 1;
    while (k + run < p.n_ks && base_of(k + run) == base_of(k) && cp.chunk_off[k + run] == cp.chunk_off[k] + 16 * run) ++run;
    int off = cp.chunk_off[k];
    if (off % 8 || pitch_of(k) % 8 || reinterpret_cast<uintptr_t>(base_of(k)) % 16) return "input not 16-byte aligned";
    for (int left = run; left > 0;) {
      const int take = left >= 4 ? 4 : left >= 2 ? 2 : 1;
      if (p.n_chunks == kRowMaxChunks) return "too many K-chunks";
      p.chunk_ch[p.n_chunks] = 16 * take;
      p.chunk_coff[p.n_chunks] = off;
      chunk_base[p.n_chunks] = base_of(k);
      chunk_pitch[p.n_chunks] = pitch_of(k);
      ++p.n_chunks;
      off += 16 * take;
      left -= take;
    }
    k += run;
  }
  // ---- crops per group: the best lane utilisation that fits TMEM (4 slots x Cout columns per tile) and shared memory (>= 2 stages)
  const int nt_max = rowconv_max_tiles(w.cout);
  const size_t misc = 128 + 32 * 4 + (kRowMaxStages + 8 * kRowMaxTiles + 2) * 8 + 3 * kMaxChunks * 16 + 64;
  const size_t budget = 227 * 1024 - 1024 - ((w.bytes + 127) & ~127u) - misc;
  // sub-rows: consecutive chunks of together <= cap channels (64 bytes per pixel, or the widest chunk)
  int cap_ch = 32;
  for (int c = 0; c < p.n_chunks; ++c) cap_ch = std::max(cap_ch, p.chunk_ch[c]);
  p.n_sub = 0;
  p.sub_c0[0] = 0;
  p.sub_step0[0] = 0;
  {
    int acc = 0, steps_so_far = 0;
    for (int c = 0; c < p.n_chunks; ++c) {
      if (acc > 0 && acc + p.chunk_ch[c] > cap_ch) {
        if (p.n_sub + 1 >= kRowMaxSub) return "too many sub-rows";
        ++p.n_sub;
        p.sub_c0[p.n_sub] = c;
        p.sub_step0[p.n_sub] = steps_so_far;
        acc = 0;
      }
      acc += p.chunk_ch[c];
      steps_so_far += 3 * (p.chunk_ch[c] / 16);
    }
    ++p.n_sub;
    p.sub_c0[p.n_sub] = p.n_chunks;
    p.sub_step0[p.n_sub] = steps_so_far;
  }
  auto entry_bytes = [&](int rows) {
    size_t e = 0;
    for (int sb = 0; sb < p.n_sub; ++sb) {
      size_t b = 0;
      for (int c = p.sub_c0[sb]; c < p.sub_c0[sb + 1]; ++c) b += ((size_t)rows * p.chunk_ch[c] * 2 + 1023) & ~(size_t)1023;
      e = std::max(e, b);
    }
    return e;
  };
  // a ring entry always holds the same sub-row (entries = a multiple of n_sub): the zero rows in front of every chunk are then never overwritten
  const size_t min_entries = (size_t)std::max(2, p.n_sub);
  double best_util = 0.0;
  int bestG = 0;
  for (int G = 1; G <= (n_xbox > 1 ? 1 : 8) && G <= std::max(1, cp.B); ++G) {
    const int lanes = G * P8, NT = (lanes + 127) / 128;
    if (NT > nt_max) break;
    const int rows = (std::max(kRowLead + NT * 128 + 1, kRowLead + lanes) + 7) / 8 * 8;
    if (min_entries * entry_bytes(rows) > budget) continue;
    const double util = (double)G * cp.W / (NT * 128.0);
    if (util > best_util - 1e-9) { best_util = util; bestG = G; }   // ties: more tiles per row hide the per-tile drain latency
  }
  if (!bestG) return "no crop grouping fits shared memory";
  p.G = bestG; p.P8 = P8;
  p.n_xbox = n_xbox; p.box_px = P8 / n_xbox;
  p.NT = (bestG * P8 + 127) / 128;
  const int rows = (std::max(kRowLead + p.NT * 128 + 1, kRowLead + bestG * P8) + 7) / 8 * 8;
  for (int sb = 0; sb < p.n_sub; ++sb) {
    uint32_t off = 0;
    p.sub_tx[sb] = 0;
    for (int c = p.sub_c0[sb]; c < p.sub_c0[sb + 1]; ++c) {
      p.chunk_smem[c] = off;
      off += (uint32_t)(((size_t)rows * p.chunk_ch[c] * 2 + 1023) & ~(size_t)1023);
      p.sub_tx[sb] += (uint32_t)(bestG * P8 * p.chunk_ch[c] * 2);
    }
  }
  p.stage_bytes = (uint32_t)entry_bytes(rows);
  // at most 4 staged rows in flight: the producer waits on step_done[.][s & 3] of the row that used an entry last, and must not fall two
  // phases behind on that barrier
  p.n_stages = (int)std::min<size_t>(std::min(kRowMaxStages, 4 * p.n_sub), budget / p.stage_bytes);
  p.n_stages = p.n_stages / p.n_sub * p.n_sub;
  p.w = w.w; p.w_bytes = w.bytes; p.bias = w.bias;
  p.out = cp.out; p.out_pitch = cp.out_pitch; p.out_off = cp.out_off;
  p.res = cp.res; p.res_pitch = cp.res_pitch; p.res_off = cp.res_off;
  p.B = cp.B; p.H = cp.H; p.W = cp.W;
  p.n_groups = (cp.B + bestG - 1) / bestG;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("LPSR_UMMA_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.debug = dbg;
    static int tr_on = -1;
    if (tr_on < 0) { const char* e = getenv("LPSR_ROW_TRACE"); tr_on = (e && e[0] == '1') ? 1 : 0; }
    p.trace = nullptr;
    if (tr_on) {
      long long*& buf = umma_trace_buffer();
      if (!buf) cudaMalloc(&buf, 512 * 8 * sizeof(long long));
      if (buf) { cudaMemsetAsync(buf, 0, 512 * 8 * sizeof(long long)); p.trace = buf; }
    }
  }
  const long long total_u = (long long)p.n_groups * (cp.H / 4);
  // small batches stay on the tile kernels of umma_conv.cuh: a CTA here walks its rows one after the other (~1 us per row), so with fewer than ~16
  // rows per SM the launch is latency bound (measured: B = 1 at 32 x 192, 250 us per forward against 190 us; B = 3: 210 against 76 us per plate)
  const char* mu = getenv("LPSR_ROWCONV_MIN_UNITS");            // read per call: the tests force the kernel on small shapes with 0
  const int min_units = mu ? atoi(mu) : 4;
  if (total_u < (long long)min_units * num_sms) return "too little work for the row-streaming kernel";
  plan.grid = (int)std::min<long long>(num_sms, total_u);
  plan.threads = (kRowEpiWarp0 + 4 * p.NT) * 32;
  plan.smem_bytes = 1024 + (size_t)p.n_stages * p.stage_bytes + ((w.bytes + 127) & ~127u) + misc;
  PFN_lpsr_tmapEncodeTiled enc = umma_encode_fn();
  if (!enc) return "cuTensorMapEncodeTiled entry point not found";
  for (int c = 0; c < p.n_chunks; ++c) {
    const int ch = p.chunk_ch[c], pitch = chunk_pitch[c];
    const CUtensorMapSwizzle sw = ch == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : ch == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    const cuuint64_t gdim[4] = {(cuuint64_t)pitch, (cuuint64_t)cp.W, (cuuint64_t)cp.H, (cuuint64_t)cp.B};
    const cuuint64_t gstr[3] = {(cuuint64_t)pitch * 2, (cuuint64_t)cp.W * pitch * 2, (cuuint64_t)cp.H * cp.W * pitch * 2};
    const cuuint32_t box[4] = {(cuuint32_t)ch, (cuuint32_t)(P8 / n_xbox), 1, (cuuint32_t)bestG};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    const CUresult r = enc(&plan.tm.m[c], fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(chunk_base[c]),
                           gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return "cuTensorMapEncodeTiled failed";
  }
  return nullptr;
}

template <typename T, int COUT, int EPI>
inline const char* rowconv_launch_inst(const RowPlan& plan, cudaStream_t st) {
  static bool configured[kMaxDevices] = {};
  bool* flag = func_configured_flag(configured);
  if (!flag || !*flag) {
    cudaError_t e = cudaFuncSetAttribute(rowconv_kernel<T, COUT, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    if (flag) *flag = true;
  }
  cudaError_t e = launch_pdl(rowconv_kernel<T, COUT, EPI>, dim3(plan.grid), dim3(plan.threads), plan.smem_bytes, st, plan.p, plan.tm);
  if (e == cudaSuccess) e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

// "unhandled": the shape / epilogue is not covered (fall back); any other non-null string is a launch error
template <typename T>
inline const char* rowconv_launch(const RowWeights& w, const ConvParams& cp, int num_sms, cudaStream_t st, bool* handled, bool final_sigmoid = false) {
  *handled = false;
  if constexpr (sizeof(T) != 2) {
    return nullptr;
  } else {
    if (!rowconv_enabled() || !w.packed) return nullptr;
    if (cp.relu && cp.res) return nullptr;
    RowPlan plan;
    if (rowconv_plan(plan, w, cp, num_sms, !IsBf16<T>::value, final_sigmoid)) return nullptr;
    *handled = true;
    if (final_sigmoid) {
      if (w.cout == 16 && !cp.relu && !cp.res) return rowconv_launch_inst<T, 16, kEpiFinalSigmoid>(plan, st);
      *handled = false;
      return nullptr;
    }
    if (w.cout == 16) {
      if (cp.relu) return rowconv_launch_inst<T, 16, kEpiRelu>(plan, st);
      if (!cp.res) return rowconv_launch_inst<T, 16, kEpiPlain>(plan, st);
    } else if (w.cout == 32) {
      if (cp.relu) return rowconv_launch_inst<T, 32, kEpiRelu>(plan, st);
      if (cp.res) return rowconv_launch_inst<T, 32, kEpiResidual>(plan, st);
      return rowconv_launch_inst<T, 32, kEpiPlain>(plan, st);
    }
    *handled = false;
    return nullptr;
  }
}

}  // namespace lpsr
