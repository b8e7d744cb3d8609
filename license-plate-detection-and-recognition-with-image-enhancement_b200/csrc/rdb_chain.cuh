// rdb_chain.cuh -- one launch per residual dense block (reference my_models/lpsr.py:43-61): the four dense 3x3 layers, the 1x1
// local feature fusion, alpha and the residual, with the growing concatenation kept OUT of HBM.  sm_100a only, 16-bit modes.
//
// Layer by layer an RDB moves 289 channel-planes per pixel through HBM (x is read 5 times, g0 4 times, ...) for 64 algorithmic ones.
// Shared memory cannot hold a useful tile of the 96-channel concatenation next to 80 KB of weights, so the intermediate growths g0..g2
// live in a small PER-CTA scratch in global memory that is rewritten for every band and therefore stays resident in the 126 MB L2:
//   * a CTA owns a BAND of one crop (Rb rows x Cb columns, default the full height x 32 columns) and runs the four layers on it back to
//     back; layer l produces its 16 channels on the band grown by the halo the later layers still need (3 - l pixels per side; none along
//     a dimension the band spans completely, where the TMA's out-of-bounds zero fill of the scratch tensor IS the layer's zero padding);
//   * layers 0..2 store ReLU(conv) into the scratch (zeros where the halo falls outside the image: every layer is zero padded,
//     lpsr.py:34), the next layer's TMA producer waits on an mbarrier per (layer, item) that the epilogue warps arrive on after their
//     stores (generic -> async proxy fence), and loads x from the block input and g0.. from the scratch as ordinary haloed boxes;
//   * layer 3 is the fused "last dense layer + lff + alpha + residual" of umma_conv.cuh (kConv3x3FoldLff) and writes the block output.
// HBM sees x once (plus the horizontal halo of the band) and the output once; everything else is L2 traffic.  The MMA / epilogue
// machinery (dx folded into N = 48, warp-shuffle shifted sum, one MMA warp and one epilogue group per TMEM accumulator, TMA-swizzled
// operands addressed by descriptor start address per tap) is the one of umma_conv_kernel; items of the four layers simply follow each
// other through the same shared-memory ring, so the pipeline never drains between layers.
#pragma once
#include "umma_conv.cuh"

namespace lpsr {

constexpr int kChainLayers = 4;
constexpr int kChainMaxItems = 48;          // items per (band, layer)
constexpr int kChainAccCols = kLffCols;     // TMEM columns per accumulator (layer 3 needs 48 + 32)
constexpr int kChainStepsPerLayer = 8;

struct ChainLayer {
  int hy, hx;                  // halo of this layer's output region around the band core (rows / columns)
  int Hv, Wv;                  // output region of the layer = the "virtual image" its tiles walk
  int pitch;                   // Wv + 2: one strip with one halo column per side
  int k, k_last, n_items;      // tiles per item, tiles of the last item, items per band
  int rbox, npx;               // TMA box rows, staged slots per item
  int n_ks;                    // K-slices of 16 channels: 2 (x) + l
  uint32_t chunk_smem[4];      // byte offsets of the boxes {x, g0, g1, g2} inside an item buffer (1024-aligned)
  uint32_t w_off, w_bytes;     // this layer's packed weights inside the shared-memory weight block
  uint32_t item_bytes;         // TMA transaction bytes per item
  int done0;                   // index of the layer's first "item stored" barrier (layers 0..2)
  signed char dep[kChainMaxItems];   // last item of layer l-1 whose stores item j's loads need (-1: none)
};

struct ChainParams {
  ChainLayer L[kChainLayers];
  int B, H, W;                 // crops (padded size)
  int Rb, Cb, bands_y, bands_x;
  int n_bands;
  const uint16_t* w[kChainLayers];
  const float* bias[kChainLayers];    // [16] for layers 0..2; layer 3: [16 | 32] (b3, alpha * lff bias)
  void* out;                   // block output, dense NHWC [B][H][W][32]
  void* scr[3];                // g0, g1, g2: [gridDim.x][Hv][Wv][16]
  uint32_t buf_bytes;
  int n_bufs;
  uint32_t w_total;            // bytes of the weight block (multiple of 128)
  int n_done;
  int debug;                   // bit 0: skip the gpu-scope fence before publishing scratch stores (experiment)
};

struct ChainTmaps {
  CUtensorMap x[kChainLayers];   // block input, box [rbox_l][pitch_l][32]
  CUtensorMap g[6];              // consumer layer l, producer m < l: index l*(l-1)/2 + m, box [rbox_l][pitch_l][16]
};

template <typename T>
__global__ void __launch_bounds__(kUmmaThreads, 1) rdb_chain_kernel(const __grid_constant__ ChainParams p, const __grid_constant__ ChainTmaps tm) {
  static_assert(sizeof(T) == 2, "16-bit operands");
  constexpr int G = kEpiGroups, NOUT = 16, CH = 16;
  constexpr int XCH = G * 2 * 4 * 2 * NOUT;                      // floats of warp-boundary exchange (folded epilogue)
  constexpr uint32_t kTmemCols = (G * kChainAccCols <= 256) ? 256 : 512;
  static_assert(G * kChainAccCols <= 512, "accumulators exceed TMEM");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R = p.n_bufs;
  const uint32_t buf_bytes = p.buf_bytes;
  uint8_t* a_smem = smem;
  uint8_t* w_smem = smem + (size_t)R * buf_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_smem + p.w_total);
  const uint32_t bar0 = ptx::smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (R + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * R + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * R + G + a); };
  auto a2full_bar = [&](int a) { return bar0 + 8u * (2 * R + 2 * G + a); };
  auto tfull2_bar = [&](int a) { return bar0 + 8u * (2 * R + 3 * G + a); };
  auto done_bar = [&](int i) { return bar0 + 8u * (2 * R + 4 * G + i); };
  const int n_bars = 2 * R + 4 * G + p.n_done;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + n_bars);
  float* xchg = reinterpret_cast<float*>(bars + ((n_bars + 2) & ~1));               // 16-byte aligned
  uint4* steps = reinterpret_cast<uint4*>(xchg + XCH);                               // [layer][K-slice]
  int* slot_base_s = reinterpret_cast<int*>(steps + kChainLayers * kChainStepsPerLayer);
  uint8_t* a2_all = reinterpret_cast<uint8_t*>(slot_base_s + 8);                     // per group: g3 as a K = 16 operand (4 KB)
  uint8_t* ones_s = a2_all + (size_t)G * 4096;                                       // [2 planes][128 rows][8]: k = 0, 1 are 1.0
  uint8_t* lffb_s = ones_s + 2 * 128 * 16;                                           // [2][80][8] biases {hi, lo}
  uint8_t* lffid_s = lffb_s + 2 * kLffCols * 16;                                     // [4][32][8] identity (residual through the tensor core)

  // ---- one-time setup: all four layers' weights, the fused layer's constant operands, barriers, TMEM ------------------------------
  for (int l = 0; l < kChainLayers; ++l) {
    const uint4* src = reinterpret_cast<const uint4*>(p.w[l]);
    uint4* dst = reinterpret_cast<uint4*>(w_smem + p.L[l].w_off);
    for (uint32_t i = threadIdx.x; i < p.L[l].w_bytes / 16; i += kUmmaThreads) dst[i] = __ldg(src + i);
  }
  {
    T* ones = reinterpret_cast<T*>(ones_s);
    for (uint32_t i = threadIdx.x; i < 2 * 128 * 8; i += kUmmaThreads) ones[i] = from_f32<T>((i < 128 * 8 && (i & 7) < 2) ? 1.f : 0.f);
    T* bb = reinterpret_cast<T*>(lffb_s);
    const float* bias3 = p.bias[3];
    for (uint32_t i = threadIdx.x; i < 2 * kLffCols * 8; i += kUmmaThreads) {
      const uint32_t k = i & 7, col = (i >> 3) % kLffCols, plane = (i >> 3) / kLffCols;
      float v = 0.f;
      if (plane == 0 && k < 2 && ((col >= 16 && col < 32) || col >= 48)) {
        const float bv = __ldg(bias3 + (col < 32 ? col - 16 : col - 32));              // [b3 (16) | lff bias (32)]
        const float hi = to_f32<T>(from_f32<T>(bv));
        v = k == 0 ? hi : bv - hi;
      }
      bb[i] = from_f32<T>(v);
    }
    T* idm = reinterpret_cast<T*>(lffid_s);
    for (uint32_t i = threadIdx.x; i < 4 * 32 * 8; i += kUmmaThreads) idm[i] = from_f32<T>(((i >> 8) * 8 + (i & 7)) == ((i >> 3) & 31) ? 1.f : 0.f);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < R; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), G);
    }
    for (int a = 0; a < G; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);
      ptx::mbar_init(tempty_bar(a), 4);
      ptx::mbar_init(a2full_bar(a), 4);
      ptx::mbar_init(tfull2_bar(a), 1);
    }
    for (int l = 0; l < kChainLayers - 1; ++l)
      for (int j = 0; j < p.L[l].n_items; ++j)                                        // one arrival per epilogue warp and tile of the item
        ptx::mbar_init(done_bar(p.L[l].done0 + j), 4u * (uint32_t)(j == p.L[l].n_items - 1 ? p.L[l].k_last : p.L[l].k));
    ptx::fence_mbar_init();
  }
  if (warp == kUmmaMmaWarp) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), kTmemCols);
    ptx::tmem_relinquish();
    if (lane == 0) {
      for (int l = 0; l < kChainLayers; ++l) {
        const ChainLayer& Ly = p.L[l];
        uint4* st = steps + l * kChainStepsPerLayer;
        // K-slices 0, 1: the 32 channels of x (64-byte rows, SWIZZLE_64B); K-slice 2 + m: g_m (32-byte rows, SWIZZLE_32B)
        for (int kk = 0; kk < 2; ++kk)
          st[kk] = make_uint4((Ly.chunk_smem[0] >> 4) + 2u * (uint32_t)kk, 4u, umma_desc_hi_swizzled(64), (uint32_t)Ly.pitch * 4u);
        for (int m = 0; m < l; ++m) st[2 + m] = make_uint4(Ly.chunk_smem[1 + m] >> 4, 2u, umma_desc_hi_swizzled(32), (uint32_t)Ly.pitch * 2u);
      }
    }
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::griddep_wait();

  const int n_my_bands = (p.n_bands - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int bands_per_crop = p.bands_y * p.bands_x;

  if (warp >= kUmmaFirstLoaderWarp) {
    // =================================== TMA producer ==============================================
    if (warp == kUmmaFirstLoaderWarp && ptx::elect_one()) {
      for (int l = 0; l < kChainLayers; ++l) ptx::prefetch_tmap(&tm.x[l]);
      for (int i = 0; i < 6; ++i) ptx::prefetch_tmap(&tm.g[i]);
      int buf = 0;
      uint32_t ph = 0;
      for (int bi = 0; bi < n_my_bands; ++bi) {
        const int band = blockIdx.x + bi * gridDim.x;
        const int n = band / bands_per_crop, rem = band - n * bands_per_crop;
        const int by0 = (rem / p.bands_x) * p.Rb, bx0 = (rem % p.bands_x) * p.Cb;
        const uint32_t band_par = (uint32_t)bi & 1u;
        for (int l = 0; l < kChainLayers; ++l) {
          const ChainLayer& Ly = p.L[l];
          const int rows_per_item = Ly.k * 126;
          int waited = 0;                                         // items of layer l-1 known to be stored
          for (int j = 0; j < Ly.n_items; ++j, (++buf == R ? (buf = 0, ph ^= 1u) : 0u)) {
            if (l > 0) {
              for (; waited <= (int)Ly.dep[j]; ++waited) ptx::mbar_wait(done_bar(p.L[l - 1].done0 + waited), band_par);
            }
            ptx::mbar_wait(empty_bar(buf), ph ^ 1u);
            const uint32_t dst0 = ptx::smem_u32(a_smem + (size_t)buf * buf_bytes);
            const uint32_t bar = full_bar(buf);
            const int qlo = j * rows_per_item - (Ly.pitch + 1);   // first linear position any tap of this item reads
            const int y_lo = (qlo + 2 * Ly.pitch) / Ly.pitch - 2; // floor(qlo / pitch), qlo >= -(pitch + 1)
            slot_base_s[buf] = qlo - y_lo * Ly.pitch;
            // block input: image coordinates of the layer's virtual pixel (0, 0) are (by0 - hy, bx0 - hx); strip halo column: -1
            ptx::tma_load_4d(dst0 + Ly.chunk_smem[0], &tm.x[l], bar, 0, bx0 - Ly.hx - 1, by0 - Ly.hy + y_lo, n);
            for (int m = 0; m < l; ++m) {                         // g_m from this CTA's scratch: its region is (hy_m - hy_l) larger per side
              const int dy = p.L[m].hy - Ly.hy, dx = p.L[m].hx - Ly.hx;
              ptx::tma_load_4d(dst0 + Ly.chunk_smem[1 + m], &tm.g[l * (l - 1) / 2 + m], bar, 0, dx - 1, y_lo + dy, (int)blockIdx.x);
            }
            ptx::mbar_arrive_expect_tx(bar, Ly.item_bytes);
          }
        }
      }
    }
  } else if (warp >= kUmmaMmaWarp) {
    // =================================== MMA issuers ===============================================
    const int mg = warp - kUmmaMmaWarp;
    const bool leader = ptx::elect_one();
    constexpr uint32_t idesc48 = umma_idesc_f16(IsBf16<T>::value, 48), idesc80 = umma_idesc_f16(IsBf16<T>::value, kLffCols),
                       idesc32 = umma_idesc_f16(IsBf16<T>::value, kLffN);
    const uint32_t a2_lo = umma_desc_lo(ptx::smem_u32(a2_all) + (uint32_t)mg * 4096u, 2048u);
    const uint32_t ones_lo = umma_desc_lo(ptx::smem_u32(ones_s), 2048u), lffb_lo = umma_desc_lo(ptx::smem_u32(lffb_s), (uint32_t)kLffCols * 16),
                   lffid_lo = umma_desc_lo(ptx::smem_u32(lffid_s), 32u * 16);
    const uint32_t a_smem16 = ptx::smem_u32(a_smem) >> 4, buf16_sz = buf_bytes >> 4;
    __syncwarp();
    const uint32_t acc = (uint32_t)mg;
    uint32_t acc_par = 1, a2_par = 0;
    int turn = 0, buf = 0;
    uint32_t buf_par = 0;
    for (int bi = 0; bi < n_my_bands; ++bi) {
      for (int l = 0; l < kChainLayers; ++l) {
        const ChainLayer& Ly = p.L[l];
        const bool last = (l == kChainLayers - 1);
        const uint32_t NM = last ? (uint32_t)kLffCols : 48u;       // weight rows per K core-matrix
        const uint32_t w_addr = ptx::smem_u32(w_smem) + Ly.w_off;
        const uint32_t w_lo = umma_desc_lo(w_addr, NM * 16);
        const uint32_t cgn = (uint32_t)Ly.n_ks * 2u * NM;          // 16-byte units between dy taps
        const uint32_t w2_lo = umma_desc_lo(w_addr + 3u * cgn * 16u, (uint32_t)kLffN * 16);   // fused layer: lff's g3 slice follows the taps
        const uint4* st = steps + l * kChainStepsPerLayer;
        const int n_ks = Ly.n_ks;
        for (int j = 0; j < Ly.n_items; ++j) {
          const int kt = (j == Ly.n_items - 1) ? Ly.k_last : Ly.k;
          ptx::mbar_wait(full_bar(buf), buf_par);
          ptx::tc_fence_after();
          uint32_t slot = (uint32_t)slot_base_s[buf];
          const uint32_t buf16 = a_smem16 + (uint32_t)buf * buf16_sz;
          for (int m = 0; m < kt; ++m, slot += 126u) {
            const bool mine = (turn == mg);
            if (++turn == G) turn = 0;
            if (!mine) continue;
            ptx::mbar_wait(tempty_bar(acc), acc_par);
            ptx::tc_fence_after();
            if (leader) {
              const uint32_t d = tmem_base + acc * kChainAccCols;
              uint32_t b_lo = w_lo;
              if (!last) {
#pragma unroll 1
                for (int ks = 0; ks < n_ks; ++ks) {
                  const uint4 e = st[ks];
                  const uint32_t a0 = (buf16 + e.x + slot * e.y) | (1u << 16);
#pragma unroll
                  for (int t = 0; t < 3; ++t)
                    ptx::tc_mma_f16_lohi(d, a0 + (uint32_t)t * e.w, e.z, b_lo + (uint32_t)t * cgn, kUmmaDescHi, idesc48, (uint32_t)(ks | t));
                  b_lo += 2 * 48;
                }
              } else {
                ptx::tc_mma_f16_lohi(d, ones_lo, kUmmaDescHi, lffb_lo, kUmmaDescHi, idesc80, 0u);   // accumulator := biases
#pragma unroll 1
                for (int ks = 0; ks < n_ks; ++ks) {
                  const uint4 e = st[ks];
                  const uint32_t a0 = (buf16 + e.x + slot * e.y) | (1u << 16);
                  if (ks < 2)   // + x (centre tap) times the identity -> lff columns: the residual
                    ptx::tc_mma_f16_lohi(d + 48, a0 + e.w, e.z, lffid_lo + (uint32_t)(ks * 2 * 32), kUmmaDescHi, idesc32, 1u);
#pragma unroll
                  for (int t = 0; t < 3; ++t) {
                    const uint32_t dyv = t == 0 ? 1u : (t == 1 ? 0u : 2u);   // dy = 1 first: its 80 columns carry lff over this K-slice
                    ptx::tc_mma_f16_lohi(d, a0 + dyv * e.w, e.z, b_lo + dyv * cgn, kUmmaDescHi, t == 0 ? idesc80 : idesc48, 1u);
                  }
                  b_lo += 2 * kLffCols;
                }
              }
              ptx::tc_commit(tfull_bar(acc));
            }
            __syncwarp();
            if (last) {
              ptx::mbar_wait(a2full_bar(acc), a2_par);              // the epilogue group wrote g3 as a K = 16 operand
              a2_par ^= 1u;
              ptx::tc_fence_after();
              if (leader) {
                ptx::tc_mma_f16_lohi(tmem_base + acc * kChainAccCols + 48, a2_lo, kUmmaDescHi, w2_lo, kUmmaDescHi, idesc32, 1u);
                ptx::tc_commit(tfull2_bar(acc));
              }
              __syncwarp();
            }
            acc_par ^= 1u;
          }
          if (leader) ptx::tc_commit(empty_bar(buf));
          __syncwarp();
          if (++buf == R) { buf = 0; buf_par ^= 1u; }
        }
      }
    }
  } else {
    // =================================== epilogue groups ============================================
    const int grp = warp >> 2, wq = warp & 3;
    const int row = wq * 32 + lane;
    float* xg = xchg + (size_t)grp * (2 * 4 * 2 * NOUT);
    const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(grp * kChainAccCols);
    const int Himg = p.H, Wimg = p.W;
    T* out = static_cast<T*>(p.out);
    uint32_t my_par = 0;                                          // parity of tmem_full: flips with every tile of this group
    uint32_t lff_par = 0;                                         // parity of tmem_full2 (lff accumulator complete): fused-layer tiles only
    int t0mod = 0;
    for (int bi = 0; bi < n_my_bands; ++bi) {
      const int band = blockIdx.x + bi * gridDim.x;
      const int n = band / bands_per_crop, rem = band - n * bands_per_crop;
      const int by0 = (rem / p.bands_x) * p.Rb, bx0 = (rem % p.bands_x) * p.Cb;
      for (int l = 0; l < kChainLayers; ++l) {
        const ChainLayer& Ly = p.L[l];
        const bool last = (l == kChainLayers - 1);
        const int pitch = Ly.pitch, Hv = Ly.Hv, Wv = Ly.Wv;
        const int iy0 = by0 - Ly.hy, ix0 = bx0 - Ly.hx - 1;            // image coordinates of virtual (row 0, strip column 0)
        const int adv_y = 126 / pitch, adv_x = 126 - adv_y * pitch;
        const int advg_y = (G * 126) / pitch, advg_x = G * 126 - advg_y * pitch;
        T* scr = last ? nullptr : static_cast<T*>(p.scr[l]) + (size_t)blockIdx.x * Hv * Wv * NOUT;
        float bias_r[NOUT];
        if (!last) {
#pragma unroll
          for (int c = 0; c < NOUT; ++c) bias_r[c] = __ldg(p.bias[l] + c);
        }
        for (int j = 0; j < Ly.n_items; ++j) {
          const int kt = (j == Ly.n_items - 1) ? Ly.k_last : Ly.k;
          const int q = j * Ly.k * 126 - 1 + row;                      // linear position of this thread's row in tile 0 (>= -1)
          int y = (q + 2 * pitch) / pitch - 2;
          int xs = q - y * pitch;
          int m = grp - t0mod;
          if (m < 0) m += G;
          t0mod += kt % G;
          if (t0mod >= G) t0mod -= G;
          for (int s = 0; s < m; ++s) {
            xs += adv_x;
            y += adv_y;
            if (xs >= pitch) { xs -= pitch; ++y; }
          }
          for (; m < kt; m += G) {
            const bool in_tile = row >= 1 && row <= 126 && (unsigned)y < (unsigned)Hv && xs >= 1 && xs <= Wv;
            const int iy = iy0 + y, ix = ix0 + xs;
            const bool in_img = (unsigned)iy < (unsigned)Himg && (unsigned)ix < (unsigned)Wimg;
            const int vy = y, vx = xs - 1;
            xs += advg_x;
            y += advg_y;
            if (xs >= pitch) { xs -= pitch; ++y; }
            ptx::mbar_wait(tfull_bar(grp), my_par);
            ptx::tc_fence_after();
            float* xb = xg + (size_t)my_par * (4 * 2 * NOUT);
            my_par ^= 1u;
            float v[CH], lf[CH], rg[CH];
            ptx::tc_ld16_nowait(taddr, lf);
            ptx::tc_ld16_nowait(taddr + NOUT, v);
            ptx::tc_ld16_nowait(taddr + 2 * NOUT, rg);
            ptx::tc_wait_ld();
            if (!last) {
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(tempty_bar(grp));       // accumulator is in registers
            }
            // out[q] = D[q-1, dx=0] + D[q, dx=1] + D[q+1, dx=2]; warp boundaries through a small exchange
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
            if (!last) {
              // g_l = relu(conv + bias) inside the image, 0 in the part of the halo that lies outside it (the next layer's zero padding)
#pragma unroll
              for (int c = 0; c < CH; ++c) v[c] = in_img ? v[c] + bias_r[c] : 0.f;
              store_chunk16<T, true>(scr, NOUT, 0, in_tile ? vy * Wv + vx : -1, v);
              // publish: the next layer's TMA (async proxy) reads these generic-proxy stores through L2
              if (!(p.debug & 1)) __threadfence();
              asm volatile("fence.proxy.async;" ::: "memory");
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(done_bar(Ly.done0 + j));
            } else {
              // ---- stage 1: g3 = relu(conv + b3) (bias already in the accumulator) becomes the K = 16 operand of lff's last slice
              {
                uint8_t* a2 = a2_all + (size_t)grp * 4096;
                *reinterpret_cast<uint4*>(a2 + row * 16) =
                    make_uint4(pack2<T, true>(v[0], v[1]), pack2<T, true>(v[2], v[3]), pack2<T, true>(v[4], v[5]), pack2<T, true>(v[6], v[7]));
                *reinterpret_cast<uint4*>(a2 + 2048 + row * 16) =
                    make_uint4(pack2<T, true>(v[8], v[9]), pack2<T, true>(v[10], v[11]), pack2<T, true>(v[12], v[13]), pack2<T, true>(v[14], v[15]));
              }
              ptx::fence_proxy_async();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(a2full_bar(grp));
              // ---- stage 2: columns 48..79 = x + alpha * lff(cat[x, g0..g3]) -> block output
              ptx::mbar_wait(tfull2_bar(grp), lff_par);
              lff_par ^= 1u;
              ptx::tc_fence_after();
              float o[kLffN];
#pragma unroll
              for (int hh = 0; hh < kLffN / CH; ++hh) ptx::tc_ld16_nowait(taddr + 48 + hh * CH, o + hh * CH);
              ptx::tc_wait_ld();
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(tempty_bar(grp));
              const int pix = (in_tile && in_img) ? (n * Himg + iy) * Wimg + ix : -1;
#pragma unroll
              for (int hh = 0; hh < kLffN / CH; ++hh) {
                float oc[CH];
#pragma unroll
                for (int c = 0; c < CH; ++c) oc[c] = o[hh * CH + c];
                store_chunk16<T>(out, kLffN, hh * CH, pix, oc);
              }
            }
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kUmmaMmaWarp) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------------------------------
// host: geometry, plan, launch
// ---------------------------------------------------------------------------------------------------
struct ChainGeom {
  bool ok = false;
  int Rb = 0, Cb = 0, bands_y = 0, bands_x = 0;
  int hy[kChainLayers], hx[kChainLayers], Hv[kChainLayers], Wv[kChainLayers];
  size_t scr_off[3];             // byte offsets of g0, g1, g2 inside the scratch (for `ctas` CTAs)
  size_t scr_bytes = 0;
};

inline int chain_env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && e[0]) ? atoi(e) : dflt;
}
inline bool chain_enabled() {
  static int v = -1;
  if (v < 0) v = chain_env_int("LPSR_RDB_CHAIN", 1) ? 1 : 0;
  return v == 1;
}

// Bands of one crop: the full height when it fits (<= LPSR_CHAIN_RB rows, default 64), LPSR_CHAIN_CB (default 32) columns.
inline ChainGeom chain_geometry(int H, int W, int ctas) {
  ChainGeom g;
  static int rb_max = 0, cb_req = 0;
  if (!rb_max) { rb_max = std::max(8, chain_env_int("LPSR_CHAIN_RB", 64)); cb_req = std::max(8, chain_env_int("LPSR_CHAIN_CB", 32)); }
  g.bands_y = (H + rb_max - 1) / rb_max;
  g.Rb = (H + g.bands_y - 1) / g.bands_y;
  g.bands_x = (W + cb_req - 1) / cb_req;
  g.Cb = (W + g.bands_x - 1) / g.bands_x;
  size_t off = 0;
  for (int l = 0; l < kChainLayers; ++l) {
    g.hy[l] = g.bands_y == 1 ? 0 : kChainLayers - 1 - l;
    g.hx[l] = g.bands_x == 1 ? 0 : kChainLayers - 1 - l;
    g.Hv[l] = g.Rb + 2 * g.hy[l];
    g.Wv[l] = g.Cb + 2 * g.hx[l];
    if (l < 3) {
      g.scr_off[l] = off;
      off += ((size_t)ctas * g.Hv[l] * g.Wv[l] * 16 * 2 + 255) & ~(size_t)255;
    }
  }
  g.scr_bytes = off;
  g.ok = g.Wv[0] + 2 <= 256;
  return g;
}

struct ChainPlan {
  ChainParams p;
  ChainTmaps tm;
  size_t smem_bytes;
  int grid;
};

inline size_t chain_fixed_smem(uint32_t w_total, int n_done) {
  const int G = kEpiGroups;
  return (size_t)w_total + (size_t)(2 * kUmmaMaxBufs + 4 * G + n_done + 4) * 8 + (size_t)G * 2 * 4 * 2 * 16 * 4 /*xchg*/ +
         kChainLayers * kChainStepsPerLayer * 16 + 64 /*slot_base*/ + (size_t)G * 4096 + 2 * 128 * 16 + 2 * kLffCols * 16 + 4 * 32 * 16 + 1024 /*alignment*/ + 256;
}

// x: block input (dense 32 ch); w[l]: the packed weights of the dense layers (l < 3: folded Cout = 16; l = 3: the fused layer + lff)
inline const char* rdb_chain_plan(ChainPlan& plan, const UmmaWeights* const w[kChainLayers], const void* x, void* out, void* scratch, int B, int H,
                                  int W, int num_sms, bool fp16) {
  ChainParams& p = plan.p;
  p = ChainParams{};
  memset(&plan.tm, 0, sizeof plan.tm);
  if ((long long)B * H * W >= (1LL << 31)) return "batch too large for 32-bit pixel indices";
  if (reinterpret_cast<uintptr_t>(x) % 16 || reinterpret_cast<uintptr_t>(out) % 32 || reinterpret_cast<uintptr_t>(scratch) % 256) return "chain: misaligned tensors";
  const ChainGeom g = chain_geometry(H, W, num_sms);
  if (!g.ok) return "chain: band too wide";
  p.B = B; p.H = H; p.W = W;
  p.Rb = g.Rb; p.Cb = g.Cb; p.bands_y = g.bands_y; p.bands_x = g.bands_x;
  const long long n_bands = (long long)B * g.bands_y * g.bands_x;
  if (n_bands >= (1LL << 30)) return "chain: too many bands";
  p.n_bands = (int)n_bands;
  plan.grid = (int)std::min<long long>(n_bands, num_sms);
  p.out = out;
  for (int m = 0; m < 3; ++m) p.scr[m] = static_cast<char*>(scratch) + g.scr_off[m];
  uint32_t w_off = 0;
  for (int l = 0; l < kChainLayers; ++l) {
    const UmmaWeights& u = *w[l];
    const int cin = 32 + 16 * l;
    if (!u.packed || u.ks != 3 || u.cout != 16 || u.cin != cin || u.fused_lff != (l == 3) || u.wsplit) return "chain: unexpected layer weights";
    p.w[l] = u.w; p.bias[l] = u.bias;
    p.L[l].w_off = w_off;
    p.L[l].w_bytes = (uint32_t)(3 * cin * (l == 3 ? kLffCols : 48) * 2 + (l == 3 ? 2 * kLffN * 16 : 0));
    w_off += (p.L[l].w_bytes + 127u) & ~127u;
  }
  p.w_total = w_off;
  // ---- per-layer tiling: one strip (pitch = Wv + 2), items of k tiles of 126 positions
  // first pass with the worst-case barrier count to find the buffer capacity
  const size_t fixed = chain_fixed_smem(p.w_total, 3 * kChainMaxItems);
  if (fixed + 2 * 16 * 1024 > 227 * 1024) return "chain: weights leave no room for item buffers";
  const int n_bufs = 2;
  const size_t buf_cap = ((227 * 1024 - fixed) / n_bufs) & ~(size_t)1023;
  size_t buf_bytes = 0;
  int done = 0;
  for (int l = 0; l < kChainLayers; ++l) {
    ChainLayer& Ly = p.L[l];
    Ly.hy = g.hy[l]; Ly.hx = g.hx[l]; Ly.Hv = g.Hv[l]; Ly.Wv = g.Wv[l];
    Ly.pitch = Ly.Wv + 2;
    Ly.n_ks = 2 + l;
    const int total_tiles = (Ly.Hv * Ly.pitch + 125) / 126;
    auto item_bytes = [&](int k, int& rbox, uint32_t* chunk_smem) {
      rbox = (k * 126 + 3 * Ly.pitch + 1 + Ly.pitch - 1) / Ly.pitch;
      const size_t npx = (size_t)rbox * Ly.pitch;
      size_t off = 0;
      for (int c = 0; c <= l; ++c) {
        if (chunk_smem) chunk_smem[c] = (uint32_t)off;
        off += (npx * (c == 0 ? 64 : 32) + 1023) & ~(size_t)1023;
      }
      return off;
    };
    int k = 0, rbox = 0;
    for (int kk = 1; kk <= std::min(kUmmaMaxK, total_tiles); ++kk) {
      int rb;
      if (item_bytes(kk, rb, nullptr) > buf_cap || rb > 256) break;
      k = kk;
    }
    if (!k) return "chain: one tile does not fit the item buffer";
    int n_items = (total_tiles + k - 1) / k;
    k = (total_tiles + n_items - 1) / n_items;                   // balance the items
    n_items = (total_tiles + k - 1) / k;
    if (n_items > kChainMaxItems) return "chain: too many items per band";
    Ly.k = k; Ly.n_items = n_items; Ly.k_last = total_tiles - (n_items - 1) * k;
    const size_t ib = item_bytes(k, rbox, Ly.chunk_smem);
    Ly.rbox = rbox; Ly.npx = rbox * Ly.pitch;
    Ly.item_bytes = (uint32_t)((size_t)Ly.npx * (64 + 32 * l));
    buf_bytes = std::max(buf_bytes, ib);
    Ly.done0 = done;
    if (l < 3) done += n_items;
    // dependencies on layer l-1: item j stages rows [y_lo, y_lo + rbox) of its own region
    for (int j = 0; j < n_items; ++j) {
      Ly.dep[j] = -1;
      if (l == 0) continue;
      const ChainLayer& Lp = p.L[l - 1];
      const int qlo = j * k * 126 - (Ly.pitch + 1);
      const int y_lo = (qlo + 2 * Ly.pitch) / Ly.pitch - 2;
      int r_max = y_lo + rbox - 1 + (Lp.hy - Ly.hy);               // last row of layer l-1's region the box touches
      if (r_max < 0) continue;
      r_max = std::min(r_max, Lp.Hv - 1);
      const int q_last = r_max * Lp.pitch + Lp.Wv;                  // linear position of the row's last pixel (strip column Wv)
      const int item = std::min((q_last / 126) / Lp.k, Lp.n_items - 1);
      Ly.dep[j] = (signed char)item;
    }
  }
  p.n_done = done;
  p.buf_bytes = (uint32_t)buf_bytes;
  p.n_bufs = n_bufs;
  {
    static int dbg = -1;
    if (dbg < 0) dbg = chain_env_int("LPSR_CHAIN_DEBUG", 0);
    p.debug = dbg;
  }
  plan.smem_bytes = (size_t)n_bufs * buf_bytes + chain_fixed_smem(p.w_total, done);
  if (plan.smem_bytes > 227 * 1024) return "chain: shared memory budget exceeded";
  for (int l = 0; l < kChainLayers; ++l) {
    const ChainLayer& Ly = p.L[l];
    if (const char* msg = umma_make_tmap(&plan.tm.x[l], x, fp16, 32, 32, true, B, H, W, Ly.pitch, Ly.rbox, 0)) return msg;
    for (int m = 0; m < l; ++m)
      if (const char* msg = umma_make_tmap(&plan.tm.g[l * (l - 1) / 2 + m], p.scr[m], fp16, 16, 16, true, num_sms, g.Hv[m], g.Wv[m], Ly.pitch, Ly.rbox, 0))
        return msg;
  }
  return nullptr;
}

template <typename T>
inline const char* rdb_chain_launch(const ChainPlan& plan, cudaStream_t st) {
  static bool configured[kMaxDevices] = {};
  bool* flag = func_configured_flag(configured);
  if (!flag || !*flag) {
    cudaError_t e = cudaFuncSetAttribute(rdb_chain_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cudaGetErrorString(e);
    if (flag) *flag = true;
  }
  cudaError_t e = launch_pdl(rdb_chain_kernel<T>, dim3(plan.grid), dim3(kUmmaThreads), plan.smem_bytes, st, plan.p, plan.tm);
  if (e == cudaSuccess) e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
template <> inline const char* rdb_chain_launch<float>(const ChainPlan&, cudaStream_t) { return "tensor-core path is 16-bit only"; }

}  // namespace lpsr
