// forward_impl.cuh -- the launch plan of one LPSR forward (reference: my_models/lpsr.py:269-274), templated on
// the activation storage type T.  Included only by inst_*.cu, which explicitly instantiate it per dtype so the
// three dtypes compile in parallel.
#pragma once
#include <algorithm>

#include "cuda_core_kernels.cuh"
#include "engine_internal.h"
#include "umma_conv.cuh"
#include "rowconv.cuh"
#include "rowchain.cuh"
#include "csar_tail_umma.cuh"

namespace lpsr {

// ---- launch helpers ------------------------------------------------------------------------------------
struct Ctx {
  lpsr_handle* h;
  cudaStream_t st;
  bool dry;
  int launches = 0;
  int rc = LPSR_OK;
  LaunchProfile* prof = nullptr;   // optional: one cudaEvent before every launch (lpsr_forward_profiled)
  const char* tag = "";            // layer label for the profile
  void begin(const char* kind) {
    launches++;
    if (prof && !dry && rc == LPSR_OK) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) == cudaSuccess) {
        cudaEventRecord(e, st);
        prof->events.push_back(e);
        prof->names.push_back(std::string(tag) + ":" + kind);
      }
    }
  }
};

template <typename T, int KS, int CCH, int COUT, bool IN_NCHW, bool OUT_SIG>
void launch_direct(Ctx& c, const ConvParams& p) {
  c.begin("conv_direct");
  if (c.dry || c.rc != LPSR_OK) return;
  dim3 grid((p.W + kTileW - 1) / kTileW, (p.H + kTileH - 1) / kTileH, p.B);
  conv_direct_kernel<T, KS, CCH, COUT, IN_NCHW, OUT_SIG><<<grid, kThreads, 0, c.st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) c.rc = fail(c.h, LPSR_ERR_CUDA, "conv_direct<%d,%d,%d> launch: %s", KS, CCH, COUT, cudaGetErrorString(e));
}

struct Seg { const void* ptr; int pitch, off, nch; };   // nch channels [off, off+nch) of a dense NHWC tensor

// input = concatenation of channel segments (possibly of different tensors), consumed in chunks of cch channels
inline ConvParams conv_params(const ConvW& w, std::initializer_list<Seg> segs, int cch, void* out, int out_pitch, int out_off,
                              int B, int H, int Wd, bool relu, const void* res = nullptr, int res_pitch = 0, int res_off = 0) {
  ConvParams p{};
  p.in = segs.begin()->ptr; p.in_pitch = segs.begin()->pitch;
  p.n_chunks = 0;
  for (const Seg& s : segs)
    for (int c = 0; c < s.nch; c += cch) {
      if (p.n_chunks < kMaxChunks) {
        p.chunk_ptr[p.n_chunks] = s.ptr;
        p.chunk_pitch[p.n_chunks] = s.pitch;
        p.chunk_off[p.n_chunks] = s.off + c;
      }
      ++p.n_chunks;
    }
  p.w = w.w; p.bias = w.b;
  p.out = out; p.out_pitch = out_pitch; p.out_off = out_off;
  p.res = res; p.res_pitch = res_pitch; p.res_off = res_off;
  p.B = B; p.H = H; p.W = Wd; p.inH = H; p.inW = Wd;
  p.relu = relu ? 1 : 0;
  return p;
}
inline ConvParams conv_params(const ConvW& w, const void* in, int in_pitch, int in_off, int cch, void* out, int out_pitch, int out_off,
                              int B, int H, int Wd, bool relu, const void* res = nullptr, int res_pitch = 0, int res_off = 0) {
  return conv_params(w, {Seg{in, in_pitch, in_off, w.cin}}, cch, out, out_pitch, out_off, B, H, Wd, relu, res, res_pitch, res_off);
}

// dense 3x3 / 1x1 conv with Cin % 16 == 0: tensor cores in the 16-bit modes, FFMA in fp32 mode
template <typename T>
void dense_conv(Ctx& c, const ConvW& w, ConvParams p) {
  if constexpr (sizeof(T) == 2) {
    if (w.u.packed) {
      if (w.rw.packed && w.ks == 3 && rowconv_enabled() && !(p.relu && p.res)) {
        RowPlan rp;
        if (!rowconv_plan(rp, w.rw, p, c.h->num_sms, !IsBf16<T>::value)) {   // shapes it does not cover fall through to umma_conv
          c.begin("rowconv");
          if (c.dry || c.rc != LPSR_OK) return;
          bool handled = false;
          const char* msg = rowconv_launch<T>(w.rw, p, c.h->num_sms, c.st, &handled);
          if (msg) c.rc = fail(c.h, LPSR_ERR_CUDA, "rowconv launch (cin=%d cout=%d): %s", w.cin, w.cout, msg);
          if (handled || msg) return;
          c.launches--;
        }
      }
      c.begin("umma_conv");
      if (c.dry || c.rc != LPSR_OK) return;
      const char* msg = umma_conv_launch<T>(w.u, p, c.h->num_sms, c.st);
      if (msg) c.rc = fail(c.h, LPSR_ERR_CUDA, "umma_conv launch (ks=%d cin=%d cout=%d): %s", w.ks, w.cin, w.cout, msg);
      return;
    }
  }
  if (w.ks == 3 && w.cout == 16) return launch_direct<T, 3, 16, 16, false, false>(c, p);
  if (w.ks == 3 && w.cout == 32) return launch_direct<T, 3, 16, 32, false, false>(c, p);
  if (w.ks == 1 && w.cout == 32) return launch_direct<T, 1, 16, 32, false, false>(c, p);
  if (w.ks == 1 && w.cout == 16) return launch_direct<T, 1, 16, 16, false, false>(c, p);
  if (w.ks == 1 && w.cout == 64) return launch_direct<T, 1, 16, 64, false, false>(c, p);
  c.rc = fail(c.h, LPSR_ERR_UNSUPPORTED, "no kernel for conv ks=%d cin=%d cout=%d", w.ks, w.cin, w.cout);
}

template <typename T, int CIN, int COUT, int MODE, bool ADD>
void launch_dconv(Ctx& c, const DConvW& w, const void* in, int in_pitch, void* out, int out_pitch, const void* res, int res_pitch,
                  int B, int H, int Wd) {
  c.begin("dconv_fused");
  if (c.dry || c.rc != LPSR_OK) return;
  DConvParams p{};
  p.in = in; p.in_pitch = in_pitch; p.in_off = 0;
  p.dw_w = w.dw_w; p.dw_b = w.dw_b; p.pw_w = w.pw_w; p.pw_b = w.pw_b;
  p.out = out; p.out_pitch = out_pitch; p.out_off = 0;
  p.res = res; p.res_pitch = res_pitch; p.res_off = 0;
  p.B = B; p.H = H; p.W = Wd;
  const int tiles_x = (Wd + kDcTileW - 1) / kDcTileW, tiles_y = (H + kDcTileH - 1) / kDcTileH;
  const size_t smem = dconv_smem_bytes<T, CIN>(COUT);
  static bool configured[kMaxDevices] = {};   // the attribute is per device
  bool* flag = func_configured_flag(configured);
  if (!flag || !*flag) {
    if (cudaFuncSetAttribute(dconv_fused_kernel<T, CIN, COUT, MODE, ADD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess && flag)
      *flag = true;
  }
  const long long n_tiles = (long long)tiles_x * tiles_y * B;
  const int grid = (int)std::min<long long>(n_tiles, (long long)c.h->num_sms * 2);   // persistent: weights staged once per CTA
  dconv_fused_kernel<T, CIN, COUT, MODE, ADD><<<grid, kThreads, smem, c.st>>>(p, tiles_x, tiles_y);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) c.rc = fail(c.h, LPSR_ERR_CUDA, "dconv_fused<%d,%d> launch: %s", CIN, COUT, cudaGetErrorString(e));
}

template <typename T>
void csar_block(Ctx& c, const WsLayout& L, char* ws, int B, size_t in_t, size_t out_t, const char* tag_conv, const char* tag_tail) {
  lpsr_handle* h = c.h;
  T* xres = reinterpret_cast<T*>(ws + in_t);      // CSAR input x (dense 32-channel tensor)
  T* yout = reinterpret_cast<T*>(ws + out_t);
  T* t = reinterpret_cast<T*>(ws + L.t);
  T* xin = reinterpret_cast<T*>(ws + L.xin);
  float* pool = reinterpret_cast<float*>(ws + L.pool);
  c.tag = tag_conv;
  // x_in = conv_in.2(relu(conv_in.0(x)))                                               (lpsr.py:159-172,181)
  dense_conv<T>(c, h->csar_c1, conv_params(h->csar_c1, xres, 32, 0, 16, t, 32, 0, B, L.Hp, L.Wp, true));
  int pool_slots = L.S;                         // partial sums per crop feeding the channel gate
  bool pooled = false;
  if constexpr (sizeof(T) == 2) {
    static int fuse_pool = -1;
    if (fuse_pool < 0) { const char* e = getenv("LPSR_POOL_FUSED"); fuse_pool = (e && e[0] == '0') ? 0 : 1; }
    if (fuse_pool && h->csar_c2.u.packed && L.pool_slots > 0) {
      // conv_in.2 with AdaptiveAvgPool2d(1)'s partial sums (lpsr.py:124) taken from its own fp32 results: x_in is not read back
      c.begin("umma_conv");
      pooled = true;
      if (!c.dry && c.rc == LPSR_OK) {
        UmmaGate g{};
        g.epi = kEpiPool;
        g.pool = pool;
        g.pool_slots_per_crop = L.pool_slots;
        g.out_slots_per_crop = &pool_slots;
        const char* msg = umma_conv_launch<T>(h->csar_c2.u, conv_params(h->csar_c2, t, 32, 0, 16, xin, 32, 0, B, L.Hp, L.Wp, false), h->num_sms, c.st, &g);
        if (msg) c.rc = fail(h, LPSR_ERR_CUDA, "umma_conv conv_in.2 + pool launch: %s", msg);
      }
    }
  }
  if (!pooled) dense_conv<T>(c, h->csar_c2, conv_params(h->csar_c2, t, 32, 0, 16, xin, 32, 0, B, L.Hp, L.Wp, false));
  c.tag = tag_tail;
  if (!pooled) {
    // AdaptiveAvgPool2d(1) partial sums                                                    (lpsr.py:124)
    c.begin("gap_partial");
    if (!c.dry && c.rc == LPSR_OK) {
      gap_partial_kernel<T><<<dim3(L.S, B), kThreads, 0, c.st>>>(xin, 32, 0, L.P, L.S, pool);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) c.rc = fail(h, LPSR_ERR_CUDA, "gap_partial launch: %s", cudaGetErrorString(e));
    }
  }
  // gates + conv_out + residual                                                          (lpsr.py:182-186)
  if constexpr (sizeof(T) == 2) {
    if (h->csar_sa1.u.packed && h->csar_sa2.u.packed && h->csar_co.u.packed) {
      // tensor-core tail: s_c (tiny kernel) -> 1x1 32->64 ReLU -> 1x1 64->32 with the gate epilogue -> 1x1 64->32 + x
      T* hid = reinterpret_cast<T*>(ws + L.hid);
      T* gate = reinterpret_cast<T*>(ws + L.gate);
      float* sc = reinterpret_cast<float*>(ws + L.sc);
      c.begin("channel_gate");
      if (!c.dry && c.rc == LPSR_OK) {
        cudaError_t e = launch_pdl(channel_gate_kernel, dim3(B), dim3(256), 0, c.st, (const float*)pool, pool_slots, L.P, (const float*)h->ca_w1,
                                   (const float*)h->ca_b1, (const float*)h->ca_w2, (const float*)h->ca_b2, sc, 1.f / kCsarChanScale);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) c.rc = fail(h, LPSR_ERR_CUDA, "channel_gate launch: %s", cudaGetErrorString(e));
      }
      static int fused = -1;
      if (fused < 0) { const char* e = getenv("LPSR_TAIL_FUSED"); fused = (e && e[0] == '0') ? 0 : 1; }
      if (fused) {   // one kernel: three chained GEMMs, hidden and gated maps stay in shared memory / TMEM
        c.begin("csar_tail_umma");
        if (!c.dry && c.rc == LPSR_OK) {
          TailUmmaParams tp{};
          tp.x_in = xin;
          tp.res = xres; tp.res_pitch = 32; tp.res_off = 0;
          tp.out = yout; tp.out_pitch = 32; tp.out_off = 0;
          tp.w3 = h->csar_sa1.u.w; tp.b3 = h->csar_sa1.u.bias;
          tp.w4 = h->csar_sa2.u.w; tp.b4 = h->csar_sa2.u.bias;
          tp.wo = h->csar_co.u.w; tp.bo = h->csar_co.u.bias;
          tp.s_c = sc;
          tp.total_px = (long long)B * L.P;
          tp.px_per_crop = L.P;
          const char* msg = csar_tail_umma_launch<T>(tp, h->cfg.precision == LPSR_PREC_FP16, h->num_sms, c.st);
          if (msg) c.rc = fail(h, LPSR_ERR_CUDA, "csar_tail_umma launch: %s", msg);
        }
        return;
      }
      dense_conv<T>(c, h->csar_sa1, conv_params(h->csar_sa1, xin, 32, 0, 16, hid, 64, 0, B, L.Hp, L.Wp, true));
      c.begin("umma_conv_gate");
      if (!c.dry && c.rc == LPSR_OK) {
        UmmaGate g{xin, 32, 0, sc, 32};
        const char* msg = umma_conv_launch<T>(h->csar_sa2.u, conv_params(h->csar_sa2, hid, 64, 0, 16, gate, 64, 0, B, L.Hp, L.Wp, false),
                                              h->num_sms, c.st, &g);
        if (msg) c.rc = fail(h, LPSR_ERR_CUDA, "umma_conv gate launch: %s", msg);
      }
      dense_conv<T>(c, h->csar_co, conv_params(h->csar_co, gate, 64, 0, 16, yout, 32, 0, B, L.Hp, L.Wp, false, xres, 32, 0));
      return;
    }
  }
  c.begin("csar_tail");
  if (!c.dry && c.rc == LPSR_OK) {
    TailParams p{};
    p.x_in = xin; p.xin_pitch = 32; p.xin_off = 0;
    p.res = xres; p.res_pitch = 32; p.res_off = 0;
    p.out = yout; p.out_pitch = 32; p.out_off = 0;
    p.out2 = nullptr;
    p.pool_partial = pool; p.S = L.S;
    p.ca_w1 = h->ca_w1; p.ca_b1 = h->ca_b1; p.ca_w2 = h->ca_w2; p.ca_b2 = h->ca_b2;
    p.sa_w1 = h->sa_w1; p.sa_b1 = h->sa_b1; p.sa_w2 = h->sa_w2; p.sa_b2 = h->sa_b2;
    p.co_w = h->co_w; p.co_b = h->co_b;
    p.B = B; p.P = L.P;
    csar_tail_kernel<T><<<dim3((L.P + kThreads - 1) / kThreads, B), kThreads, 0, c.st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) c.rc = fail(h, LPSR_ERR_CUDA, "csar_tail launch: %s", cudaGetErrorString(e));
  }
}

template <typename T>
void rdb_block(Ctx& c, const WsLayout& L, char* ws, int B, int r, size_t x_t, size_t out_t) {
  lpsr_handle* h = c.h;
  T* x = reinterpret_cast<T*>(ws + x_t);
  T* out = reinterpret_cast<T*>(ws + out_t);
  T* g[4];
  for (int i = 0; i < 4; ++i) g[i] = reinterpret_cast<T*>(ws + L.grow[r][i]);
  // dense layers: conv3x3(cat[x, g0..g(i-1)]) -> ReLU -> its own 16-channel tensor g[i]              (lpsr.py:31-40)
  dense_conv<T>(c, h->rdb[r][0], conv_params(h->rdb[r][0], {Seg{x, 32, 0, 32}}, 16, g[0], 16, 0, B, L.Hp, L.Wp, true));
  dense_conv<T>(c, h->rdb[r][1], conv_params(h->rdb[r][1], {Seg{x, 32, 0, 32}, Seg{g[0], 16, 0, 16}}, 16, g[1], 16, 0, B, L.Hp, L.Wp, true));
  dense_conv<T>(c, h->rdb[r][2], conv_params(h->rdb[r][2], {Seg{x, 32, 0, 32}, Seg{g[0], 16, 0, 16}, Seg{g[1], 16, 0, 16}}, 16, g[2], 16, 0, B,
                                             L.Hp, L.Wp, true));
  if constexpr (sizeof(T) == 2) {
    if (h->rdb_fused[r].packed) {
      // last dense layer + lff + alpha + residual in one launch: g3 stays in shared memory               (lpsr.py:31-40,52-61)
      c.begin("umma_conv_lff");
      if (c.dry || c.rc != LPSR_OK) return;
      const char* msg = umma_conv_launch<T>(h->rdb_fused[r], conv_params(h->rdb[r][3], {Seg{x, 32, 0, 32}, Seg{g[0], 16, 0, 16}, Seg{g[1], 16, 0, 16},
                                            Seg{g[2], 16, 0, 16}}, 16, out, 32, 0, B, L.Hp, L.Wp, false, x, 32, 0), h->num_sms, c.st);
      if (msg) c.rc = fail(h, LPSR_ERR_CUDA, "umma_conv fused layer+lff launch: %s", msg);
      return;
    }
  }
  dense_conv<T>(c, h->rdb[r][3], conv_params(h->rdb[r][3], {Seg{x, 32, 0, 32}, Seg{g[0], 16, 0, 16}, Seg{g[1], 16, 0, 16}, Seg{g[2], 16, 0, 16}},
                                             16, g[3], 16, 0, B, L.Hp, L.Wp, true));
  // x + alpha*lff(cat): alpha is folded into the packed lff weights/bias                               (lpsr.py:52-61)
  dense_conv<T>(c, h->lff[r], conv_params(h->lff[r], {Seg{x, 32, 0, 32}, Seg{g[0], 16, 0, 16}, Seg{g[1], 16, 0, 16}, Seg{g[2], 16, 0, 16},
                                                     Seg{g[3], 16, 0, 16}}, 16, out, 32, 0, B, L.Hp, L.Wp, false, x, 32, 0));
}

// AutoEncoder on tensor cores (16-bit modes).  Every stage is one dense convolution (DConv = depthwise x pointwise composed at pack
// time) evaluated on the half (H/2 x W/2) or quarter grid, where PixelUnshuffle / PixelShuffle are channel relabelings:
//   xu  [H/2][W/2][16]  = unshuffle(pad4(x))                       ae_unshuffle_in_kernel
//   c0u [H/2][W/2][48]  = unshuffle(conv_in(x))                    3x3 taps on the half grid, K=16, N=48     (lpsr.py:67-69)
//   e0  [H/2][W/2][48]  = relu(unshuffle(DConv(c0)))               3x3 taps on the half grid, K=48, N=48     (lpsr.py:71-73)
//   e1  [H/4][W/4][48]  = relu(unshuffle(DConv(e0)))               5x5 taps on the half grid, N=12, unshuffling store (74-80)
//   d0  [H/2][W/2][16]  = relu(shuffle(DConv(e1))) (12 real ch)    5x5 taps on the quarter grid, N=4x16, 2x up-store (83-89)
//   s   [H/2][W/2][48]  = c0u + relu(unshuffle(shuffle(DConv(d0))))  5x5 taps on the half grid, K=16, N=48: the 48 outputs per
//                         half-grid pixel are already PixelUnshuffle(full-resolution 12 ch)                  (lpsr.py:90-96,115)
//   aeu [H/2][W/2][16]  = unshuffle(conv_out(c0 + d))              folded 3x3 on the half grid, N=12 of 16   (lpsr.py:102-104,116)
//   sfe1 [H][W][32]     = shallowF1(ae_out) (7x7)                  5x5 taps on the half grid, K=16, N=4x32, 2x up-store (195-197)
// Operands and intermediate tensors of the AutoEncoder are fp16 in BOTH 16-bit modes (its activations are O(1), and the extra three
// mantissa bits matter: the trunk amplifies AutoEncoder rounding ~20x); only the last stage converts to the trunk's type T.
template <typename T>
void ae_forward_tc(Ctx& c, const WsLayout& L, char* ws, const float* x, int B, int H, int W) {
  using TA = __half;
  lpsr_handle* h = c.h;
  TA* xu = reinterpret_cast<TA*>(ws + L.xu);
  TA* c0u = reinterpret_cast<TA*>(ws + L.c0);
  TA* e0 = reinterpret_cast<TA*>(ws + L.e0);
  TA* e1 = reinterpret_cast<TA*>(ws + L.e1);
  TA* d0 = reinterpret_cast<TA*>(ws + L.d0);
  TA* s = reinterpret_cast<TA*>(ws + L.s);
  TA* aeu = reinterpret_cast<TA*>(ws + L.ae);                  // PixelUnshuffle(AutoEncoder output): 12 real of 16 channels
  TA* sfe1 = reinterpret_cast<TA*>(ws + L.sfe1);               // written as T by the Up2Store epilogue (same element size)
  const int H2 = L.Hp / 2, W2 = L.Wp / 2, H4 = L.Hp / 4, W4 = L.Wp / 4;
  auto run = [&](const char* what, const UmmaWeights& u, const ConvParams& p, const UmmaGate* g, bool trunk_out = false) {
    c.begin("umma_conv_ae");
    if (c.dry || c.rc != LPSR_OK) return;
    // trunk_out: the stage that feeds the trunk converts to the trunk's element type in its (Up2Store) epilogue
    const char* msg = trunk_out ? umma_conv_launch<TA, T>(u, p, h->num_sms, c.st, g) : umma_conv_launch<TA>(u, p, h->num_sms, c.st, g);
    if (msg) c.rc = fail(h, LPSR_ERR_CUDA, "umma_conv %s launch: %s", what, msg);
  };
  auto shape = [](int ks, int cin, int cout) { ConvW w; w.ks = ks; w.cin = cin; w.cout = cout; return w; };
  c.tag = "ae.conv_in";
  c.begin("ae_unshuffle_in");
  if (!c.dry && c.rc == LPSR_OK) {
    const long long total = (long long)B * H2 * W2;
    const int grid = (int)std::min<long long>((total + kThreads - 1) / kThreads, (long long)h->num_sms * 16);
    ae_unshuffle_in_kernel<TA><<<grid, kThreads, 0, c.st>>>(x, xu, B, L.Hp, L.Wp, H, W);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) c.rc = fail(h, LPSR_ERR_CUDA, "ae_unshuffle_in launch: %s", cudaGetErrorString(e));
  }
  run("ae.conv_in", h->aet_in, conv_params(shape(3, 16, 48), xu, 16, 0, 16, c0u, 48, 0, B, H2, W2, false), nullptr);
  c.tag = "ae.enc0";
  run("ae.enc0", h->aet_enc0, conv_params(shape(3, 48, 48), c0u, 48, 0, 16, e0, 48, 0, B, H2, W2, true), nullptr);
  c.tag = "ae.enc1";
  {
    UmmaGate g{};
    g.epi = kEpiUnshuffleRelu;
    run("ae.enc1", h->aet_enc1, conv_params(shape(5, 48, 16), e0, 48, 0, 16, e1, 48, 0, B, H2, W2, true), &g);
  }
  c.tag = "ae.dec0";
  {
    UmmaGate g{};
    g.epi = kEpiUp2Store;                                      // PixelShuffle + ReLU in the store: four half-grid pixels per row
    run("ae.dec0", h->aet_dec0, conv_params(shape(5, 48, 64), e1, 48, 0, 16, d0, 16, 0, B, H4, W4, true), &g);
  }
  c.tag = "ae.dec1";
  run("ae.dec1", h->aet_dec1, conv_params(shape(5, 16, 48), d0, 16, 0, 16, s, 48, 0, B, H2, W2, true, c0u, 48, 0), nullptr);
  c.tag = "ae.conv_out";
  run("ae.conv_out", h->aet_out, conv_params(shape(3, 48, 16), s, 48, 0, 16, aeu, 16, 0, B, H2, W2, false), nullptr);
  // RDN shallowF1: the 7x7 over the 3-channel AutoEncoder output, evaluated on the half grid (25 coarse taps, K = 16, N = 4 x 32)
  c.tag = "rdn.shallowF1";
  {
    UmmaGate g{};
    g.epi = kEpiUp2Store;
    run("rdn.shallowF1", h->aet_sfe1, conv_params(shape(5, 16, 128), aeu, 16, 0, 16, sfe1, 32, 0, B, H2, W2, false), &g, true);
  }
}

template <typename T>
int forward_impl(lpsr_handle* h, const float* x, float* y, int B, int H, int W, char* ws, cudaStream_t st, bool dry, int* n_launch,
                 LaunchProfile* prof) {
  const WsLayout L = ws_layout(h, B, H, W);
  Ctx c{h, st, dry};
  c.prof = prof;
  T* c0 = reinterpret_cast<T*>(ws + L.c0);
  T* e0 = reinterpret_cast<T*>(ws + L.e0);
  T* e1 = reinterpret_cast<T*>(ws + L.e1);
  T* d0 = reinterpret_cast<T*>(ws + L.d0);
  T* s = reinterpret_cast<T*>(ws + L.s);
  T* ae = reinterpret_cast<T*>(ws + L.ae);
  T* sfe1 = reinterpret_cast<T*>(ws + L.sfe1);
  T* x0 = reinterpret_cast<T*>(ws + L.x0);
  T* g0 = reinterpret_cast<T*>(ws + L.g0);
  T* g = reinterpret_cast<T*>(ws + L.g);
  const int Hp = L.Hp, Wp = L.Wp;

  // ---- AutoEncoder (lpsr.py:106-117) ------------------------------------------------------------------
  bool ae_tc = false;
  if constexpr (sizeof(T) == 2) ae_tc = h->ae_tc;
  const bool sfe1_tc = (sizeof(T) == 2) && h->sfe1_u.packed && h->ae_out_u.packed;
  if (ae_tc) {
    if constexpr (sizeof(T) == 2) ae_forward_tc<T>(c, L, ws, x, B, H, W);
  } else {
  c.tag = "ae.conv_in";
  {  // conv_in 3->12 reads the caller's NCHW fp32 tensor; zero beyond (H,W) == pad-to-4 (lpsr.py:107-111)
    c.begin("ae_conv_in");
    if (!c.dry && c.rc == LPSR_OK) {
      const long long total = (long long)B * Hp * Wp;
      const int grid = (int)std::min<long long>((total + kThreads - 1) / kThreads, (long long)h->num_sms * 16);
      ae_conv_in_kernel<T><<<grid, kThreads, 0, st>>>(x, h->ae_in.w, c0, B, Hp, Wp, H, W);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) c.rc = fail(h, LPSR_ERR_CUDA, "ae_conv_in launch: %s", cudaGetErrorString(e));
    }
  }
  c.tag = "ae.enc0";
  launch_dconv<T, 12, 12, kShuffleDown, false>(c, h->dc[0], c0, 12, e0, 48, nullptr, 0, B, Hp, Wp);          // -> [48,H/2,W/2]
  c.tag = "ae.enc1";
  launch_dconv<T, 48, 12, kShuffleDown, false>(c, h->dc[1], e0, 48, e1, 48, nullptr, 0, B, Hp / 2, Wp / 2);  // -> [48,H/4,W/4]
  c.tag = "ae.dec0";
  launch_dconv<T, 48, 48, kShuffleUp, false>(c, h->dc[2], e1, 48, d0, 12, nullptr, 0, B, Hp / 4, Wp / 4);    // -> [12,H/2,W/2]
  c.tag = "ae.dec1";
  const int s_pitch = sfe1_tc ? 16 : 12;                       // tensor-core conv_out reads a 16-channel (zero padded) operand
  launch_dconv<T, 12, 48, kShuffleUp, true>(c, h->dc[3], d0, 12, s, s_pitch, c0, 12, B, Hp / 2, Wp / 2);     // -> c0 + [12,H,W]
  c.tag = "ae.conv_out";
  if constexpr (sizeof(T) == 2) {
    if (sfe1_tc) {
      c.begin("umma_conv");
      if (!c.dry && c.rc == LPSR_OK) {
        ConvW wo;
        wo.ks = 3; wo.cin = 16; wo.cout = 16;
        const char* msg = umma_conv_launch<T>(h->ae_out_u, conv_params(wo, s, 16, 0, 16, ae, 16, 0, B, Hp, Wp, false), h->num_sms, c.st);
        if (msg) c.rc = fail(h, LPSR_ERR_CUDA, "umma_conv ae.conv_out launch: %s", msg);
      }
    }
  }
  if (!sfe1_tc) launch_direct<T, 3, 12, 3, false, false>(c, conv_params(h->ae_out, s, 12, 0, 12, ae, 3, 0, B, Hp, Wp, false));
  }

  // ---- RDN (lpsr.py:214-225) ---------------------------------------------------------------------------
  c.tag = "rdn.shallowF1";
  if constexpr (sizeof(T) == 2) {
    if (sfe1_tc && !ae_tc) {
      c.begin("umma_conv7x7");
      if (!c.dry && c.rc == LPSR_OK) {
        ConvW w7;
        w7.ks = 7; w7.cin = 448; w7.cout = 32;
        const char* msg = umma_conv_launch<T>(h->sfe1_u, conv_params(w7, {Seg{ae, 16, 0, 8}}, 8, sfe1, 32, 0, B, Hp, Wp, false), h->num_sms, c.st);
        if (msg) c.rc = fail(h, LPSR_ERR_CUDA, "umma_conv 7x7 launch: %s", msg);
      }
    }
  }
  if (!sfe1_tc && !ae_tc) launch_direct<T, 7, 3, 32, false, false>(c, conv_params(h->sfe1, ae, 3, 0, 3, sfe1, 32, 0, B, Hp, Wp, false));
  c.tag = "rdn.shallowF2";
  dense_conv<T>(c, h->sfe2, conv_params(h->sfe2, sfe1, 32, 0, 16, x0, 32, 0, B, Hp, Wp, false));
  c.tag = "rdb0";
  rdb_block<T>(c, L, ws, B, 0, L.x0, L.f[0]);   // rdbs[0]
  csar_block<T>(c, L, ws, B, L.f[0], L.f[1], "csar1.conv_in", "csar1.tail");         // rdbs[1] = shared CSAR
  c.tag = "rdb2";
  rdb_block<T>(c, L, ws, B, 1, L.f[1], L.f[2]); // rdbs[2]
  csar_block<T>(c, L, ws, B, L.f[2], L.f[3], "csar3.conv_in", "csar3.tail");         // rdbs[3] = same CSAR weights
  c.tag = "rdn.gff0";
  {  // gff.0 1x1 over cat(local features): a gather list over the four block outputs, no 128-channel copy (lpsr.py:207-210,224)
    T* f[4];
    for (int i = 0; i < 4; ++i) f[i] = reinterpret_cast<T*>(ws + L.f[i]);
    dense_conv<T>(c, h->gff0, conv_params(h->gff0, {Seg{f[0], 32, 0, 32}, Seg{f[1], 32, 0, 32}, Seg{f[2], 32, 0, 32}, Seg{f[3], 32, 0, 32}}, 16,
                                          g0, 32, 0, B, Hp, Wp, false));
  }
  // gff.1 3x3 + global residual sfe1 (lpsr.py:211,224)
  c.tag = "rdn.gff1";
  bool fin_done = false;
  if constexpr (sizeof(T) == 2) {
    // gff.1 -> final conv + sigmoid chained in ONE row-streaming kernel (rowchain.cuh): the 32-channel tensor between them stays in shared memory
    if (h->gff1.rw.packed && h->fin_rw.packed) {
      bool handled = false;
      const char* msg = rowchain_launch<T>(h->gff1.rw, h->fin_rw, g0, 32, 0, sfe1, 32, 0, y, B, Hp, Wp, h->num_sms, c.st, &handled, /*dry=*/true);
      if (handled && !msg) {
        c.begin("rowconv_chain");
        fin_done = true;
        if (!c.dry && c.rc == LPSR_OK) {
          msg = rowchain_launch<T>(h->gff1.rw, h->fin_rw, g0, 32, 0, sfe1, 32, 0, y, B, Hp, Wp, h->num_sms, c.st, &handled);
          if (msg) c.rc = fail(h, LPSR_ERR_CUDA, "rowchain gff.1 + final launch: %s", msg);
        }
      }
    }
  }
  if (!fin_done) dense_conv<T>(c, h->gff1, conv_params(h->gff1, g0, 32, 0, 16, g, 32, 0, B, Hp, Wp, false, sfe1, 32, 0));
  // ---- final conv + sigmoid, NCHW fp32 out (lpsr.py:273-274) -----------------------------------------------
  c.tag = "final_conv";
  if constexpr (sizeof(T) == 2) {
    if (h->fin_u.packed && !fin_done) {   // tensor cores: Cout padded to 16, epilogue keeps channel 0 -> sigmoid -> fp32
      ConvW fw;
      fw.ks = 3; fw.cin = 32; fw.cout = 16;
      const ConvParams fp = conv_params(fw, g, 32, 0, 16, y, 1, 0, B, Hp, Wp, false);
      RowPlan rp;
      const bool row = h->fin_rw.packed && rowconv_enabled() && !rowconv_plan(rp, h->fin_rw, fp, h->num_sms, !IsBf16<T>::value, true);
      c.begin(row ? "rowconv_final" : "umma_conv_final");
      fin_done = true;
      if (!c.dry && c.rc == LPSR_OK) {
        const char* msg = nullptr;
        bool handled = false;
        if (row) msg = rowconv_launch<T>(h->fin_rw, fp, h->num_sms, c.st, &handled, true);
        if (!handled && !msg) {
          UmmaGate fg{};
          fg.final_sigmoid = 1;
          msg = umma_conv_launch<T>(h->fin_u, fp, h->num_sms, c.st, &fg);
        }
        if (msg) c.rc = fail(h, LPSR_ERR_CUDA, "final conv launch: %s", msg);
      }
    }
  }
  if (!fin_done) launch_direct<T, 3, 16, 1, false, true>(c, conv_params(h->fin, g, 32, 0, 16, y, 1, 0, B, Hp, Wp, false));
  if (n_launch) *n_launch = c.launches;
  if (prof && !dry) {   // closing event
    cudaEvent_t e;
    if (cudaEventCreate(&e) == cudaSuccess) { cudaEventRecord(e, st); prof->events.push_back(e); }
  }
  return c.rc;
}

template <typename T>
int op_conv_impl(lpsr_handle* h, const float* x, const float* w_oihw, const float* bias, float* y, int B, int Cin, int Cout, int ks,
                        int H, int W, int relu, cudaStream_t st) {
  const size_t npix = (size_t)B * H * W;
  std::vector<float> wh((size_t)Cout * Cin * ks * ks), bh(Cout, 0.f);
  CUDA_TRY(h, cudaMemcpy(wh.data(), w_oihw, wh.size() * 4, cudaMemcpyDeviceToHost));
  if (bias) CUDA_TRY(h, cudaMemcpy(bh.data(), bias, Cout * 4, cudaMemcpyDeviceToHost));
  std::vector<float> pw((size_t)ks * ks * Cin * Cout);
  for (int co = 0; co < Cout; ++co)
    for (int ci = 0; ci < Cin; ++ci)
      for (int t = 0; t < ks * ks; ++t) pw[((size_t)t * Cin + ci) * Cout + co] = wh[((size_t)co * Cin + ci) * ks * ks + t];
  // private scratch: [packed weights | bias | umma weights | in NHWC | out NHWC]
  void* scratch = nullptr;
  const size_t wbytes = align_up(pw.size() * 4, 256) + align_up(Cout * 4, 256) + align_up(pw.size() * 2 + 4096, 256) + align_up(pw.size() * 2 * 7 / 3 + 4096, 256) + 1024;
  const size_t total = wbytes + align_up(npix * Cin * sizeof(T), 256) + align_up(npix * Cout * sizeof(T), 256);
  CUDA_TRY(h, cudaMalloc(&scratch, total));
  char* sp = static_cast<char*>(scratch);
  size_t off = 0;
  auto put = [&](const void* src, size_t bytes) -> void* {
    void* d = sp + off;
    cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice);
    off = align_up(off + bytes, 256);
    return d;
  };
  ConvW cw;
  cw.ks = ks; cw.cin = Cin; cw.cout = Cout;
  cw.w = static_cast<float*>(put(pw.data(), pw.size() * 4));
  cw.b = bias ? static_cast<float*>(put(bh.data(), Cout * 4)) : nullptr;
  if (sizeof(T) == 2 && umma_supported(ks, Cin, Cout)) {
    bool ok = umma_pack_weights(cw.u, pw.data(), bias ? bh.data() : nullptr, ks, Cin, Cout, h->cfg.precision == LPSR_PREC_FP16,
                                [&](const std::vector<uint16_t>& v) { return static_cast<uint16_t*>(put(v.data(), v.size() * 2)); },
                                [&](const std::vector<float>& v) { return static_cast<float*>(put(v.data(), v.size() * 4)); });
    if (ok && ks == 3 && (Cout == 16 || Cout == 32))
      ok = rowconv_pack_weights(cw.rw, pw.data(), bias ? bh.data() : nullptr, Cin, Cout, h->cfg.precision == LPSR_PREC_FP16,
                                [&](const std::vector<uint16_t>& v) { return static_cast<uint16_t*>(put(v.data(), v.size() * 2)); },
                                [&](const std::vector<float>& v) { return static_cast<float*>(put(v.data(), v.size() * 4)); });
    if (!ok || off > wbytes) { cudaFree(scratch); return fail(h, LPSR_ERR_CUDA, "op_conv: umma weight packing failed"); }
  }
  T* in = reinterpret_cast<T*>(sp + wbytes);
  T* out = reinterpret_cast<T*>(sp + wbytes + align_up(npix * Cin * sizeof(T), 256));
  const int blocks = (int)std::min<size_t>(8192, (npix * Cin + 255) / 256);
  nchw_to_nhwc_kernel<T><<<blocks, 256, 0, st>>>(x, in, B, Cin, H, W, Cin, 0);
  Ctx c{h, st, false};
  dense_conv<T>(c, cw, conv_params(cw, in, Cin, 0, 16, out, Cout, 0, B, H, W, relu != 0));
  if (c.rc == LPSR_OK) {
    const int blocks2 = (int)std::min<size_t>(8192, (npix * Cout + 255) / 256);
    nhwc_to_nchw_kernel<T><<<blocks2, 256, 0, st>>>(out, y, B, Cout, H, W, Cout, 0, 0);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(scratch);
  if (c.rc != LPSR_OK) return c.rc;
  if (e != cudaSuccess) return fail(h, LPSR_ERR_CUDA, "op_conv2d: %s", cudaGetErrorString(e));
  return LPSR_OK;
}


template <typename T>
int tap_copy_impl(lpsr_handle* h, const void* src, float* dst, int B, int C, int H, int W, int pitch, int off, int unshuffled, cudaStream_t st) {
  const long long need = (long long)B * C * H * W;
  const int blocks = (int)std::min<long long>(4096, (need + 255) / 256);
  nhwc_to_nchw_kernel<T><<<blocks, 256, 0, st>>>(static_cast<const T*>(src), dst, B, C, H, W, pitch, off, unshuffled);
  CUDA_TRY(h, cudaGetLastError());
  return LPSR_OK;
}

#define LPSR_INSTANTIATE(T)                                                                                             \
  template int forward_impl<T>(lpsr_handle*, const float*, float*, int, int, int, char*, cudaStream_t, bool, int*,      \
                               LaunchProfile*);                                                                         \
  template int op_conv_impl<T>(lpsr_handle*, const float*, const float*, const float*, float*, int, int, int, int, int, \
                               int, int, cudaStream_t);                                                                 \
  template int tap_copy_impl<T>(lpsr_handle*, const void*, float*, int, int, int, int, int, int, int, cudaStream_t);

}  // namespace lpsr
