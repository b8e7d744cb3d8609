// forward_split.cuh -- the fp32-ACCURACY forward on tensor cores (LPSR_PREC_FP32 with split tensors; BASELINE.json configs[1]).
//
// The reference computes in fp32 (my_models/lpsr.py:269-274).  tcgen05 has no fp32 x fp32 mode, so every dense trunk convolution runs on
// "split" (double-fp16) operands: a real value is hi + lo, two fp16 numbers with 22 significant bits together, activations are stored as
// [16 hi | 16 lo] per 16-channel chunk (umma_conv.cuh: store_chunk16_split), weights as hi and lo K-rows, and each layer accumulates
// A_hi W_hi + A_lo W_hi + A_hi W_lo in fp32 (TMEM).  Held to <= 1e-4 against the reference on every golden case (tests/test_gpu_parity.py),
// at several times the speed of the FFMA path (forward_impl<float>, kept behind LPSR_FP32_FFMA=1).
//   tensor cores, split operands : shallowF2, the 8 dense RDB layers, both lff 1x1s (+ residual), the CSAR conv_in pair (+ pool partial sums),
//                                  gff.0, gff.1 (+ global residual), final conv (+ sigmoid)
//   CUDA cores, fp32             : the AutoEncoder and shallowF1 (3 input channels: 1.3 % of the FLOPs), the CSAR tail (reads and writes
//                                  split tensors directly; its three per-pixel 1x1s are 2 % of the FLOPs)
// The fused "last dense layer + lff" kernel is not used here: with 3x the K rows its weights alone would take 115 KB of shared memory.
#pragma once
#include "csar_tail_split_umma.cuh"
#include "forward_impl.cuh"

namespace lpsr {

inline int forward_split(lpsr_handle* h, const float* x, float* y, int B, int H, int W, char* ws, cudaStream_t st, bool dry, int* n_launch,
                         LaunchProfile* prof) {
  using TH = __half;
  const WsLayout L = ws_layout(h, B, H, W);
  Ctx c{h, st, dry};
  c.prof = prof;
  const int Hp = L.Hp, Wp = L.Wp;
  float* c0 = reinterpret_cast<float*>(ws + L.c0);
  float* e0 = reinterpret_cast<float*>(ws + L.e0);
  float* e1 = reinterpret_cast<float*>(ws + L.e1);
  float* d0 = reinterpret_cast<float*>(ws + L.d0);
  float* s = reinterpret_cast<float*>(ws + L.s);
  float* ae = reinterpret_cast<float*>(ws + L.ae);
  float* sfe1_f = reinterpret_cast<float*>(ws + L.sfe1);
  // every trunk tensor below is a split tensor: 2 x (real channels) 16-bit channels per pixel, the same bytes as its fp32 twin
  TH* sfe1 = reinterpret_cast<TH*>(ws + L.sfe1);
  TH* x0 = reinterpret_cast<TH*>(ws + L.x0);
  TH* g0 = reinterpret_cast<TH*>(ws + L.g0);
  TH* g = reinterpret_cast<TH*>(ws + L.g);
  TH* t = reinterpret_cast<TH*>(ws + L.t);
  TH* xin = reinterpret_cast<TH*>(ws + L.xin);
  float* pool = reinterpret_cast<float*>(ws + L.pool);
  TH* f[4];
  for (int i = 0; i < 4; ++i) f[i] = reinterpret_cast<TH*>(ws + L.f[i]);

  // ---- AutoEncoder + shallowF1 on CUDA cores, fp32 (lpsr.py:106-117, 195-197) ----------------------------------------------------------
  c.tag = "ae.conv_in";
  {
    c.begin("ae_conv_in");
    if (!c.dry && c.rc == LPSR_OK) {
      const long long total = (long long)B * Hp * Wp;
      const int grid = (int)std::min<long long>((total + kThreads - 1) / kThreads, (long long)h->num_sms * 16);
      ae_conv_in_kernel<float><<<grid, kThreads, 0, st>>>(x, h->ae_in.w, c0, B, Hp, Wp, H, W);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) c.rc = fail(h, LPSR_ERR_CUDA, "ae_conv_in launch: %s", cudaGetErrorString(e));
    }
  }
  c.tag = "ae.enc0";
  launch_dconv<float, 12, 12, kShuffleDown, false>(c, h->dc[0], c0, 12, e0, 48, nullptr, 0, B, Hp, Wp);
  c.tag = "ae.enc1";
  launch_dconv<float, 48, 12, kShuffleDown, false>(c, h->dc[1], e0, 48, e1, 48, nullptr, 0, B, Hp / 2, Wp / 2);
  c.tag = "ae.dec0";
  launch_dconv<float, 48, 48, kShuffleUp, false>(c, h->dc[2], e1, 48, d0, 12, nullptr, 0, B, Hp / 4, Wp / 4);
  c.tag = "ae.dec1";
  launch_dconv<float, 12, 48, kShuffleUp, true>(c, h->dc[3], d0, 12, s, 12, c0, 12, B, Hp / 2, Wp / 2);
  c.tag = "ae.conv_out";
  launch_direct<float, 3, 12, 3, false, false>(c, conv_params(h->ae_out, s, 12, 0, 12, ae, 3, 0, B, Hp, Wp, false));
  c.tag = "rdn.shallowF1";
  static const bool sfe1_ffma = getenv("LPSR_SPLIT_SFE1_FFMA") != nullptr;
  if (h->sfe1_us.packed && !sfe1_ffma) {
    // 7x7 on tensor cores (umma_conv.cuh kConv7x7, hi + lo weights): the operand is the AutoEncoder output as 16-byte split pixels, staged in the
    // region of x0 (written only afterwards, by shallowF2)
    TH* ae8 = reinterpret_cast<TH*>(ws + L.x0);
    c.begin("ae_to_pix8_split");
    if (!c.dry && c.rc == LPSR_OK) {
      const long long n_px = (long long)B * L.P;
      ae_to_pix8_split_kernel<<<(int)std::min<long long>((n_px + 255) / 256, (long long)h->num_sms * 16), 256, 0, st>>>(ae, ae8, n_px);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) c.rc = fail(h, LPSR_ERR_CUDA, "ae_to_pix8_split launch: %s", cudaGetErrorString(e));
    }
    c.begin("umma_conv7x7_split");
    if (!c.dry && c.rc == LPSR_OK) {
      ConvW w7;
      w7.ks = 7; w7.cin = 896; w7.cout = 32;
      const char* msg = umma_conv_launch<TH>(h->sfe1_us, conv_params(w7, {Seg{ae8, 8, 0, 8}}, 8, sfe1, 64, 0, B, Hp, Wp, false), h->num_sms, c.st);
      if (msg) c.rc = fail(h, LPSR_ERR_CUDA, "umma_conv 7x7 (split) launch: %s", msg);
    }
  } else {
    launch_direct<float, 7, 3, 32, false, false>(c, conv_params(h->sfe1, ae, 3, 0, 3, sfe1_f, 32, 0, B, Hp, Wp, false));
    c.begin("f32_to_split");
    if (!c.dry && c.rc == LPSR_OK) {
      const long long n_chunks = (long long)B * L.P * 2;
      f32_to_split_inplace_kernel<<<(int)std::min<long long>((n_chunks + 255) / 256, (long long)h->num_sms * 16), 256, 0, st>>>(sfe1_f, n_chunks);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) c.rc = fail(h, LPSR_ERR_CUDA, "f32_to_split launch: %s", cudaGetErrorString(e));
    }
  }

  // ---- trunk on tensor cores, split operands ---------------------------------------------------------------------------------------------
  auto conv = [&](const char* what, const UmmaWeights& u, const ConvParams& p, const UmmaGate* gate = nullptr) {
    c.begin("umma_conv_split");
    if (c.dry || c.rc != LPSR_OK) return;
    const char* msg = u.packed ? umma_conv_launch<TH>(u, p, h->num_sms, c.st, gate) : "split weights not packed";
    if (msg) c.rc = fail(h, LPSR_ERR_CUDA, "umma_conv (split) %s launch: %s", what, msg);
  };
  auto S32 = [](const TH* p) { return Seg{p, 64, 0, 64}; };    // 32 real channels
  auto S16 = [](const TH* p) { return Seg{p, 32, 0, 32}; };    // 16 real channels
  c.tag = "rdn.shallowF2";
  conv("shallowF2", h->sfe2.us, conv_params(h->sfe2, {S32(sfe1)}, 16, x0, 64, 0, B, Hp, Wp, false));
  auto rdb = [&](int r, const TH* xb, TH* out) {   // lpsr.py:43-61
    TH* gr[4];
    for (int i = 0; i < 4; ++i) gr[i] = reinterpret_cast<TH*>(ws + L.grow[r][i]);
    conv("rdb.0", h->rdb[r][0].us, conv_params(h->rdb[r][0], {S32(xb)}, 16, gr[0], 32, 0, B, Hp, Wp, true));
    conv("rdb.1", h->rdb[r][1].us, conv_params(h->rdb[r][1], {S32(xb), S16(gr[0])}, 16, gr[1], 32, 0, B, Hp, Wp, true));
    conv("rdb.2", h->rdb[r][2].us, conv_params(h->rdb[r][2], {S32(xb), S16(gr[0]), S16(gr[1])}, 16, gr[2], 32, 0, B, Hp, Wp, true));
    conv("rdb.3", h->rdb[r][3].us, conv_params(h->rdb[r][3], {S32(xb), S16(gr[0]), S16(gr[1]), S16(gr[2])}, 16, gr[3], 32, 0, B, Hp, Wp, true));
    conv("rdb.lff", h->lff[r].us, conv_params(h->lff[r], {S32(xb), S16(gr[0]), S16(gr[1]), S16(gr[2]), S16(gr[3])}, 16, out, 64, 0, B, Hp, Wp, false,
                                              xb, 64, 0));
  };
  auto csar = [&](const TH* xb, TH* out, const char* tag_conv, const char* tag_tail) {   // lpsr.py:156-186
    c.tag = tag_conv;
    conv("csar.conv_in.0", h->csar_c1.us, conv_params(h->csar_c1, {S32(xb)}, 16, t, 64, 0, B, Hp, Wp, true));
    int pool_slots = 0;
    UmmaGate pg{};
    pg.epi = kEpiPool;
    pg.pool = pool;
    pg.pool_slots_per_crop = L.pool_slots;
    pg.out_slots_per_crop = &pool_slots;
    conv("csar.conv_in.2 + pool", h->csar_c2.us, conv_params(h->csar_c2, {S32(t)}, 16, xin, 64, 0, B, Hp, Wp, false), &pg);
    c.tag = tag_tail;
    float* sc = reinterpret_cast<float*>(ws + L.sc);
    // the tail on tensor cores (csar_tail_split_umma.cuh) carries the channel branch as x_in^2 * s_c / kCsarChanScale (fp16 hi parts must not
    // overflow); conv_out's channel-branch weight rows are multiplied by the same power of two at pack time.  LPSR_SPLIT_TAIL_FFMA=1: CUDA cores.
    static const bool tail_tc = !getenv("LPSR_SPLIT_TAIL_FFMA");
    const bool use_tc = tail_tc && h->csar_sa1.us.packed && h->csar_sa2.us.packed && h->csar_co_us_scaled.packed;
    c.begin("channel_gate");   // s_c once per crop from the pooled partial sums (lpsr.py:120-135)
    if (!c.dry && c.rc == LPSR_OK) {
      channel_gate_kernel<<<dim3(B), dim3(256), 0, st>>>((const float*)pool, pool_slots, L.P, (const float*)h->ca_w1, (const float*)h->ca_b1,
                                                        (const float*)h->ca_w2, (const float*)h->ca_b2, sc, use_tc ? 1.f / kCsarChanScale : 1.f);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) c.rc = fail(h, LPSR_ERR_CUDA, "channel_gate launch: %s", cudaGetErrorString(e));
    }
    if (use_tc) {
      c.begin("csar_tail_split_umma");
      if (!c.dry && c.rc == LPSR_OK) {
        TailSplitParams p{};
        p.x_in = xin;
        p.res = xb; p.res_pitch = 64; p.res_off = 0;
        p.out = out; p.out_pitch = 64; p.out_off = 0;
        p.w3 = h->csar_sa1.us.w; p.b3 = h->csar_sa1.us.bias; p.s3 = h->csar_sa1.us.acc_scale;
        p.w4 = h->csar_sa2.us.w; p.b4 = h->csar_sa2.us.bias; p.s4 = h->csar_sa2.us.acc_scale;
        p.wo = h->csar_co_us_scaled.w; p.bo = h->csar_co_us_scaled.bias; p.so = h->csar_co_us_scaled.acc_scale;
        p.s_c = sc;
        p.total_px = (long long)B * L.P;
        p.px_per_crop = L.P;
        if (const char* msg = csar_tail_split_umma_launch(p, h->num_sms, st)) c.rc = fail(h, LPSR_ERR_CUDA, "csar_tail (split, tensor cores): %s", msg);
      }
      return;
    }
    c.begin("csar_tail_split");
    if (!c.dry && c.rc == LPSR_OK) {
      TailParams p{};
      p.x_in = xin; p.xin_pitch = 64; p.xin_off = 0;
      p.res = xb; p.res_pitch = 64; p.res_off = 0;
      p.out = out; p.out_pitch = 64; p.out_off = 0;
      p.out2 = nullptr;
      p.pool_partial = pool; p.S = pool_slots;
      p.sc = sc;
      p.ca_w1 = h->ca_w1; p.ca_b1 = h->ca_b1; p.ca_w2 = h->ca_w2; p.ca_b2 = h->ca_b2;
      p.sa_w1 = h->sa_w1; p.sa_b1 = h->sa_b1; p.sa_w2 = h->sa_w2; p.sa_b2 = h->sa_b2;
      p.co_w = h->co_w; p.co_b = h->co_b;
      p.B = B; p.P = L.P;
      csar_tail_kernel<TH, true><<<dim3((L.P + kThreads - 1) / kThreads, B), kThreads, 0, st>>>(p);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) c.rc = fail(h, LPSR_ERR_CUDA, "csar_tail (split) launch: %s", cudaGetErrorString(e));
    }
  };
  c.tag = "rdb0";
  rdb(0, x0, f[0]);
  csar(f[0], f[1], "csar1.conv_in", "csar1.tail");
  c.tag = "rdb2";
  rdb(1, f[1], f[2]);
  csar(f[2], f[3], "csar3.conv_in", "csar3.tail");
  c.tag = "rdn.gff0";
  conv("gff.0", h->gff0.us, conv_params(h->gff0, {S32(f[0]), S32(f[1]), S32(f[2]), S32(f[3])}, 16, g0, 64, 0, B, Hp, Wp, false));
  c.tag = "rdn.gff1";
  conv("gff.1", h->gff1.us, conv_params(h->gff1, {S32(g0)}, 16, g, 64, 0, B, Hp, Wp, false, sfe1, 64, 0));
  c.tag = "final_conv";
  {
    ConvW fw;
    fw.ks = 3; fw.cin = 32; fw.cout = 16;
    UmmaGate fg{};
    fg.final_sigmoid = 1;
    conv("final", h->fin_us, conv_params(fw, {S32(g)}, 16, y, 1, 0, B, Hp, Wp, false), &fg);
  }
  if (n_launch) *n_launch = c.launches;
  if (prof && !dry) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) == cudaSuccess) { cudaEventRecord(e, st); prof->events.push_back(e); }
  }
  return c.rc;
}

}  // namespace lpsr
