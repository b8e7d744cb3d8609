// engine.cu -- host side of the C ABI declared in include/lpsr_b200.h: weight packing, workspace layout,
// the launch plan of one LPSR forward (reference: my_models/lpsr.py:269-274) and the op-level entry points.
#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "cuda_core_kernels.cuh"
#include "engine_internal.h"
#include "preprocess.cuh"
#include "nms.cuh"
#include "umma_conv.cuh"

using namespace lpsr;

namespace lpsr {

std::mutex g_err_mu;
char g_err[512] = "";

int fail(lpsr_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) snprintf(h->err, sizeof h->err, "%s", buf);
  std::lock_guard<std::mutex> lk(g_err_mu);
  snprintf(g_err, sizeof g_err, "%s", buf);
  return code;
}

}  // namespace lpsr

namespace {

// ---- live tensor table (names = reference state_dict keys; order = ctor order, lpsr.py) --------------
void add_live(lpsr_handle* h, const std::string& name, int64_t numel) {
  h->live_index[name] = (int)h->live.size();
  LiveTensor t;
  t.name = name;
  t.numel = numel;
  h->live.push_back(std::move(t));
}

void build_live_table(lpsr_handle* h) {
  const int C = h->cfg.num_channels, F = h->cfg.num_features, G = h->cfg.growth_rate, L = h->cfg.num_layers;
  const int E = 4 * C, O = h->cfg.out_channels;
  add_live(h, "auto_encoder.conv_in.weight", (int64_t)E * C * 9);
  struct { const char* blk; int cin, cout; } dcs[4] = {{"encoder.0", E, E}, {"encoder.3", 4 * E, E}, {"decoder.0", 4 * E, 4 * E}, {"decoder.3", E, 4 * E}};
  for (auto& d : dcs) {
    std::string p = std::string("auto_encoder.") + d.blk + ".dConv.";
    add_live(h, p + "0.weight", (int64_t)d.cin * 25);
    add_live(h, p + "0.bias", d.cin);
    add_live(h, p + "1.weight", (int64_t)d.cout * d.cin);
    add_live(h, p + "1.bias", d.cout);
  }
  add_live(h, "auto_encoder.conv_out.weight", (int64_t)C * E * 9);
  add_live(h, "rdn.shallowF1.weight", (int64_t)F * C * 49);
  add_live(h, "rdn.shallowF1.bias", F);
  add_live(h, "rdn.shallowF2.weight", (int64_t)F * F * 9);
  add_live(h, "rdn.shallowF2.bias", F);
  add_live(h, "rdn.csar.conv_in.0.weight", (int64_t)F * F * 9);
  add_live(h, "rdn.csar.conv_in.0.bias", F);
  add_live(h, "rdn.csar.conv_in.2.weight", (int64_t)F * F * 9);
  add_live(h, "rdn.csar.conv_in.2.bias", F);
  add_live(h, "rdn.csar.ca.block.2.weight", (int64_t)(F / 4) * F);
  add_live(h, "rdn.csar.ca.block.2.bias", F / 4);
  add_live(h, "rdn.csar.ca.block.4.weight", (int64_t)F * (F / 4));
  add_live(h, "rdn.csar.ca.block.4.bias", F);
  add_live(h, "rdn.csar.sa.block.0.weight", (int64_t)2 * F * F);
  add_live(h, "rdn.csar.sa.block.0.bias", 2 * F);
  add_live(h, "rdn.csar.sa.block.2.weight", (int64_t)F * 2 * F);
  add_live(h, "rdn.csar.sa.block.2.bias", F);
  add_live(h, "rdn.csar.conv_out.weight", (int64_t)F * 2 * F);
  add_live(h, "rdn.csar.conv_out.bias", F);
  for (int r = 0; r < h->cfg.num_blocks; r += 2) {   // executed RDBs are the even module indices
    std::string p = "rdn.rdbs." + std::to_string(r) + ".";
    add_live(h, p + "alpha", 1);
    for (int i = 0; i < L; ++i) {
      add_live(h, p + "layers." + std::to_string(i) + ".conv.weight", (int64_t)G * (F + G * i) * 9);
      add_live(h, p + "layers." + std::to_string(i) + ".conv.bias", G);
    }
    add_live(h, p + "lff.weight", (int64_t)F * (F + G * L));
    add_live(h, p + "lff.bias", F);
  }
  add_live(h, "rdn.gff.0.weight", (int64_t)F * F * h->cfg.num_blocks);
  add_live(h, "rdn.gff.0.bias", F);
  add_live(h, "rdn.gff.1.weight", (int64_t)F * F * 9);
  add_live(h, "rdn.gff.1.bias", F);
  add_live(h, "final_conv.weight", (int64_t)O * F * 9);
  add_live(h, "final_conv.bias", O);
}

const std::vector<float>& W(const lpsr_handle* h, const std::string& name) {
  return h->live[h->live_index.at(name)].host;
}

// ---- device arena ------------------------------------------------------------------------------------
template <typename U>
U* arena_put(lpsr_handle* h, const std::vector<U>& v) {
  DeviceArena& a = h->arena;
  a.used = align_up(a.used, 256);
  const size_t bytes = v.size() * sizeof(U);
  if (a.used + bytes > a.cap) return nullptr;
  U* p = reinterpret_cast<U*>(static_cast<char*>(a.base) + a.used);
  if (cudaMemcpy(p, v.data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  a.used += bytes;
  return p;
}

// conv weight OIHW -> [ks*ks][cin][cout], optional scale (alpha fold, lpsr.py:58-61)
bool pack_conv(lpsr_handle* h, ConvW& cw, const std::string& prefix, int cin, int cout, int ks, bool bias, float scale = 1.f) {
  const std::vector<float>& w = W(h, prefix + ".weight");
  std::vector<float> pw((size_t)ks * ks * cin * cout);
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < cin; ++ci)
      for (int t = 0; t < ks * ks; ++t)
        pw[((size_t)t * cin + ci) * cout + co] = scale * w[((size_t)co * cin + ci) * ks * ks + t];
  cw.ks = ks; cw.cin = cin; cw.cout = cout;
  cw.w = arena_put(h, pw);
  if (!cw.w) return false;
  std::vector<float> pb;
  if (bias) {
    pb = W(h, prefix + ".bias");
    for (auto& v : pb) v *= scale;
    cw.b = arena_put(h, pb);
    if (!cw.b) return false;
  } else {
    cw.b = nullptr;
  }
  if (half_mode(h) && umma_supported(ks, cin, cout)) {
    quantize_conv_sum_preserving(pw, ks * ks, cin, cout, h->cfg.precision == LPSR_PREC_FP16);   // 16-bit packing only (cw.w stays exact)
    if (!umma_pack_weights(cw.u, pw.data(), bias ? pb.data() : nullptr, ks, cin, cout,
                           h->cfg.precision == LPSR_PREC_FP16, [&](const std::vector<uint16_t>& v) { return arena_put(h, v); },
                           [&](const std::vector<float>& v) { return arena_put(h, v); }))
      return false;
    // row-streaming kernel (rowconv.cuh): every 3x3 layer with 16 or 32 output channels that keeps single-rounded weights
    cw.rw = RowWeights{};
    if (ks == 3 && (cout == 16 || cout == 32) && cin <= 16 * kMaxChunks &&
        !rowconv_pack_weights(cw.rw, pw.data(), bias ? pb.data() : nullptr, cin, cout, h->cfg.precision == LPSR_PREC_FP16,
                              [&](const std::vector<uint16_t>& v) { return arena_put(h, v); }, [&](const std::vector<float>& v) { return arena_put(h, v); }))
      return false;
  }
  cw.us = UmmaWeights{};
  if (h->fp32_split && umma_supported(ks, cin, cout) && 3 * cin / 16 <= kUmmaMaxSteps &&
      !umma_pack_weights_split(cw.us, pw.data(), bias ? pb.data() : nullptr, ks, cin, cout, [&](const std::vector<uint16_t>& v) { return arena_put(h, v); },
                               [&](const std::vector<float>& v) { return arena_put(h, v); }))
    return false;
  return true;
}

// ---- tensor-core AutoEncoder weights ---------------------------------------------------------------------
// One stage of the AutoEncoder as a dense ks x ks convolution on its own grid: w[co][ci][ks*ks], b[co]
struct DenseConv {
  int cin = 0, cout = 0, ks = 0;
  std::vector<float> w, b;
};

// DConv (lpsr.py:8-28) = depthwise ks x ks (+bias) then pointwise 1x1 (+bias) with NO activation in between, i.e. one dense
// conv: w[co][ci][t] = pw[co][ci] * dw[ci][t]; the depthwise bias is added at every pixel ('same' zero padding pads the INPUT),
// so it folds exactly: b[co] = pw_b[co] + sum_ci pw[co][ci] * dw_b[ci]
DenseConv compose_dconv(const lpsr_handle* h, const std::string& p, int cin, int cout, int ks) {
  DenseConv d;
  d.cin = cin; d.cout = cout; d.ks = ks;
  const std::vector<float>&dw = W(h, p + "0.weight"), &dwb = W(h, p + "0.bias"), &pw = W(h, p + "1.weight"), &pwb = W(h, p + "1.bias");
  d.w.resize((size_t)cout * cin * ks * ks);
  d.b.resize(cout);
  for (int co = 0; co < cout; ++co) {
    double acc = pwb[co];
    for (int ci = 0; ci < cin; ++ci) {
      acc += (double)pw[(size_t)co * cin + ci] * dwb[ci];
      for (int t = 0; t < ks * ks; ++t)
        d.w[((size_t)co * cin + ci) * ks * ks + t] = pw[(size_t)co * cin + ci] * dw[(size_t)ci * ks * ks + t];
    }
    d.b[co] = (float)acc;
  }
  return d;
}

// The same convolution evaluated on the 2x COARSER grid (space-to-depth): a coarse pixel carries its 2x2 fine pixels as channels,
// fine output row 2h + I reads fine rows 2h + I + dy - R = 2(h + th) + i, th in {-1,0,1} for any ks <= 5, so the coarse kernel is
// 3x3 (5x5 for a 7x7 fine kernel: th in {-2..2}) with w'[n(co,I,J)][k(ci,i,j)][th][tw] = w[co][ci][2th + i - I + R][2tw + j - J + R]
// (zero where that tap does not exist).  in_index / out_index give the operand's channel order (out_index < 0: column not produced
// by this launch).  Returns [KC*KC][cin_cols][cout_cols] fp32 (the layout umma_pack_weights takes) and the per-column bias.
template <typename InIdx, typename OutIdx>
void s2d_weights(const DenseConv& d, int cin_cols, int cout_cols, InIdx in_index, OutIdx out_index, std::vector<float>& pw,
                 std::vector<float>& pb) {
  const int R = d.ks / 2, RC = d.ks <= 5 ? 1 : 2, KC = 2 * RC + 1;
  pw.assign((size_t)KC * KC * cin_cols * cout_cols, 0.f);
  pb.assign(cout_cols, 0.f);
  for (int co = 0; co < d.cout; ++co)
    for (int I = 0; I < 2; ++I)
      for (int J = 0; J < 2; ++J) {
        const int n = out_index(co, I, J);
        if (n < 0) continue;
        pb[n] = d.b.empty() ? 0.f : d.b[co];
        for (int ci = 0; ci < d.cin; ++ci)
          for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) {
              const int k = in_index(ci, i, j);
              for (int th = -RC; th <= RC; ++th) {
                const int dy = 2 * th + i - I + R;
                if (dy < 0 || dy >= d.ks) continue;
                for (int tw = -RC; tw <= RC; ++tw) {
                  const int dx = 2 * tw + j - J + R;
                  if (dx < 0 || dx >= d.ks) continue;
                  pw[((size_t)((th + RC) * KC + (tw + RC)) * cin_cols + k) * cout_cols + n] = d.w[((size_t)co * d.cin + ci) * d.ks * d.ks + dy * d.ks + dx];
                }
              }
            }
      }
}

// dense conv at its own grid: [ks*ks][cin_rows][cout_cols]; output channel co goes to column col_of(co), unused rows / columns are zero
template <typename ColOf>
void plain_weights(const DenseConv& d, int cin_rows, int cout_cols, ColOf col_of, std::vector<float>& pw, std::vector<float>& pb) {
  pw.assign((size_t)d.ks * d.ks * cin_rows * cout_cols, 0.f);
  pb.assign(cout_cols, 0.f);
  for (int co = 0; co < d.cout; ++co) {
    const int n = col_of(co);
    pb[n] = d.b.empty() ? 0.f : d.b[co];
    for (int ci = 0; ci < d.cin; ++ci)
      for (int t = 0; t < d.ks * d.ks; ++t) pw[((size_t)t * cin_rows + ci) * cout_cols + n] = d.w[((size_t)co * d.cin + ci) * d.ks * d.ks + t];
  }
}

bool pack_ae_tensor_core(lpsr_handle* h) {
  const bool fp16 = true;   // the AutoEncoder stages run fp16 operands in both 16-bit modes (ae_forward_tc)
  auto put16 = [&](const std::vector<uint16_t>& v) { return arena_put(h, v); };
  auto put32 = [&](const std::vector<float>& v) { return arena_put(h, v); };
  auto unshuffle_idx = [](int c, int i, int j) { return c * 4 + i * 2 + j; };   // PixelUnshuffle / PixelShuffle channel order
  std::vector<float> pw, pb;
  bool ok = true;
  // conv_in 3 -> 12 (3x3, no bias) on the half grid: operand = ae_unshuffle_in_kernel's [(i*2+j)*3 + c] (12 real of 16), output =
  // PixelUnshuffle(c0): 48 channels
  // Weight rounding is what the trained checkpoint amplifies most on smooth crops (CPU emulation: conv_in's fp16 weights alone -> 1.6e-2
  // at the output, conv_out 1.4e-3): those two stages carry their weights as hi + lo (their K is 16 / 48: the extra MMAs are free);
  // the composed DConv stages and shallowF1 use sum-preserving rounding of the fine kernel (umma_weights.h).
  auto sum_preserve = [&](DenseConv& d) {   // d.w: [co][ci][ks*ks]
    if (!sum_preserving_enabled()) return;
    for (int f = 0; f < d.cout * d.cin; ++f) quantize_taps_sum_preserving(d.w.data() + (size_t)f * d.ks * d.ks, d.ks * d.ks, 1, fp16);
  };
  DenseConv cin;
  cin.cin = 3; cin.cout = 12; cin.ks = 3; cin.w = W(h, "auto_encoder.conv_in.weight");
  s2d_weights(cin, 16, 48, [](int c, int i, int j) { return (i * 2 + j) * 3 + c; }, unshuffle_idx, pw, pb);
  ok &= umma_pack_weights_wsplit(h->aet_in, pw.data(), nullptr, 3, 16, 48, fp16, put16, put32);
  // encoder.0: DConv 12 -> 12 at full resolution + PixelUnshuffle == 48 -> 48 on the half grid (lpsr.py:71-73)
  DenseConv e0 = compose_dconv(h, "auto_encoder.encoder.0.dConv.", 12, 12, 5);
  sum_preserve(e0);
  s2d_weights(e0, 48, 48, unshuffle_idx, unshuffle_idx, pw, pb);
  ok &= umma_pack_weights(h->aet_enc0, pw.data(), pb.data(), 3, 48, 48, fp16, put16, put32);
  // encoder.3: DConv 48 -> 12 on the half grid (5x5 taps); PixelUnshuffle + ReLU in the store (lpsr.py:74-80)
  DenseConv e1 = compose_dconv(h, "auto_encoder.encoder.3.dConv.", 48, 12, 5);
  sum_preserve(e1);
  plain_weights(e1, 48, 16, [](int co) { return co; }, pw, pb);
  ok &= umma_pack_weights(h->aet_enc1, pw.data(), pb.data(), 5, 48, 16, fp16, put16, put32);
  // decoder.0: DConv 48 -> 48 on the quarter grid + PixelShuffle + ReLU (lpsr.py:83-89): output channel c*4 + i*2 + j goes to column
  // (i*2 + j)*16 + c, so the epilogue stores four 16-channel (12 real) half-grid pixels per quarter-grid pixel
  DenseConv d0 = compose_dconv(h, "auto_encoder.decoder.0.dConv.", 48, 48, 5);
  sum_preserve(d0);
  plain_weights(d0, 48, 64, [](int co) { return (co & 3) * 16 + (co >> 2); }, pw, pb);
  ok &= umma_pack_weights(h->aet_dec0, pw.data(), pb.data(), 5, 48, 64, fp16, put16, put32);
  // decoder.3: DConv 12 -> 48 on the half grid (lpsr.py:90-96); its 48 outputs per half-grid pixel ARE PixelUnshuffle of the
  // full-resolution 12 channels, the layout of c0 (residual, lpsr.py:115) and of conv_out's operand: plain ReLU + residual store
  DenseConv d1 = compose_dconv(h, "auto_encoder.decoder.3.dConv.", 12, 48, 5);
  sum_preserve(d1);
  plain_weights(d1, 16, 48, [](int co) { return co; }, pw, pb);
  ok &= umma_pack_weights(h->aet_dec1, pw.data(), pb.data(), 5, 16, 48, fp16, put16, put32);
  // conv_out 12 -> 3 (3x3, no bias) on the half grid: the output stays PixelUnshuffle(ae_out), 12 real of 16 columns
  DenseConv co;
  co.cin = 12; co.cout = 3; co.ks = 3; co.w = W(h, "auto_encoder.conv_out.weight");
  s2d_weights(co, 48, 16, unshuffle_idx, unshuffle_idx, pw, pb);
  ok &= umma_pack_weights_wsplit(h->aet_out, pw.data(), nullptr, 3, 48, 16, fp16, put16, put32);
  // RDN shallowF1 (7x7, 3 -> 32, lpsr.py:195-197) on the half grid: 5x5 coarse taps over the 12 (of 16) unshuffled channels, four
  // 32-channel full-resolution pixels per half-grid pixel (N = 128), written by the epilogue in the trunk's element type
  DenseConv s1;
  s1.cin = 3; s1.cout = 32; s1.ks = 7; s1.w = W(h, "rdn.shallowF1.weight"); s1.b = W(h, "rdn.shallowF1.bias");
  sum_preserve(s1);
  s2d_weights(s1, 16, 128, unshuffle_idx, [](int c, int I, int J) { return (I * 2 + J) * 32 + c; }, pw, pb);
  ok &= umma_pack_weights(h->aet_sfe1, pw.data(), pb.data(), 5, 16, 128, fp16, put16, put32);
  return ok;
}

int pack_all(lpsr_handle* h) {
  const int C = h->cfg.num_channels, F = h->cfg.num_features, G = h->cfg.growth_rate, L = h->cfg.num_layers, E = 4 * C;
  if (!h->arena.base) {
    h->arena.cap = 8u << 20;
    CUDA_TRY(h, cudaMalloc(&h->arena.base, h->arena.cap));
  }
  h->arena.used = 0;
  bool ok = true;
  ok &= pack_conv(h, h->ae_in, "auto_encoder.conv_in", C, E, 3, false);
  ok &= pack_conv(h, h->ae_out, "auto_encoder.conv_out", E, C, 3, false);
  const char* blks[4] = {"encoder.0", "encoder.3", "decoder.0", "decoder.3"};
  const int dcin[4] = {E, 4 * E, 4 * E, E}, dcout[4] = {E, E, 4 * E, 4 * E};
  for (int i = 0; i < 4; ++i) {
    std::string p = std::string("auto_encoder.") + blks[i] + ".dConv.";
    DConvW& d = h->dc[i];
    d.cin = dcin[i]; d.cout = dcout[i];
    d.dw_w = arena_put(h, W(h, p + "0.weight"));      // [cin][1][5][5] == [cin][25]
    d.dw_b = arena_put(h, W(h, p + "0.bias"));
    const std::vector<float>& pw = W(h, p + "1.weight");  // [cout][cin] -> [cin][cout]
    std::vector<float> t((size_t)d.cin * d.cout);
    for (int co = 0; co < d.cout; ++co)
      for (int ci = 0; ci < d.cin; ++ci) t[(size_t)ci * d.cout + co] = pw[(size_t)co * d.cin + ci];
    d.pw_w = arena_put(h, t);
    d.pw_b = arena_put(h, W(h, p + "1.bias"));
    ok &= d.dw_w && d.dw_b && d.pw_w && d.pw_b;
  }
  ok &= pack_conv(h, h->sfe1, "rdn.shallowF1", C, F, 7, true);
  h->sfe1_u.packed = false;
  if (half_mode(h) && umma_enabled()) {
    // 7x7 on tensor cores: K-step (dy, dx-pair) = 16 values = {pixel dx: 8 ch (3 real), pixel dx+1: 8 ch}; dx = 7 and ch >= 3 are zero
    const std::vector<float>& w = W(h, "rdn.shallowF1.weight");   // [F][C][7][7]
    std::vector<float> pw((size_t)448 * F, 0.f);
    for (int dy = 0; dy < 7; ++dy)
      for (int pr = 0; pr < 4; ++pr)
        for (int half = 0; half < 2; ++half) {
          const int dx = 2 * pr + half;
          if (dx >= 7) continue;
          for (int ci = 0; ci < C; ++ci)
            for (int co = 0; co < F; ++co)
              pw[((size_t)((dy * 4 + pr) * 16 + half * 8 + ci)) * F + co] = w[(((size_t)co * C + ci) * 7 + dy) * 7 + dx];
        }
    ok &= umma_pack_weights(h->sfe1_u, pw.data(), W(h, "rdn.shallowF1.bias").data(), 1, 448, F, h->cfg.precision == LPSR_PREC_FP16,
                            [&](const std::vector<uint16_t>& v) { return arena_put(h, v); }, [&](const std::vector<float>& v) { return arena_put(h, v); });
    h->sfe1_u.ks = 7;
    // AutoEncoder conv_out 12 -> 3 on tensor cores: input (c0 + decoder) is written as a 16-channel padded tensor, the output
    // as a 16-channel padded tensor whose first 8 channels (3 real) are the 16-byte pixels the 7x7 reads
    const std::vector<float>& wo = W(h, "auto_encoder.conv_out.weight");   // [C][E][3][3]
    std::vector<float> p16((size_t)9 * 16 * 16, 0.f);
    for (int co = 0; co < C; ++co)
      for (int ci = 0; ci < E; ++ci)
        for (int t = 0; t < 9; ++t) p16[((size_t)t * 16 + ci) * 16 + co] = wo[((size_t)co * E + ci) * 9 + t];
    ok &= umma_pack_weights(h->ae_out_u, p16.data(), nullptr, 3, 16, 16, h->cfg.precision == LPSR_PREC_FP16,
                            [&](const std::vector<uint16_t>& v) { return arena_put(h, v); }, [&](const std::vector<float>& v) { return arena_put(h, v); });
  }
  h->sfe1_us = UmmaWeights{};
  if (h->fp32_split && C == 3 && F == 32) {
    // fp32-accuracy mode: the same pixel-pair K-steps, but a pixel's 8 slots are [hi(c0 c1 c2) | lo(c0 c1 c2) | 0 0] (ae_to_pix8_split_kernel), so
    // one MMA per K-step covers A_hi W_hi + A_lo W_hi (rows 0..5 carry hi(W) twice) and a second pass over the same pixels with lo(W) in
    // rows 0..2 adds A_hi W_lo: 56 K-steps.  Weights times a power of two so that the lo terms stay normal fp16 numbers (acc_scale undoes it).
    const std::vector<float>& w = W(h, "rdn.shallowF1.weight");   // [F][C][7][7]
    const float sc = split_weight_scale(w.data(), w.size());
    std::vector<float> pw((size_t)896 * F, 0.f);
    for (int dy = 0; dy < 7; ++dy)
      for (int pr = 0; pr < 4; ++pr)
        for (int half = 0; half < 2; ++half) {
          const int dx = 2 * pr + half;
          if (dx >= 7) continue;
          for (int ci = 0; ci < C; ++ci)
            for (int co = 0; co < F; ++co) {
              const float wv = sc * w[(((size_t)co * C + ci) * 7 + dy) * 7 + dx];
              const float hi = bits16_to_f32(f32_to_bits16(wv, true), true);
              const size_t row = (size_t)((dy * 4 + pr) * 16 + half * 8);
              pw[(row + ci) * F + co] = hi;              // x hi(activation)
              pw[(row + 3 + ci) * F + co] = hi;          // x lo(activation)
              pw[(448 + row + ci) * F + co] = wv - hi;   // lo(W) x hi(activation)
            }
        }
    ok &= umma_pack_weights(h->sfe1_us, pw.data(), W(h, "rdn.shallowF1.bias").data(), 1, 896, F, true,
                            [&](const std::vector<uint16_t>& v) { return arena_put(h, v); }, [&](const std::vector<float>& v) { return arena_put(h, v); });
    h->sfe1_us.ks = 7;
    h->sfe1_us.split = true;      // split output tensor, acc_scale in the epilogue
    h->sfe1_us.n_real = 28;
    h->sfe1_us.wlo_step = 1;
    h->sfe1_us.acc_scale = 1.f / sc;
  }
  h->ae_tc = false;
  if (half_mode(h) && umma_enabled() && C == 3 && h->sfe1_u.packed && !getenv("LPSR_AE_CUDA_CORES")) {
    ok &= pack_ae_tensor_core(h);
    h->ae_tc = ok;
  }
  ok &= pack_conv(h, h->sfe2, "rdn.shallowF2", F, F, 3, true);
  for (int r = 0; r < 2; ++r) {
    std::string p = "rdn.rdbs." + std::to_string(2 * r);
    const float alpha = W(h, p + ".alpha")[0];
    for (int i = 0; i < L; ++i) ok &= pack_conv(h, h->rdb[r][i], p + ".layers." + std::to_string(i) + ".conv", F + G * i, G, 3, true);
    ok &= pack_conv(h, h->lff[r], p + ".lff", F + G * L, F, 1, true, alpha);
  }
  for (int r = 0; r < 2; ++r) h->rdb_fused[r] = UmmaWeights{};
  if (half_mode(h) && umma_enabled() && F == 32 && G == 16 && L >= 1 && (F + G * (L - 1)) % 16 == 0 && F + G * (L - 1) <= 16 * kUmmaMaxKChunks &&
      !getenv("LPSR_NO_LFF_FUSION")) {
    const int cin = F + G * (L - 1);
    for (int r = 0; r < 2; ++r) {
      const std::string p = "rdn.rdbs." + std::to_string(2 * r);
      const float alpha = W(h, p + ".alpha")[0];
      const std::vector<float>& w3 = W(h, p + ".layers." + std::to_string(L - 1) + ".conv.weight");   // [G][cin][3][3]
      const std::vector<float>& wl = W(h, p + ".lff.weight");                                          // [F][cin + G]
      std::vector<float> p3((size_t)9 * cin * G), pl((size_t)(cin + G) * F), bl = W(h, p + ".lff.bias");
      for (int co = 0; co < G; ++co)
        for (int ci = 0; ci < cin; ++ci)
          for (int t = 0; t < 9; ++t) p3[((size_t)t * cin + ci) * G + co] = w3[((size_t)co * cin + ci) * 9 + t];
      for (int co = 0; co < F; ++co)
        for (int ci = 0; ci < cin + G; ++ci) pl[(size_t)ci * F + co] = alpha * wl[(size_t)co * (cin + G) + ci];
      for (auto& v : bl) v *= alpha;
      quantize_conv_sum_preserving(p3, 9, cin, G, h->cfg.precision == LPSR_PREC_FP16);
      ok &= umma_pack_fused_lff(h->rdb_fused[r], p3.data(), W(h, p + ".layers." + std::to_string(L - 1) + ".conv.bias").data(), pl.data(), bl.data(),
                                cin, h->cfg.precision == LPSR_PREC_FP16, [&](const std::vector<uint16_t>& v) { return arena_put(h, v); },
                                [&](const std::vector<float>& v) { return arena_put(h, v); });
    }
  }
  ok &= pack_conv(h, h->csar_c1, "rdn.csar.conv_in.0", F, F, 3, true);
  ok &= pack_conv(h, h->csar_c2, "rdn.csar.conv_in.2", F, F, 3, true);
  if (half_mode(h) && h->csar_c1.u.packed && h->csar_c2.u.packed && F == 32 && !getenv("LPSR_NO_NSPLIT")) {
    // the CSAR conv_in pair feeds x_in, which the channel branch squares (lpsr.py:133-135): its weight rounding is, after lff's, the largest
    // remaining contribution to the output error (tools/parity_report.py, CPU emulation).  hi + lo weights along N: 48 instead of 45 clk per MMA.
    auto repack = [&](lpsr::ConvW& cw, const std::string& name) {
      const std::vector<float>& w = W(h, name + ".weight");   // [F][F][3][3]
      std::vector<float> pw((size_t)9 * F * F);
      for (int co = 0; co < F; ++co)
        for (int ci = 0; ci < F; ++ci)
          for (int t = 0; t < 9; ++t) pw[((size_t)t * F + ci) * F + co] = w[((size_t)co * F + ci) * 9 + t];
      return umma_pack_weights_nsplit(cw.u, pw.data(), W(h, name + ".bias").data(), 3, F, F, h->cfg.precision == LPSR_PREC_FP16,
                                      [&](const std::vector<uint16_t>& v) { return arena_put(h, v); }, [&](const std::vector<float>& v) { return arena_put(h, v); });
    };
    ok &= repack(h->csar_c1, "rdn.csar.conv_in.0");
    ok &= repack(h->csar_c2, "rdn.csar.conv_in.2");
    h->csar_c1.rw.packed = h->csar_c2.rw.packed = false;   // hi + lo weights: these two stay on the per-tap kernel
  }
  ok &= pack_conv(h, h->csar_sa1, "rdn.csar.sa.block.0", F, 2 * F, 1, true);    // tensor-core CSAR tail (16-bit modes)
  ok &= pack_conv(h, h->csar_sa2, "rdn.csar.sa.block.2", 2 * F, F, 1, true);
  ok &= pack_conv(h, h->csar_co, "rdn.csar.conv_out", 2 * F, F, 1, true);
  if (half_mode(h) && h->csar_co.u.packed) {
    // tensor-core packing only: input columns [0, F) (channel branch) x kCsarChanScale; the channel gate kernel writes s_c / kCsarChanScale
    const std::vector<float>& w = W(h, "rdn.csar.conv_out.weight");   // [F][2F]
    std::vector<float> pw((size_t)2 * F * F);
    for (int co = 0; co < F; ++co)
      for (int ci = 0; ci < 2 * F; ++ci) pw[(size_t)ci * F + co] = (ci < F ? kCsarChanScale : 1.f) * w[(size_t)co * 2 * F + ci];
    // hi + lo weights (K doubled): conv_out multiplies the largest operands of the network (x_in^2 * s_c reaches thousands) and its
    // weight rounding is one of the two largest single contributions to the output error (CPU emulation); the tail is not MMA bound
    ok &= umma_pack_weights_wsplit(h->csar_co.u, pw.data(), W(h, "rdn.csar.conv_out.bias").data(), 1, 2 * F, F, h->cfg.precision == LPSR_PREC_FP16,
                                   [&](const std::vector<uint16_t>& v) { return arena_put(h, v); }, [&](const std::vector<float>& v) { return arena_put(h, v); });
  }
  h->csar_co_us_scaled = UmmaWeights{};
  if (h->fp32_split) {   // split tensor-core tail: same channel-branch scaling as the 16-bit tail (exact: a power of two)
    const std::vector<float>& w = W(h, "rdn.csar.conv_out.weight");   // [F][2F]
    std::vector<float> pw((size_t)2 * F * F);
    for (int co = 0; co < F; ++co)
      for (int ci = 0; ci < 2 * F; ++ci) pw[(size_t)ci * F + co] = (ci < F ? kCsarChanScale : 1.f) * w[(size_t)co * 2 * F + ci];
    ok &= umma_pack_weights_split(h->csar_co_us_scaled, pw.data(), W(h, "rdn.csar.conv_out.bias").data(), 1, 2 * F, F,
                                  [&](const std::vector<uint16_t>& v) { return arena_put(h, v); }, [&](const std::vector<float>& v) { return arena_put(h, v); });
  }
  h->ca_w1 = arena_put(h, W(h, "rdn.csar.ca.block.2.weight"));
  h->ca_b1 = arena_put(h, W(h, "rdn.csar.ca.block.2.bias"));
  h->ca_w2 = arena_put(h, W(h, "rdn.csar.ca.block.4.weight"));
  h->ca_b2 = arena_put(h, W(h, "rdn.csar.ca.block.4.bias"));
  auto transpose_1x1 = [&](const std::string& name, int cout, int cin) {
    const std::vector<float>& w = W(h, name);
    std::vector<float> t((size_t)cin * cout);
    for (int co = 0; co < cout; ++co)
      for (int ci = 0; ci < cin; ++ci) t[(size_t)ci * cout + co] = w[(size_t)co * cin + ci];
    return arena_put(h, t);
  };
  h->sa_w1 = transpose_1x1("rdn.csar.sa.block.0.weight", 2 * F, F);
  h->sa_b1 = arena_put(h, W(h, "rdn.csar.sa.block.0.bias"));
  h->sa_w2 = transpose_1x1("rdn.csar.sa.block.2.weight", F, 2 * F);
  h->sa_b2 = arena_put(h, W(h, "rdn.csar.sa.block.2.bias"));
  h->co_w = transpose_1x1("rdn.csar.conv_out.weight", F, 2 * F);
  h->co_b = arena_put(h, W(h, "rdn.csar.conv_out.bias"));
  ok &= h->ca_w1 && h->ca_b1 && h->ca_w2 && h->ca_b2 && h->sa_w1 && h->sa_b1 && h->sa_w2 && h->sa_b2 && h->co_w && h->co_b;
  ok &= pack_conv(h, h->gff0, "rdn.gff.0", F * h->cfg.num_blocks, F, 1, true);
  if (half_mode(h) && h->gff0.u.packed && !getenv("LPSR_NO_WSPLIT")) {
    // 1x1 over the 128-channel concatenation of block outputs (|f| up to ~800 with the shipped checkpoint): weights as hi + lo.  The
    // launch is HBM bound (96 % of peak), the 8 extra MMAs per tile are free.
    const std::vector<float>& w = W(h, "rdn.gff.0.weight");   // [F][4F]
    const int cin = F * h->cfg.num_blocks;
    std::vector<float> pw((size_t)cin * F);
    for (int co = 0; co < F; ++co)
      for (int ci = 0; ci < cin; ++ci) pw[(size_t)ci * F + co] = w[(size_t)co * cin + ci];
    ok &= umma_pack_weights_wsplit(h->gff0.u, pw.data(), W(h, "rdn.gff.0.bias").data(), 1, cin, F, h->cfg.precision == LPSR_PREC_FP16,
                                   [&](const std::vector<uint16_t>& v) { return arena_put(h, v); }, [&](const std::vector<float>& v) { return arena_put(h, v); });
  }
  ok &= pack_conv(h, h->gff1, "rdn.gff.1", F, F, 3, true);
  ok &= pack_conv(h, h->fin, "final_conv", F, h->cfg.out_channels, 3, true);
  if (half_mode(h) && umma_supported(3, F, 16)) {   // tensor-core final conv: pad Cout 1 -> 16 with zero filters
    const std::vector<float>& w = W(h, "final_conv.weight");
    std::vector<float> pw((size_t)9 * F * 16, 0.f), pb(16, 0.f);
    for (int ci = 0; ci < F; ++ci)
      for (int t = 0; t < 9; ++t) pw[((size_t)t * F + ci) * 16] = w[(size_t)ci * 9 + t];
    pb[0] = W(h, "final_conv.bias")[0];
    quantize_conv_sum_preserving(pw, 9, F, 16, h->cfg.precision == LPSR_PREC_FP16);
    ok &= umma_pack_weights(h->fin_u, pw.data(), pb.data(), 3, F, 16, h->cfg.precision == LPSR_PREC_FP16,
                            [&](const std::vector<uint16_t>& v) { return arena_put(h, v); }, [&](const std::vector<float>& v) { return arena_put(h, v); });
    h->fin_rw = RowWeights{};
    ok &= rowconv_pack_weights(h->fin_rw, pw.data(), pb.data(), F, 16, h->cfg.precision == LPSR_PREC_FP16,
                               [&](const std::vector<uint16_t>& v) { return arena_put(h, v); }, [&](const std::vector<float>& v) { return arena_put(h, v); });
  }
  h->fin_us = UmmaWeights{};
  if (h->fp32_split) {   // final conv for split operands: Cout padded 1 -> 16 with zero filters
    const std::vector<float>& w = W(h, "final_conv.weight");
    std::vector<float> pw((size_t)9 * F * 16, 0.f), pb(16, 0.f);
    for (int ci = 0; ci < F; ++ci)
      for (int t = 0; t < 9; ++t) pw[((size_t)t * F + ci) * 16] = w[(size_t)ci * 9 + t];
    pb[0] = W(h, "final_conv.bias")[0];
    ok &= umma_pack_weights_split(h->fin_us, pw.data(), pb.data(), 3, F, 16, [&](const std::vector<uint16_t>& v) { return arena_put(h, v); },
                                  [&](const std::vector<float>& v) { return arena_put(h, v); });
    ok &= h->sfe2.us.packed && h->gff0.us.packed && h->gff1.us.packed && h->csar_c1.us.packed && h->csar_c2.us.packed;
    for (int r = 0; r < 2; ++r) {
      ok &= h->lff[r].us.packed;
      for (int i = 0; i < L; ++i) ok &= h->rdb[r][i].us.packed;
    }
  }
  if (!ok) return fail(h, LPSR_ERR_CUDA, "weight packing failed (arena %zu/%zu bytes): %s", h->arena.used, h->arena.cap,
                       cudaGetErrorString(cudaGetLastError()));
  h->packed = true;
  return LPSR_OK;
}

}  // namespace

namespace lpsr {
WsLayout ws_layout(const lpsr_handle* h, int B, int H, int W) {
  WsLayout L{};
  L.Hp = (H + 3) / 4 * 4;
  L.Wp = (W + 3) / 4 * 4;
  L.P = L.Hp * L.Wp;
  int S = L.P / 256;   // pooling slices per crop: a function of the crop size only, so a crop's result does not depend on B
  if (S > 64) S = 64;
  if (S < 1) S = 1;
  L.S = S;
  const size_t es = elem_size(h), BP = (size_t)B * L.P;
  size_t off = 0;
  auto take = [&](size_t elems, size_t esz) { size_t o = off; off = align_up(off + elems * esz, 256); return o; };
  L.xu = take(BP / 4 * 16, es);   // tensor-core AutoEncoder: the input crop as a half-grid space-to-depth operand
  L.c0 = take(BP * 12, es);
  L.e0 = take(BP / 4 * 48, es);
  L.e1 = take(BP / 16 * 48, es);
  L.d0 = take(BP / 4 * 16, es);   // 12 channels on the half grid; the tensor-core path pads the pitch to 16
  L.s = take(BP * 16, es);   // c0 + decoder output: 12 channels, pitch 16 (zero padded) on the tensor-core path
  L.ae = take(BP * 16, es);  // AutoEncoder output: 3 channels (pitch 3), or padded to 16 channels on the tensor-core path
  L.sfe1 = take(BP * 32, es);
  L.x0 = take(BP * 32, es);
  for (int r = 0; r < 2; ++r)
    for (int i = 0; i < 4; ++i) L.grow[r][i] = take(BP * 16, es);
  for (int i = 0; i < 4; ++i) L.f[i] = take(BP * 32, es);
  L.t = take(BP * 32, es);
  L.xin = take(BP * 32, es);
  L.g0 = take(BP * 32, es);
  L.g = take(BP * 32, es);
  // 16-bit modes: one partial per (tile, epilogue warp) of conv_in.2; a crop has at most ~1.3 x P/126 tiles (halo rows of its strips)
  L.pool_slots = (half_mode(h) || h->fp32_split) ? 4 * (2 * (L.P / 128) + 40) : 0;
  L.pool = take((size_t)B * std::max(S, L.pool_slots) * 32, 4);
  L.sc = take((size_t)B * 32, 4);
  if (half_mode(h)) {   // tensor-core CSAR tail: 64-channel hidden map and the gated concat [x_in^2*s_c | x_in*s_s]
    L.hid = take(BP * 64, es);
    L.gate = take(BP * 64, es);
  }
  L.total = off;
  return L;
}
}  // namespace lpsr

namespace {

int forward_dispatch(lpsr_handle* h, const float* x, float* y, int B, int H, int W, char* ws, cudaStream_t st, bool dry, int* nl,
                     LaunchProfile* prof = nullptr) {
  switch (h->cfg.precision) {
    case LPSR_PREC_FP32:
      if (h->fp32_split) return forward_split_entry(h, x, y, B, H, W, ws, st, dry, nl, prof);
      return forward_impl<float>(h, x, y, B, H, W, ws, st, dry, nl, prof);
    case LPSR_PREC_BF16: return forward_impl<__nv_bfloat16>(h, x, y, B, H, W, ws, st, dry, nl, prof);
    case LPSR_PREC_FP16: return forward_impl<__half>(h, x, y, B, H, W, ws, st, dry, nl, prof);
    default: return fail(h, LPSR_ERR_UNSUPPORTED, "precision mode %d not built", h->cfg.precision);
  }
}

int check_shape(lpsr_handle* h, int B, int H, int W) {
  if (B < 1 || H < 1 || W < 1) return fail(h, LPSR_ERR_INVALID_ARG, "bad shape B=%d H=%d W=%d", B, H, W);
  if (B > 65535) return fail(h, LPSR_ERR_INVALID_ARG, "B=%d exceeds the per-call limit 65535 (split the batch)", B);
  return LPSR_OK;
}

}  // namespace

namespace {
// ---- op-level entry points ------------------------------------------------------------------------------
static int pixel_remap(const float* x, float* y, int B, int C, int H, int W, void* stream, int mode) {
  if (!x || !y || B < 1 || C < 1 || H < 1 || W < 1) return fail(nullptr, LPSR_ERR_INVALID_ARG, "bad argument");
  if (mode == kShuffleDown && ((H | W) & 1)) return fail(nullptr, LPSR_ERR_INVALID_ARG, "PixelUnshuffle(2) needs even H,W");
  if (mode == kShuffleUp && (C % 4)) return fail(nullptr, LPSR_ERR_INVALID_ARG, "PixelShuffle(2) needs C %% 4 == 0");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n = (size_t)B * C * H * W;
  float *a = nullptr, *b = nullptr;
  CUDA_TRY(nullptr, cudaMalloc(&a, n * 4));
  if (cudaMalloc(&b, n * 4) != cudaSuccess) { cudaFree(a); return fail(nullptr, LPSR_ERR_CUDA, "cudaMalloc failed"); }
  const int blocks = (int)((n + 255) / 256 > 8192 ? 8192 : (n + 255) / 256);
  nchw_to_nhwc_kernel<float><<<blocks, 256, 0, st>>>(x, a, B, C, H, W, C, 0);
  if (mode == kShuffleDown) {
    pixel_remap_nhwc_kernel<kShuffleDown><<<blocks, 256, 0, st>>>(a, b, B, C, H, W);
    nhwc_to_nchw_kernel<float><<<blocks, 256, 0, st>>>(b, y, B, C * 4, H / 2, W / 2, C * 4, 0, 0);
  } else {
    pixel_remap_nhwc_kernel<kShuffleUp><<<blocks, 256, 0, st>>>(a, b, B, C, H, W);
    nhwc_to_nchw_kernel<float><<<blocks, 256, 0, st>>>(b, y, B, C / 4, H * 2, W * 2, C / 4, 0, 0);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(a);
  cudaFree(b);
  if (e != cudaSuccess) return fail(nullptr, LPSR_ERR_CUDA, "pixel remap: %s", cudaGetErrorString(e));
  return LPSR_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" {

int lpsr_abi_version(void) { return LPSR_B200_ABI_VERSION; }

const char* lpsr_last_error(const lpsr_handle* h) { return h ? h->err : g_err; }

int lpsr_create(lpsr_handle** out, const lpsr_config* cfg) {
  if (!out || !cfg) return fail(nullptr, LPSR_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  if (cfg->abi_version != LPSR_B200_ABI_VERSION) return fail(nullptr, LPSR_ERR_INVALID_ARG, "ABI version %d != %d", cfg->abi_version, LPSR_B200_ABI_VERSION);
  // kernels are specialised on the canonical ctor literals every reference call site uses
  // (inference/run.py:124, evaluation/eval.py:77, my_utils/export_onnx.py:40-47)
  if (cfg->num_channels != 3 || cfg->num_features != 32 || cfg->growth_rate != 16 || cfg->num_blocks != 4 || cfg->num_layers != 4 ||
      cfg->out_channels != 1)
    return fail(nullptr, LPSR_ERR_INVALID_ARG,
                "unsupported LPSR dims (%d,%d,%d,%d,%d,out=%d): this build is specialised on (3,32,16,4,4,out=1)", cfg->num_channels,
                cfg->num_features, cfg->growth_rate, cfg->num_blocks, cfg->num_layers, cfg->out_channels);
  if (cfg->precision < LPSR_PREC_FP32 || cfg->precision > LPSR_PREC_FP16)
    return fail(nullptr, LPSR_ERR_UNSUPPORTED, "precision mode %d not built", cfg->precision);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail(nullptr, LPSR_ERR_CUDA, "no CUDA device (%s); there is no CPU fallback", cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, LPSR_ERR_INVALID_ARG, "device %d out of range [0,%d)", cfg->device, ndev);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, cfg->device)) != cudaSuccess) return fail(nullptr, LPSR_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10) return fail(nullptr, LPSR_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major, prop.minor);
  if ((e = cudaSetDevice(cfg->device)) != cudaSuccess) return fail(nullptr, LPSR_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  lpsr_handle* h = new lpsr_handle();
  h->cfg = *cfg;
  h->sm = prop.major * 10 + prop.minor;
  h->num_sms = prop.multiProcessorCount;
  // fp32 mode = tensor cores with split (double-fp16) operands; LPSR_FP32_FFMA=1 (or LPSR_UMMA=0) keeps the CUDA-core FFMA path
  h->fp32_split = cfg->precision == LPSR_PREC_FP32 && umma_enabled() && !getenv("LPSR_FP32_FFMA");
  build_live_table(h);
  *out = h;
  return LPSR_OK;
}

int lpsr_destroy(lpsr_handle* h) {
  if (!h) return LPSR_OK;
  cudaSetDevice(h->cfg.device);
  if (h->arena.base) cudaFree(h->arena.base);
  if (h->host_x) cudaFree(h->host_x);
  if (h->host_y) cudaFree(h->host_y);
  if (h->pre_host) cudaFreeHost(h->pre_host);
  if (h->pre_dev) cudaFree(h->pre_dev);
  if (h->pre_ev) cudaEventDestroy(h->pre_ev);
  if (h->host_ws) cudaFree(h->host_ws);
  if (h->host_ws2) cudaFree(h->host_ws2);
  if (h->host_stream2) cudaStreamDestroy(h->host_stream2);
  if (h->host_stream) {
    cudaStreamDestroy(h->host_stream);
    cudaStreamDestroy(h->copy_in_stream);
    cudaStreamDestroy(h->copy_out_stream);
    for (int i = 0; i < 2 * kHostChunksMax; ++i) cudaEventDestroy(h->host_ev[i]);
  }
  delete h;
  return LPSR_OK;
}

int lpsr_device_sm(const lpsr_handle* h) { return h ? h->sm : 0; }
int lpsr_num_live_tensors(const lpsr_handle* h) { return h ? (int)h->live.size() : 0; }
const char* lpsr_live_tensor_name(const lpsr_handle* h, int32_t i) {
  return (h && i >= 0 && i < (int)h->live.size()) ? h->live[i].name.c_str() : nullptr;
}
int64_t lpsr_live_tensor_numel(const lpsr_handle* h, int32_t i) { return (h && i >= 0 && i < (int)h->live.size()) ? h->live[i].numel : -1; }

int lpsr_load_weights(lpsr_handle* h, const lpsr_tensor_desc* tensors, int32_t n) {
  if (!h || (!tensors && n > 0)) return fail(h, LPSR_ERR_INVALID_ARG, "null argument");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  for (int i = 0; i < n; ++i) {
    const lpsr_tensor_desc& t = tensors[i];
    if (!t.name || !t.data) return fail(h, LPSR_ERR_INVALID_ARG, "tensor %d: null name/data", i);
    auto it = h->live_index.find(t.name);
    if (it == h->live_index.end()) return fail(h, LPSR_ERR_INVALID_ARG, "unexpected tensor '%s' (not on the executed path)", t.name);
    LiveTensor& lt = h->live[it->second];
    if (t.numel != lt.numel) return fail(h, LPSR_ERR_INVALID_ARG, "size mismatch for '%s': got %lld, expected %lld", t.name, (long long)t.numel, (long long)lt.numel);
    lt.host.resize(lt.numel);
    if (t.on_device) CUDA_TRY(h, cudaMemcpy(lt.host.data(), t.data, lt.numel * sizeof(float), cudaMemcpyDeviceToHost));
    else memcpy(lt.host.data(), t.data, lt.numel * sizeof(float));
    lt.loaded = true;
  }
  for (const LiveTensor& lt : h->live)
    if (!lt.loaded) { h->packed = false; return LPSR_OK; }   // partial load: pack once everything arrived
  CUDA_TRY(h, cudaDeviceSynchronize());   // packed buffers may still be in use by earlier forwards
  return pack_all(h);
}

int lpsr_output_shape(const lpsr_handle* h, int32_t B, int32_t H, int32_t W, int32_t* oc, int32_t* oh, int32_t* ow) {
  if (!h) return LPSR_ERR_INVALID_ARG;
  (void)B;
  if (oc) *oc = h->cfg.out_channels;
  if (oh) *oh = (H + 3) / 4 * 4;
  if (ow) *ow = (W + 3) / 4 * 4;
  return LPSR_OK;
}

size_t lpsr_workspace_bytes(const lpsr_handle* h, int32_t B, int32_t H, int32_t W) {
  if (!h || B < 1 || H < 1 || W < 1) return 0;
  return ws_layout(h, B, H, W).total;
}

int lpsr_forward_launch_count(const lpsr_handle* h, int32_t B, int32_t H, int32_t W) {
  if (!h || !h->packed) return -1;
  int n = 0;
  forward_dispatch(const_cast<lpsr_handle*>(h), nullptr, nullptr, B, H, W, nullptr, nullptr, true, &n);
  return n;
}

int lpsr_forward(lpsr_handle* h, const float* x, float* y, int32_t B, int32_t H, int32_t W, void* ws, size_t ws_bytes, void* stream) {
  if (!h || !x || !y || !ws) return fail(h, LPSR_ERR_INVALID_ARG, "null argument");
  int rc = check_shape(h, B, H, W);
  if (rc) return rc;
  if (!h->packed) {
    std::string missing;
    for (const LiveTensor& lt : h->live) if (!lt.loaded) { missing = lt.name; break; }
    return fail(h, LPSR_ERR_NO_WEIGHTS, "forward before weights were loaded (first missing: %s)", missing.c_str());
  }
  if (ws_bytes < ws_layout(h, B, H, W).total) return fail(h, LPSR_ERR_WORKSPACE, "workspace %zu < required %zu bytes", ws_bytes, ws_layout(h, B, H, W).total);
  if (reinterpret_cast<uintptr_t>(ws) % 256) return fail(h, LPSR_ERR_WORKSPACE, "workspace must be 256-byte aligned");
  int cur = -1;
  CUDA_TRY(h, cudaGetDevice(&cur));
  if (cur != h->cfg.device) CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  rc = forward_dispatch(h, x, y, B, H, W, static_cast<char*>(ws), static_cast<cudaStream_t>(stream), false, nullptr);
  if (cur != h->cfg.device && cur >= 0) cudaSetDevice(cur);
  return rc;
}

int lpsr_forward_profiled(lpsr_handle* h, const float* x, float* y, int32_t B, int32_t H, int32_t W, void* ws, size_t ws_bytes,
                          void* stream, float* ms_out, char* names_out, int32_t name_stride, int32_t capacity) {
  if (!h || !x || !y || !ws || !ms_out) return fail(h, LPSR_ERR_INVALID_ARG, "null argument");
  int rc = check_shape(h, B, H, W);
  if (rc) return rc;
  if (!h->packed) return fail(h, LPSR_ERR_NO_WEIGHTS, "forward before weights were loaded");
  if (ws_bytes < ws_layout(h, B, H, W).total) return fail(h, LPSR_ERR_WORKSPACE, "workspace too small");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  LaunchProfile prof;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  rc = forward_dispatch(h, x, y, B, H, W, static_cast<char*>(ws), st, false, nullptr, &prof);
  cudaError_t e = cudaStreamSynchronize(st);
  const int n = (int)prof.names.size();
  if (rc == LPSR_OK && e == cudaSuccess && (int)prof.events.size() == n + 1 && n <= capacity) {
    for (int i = 0; i < n; ++i) {
      cudaEventElapsedTime(&ms_out[i], prof.events[i], prof.events[i + 1]);
      if (names_out && name_stride > 0) snprintf(names_out + (size_t)i * name_stride, name_stride, "%s", prof.names[i].c_str());
    }
  } else if (rc == LPSR_OK) {
    rc = fail(h, e != cudaSuccess ? LPSR_ERR_CUDA : LPSR_ERR_INVALID_ARG, "profiled forward: %s (launches %d, capacity %d)",
              cudaGetErrorString(e), n, capacity);
  }
  for (cudaEvent_t ev : prof.events) cudaEventDestroy(ev);
  return rc == LPSR_OK ? n : rc;
}

int lpsr_forward_host(lpsr_handle* h, const float* x_host, float* y_host, int32_t B, int32_t H, int32_t W) {
  if (!h || !x_host || !y_host) return fail(h, LPSR_ERR_INVALID_ARG, "null argument");
  int rc = check_shape(h, B, H, W);
  if (rc) return rc;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  if (!h->host_stream) {
    CUDA_TRY(h, cudaStreamCreateWithFlags(&h->host_stream, cudaStreamNonBlocking));
    CUDA_TRY(h, cudaStreamCreateWithFlags(&h->host_stream2, cudaStreamNonBlocking));
    CUDA_TRY(h, cudaStreamCreateWithFlags(&h->copy_in_stream, cudaStreamNonBlocking));
    CUDA_TRY(h, cudaStreamCreateWithFlags(&h->copy_out_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2 * kHostChunksMax; ++i) CUDA_TRY(h, cudaEventCreateWithFlags(&h->host_ev[i], cudaEventDisableTiming));
  }
  // Large batches are cut into chunks so the H2D copy of chunk i+1 and the D2H copy of chunk i-1 (separate copy streams)
  // overlap the forward of chunk i (compute stream); crops are independent, so chunking does not change any result
  // except through the batch-size independent pooling slices (bit-identical).  The first and the last chunk are small: the
  // first H2D copy and the last D2H copy are the only ones nothing overlaps.
  // (measured on B200, B = 1024, uniform chunks: 2 -> 14.5 ms, 4 -> 14.3 ms, 8 -> 15.2 ms: every chunk costs ~30 launch ramps)
  int sizes[kHostChunksMax];
  int nchunk = 0;
  const char* env = getenv("LPSR_HOST_CHUNK_SIZES");               // tuning knob: comma separated crops per chunk (last one repeats)
  if (env && env[0]) {
    int left = B, last = B;
    for (const char* q = env; left > 0 && nchunk < kHostChunksMax;) {
      int v = atoi(q);
      if (v > 0) last = v;
      const int take = (nchunk == kHostChunksMax - 1) ? left : std::min(left, last);
      sizes[nchunk++] = take;
      left -= take;
      const char* comma = strchr(q, ',');
      if (comma) q = comma + 1;
    }
  } else if (B >= 512) {
    // measured on B200 (B = 1024, ms per call): 4 x 256 -> 11.99, 64/448/448/64 -> 11.60, 128/384/384/128 -> 11.46
    const int edge = B / 8, mid = (B - 2 * edge) / 2;
    sizes[0] = edge;
    sizes[1] = mid;
    sizes[2] = B - 2 * edge - mid;
    sizes[3] = edge;
    nchunk = 4;
  } else if (B >= 128) {
    sizes[0] = B / 2; sizes[1] = B - B / 2; nchunk = 2;
  } else {
    sizes[0] = B; nchunk = 1;
  }
  int cb = 0;                                               // largest chunk sizes the workspace
  for (int i = 0; i < nchunk; ++i) cb = std::max(cb, sizes[i]);
  const WsLayout L = ws_layout(h, cb, H, W);
  const size_t x_crop = (size_t)3 * H * W * 4, y_crop = (size_t)h->cfg.out_channels * L.P * 4;
  auto grow = [&](void** p, size_t* cap, size_t need) -> cudaError_t {
    if (*cap >= need) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    cudaError_t e = cudaMalloc(p, need);
    if (e == cudaSuccess) *cap = need;
    return e;
  };
  CUDA_TRY(h, grow(&h->host_x, &h->host_x_cap, x_crop * B));
  CUDA_TRY(h, grow(&h->host_y, &h->host_y_cap, y_crop * B));
  CUDA_TRY(h, grow(&h->host_ws, &h->host_ws_cap, L.total));
  // Chunks alternate between two compute streams (each with its own workspace): the persistent kernels of chunk i+1 start on SMs the
  // kernels of chunk i have already left, so the per-launch ramp-down / ramp-up of ~28 launches per chunk overlaps instead of adding up.
  static int two_streams = -1;
  if (two_streams < 0) { const char* e2 = getenv("LPSR_HOST_STREAMS"); two_streams = (e2 && e2[0] == '1') ? 0 : 1; }
  const bool dual = two_streams && nchunk > 1;
  if (dual) CUDA_TRY(h, grow(&h->host_ws2, &h->host_ws2_cap, L.total));
  char* dx = static_cast<char*>(h->host_x);
  char* dy = static_cast<char*>(h->host_y);
  const char* hx = reinterpret_cast<const char*>(x_host);
  char* hy = reinterpret_cast<char*>(y_host);
  for (int i = 0, lo = 0; i < nchunk; lo += sizes[i], ++i) {
    const int n = sizes[i];
    if (n <= 0) continue;
    CUDA_TRY(h, cudaMemcpyAsync(dx + x_crop * lo, hx + x_crop * lo, x_crop * n, cudaMemcpyHostToDevice, h->copy_in_stream));
    CUDA_TRY(h, cudaEventRecord(h->host_ev[2 * i], h->copy_in_stream));
    cudaStream_t cs = (dual && (i & 1)) ? h->host_stream2 : h->host_stream;
    CUDA_TRY(h, cudaStreamWaitEvent(cs, h->host_ev[2 * i], 0));
    rc = lpsr_forward(h, reinterpret_cast<const float*>(dx + x_crop * lo), reinterpret_cast<float*>(dy + y_crop * lo), n, H, W,
                      (dual && (i & 1)) ? h->host_ws2 : h->host_ws, (dual && (i & 1)) ? h->host_ws2_cap : h->host_ws_cap, cs);
    if (rc) {   // earlier chunks' async copies still reference the caller's buffers: drain them before reporting the error
      cudaStreamSynchronize(h->copy_in_stream);
      cudaStreamSynchronize(h->host_stream);
      cudaStreamSynchronize(h->host_stream2);
      cudaStreamSynchronize(h->copy_out_stream);
      return rc;
    }
    CUDA_TRY(h, cudaEventRecord(h->host_ev[2 * i + 1], cs));
    CUDA_TRY(h, cudaStreamWaitEvent(h->copy_out_stream, h->host_ev[2 * i + 1], 0));
    CUDA_TRY(h, cudaMemcpyAsync(hy + y_crop * lo, dy + y_crop * lo, y_crop * n, cudaMemcpyDeviceToHost, h->copy_out_stream));
  }
  CUDA_TRY(h, cudaStreamSynchronize(h->copy_out_stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->host_stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->host_stream2));
  return LPSR_OK;
}

int lpsr_debug_read_tap(lpsr_handle* h, const char* name, float* dst, int64_t dst_numel, int32_t B, int32_t H, int32_t W, void* wsv,
                        void* stream) {
  if (!h || !name || !dst || !wsv) return fail(h, LPSR_ERR_INVALID_ARG, "null argument");
  const WsLayout L = ws_layout(h, B, H, W);
  struct Tap { const char* name; size_t off; int pitch, choff, C, div, unshuffled, ae_half; };
  const bool tc = h->ae_tc;    // tensor-core AutoEncoder: full-resolution tensors are stored as their PixelUnshuffle on the half grid
  const Tap taps[] = {
      {"ae.c0", L.c0, tc ? 48 : 12, 0, 12, 1, tc, tc},  {"ae.enc0", L.e0, 48, 0, 48, 2, 0, tc},   {"ae.enc1", L.e1, 48, 0, 48, 4, 0, tc},
      {"ae.dec0", L.d0, tc ? 16 : 12, 0, 12, 2, 0, tc}, {"ae.sum", L.s, tc ? 48 : (h->sfe1_u.packed ? 16 : 12), 0, 12, 1, tc, tc},
      {"ae.out", L.ae, tc ? 16 : (h->sfe1_u.packed ? 16 : 3), 0, 3, 1, tc, tc},
      {"rdn.sfe1", L.sfe1, 32, 0, 32, 1, 0}, {"rdn.sfe2", L.x0, 32, 0, 32, 1, 0},
      {"rdn.block0", L.f[0], 32, 0, 32, 1, 0}, {"rdn.block1", L.f[1], 32, 0, 32, 1, 0},
      {"rdn.block2", L.f[2], 32, 0, 32, 1, 0}, {"rdn.block3", L.f[3], 32, 0, 32, 1, 0},
      {"rdb0.growth2", L.grow[0][2], 16, 0, 16, 1, 0},   {"csar3.x_in", L.xin, 32, 0, 32, 1, 0},
      {"rdn.gff0", L.g0, 32, 0, 32, 1, 0},   {"rdn.out", L.g, 32, 0, 32, 1, 0}};
  for (const Tap& t : taps) {
    if (strcmp(t.name, name)) continue;
    const int Ht = L.Hp / t.div, Wt = L.Wp / t.div;
    const int64_t need = (int64_t)B * t.C * Ht * Wt;
    if (dst_numel != need) return fail(h, LPSR_ERR_INVALID_ARG, "tap '%s' has %lld elements, buffer has %lld", name, (long long)need, (long long)dst_numel);
    char* ws = static_cast<char*>(wsv);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (h->fp32_split && strncmp(t.name, "ae.", 3))   // trunk tensors of the fp32 mode are split (double-fp16) tensors
      return tap_copy_split(h, ws + t.off, dst, B, t.C, Ht, Wt, st);
    if (t.ae_half)   // tensor-core AutoEncoder intermediates are fp16 in both 16-bit modes
      return tap_copy_impl<__half>(h, ws + t.off, dst, B, t.C, Ht, Wt, t.pitch, t.choff, t.unshuffled, st);
    switch (h->cfg.precision) {
      case LPSR_PREC_FP32: return tap_copy_impl<float>(h, ws + t.off, dst, B, t.C, Ht, Wt, t.pitch, t.choff, t.unshuffled, st);
      case LPSR_PREC_BF16: return tap_copy_impl<__nv_bfloat16>(h, ws + t.off, dst, B, t.C, Ht, Wt, t.pitch, t.choff, t.unshuffled, st);
      default: return tap_copy_impl<__half>(h, ws + t.off, dst, B, t.C, Ht, Wt, t.pitch, t.choff, t.unshuffled, st);
    }
    return LPSR_OK;
  }
  return fail(h, LPSR_ERR_INVALID_ARG, "unknown tap '%s'", name);
}

int lpsr_op_pixel_unshuffle2(const float* x, float* y, int32_t B, int32_t C, int32_t H, int32_t W, void* stream) {
  return pixel_remap(x, y, B, C, H, W, stream, kShuffleDown);
}
int lpsr_op_pixel_shuffle2(const float* x, float* y, int32_t B, int32_t C, int32_t H, int32_t W, void* stream) {
  return pixel_remap(x, y, B, C, H, W, stream, kShuffleUp);
}

// profiling experiment hook (LPSR_UMMA_TRACE): copy the clock64 trace of the last traced tensor-core launch, 512 x 8 stamps
int lpsr_debug_umma_trace(long long* dst_host) {
  long long* t = umma_trace_buffer();
  if (!t || !dst_host) return LPSR_ERR_INVALID_ARG;
  return cudaMemcpy(dst_host, t, 512 * 8 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? LPSR_OK : LPSR_ERR_CUDA;
}

size_t lpsr_op_yolo_nms_workspace_bytes(int32_t B, int32_t max_candidates, int32_t max_det, int32_t n_classes) {
  if (B < 0 || max_candidates < 64 || max_candidates % 64 || max_candidates > kNmsMaxCap || max_det < 1 || n_classes < 0) return 0;
  return nms_workspace_bytes(B, max_candidates, max_det, n_classes);
}

int lpsr_op_yolo_nms(lpsr_handle* h, const float* pred, int32_t B, int32_t N, int32_t nc, float conf_thres, float iou_thres,
                     const int32_t* classes_host, int32_t n_classes, int32_t agnostic, int32_t multi_label, int32_t max_det,
                     int32_t max_candidates, float* out, int32_t* out_count, void* workspace, size_t workspace_bytes, void* stream) {
  if (B < 0 || N < 0 || nc < 1 || max_det < 1 || n_classes < 0) return fail(h, LPSR_ERR_INVALID_ARG, "yolo_nms: bad shape");
  if (!(conf_thres >= 0.f && conf_thres <= 1.f) || !(iou_thres >= 0.f && iou_thres <= 1.f))      // general.py:689-690
    return fail(h, LPSR_ERR_INVALID_ARG, "yolo_nms: thresholds must be in [0, 1]");
  if (max_candidates < 64 || max_candidates % 64 || max_candidates > kNmsMaxCap)
    return fail(h, LPSR_ERR_INVALID_ARG, "yolo_nms: max_candidates must be a multiple of 64 in [64, %d]", kNmsMaxCap);
  if (B == 0) return LPSR_OK;
  if (!pred || !out || !out_count || !workspace || (n_classes && !classes_host)) return fail(h, LPSR_ERR_INVALID_ARG, "yolo_nms: null argument");
  const size_t need = lpsr_op_yolo_nms_workspace_bytes(B, max_candidates, max_det, n_classes);
  char* ws = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) / 256 * 256);
  if (workspace_bytes < need) return fail(h, LPSR_ERR_INVALID_ARG, "yolo_nms: workspace of %zu bytes, %zu needed", workspace_bytes, need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  NmsParams p{};
  p.pred = pred; p.B = B; p.N = N; p.nc = nc;
  p.conf_thres = conf_thres; p.iou_thres = iou_thres;
  p.classes = nullptr; p.n_classes = n_classes;
  if (n_classes) {
    CUDA_TRY(h, cudaMemcpyAsync(ws, classes_host, (size_t)n_classes * 4, cudaMemcpyHostToDevice, st));
    p.classes = reinterpret_cast<const int*>(ws);
  }
  ws += align_up((size_t)n_classes * 4, 256);
  p.agnostic = agnostic ? 1 : 0; p.multi_label = multi_label ? 1 : 0; p.max_det = max_det;
  p.cap = max_candidates;
  const size_t cap = (size_t)max_candidates;
  // per-image slices, array by array
  p.det = reinterpret_cast<float*>(ws);                            ws += align_up((size_t)B * cap * 6 * 4, 256);
  p.order = reinterpret_cast<int*>(ws);                            ws += align_up((size_t)B * cap * 4, 256);
  p.keep = reinterpret_cast<int*>(ws);                             ws += align_up((size_t)B * max_det * 4, 256);
  p.mask = reinterpret_cast<unsigned long long*>(ws);
  p.out = out; p.out_count = out_count;
  const size_t smem = cap * 5 * sizeof(float);
  static bool configured[kMaxDevices] = {};
  bool* flag = func_configured_flag(configured);
  if (!flag || !*flag) {
    CUDA_TRY(h, cudaFuncSetAttribute(yolo_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kNmsMaxCap * 5 * sizeof(float))));
    if (flag) *flag = true;
  }
  yolo_nms_kernel<<<B, kNmsThreads, smem, st>>>(p);
  CUDA_TRY(h, cudaGetLastError());
  return LPSR_OK;
}

int lpsr_preprocess_resize(lpsr_handle* h, const uint8_t* crops, const int64_t* offsets, const int32_t* heights, const int32_t* widths,
                           int32_t B, int32_t out_h, int32_t out_w, float* x_out, void* stream) {
  if (!h || !crops || !offsets || !heights || !widths || !x_out) return fail(h, LPSR_ERR_INVALID_ARG, "null argument");
  if (B < 0 || out_h < 1 || out_w < 1) return fail(h, LPSR_ERR_INVALID_ARG, "bad shape");
  if (B == 0) return LPSR_OK;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // host side: Pillow's weight tables (double precision, Pillow's operation order) per distinct source size, crop descriptors
  std::vector<int> tab;
  std::vector<ResampleCrop> desc((size_t)B);
  std::map<std::pair<int, int>, std::array<int, 3>> cache;   // (in, out) -> {bounds offset, weights offset, ksize}
  auto coeffs = [&](int in_size, int out_size) {
    auto it = cache.find({in_size, out_size});
    if (it != cache.end()) return it->second;
    std::array<int, 3> r{};
    resample_coeffs(in_size, out_size, tab, r[0], r[1], r[2]);
    cache[{in_size, out_size}] = r;
    return r;
  };
  long long tmp_total = 0;
  for (int i = 0; i < B; ++i) {
    const int H = heights[i], W = widths[i];
    if (H < 1 || W < 1) return fail(h, LPSR_ERR_INVALID_ARG, "crop %d has size %dx%d", i, H, W);
    ResampleCrop& d = desc[i];
    d.src = offsets[i]; d.H = H; d.W = W;
    d.need_h = (W != out_w); d.need_v = (H != out_h);
    d.y0 = 0; d.rows = H;
    if (d.need_v) {
      const auto v = coeffs(H, out_h);
      d.vb = v[0]; d.vk = v[1]; d.vks = v[2];
    }
    if (d.need_h) {
      const auto hc = coeffs(W, out_w);
      d.hb = hc[0]; d.hk = hc[1]; d.hks = hc[2];
    }
    d.tmp = tmp_total;
    tmp_total += (long long)H * out_w * 3;
  }
  // Pillow's horizontal pass only produces the rows the vertical pass reads; with the full-image box that is every row whose
  // weight can be non-zero, so processing all H rows gives the same values (the vertical bounds index the unshifted rows)
  const size_t tab_bytes = tab.size() * sizeof(int), desc_bytes = desc.size() * sizeof(ResampleCrop);
  const size_t need_dev = align_up(tab_bytes, 256) + align_up(desc_bytes, 256) + (size_t)tmp_total;
  if (h->pre_ev) CUDA_TRY(h, cudaEventSynchronize(h->pre_ev));            // the previous call's kernel is done with the staging buffer, the tables and the scratch
  else CUDA_TRY(h, cudaEventCreateWithFlags(&h->pre_ev, cudaEventDisableTiming));
  if (h->pre_host_cap < tab_bytes + desc_bytes) {
    if (h->pre_host) cudaFreeHost(h->pre_host);
    h->pre_host = nullptr; h->pre_host_cap = 0;
    CUDA_TRY(h, cudaMallocHost(&h->pre_host, (tab_bytes + desc_bytes) * 2));
    h->pre_host_cap = (tab_bytes + desc_bytes) * 2;
  }
  if (h->pre_dev_cap < need_dev) {
    CUDA_TRY(h, cudaDeviceSynchronize());                                  // earlier launches may still read the old buffer
    if (h->pre_dev) cudaFree(h->pre_dev);
    h->pre_dev = nullptr; h->pre_dev_cap = 0;
    CUDA_TRY(h, cudaMalloc(&h->pre_dev, need_dev * 2));
    h->pre_dev_cap = need_dev * 2;
  }
  char* hp = static_cast<char*>(h->pre_host);
  memcpy(hp, tab.data(), tab_bytes);
  memcpy(hp + tab_bytes, desc.data(), desc_bytes);
  char* dp = static_cast<char*>(h->pre_dev);
  int* d_tab = reinterpret_cast<int*>(dp);
  ResampleCrop* d_desc = reinterpret_cast<ResampleCrop*>(dp + align_up(tab_bytes, 256));
  uint8_t* d_tmp = reinterpret_cast<uint8_t*>(dp + align_up(tab_bytes, 256) + align_up(desc_bytes, 256));
  CUDA_TRY(h, cudaMemcpyAsync(d_tab, hp, tab_bytes, cudaMemcpyHostToDevice, st));
  CUDA_TRY(h, cudaMemcpyAsync(d_desc, hp + tab_bytes, desc_bytes, cudaMemcpyHostToDevice, st));
  preprocess_resize_kernel<<<B, 256, 0, st>>>(crops, d_desc, d_tab, d_tmp, x_out, out_h, out_w);
  CUDA_TRY(h, cudaGetLastError());
  // recorded AFTER the kernel: the next call (on any stream) waits until this one has finished reading the tables, the descriptors and
  // the scratch before it overwrites them, not only until the staging buffer was copied
  CUDA_TRY(h, cudaEventRecord(h->pre_ev, st));
  return LPSR_OK;
}

int lpsr_op_conv2d(lpsr_handle* h, const float* x, const float* w, const float* bias, float* y, int32_t B, int32_t Cin, int32_t Cout,
                   int32_t ks, int32_t H, int32_t W, int32_t relu, void* stream) {
  if (!h || !x || !w || !y) return fail(h, LPSR_ERR_INVALID_ARG, "null argument");
  if ((ks != 1 && ks != 3) || Cin % 16 || Cin < 16 || Cin > 16 * kMaxChunks || (Cout != 16 && Cout != 32 && Cout != 64))
    return fail(h, LPSR_ERR_UNSUPPORTED, "op_conv2d supports ks in {1,3}, Cin %% 16 == 0 (<= %d), Cout in {16,32,64}", 16 * kMaxChunks);
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (h->cfg.precision) {
    case LPSR_PREC_FP32: return op_conv_impl<float>(h, x, w, bias, y, B, Cin, Cout, ks, H, W, relu, st);
    case LPSR_PREC_BF16: return op_conv_impl<__nv_bfloat16>(h, x, w, bias, y, B, Cin, Cout, ks, H, W, relu, st);
    case LPSR_PREC_FP16: return op_conv_impl<__half>(h, x, w, bias, y, B, Cin, Cout, ks, H, W, relu, st);
    default: return fail(h, LPSR_ERR_UNSUPPORTED, "precision mode %d not built", h->cfg.precision);
  }
}

}  // extern "C"
