// engine_internal.h -- types shared by engine.cu (C ABI, packing) and the per-dtype forward instantiations.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/lpsr_b200.h"
#include "common.cuh"
#include "umma_weights.h"
#include "rowconv_weights.h"

namespace lpsr {

struct LiveTensor {
  std::string name;
  int64_t numel;
  std::vector<float> host;   // raw fp32 as given (PyTorch layout)
  bool loaded = false;
};

struct DeviceArena {       // one cudaMalloc for all packed weights
  void* base = nullptr;
  size_t cap = 0, used = 0;
};

// packed weights of one dense conv
struct ConvW {
  int ks = 0, cin = 0, cout = 0;
  float* w = nullptr;      // [ks*ks][cin][cout] fp32 (CUDA-core kernels)
  float* b = nullptr;      // [cout] fp32 (nullptr: no bias)
  UmmaWeights u;           // tensor-core packing (16-bit modes)
  RowWeights rw;           // row-streaming 3x3 kernel (rowconv.cuh), 16-bit modes
  UmmaWeights us;          // tensor-core packing for split (double-fp16) operands: the fp32-accuracy mode (forward_split.cuh)
};

// Tensor-core CSAR tail: the channel branch x_in^2 * s_c (lpsr.py:133-135,182-184) reaches a few thousand with the shipped checkpoint
// (|x_in| ~ 100).  It is carried as x_in^2 * (s_c / 16) with conv_out's first 32 input columns multiplied by 16 at pack time: exact
// (powers of two) in bf16 and fp16, and it keeps 16x more headroom below the fp16 maximum (65504) on adversarial inputs.
constexpr float kCsarChanScale = 16.f;

struct DConvW { float *dw_w, *dw_b, *pw_w, *pw_b; int cin, cout; };

}  // namespace lpsr

constexpr int kHostChunksMax = 8;   // lpsr_forward_host: H2D / forward / D2H pipeline depth

struct lpsr_handle {
  lpsr_config cfg{};
  int sm = 0;
  int num_sms = 0;
  std::vector<lpsr::LiveTensor> live;
  std::map<std::string, int> live_index;
  bool packed = false;
  lpsr::DeviceArena arena;
  // packed layers
  lpsr::ConvW ae_in, ae_out, sfe1, sfe2, rdb[2][4], lff[2], csar_c1, csar_c2, csar_sa1, csar_sa2, csar_co, gff0, gff1, fin;
  lpsr::UmmaWeights fin_u;   // final conv with Cout padded 1 -> 16 for the tensor-core path
  lpsr::RowWeights fin_rw;   // the same for the row-streaming kernel (rowconv.cuh)
  lpsr::UmmaWeights fin_us;  // the same for split operands
  lpsr::UmmaWeights sfe1_us;   // shallowF1 7x7 for the fp32-accuracy mode: 8-slot pixels [hi(3) | lo(3) | 0 0], hi + lo weights (56 K-steps)
  lpsr::UmmaWeights csar_co_us_scaled;   // conv_out for the split tensor-core tail: channel-branch rows x kCsarChanScale
  bool fp32_split = false;   // LPSR_PREC_FP32 runs its dense trunk layers on tensor cores with split operands (LPSR_FP32_FFMA=1: CUDA cores)
  lpsr::UmmaWeights sfe1_u;  // shallowF1 7x7 as 28 pixel-pair K-steps over an 8-channel padded input (tensor-core path)
  lpsr::UmmaWeights ae_out_u; // AutoEncoder conv_out 12 -> 3 on tensor cores: Cin padded to 16, Cout padded to 16 (zeros)
  // Tensor-core AutoEncoder (16-bit modes): every stage is ONE dense convolution on the half / quarter grid, the pixel
  // (un)shuffles are channel relabelings of space-to-depth operands (DESIGN.md "AutoEncoder on tensor cores")
  lpsr::UmmaWeights rdb_fused[2];   // last dense layer + lff + residual of each RDB as one tensor-core launch (16-bit modes)
  bool ae_tc = false;
  lpsr::UmmaWeights aet_in, aet_enc0, aet_enc1, aet_dec0, aet_dec1, aet_out, aet_sfe1;
  lpsr::DConvW dc[4];
  float *ca_w1 = nullptr, *ca_b1 = nullptr, *ca_w2 = nullptr, *ca_b2 = nullptr;
  float *sa_w1 = nullptr, *sa_b1 = nullptr, *sa_w2 = nullptr, *sa_b2 = nullptr, *co_w = nullptr, *co_b = nullptr;
  // host-call path (lpsr_forward_host)
  cudaStream_t host_stream = nullptr, host_stream2 = nullptr, copy_in_stream = nullptr, copy_out_stream = nullptr;
  cudaEvent_t host_ev[16] = {};
  void* host_x = nullptr; void* host_y = nullptr; void* host_ws = nullptr; void* host_ws2 = nullptr;
  size_t host_x_cap = 0, host_y_cap = 0, host_ws_cap = 0, host_ws2_cap = 0;
  // lpsr_preprocess_resize: pinned staging + device buffer for the weight tables, crop descriptors and the inter-pass scratch
  void* pre_host = nullptr; void* pre_dev = nullptr; size_t pre_host_cap = 0, pre_dev_cap = 0;
  cudaEvent_t pre_ev = nullptr;
  char err[512] = "";
};

namespace lpsr {

int fail(lpsr_handle* h, int code, const char* fmt, ...);

struct LaunchProfile {            // filled by forward_impl when profiling: events[i] precedes launch i, +1 closing event
  std::vector<cudaEvent_t> events;
  std::vector<std::string> names;
};

#define CUDA_TRY(h, expr)                                                                         \
  do {                                                                                            \
    cudaError_t e_ = (expr);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      return fail(h, LPSR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline bool half_mode(const lpsr_handle* h) { return h->cfg.precision == LPSR_PREC_BF16 || h->cfg.precision == LPSR_PREC_FP16; }
inline size_t elem_size(const lpsr_handle* h) { return half_mode(h) ? 2 : 4; }


// ---- workspace layout ----------------------------------------------------------------------------------
// Trunk tensors: every activation is its OWN dense NHWC tensor (pitch == channels), so each layer streams contiguous
// rows in and out:  x0 = sfe2 out (32 ch) = RDB#0 input;  grow[r][i] = growth i of RDB r (16 ch each);
// f[0..3] = block0..3 outputs (32 ch each; f[1] is also RDB#2's input).  Dense concatenation (lpsr.py:39-40) and the final
// torch.cat (lpsr.py:224) are channel-chunk gather lists over these tensors (ConvParams::chunk_ptr), never copies.
struct WsLayout {
  size_t xu, c0, e0, e1, d0, s, ae, sfe1, x0, f[4], grow[2][4], t, xin, g0, g, pool, hid, gate, sc, total;
  int Hp, Wp, P, S;
  int pool_slots;   // 16-bit modes: capacity (per crop) of the pooled partial sums written by conv_in.2's epilogue (0: not used)
};

WsLayout ws_layout(const lpsr_handle* h, int B, int H, int W);

// per-dtype entry points, explicitly instantiated in inst_{f32,bf16,f16}.cu
template <typename T>
int forward_impl(lpsr_handle* h, const float* x, float* y, int B, int H, int W, char* ws, cudaStream_t st, bool dry, int* n_launch,
                 LaunchProfile* prof);
int forward_split_entry(lpsr_handle* h, const float* x, float* y, int B, int H, int W, char* ws, cudaStream_t st, bool dry, int* n_launch,
                        LaunchProfile* prof);
int tap_copy_split(lpsr_handle* h, const void* src, float* dst, int B, int C, int H, int W, cudaStream_t st);
template <typename T>
int op_conv_impl(lpsr_handle* h, const float* x, const float* w_oihw, const float* bias, float* y, int B, int Cin, int Cout, int ks,
                 int H, int W, int relu, cudaStream_t st);
template <typename T>
int tap_copy_impl(lpsr_handle* h, const void* src, float* dst, int B, int C, int H, int W, int pitch, int off, int unshuffled,
                  cudaStream_t st);

}  // namespace lpsr
