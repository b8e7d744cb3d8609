/*
 * lpsr_b200.h -- C ABI of the B200-native LPSR forward pass (sm_100a).
 *
 * The reference (nhanth301/License-Plate-Detection-and-Recognition-with-Image-Enhancement) has no
 * FFI of its own: its boundary for this path is the Python class my_models.lpsr.LPSR used as a stock
 * torch nn.Module (my_models/lpsr.py:252-274), whose arithmetic is done by ATen (oneDNN / cuDNN).
 * This header is the boundary that replaces that ATen layer: plain pointers and sizes, no torch types.
 * Each entry point names the reference interface it stands in for.  The Python mirror of the
 * reference class (same ctor / forward / state_dict layout) binds these with ctypes; the binding a
 * maintainer of the reference would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns LPSR_OK (0) or a negative lpsr_status; nothing throws across the ABI;
 *     lpsr_last_error() gives the message for the last failure on that handle (or globally).
 *   - device buffers are owned by the caller (PyTorch's caching allocator in the Python host);
 *     the handle owns only the packed weights.  lpsr_forward() allocates nothing.
 *   - all work is stream ordered on the cudaStream_t passed as void* (0 = legacy default stream)
 *     and is CUDA-graph capturable.
 *   - there is NO CPU fallback: every entry point fails with LPSR_ERR_CUDA if no sm_100 device.
 */
#ifndef LPSR_B200_H_
#define LPSR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LPSR_B200_ABI_VERSION 1

typedef enum lpsr_status {
  LPSR_OK = 0,
  LPSR_ERR_INVALID_ARG = -1,   /* bad pointer / shape / unsupported ctor dims              */
  LPSR_ERR_CUDA = -2,          /* CUDA runtime/driver error (message has cudaGetErrorString) */
  LPSR_ERR_NO_WEIGHTS = -3,    /* forward before all live tensors were loaded               */
  LPSR_ERR_WORKSPACE = -4,     /* workspace too small / misaligned                          */
  LPSR_ERR_UNSUPPORTED = -5    /* feature not built for this mode                           */
} lpsr_status;

/* Arithmetic mode of the dense convolutions.
 *   LPSR_PREC_FP32 : fp32 storage, fp32 FFMA accumulate -- parity mode, max|err| <= 1e-4 vs reference CPU fp32
 *   LPSR_PREC_BF16 : bf16 NHWC activations + bf16 weights on tcgen05 tensor cores, fp32 accumulate in TMEM
 *   LPSR_PREC_FP16 : same kernels with IEEE fp16 operands (8x finer mantissa; saturating convert)
 *   LPSR_PREC_FP32_SPLIT : fp32 accuracy on tensor cores: bf16 hi/lo operand split, 3 MMAs per product
 */
typedef enum lpsr_precision {
  LPSR_PREC_FP32 = 0,
  LPSR_PREC_BF16 = 1,
  LPSR_PREC_FP16 = 2,
  LPSR_PREC_FP32_SPLIT = 3
} lpsr_precision;

/* Mirrors the ctor of the reference class LPSR(num_channels, num_features, growth_rate, num_blocks,
 * num_layers, scale_factor, out_channels=1)  (my_models/lpsr.py:253-255).  scale_factor is accepted and
 * ignored exactly as in the reference (the UpScaling stage is commented out, lpsr.py:261,272). */
typedef struct lpsr_config {
  int32_t abi_version;   /* LPSR_B200_ABI_VERSION */
  int32_t device;        /* CUDA device ordinal */
  int32_t num_channels;  /* 3  */
  int32_t num_features;  /* 32 */
  int32_t growth_rate;   /* 16 */
  int32_t num_blocks;    /* 4  (only rdbs[0..num_blocks-1] execute, lpsr.py:220-222) */
  int32_t num_layers;    /* 4  */
  int32_t out_channels;  /* 1  */
  int32_t precision;     /* lpsr_precision */
  int32_t reserved[7];
} lpsr_config;

/* One named fp32 tensor of the reference state_dict (host or device pointer, contiguous, PyTorch
 * layout: conv weight [Cout,Cin/groups,kh,kw], linear weight [out,in], alpha 0-dim). */
typedef struct lpsr_tensor_desc {
  const char* name;      /* reference state_dict key of a LIVE tensor, e.g. "rdn.rdbs.0.layers.1.conv.weight" */
  const float* data;
  int64_t numel;
  int32_t on_device;     /* 0 = host pointer, 1 = device pointer on the handle's device */
  int32_t reserved;
} lpsr_tensor_desc;

typedef struct lpsr_handle lpsr_handle;

/* -- lifetime ----------------------------------------------------------------------------------- */
/* replaces: LPSR.__init__ + .to(device)                     (lpsr.py:253-267, inference/run.py:124) */
int lpsr_create(lpsr_handle** out, const lpsr_config* cfg);
int lpsr_destroy(lpsr_handle* h);
/* replaces: nn.Module.load_state_dict (strict) for the 64 live tensors (run.py:125-127).  Packs the
 * weights to kernel layout, folds each RDB's alpha into its lff weights/bias (lpsr.py:58-61).  May be
 * called again after parameters change; names not on the executed path are rejected. */
int lpsr_load_weights(lpsr_handle* h, const lpsr_tensor_desc* tensors, int32_t n_tensors);
/* number of live tensors the handle expects, and the i-th name / element count (for strict checking) */
int lpsr_num_live_tensors(const lpsr_handle* h);
const char* lpsr_live_tensor_name(const lpsr_handle* h, int32_t i);
int64_t lpsr_live_tensor_numel(const lpsr_handle* h, int32_t i);

/* -- forward ------------------------------------------------------------------------------------ */
/* output spatial size: H,W rounded up to a multiple of 4, never cropped back (lpsr.py:107-111) */
int lpsr_output_shape(const lpsr_handle* h, int32_t B, int32_t H, int32_t W,
                      int32_t* outC, int32_t* outH, int32_t* outW);
/* bytes of device scratch lpsr_forward needs for a [B,3,H,W] batch (256-byte aligned base required) */
size_t lpsr_workspace_bytes(const lpsr_handle* h, int32_t B, int32_t H, int32_t W);
/* replaces: LPSR.forward (lpsr.py:269-274) on device tensors.
 *   x_nchw : device fp32 [B,num_channels,H,W] contiguous, values in [0,1]
 *   y_nchw : device fp32 [B,out_channels,ceil4(H),ceil4(W)] contiguous */
int lpsr_forward(lpsr_handle* h, const float* x_nchw, float* y_nchw,
                 int32_t B, int32_t H, int32_t W,
                 void* workspace, size_t workspace_bytes, void* cuda_stream);
/* lpsr_forward with a cudaEvent between every launch: fills ms_out[i] (device time of launch i, warm, serialised
 * on the stream) and names_out[i*name_stride] ("<layer>:<kernel>") and returns the number of launches (>0) or a
 * negative status.  Synchronises the stream.  Used by bench.py for the live per-kernel roofline. */
int lpsr_forward_profiled(lpsr_handle* h, const float* x_nchw, float* y_nchw, int32_t B, int32_t H, int32_t W,
                          void* workspace, size_t workspace_bytes, void* cuda_stream,
                          float* ms_out, char* names_out, int32_t name_stride, int32_t capacity);
/* replaces: `sr_model(x.to(device)).cpu()` at the reference call site (inference/run.py:201-202):
 * HOST fp32 buffers in and out; the H2D copy, forward and D2H copy are issued on an internal stream
 * (scratch owned by the handle, grown on demand) and the call returns when y_host is valid. */
int lpsr_forward_host(lpsr_handle* h, const float* x_host, float* y_host,
                      int32_t B, int32_t H, int32_t W);

/* -- introspection / errors --------------------------------------------------------------------- */
/* number of kernels one lpsr_forward launches for this shape (what bench.py reports as gpu_launches) */
int lpsr_forward_launch_count(const lpsr_handle* h, int32_t B, int32_t H, int32_t W);
/* copy a named intermediate of the LAST forward out of the workspace as fp32 NCHW (tests only):
 * names follow tests/golden/taps_*.npz ("ae.c0", "ae.out", "rdn.sfe1", "rdn.block0", ...). */
int lpsr_debug_read_tap(lpsr_handle* h, const char* name, float* dst_nchw_dev, int64_t dst_numel,
                        int32_t B, int32_t H, int32_t W, void* workspace, void* cuda_stream);
/* profiling builds only (-DLPSR_UMMA_TRACE_BUILD, env LPSR_UMMA_TRACE=1): copies the 512 x 8 clock64 stamps the last traced
 * tensor-core launch recorded for CTA 0 to host memory; LPSR_ERR_INVALID_ARG when tracing is not compiled in / enabled. */
int lpsr_debug_umma_trace(long long* dst_host);
const char* lpsr_last_error(const lpsr_handle* h);   /* h may be NULL: last global (create-time) error */
int lpsr_abi_version(void);
/* compute capability major*10+minor of the handle's device, 0 on error */
int lpsr_device_sm(const lpsr_handle* h);

/* -- pre-processing in front of the forward ------------------------------------- */
/* Batched replacement of the reference's per-plate `preprocess_for_sr` (inference/run.py:80-96): B BGR uint8 crops of arbitrary
 * sizes (packed HWC, device memory; crop i starts at byte offsets[i], heights[i] x widths[i]; the three arrays are host memory)
 * -> RGB -> Pillow's antialiased bicubic resize to out_w x out_h (bit-exact with PIL.Image.resize(..., Image.BICUBIC)) -> / 255
 * -> fp32 NCHW [B,3,out_h,out_w] on the device, i.e. exactly the tensor the reference feeds to LPSR.forward.  Asynchronous on
 * `cuda_stream`; the host arrays may be reused when the call returns. */
int lpsr_preprocess_resize(lpsr_handle* h, const uint8_t* crops_bgr_dev, const int64_t* offsets, const int32_t* heights,
                           const int32_t* widths, int32_t B, int32_t out_h, int32_t out_w, float* x_nchw_dev, void* cuda_stream);

/* -- operator-level entry points (unit parity tests; same kernels the forward uses) -------------- */
/* nn.PixelUnshuffle(2) / nn.PixelShuffle(2) on device fp32 NCHW tensors (lpsr.py:72,79,88,95) through the
 * NHWC address maps the fused DConv kernels use; bit-exact index remaps. */
int lpsr_op_pixel_unshuffle2(const float* x_nchw, float* y_nchw, int32_t B, int32_t C, int32_t H, int32_t W, void* cuda_stream);
int lpsr_op_pixel_shuffle2(const float* x_nchw, float* y_nchw, int32_t B, int32_t C, int32_t H, int32_t W, void* cuda_stream);
/* nn.Conv2d(Cin,Cout,k,padding='same') (+ReLU) on device fp32 NCHW tensors, dispatched to the kernel the
 * handle's precision mode would use for that shape (lpsr.py:34-36 and the other conv sites). */
int lpsr_op_conv2d(lpsr_handle* h, const float* x_nchw, const float* w_oihw, const float* bias,
                   float* y_nchw, int32_t B, int32_t Cin, int32_t Cout, int32_t ksize,
                   int32_t H, int32_t W, int32_t relu, void* cuda_stream);


/* -- detection post-processing around the path ---------------------------------- */
/* replaces: `non_max_suppression(pred, conf_thres, iou_thres, classes, agnostic, multi_label, max_det=...)` of
 * yolov5/utils/general.py:677-760 as called per plate by my_models/detection.py:44-54, for a whole batch of images in one launch.
 *   pred_dev      : device fp32 [B,N,5+nc] (cx, cy, w, h, objectness, class scores) -- the YOLOv5 head's output; not modified
 *   classes_host  : optional class filter (host int32[n_classes], NULL = none)
 *   out_dev       : device fp32 [B,max_det,6] rows (x1, y1, x2, y2, conf, cls), image b's first out_count_dev[b] rows are valid,
 *                   in the reference's order (descending confidence; equal confidences in candidate order)
 *   out_count_dev : device int32 [B]; -1 = the image had more than max_candidates detections before NMS (nothing is truncated
 *                   silently: call again with a larger max_candidates, a multiple of 64, at most 8192)
 * Bit-identical to the reference on CPU (same float32 operations in the same order).  Asynchronous on `cuda_stream`. */
size_t lpsr_op_yolo_nms_workspace_bytes(int32_t B, int32_t max_candidates, int32_t max_det, int32_t n_classes);
int lpsr_op_yolo_nms(lpsr_handle* h, const float* pred_dev, int32_t B, int32_t N, int32_t nc, float conf_thres, float iou_thres,
                     const int32_t* classes_host, int32_t n_classes, int32_t agnostic, int32_t multi_label, int32_t max_det,
                     int32_t max_candidates, float* out_dev, int32_t* out_count_dev, void* workspace, size_t workspace_bytes,
                     void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* LPSR_B200_H_ */
