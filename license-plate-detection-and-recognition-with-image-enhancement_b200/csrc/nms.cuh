// nms.cuh -- YOLOv5 detection post-processing on the device: the reference's `non_max_suppression`
// (/root/reference/yolov5/utils/general.py:677-760, called per plate by my_models/detection.py:44-54) for a whole batch of images in ONE
// launch, one CTA per image.  Index / compare work: every float operation is the reference's float32 operation in the reference's order
// (round-to-nearest intrinsics, no fused multiply-add), so the result is bit-identical to the CPU reference, including the order of the
// kept boxes (descending score, lower candidate index first among equal scores -- torchvision.ops.nms' stable sort).
//
// Phases of a CTA (all state of an image lives in its slice of a caller-provided workspace):
//   A  candidate rows (objectness > conf_thres on the ORIGINAL prediction, general.py:686) -> detections in (row, class) order (multi-label,
//      :726-728) or best class per row (:729-731), width/height constraint (:704), class filter (:734-735); a block-wide exclusive scan keeps
//      the reference's order.  More detections than `cap`: the image is flagged (count = -1), never truncated silently.
//   B  stable descending rank by score (O(n^2) compares against the scores in shared memory; n is a few hundred for a plate).
//   C  suppression bit matrix in rank order: mask[i] bit j = IoU(i, j) > iou_thres for j > i, torchvision's CPU formula.
//   D  one warp walks the ranks: a box not yet removed is kept and ORs its mask row into the removed set (:750-753, max_det cut).
//   E  kept rows -> out[b][k] = (x1, y1, x2, y2, conf, cls).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lpsr {

constexpr int kNmsThreads = 1024;
constexpr float kNmsMinWh = 2.f, kNmsMaxWh = 7680.f;   // general.py:692

struct NmsParams {
  const float* pred;       // [B][N][5 + nc]
  int B, N, nc;
  float conf_thres, iou_thres;
  const int* classes;      // device, n_classes entries (nullptr: no class filter)
  int n_classes;
  int agnostic, multi_label, max_det;
  int cap;                 // detections per image the workspace holds
  float* det;              // [B][cap][6]   detections in the reference's order, boxes WITHOUT the class offset
  int* order;              // [B][cap]      rank -> detection index
  unsigned long long* mask;  // [B][cap][cap / 64]
  int* keep;               // [B][max_det]
  float* out;              // [B][max_det][6]
  int* out_count;          // [B]  (-1: more than cap candidates)
};

constexpr int kNmsMaxCap = 8192;   // 4 removed-words per lane in phase D; 160 KB of shared memory for the ranked boxes

// workspace: [classes | det B x cap x 6 | order B x cap | keep B x max_det | mask B x cap x cap/64], every array 256-byte aligned
inline size_t nms_align(size_t v) { return (v + 255) / 256 * 256; }
inline size_t nms_workspace_bytes(int B, int cap, int max_det, int n_classes) {
  return 256 + nms_align((size_t)n_classes * 4) + nms_align((size_t)B * cap * 6 * 4) + nms_align((size_t)B * cap * 4) + nms_align((size_t)B * max_det * 4) +
         nms_align((size_t)B * cap * (cap / 64) * 8);
}

__device__ __forceinline__ bool nms_class_ok(const int* classes, int n_classes, int j) {
  if (!classes) return true;
  for (int k = 0; k < n_classes; ++k)
    if (classes[k] == j) return true;
  return false;
}

__global__ void __launch_bounds__(kNmsThreads, 1) yolo_nms_kernel(const NmsParams p) {
  extern __shared__ float nms_smem[];                 // [cap] scores (phase B), then [cap][5] ranked boxes (phase C)
  __shared__ int s_warp[kNmsThreads / 32];
  __shared__ int s_total, s_kept;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row_len = 5 + p.nc;
  const float* pred = p.pred + (size_t)b * p.N * row_len;
  float* det = p.det + (size_t)b * p.cap * 6;
  int* order = p.order + (size_t)b * p.cap;
  const int words = p.cap / 64;
  unsigned long long* mask = p.mask + (size_t)b * p.cap * words;
  int* keep = p.keep + (size_t)b * p.max_det;
  const bool multi = p.multi_label && p.nc > 1;     // general.py:697
  if (tid == 0) { s_total = 0; s_kept = 0; }
  __syncthreads();

  // ---- phase A: detections in the reference's order ----------------------------------------------------------------------
  for (int r0 = 0; r0 < p.N; r0 += kNmsThreads) {
    const int r = r0 + tid;
    int cnt = 0;
    float obj = 0.f;
    const float* x = pred + (size_t)r * row_len;
    if (r < p.N && x[4] > p.conf_thres) {            // candidate on the original objectness (:686, :705)
      const float w = x[2], h = x[3];
      obj = (w < kNmsMinWh || h < kNmsMinWh || w > kNmsMaxWh || h > kNmsMaxWh) ? 0.f : x[4];   // :704
      if (multi) {
        for (int j = 0; j < p.nc; ++j)
          if (__fmul_rn(x[5 + j], obj) > p.conf_thres && nms_class_ok(p.classes, p.n_classes, j)) ++cnt;
      } else {
        float best = __fmul_rn(x[5], obj);
        int bj = 0;
        for (int j = 1; j < p.nc; ++j) {
          const float c = __fmul_rn(x[5 + j], obj);
          if (c > best) { best = c; bj = j; }          // first maximum
        }
        cnt = (best > p.conf_thres && nms_class_ok(p.classes, p.n_classes, bj)) ? 1 : 0;
      }
    }
    // exclusive scan of cnt over the 1024 rows of this pass
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int v = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
      }
      s_warp[lane] = v;                               // inclusive over warps
    }
    __syncthreads();
    const int base = s_total + (warp ? s_warp[warp - 1] : 0) + incl - cnt;
    const int pass_total = s_warp[kNmsThreads / 32 - 1];
    if (cnt && base + cnt <= p.cap) {
      const float hw = __fmul_rn(x[2], 0.5f), hh = __fmul_rn(x[3], 0.5f);                      // xywh2xyxy, general.py:582-589
      const float x1 = __fsub_rn(x[0], hw), y1 = __fsub_rn(x[1], hh), x2 = __fadd_rn(x[0], hw), y2 = __fadd_rn(x[1], hh);
      int k = base;
      if (multi) {
        for (int j = 0; j < p.nc; ++j) {
          const float c = __fmul_rn(x[5 + j], obj);
          if (c > p.conf_thres && nms_class_ok(p.classes, p.n_classes, j)) {
            float* d = det + (size_t)k * 6;
            d[0] = x1; d[1] = y1; d[2] = x2; d[3] = y2; d[4] = c; d[5] = (float)j;
            ++k;
          }
        }
      } else {
        float best = __fmul_rn(x[5], obj);
        int bj = 0;
        for (int j = 1; j < p.nc; ++j) {
          const float c = __fmul_rn(x[5 + j], obj);
          if (c > best) { best = c; bj = j; }
        }
        float* d = det + (size_t)k * 6;
        d[0] = x1; d[1] = y1; d[2] = x2; d[3] = y2; d[4] = best; d[5] = (float)bj;
      }
    }
    __syncthreads();
    if (tid == 0) s_total += pass_total;
    __syncthreads();
  }
  const int n = s_total;
  if (n > p.cap) {                                     // reported to the host: the caller re-runs with a larger capacity
    if (tid == 0) p.out_count[b] = -1;
    return;
  }
  __threadfence_block();
  __syncthreads();

  // ---- phase B: stable descending rank by score ------------------------------------------------------------------------------
  float* s_score = nms_smem;
  for (int k = tid; k < n; k += kNmsThreads) s_score[k] = det[(size_t)k * 6 + 4];
  __syncthreads();
  for (int k = tid; k < n; k += kNmsThreads) {
    const float s = s_score[k];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const float t = s_score[j];
      rank += (t > s || (t == s && j < k)) ? 1 : 0;
    }
    order[rank] = k;
  }
  __threadfence_block();
  __syncthreads();

  // ---- phase C: ranked boxes (+ class offset, :749-750), areas, suppression matrix --------------------------------------
  float* s_box = nms_smem;                             // [n][5]
  for (int i = tid; i < n; i += kNmsThreads) {
    const float* d = det + (size_t)order[i] * 6;
    const float c = __fmul_rn(d[5], p.agnostic ? 0.f : kNmsMaxWh);
    const float x1 = __fadd_rn(d[0], c), y1 = __fadd_rn(d[1], c), x2 = __fadd_rn(d[2], c), y2 = __fadd_rn(d[3], c);
    s_box[i * 5 + 0] = x1; s_box[i * 5 + 1] = y1; s_box[i * 5 + 2] = x2; s_box[i * 5 + 3] = y2;
    s_box[i * 5 + 4] = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
  }
  __syncthreads();
  const int nw = (n + 63) / 64;
  for (int idx = tid; idx < n * nw; idx += kNmsThreads) {
    const int i = idx / nw, w = idx - i * nw;
    unsigned long long bits = 0ull;
    const int j0 = w * 64;
    if (j0 + 63 > i) {
      const float ix1 = s_box[i * 5], iy1 = s_box[i * 5 + 1], ix2 = s_box[i * 5 + 2], iy2 = s_box[i * 5 + 3], ia = s_box[i * 5 + 4];
      for (int jj = 0; jj < 64; ++jj) {
        const int j = j0 + jj;
        if (j <= i || j >= n) continue;
        const float xx1 = fmaxf(ix1, s_box[j * 5]), yy1 = fmaxf(iy1, s_box[j * 5 + 1]);
        const float xx2 = fminf(ix2, s_box[j * 5 + 2]), yy2 = fminf(iy2, s_box[j * 5 + 3]);
        const float ww = fmaxf(0.f, __fsub_rn(xx2, xx1)), hh = fmaxf(0.f, __fsub_rn(yy2, yy1));
        const float inter = __fmul_rn(ww, hh);
        const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ia, s_box[j * 5 + 4]), inter));   // torchvision nms_kernel.cpp
        if (ovr > p.iou_thres) bits |= 1ull << jj;
      }
    }
    mask[(size_t)i * words + w] = bits;
  }
  __threadfence_block();
  __syncthreads();

  // ---- phase D: greedy walk in rank order (one warp; lane l owns the removed words l, l + 32, ...) ------------------------
  if (warp == 0) {
    constexpr int kMaxWordsPerLane = 4;                // cap <= 8192
    unsigned long long removed[kMaxWordsPerLane] = {0ull, 0ull, 0ull, 0ull};
    int kept = 0;
    for (int i = 0; i < n && kept < p.max_det; ++i) {
      const int w = i >> 6;
      unsigned long long word = 0ull;
#pragma unroll
      for (int q = 0; q < kMaxWordsPerLane; ++q)
        if ((w >> 5) == q) word = removed[q];
      word = __shfl_sync(0xffffffffu, word, w & 31);
      if ((word >> (i & 63)) & 1ull) continue;
      if (lane == 0) keep[kept] = order[i];
      ++kept;
#pragma unroll
      for (int q = 0; q < kMaxWordsPerLane; ++q) {
        const int ww = lane + 32 * q;
        if (ww < nw) removed[q] |= mask[(size_t)i * words + ww];
      }
    }
    if (lane == 0) s_kept = kept;
  }
  __threadfence_block();
  __syncthreads();

  // ---- phase E: output -------------------------------------------------------------------------------------------------------------
  const int kept = s_kept;
  float* out = p.out + (size_t)b * p.max_det * 6;
  for (int idx = tid; idx < kept * 6; idx += kNmsThreads) out[idx] = det[(size_t)keep[idx / 6] * 6 + idx % 6];
  if (tid == 0) p.out_count[b] = kept;
}

}  // namespace lpsr
