// Shared helpers for the sm_100a kernels behind include/wssdl_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "../../include/wssdl_b200.h"

#define WSSDL_NUM_SMS 148  // B200: 2 dies x 74 SMs

#define WSSDL_RETURN_IF_CUDA(expr)              \
  do {                                          \
    cudaError_t _e = (expr);                    \
    if (_e != cudaSuccess) return (int)_e;      \
  } while (0)

// Launch errors surface here without synchronising the stream.
#define WSSDL_CHECK_LAUNCH()                    \
  do {                                          \
    cudaError_t _e = cudaGetLastError();        \
    if (_e != cudaSuccess) return (int)_e;      \
  } while (0)

// Process-wide tuning switch (capi.cu; include/wssdl_b200.h: wssdl_set_tuning).
int wssdl_tuning(int key);

// Internal forms of two public entries, used by the fused hot-path entry (hot_path.cu):
// the RoI-pool forward with image-major RoIs (roi_pool.cu) and the proposals launch with a chosen
// batch index for the padding rows behind an image's RoIs (proposal.cu).
int wssdl_roi_pool_fwd_impl(const float* bottom, const float* rois, int B, int H, int W, int C,
                            int R, int PH, int PW, float spatial_scale, int bin_mode, float* top,
                            int* argmax, void* workspace, size_t workspace_bytes,
                            wssdl_stream_t stream, int grouped_stride);
int wssdl_proposals_impl(const float* cls_prob, const float* bbox_pred, const float* im_info,
                         int info_stride, int B, int H, int W, int A, const float* base_anchors,
                         int feat_stride, int pre_nms_topN, int post_nms_topN, double nms_thresh,
                         int nms_mode, float min_size, float* rois, float* scores, int* anchor_idx,
                         int* counts, float* decoded, wssdl_stream_t stream, float pad_batch_index,
                         int one_cta_per_image);

static inline cudaStream_t to_cuda(wssdl_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// inter / den for an IoU whose numerator is very often exactly zero (boxes that do not
// touch).  __fdiv_rn sends a zero numerator down its slow path (FCHK fails -> CALL), which
// made the NMS mask kernel spend half of its instructions there (ncu: 85 instructions per
// pair).  The shortcut returns what IEEE division would, up to the sign of zero, which no
// caller observes (the quotient is only compared against a threshold): 0/0 and 0/NaN = NaN.
__device__ __forceinline__ float iou_quotient(float inter, float den) {
  if (inter == 0.0f) return (den == 0.0f || den != den) ? __int_as_float(0x7fc00000) : 0.0f;
  return __fdiv_rn(inter, den);
}

// Exactly `iou_quotient(inter, den) >= thr` (thr > 0 finite), without the division in all but a
// sliver of cases.  Round-to-nearest is monotone and thr is a float, so RN(inter/den) >= thr iff
// inter/den >= t* for a real t* in [thr*(1 - 2^-24), thr]; with p = RN(thr*den) (relative error
// 2^-24) everything above p*(1 + 2^-19) is certainly >= t* * den and everything below
// p*(1 - 2^-19) certainly is not.  Only pairs inside that band -- or with a non-positive, tiny or
// non-finite denominator, where the error bounds do not hold -- pay for the IEEE division (which
// cost ~30 of the ~45 instructions of an IoU test: ncu on the NMS mask and proposals kernels).
__device__ __forceinline__ bool iou_ge_exact(float inter, float den, float thr) {
  const float p = __fmul_rn(thr, den);
  if (den > 0.0f && p > 1e-30f) {
    if (inter > __fmul_rn(p, 1.0000019073486328f)) return true;    // 1 + 2^-19
    if (inter < __fmul_rn(p, 0.9999980926513672f)) return false;   // 1 - 2^-19 (p = +inf lands here)
  }
  return iou_quotient(inter, den) >= thr;
}

// RoI geometry in feature-map cells, computed exactly as the reference does
// (roi_pooling_op.cc:153-165): C round() = half away from zero, float products.
struct RoiCells {
  int batch;
  int start_w, start_h, end_w, end_h;
  float bin_h, bin_w;
};

__device__ __forceinline__ RoiCells roi_cells(const float* __restrict__ roi, float spatial_scale,
                                              int pooled_h, int pooled_w) {
  RoiCells g;
  g.batch = (int)roi[0];
  g.start_w = (int)roundf(__fmul_rn(roi[1], spatial_scale));
  g.start_h = (int)roundf(__fmul_rn(roi[2], spatial_scale));
  g.end_w = (int)roundf(__fmul_rn(roi[3], spatial_scale));
  g.end_h = (int)roundf(__fmul_rn(roi[4], spatial_scale));
  int roi_w = max(g.end_w - g.start_w + 1, 1);  // malformed RoIs become 1x1 (:160-161)
  int roi_h = max(g.end_h - g.start_h + 1, 1);
  g.bin_h = __fdiv_rn((float)roi_h, (float)pooled_h);
  g.bin_w = __fdiv_rn((float)roi_w, (float)pooled_w);
  return g;
}

// Bin edge before the RoI offset is added.  CPU_TRUNC: the reference casts the float
// product to int before floor/ceil (roi_pooling_op.cc:167-170) so both edges truncate;
// GPU_CEIL: floor / ceil of the float product (roi_pooling_op_gpu.cu.cc:51-58).
template <int BIN_MODE>
__device__ __forceinline__ int bin_lo(int p, float bin) {
  float v = __fmul_rn((float)p, bin);
  return BIN_MODE == WSSDL_BIN_CPU_TRUNC ? (int)v : (int)floorf(v);
}
template <int BIN_MODE>
__device__ __forceinline__ int bin_hi(int p, float bin) {
  float v = __fmul_rn((float)(p + 1), bin);
  return BIN_MODE == WSSDL_BIN_CPU_TRUNC ? (int)v : (int)ceilf(v);
}
