// Box regression arithmetic shared by bbox.cu and detect.cu.
#pragma once
#include "common.cuh"

namespace {

// exp in fp64, rounded once to fp32: correctly rounded expf (np.exp on fp32 is within 1 ulp)
__device__ __forceinline__ float exp_cr(float x) { return (float)exp((double)x); }

// bbox_transform_inv for one (box, delta) group (fast_rcnn/bbox_transform.py:36-59), one
// fp32 rounding per operation like numpy.
__device__ __forceinline__ float4 decode_box(float x1, float y1, float x2, float y2, float dx,
                                             float dy, float dw, float dh) {
  const float w = __fadd_rn(__fsub_rn(x2, x1), 1.0f);                 // :36
  const float h = __fadd_rn(__fsub_rn(y2, y1), 1.0f);
  const float cx = __fadd_rn(x1, __fmul_rn(0.5f, w));                 // :38
  const float cy = __fadd_rn(y1, __fmul_rn(0.5f, h));
  const float pcx = __fadd_rn(__fmul_rn(dx, w), cx);                  // :46
  const float pcy = __fadd_rn(__fmul_rn(dy, h), cy);
  const float pw = __fmul_rn(exp_cr(dw), w);                          // :48
  const float ph = __fmul_rn(exp_cr(dh), h);
  return make_float4(__fsub_rn(pcx, __fmul_rn(0.5f, pw)), __fsub_rn(pcy, __fmul_rn(0.5f, ph)),
                     __fadd_rn(pcx, __fmul_rn(0.5f, pw)), __fadd_rn(pcy, __fmul_rn(0.5f, ph)));  // :53-59
}

}  // namespace
