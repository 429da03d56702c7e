// NMS predicate and sort-key helpers shared by nms.cu and detect.cu.
#pragma once
#include <math.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ unsigned orderable(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct Thresh {
  float ge;       // mode GE_F64: smallest float whose double value is >= thresh
  float gt;       // mode GT_F32: (float)thresh, test is '>'
  int use_gt;
  int contain;    // nms_new's extra containment tests
  float eff;      // either mode as ONE test: suppressed iff iou >= eff (x > gt  <=>  x >= next float up)
};

// max/min exactly as cpu_nms.pyx:11-15 (NaN-asymmetric like the reference)
__device__ __forceinline__ float rmax(float a, float b) { return a >= b ? a : b; }
__device__ __forceinline__ float rmin(float a, float b) { return a <= b ? a : b; }

// true iff box j (lower score) is suppressed by box i (higher score); *zero is set when
// the union is exactly zero (the reference raises ZeroDivisionError when it visits such a
// pair, cpu_nms.c:2480-2483).
__device__ __forceinline__ bool suppresses(const float4 bi, float ai, const float4 bj, float aj,
                                           const Thresh th, bool* zero) {
  const float xx1 = rmax(bi.x, bj.x), yy1 = rmax(bi.y, bj.y);
  const float xx2 = rmin(bi.z, bj.z), yy2 = rmin(bi.w, bj.w);
  const float w = rmax(0.0f, __fadd_rn(__fsub_rn(xx2, xx1), 1.0f));
  const float h = rmax(0.0f, __fadd_rn(__fsub_rn(yy2, yy1), 1.0f));
  const float inter = __fmul_rn(w, h);
  const float den = __fsub_rn(__fadd_rn(ai, aj), inter);
  if (den == 0.0f) *zero = true;
  // (the IEEE division only where the quotient is within 2^-19 of the threshold: common.cuh)
  bool s = iou_ge_exact(inter, den, th.eff);
  if (th.contain) {
    // nms.pyx:117-120: float division, compared against the double 0.95: x > 0.95 for a float x
    // is x > 0.949999988f (the largest float below 0.95), i.e. x >= 0.95000005f, the next float
    const float c95 = 0.95000004768371582031f;
    if (ai == 0.0f || aj == 0.0f) *zero = true;   // the reference raises there as well
    s = s || iou_ge_exact(inter, ai, c95) || iou_ge_exact(inter, aj, c95);
  }
  return s;
}

inline Thresh make_thresh(double thresh, int mode) {
  Thresh t;
  t.use_gt = (mode & 3) == WSSDL_NMS_GT_F32;
  t.contain = (mode & WSSDL_NMS_CONTAIN) ? 1 : 0;
  t.gt = (float)thresh;
  // smallest float f with (double)f >= thresh
  float f = (float)thresh;
  if ((double)f < thresh) f = nextafterf(f, INFINITY);
  t.ge = f;
  t.eff = t.use_gt ? nextafterf(t.gt, INFINITY) : t.ge;
  return t;
}


}  // namespace
