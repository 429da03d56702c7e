/*
 * oracle/hotpath_ref.c -- CPU restatement of the reference's hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: it is compiled by
 * oracle/Makefile into oracle/liboracle.so and may be loaded only by tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline legs.  The product
 * (wssdl_bus_b200/) never links, loads or calls it.
 *
 * Parity status
 *   roi_pool_*      : PINNED.  bin_mode CPU_TRUNC is checked bit for bit (forward, backward,
 *                     adversarial / malformed RoIs, arbitrary argmax tensors) against the
 *                     reference's own RoiPool / RoiPoolGrad CPU kernels: roi_pooling_op.cc
 *                     compiled UNMODIFIED against oracle/tf_stub (a stand-in for the TF op
 *                     framework that holds no arithmetic) into oracle/_ref/ref_roi_pool.so.
 *                     bin_mode GPU_CEIL is checked bit for bit against the reference's CUDA
 *                     kernel source roi_pooling_op_gpu.cu.cc built for the host
 *                     (oracle/_ref/ref_roi_pool_cudatwin.so).  Recipes: oracle/build_ref.py;
 *                     tests: tests/test_oracle.py.  (The reference itself ships no golden
 *                     vectors for the op: roi_pooling_op_test.py asserts nothing.)
 *   nms_ref, bbox_overlaps_ref, bbox_overlaps_ui_ref
 *                   : pinned -- checked bit-for-bit against the reference's own
 *                     Cython modules compiled into oracle/_ref (tests/test_oracle.py)
 *                     and against fixtures in tests/golden generated from them.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp (no -march=native, no -ffast-math): every
 * float expression below is evaluated with one IEEE rounding per operation, exactly
 * as x86-64 SSE2 code generated from the reference would.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define IMIN(a, b) ((a) < (b) ? (a) : (b))
#define IMAX(a, b) ((a) > (b) ? (a) : (b))

/* ------------------------------------------------------------------------- */
/* RoiPool forward.                                                          */
/*   bin_mode 0 (CPU_TRUNC): follows roi_pooling_op.cc:141-195 (the CPU op,   */
/*     parity target): the int cast happens BEFORE floor/ceil (:167-170).    */
/*   bin_mode 1 (GPU_CEIL) : follows roi_pooling_op_gpu.cu.cc:25-84 (the CUDA */
/*     twin; differential tests only).                                        */
/*   bottom [B,H,W,C] f32, rois [R,5] f32 -> top [R,PH,PW,C] f32, argmax i32  */
/*   argmax is the per-image NHWC flat index (h*W+w)*C+c (cc:186).           */
/* ------------------------------------------------------------------------- */
void roi_pool_fwd_ref(const float* bottom_data_flat, const float* bottom_rois_flat,
                      int batch_size, int data_height, int data_width,
                      int num_channels, int num_rois, int pooled_height,
                      int pooled_width, float spatial_scale, int bin_mode,
                      float* output, int* argmax, int num_threads)
{
  (void)batch_size;
  const int64_t total =
      (int64_t)num_rois * pooled_height * pooled_width * num_channels;
  if (num_threads < 1) num_threads = 1;
#pragma omp parallel for schedule(static) num_threads(num_threads)
  for (int64_t b = 0; b < total; ++b) {
    /* (n, ph, pw, c) is an element in the pooled output  (cc:143-150) */
    int n = (int)b;
    int c = n % num_channels;
    n /= num_channels;
    int pw = n % pooled_width;
    n /= pooled_width;
    int ph = n % pooled_height;
    n /= pooled_height;

    const float* bottom_rois = bottom_rois_flat + n * 5;
    int roi_batch_ind = (int)bottom_rois[0];                       /* cc:153 */
    int roi_start_w = (int)round(bottom_rois[1] * spatial_scale);  /* cc:154 */
    int roi_start_h = (int)round(bottom_rois[2] * spatial_scale);
    int roi_end_w = (int)round(bottom_rois[3] * spatial_scale);
    int roi_end_h = (int)round(bottom_rois[4] * spatial_scale);

    /* Force malformed ROIs to be 1x1  (cc:160-161) */
    int roi_width = IMAX(roi_end_w - roi_start_w + 1, 1);
    int roi_height = IMAX(roi_end_h - roi_start_h + 1, 1);
    const float bin_size_h = (float)roi_height / (float)pooled_height;
    const float bin_size_w = (float)roi_width / (float)pooled_width;

    int hstart, wstart, hend, wend;
    if (bin_mode == 0) {
      /* cc:167-170: floor(static_cast<int>(ph * bin_size_h)) etc. */
      hstart = (int)floor((double)(int)((float)ph * bin_size_h));
      wstart = (int)floor((double)(int)((float)pw * bin_size_w));
      hend = (int)ceil((double)(int)((float)(ph + 1) * bin_size_h));
      wend = (int)ceil((double)(int)((float)(pw + 1) * bin_size_w));
    } else {
      /* cu.cc:51-58: static_cast<int>(floor(static_cast<Dtype>(ph) * bin_size_h)) */
      hstart = (int)floorf((float)ph * bin_size_h);
      wstart = (int)floorf((float)pw * bin_size_w);
      hend = (int)ceilf((float)(ph + 1) * bin_size_h);
      wend = (int)ceilf((float)(pw + 1) * bin_size_w);
    }

    /* Add roi offsets and clip to input boundaries  (cc:173-176) */
    hstart = IMIN(IMAX(hstart + roi_start_h, 0), data_height);
    hend = IMIN(IMAX(hend + roi_start_h, 0), data_height);
    wstart = IMIN(IMAX(wstart + roi_start_w, 0), data_width);
    wend = IMIN(IMAX(wend + roi_start_w, 0), data_width);
    int is_empty = (hend <= hstart) || (wend <= wstart);

    float maxval = is_empty ? 0 : -FLT_MAX;                        /* cc:180 */
    int maxidx = -1;
    const float* bottom_data =
        bottom_data_flat + roi_batch_ind * num_channels * data_height * data_width;
    for (int h = hstart; h < hend; ++h) {
      for (int w = wstart; w < wend; ++w) {
        int bottom_index = (h * data_width + w) * num_channels + c;
        if (bottom_data[bottom_index] > maxval) {                  /* strict > */
          maxval = bottom_data[bottom_index];
          maxidx = bottom_index;
        }
      }
    }
    output[b] = maxval;
    if (argmax) argmax[b] = maxidx;
  }
}

/* ------------------------------------------------------------------------- */
/* RoiPoolGrad, literal: follows roi_pooling_op.cc:387-457 loop for loop.     */
/* O(B*H*W*C*R): use on small cases; roi_pool_bwd_ref_fast below is the same  */
/* arithmetic in the same per-element order for large ones.                   */
/* ------------------------------------------------------------------------- */
void roi_pool_bwd_ref(const float* out_backprop_flat, const int* argmax_data_flat,
                      const float* bottom_rois_flat, int batch_size,
                      int data_height, int data_width, int num_channels,
                      int num_rois, int pooled_height, int pooled_width,
                      float spatial_scale, float* output, int num_threads)
{
  const int64_t total =
      (int64_t)batch_size * data_height * data_width * num_channels;
  if (num_threads < 1) num_threads = 1;
#pragma omp parallel for schedule(static) num_threads(num_threads)
  for (int64_t b = 0; b < total; ++b) {
    /* (n, h, w, c) coords in bottom data  (cc:389-396) */
    int n = (int)b;
    int c = n % num_channels;
    n /= num_channels;
    int w = n % data_width;
    n /= data_width;
    int h = n % data_height;
    n /= data_height;

    float gradient = 0.0;
    for (int roi_n = 0; roi_n < num_rois; ++roi_n) {
      const float* offset_bottom_rois = bottom_rois_flat + roi_n * 5;
      int roi_batch_ind = (int)offset_bottom_rois[0];
      if (n != roi_batch_ind) continue;                            /* cc:405 */

      int roi_start_w = (int)round(offset_bottom_rois[1] * spatial_scale);
      int roi_start_h = (int)round(offset_bottom_rois[2] * spatial_scale);
      int roi_end_w = (int)round(offset_bottom_rois[3] * spatial_scale);
      int roi_end_h = (int)round(offset_bottom_rois[4] * spatial_scale);

      /* cc:415-419: test on the un-clipped rounded RoI */
      const int in_roi = (w >= roi_start_w && w <= roi_end_w &&
                          h >= roi_start_h && h <= roi_end_h);
      if (!in_roi) continue;

      int offset = roi_n * pooled_height * pooled_width * num_channels;
      const float* offset_top_diff = out_backprop_flat + offset;
      const int* offset_argmax_data = argmax_data_flat + offset;

      int roi_width = IMAX(roi_end_w - roi_start_w + 1, 1);
      int roi_height = IMAX(roi_end_h - roi_start_h + 1, 1);
      const float bin_size_h = (float)roi_height / (float)pooled_height;
      const float bin_size_w = (float)roi_width / (float)pooled_width;

      /* cc:437-440: int / float -> float, then floor / ceil */
      int phstart = (int)floorf((float)(int)(h - roi_start_h) / bin_size_h);
      int phend = (int)ceilf((float)(int)(h - roi_start_h + 1) / bin_size_h);
      int pwstart = (int)floorf((float)(int)(w - roi_start_w) / bin_size_w);
      int pwend = (int)ceilf((float)(int)(w - roi_start_w + 1) / bin_size_w);

      phstart = IMIN(IMAX(phstart, 0), pooled_height);
      phend = IMIN(IMAX(phend, 0), pooled_height);
      pwstart = IMIN(IMAX(pwstart, 0), pooled_width);
      pwend = IMIN(IMAX(pwend, 0), pooled_width);

      for (int ph = phstart; ph < phend; ++ph) {
        for (int pw = pwstart; pw < pwend; ++pw) {
          if (offset_argmax_data[(ph * pooled_width + pw) * num_channels + c] ==
              (h * data_width + w) * num_channels + c) {
            gradient += offset_top_diff[(ph * pooled_width + pw) * num_channels + c];
          }
        }
      }
    }
    output[b] = gradient;
  }
}

/* Same arithmetic, loops interchanged (roi outermost per cell, channel innermost)
 * so the cost is O(sum_roi area*bins*C) instead of O(B*H*W*C*R).  For every output
 * element the sequence of float additions is still (roi asc, ph asc, pw asc), hence
 * bit-identical to roi_pool_bwd_ref (asserted in tests/test_oracle.py). */
void roi_pool_bwd_ref_fast(const float* out_backprop_flat, const int* argmax_data_flat,
                           const float* bottom_rois_flat, int batch_size,
                           int data_height, int data_width, int num_channels,
                           int num_rois, int pooled_height, int pooled_width,
                           float spatial_scale, float* output, int num_threads)
{
  const int64_t cells = (int64_t)batch_size * data_height * data_width;
  if (num_threads < 1) num_threads = 1;
  int* rb = (int*)malloc(sizeof(int) * 5 * (size_t)(num_rois > 0 ? num_rois : 1));
  for (int r = 0; r < num_rois; ++r) {
    const float* q = bottom_rois_flat + r * 5;
    rb[r * 5 + 0] = (int)q[0];
    rb[r * 5 + 1] = (int)round(q[1] * spatial_scale);
    rb[r * 5 + 2] = (int)round(q[2] * spatial_scale);
    rb[r * 5 + 3] = (int)round(q[3] * spatial_scale);
    rb[r * 5 + 4] = (int)round(q[4] * spatial_scale);
  }
#pragma omp parallel for schedule(dynamic, 16) num_threads(num_threads)
  for (int64_t cell = 0; cell < cells; ++cell) {
    int n = (int)cell;
    int w = n % data_width;
    n /= data_width;
    int h = n % data_height;
    n /= data_height;
    float* grad = output + cell * num_channels;
    for (int c = 0; c < num_channels; ++c) grad[c] = 0.0f;
    const int base_index = (h * data_width + w) * num_channels;
    for (int roi_n = 0; roi_n < num_rois; ++roi_n) {
      const int* q = rb + roi_n * 5;
      if (n != q[0]) continue;
      int roi_start_w = q[1], roi_start_h = q[2], roi_end_w = q[3], roi_end_h = q[4];
      if (!(w >= roi_start_w && w <= roi_end_w && h >= roi_start_h && h <= roi_end_h))
        continue;
      int offset = roi_n * pooled_height * pooled_width * num_channels;
      const float* offset_top_diff = out_backprop_flat + offset;
      const int* offset_argmax_data = argmax_data_flat + offset;
      int roi_width = IMAX(roi_end_w - roi_start_w + 1, 1);
      int roi_height = IMAX(roi_end_h - roi_start_h + 1, 1);
      const float bin_size_h = (float)roi_height / (float)pooled_height;
      const float bin_size_w = (float)roi_width / (float)pooled_width;
      int phstart = (int)floorf((float)(int)(h - roi_start_h) / bin_size_h);
      int phend = (int)ceilf((float)(int)(h - roi_start_h + 1) / bin_size_h);
      int pwstart = (int)floorf((float)(int)(w - roi_start_w) / bin_size_w);
      int pwend = (int)ceilf((float)(int)(w - roi_start_w + 1) / bin_size_w);
      phstart = IMIN(IMAX(phstart, 0), pooled_height);
      phend = IMIN(IMAX(phend, 0), pooled_height);
      pwstart = IMIN(IMAX(pwstart, 0), pooled_width);
      pwend = IMIN(IMAX(pwend, 0), pooled_width);
      for (int ph = phstart; ph < phend; ++ph) {
        for (int pw = pwstart; pw < pwend; ++pw) {
          const int* am = offset_argmax_data + (ph * pooled_width + pw) * num_channels;
          const float* td = offset_top_diff + (ph * pooled_width + pw) * num_channels;
          for (int c = 0; c < num_channels; ++c) {
            if (am[c] == base_index + c) grad[c] += td[c];
          }
        }
      }
    }
  }
  free(rb);
}

/* ------------------------------------------------------------------------- */
/* cpu_nms restatement: follows nms/cpu_nms.pyx:17-68 (== utils/nms.pyx:17-68) */
/* with the generated C's exact arithmetic (cpu_nms.c:2442-2495):             */
/*   areas (numpy, f32): ((x2-x1)+1)*((y2-y1)+1), one rounding per op  (:24)  */
/*   w = max(0, (float)((double)(xx2-xx1)+1.0))                       (:61)   */
/*   ovr = inter / ((iarea + areas[j]) - inter)   all f32             (:64)   */
/*   suppress iff (double)ovr >= thresh (thresh is a Python float)    (:65)   */
/* `order` (argsort()[::-1] of the scores, :25) is supplied by the caller so   */
/* that numpy's tie order is whatever numpy does.                             */
/* variant 1 = nms_new (utils/nms.pyx:70-123): also suppress when             */
/*   inter/iarea > 0.95 or inter/areas[j] > 0.95 (f32 division, f64 compare). */
/* Returns the number kept; -1 when a visited pair has a zero union (the      */
/* reference raises ZeroDivisionError there, cpu_nms.c:2480-2483).            */
/* ------------------------------------------------------------------------- */
static inline float fmax_ref(float a, float b) { return a >= b ? a : b; } /* pyx:11 */
static inline float fmin_ref(float a, float b) { return a <= b ? a : b; } /* pyx:14 */

int nms_ref(const float* dets, int ndets, int stride, const int64_t* order,
            double thresh, int variant, int64_t* keep)
{
  float* areas = (float*)malloc(sizeof(float) * (size_t)(ndets > 0 ? ndets : 1));
  unsigned char* suppressed = (unsigned char*)calloc((size_t)(ndets > 0 ? ndets : 1), 1);
  for (int k = 0; k < ndets; ++k) {
    const float* d = dets + (size_t)k * stride;
    float a = d[2] - d[0];
    a = a + 1.0f;
    float b = d[3] - d[1];
    b = b + 1.0f;
    areas[k] = a * b;
  }
  int nkeep = 0;
  for (int _i = 0; _i < ndets; ++_i) {
    int i = (int)order[_i];
    if (suppressed[i] == 1) continue;
    keep[nkeep++] = i;
    const float* di = dets + (size_t)i * stride;
    float ix1 = di[0], iy1 = di[1], ix2 = di[2], iy2 = di[3];
    float iarea = areas[i];
    for (int _j = _i + 1; _j < ndets; ++_j) {
      int j = (int)order[_j];
      if (suppressed[j] == 1) continue;
      const float* dj = dets + (size_t)j * stride;
      float xx1 = fmax_ref(ix1, dj[0]);
      float yy1 = fmax_ref(iy1, dj[1]);
      float xx2 = fmin_ref(ix2, dj[2]);
      float yy2 = fmin_ref(iy2, dj[3]);
      float w = fmax_ref(0.0f, (float)((double)(float)(xx2 - xx1) + 1.0));
      float h = fmax_ref(0.0f, (float)((double)(float)(yy2 - yy1) + 1.0));
      float inter = w * h;
      float s = iarea + areas[j];
      float den = s - inter;
      if (den == 0) { free(areas); free(suppressed); return -1; }
      float ovr = inter / den;
      int sup = ((double)ovr >= thresh);
      if (variant == 1) {
        /* pyx:117-120; iarea == 0 would raise in the reference as well */
        if (iarea == 0 || areas[j] == 0) { free(areas); free(suppressed); return -1; }
        float ovr1 = inter / iarea;
        float ovr2 = inter / areas[j];
        sup = sup || ((double)ovr1 > 0.95) || ((double)ovr2 > 0.95);
      }
      if (sup) suppressed[j] = 1;
    }
  }
  free(areas);
  free(suppressed);
  return nkeep;
}

/* ------------------------------------------------------------------------- */
/* bbox_overlaps (utils/bbox.pyx:15-55) and bbox_overlaps_ui                  */
/* (utils/bbox_ui.pyx:12-46), fp64.  Association follows the generated C       */
/* (bbox.c:2068): ua = ((bw*bh) + box_area) - (iw*ih); o = (iw*ih)/ua.        */
/* ------------------------------------------------------------------------- */
static inline double dmin_ref(double a, double b) { return a < b ? a : b; }
static inline double dmax_ref(double a, double b) { return a > b ? a : b; }

void bbox_overlaps_ref(const double* boxes, int N, const double* query_boxes, int K,
                       double* overlaps)
{
  memset(overlaps, 0, sizeof(double) * (size_t)N * (size_t)K);
  for (int k = 0; k < K; ++k) {
    const double* q = query_boxes + 4 * (size_t)k;
    double box_area = (q[2] - q[0] + 1) * (q[3] - q[1] + 1);
    for (int n = 0; n < N; ++n) {
      const double* b = boxes + 4 * (size_t)n;
      double iw = dmin_ref(b[2], q[2]) - dmax_ref(b[0], q[0]) + 1;
      if (iw > 0) {
        double ih = dmin_ref(b[3], q[3]) - dmax_ref(b[1], q[1]) + 1;
        if (ih > 0) {
          double ua = (b[2] - b[0] + 1) * (b[3] - b[1] + 1) + box_area - iw * ih;
          overlaps[(size_t)n * K + k] = iw * ih / ua;
        }
      }
    }
  }
}

void bbox_overlaps_ui_ref(const double* boxes, int N, const double* query_boxes, int K,
                          double* overlaps)
{
  memset(overlaps, 0, sizeof(double) * (size_t)N * (size_t)K);
  for (int n = 0; n < N; ++n) {
    const double* b = boxes + 4 * (size_t)n;
    double box_area = (b[2] - b[0] + 1) * (b[3] - b[1] + 1);
    for (int k = 0; k < K; ++k) {
      const double* q = query_boxes + 4 * (size_t)k;
      double iw = dmin_ref(b[2], q[2]) - dmax_ref(b[0], q[0]) + 1;
      if (iw > 0) {
        double ih = dmin_ref(b[3], q[3]) - dmax_ref(b[1], q[1]) + 1;
        if (ih > 0) overlaps[(size_t)n * K + k] = iw * ih / box_area;
      }
    }
  }
}
