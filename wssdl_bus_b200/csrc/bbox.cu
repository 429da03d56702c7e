// IoU matrices and box transforms for sm_100a.
//
//   wssdl_bbox_overlaps_f64/_f32 : bbox_overlaps (utils/bbox.pyx:15-55) and
//                                  bbox_overlaps_ui (utils/bbox_ui.pyx:12-46)
//   wssdl_bbox_transform_inv     : fast_rcnn/bbox_transform.py:30-61
//   wssdl_clip_boxes             : fast_rcnn/bbox_transform.py:63-77
//   wssdl_bbox_transform         : fast_rcnn/bbox_transform.py:10-28
//
// All of these are elementwise / outer-product kernels bound by the output stream
// (N*K*8 B for the fp64 IoU matrix).  One thread per output element, k fastest, so stores
// are fully coalesced; the query boxes (K*32 B) are staged in shared memory per CTA when
// they fit, the row box is a broadcast load.  Every arithmetic operation is written with
// the round-to-nearest intrinsics so the compiler cannot contract mul+add into FMA: the
// reference's expression tree (bbox.c:2068: ((bw*bh)+qarea)-(iw*ih), then (iw*ih)/ua) is
// evaluated with exactly one rounding per operation, which makes the fp64 path bit-exact.
#include "common.cuh"

namespace {

template <typename T> struct Ops;
template <> struct Ops<double> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};
template <> struct Ops<float> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};

template <typename T> struct Box { T x1, y1, x2, y2; };

// Cython lowers min(a,b)/max(a,b) on C doubles to (b < a ? b : a) / (b > a ? b : a)
template <typename T> __device__ __forceinline__ T cmin(T a, T b) { return b < a ? b : a; }
template <typename T> __device__ __forceinline__ T cmax(T a, T b) { return b > a ? b : a; }

template <typename T, int KIND>
__device__ __forceinline__ T overlap(const Box<T>& b, const Box<T>& q) {
  using O = Ops<T>;
  const T one = (T)1;
  const T iw = O::add(O::sub(cmin(b.x2, q.x2), cmax(b.x1, q.x1)), one);
  if (!(iw > (T)0)) return (T)0;
  const T ih = O::add(O::sub(cmin(b.y2, q.y2), cmax(b.y1, q.y1)), one);
  if (!(ih > (T)0)) return (T)0;
  const T barea = O::mul(O::add(O::sub(b.x2, b.x1), one), O::add(O::sub(b.y2, b.y1), one));
  const T inter = O::mul(iw, ih);
  if (KIND == WSSDL_IOU_UI) return O::div(inter, barea);                  // bbox_ui.pyx:45
  const T qarea = O::mul(O::add(O::sub(q.x2, q.x1), one), O::add(O::sub(q.y2, q.y1), one));
  const T ua = O::sub(O::add(barea, qarea), inter);                       // bbox.pyx:49-53
  return O::div(inter, ua);
}

constexpr int OV_THREADS = 256;
constexpr int OV_PER_THREAD = 4;
constexpr int OV_SMEM_K = 1024;   // query boxes staged in shared memory when K <= this

template <typename T, int KIND>
__global__ void __launch_bounds__(OV_THREADS)
bbox_overlaps_kernel(const T* __restrict__ boxes, long long N, const T* __restrict__ query,
                     int K, T* __restrict__ out) {
  __shared__ Box<T> s_q[OV_SMEM_K];
  const bool staged = K <= OV_SMEM_K;
  if (staged) {
    for (int k = threadIdx.x; k < K; k += OV_THREADS) {
      Box<T> q;
      q.x1 = query[4 * k]; q.y1 = query[4 * k + 1]; q.x2 = query[4 * k + 2]; q.y2 = query[4 * k + 3];
      s_q[k] = q;
    }
    __syncthreads();
  }
  const long long total = N * K;
  const long long base = ((long long)blockIdx.x * OV_PER_THREAD) * OV_THREADS + threadIdx.x;
#pragma unroll
  for (int u = 0; u < OV_PER_THREAD; ++u) {
    const long long idx = base + (long long)u * OV_THREADS;
    if (idx >= total) break;
    const long long n = idx / K;
    const int k = (int)(idx - n * K);
    Box<T> b;
    b.x1 = boxes[4 * n]; b.y1 = boxes[4 * n + 1]; b.x2 = boxes[4 * n + 2]; b.y2 = boxes[4 * n + 3];
    Box<T> q;
    if (staged) q = s_q[k];
    else { q.x1 = query[4 * k]; q.y1 = query[4 * k + 1]; q.x2 = query[4 * k + 2]; q.y2 = query[4 * k + 3]; }
    out[idx] = overlap<T, KIND>(b, q);
  }
}

template <typename T>
int launch_overlaps(const T* boxes, int N, const T* query, int K, int kind, T* out,
                    cudaStream_t s) {
  if (N < 0 || K < 0 || (kind != WSSDL_IOU && kind != WSSDL_IOU_UI)) return WSSDL_EINVAL;
  const long long total = (long long)N * K;
  if (total == 0) return WSSDL_OK;
  if (!boxes || !query || !out) return WSSDL_EINVAL;
  const long long per_cta = (long long)OV_THREADS * OV_PER_THREAD;
  const long long ctas = (total + per_cta - 1) / per_cta;
  if (ctas > 0x7fffffffll) return WSSDL_ELIMIT;
  if (kind == WSSDL_IOU)
    bbox_overlaps_kernel<T, WSSDL_IOU><<<(unsigned)ctas, OV_THREADS, 0, s>>>(boxes, N, query, K, out);
  else
    bbox_overlaps_kernel<T, WSSDL_IOU_UI><<<(unsigned)ctas, OV_THREADS, 0, s>>>(boxes, N, query, K, out);
  WSSDL_CHECK_LAUNCH();
  return WSSDL_OK;
}

// exp in fp64, rounded once to fp32: correctly rounded expf (np.exp on fp32 is within 1 ulp)
__device__ __forceinline__ float exp_cr(float x) { return (float)exp((double)x); }

// bbox_transform_inv: one thread per (row, class) group of 4 deltas
__global__ void bbox_transform_inv_kernel(const float* __restrict__ boxes,
                                          const float* __restrict__ deltas, long long N, int k,
                                          float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * k) return;
  const long long n = idx / k;
  const float x1 = boxes[4 * n], y1 = boxes[4 * n + 1], x2 = boxes[4 * n + 2], y2 = boxes[4 * n + 3];
  const float w = __fadd_rn(__fsub_rn(x2, x1), 1.0f);                 // :36
  const float h = __fadd_rn(__fsub_rn(y2, y1), 1.0f);
  const float cx = __fadd_rn(x1, __fmul_rn(0.5f, w));                 // :38
  const float cy = __fadd_rn(y1, __fmul_rn(0.5f, h));
  const float dx = deltas[4 * idx], dy = deltas[4 * idx + 1];
  const float dw = deltas[4 * idx + 2], dh = deltas[4 * idx + 3];
  const float pcx = __fadd_rn(__fmul_rn(dx, w), cx);                  // :46
  const float pcy = __fadd_rn(__fmul_rn(dy, h), cy);
  const float pw = __fmul_rn(exp_cr(dw), w);                          // :48
  const float ph = __fmul_rn(exp_cr(dh), h);
  out[4 * idx] = __fsub_rn(pcx, __fmul_rn(0.5f, pw));                 // :53-59
  out[4 * idx + 1] = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
  out[4 * idx + 2] = __fadd_rn(pcx, __fmul_rn(0.5f, pw));
  out[4 * idx + 3] = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
}

__device__ __forceinline__ float clipf(float v, float hi) {
  // np.maximum(np.minimum(v, hi), 0)  (NaN propagates in numpy; fminf/fmaxf would drop it)
  float t = (v != v) ? v : (v < hi ? v : hi);
  return (t != t) ? t : (t > 0.f ? t : 0.f);
}

__global__ void clip_boxes_kernel(float* __restrict__ boxes, long long groups, float xmax,
                                  float ymax) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= groups) return;
  float* b = boxes + 4 * idx;
  b[0] = clipf(b[0], xmax);
  b[1] = clipf(b[1], ymax);
  b[2] = clipf(b[2], xmax);
  b[3] = clipf(b[3], ymax);
}

// bbox_transform (regression targets), fp32; log evaluated in fp64 and rounded once
__global__ void bbox_transform_kernel(const float* __restrict__ ex, const float* __restrict__ gt,
                                      long long N, float* __restrict__ t) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float ew = __fadd_rn(__fsub_rn(ex[4 * n + 2], ex[4 * n]), 1.0f);
  const float eh = __fadd_rn(__fsub_rn(ex[4 * n + 3], ex[4 * n + 1]), 1.0f);
  const float ecx = __fadd_rn(ex[4 * n], __fmul_rn(0.5f, ew));
  const float ecy = __fadd_rn(ex[4 * n + 1], __fmul_rn(0.5f, eh));
  const float gw = __fadd_rn(__fsub_rn(gt[4 * n + 2], gt[4 * n]), 1.0f);
  const float gh = __fadd_rn(__fsub_rn(gt[4 * n + 3], gt[4 * n + 1]), 1.0f);
  const float gcx = __fadd_rn(gt[4 * n], __fmul_rn(0.5f, gw));
  const float gcy = __fadd_rn(gt[4 * n + 1], __fmul_rn(0.5f, gh));
  t[4 * n] = __fdiv_rn(__fsub_rn(gcx, ecx), ew);
  t[4 * n + 1] = __fdiv_rn(__fsub_rn(gcy, ecy), eh);
  t[4 * n + 2] = (float)log((double)__fdiv_rn(gw, ew));
  t[4 * n + 3] = (float)log((double)__fdiv_rn(gh, eh));
}

}  // namespace

extern "C" int wssdl_bbox_overlaps_f64(const double* boxes, int N, const double* query, int K,
                                       int kind, double* out, wssdl_stream_t stream) {
  return launch_overlaps<double>(boxes, N, query, K, kind, out, to_cuda(stream));
}

extern "C" int wssdl_bbox_overlaps_f32(const float* boxes, int N, const float* query, int K,
                                       int kind, float* out, wssdl_stream_t stream) {
  return launch_overlaps<float>(boxes, N, query, K, kind, out, to_cuda(stream));
}

extern "C" int wssdl_bbox_transform_inv(const float* boxes, const float* deltas, int N, int k,
                                        float* out, wssdl_stream_t stream) {
  if (N < 0 || k < 0) return WSSDL_EINVAL;
  const long long groups = (long long)N * k;
  if (groups == 0) return WSSDL_OK;
  if (!boxes || !deltas || !out) return WSSDL_EINVAL;
  bbox_transform_inv_kernel<<<ceil_div(groups, 256), 256, 0, to_cuda(stream)>>>(boxes, deltas, N,
                                                                               k, out);
  WSSDL_CHECK_LAUNCH();
  return WSSDL_OK;
}

extern "C" int wssdl_clip_boxes(float* boxes, int N, int k, float im_h, float im_w,
                                wssdl_stream_t stream) {
  if (N < 0 || k < 0) return WSSDL_EINVAL;
  const long long groups = (long long)N * k;
  if (groups == 0) return WSSDL_OK;
  if (!boxes) return WSSDL_EINVAL;
  // im_shape[1] - 1 / im_shape[0] - 1 in fp32 (bbox_transform.py:69-75)
  clip_boxes_kernel<<<ceil_div(groups, 256), 256, 0, to_cuda(stream)>>>(boxes, groups,
                                                                       im_w - 1.0f, im_h - 1.0f);
  WSSDL_CHECK_LAUNCH();
  return WSSDL_OK;
}

extern "C" int wssdl_bbox_transform(const float* ex_rois, const float* gt_rois, int N,
                                    float* targets, wssdl_stream_t stream) {
  if (N < 0) return WSSDL_EINVAL;
  if (N == 0) return WSSDL_OK;
  if (!ex_rois || !gt_rois || !targets) return WSSDL_EINVAL;
  bbox_transform_kernel<<<ceil_div(N, 256), 256, 0, to_cuda(stream)>>>(ex_rois, gt_rois, N, targets);
  WSSDL_CHECK_LAUNCH();
  return WSSDL_OK;
}
