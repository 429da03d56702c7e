// IoU matrices and box transforms for sm_100a.
//
//   wssdl_bbox_overlaps_f64/_f32 : bbox_overlaps (utils/bbox.pyx:15-55) and
//                                  bbox_overlaps_ui (utils/bbox_ui.pyx:12-46)
//   wssdl_bbox_transform_inv     : fast_rcnn/bbox_transform.py:30-61
//   wssdl_clip_boxes             : fast_rcnn/bbox_transform.py:63-77
//   wssdl_bbox_transform         : fast_rcnn/bbox_transform.py:10-28
//
// All of these are elementwise / outer-product kernels bound by the output stream
// (N*K*8 B for the fp64 IoU matrix); see the tiling note above bbox_overlaps_kernel.
// Every arithmetic operation is written with
// the round-to-nearest intrinsics so the compiler cannot contract mul+add into FMA: the
// reference's expression tree (bbox.c:2068: ((bw*bh)+qarea)-(iw*ih), then (iw*ih)/ua) is
// evaluated with exactly one rounding per operation, which makes the fp64 path bit-exact.
#include "common.cuh"
#include "box_common.cuh"

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

// Cython lowers min(a,b)/max(a,b) on C doubles to (b < a ? b : a) / (b > a ? b : a)
template <typename T> __device__ __forceinline__ T cmin(T a, T b) { return b < a ? b : a; }
template <typename T> __device__ __forceinline__ T cmax(T a, T b) { return b > a ? b : a; }

// Output-stream bound for large matrices (N*K*sizeof(T) written once, the boxes are
// noise), so the kernel is organised around the store: a CTA owns a tile of OV_ROWS rows x
// (KX * VEC) columns, a thread owns VEC consecutive columns (one 16-byte store per row) and
// walks down the rows of the tile.  Its VEC query boxes and their areas live in registers
// for the whole tile; the tile's row boxes and areas are staged once in shared memory and
// read as broadcasts.  No integer division, no recomputed areas, one vector store per VEC
// results.  The block shape adapts to K (KX = pow2 >= K/VEC, capped at OV_THREADS) so the
// detector's skinny matrices (17100 x 20) still fill their warps.
constexpr int OV_THREADS = 256;
constexpr int OV_ROWS = 64;       // rows per CTA tile

template <typename T> struct BoxA { T x1, y1, x2, y2, area; };

template <typename T, int KIND>
__device__ __forceinline__ T overlap_a(const BoxA<T>& b, const BoxA<T>& q) {
  using O = Ops<T>;
  const T one = (T)1;
  const T iw = O::add(O::sub(cmin(b.x2, q.x2), cmax(b.x1, q.x1)), one);     // bbox.pyx:38-41
  const T ih = O::add(O::sub(cmin(b.y2, q.y2), cmax(b.y1, q.y1)), one);     // :43-46
  if (!(iw > (T)0) || !(ih > (T)0)) return (T)0;
  const T inter = O::mul(iw, ih);
  if (KIND == WSSDL_IOU_UI) return O::div(inter, b.area);                   // bbox_ui.pyx:45
  const T ua = O::sub(O::add(b.area, q.area), inter);                       // bbox.pyx:49-53
  return O::div(inter, ua);
}

template <typename T>
__device__ __forceinline__ BoxA<T> load_box(const T* __restrict__ p) {
  using O = Ops<T>;
  BoxA<T> b;
  b.x1 = p[0]; b.y1 = p[1]; b.x2 = p[2]; b.y2 = p[3];
  b.area = O::mul(O::add(O::sub(b.x2, b.x1), (T)1), O::add(O::sub(b.y2, b.y1), (T)1));
  return b;
}

template <typename T, int VEC> struct OutVec;
template <> struct OutVec<float, 4> { using V = float4; };
template <> struct OutVec<double, 2> { using V = double2; };
template <> struct OutVec<float, 1> { using V = float; };
template <> struct OutVec<double, 1> { using V = double; };

template <typename T, int KIND, int VEC>
__global__ void __launch_bounds__(OV_THREADS)
bbox_overlaps_kernel(const T* __restrict__ boxes, int N, const T* __restrict__ query, int K,
                     int kx_log2, T* __restrict__ out) {
  __shared__ BoxA<T> s_b[OV_ROWS];
  const int KX = 1 << kx_log2;                  // threads along k
  const int RY = OV_THREADS >> kx_log2;         // rows in flight per CTA
  const int tx = threadIdx.x & (KX - 1), ty = threadIdx.x >> kx_log2;
  const int n0 = blockIdx.y * OV_ROWS;
  const int rows = min(OV_ROWS, N - n0);
  for (int r = threadIdx.x; r < rows; r += OV_THREADS) s_b[r] = load_box(boxes + 4 * (size_t)(n0 + r));
  const int k0 = (blockIdx.x * KX + tx) * VEC;
  BoxA<T> q[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v)
    if (k0 + v < K) q[v] = load_box(query + 4 * (size_t)(k0 + v));
  __syncthreads();
  if (k0 >= K) return;
  T* orow = out + (size_t)(n0 + ty) * K + k0;
  const size_t ostep = (size_t)RY * K;
  for (int r = ty; r < rows; r += RY, orow += ostep) {
    const BoxA<T> b = s_b[r];
    T o[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) o[v] = overlap_a<T, KIND>(b, q[v]);
    if constexpr (VEC == 1) {
      __stcs(orow, o[0]);
    } else {
      using V = typename OutVec<T, VEC>::V;      // K % VEC == 0 on this path: aligned, in range
      V ov;
      if constexpr (VEC == 4) ov = make_float4(o[0], o[1], o[2], o[3]);
      else ov = make_double2(o[0], o[1]);
      __stcs(reinterpret_cast<V*>(orow), ov);
    }
  }
}

template <typename T>
int launch_overlaps(const T* boxes, int N, const T* query, int K, int kind, T* out,
                    cudaStream_t s) {
  if (N < 0 || K < 0 || (kind != WSSDL_IOU && kind != WSSDL_IOU_UI)) return WSSDL_EINVAL;
  const long long total = (long long)N * K;
  if (total == 0) return WSSDL_OK;
  if (!boxes || !query || !out) return WSSDL_EINVAL;
  constexpr int VECMAX = 16 / (int)sizeof(T);
  const bool vec = (K % VECMAX == 0) && aligned16(out);
  const int V = vec ? VECMAX : 1;
  const int kthreads = (K + V - 1) / V;
  int kx_log2 = 0;
  while ((1 << kx_log2) < kthreads && (1 << kx_log2) < OV_THREADS) ++kx_log2;
  const long long gx = (kthreads + (1 << kx_log2) - 1) >> kx_log2;
  const long long gy = ((long long)N + OV_ROWS - 1) / OV_ROWS;
  if (gx > 0x7fffffffll || gy > 65535) {
    // more than 4.19 M rows: fold the row tiles over grid.x of several launches
    return WSSDL_ELIMIT;
  }
  dim3 grid((unsigned)gx, (unsigned)gy);
#define LAUNCH_OV(KIND_, V_)                                                                  \
  bbox_overlaps_kernel<T, KIND_, V_><<<grid, OV_THREADS, 0, s>>>(boxes, N, query, K, kx_log2, out)
  if (kind == WSSDL_IOU) { if (vec) LAUNCH_OV(WSSDL_IOU, VECMAX); else LAUNCH_OV(WSSDL_IOU, 1); }
  else { if (vec) LAUNCH_OV(WSSDL_IOU_UI, VECMAX); else LAUNCH_OV(WSSDL_IOU_UI, 1); }
#undef LAUNCH_OV
  WSSDL_CHECK_LAUNCH();
  return WSSDL_OK;
}

// bbox_transform_inv: one thread per (row, class) group of 4 deltas
__global__ void bbox_transform_inv_kernel(const float* __restrict__ boxes,
                                          const float* __restrict__ deltas, long long N, int k,
                                          float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * k) return;
  const long long n = idx / k;
  const float x1 = boxes[4 * n], y1 = boxes[4 * n + 1], x2 = boxes[4 * n + 2], y2 = boxes[4 * n + 3];
  const float4 o = decode_box(x1, y1, x2, y2, deltas[4 * idx], deltas[4 * idx + 1],
                              deltas[4 * idx + 2], deltas[4 * idx + 3]);
  out[4 * idx] = o.x; out[4 * idx + 1] = o.y; out[4 * idx + 2] = o.z; out[4 * idx + 3] = o.w;
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
