// Deterministic part of anchor_target_layer[_joint] for sm_100a
// (rpn_msr/anchor_target_layer_tf_bus.py:410-509): inside filter, fp64 IoU of every inside
// anchor against the foreground GT rows, uni-directional overlap against the explicit
// background rows, row max/argmax, per-GT column max, labels.  Nothing here is sampled;
// npr.choice (:512-527) stays on the host.
//
// Two passes over the (anchor, gt) pairs, anchors generated on the fly from (h, w, a):
//   pass 1: column maxima via 64-bit atomicMax on the IEEE bit pattern (IoU >= 0, so the
//           unsigned order of the bits is the numeric order);
//   pass 2: recomputes the same IoUs (bit-identical) and applies the label rules in the
//           reference's order.
// The IoU expression tree is the one of bbox.pyx:39-54 / bbox_ui.pyx:35-45, one rounding
// per operation, so labels are bit-exact against the fp64 Cython path.
#include "common.cuh"

namespace {

constexpr int MAX_ANCHORS = 32;
constexpr int MAX_GT = 64;

struct AtParams {
  const float* gt_boxes;   // [B,max_gt,5]
  const int* num_gt;       // [B]
  int max_gt;
  const float* im_info;
  int info_stride;
  int H, W, A, NA;
  int feat_stride;
  int dataset_mode;        // 0: SNUBH (fg rows + explicit bg rows), 1: fg-only datasets
  double positive_overlap, negative_overlap;
  int clobber_positives;
  float* labels;
  int* argmax_gt;
  double* max_overlap;
  unsigned long long* colmax;   // [B,max_gt] bit patterns
  float base[MAX_ANCHORS * 4];
};

struct DBox { double x1, y1, x2, y2; };

__device__ __forceinline__ double cmin(double a, double b) { return b < a ? b : a; }
__device__ __forceinline__ double cmax(double a, double b) { return b > a ? b : a; }

__device__ __forceinline__ double iou64(const DBox& b, const DBox& q, bool ui) {
  const double iw = __dadd_rn(__dsub_rn(cmin(b.x2, q.x2), cmax(b.x1, q.x1)), 1.0);
  if (!(iw > 0)) return 0.0;
  const double ih = __dadd_rn(__dsub_rn(cmin(b.y2, q.y2), cmax(b.y1, q.y1)), 1.0);
  if (!(ih > 0)) return 0.0;
  const double barea = __dmul_rn(__dadd_rn(__dsub_rn(b.x2, b.x1), 1.0),
                                 __dadd_rn(__dsub_rn(b.y2, b.y1), 1.0));
  const double inter = __dmul_rn(iw, ih);
  if (ui) return __ddiv_rn(inter, barea);
  const double qarea = __dmul_rn(__dadd_rn(__dsub_rn(q.x2, q.x1), 1.0),
                                 __dadd_rn(__dsub_rn(q.y2, q.y1), 1.0));
  return __ddiv_rn(inter, __dsub_rn(__dadd_rn(barea, qarea), inter));
}

__device__ __forceinline__ DBox anchor_box(const AtParams& p, int a) {
  const int cell = a / p.A;
  const int an = a - cell * p.A;
  const int y = cell / p.W;
  const int x = cell - y * p.W;
  const double sx = (double)(x * p.feat_stride), sy = (double)(y * p.feat_stride);
  DBox b;
  b.x1 = (double)p.base[4 * an] + sx;
  b.y1 = (double)p.base[4 * an + 1] + sy;
  b.x2 = (double)p.base[4 * an + 2] + sx;
  b.y2 = (double)p.base[4 * an + 3] + sy;
  return b;
}

struct GtSet {
  int n_all, n_pos;
};

// loads this image's GT rows into shared memory, returns counts (fg rows first, :434-436)
__device__ __forceinline__ GtSet load_gt(const AtParams& p, int img, DBox* s_gt) {
  __shared__ int s_npos;
  const int n_all = min(max(p.num_gt[img], 0), min(p.max_gt, MAX_GT));
  if (threadIdx.x == 0) s_npos = 0;
  __syncthreads();
  if (threadIdx.x < n_all) {
    const float* g = p.gt_boxes + ((size_t)img * p.max_gt + threadIdx.x) * 5;
    DBox b;
    b.x1 = g[0]; b.y1 = g[1]; b.x2 = g[2]; b.y2 = g[3];
    s_gt[threadIdx.x] = b;
    if (g[4] != 0.f) atomicAdd(&s_npos, 1);       // num_pos = sum(cls != 0)
  }
  __syncthreads();
  GtSet r;
  r.n_all = n_all;
  r.n_pos = p.dataset_mode == 2 ? n_all : s_npos;
  return r;
}

__device__ __forceinline__ bool inside_image(const DBox& b, float im_h, float im_w) {
  // :410-415 with _allowed_border = 0; fp64 anchor against the fp32 im_info entries
  return b.x1 >= 0.0 && b.y1 >= 0.0 && b.x2 < (double)im_w && b.y2 < (double)im_h;
}

template <int PASS>
__global__ void __launch_bounds__(256)
anchor_labels_kernel(const AtParams p) {
  __shared__ DBox s_gt[MAX_GT];
  __shared__ double s_colmax[MAX_GT];
  const int img = blockIdx.y;
  const GtSet gs = load_gt(p, img, s_gt);
  // dataset_mode 2 (plain fg-only, e.g. UDIAT) uses all rows; 0/1 use the fg prefix
  const int n_pos = gs.n_pos;
  const float* info = p.im_info + (size_t)img * p.info_stride;
  const float im_h = info[0], im_w = info[1];
  if (PASS == 2) {
    if (threadIdx.x < n_pos)
      s_colmax[threadIdx.x] =
          __longlong_as_double((long long)p.colmax[(size_t)img * p.max_gt + threadIdx.x]);
    __syncthreads();
  }
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= p.NA) return;
  const DBox b = anchor_box(p, a);
  const bool inside = inside_image(b, im_h, im_w);
  const size_t o = (size_t)img * p.NA + a;
  if (PASS == 1) {
    if (!inside) return;
    for (int k = 0; k < n_pos; ++k) {
      const double ov = iou64(b, s_gt[k], false);
      if (ov > 0.0)
        atomicMax(&p.colmax[(size_t)img * p.max_gt + k],
                  (unsigned long long)__double_as_longlong(ov));
    }
    return;
  }
  if (!inside) {
    p.labels[o] = -1.f;                 // _unmap fill (:568)
    p.argmax_gt[o] = -1;
    if (p.max_overlap) p.max_overlap[o] = 0.0;
    return;
  }
  // row max / first argmax over the fg rows (:444-445), "is a per-GT best" flag (:446-449)
  double best = -1.0;
  int best_k = 0;
  bool is_gt_best = false;
  for (int k = 0; k < n_pos; ++k) {
    const double ov = iou64(b, s_gt[k], false);
    if (ov > best) { best = ov; best_k = k; }
    if (ov == s_colmax[k]) is_gt_best = true;
  }
  if (n_pos == 0) best = 0.0;
  float label = -1.f;
  if (p.dataset_mode == 0) {
    // explicit background boxes: uni-directional overlap >= positive_overlap -> 0 (:451-461)
    const int n_neg = gs.n_all - n_pos;
    if (n_neg > 0 && !p.clobber_positives) {
      double bneg = 0.0;
      for (int k = n_pos; k < gs.n_all; ++k) {
        const double ov = iou64(b, s_gt[k], true);
        if (ov > bneg) bneg = ov;
      }
      if (bneg >= p.positive_overlap) label = 0.f;
    }
    if (is_gt_best) label = 1.f;                       // :464
    if (best >= p.positive_overlap) label = 1.f;       // :467
  } else {
    if (!p.clobber_positives && best < p.negative_overlap) label = 0.f;   // :497-499
    if (is_gt_best) label = 1.f;                                           // :502
    if (best >= p.positive_overlap) label = 1.f;                           // :505
    if (p.clobber_positives && best < p.negative_overlap) label = 0.f;    // :507-509
  }
  p.labels[o] = label;
  p.argmax_gt[o] = best_k;
  if (p.max_overlap) p.max_overlap[o] = best;
}

}  // namespace

extern "C" size_t wssdl_anchor_labels_workspace_bytes(int B, int H, int W, int A, int max_gt) {
  (void)H; (void)W; (void)A;
  if (B <= 0 || max_gt <= 0) return 256;
  return ((sizeof(unsigned long long) * (size_t)B * (size_t)max_gt) + 255) & ~(size_t)255;
}

extern "C" int wssdl_anchor_labels(const float* gt_boxes, const int* num_gt, int max_gt,
                                   const float* im_info, int info_stride, int B, int H, int W,
                                   int A, const float* base_anchors, int feat_stride,
                                   int dataset_mode, double positive_overlap,
                                   double negative_overlap, int clobber_positives, float* labels,
                                   int* argmax_gt, double* max_overlap, void* workspace,
                                   size_t workspace_bytes, wssdl_stream_t stream) {
  if (B < 0 || H <= 0 || W <= 0 || A <= 0 || max_gt <= 0 || info_stride < 2) return WSSDL_EINVAL;
  if (dataset_mode < 0 || dataset_mode > 2) return WSSDL_EINVAL;
  if (B == 0) return WSSDL_OK;
  if (!gt_boxes || !num_gt || !im_info || !base_anchors || !labels || !argmax_gt || !workspace)
    return WSSDL_EINVAL;
  if (A > MAX_ANCHORS || max_gt > MAX_GT || (long long)H * W * A >= (1ll << 31) || B > 65535)
    return WSSDL_ELIMIT;
  const size_t need = wssdl_anchor_labels_workspace_bytes(B, H, W, A, max_gt);
  if (workspace_bytes < need) return WSSDL_EWORKSPACE;
  cudaStream_t s = to_cuda(stream);
  AtParams p;
  p.gt_boxes = gt_boxes; p.num_gt = num_gt; p.max_gt = max_gt; p.im_info = im_info;
  p.info_stride = info_stride; p.H = H; p.W = W; p.A = A; p.NA = H * W * A;
  p.feat_stride = feat_stride; p.dataset_mode = dataset_mode;
  p.positive_overlap = positive_overlap; p.negative_overlap = negative_overlap;
  p.clobber_positives = clobber_positives;
  p.labels = labels; p.argmax_gt = argmax_gt; p.max_overlap = max_overlap;
  p.colmax = static_cast<unsigned long long*>(workspace);
  for (int i = 0; i < MAX_ANCHORS * 4; ++i) p.base[i] = i < 4 * A ? base_anchors[i] : 0.f;
  WSSDL_RETURN_IF_CUDA(cudaMemsetAsync(workspace, 0, need, s));
  dim3 grid((unsigned)ceil_div(p.NA, 256), (unsigned)B);
  anchor_labels_kernel<1><<<grid, 256, 0, s>>>(p);
  anchor_labels_kernel<2><<<grid, 256, 0, s>>>(p);
  WSSDL_CHECK_LAUNCH();
  return WSSDL_OK;
}
