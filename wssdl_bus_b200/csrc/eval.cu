// Detection-vs-ground-truth matching for the VOC AP / CorLoc / FROC evaluation, on the
// fixed-stride detection blob the ranks all-gather (csrc/detect.cu).
//
// Semantics: the per-detection loop of voc_eval_bus (datasets/voc_eval_bus.py:206-247) and its
// CorLoc pass (:161-204).  The reference walks ALL detections of a class in descending
// confidence; the only state it carries is the per-image "this GT is already detected" flag,
// so the walk factors into independent per-image walks in the same relative order -- which is
// the order the blob already has (class lists are sorted by descending score).  One warp per
// (image, class): lanes hold the image's GT boxes of that class, every detection is scored
// against them in fp64 with the reference's operation order (one rounding per operation, no
// FMA), the warp reduces (max overlap, FIRST index) like np.max / np.argmax, lane 0 applies
// the TP / FP / difficult / duplicate rules.  The O(n log n) part of the metric (global
// argsort by confidence, cumulative sums, AP) stays on the host in numpy, with the
// reference's own calls (wssdl_bus_b200/datasets/voc_eval_bus.py).
#include "common.cuh"

namespace {

constexpr int EV_MAX_GT = 64;   // GT boxes of one class in one image

struct EvalArgs {
  const float* dets; const int* det_counts; int B, K, S;
  const float* gt_boxes; const int* num_gt; const unsigned char* difficult; int G;
  double ovthresh; float score_thresh;
  unsigned char* tp; unsigned char* fp; unsigned char* fp_froc;
  int* img_stats; int* npos;
};

__device__ __forceinline__ double np_max(double a, double b) { return (a > b || a != a) ? a : b; }  // np.maximum
__device__ __forceinline__ double np_min(double a, double b) { return (a < b || a != a) ? a : b; }  // np.minimum

// inters / uni of voc_eval_bus.py:223-238 for one (detection, GT) pair
__device__ __forceinline__ double voc_overlap(const double bb[4], const double gt[4]) {
  const double ixmin = np_max(gt[0], bb[0]), iymin = np_max(gt[1], bb[1]);
  const double ixmax = np_min(gt[2], bb[2]), iymax = np_min(gt[3], bb[3]);
  const double iw = np_max(__dadd_rn(__dsub_rn(ixmax, ixmin), 1.0), 0.0);
  const double ih = np_max(__dadd_rn(__dsub_rn(iymax, iymin), 1.0), 0.0);
  const double inters = __dmul_rn(iw, ih);
  const double abb = __dmul_rn(__dadd_rn(__dsub_rn(bb[2], bb[0]), 1.0), __dadd_rn(__dsub_rn(bb[3], bb[1]), 1.0));
  const double agt = __dmul_rn(__dadd_rn(__dsub_rn(gt[2], gt[0]), 1.0), __dadd_rn(__dsub_rn(gt[3], gt[1]), 1.0));
  const double uni = __dsub_rn(__dadd_rn(abb, agt), inters);
  return __ddiv_rn(inters, uni);
}

__global__ void __launch_bounds__(32)
eval_match_kernel(const EvalArgs a) {
  __shared__ double s_gt[EV_MAX_GT][4];
  __shared__ unsigned char s_diff[EV_MAX_GT];
  const int b = blockIdx.x, j = blockIdx.y + 1;        // class 0 is background
  const int lane = threadIdx.x;
  // ---- this image's GT boxes of class j, in annotation order (jmax = first maximum)
  const int ng_all = min(max(a.num_gt[b], 0), a.G);
  int ng = 0;
  for (int g0 = 0; g0 < ng_all; g0 += 32) {
    const int g = g0 + lane;
    const float* row = a.gt_boxes + ((size_t)b * a.G + min(g, ng_all - 1)) * 5;
    const bool mine = g < ng_all && (int)row[4] == j;
    const unsigned bal = __ballot_sync(0xffffffffu, mine);
    const int pos = ng + __popc(bal & ((1u << lane) - 1u));
    if (mine && pos < EV_MAX_GT) {
      s_gt[pos][0] = row[0]; s_gt[pos][1] = row[1]; s_gt[pos][2] = row[2]; s_gt[pos][3] = row[3];
      s_diff[pos] = a.difficult ? a.difficult[(size_t)b * a.G + g] : 0;
    }
    ng += __popc(bal);
  }
  ng = min(ng, EV_MAX_GT);
  __syncwarp();
  int npos_local = 0;
  for (int g = lane; g < ng; g += 32) npos_local += s_diff[g] ? 0 : 1;
  for (int o = 16; o > 0; o >>= 1) npos_local += __shfl_xor_sync(0xffffffffu, npos_local, o);
  if (lane == 0 && npos_local) atomicAdd(&a.npos[j], npos_local);                 // :135

  const int cnt = min(max(a.det_counts[b * a.K + j], 0), a.S);
  const size_t base = ((size_t)b * a.K + j) * a.S;
  const float* dets = a.dets + base * 5;
  unsigned long long detected = 0ull;                  // R['det'], lane 0's copy is the truth
  bool corloc_ok = false;
  for (int i = 0; i < cnt; ++i) {
    double bb[4];
    bb[0] = dets[i * 5]; bb[1] = dets[i * 5 + 1]; bb[2] = dets[i * 5 + 2]; bb[3] = dets[i * 5 + 3];
    const float score = dets[i * 5 + 4];
    double ovmax = -INFINITY;                                                     // :215
    int jmax = 0x7fffffff;
    for (int g = lane; g < ng; g += 32) {
      const double ov = voc_overlap(bb, s_gt[g]);
      // np.max / np.argmax: NaN wins and sticks, otherwise first maximum
      if (ov > ovmax || (ov != ov && ovmax == ovmax)) { ovmax = ov; jmax = g; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ov2 = __shfl_xor_sync(0xffffffffu, ovmax, o);
      const int j2 = __shfl_xor_sync(0xffffffffu, jmax, o);
      const bool nan1 = ovmax != ovmax, nan2 = ov2 != ov2;
      const bool take = (nan1 || nan2) ? (nan2 && (!nan1 || j2 < jmax))
                                       : (ov2 > ovmax || (ov2 == ovmax && j2 < jmax));
      if (take) { ovmax = ov2; jmax = j2; }
    }
    if (lane == 0) {
      unsigned char tp = 0, fp = 0, ff = 0;
      if (ovmax > a.ovthresh) {                                                   // :240-247
        if (!s_diff[jmax]) {
          if (!((detected >> jmax) & 1ull)) { tp = 1; detected |= 1ull << jmax; }
          else fp = 1;
        }
      } else {
        fp = 1;
      }
      if (score >= a.score_thresh) {                                              // :250-252
        if (ovmax <= a.ovthresh) ff = 1;
        if (ovmax > a.ovthresh) corloc_ok = true;                                 // :199-203
      }
      a.tp[base + i] = tp; a.fp[base + i] = fp; a.fp_froc[base + i] = ff;
    }
  }
  if (lane == 0) {
    for (int i = cnt; i < a.S; ++i) { a.tp[base + i] = 0; a.fp[base + i] = 0; a.fp_froc[base + i] = 0; }
    a.img_stats[(b * a.K + j) * 2] = ng > 0;                                      // counts toward ni
    a.img_stats[(b * a.K + j) * 2 + 1] = (ng > 0 && corloc_ok) ? 1 : 0;           // counts toward nok
  }
}

}  // namespace

extern "C" int wssdl_eval_match(const float* dets, const int* det_counts, int B, int K, int S,
                                const float* gt_boxes, const int* num_gt,
                                const unsigned char* difficult, int G, double ovthresh,
                                float score_thresh, unsigned char* tp, unsigned char* fp,
                                unsigned char* fp_froc, int* img_stats, int* npos,
                                wssdl_stream_t stream) {
  if (B < 0 || K < 1 || S < 0 || G < 0) return WSSDL_EINVAL;
  cudaStream_t s = to_cuda(stream);
  if (npos) WSSDL_RETURN_IF_CUDA(cudaMemsetAsync(npos, 0, sizeof(int) * (size_t)K, s));
  if (B == 0 || K == 1) return WSSDL_OK;
  if (!det_counts || !num_gt || !tp || !fp || !fp_froc || !img_stats || !npos) return WSSDL_EINVAL;
  if ((S > 0 && !dets) || (G > 0 && !gt_boxes)) return WSSDL_EINVAL;
  if (K - 1 > 65535 || G > EV_MAX_GT) return WSSDL_ELIMIT;
  WSSDL_RETURN_IF_CUDA(cudaMemsetAsync(img_stats, 0, sizeof(int) * 2 * (size_t)B * K, s));
  // class-0 slots of the flag arrays are never written by the kernel
  EvalArgs a;
  a.dets = dets; a.det_counts = det_counts; a.B = B; a.K = K; a.S = S;
  a.gt_boxes = gt_boxes; a.num_gt = num_gt; a.difficult = difficult; a.G = G;
  a.ovthresh = ovthresh; a.score_thresh = score_thresh;
  a.tp = tp; a.fp = fp; a.fp_froc = fp_froc; a.img_stats = img_stats; a.npos = npos;
  dim3 grid((unsigned)B, (unsigned)(K - 1));
  eval_match_kernel<<<grid, 32, 0, s>>>(a);
  WSSDL_CHECK_LAUNCH();
  return WSSDL_OK;
}
