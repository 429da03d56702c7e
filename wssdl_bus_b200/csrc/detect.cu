// Per-class detection post-processing for sm_100a: the step right after the RCNN head in
// the reference's test loop, batched over images.
//
//   im_detect tail   fast_rcnn/test_bus.py:207-223  rois / im_scale -> bbox_transform_inv
//                                                   (per class) -> _clip_boxes (:124-134)
//   test_net body    fast_rcnn/test_bus.py:360-401  per class: score > thresh -> nms(0.3)
//                                                   [-> class-agnostic nms] -> cap at
//                                                   max_per_image over all classes
//   (twin in fast_rcnn/train_bus.py:453-514)
//
// One CTA per image; everything lives in shared memory (<= 1024 candidates per NMS
// problem): order-preserving compaction of the RoIs that pass the score threshold,
// decode + clip, bitonic sort on (score desc, index desc) keys, 64-bit suppression mask,
// one-warp sweep (same fixed-point resolution of the diagonal blocks as nms.cu).  The
// per-class results are written to a fixed-stride blob [B, K, S, 5] + counts [B, K]: the
// tensors the multi-GPU path all-gathers for evaluation.  Latency bound (a few hundred
// boxes per class); the point is that nothing returns to the host between the head and the
// gathered detections.
#include "box_common.cuh"
#include "common.cuh"
#include "nms_common.cuh"

namespace {

constexpr int DET_THREADS = 512;
constexpr int DET_WARPS = DET_THREADS / 32;
constexpr int DET_MAX_N = 1024;   // candidates per NMS problem

struct DetShared {
  unsigned long long* keys;   // [NP]
  float4* box;                // [nmax] compacted order
  float4* sbox;               // [nmax] sorted order
  float* score;               // [nmax]
  float* sscore;              // [nmax]
  float* sarea;               // [nmax]
  int* cls;                   // [nmax] (class-agnostic pass)
  int* scls;                  // [nmax]
  unsigned long long* mask;   // [nmax * nblk]
  unsigned long long* remv;   // [nblk]
  int* keep;                  // [nmax]
};

__device__ __forceinline__ DetShared carve(unsigned char* base, int nmax, int NP, int nblk) {
  DetShared s;
  unsigned char* p = base;
  s.keys = reinterpret_cast<unsigned long long*>(p); p += sizeof(unsigned long long) * NP;
  s.mask = reinterpret_cast<unsigned long long*>(p); p += sizeof(unsigned long long) * (size_t)nmax * nblk;
  s.remv = reinterpret_cast<unsigned long long*>(p); p += sizeof(unsigned long long) * nblk;
  p = base + ((static_cast<size_t>(p - base) + 15) & ~static_cast<size_t>(15));   // float4 arrays
  s.box = reinterpret_cast<float4*>(p); p += sizeof(float4) * nmax;
  s.sbox = reinterpret_cast<float4*>(p); p += sizeof(float4) * nmax;
  s.score = reinterpret_cast<float*>(p); p += sizeof(float) * nmax;
  s.sscore = reinterpret_cast<float*>(p); p += sizeof(float) * nmax;
  s.sarea = reinterpret_cast<float*>(p); p += sizeof(float) * nmax;
  s.cls = reinterpret_cast<int*>(p); p += sizeof(int) * nmax;
  s.scls = reinterpret_cast<int*>(p); p += sizeof(int) * nmax;
  s.keep = reinterpret_cast<int*>(p);
  return s;
}

size_t det_smem_bytes(int nmax, int NP, int nblk) {
  return sizeof(unsigned long long) * ((size_t)NP + (size_t)nmax * nblk + nblk) +
         sizeof(float4) * 2 * (size_t)nmax + sizeof(float) * 3 * (size_t)nmax +
         sizeof(int) * 3 * (size_t)nmax + 48;
}

// Position of this thread's element among the flagged ones of the whole block, in thread
// order; *total = number flagged.  s_w: DET_WARPS ints of scratch.
__device__ __forceinline__ int block_rank(bool flag, int* s_w, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, flag);
  if (lane == 0) s_w[warp] = __popc(bal);
  __syncthreads();
  int off = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < DET_WARPS; ++i) {
    const int c = s_w[i];
    if (i < warp) off += c;
    tot += c;
  }
  __syncthreads();
  *total = tot;
  return off + __popc(bal & ((1u << lane) - 1u));
}

// Ascending bitonic sort of keys[0..NP) (NP a power of two) by the whole block.
__device__ void block_sort(unsigned long long* keys, int NP) {
  for (int k = 2; k <= NP; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < NP / 2; t += blockDim.x) {
        const int i = 2 * t - (t & (j - 1));
        const bool up = (i & k) == 0;
        const unsigned long long a = keys[i], b = keys[i + j];
        if ((a > b) == up) { keys[i] = b; keys[i + j] = a; }
      }
      __syncthreads();
    }
  }
}

// Greedy NMS of n boxes already sorted by descending score (sbox, sarea in shared memory).
// keep[0..return) = kept sorted positions, ascending.  All threads must call; all get the
// count.  `zero` is raised when some pair has a zero union.
__device__ int block_nms(int n, const DetShared& s, const Thresh th, int* s_cnt, bool* zero_any) {
  const int nblk = (n + 63) >> 6;
  const int tid = threadIdx.x, lane = tid & 31;
  bool zero = false;
  for (int item = tid; item < n * nblk; item += blockDim.x) {
    const int q = item / nblk, cb = item - q * nblk;
    unsigned long long bits = 0;
    if (cb >= (q >> 6)) {
      const float4 bi = s.sbox[q];
      const float ai = s.sarea[q];
      const int c0 = max(q + 1, 64 * cb), c1 = min(n, 64 * cb + 64);
      for (int c = c0; c < c1; ++c)
        if (suppresses(bi, ai, s.sbox[c], s.sarea[c], th, &zero)) bits |= 1ull << (c & 63);
    }
    s.mask[item] = bits;
  }
  if (zero) *zero_any = true;
  if (tid < nblk) s.remv[tid] = 0ull;
  __syncthreads();
  if (tid < 32) {
    int count = 0;
    for (int b = 0; b < nblk; ++b) {
      const int r0 = 64 * b + lane, r1 = r0 + 32;
      const unsigned long long d0 = r0 < n ? s.mask[r0 * nblk + b] : 0ull;
      const unsigned long long d1 = r1 < n ? s.mask[r1 * nblk + b] : 0ull;
      const int nrow = min(64, n - 64 * b);
      const unsigned long long valid = nrow == 64 ? ~0ull : ((1ull << nrow) - 1ull);
      const unsigned long long alive = ~s.remv[b] & valid;
      unsigned long long K = alive;
      for (int it = 0; it < 64; ++it) {          // unique fixed point = greedy solution
        unsigned long long mine = 0;
        if ((K >> lane) & 1ull) mine |= d0;
        if ((K >> (lane + 32)) & 1ull) mine |= d1;
        const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)mine);
        const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(mine >> 32));
        const unsigned long long Knew = alive & ~(((unsigned long long)hi << 32) | lo);
        if (Knew == K) break;
        K = Knew;
      }
      if ((K >> lane) & 1ull) s.keep[count + __popcll(K & ((1ull << lane) - 1ull))] = r0;
      if ((K >> (lane + 32)) & 1ull)
        s.keep[count + __popcll(K & ((1ull << (lane + 32)) - 1ull))] = r1;
      count += __popcll(K);
      for (int j = b + 1 + lane; j < nblk; j += 32) {
        unsigned long long acc = 0, rem = K;
        while (rem) {
          const int i = __ffsll((long long)rem) - 1;
          rem &= rem - 1;
          acc |= s.mask[(64 * b + i) * nblk + j];
        }
        s.remv[j] |= acc;
      }
      __syncwarp();
    }
    if (lane == 0) *s_cnt = count;
  }
  __syncthreads();
  return *s_cnt;
}

// np.maximum(v, 0) / np.minimum(v, hi): NaN propagates (test_bus.py:127-133)
__device__ __forceinline__ float np_max0(float v) { return (v > 0.f || v != v) ? v : 0.f; }
__device__ __forceinline__ float np_min(float v, float hi) { return (v < hi || v != v) ? v : hi; }

struct DetArgs {
  const float* rois; const int* roi_counts; int roi_stride;
  const float* scores; const float* bbox_pred; const float* im_meta;
  int B, K; float score_thresh; Thresh nms; int max_per_image, cls_agnostic;
  int nmax, NP, nblk;
  float* dets; int* det_counts; float* pred_boxes; int* status;
};

__global__ void __launch_bounds__(DET_THREADS)
detect_postprocess_kernel(const DetArgs a) {
  extern __shared__ __align__(16) unsigned char d_smem[];
  __shared__ int s_w[DET_WARPS];
  __shared__ int s_cnt;
  __shared__ int s_class_cnt[64];
  const DetShared s = carve(d_smem, a.nmax, a.NP, a.nblk);
  const int b = blockIdx.x, tid = threadIdx.x;
  const int S = a.roi_stride, K = a.K;
  const int n_rois = a.roi_counts ? min(max(a.roi_counts[b], 0), S) : S;
  const float im_h = a.im_meta[b * 3], im_w = a.im_meta[b * 3 + 1], im_scale = a.im_meta[b * 3 + 2];
  const float x_hi = __fsub_rn(im_w, 1.0f), y_hi = __fsub_rn(im_h, 1.0f);
  float* dets_b = a.dets + (size_t)b * K * S * 5;
  bool zero_any = false;
  if (tid < 64) s_class_cnt[tid] = 0;
  __syncthreads();

  if (a.pred_boxes != nullptr) {
    // the reference regresses the background column too (:222); only reported, never used
    for (int r = tid; r < n_rois; r += DET_THREADS) {
      const size_t row = (size_t)b * S + r;
      const float* roi = a.rois + row * 5;
      const float* d = a.bbox_pred + row * 4 * K;
      float4 pb = decode_box(__fdiv_rn(roi[1], im_scale), __fdiv_rn(roi[2], im_scale),
                             __fdiv_rn(roi[3], im_scale), __fdiv_rn(roi[4], im_scale), d[0], d[1],
                             d[2], d[3]);
      pb.x = np_max0(pb.x); pb.y = np_max0(pb.y);
      pb.z = np_min(pb.z, x_hi); pb.w = np_min(pb.w, y_hi);
      *reinterpret_cast<float4*>(a.pred_boxes + row * 4 * K) = pb;
    }
  }

  for (int j = 1; j < K; ++j) {                                        // :360 skip background
    // ---- threshold, order-preserving compaction, decode + clip
    int n = 0;
    for (int r0 = 0; r0 < n_rois; r0 += DET_THREADS) {
      const int r = r0 + tid;
      bool pass = false;
      float sc = 0.f;
      float4 pb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < n_rois) {
        const size_t row = (size_t)b * S + r;
        sc = a.scores[row * K + j];
        pass = sc > a.score_thresh;                                    // :361
        if (pass || a.pred_boxes != nullptr) {
          const float* roi = a.rois + row * 5;
          const float x1 = __fdiv_rn(roi[1], im_scale), y1 = __fdiv_rn(roi[2], im_scale);   // :209
          const float x2 = __fdiv_rn(roi[3], im_scale), y2 = __fdiv_rn(roi[4], im_scale);
          const float* d = a.bbox_pred + row * 4 * K + 4 * j;
          pb = decode_box(x1, y1, x2, y2, d[0], d[1], d[2], d[3]);     // :222
          pb.x = np_max0(pb.x); pb.y = np_max0(pb.y);                  // :223 _clip_boxes
          pb.z = np_min(pb.z, x_hi); pb.w = np_min(pb.w, y_hi);
          if (a.pred_boxes != nullptr)
            *reinterpret_cast<float4*>(a.pred_boxes + row * 4 * K + 4 * j) = pb;
        }
      }
      int tot;
      const int pos = n + block_rank(pass, s_w, &tot);
      if (pass && pos < a.nmax) {
        s.box[pos] = pb;
        s.score[pos] = sc;
        s.keys[pos] = ((unsigned long long)(~orderable(sc)) << 32) | (unsigned)(~(unsigned)pos);
      }
      n += tot;
    }
    n = min(n, a.nmax);   // host guarantees nmax >= S, so nothing is ever dropped here
    int NP = 1;
    while (NP < n) NP <<= 1;
    for (int i = n + tid; i < NP; i += DET_THREADS) s.keys[i] = ~0ull;
    __syncthreads();
    block_sort(s.keys, NP);
    for (int q = tid; q < n; q += DET_THREADS) {
      const int p = (int)(~(unsigned)(s.keys[q] & 0xffffffffull));
      const float4 bx = s.box[p];
      s.sbox[q] = bx;
      s.sscore[q] = s.score[p];
      s.sarea[q] = __fmul_rn(__fadd_rn(__fsub_rn(bx.z, bx.x), 1.0f),
                             __fadd_rn(__fsub_rn(bx.w, bx.y), 1.0f));   // nms.pyx:24
    }
    __syncthreads();
    const int nk = block_nms(n, s, a.nms, &s_cnt, &zero_any);          // :366
    float* out = dets_b + (size_t)j * S * 5;
    for (int i = tid; i < S; i += DET_THREADS) {
      float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
      float sc = 0.f;
      if (i < nk) { const int q = s.keep[i]; bx = s.sbox[q]; sc = s.sscore[q]; }
      out[i * 5] = bx.x; out[i * 5 + 1] = bx.y; out[i * 5 + 2] = bx.z; out[i * 5 + 3] = bx.w;
      out[i * 5 + 4] = sc;
    }
    if (tid == 0) s_class_cnt[j] = nk;
    __syncthreads();
  }
  // class 0 (background) is never reported
  for (int i = tid; i < S * 5; i += DET_THREADS) dets_b[i] = 0.f;

  if (a.cls_agnostic) {                                                // :371-386
    // concatenate the per-class survivors in class order, NMS over all, split back
    int n = 0;
    for (int j = 1; j < K; ++j) {
      const int cj = s_class_cnt[j];
      const float* src = dets_b + (size_t)j * S * 5;
      for (int i = tid; i < cj; i += DET_THREADS) {
        const int p = n + i;
        if (p < a.nmax) {
          s.box[p] = make_float4(src[i * 5], src[i * 5 + 1], src[i * 5 + 2], src[i * 5 + 3]);
          const float sc = src[i * 5 + 4];
          s.score[p] = sc;
          s.cls[p] = j;
          s.keys[p] = ((unsigned long long)(~orderable(sc)) << 32) | (unsigned)(~(unsigned)p);
        }
      }
      n += cj;
    }
    n = min(n, a.nmax);
    int NP = 1;
    while (NP < n) NP <<= 1;
    for (int i = n + tid; i < NP; i += DET_THREADS) s.keys[i] = ~0ull;
    __syncthreads();
    block_sort(s.keys, NP);
    for (int q = tid; q < n; q += DET_THREADS) {
      const int p = (int)(~(unsigned)(s.keys[q] & 0xffffffffull));
      const float4 bx = s.box[p];
      s.sbox[q] = bx;
      s.sscore[q] = s.score[p];
      s.scls[q] = s.cls[p];
      s.sarea[q] = __fmul_rn(__fadd_rn(__fsub_rn(bx.z, bx.x), 1.0f),
                             __fadd_rn(__fsub_rn(bx.w, bx.y), 1.0f));
    }
    __syncthreads();
    const int nk = block_nms(n, s, a.nms, &s_cnt, &zero_any);
    for (int j = 1; j < K; ++j) {
      float* out = dets_b + (size_t)j * S * 5;
      int cnt = 0;
      for (int i0 = 0; i0 < nk; i0 += DET_THREADS) {
        const int i = i0 + tid;
        const int q = i < nk ? s.keep[i] : 0;
        const bool mine = i < nk && s.scls[q] == j;
        int tot;
        const int pos = cnt + block_rank(mine, s_w, &tot);
        if (mine) {
          const float4 bx = s.sbox[q];
          out[pos * 5] = bx.x; out[pos * 5 + 1] = bx.y; out[pos * 5 + 2] = bx.z;
          out[pos * 5 + 3] = bx.w; out[pos * 5 + 4] = s.sscore[q];
        }
        cnt += tot;
      }
      __syncthreads();
      for (int i = cnt + tid; i < S; i += DET_THREADS) {
        out[i * 5] = 0.f; out[i * 5 + 1] = 0.f; out[i * 5 + 2] = 0.f; out[i * 5 + 3] = 0.f;
        out[i * 5 + 4] = 0.f;
      }
      if (tid == 0) s_class_cnt[j] = cnt;
      __syncthreads();
    }
  }

  if (a.max_per_image > 0) {                                           // :394-401
    __syncthreads();
    int total = 0;
    for (int j = 1; j < K; ++j) total += s_class_cnt[j];
    if (total > a.max_per_image) {
      // image_thresh = np.sort(image_scores)[-max_per_image]
      int n = 0;
      for (int j = 1; j < K; ++j) {
        const int cj = s_class_cnt[j];
        const float* src = dets_b + (size_t)j * S * 5;
        for (int i = tid; i < cj; i += DET_THREADS)
          s.keys[n + i] = (unsigned long long)(~orderable(src[i * 5 + 4]));
        n += cj;
      }
      int NP = 1;
      while (NP < n) NP <<= 1;
      for (int i = n + tid; i < NP; i += DET_THREADS) s.keys[i] = ~0ull;
      __syncthreads();
      block_sort(s.keys, NP);
      const unsigned ot = ~(unsigned)s.keys[a.max_per_image - 1];      // orderable(image_thresh)
      const float image_thresh = __uint_as_float((ot & 0x80000000u) ? (ot ^ 0x80000000u) : ~ot);
      // every class list is in descending-score order: `score >= image_thresh` keeps a prefix
      for (int j = 1; j < K; ++j) {
        const int cj = s_class_cnt[j];
        float* out = dets_b + (size_t)j * S * 5;
        int cnt = 0;
        for (int i0 = 0; i0 < cj; i0 += DET_THREADS) {
          const int i = i0 + tid;
          const bool ok = i < cj && out[i * 5 + 4] >= image_thresh;   // :399
          int tot;
          block_rank(ok, s_w, &tot);
          cnt += tot;
        }
        for (int i = cnt + tid; i < cj; i += DET_THREADS) {
          out[i * 5] = 0.f; out[i * 5 + 1] = 0.f; out[i * 5 + 2] = 0.f; out[i * 5 + 3] = 0.f;
          out[i * 5 + 4] = 0.f;
        }
        __syncthreads();
        if (tid == 0) s_class_cnt[j] = cnt;
        __syncthreads();
      }
    }
  }
  __syncthreads();
  if (tid < K) a.det_counts[b * K + tid] = tid == 0 ? 0 : s_class_cnt[tid];
  if (zero_any && a.status != nullptr) a.status[0] = 1;
}

}  // namespace

extern "C" int wssdl_detect_postprocess(const float* rois, const int* roi_counts, int roi_stride,
                                        const float* scores, const float* bbox_pred,
                                        const float* im_meta, int B, int K, float score_thresh,
                                        double nms_thresh, int max_per_image, int cls_agnostic,
                                        float* dets, int* det_counts, float* pred_boxes,
                                        int* status, wssdl_stream_t stream) {
  if (B < 0 || K < 1 || roi_stride < 0) return WSSDL_EINVAL;
  cudaStream_t s = to_cuda(stream);
  if (status) WSSDL_RETURN_IF_CUDA(cudaMemsetAsync(status, 0, sizeof(int), s));
  if (B == 0) return WSSDL_OK;
  if (!det_counts) return WSSDL_EINVAL;
  if (roi_stride == 0 || K == 1) {
    WSSDL_RETURN_IF_CUDA(cudaMemsetAsync(det_counts, 0, sizeof(int) * (size_t)B * K, s));
    return WSSDL_OK;
  }
  if (!rois || !scores || !bbox_pred || !im_meta || !dets) return WSSDL_EINVAL;
  if (K > 64) return WSSDL_ELIMIT;
  DetArgs a;
  a.rois = rois; a.roi_counts = roi_counts; a.roi_stride = roi_stride;
  a.scores = scores; a.bbox_pred = bbox_pred; a.im_meta = im_meta;
  a.B = B; a.K = K; a.score_thresh = score_thresh;
  a.nms = make_thresh(nms_thresh, WSSDL_NMS_GE_F64);
  a.max_per_image = max_per_image; a.cls_agnostic = cls_agnostic ? 1 : 0;
  // One NMS problem at a time lives in shared memory: a class (<= S boxes) or, class-agnostic,
  // all classes together (<= (K-1)*S).  Boxes, scores and mask rows are sized for that; only the
  // sort keys also serve the max_per_image cap, which ranks up to (K-1)*S scores.
  const long long all = (long long)(K - 1) * roi_stride;
  const long long nms_n = cls_agnostic ? all : (long long)roi_stride;
  if (roi_stride > DET_MAX_N || nms_n > DET_MAX_N) return WSSDL_ELIMIT;
  a.nmax = (int)nms_n;
  const long long nkeys = (cls_agnostic || max_per_image > 0) ? all : (long long)roi_stride;
  if (nkeys > (1 << 16)) return WSSDL_ELIMIT;
  a.NP = 1;
  while (a.NP < nkeys) a.NP <<= 1;
  a.nblk = (int)((nms_n + 63) / 64);
  a.dets = dets; a.det_counts = det_counts; a.pred_boxes = pred_boxes; a.status = status;
  const size_t smem = det_smem_bytes(a.nmax, a.NP, a.nblk);
  if (smem > 227 * 1024 - 1024) return WSSDL_ELIMIT;
  static unsigned long long done = 0;
  int dev = 0;
  WSSDL_RETURN_IF_CUDA(cudaGetDevice(&dev));
  if (!(__atomic_load_n(&done, __ATOMIC_ACQUIRE) & (1ull << (dev & 63)))) {
    WSSDL_RETURN_IF_CUDA(cudaFuncSetAttribute(detect_postprocess_kernel,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              227 * 1024 - 1024));
    __atomic_fetch_or(&done, 1ull << (dev & 63), __ATOMIC_RELEASE);
  }
  detect_postprocess_kernel<<<B, DET_THREADS, smem, s>>>(a);
  WSSDL_CHECK_LAUNCH();
  return WSSDL_OK;
}
