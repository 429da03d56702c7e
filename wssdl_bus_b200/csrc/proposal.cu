// Fused RPN proposal layer for sm_100a: one CTA per image, everything between the RPN
// head outputs and the RoI blob happens in shared memory.
//
// Semantics: proposal_layer (rpn_msr/proposal_layer_tf_bus.py:19-148) =
//   anchors (generate_anchors.py + shifts :55-71) -> bbox_transform_inv
//   (fast_rcnn/bbox_transform.py:30-61) -> clip_boxes (:63-77) -> _filter_boxes
//   (proposal_layer_tf_bus.py:151-156) -> argsort desc + pre-NMS top-N (:129-133) ->
//   cpu_nms (nms/cpu_nms.pyx:17-68) -> post-NMS top-N (:139-146) -> (R,5) blob.
//
// Phases of the kernel (1024 threads, one image):
//   1. a 32-bit orderable score key per anchor in shared memory.  Boxes are NOT stored: they are
//      a pure function of (anchor index, 4 deltas) and are decoded (+clipped, +min-size filtered)
//      bit-identically whenever needed -- lazily, for the candidates of a batch only.
//   2./3. the K = min(pre_nms_topN, #valid) best anchors in the reference's order (score desc,
//      index desc for ties) are produced LAZILY, M at a time: a radix select (4 x 8-bit
//      histogram passes over the shared keys, a second select on the anchor index for ties at
//      the threshold) finds the key that bounds the next M candidates, they are compacted as
//      64-bit (~key, ~index) words and sorted by an in-shared bitonic sort.  The keep-list NMS
//      below stops as soon as post_nms_topN boxes are kept, which on the detector's shapes
//      (6000 -> 300) happens inside the first batch: sorting all 6000 (padded to 8192) up front
//      was 40 % of the kernel (ncu, profiles/r02_proposals.txt).  M = max(1024, pow2 >= 2*post).
//   4. NMS driven by the keep list instead of an N x N mask: post_nms_topN is small
//      (300 / 2000) so a candidate only has to be tested against boxes ALREADY KEPT
//      (<= post_nms_topN) and the loop stops as soon as the list is full.  Candidates are
//      taken in chunks of 256 in score order:
//        A. 4 threads per candidate scan the kept list (early exit on first hit);
//        B. for the survivors, 64-bit column masks of the chunk's own upper triangle;
//        C. one warp resolves the chunk with ballots: per 32-candidate sub-block a
//           fixed-point iteration K <- alive & ~any(col & K), whose unique fixed point is
//           the greedy answer (dependencies only point to earlier candidates);
//        D. newly kept boxes are appended to the list and written to the output blob.
//
// IoU arithmetic = cpu_nms's fp32 sequence with the double-threshold compare
// (cpu_nms.c:2442-2495), one rounding per operation.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"

namespace {

constexpr int PT = 1024;         // threads per CTA
constexpr int CHUNK = 256;       // NMS candidates per round
constexpr int MAX_ANCHORS = 32; // base anchors per cell

struct PropParams {
  const float* cls_prob;
  const float* bbox_pred;
  const float* im_info;
  int info_stride;
  int H, W, A, NA;
  int feat_stride;
  int pre_nms_topN, post_nms_topN;
  int kpad;                      // M: candidates sorted at a time (power of two)
  float thr_ge;                  // smallest float whose double value is >= nms_thresh
  float min_size;
  float pad_batch;               // batch index written into the unused rows of an image's block
  float* rois;
  float* scores;
  int* anchor_idx;
  int* counts;
  float* decoded;                // optional [B,NA,4]
  float base[MAX_ANCHORS * 4];   // base anchors, by value (constant bank, no extra copy)
};

__device__ __forceinline__ unsigned orderable_key(float f) {
  unsigned u = __float_as_uint(f);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return u == 0u ? 1u : u;       // 0 is reserved for "filtered out"
}
__device__ __forceinline__ float key_to_float(unsigned k) {
  unsigned u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

__device__ __forceinline__ float exp_cr(float x) { return (float)exp((double)x); }

__device__ __forceinline__ float clip_np(float v, float hi) {
  float t = (v != v) ? v : (v < hi ? v : hi);     // np.minimum
  return (t != t) ? t : (t > 0.f ? t : 0.f);      // np.maximum(., 0)
}

// decode + clip anchor `a` (row order (h, w, anchor)) of image `img`
__device__ __forceinline__ float4 decode_anchor(const PropParams& p, int img, int a, float im_h,
                                                float im_w) {
  const int cell = a / p.A;
  const int an = a - cell * p.A;
  const int y = cell / p.W;
  const int x = cell - y * p.W;
  const float sx = (float)(x * p.feat_stride), sy = (float)(y * p.feat_stride);
  const float ax1 = p.base[4 * an] + sx, ay1 = p.base[4 * an + 1] + sy;
  const float ax2 = p.base[4 * an + 2] + sx, ay2 = p.base[4 * an + 3] + sy;
  const float* d = p.bbox_pred + ((size_t)img * p.H * p.W + cell) * (4 * p.A) + 4 * an;
  const float dx = __ldg(d), dy = __ldg(d + 1), dw = __ldg(d + 2), dh = __ldg(d + 3);
  const float w = __fadd_rn(__fsub_rn(ax2, ax1), 1.0f);
  const float h = __fadd_rn(__fsub_rn(ay2, ay1), 1.0f);
  const float cx = __fadd_rn(ax1, __fmul_rn(0.5f, w));
  const float cy = __fadd_rn(ay1, __fmul_rn(0.5f, h));
  const float pcx = __fadd_rn(__fmul_rn(dx, w), cx);
  const float pcy = __fadd_rn(__fmul_rn(dy, h), cy);
  const float pw = __fmul_rn(exp_cr(dw), w);
  const float ph = __fmul_rn(exp_cr(dh), h);
  float4 b;
  b.x = clip_np(__fsub_rn(pcx, __fmul_rn(0.5f, pw)), __fsub_rn(im_w, 1.0f));
  b.y = clip_np(__fsub_rn(pcy, __fmul_rn(0.5f, ph)), __fsub_rn(im_h, 1.0f));
  b.z = clip_np(__fadd_rn(pcx, __fmul_rn(0.5f, pw)), __fsub_rn(im_w, 1.0f));
  b.w = clip_np(__fadd_rn(pcy, __fmul_rn(0.5f, ph)), __fsub_rn(im_h, 1.0f));
  return b;
}

__device__ __forceinline__ float rmax(float a, float b) { return a >= b ? a : b; }
__device__ __forceinline__ float rmin(float a, float b) { return a <= b ? a : b; }

__device__ __forceinline__ bool iou_ge(const float4 bi, float ai, const float4 bj, float aj,
                                       float thr_ge) {
  const float xx1 = rmax(bi.x, bj.x), yy1 = rmax(bi.y, bj.y);
  const float xx2 = rmin(bi.z, bj.z), yy2 = rmin(bi.w, bj.w);
  const float w = rmax(0.0f, __fadd_rn(__fsub_rn(xx2, xx1), 1.0f));
  const float h = rmax(0.0f, __fadd_rn(__fsub_rn(yy2, yy1), 1.0f));
  const float inter = __fmul_rn(w, h);
  const float den = __fsub_rn(__fadd_rn(ai, aj), inter);
  return iou_ge_exact(inter, den, thr_ge);
}

// K-th largest (1-based) among n keys in shared memory, considering only keys accepted by
// `accept`; returns the threshold key T and, through *n_greater, how many accepted keys
// are strictly greater than T.  4 passes of 8 bits, MSB first.  All threads must call.
template <typename KeyFn>
__device__ unsigned radix_select(KeyFn key_at, int n, int K, unsigned* s_hist /*256*/,
                                 unsigned* s_bcast /*4*/, int* n_greater, int* n_equal = nullptr) {
  unsigned prefix = 0, prefix_mask = 0;
  int remaining = K;     // rank still to be found inside the current prefix bucket
  int greater = 0;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += PT) s_hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += PT) {
      const unsigned k = key_at(i);
      if (k != 0u && (k & prefix_mask) == prefix) atomicAdd(&s_hist[(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      // scan the 256 bins from the top; lane l owns bins 255-8l .. 248-8l
      const int lane = threadIdx.x;
      unsigned c[8];
      unsigned sum = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) { c[q] = s_hist[255 - (8 * lane + q)]; sum += c[q]; }
      unsigned incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      unsigned before = incl - sum;   // keys in bins above this lane's bins
      if (before < (unsigned)remaining && (unsigned)remaining <= incl) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (before < (unsigned)remaining && (unsigned)remaining <= before + c[q]) {
            s_bcast[0] = 255 - (8 * lane + q);   // digit of the K-th key
            s_bcast[1] = before;                 // accepted keys above that digit
            s_bcast[2] = c[q];                   // accepted keys with that digit
          }
          before += c[q];
        }
      }
    }
    __syncthreads();
    const unsigned digit = s_bcast[0];
    if (n_equal) *n_equal = (int)s_bcast[2];      // after the last pass: keys equal to the result
    greater += (int)s_bcast[1];
    remaining -= (int)s_bcast[1];
    prefix |= digit << shift;
    prefix_mask |= 255u << shift;
    __syncthreads();
  }
  *n_greater = greater;
  return prefix;
}

// CS > 1: a thread-block cluster of CS CTAs (2, 4 or 8) works on ONE image (batches that leave
// SMs idle: one image's pipeline is issue bound on a single SM).  The score keys are loaded once
// per cluster (all-gather through distributed shared memory); every CTA then runs phases 2-3
// redundantly on identical state (select, sort and decode are deterministic);
// the NMS rounds are split: CTA r owns 256 / CS candidates of the round, 4 * CS threads each, for
// BOTH stage A (candidate against the kept list) and stage B (the candidate's column over the
// earlier candidates of the chunk).  Alive words and columns are exchanged through distributed
// shared memory -- every CTA stores its part into every CTA's copy, ONE cluster barrier per
// round, two slot sets so that a CTA one round ahead cannot overwrite what a peer still reads --
// and stages C and D run redundantly, so the kept lists stay identical.  CTA 0 writes the outputs.
constexpr int XCHG_WORDS = CHUNK / 32;

template <int CS>
__global__ void __launch_bounds__(PT, 1)
proposals_kernel(const PropParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned s_sup[XCHG_WORDS];            // cluster rounds: alive words of my candidates
  __shared__ int s_wcnt[PT / 32];                   // per-warp counts of a block-wide compaction
  namespace cg = cooperative_groups;
  int crank = 0;
  if constexpr (CS > 1) {
    crank = (int)cg::this_cluster().block_rank();
    // distributed shared memory may only be touched once every CTA of the cluster has started
    // (compute-sanitizer: "block that might not have entered yet"); the first remote store is
    // far away, but the guarantee has to be explicit
    cg::this_cluster().sync();
  }
  const bool writer = crank == 0;
  if (threadIdx.x < XCHG_WORDS) s_sup[threadIdx.x] = 0u;
  // programmatic dependent launch: a kernel queued behind this one with the attribute (the
  // RoI-pool pre-pass of the fused hot-path entry) may become resident now; it waits for this
  // grid's completion (griddepcontrol.wait) before it reads the RoIs
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // layout: sort buffer (kpad u64) | histogram/scalars | keys (NA u32, rounded to 4) | per-chunk
  // NMS state | kept list
  unsigned long long* s_sort = reinterpret_cast<unsigned long long*>(smem_raw);
  unsigned* s_hist = reinterpret_cast<unsigned*>(s_sort + p.kpad);             // [256]
  unsigned* s_bcast = s_hist + 256;                                            // [4]
  int* s_cnt = reinterpret_cast<int*>(s_bcast + 4);                            // [4]
  unsigned* s_keys = reinterpret_cast<unsigned*>(s_cnt + 4);                   // [NA]
  float4* s_cbox = reinterpret_cast<float4*>(s_keys + ((p.NA + 3) & ~3));      // [CHUNK]
  unsigned long long* s_col = reinterpret_cast<unsigned long long*>(s_cbox + CHUNK);  // [2][CHUNK][4]
  float* s_carea = reinterpret_cast<float*>(s_col + 2 * CHUNK * 4);            // [CHUNK]
  int* s_cidx = reinterpret_cast<int*>(s_carea + CHUNK);                       // [CHUNK]
  unsigned* s_ckey = reinterpret_cast<unsigned*>(s_cidx + CHUNK);              // [CHUNK]
  unsigned* s_alive = s_ckey + CHUNK;                                          // [2][CHUNK/32]
  unsigned* s_kmask = s_alive + 2 * (CHUNK / 32);                              // [CHUNK/32]
  float4* s_kbox = reinterpret_cast<float4*>(s_kmask + CHUNK / 32);            // [post]
  float* s_karea = reinterpret_cast<float*>(s_kbox + p.post_nms_topN);         // [post]

  const int img = blockIdx.x / CS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* info = p.im_info + (size_t)img * p.info_stride;
  const float im_h = info[0], im_w = info[1];
  const float min_size = __fmul_rn(p.min_size, info[2]);    // :123, fp32 product
  const int post = p.post_nms_topN;

  // ---- phase 1: a score key for EVERY anchor.  The min-size filter (:123) is applied lazily,
  // to the candidates of a batch: walking the anchors in score order and skipping the boxes that
  // fail the filter visits the same sequence as filtering first and sorting the rest, so only a
  // few hundred of the 17100 boxes are ever decoded (the full decode was 24 % of the kernel).
  if (tid == 0) s_cnt[1] = 0;
  // (eight independent loads in flight per thread: one at a time the 17 strided loads of a thread
  // were 12 % of the kernel, all of it load latency)
  // Cluster: CTA r loads the scores of its 1/CS share of the anchors only and stores the keys into
  // every CTA's copy (distributed shared memory all-gather, published by one cluster barrier).
  {
    const float* sc_img = p.cls_prob + (size_t)img * p.H * p.W * (2 * p.A) + p.A;
    const int a_lo = CS > 1 ? (int)((long long)p.NA * crank / CS) : 0;
    const int a_hi = CS > 1 ? (int)((long long)p.NA * (crank + 1) / CS) : p.NA;
    unsigned* peer_keys[CS];
    if constexpr (CS > 1) {
#pragma unroll
      for (int r = 0; r < CS; ++r) peer_keys[r] = cg::this_cluster().map_shared_rank(s_keys, r);
    } else {
      peer_keys[0] = s_keys;
    }
    for (int a0 = a_lo; a0 < a_hi; a0 += 8 * PT) {
      float sc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int a = a0 + q * PT + tid;
        const int cell = a / p.A;
        sc[q] = a < a_hi ? __ldg(sc_img + (size_t)a + (size_t)cell * p.A) : 0.f;   // cell*2A + A + an
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int a = a0 + q * PT + tid;
        if (a < a_hi) {
          const unsigned key = orderable_key(sc[q]);
#pragma unroll
          for (int r = 0; r < CS; ++r) peer_keys[r][a] = key;
        }
      }
    }
    if (p.decoded && writer)                  // the intermediate of :116-119, only when asked for
      for (int a = tid; a < p.NA; a += PT)
        reinterpret_cast<float4*>(p.decoded)[(size_t)img * p.NA + a] = decode_anchor(p, img, a, im_h, im_w);
  }
  if constexpr (CS > 1) cg::this_cluster().sync();
  else __syncthreads();

  // ---- phases 2-4: batches of M anchors in descending score order; decode + filter the batch;
  // keep-list NMS over its valid boxes
  const int M = p.kpad;
  int nkept = 0;
  int round = 0;
  int taken_valid = 0;                        // valid candidates consumed so far (<= pre_nms_topN)
  unsigned Tp = 0, T2p = 0;                   // boundary behind the previous batch
  for (int taken = 0; taken < p.NA && taken_valid < p.pre_nms_topN && nkept < post; taken += M) {
    const int want = min(taken + M, p.NA);    // rank of the last anchor of this batch
    // boundary (T, T2): anchor (k, a) is among the best `want` iff k > T or (k == T and
    // a + 1 >= T2): key and, among ties at T, the highest indices first
    unsigned T = 0, T2 = 0;
    if (want < p.NA) {
      int n_gt = 0, n_eq = 0;
      T = radix_select([&](int i) { return s_keys[i]; }, p.NA, want, s_hist, s_bcast, &n_gt, &n_eq);
      const int need_eq = want - n_gt;        // how many keys == T are in (highest indices)
      if (need_eq < n_eq) {                   // (all ties in: T2 = 0 admits them, no second select)
        int dummy = 0;
        T2 = radix_select([&](int i) { return s_keys[i] == T ? (unsigned)(i + 1) : 0u; }, p.NA,
                          need_eq, s_hist, s_bcast, &dummy);
      }
    }
    // compact this batch + bitonic sort (ascending in (~key, ~index))
    if (tid == 0) s_cnt[1] = 0;
    for (int i = tid; i < M; i += PT) s_sort[i] = ~0ull;
    __syncthreads();
    for (int a = tid; a < p.NA; a += PT) {
      const unsigned k = s_keys[a];
      const bool in_now = k > T || (k == T && (unsigned)(a + 1) >= T2);
      const bool in_before = taken > 0 && (k > Tp || (k == Tp && (unsigned)(a + 1) >= T2p));
      if (in_now && !in_before) {
        const int pos = atomicAdd(&s_cnt[1], 1);
        s_sort[pos] = ((unsigned long long)(~k) << 32) | (unsigned)(~(unsigned)a);
      }
    }
    __syncthreads();
    // (element i lives in lane i & 31 of its warp for every i = e * PT + tid: the stages with
    // j < 32 exchange through shuffles, one barrier per run of them instead of one per stage)
    for (int k = 2; k <= M; k <<= 1) {
      int j = k >> 1;
      for (; j >= 32; j >>= 1) {
        for (int t = tid; t < M / 2; t += PT) {
          const int i = 2 * t - (t & (j - 1));
          const bool up = ((i & k) == 0);
          const unsigned long long x = s_sort[i], y = s_sort[i + j];
          if ((x > y) == up) { s_sort[i] = y; s_sort[i + j] = x; }
        }
        __syncthreads();
      }
      for (int i = tid; i - lane < M; i += PT) {             // whole warps (M may be below 32)
        unsigned long long x = i < M ? s_sort[i] : ~0ull;
        const bool up = ((i & k) == 0);
        for (int jj = j; jj > 0; jj >>= 1) {
          const unsigned long long y = __shfl_xor_sync(0xffffffffu, x, jj);
          const bool lower = (i & jj) == 0;                  // I hold the lower index of the pair
          if ((lower == up) ? (x > y) : (x < y)) x = y;      // keep min at the lower index when up
        }
        if (i < M) s_sort[i] = x;
      }
      __syncthreads();
    }
    Tp = T;
    T2p = T2;
    // decode + min-size filter (:123, :151-156), order-preserving compaction in place (an entry
    // only ever moves towards the front, and a block of PT entries is read before it is written)
    int nvalid = 0;
    for (int base = 0; base < want - taken; base += PT) {
      const int i = base + tid;
      unsigned long long e = 0;
      bool ok = false;
      if (i < want - taken) {
        e = s_sort[i];
        const int a = (int)(~(unsigned)(e & 0xffffffffull));
        const float4 bx = decode_anchor(p, img, a, im_h, im_w);
        const float ws = __fadd_rn(__fsub_rn(bx.z, bx.x), 1.0f);
        const float hs = __fadd_rn(__fsub_rn(bx.w, bx.y), 1.0f);
        ok = (ws >= min_size) && (hs >= min_size);
      }
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) s_wcnt[warp] = __popc(bal);
      __syncthreads();
      int before = 0, total = 0;
#pragma unroll
      for (int w = 0; w < PT / 32; ++w) {
        const int c = s_wcnt[w];
        if (w < warp) before += c;
        total += c;
      }
      if (ok) s_sort[nvalid + before + __popc(bal & ((1u << lane) - 1u))] = e;
      nvalid += total;
      __syncthreads();
    }
    const int nbatch = min(nvalid, p.pre_nms_topN - taken_valid);   // :130-131
    taken_valid += nvalid;
  for (int c0 = 0; c0 < nbatch && nkept < post; c0 += CHUNK, ++round) {
    const int nc = min(CHUNK, nbatch - c0);
    if (tid < CHUNK) {
      if (tid < nc) {
        const unsigned long long e = s_sort[c0 + tid];
        const int a = (int)(~(unsigned)(e & 0xffffffffull));
        const float4 b = decode_anchor(p, img, a, im_h, im_w);
        s_cbox[tid] = b;
        s_carea[tid] = __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.0f),
                                 __fadd_rn(__fsub_rn(b.w, b.y), 1.0f));
        s_cidx[tid] = a;
        s_ckey[tid] = ~(unsigned)(e >> 32);
      }
    }
    __syncthreads();
    unsigned long long* const colp = s_col + (CS > 1 ? (round & 1) * CHUNK * 4 : 0);
    unsigned* const alivep = s_alive + (CS > 1 ? (round & 1) * (CHUNK / 32) : 0);
    if constexpr (CS == 1) {
    // A: candidate (tid>>2) against kept entries tid&3, +4, +8, ...
    {
      const int cand = tid >> 2;
      int sup = 0;
      if (cand < nc) {
        const float4 bj = s_cbox[cand];
        const float aj = s_carea[cand];
        for (int k = tid & 3; k < nkept; k += 4) {
          if (iou_ge(s_kbox[k], s_karea[k], bj, aj, p.thr_ge)) { sup = 1; break; }
        }
      }
      sup |= __shfl_xor_sync(0xffffffffu, sup, 1);
      sup |= __shfl_xor_sync(0xffffffffu, sup, 2);
      // lanes 0,4,8,...: 8 candidates per warp -> one byte of the alive bitmap
      const unsigned bal = __ballot_sync(0xffffffffu, !sup && cand < nc);
      if (lane == 0) {
        unsigned byte = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) byte |= ((bal >> (4 * q)) & 1u) << q;
        reinterpret_cast<unsigned char*>(s_alive)[warp] = (unsigned char)byte;
      }
    }
    __syncthreads();
    // B: column masks inside the chunk: bit i of candidate j's 256-bit column = alive i < j
    //    suppresses j.  The four threads of a candidate take the earlier candidates i == g (mod 4):
    //    every lane of a warp then runs about the same number of tests (j / 4 of them; with one
    //    64-candidate block per thread most lanes idled while a few ran 64 tests), and the four
    //    partial columns are OR-ed with two shuffles per word.
    {
      const int j = tid >> 2, g = tid & 3;
      unsigned res[CHUNK / 32];
#pragma unroll
      for (int wd = 0; wd < CHUNK / 32; ++wd) res[wd] = 0u;
      const bool alive_j = (alivep[j >> 5] >> (j & 31)) & 1u;
      if (alive_j) {
        const float4 bj = s_cbox[j];
        const float aj = s_carea[j];
        const unsigned mine = 0x11111111u << g;
#pragma unroll
        for (int wd = 0; wd < CHUNK / 32; ++wd) {
          if (32 * wd < j) {
            unsigned aw = alivep[wd] & mine;
            if (j - 32 * wd < 32) aw &= (1u << (j - 32 * wd)) - 1u;   // earlier candidates only
            while (aw) {
              const int bit = __ffs((int)aw) - 1;
              aw &= aw - 1;
              const int i = 32 * wd + bit;
              if (iou_ge(s_cbox[i], s_carea[i], bj, aj, p.thr_ge)) res[wd] |= 1u << bit;
            }
          }
        }
      }
#pragma unroll
      for (int wd = 0; wd < CHUNK / 32; ++wd) {
        res[wd] |= __shfl_xor_sync(0xffffffffu, res[wd], 1);
        res[wd] |= __shfl_xor_sync(0xffffffffu, res[wd], 2);
      }
      unsigned lo = 0, hi = 0;
#pragma unroll
      for (int wd = 0; wd < CHUNK / 32; wd += 2)
        if ((wd >> 1) == g) { lo = res[wd]; hi = res[wd + 1]; }
      colp[j * 4 + g] = ((unsigned long long)hi << 32) | lo;
    }
    __syncthreads();
    } else {
      // Cluster rounds: CTA r owns the candidates [r*Q, (r+1)*Q) of the chunk, TPC = 4*CS threads
      // each.  A: the kept list, strided over the TPC threads.  B: the candidate's column over
      // ALL earlier candidates i == sub (mod TPC) -- the alive bits of candidates owned by peers
      // are not known yet, and columns are only ever ANDed with kept masks, so testing dead ones
      // costs work, not correctness -- OR-reduced with shuffles.  Then ONE exchange per round:
      // every CTA stores its alive words and its columns into every CTA's copy (distributed
      // shared memory, two slot sets so that a CTA one round ahead cannot overwrite what a peer
      // still reads) and the cluster barrier publishes them; C and D run redundantly.
      constexpr int Q = CHUNK / CS, TPC = PT / Q, AW = (Q + 31) / 32;
      static_assert(TPC <= 32 && Q * TPC == PT && Q >= 32, "4*CS threads per candidate, CS <= 8");
      cg::cluster_group cluster = cg::this_cluster();
      const int jl = tid / TPC, sub = tid % TPC;
      const int j = crank * Q + jl;
      int sup = 0;
      float4 bj = make_float4(0.f, 0.f, 0.f, 0.f);
      float aj = 0.f;
      if (j < nc) {
        bj = s_cbox[j];
        aj = s_carea[j];
        for (int k = sub; k < nkept; k += TPC) {
          if (iou_ge(s_kbox[k], s_karea[k], bj, aj, p.thr_ge)) { sup = 1; break; }
        }
      }
#pragma unroll
      for (int o = 1; o < TPC; o <<= 1) sup |= __shfl_xor_sync(0xffffffffu, sup, o);
      const bool alive_j = !sup && j < nc;
      unsigned res[CHUNK / 32];
#pragma unroll
      for (int wd = 0; wd < CHUNK / 32; ++wd) res[wd] = 0u;
      if (alive_j) {
        // bits b of a word with (32*wd + b) % TPC == sub (32 % TPC == 0)
        const unsigned mine = (TPC == 8 ? 0x01010101u : (TPC == 16 ? 0x00010001u : 1u)) << sub;
#pragma unroll
        for (int wd = 0; wd < CHUNK / 32; ++wd) {
          if (32 * wd < j) {
            unsigned aw = mine;
            if (j - 32 * wd < 32) aw &= (1u << (j - 32 * wd)) - 1u;   // earlier candidates only
            while (aw) {
              const int bit = __ffs((int)aw) - 1;
              aw &= aw - 1;
              const int i = 32 * wd + bit;
              if (iou_ge(s_cbox[i], s_carea[i], bj, aj, p.thr_ge)) res[wd] |= 1u << bit;
            }
          }
        }
      }
#pragma unroll
      for (int wd = 0; wd < CHUNK / 32; ++wd) {
#pragma unroll
        for (int o = 1; o < TPC; o <<= 1) res[wd] |= __shfl_xor_sync(0xffffffffu, res[wd], o);
      }
      // alive bits of my Q candidates: one byte / halfword / word per warp (32 / TPC candidates)
      {
        const unsigned bal = __ballot_sync(0xffffffffu, alive_j && sub == 0);
        if (lane == 0) {
          unsigned bits = 0;
#pragma unroll
          for (int q = 0; q < 32 / TPC; ++q) bits |= ((bal >> (TPC * q)) & 1u) << q;
          // warp w holds candidates jl in [w * 32/TPC, (w+1) * 32/TPC)
          atomicOr(&s_sup[(warp * (32 / TPC)) >> 5], bits << ((warp * (32 / TPC)) & 31));
        }
      }
      const int set = round & 1;
      if (alive_j && sub < CS) {                      // my column into CTA `sub`'s copy
        uint4* remote = reinterpret_cast<uint4*>(cluster.map_shared_rank(s_col, sub) +
                                                 (size_t)set * CHUNK * 4 + (size_t)j * 4);
        remote[0] = make_uint4(res[0], res[1], res[2], res[3]);
        remote[1] = make_uint4(res[4], res[5], res[6], res[7]);
      }
      __syncthreads();                                // s_sup complete
      if (tid < CS * AW) {                            // my alive words into every CTA's copy
        const int peer = tid / AW, w = tid % AW;
        unsigned* remote = cluster.map_shared_rank(s_alive, peer);
        remote[set * (CHUNK / 32) + crank * AW + w] = s_sup[w];
      }
      cluster.sync();
      if (tid < AW) s_sup[tid] = 0u;                  // for the next round (read after 2 barriers)
    }
    // C: warp 0 resolves the chunk, 32 candidates at a time
    if (warp == 0) {
      unsigned kw[CHUNK / 32];
#pragma unroll
      for (int sb = 0; sb < CHUNK / 32; ++sb) kw[sb] = 0;
#pragma unroll
      for (int sb = 0; sb < CHUNK / 32; ++sb) {
        const int j = 32 * sb + lane;
        const bool alive_j = (alivep[sb] >> lane) & 1u;
        const unsigned long long c0w = colp[j * 4 + 0], c1w = colp[j * 4 + 1];
        const unsigned long long c2w = colp[j * 4 + 2], c3w = colp[j * 4 + 3];
        const unsigned colw[8] = {(unsigned)c0w, (unsigned)(c0w >> 32), (unsigned)c1w,
                                  (unsigned)(c1w >> 32), (unsigned)c2w, (unsigned)(c2w >> 32),
                                  (unsigned)c3w, (unsigned)(c3w >> 32)};
        bool dead = !alive_j;
#pragma unroll
        for (int q = 0; q < CHUNK / 32; ++q)
          if (q < sb) dead = dead || ((colw[q] & kw[q]) != 0u);
        unsigned own = 0;
#pragma unroll
        for (int q = 0; q < CHUNK / 32; ++q)
          if (q == sb) own = colw[q];
        unsigned Ks = __ballot_sync(0xffffffffu, !dead);
        for (int it = 0; it < 32; ++it) {
          const unsigned Kn = __ballot_sync(0xffffffffu, !dead && (own & Ks) == 0u);
          if (Kn == Ks) break;
          Ks = Kn;
        }
        kw[sb] = Ks;
      }
      if (lane == 0) {
#pragma unroll
        for (int sb = 0; sb < CHUNK / 32; ++sb) s_kmask[sb] = kw[sb];
      }
    }
    __syncthreads();
    // D: append the newly kept boxes (in order) and emit them
    int total_new = 0;
    {
      int before = 0;   // kept candidates before mine in this chunk
      const int j = tid;
#pragma unroll
      for (int sb = 0; sb < CHUNK / 32; ++sb) {
        const unsigned m = s_kmask[sb];
        total_new += __popc(m);
        if (j < CHUNK) {
          if (sb < (j >> 5)) before += __popc(m);
          else if (sb == (j >> 5)) before += __popc(m & ((1u << (j & 31)) - 1u));
        }
      }
      if (j < CHUNK && ((s_kmask[j >> 5] >> (j & 31)) & 1u)) {
        const int pos = nkept + before;
        if (pos < post) {
          const float4 b = s_cbox[j];
          s_kbox[pos] = b;
          s_karea[pos] = s_carea[j];
          if (writer) {
            float* r = p.rois + ((size_t)img * post + pos) * 5;
            r[0] = (float)img; r[1] = b.x; r[2] = b.y; r[3] = b.z; r[4] = b.w;
            const int a = s_cidx[j];
            if (p.scores) p.scores[(size_t)img * post + pos] = key_to_float(s_ckey[j]);
            if (p.anchor_idx) p.anchor_idx[(size_t)img * post + pos] = a;
          }
        }
      }
    }
    nkept = min(nkept + total_new, post);
    __syncthreads();
  }
  }
  if constexpr (CS > 1) cg::this_cluster().sync();   // no peer may still write into my slots
  if (!writer) return;
  // zero-fill the unused tail so the blob is deterministic
  for (int i = nkept * 5 + tid; i < post * 5; i += PT)
    p.rois[(size_t)img * post * 5 + i] = (i % 5 == 0) ? p.pad_batch : 0.f;
  for (int i = nkept + tid; i < post; i += PT) {
    if (p.scores) p.scores[(size_t)img * post + i] = 0.f;
    if (p.anchor_idx) p.anchor_idx[(size_t)img * post + i] = -1;
  }
  if (tid == 0) p.counts[img] = nkept;
}

size_t prop_smem_bytes(int NA, int kpad, int post) {
  const size_t na_pad = (size_t)((NA + 3) & ~3);
  size_t total = sizeof(unsigned long long) * (size_t)kpad;          // s_sort
  total += sizeof(unsigned) * (256 + 4) + sizeof(int) * 4;           // s_hist, s_bcast, s_cnt
  total += sizeof(unsigned) * na_pad;                                // score keys
  total += sizeof(float4) * CHUNK + sizeof(unsigned long long) * CHUNK * 4 * 2 +
           sizeof(float) * CHUNK + sizeof(int) * CHUNK + sizeof(unsigned) * CHUNK +
           sizeof(unsigned) * (CHUNK / 32) * 3;                      // per-chunk NMS state
  total += (sizeof(float4) + sizeof(float)) * (size_t)post;          // kept list
  return total;
}

int pow2ceil(int v) {
  int p = 2;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace

extern "C" size_t wssdl_proposals_workspace_bytes(int, int, int, int, int, int) {
  return 256;   // the fused kernel keeps all of its state in shared memory
}

extern "C" int wssdl_proposals(const float* cls_prob, const float* bbox_pred,
                               const float* im_info, int info_stride, int B, int H, int W, int A,
                               const float* base_anchors, int feat_stride, int pre_nms_topN,
                               int post_nms_topN, double nms_thresh, int nms_mode, float min_size,
                               float* rois,
                               float* scores, int* anchor_idx, int* counts, float* decoded,
                               void* workspace, size_t workspace_bytes, wssdl_stream_t stream) {
  (void)workspace; (void)workspace_bytes;
  return wssdl_proposals_impl(cls_prob, bbox_pred, im_info, info_stride, B, H, W, A, base_anchors,
                              feat_stride, pre_nms_topN, post_nms_topN, nms_thresh, nms_mode,
                              min_size, rois, scores, anchor_idx, counts, decoded, stream, 0.f, 0);
}

int wssdl_proposals_impl(const float* cls_prob, const float* bbox_pred, const float* im_info,
                         int info_stride, int B, int H, int W, int A, const float* base_anchors,
                         int feat_stride, int pre_nms_topN, int post_nms_topN, double nms_thresh,
                         int nms_mode, float min_size, float* rois, float* scores, int* anchor_idx,
                         int* counts, float* decoded, wssdl_stream_t stream, float pad_batch_index,
                         int one_cta_per_image) {
  if (B < 0 || H <= 0 || W <= 0 || A <= 0 || info_stride < 3) return WSSDL_EINVAL;
  if (nms_mode != WSSDL_NMS_GE_F64 && nms_mode != WSSDL_NMS_GT_F32) return WSSDL_EINVAL;
  if (B == 0) return WSSDL_OK;
  if (!cls_prob || !bbox_pred || !im_info || !base_anchors || !rois || !counts) return WSSDL_EINVAL;
  const long long NA = (long long)H * W * A;
  if (A > MAX_ANCHORS || NA > 32768) return WSSDL_ELIMIT;
  if (pre_nms_topN <= 0) pre_nms_topN = (int)NA;        // :130: no truncation
  if (post_nms_topN <= 0 || post_nms_topN > 4096) return WSSDL_ELIMIT;
  if (decoded && !aligned16(decoded)) return WSSDL_EALIGN;
  // candidates sorted at a time: the first batch should usually hold enough to keep `post`
  const int kmax = pow2ceil((int)((long long)pre_nms_topN < NA ? pre_nms_topN : NA));
  int kpad = pow2ceil(2 * post_nms_topN);
  if (kpad < 1024) kpad = 1024;
  if (kpad > kmax) kpad = kmax;
  // (1 KB of the 227 KB carve-out is left to the kernels' static arrays)
  const size_t smem_max = 226 * 1024;
  while (kpad > 1024 && prop_smem_bytes((int)NA, kpad, post_nms_topN) > smem_max) kpad >>= 1;
  const size_t smem = prop_smem_bytes((int)NA, kpad, post_nms_topN);
  if (smem > smem_max) return WSSDL_ELIMIT;
  cudaStream_t s = to_cuda(stream);
  // Batches that leave SMs idle: a cluster of 2 / 4 / 8 CTAs per image (see the kernel), the
  // largest size whose B clusters are resident at once (cudaOccupancyMaxActiveClusters knows the
  // GPC layout); a batch that fills the machine keeps one CTA per image.
  // WSSDL_TUNE_PROPOSALS_CLUSTER: -1 by shape, 0 one CTA per image, 1 by shape (as -1), 2 / 4 / 8
  // that cluster size.
  auto set_smem = [&](auto kernel) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  };
  auto fits = [&](auto kernel, int cs) {
    if ((long long)B * cs > WSSDL_NUM_SMS) return false;
    if (set_smem(kernel) != cudaSuccess) return false;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * cs));
    cfg.blockDim = dim3(PT);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return false; }
    return n >= B;
  };
  // (one_cta_per_image: the caller overlaps this launch with other kernels and wants the least
  // SM time, not the least latency)
  const int ctune = one_cta_per_image ? 0 : wssdl_tuning(WSSDL_TUNE_PROPOSALS_CLUSTER);
  int cs = 1;
  if (ctune == 2 || ctune == 4 || ctune == 8) {
    cs = ctune;
  } else if (ctune != 0) {
    // (cached per (B, smem): the occupancy query is a driver call)
    static std::mutex c_mu;
    static int c_B = -1, c_cs = 1;
    static size_t c_smem = 0;
    std::lock_guard<std::mutex> lock(c_mu);
    if (c_B == B && c_smem == smem) {
      cs = c_cs;
    } else {
      if (fits(proposals_kernel<8>, 8)) cs = 8;
      else if (fits(proposals_kernel<4>, 4)) cs = 4;
      else if (fits(proposals_kernel<2>, 2)) cs = 2;
      c_B = B; c_smem = smem; c_cs = cs;
    }
  }
  if (cs == 8) WSSDL_RETURN_IF_CUDA(set_smem(proposals_kernel<8>));
  else if (cs == 4) WSSDL_RETURN_IF_CUDA(set_smem(proposals_kernel<4>));
  else if (cs == 2) WSSDL_RETURN_IF_CUDA(set_smem(proposals_kernel<2>));
  else WSSDL_RETURN_IF_CUDA(set_smem(proposals_kernel<1>));
  PropParams p;
  p.cls_prob = cls_prob; p.bbox_pred = bbox_pred; p.im_info = im_info;
  p.info_stride = info_stride; p.H = H; p.W = W; p.A = A; p.NA = (int)NA;
  p.feat_stride = feat_stride; p.pre_nms_topN = pre_nms_topN; p.post_nms_topN = post_nms_topN;
  p.kpad = kpad;
  // cpu_nms: (double)iou >= thresh  <=>  iou >= the smallest float not below thresh;
  // gpu_nms / py_cpu_nms: iou > (float)thresh  <=>  iou >= the next float above (float)thresh
  float f = (float)nms_thresh;
  if (nms_mode == WSSDL_NMS_GT_F32) f = nextafterf(f, INFINITY);
  else if ((double)f < nms_thresh) f = nextafterf(f, INFINITY);
  p.thr_ge = f;
  p.min_size = min_size;
  p.pad_batch = pad_batch_index;
  p.rois = rois; p.scores = scores; p.anchor_idx = anchor_idx; p.counts = counts;
  p.decoded = decoded;
  // base anchors: HOST pointer (generate_anchors runs on the host, as in the reference)
  for (int i = 0; i < MAX_ANCHORS * 4; ++i) p.base[i] = i < 4 * A ? base_anchors[i] : 0.f;
  if (cs > 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * cs));
    cfg.blockDim = dim3(PT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cs == 8) WSSDL_RETURN_IF_CUDA(cudaLaunchKernelEx(&cfg, proposals_kernel<8>, p));
    else if (cs == 4) WSSDL_RETURN_IF_CUDA(cudaLaunchKernelEx(&cfg, proposals_kernel<4>, p));
    else WSSDL_RETURN_IF_CUDA(cudaLaunchKernelEx(&cfg, proposals_kernel<2>, p));
  } else {
    proposals_kernel<1><<<B, PT, smem, s>>>(p);
  }
  WSSDL_CHECK_LAUNCH();
  return WSSDL_OK;
}
