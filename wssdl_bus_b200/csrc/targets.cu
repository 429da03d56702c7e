// Sampled part of anchor_target_layer[_joint] for sm_100a, device resident
// (rpn_msr/anchor_target_layer_tf_bus.py:512-611): fg / bg subsampling, regression targets,
// inside / outside weights, `_unmap` and the final (B,1,A*H,W) / (B,4A,H,W) layouts in ONE
// kernel for the whole batch, fed by the labels kernel (anchor_target.cu) without a host trip.
//
// Subsampling.  The reference disables `npr.choice(fg_inds, size=n_fg - num_fg, replace=False)`
// (:514-518) and the same for bg (:524-528).  A draw of choice(replace=False) depends on the
// SIZE of its population only, so two modes give the same kind of result:
//   WSSDL_SAMPLE_RANKS   the caller drew on the host (same RandomState stream as the reference:
//                        the parity mode) and passes, per image, the RANKS of the fg / bg anchors
//                        to disable -- rank = position in the ascending list of inside anchors
//                        with that label; the only thing it needed from the device were the two
//                        counts per image (wssdl_anchor_label_counts);
//   WSSDL_SAMPLE_PHILOX  no host trip at all: anchor i of image b gets the key
//                        philox4x32-10(counter = (i, b, which, 0), key = seed).x, which = 0 fg /
//                        1 bg, and the k anchors with the smallest (key, i) are disabled.
//
// Targets (:533, :645-653 -> fast_rcnn/bbox_transform.py:10-28) with the reference's dtypes:
// anchors are float64, GT rows float32, so GT widths / centres are fp32 expressions, everything
// that touches an anchor is fp64, np.log runs in fp64, the result is cast to fp32 once.
#include "common.cuh"

namespace {

constexpr int TG_THREADS = 1024;
constexpr int MAX_ANCHORS = 32;

struct TgParams {
  const float* labels_pre;   // [Bs,NA] pre-subsample labels (-1 / 0 / 1; outside anchors -1)
  const int* argmax_gt;      // [Bs,NA] best fg GT row, -1 for anchors outside the image
  const float* gt_boxes;     // [Bs,max_gt,5]
  int max_gt;
  int Bs, B_total;           // supervised images, all images (the rest: weakly supervised)
  int H, W, A, NA;
  int feat_stride;
  int num_fg, batchsize;     // int(RPN_FG_FRACTION * RPN_BATCHSIZE), RPN_BATCHSIZE
  int mode;
  const int* ranks;          // RANKS mode: disable lists, all images back to back
  const int* rank_off;       // [2*Bs+1]: image b: fg ranks [off[2b], off[2b+1]), bg [off[2b+1], off[2b+2])
  unsigned long long seed;   // PHILOX mode
  float inside_w[4];         // RPN_BBOX_INSIDE_WEIGHTS
  double positive_weight;    // RPN_POSITIVE_WEIGHT (< 0: uniform)
  float* labels_out;         // [B_total, A*H*W]      = (B,1,A*H,W)
  float* targets;            // [B_total, 4A*H*W]     = (B,4A,H,W)
  float* inside;             // same
  float* outside;            // same
  int* final_counts;         // [Bs,2] (may be NULL): fg, bg after subsampling
  float base[MAX_ANCHORS * 4];
};

__device__ __forceinline__ unsigned mulhilo(unsigned a, unsigned b, unsigned* hi) {
  const unsigned long long p = (unsigned long long)a * b;
  *hi = (unsigned)(p >> 32);
  return (unsigned)p;
}
// Philox4x32-10 (Salmon et al., SC'11), the standard constants
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    unsigned hi0, hi1;
    const unsigned lo0 = mulhilo(0xD2511F53u, c.x, &hi0);
    const unsigned lo1 = mulhilo(0xCD9E8D57u, c.z, &hi1);
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

__global__ void __launch_bounds__(256)
anchor_label_counts_kernel(const float* __restrict__ labels, int NA, int* __restrict__ counts) {
  __shared__ int s_fg, s_bg;
  if (threadIdx.x == 0) { s_fg = 0; s_bg = 0; }
  __syncthreads();
  int fg = 0, bg = 0;
  const float* l = labels + (size_t)blockIdx.x * NA;
  for (int i = threadIdx.x; i < NA; i += blockDim.x) {
    const float v = l[i];
    fg += v == 1.0f;
    bg += v == 0.0f;
  }
  fg = __reduce_add_sync(0xffffffffu, fg);
  bg = __reduce_add_sync(0xffffffffu, bg);
  if ((threadIdx.x & 31) == 0) { atomicAdd(&s_fg, fg); atomicAdd(&s_bg, bg); }
  __syncthreads();
  if (threadIdx.x == 0) { counts[2 * blockIdx.x] = s_fg; counts[2 * blockIdx.x + 1] = s_bg; }
}

// block-wide sum of one int per thread (all threads get the total)
__device__ __forceinline__ int block_sum(int v, int* s_w) {
  v = __reduce_add_sync(0xffffffffu, v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int w = 0; w < TG_THREADS / 32; ++w) t += s_w[w];
  return t;
}

__global__ void __launch_bounds__(TG_THREADS, 1) anchor_targets_kernel(const TgParams p) {
  extern __shared__ __align__(16) unsigned char tg_smem[];
  __shared__ int s_w[TG_THREADS / 32];
  __shared__ int s_base[2][TG_THREADS / 32 + 1];
  signed char* s_lab = reinterpret_cast<signed char*>(tg_smem);                 // [NA] final labels
  unsigned* s_key = reinterpret_cast<unsigned*>(tg_smem + ((p.NA + 15) & ~15)); // [NA] philox / rank
  unsigned* s_dis = s_key + p.NA;                                               // 2 bitmaps of NA bits
  const int nwords = (p.NA + 31) >> 5;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int img = blockIdx.x;
  const int HW = p.H * p.W;
  float* lab_o = p.labels_out + (size_t)img * p.NA;
  float* tgt_o = p.targets + (size_t)img * 4 * p.NA;
  float* in_o = p.inside + (size_t)img * 4 * p.NA;
  float* out_o = p.outside + (size_t)img * 4 * p.NA;
  if (img >= p.Bs) {
    // weakly supervised image: no RPN supervision (:613-626)
    for (int i = tid; i < p.NA; i += TG_THREADS) lab_o[i] = -1.0f;
    for (int i = tid; i < 4 * p.NA; i += TG_THREADS) { tgt_o[i] = 0.f; in_o[i] = 0.f; out_o[i] = 0.f; }
    return;
  }
  const float* lpre = p.labels_pre + (size_t)img * p.NA;
  const int* amax = p.argmax_gt + (size_t)img * p.NA;

  // ---- labels into shared memory; every thread owns a CONTIGUOUS run of anchors so that ranks
  // (positions in the ascending lists of fg / bg anchors) are running sums
  const int per = (p.NA + TG_THREADS - 1) / TG_THREADS;
  const int a0 = min(tid * per, p.NA), a1 = min(a0 + per, p.NA);
  int my_fg = 0, my_bg = 0;
  for (int i = a0; i < a1; ++i) {
    const float v = lpre[i];
    const int l = amax[i] < 0 ? -1 : (v == 1.0f ? 1 : (v == 0.0f ? 0 : -1));
    s_lab[i] = (signed char)l;
    my_fg += l == 1;
    my_bg += l == 0;
  }
  for (int i = tid; i < 2 * nwords; i += TG_THREADS) s_dis[i] = 0u;
  // exclusive prefix of (my_fg, my_bg) over the threads
  int xf = my_fg, xb = my_bg;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int yf = __shfl_up_sync(0xffffffffu, xf, d), yb = __shfl_up_sync(0xffffffffu, xb, d);
    if (lane >= d) { xf += yf; xb += yb; }
  }
  if (lane == 31) { s_base[0][warp + 1] = xf; s_base[1][warp + 1] = xb; }
  __syncthreads();
  if (tid == 0) {
    s_base[0][0] = 0; s_base[1][0] = 0;
    for (int w = 1; w <= TG_THREADS / 32; ++w) { s_base[0][w] += s_base[0][w - 1]; s_base[1][w] += s_base[1][w - 1]; }
  }
  __syncthreads();
  const int n_fg = s_base[0][TG_THREADS / 32], n_bg = s_base[1][TG_THREADS / 32];
  int rf = s_base[0][warp] + xf - my_fg, rb = s_base[1][warp] + xb - my_bg;   // my first ranks
  // how many to disable (:513-528): fg first, the bg quota depends on the fg that remain
  const int kill_fg = max(n_fg - p.num_fg, 0);
  const int fg_final = n_fg - kill_fg;
  const int num_bg = p.batchsize - fg_final;
  const int kill_bg = max(n_bg - max(num_bg, 0), 0);

  if (p.mode == WSSDL_SAMPLE_RANKS) {
    // the ranks the host drew -> two bitmaps over ranks
    const int o0 = p.rank_off[2 * img], o1 = p.rank_off[2 * img + 1], o2 = p.rank_off[2 * img + 2];
    for (int i = o0 + tid; i < o2; i += TG_THREADS) {
      const int r = p.ranks[i];
      const bool is_fg = i < o1;
      if (r >= 0 && r < (is_fg ? n_fg : n_bg))
        atomicOr(&s_dis[(is_fg ? 0 : nwords) + (r >> 5)], 1u << (r & 31));
    }
    __syncthreads();
    for (int i = a0; i < a1; ++i) {
      const int l = s_lab[i];
      if (l == 1) { if ((s_dis[rf >> 5] >> (rf & 31)) & 1u) s_lab[i] = -1; ++rf; }
      else if (l == 0) { if ((s_dis[nwords + (rb >> 5)] >> (rb & 31)) & 1u) s_lab[i] = -1; ++rb; }
    }
  } else {
    // Philox keys; the k smallest (key, index) composites of each label go: bisection on the
    // 47-bit composite (key << 15 | index, NA <= 32768), one block-wide count per step
    for (int which = 0; which < 2; ++which) {
      const int kill = which == 0 ? kill_fg : kill_bg;
      if (kill <= 0) continue;                                // (uniform over the block)
      const int want = which == 0 ? 1 : 0;
      for (int i = a0; i < a1; ++i)
        s_key[i] = philox4x32_10(make_uint4((unsigned)i, (unsigned)img, (unsigned)which, 0u),
                                 make_uint2((unsigned)p.seed, (unsigned)(p.seed >> 32))).x;
      // largest T with count(composite < T) <= kill  ->  exactly `kill` composites are < T
      unsigned long long lo = 0, hi = 1ull << 48;             // invariant: count(< lo) <= kill; every composite is < 2^47
      while (hi - lo > 1) {
        const unsigned long long mid = lo + ((hi - lo) >> 1);
        int c = 0;
        for (int i = a0; i < a1; ++i)
          c += s_lab[i] == want && (((unsigned long long)s_key[i] << 15) | (unsigned)i) < mid;
        c = block_sum(c, s_w);
        if (c <= kill) lo = mid; else hi = mid;
      }
      for (int i = a0; i < a1; ++i)
        if (s_lab[i] == want && (((unsigned long long)s_key[i] << 15) | (unsigned)i) < lo) s_lab[i] = -1;
      __syncthreads();
    }
  }
  __syncthreads();
  const int bg_final = n_bg - kill_bg;
  if (tid == 0 && p.final_counts) { p.final_counts[2 * img] = fg_final; p.final_counts[2 * img + 1] = bg_final; }
  // outside weights (:540-553), computed in fp64 like numpy and rounded when stored
  float pos_w, neg_w;
  if (p.positive_weight < 0) {
    pos_w = neg_w = (float)(1.0 / (double)(fg_final + bg_final));
  } else {
    pos_w = (float)(p.positive_weight / (double)fg_final);
    neg_w = (float)((1.0 - p.positive_weight) / (double)bg_final);
  }

  // ---- outputs in their final layouts: `_unmap` (fill -1 / 0), reshape (1,H,W,A[*4]),
  // transpose (0,3,1,2) (:568-598): element [a*? + c][h][w] comes from anchor (h*W + w)*A + a
  for (int o = tid; o < p.NA; o += TG_THREADS) {
    const int a = o / HW, cell = o - a * HW;
    lab_o[o] = (float)s_lab[cell * p.A + a];
  }
  const float* gt = p.gt_boxes + (size_t)img * p.max_gt * 5;
  for (int o = tid; o < 4 * p.NA; o += TG_THREADS) {
    const int ch = o / HW, cell = o - ch * HW;
    const int a = ch >> 2, c = ch & 3;
    const int i = cell * p.A + a;
    const int l = s_lab[i];
    const int g = amax[i];
    float t = 0.f;
    if (g >= 0) {                                   // inside anchors: targets for ALL of them (:533)
      const int y = cell / p.W, x = cell - y * p.W;
      const bool is_x = (c & 1) == 0;
      const double sh = (double)((is_x ? x : y) * p.feat_stride);
      const double e1 = (double)p.base[4 * a + (is_x ? 0 : 1)] + sh;
      const double e2 = (double)p.base[4 * a + (is_x ? 2 : 3)] + sh;
      const float g1 = gt[g * 5 + (is_x ? 0 : 1)], g2 = gt[g * 5 + (is_x ? 2 : 3)];
      const double ex_len = __dadd_rn(__dsub_rn(e2, e1), 1.0);               // ex_widths / heights
      const float gt_len = __fadd_rn(__fsub_rn(g2, g1), 1.0f);               // fp32: GT rows are fp32
      if (c < 2) {
        const double ex_ctr = __dadd_rn(e1, __dmul_rn(0.5, ex_len));
        const float gt_ctr = __fadd_rn(g1, __fmul_rn(0.5f, gt_len));
        t = (float)__ddiv_rn(__dsub_rn((double)gt_ctr, ex_ctr), ex_len);     // :22-23
      } else {
        t = (float)log(__ddiv_rn((double)gt_len, ex_len));                   // :24-25
      }
    }
    tgt_o[o] = t;
    in_o[o] = l == 1 ? p.inside_w[c] : 0.f;
    out_o[o] = l == 1 ? pos_w : (l == 0 ? neg_w : 0.f);
  }
}


// ------------------------------------------------------------------------------------------
// proposal_target_layer[_joint], device resident (rpn_msr/proposal_target_layer_tf_bus.py:15-184,
// _sample_rois :228-280).  Two kernels, one CTA per supervised image:
//   roi_match_kernel    the image's candidate RoIs in their original order (+ its fg GT rows when
//                       training, :45-50), fp64 IoU against the fg GT rows exactly as bbox.pyx,
//                       max / first argmax (:233-235), fg / bg candidacy (:239, :252-253) and the
//                       RANK of every candidate among the fg / bg candidates; counts to the caller;
//   roi_targets_kernel  the selected candidates, in selection order, into the output rows: RoI,
//                       label (:264-266), regression targets (fp32 bbox_transform + optional
//                       normalisation, :213-226) expanded to the 4-of-4K layout with inside /
//                       outside weights (:187-210, :84).
// The selection itself -- npr.choice(fg_inds, fg_this, replace=False) then the same for bg (:250,
// :261) -- depends on the candidate COUNTS only: drawn on the host from numpy.random (parity mode)
// or on the device from Philox keys (the k smallest (key, index) in ascending order).
constexpr int PT_THREADS = 1024;
constexpr int MAX_GT_ROWS = 64;

struct RmParams {
  const float* rois;         // [R,5] (batch, x1, y1, x2, y2), any order
  int R;
  const float* gt_boxes;     // [Bs,max_gt,5]
  const int* num_gt;         // [Bs]
  int max_gt;
  int add_gt;                // append the image's fg GT rows to its candidates
  int cap;                   // candidate slots per image in the workspace (>= R_img + max_gt)
  double fg_thresh, bg_hi, bg_lo;
  int* cand;                 // [Bs,cap]  RoI row (>= 0) or -(g+1) for GT row g
  int* assign;               // [Bs,cap]  first argmax over the fg GT rows
  int* fg_list;              // [Bs,cap]  candidate slot of the fg candidate of rank r
  int* bg_list;              // [Bs,cap]  the same for the bg candidates
  int* counts;               // [Bs,4]    candidates, fg candidates, bg candidates, fg GT rows
};

__global__ void __launch_bounds__(PT_THREADS, 1) roi_match_kernel(const RmParams p) {
  __shared__ int s_w[3][PT_THREADS / 32];
  __shared__ double s_gt[MAX_GT_ROWS][4];
  __shared__ int s_npos;
  const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* gt = p.gt_boxes + (size_t)img * p.max_gt * 5;
  if (tid == 0) {                                   // fg rows come first (:38-40): count them
    int np_ = 0;
    const int n = min(p.num_gt[img], p.max_gt);
    for (int g = 0; g < n; ++g) np_ += gt[g * 5 + 4] != 0.f;
    s_npos = np_;
  }
  __syncthreads();
  const int npos = s_npos;
  for (int i = tid; i < npos * 4; i += PT_THREADS) s_gt[i >> 2][i & 3] = (double)gt[(i >> 2) * 5 + (i & 3)];
  int* cand = p.cand + (size_t)img * p.cap;
  int* assign = p.assign + (size_t)img * p.cap;
  int* fg_list = p.fg_list + (size_t)img * p.cap;
  int* bg_list = p.bg_list + (size_t)img * p.cap;
  // ---- candidates: the RoIs of this image in their original order, then the fg GT rows
  int ncand = 0;
  for (int r0 = 0; r0 < p.R; r0 += PT_THREADS) {
    const int r = r0 + tid;
    const bool mine = r < p.R && (int)p.rois[(size_t)r * 5] == img && p.rois[(size_t)r * 5] == (float)img;
    const unsigned bal = __ballot_sync(0xffffffffu, mine);
    if (lane == 0) s_w[0][warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < PT_THREADS / 32; ++w) { const int c = s_w[0][w]; if (w < warp) before += c; total += c; }
    if (mine) {
      const int pos = ncand + before + __popc(bal & ((1u << lane) - 1u));
      if (pos < p.cap) cand[pos] = r;
    }
    ncand += total;
    __syncthreads();
  }
  if (p.add_gt) {
    for (int g = tid; g < npos; g += PT_THREADS) if (ncand + g < p.cap) cand[ncand + g] = -(g + 1);
    ncand += npos;
  }
  ncand = min(ncand, p.cap);
  __syncthreads();
  // ---- IoU (bbox.pyx:15-55, fp64), max / first argmax, candidacy, ranks
  int fg_total = 0, bg_total = 0;
  for (int c0 = 0; c0 < ncand; c0 += PT_THREADS) {
    const int c = c0 + tid;
    bool is_fg = false, is_bg = false;
    if (c < ncand) {
      const int id = cand[c];
      double b[4];
      if (id >= 0) { for (int k = 0; k < 4; ++k) b[k] = (double)p.rois[(size_t)id * 5 + 1 + k]; }
      else { for (int k = 0; k < 4; ++k) b[k] = s_gt[-id - 1][k]; }
      double best = 0.0;
      int arg = 0;
      bool first = true;
      for (int g = 0; g < npos; ++g) {
        double ov = 0.0;
        const double iw = __dadd_rn(__dsub_rn(fmin(b[2], s_gt[g][2]), fmax(b[0], s_gt[g][0])), 1.0);
        if (iw > 0) {
          const double ih = __dadd_rn(__dsub_rn(fmin(b[3], s_gt[g][3]), fmax(b[1], s_gt[g][1])), 1.0);
          if (ih > 0) {
            const double qa = __dmul_rn(__dadd_rn(__dsub_rn(s_gt[g][2], s_gt[g][0]), 1.0),
                                        __dadd_rn(__dsub_rn(s_gt[g][3], s_gt[g][1]), 1.0));
            const double ba = __dmul_rn(__dadd_rn(__dsub_rn(b[2], b[0]), 1.0), __dadd_rn(__dsub_rn(b[3], b[1]), 1.0));
            const double inter = __dmul_rn(iw, ih);
            ov = __ddiv_rn(inter, __dsub_rn(__dadd_rn(ba, qa), inter));
          }
        }
        if (first || ov > best) { best = ov; arg = g; first = false; }   // argmax: first maximum
      }
      assign[c] = arg;
      is_fg = best >= p.fg_thresh;                                       // :239
      is_bg = best < p.bg_hi && best >= p.bg_lo;                         // :252-253
    }
    const unsigned bf = __ballot_sync(0xffffffffu, is_fg), bb = __ballot_sync(0xffffffffu, is_bg);
    if (lane == 0) { s_w[1][warp] = __popc(bf); s_w[2][warp] = __popc(bb); }
    __syncthreads();
    int f_before = 0, f_tot = 0, b_before = 0, b_tot = 0;
    for (int w = 0; w < PT_THREADS / 32; ++w) {
      const int cf = s_w[1][w], cb = s_w[2][w];
      if (w < warp) { f_before += cf; b_before += cb; }
      f_tot += cf; b_tot += cb;
    }
    if (c < ncand) {
      const unsigned lt = (1u << lane) - 1u;
      if (is_fg) fg_list[fg_total + f_before + __popc(bf & lt)] = c;
      if (is_bg) bg_list[bg_total + b_before + __popc(bb & lt)] = c;
    }
    fg_total += f_tot;
    bg_total += b_tot;
    __syncthreads();
  }
  if (tid == 0) {
    p.counts[4 * img] = ncand; p.counts[4 * img + 1] = fg_total;
    p.counts[4 * img + 2] = bg_total; p.counts[4 * img + 3] = npos;
  }
}

struct RtParams {
  const float* rois;
  const float* gt_boxes;
  int max_gt, cap, K;
  const int* cand;
  const int* assign;
  const int* fg_list;
  const int* bg_list;
  const int* counts;         // [Bs,4] from roi_match_kernel
  int mode;
  const int* sel;            // RANKS mode: per image fg ranks in selection order, then bg ranks
  const int* sel_off;        // [2*Bs+1]
  const int* row_off;        // [Bs+1] first output row of every image (RANKS mode)
  int fg_quota, rois_per_image;   // PHILOX mode: fixed stride rois_per_image rows per image
  unsigned long long seed;
  int normalize;
  double means[4], stds[4];
  float inside_w[4];
  float* out_rois;           // [N,5]
  float* out_labels;         // [N]
  float* out_targets;        // [N,4K]
  float* out_inside;         // [N,4K]
  float* out_outside;        // [N,4K]
  int* out_counts;           // [Bs,2] (PHILOX mode): fg, bg rows of every image
};

__global__ void __launch_bounds__(PT_THREADS, 1) roi_targets_kernel(const RtParams p) {
  extern __shared__ __align__(16) unsigned char rt_smem[];
  const int img = blockIdx.x, tid = threadIdx.x;
  const int ncand = p.counts[4 * img], n_fg = p.counts[4 * img + 1], n_bg = p.counts[4 * img + 2];
  int* s_sel = reinterpret_cast<int*>(rt_smem);     // [rois_per_image] (PHILOX) selected slots, in order
  const int* cand = p.cand + (size_t)img * p.cap;
  const int* assign = p.assign + (size_t)img * p.cap;
  const int* fg_of = p.fg_list + (size_t)img * p.cap;   // candidate slot of fg rank r (global memory)
  const int* bg_of = p.bg_list + (size_t)img * p.cap;
  (void)ncand;
  int fg_this, bg_this, row0;
  if (p.mode == WSSDL_SAMPLE_RANKS) {
    fg_this = p.sel_off[2 * img + 1] - p.sel_off[2 * img];
    bg_this = p.sel_off[2 * img + 2] - p.sel_off[2 * img + 1];
    row0 = p.row_off[img];
  } else {
    fg_this = min(p.fg_quota, n_fg);                                    // :243
    bg_this = min(p.rois_per_image - fg_this, n_bg);                    // :256-258
    row0 = img * p.rois_per_image;
    // the k candidates with the smallest (philox key, rank), in ascending order: each candidate
    // counts the candidates of its kind that precede it in that order (n_fg, n_bg <= a few thousand)
    unsigned* s_key = reinterpret_cast<unsigned*>(s_sel + p.rois_per_image);   // [cap]
    const uint2 key = make_uint2((unsigned)p.seed, (unsigned)(p.seed >> 32));
    for (int which = 0; which < 2; ++which) {
      const int n = which == 0 ? n_fg : n_bg, k = which == 0 ? fg_this : bg_this;
      const int base = which == 0 ? 0 : fg_this;
      __syncthreads();
      for (int r = tid; r < n; r += PT_THREADS)
        s_key[r] = philox4x32_10(make_uint4((unsigned)r, (unsigned)img, 2u + which, 0u), key).x;
      __syncthreads();
      for (int r = tid; r < n; r += PT_THREADS) {
        const unsigned mine = s_key[r];
        int before = 0;
        for (int q = 0; q < n; ++q) {
          const unsigned other = s_key[q];
          before += (other < mine) || (other == mine && q < r);
        }
        if (before < k) s_sel[base + before] = which == 0 ? fg_of[r] : bg_of[r];
      }
    }
    __syncthreads();
    if (tid == 0 && p.out_counts) { p.out_counts[2 * img] = fg_this; p.out_counts[2 * img + 1] = bg_this; }
  }
  const int nrows = fg_this + bg_this;
  const float* gt = p.gt_boxes + (size_t)img * p.max_gt * 5;
  const int K4 = 4 * p.K;
  for (int j = tid; j < nrows; j += PT_THREADS) {
    int c;
    if (p.mode == WSSDL_SAMPLE_RANKS) {
      const int rk = p.sel[p.sel_off[2 * img] + j];
      c = j < fg_this ? fg_of[rk] : bg_of[rk];
    } else {
      c = s_sel[j];
    }
    const int id = cand[c];
    float roi[5];
    if (id >= 0) { for (int k = 0; k < 5; ++k) roi[k] = p.rois[(size_t)id * 5 + k]; }
    else { roi[0] = (float)img; for (int k = 0; k < 4; ++k) roi[1 + k] = gt[(-id - 1) * 5 + k]; }
    const float* g = gt + assign[c] * 5;
    const float label = j < fg_this ? g[4] : 0.f;                       // :264-266
    float* o = p.out_rois + (size_t)(row0 + j) * 5;
    for (int k = 0; k < 5; ++k) o[k] = roi[k];
    p.out_labels[row0 + j] = label;
    // bbox_transform in fp32 (all operands are float32 arrays, :213-219), log via fp64
    const float ew = __fadd_rn(__fsub_rn(roi[3], roi[1]), 1.0f), eh = __fadd_rn(__fsub_rn(roi[4], roi[2]), 1.0f);
    const float ecx = __fadd_rn(roi[1], __fmul_rn(0.5f, ew)), ecy = __fadd_rn(roi[2], __fmul_rn(0.5f, eh));
    const float gw = __fadd_rn(__fsub_rn(g[2], g[0]), 1.0f), gh = __fadd_rn(__fsub_rn(g[3], g[1]), 1.0f);
    const float gcx = __fadd_rn(g[0], __fmul_rn(0.5f, gw)), gcy = __fadd_rn(g[1], __fmul_rn(0.5f, gh));
    float t[4] = {__fdiv_rn(__fsub_rn(gcx, ecx), ew), __fdiv_rn(__fsub_rn(gcy, ecy), eh),
                  (float)log((double)__fdiv_rn(gw, ew)), (float)log((double)__fdiv_rn(gh, eh))};
    if (p.normalize)                                                     // :221-224, fp64 like numpy
      for (int k = 0; k < 4; ++k) t[k] = (float)__ddiv_rn(__dsub_rn((double)t[k], p.means[k]), p.stds[k]);
    float* ot = p.out_targets + (size_t)(row0 + j) * K4;
    float* oi = p.out_inside + (size_t)(row0 + j) * K4;
    float* oo = p.out_outside + (size_t)(row0 + j) * K4;
    const int cls = (int)label;
    for (int k = 0; k < K4; ++k) {
      const bool on = label > 0.f && cls < p.K && (k >> 2) == cls;      // :200-208
      ot[k] = on ? t[k & 3] : 0.f;
      oi[k] = on ? p.inside_w[k & 3] : 0.f;
      oo[k] = (on && p.inside_w[k & 3] > 0.f) ? 1.f : 0.f;              // :84
    }
  }
  if (p.mode == WSSDL_SAMPLE_PHILOX) {                                  // fixed stride: clear the rest
    for (int j = nrows + tid; j < p.rois_per_image; j += PT_THREADS) {
      float* o = p.out_rois + (size_t)(row0 + j) * 5;
      for (int k = 0; k < 5; ++k) o[k] = 0.f;
      p.out_labels[row0 + j] = 0.f;
      for (int k = 0; k < K4; ++k) {
        p.out_targets[(size_t)(row0 + j) * K4 + k] = 0.f;
        p.out_inside[(size_t)(row0 + j) * K4 + k] = 0.f;
        p.out_outside[(size_t)(row0 + j) * K4 + k] = 0.f;
      }
    }
  }
}

}  // namespace

extern "C" int wssdl_anchor_label_counts(const float* labels, int B, int NA, int* counts,
                                         wssdl_stream_t stream) {
  if (B < 0 || NA < 0) return WSSDL_EINVAL;
  if (B == 0) return WSSDL_OK;
  if (!labels || !counts) return WSSDL_EINVAL;
  anchor_label_counts_kernel<<<B, 256, 0, to_cuda(stream)>>>(labels, NA, counts);
  WSSDL_CHECK_LAUNCH();
  return WSSDL_OK;
}

extern "C" int wssdl_anchor_targets(const float* labels_pre, const int* argmax_gt,
                                    const float* gt_boxes, int max_gt, int B_supervised,
                                    int B_total, int H, int W, int A, const float* base_anchors,
                                    int feat_stride, int num_fg, int batchsize, int sample_mode,
                                    const int* ranks, const int* rank_off,
                                    unsigned long long seed, const float* inside_weights,
                                    double positive_weight, float* labels_out, float* targets,
                                    float* inside, float* outside, int* final_counts,
                                    wssdl_stream_t stream) {
  if (B_supervised < 0 || B_total < B_supervised || H <= 0 || W <= 0 || A <= 0 || max_gt <= 0)
    return WSSDL_EINVAL;
  if (sample_mode != WSSDL_SAMPLE_RANKS && sample_mode != WSSDL_SAMPLE_PHILOX) return WSSDL_EINVAL;
  if (B_total == 0) return WSSDL_OK;
  if (!base_anchors || !inside_weights || !labels_out || !targets || !inside || !outside)
    return WSSDL_EINVAL;
  if (B_supervised > 0 && (!labels_pre || !argmax_gt || !gt_boxes)) return WSSDL_EINVAL;
  if (sample_mode == WSSDL_SAMPLE_RANKS && B_supervised > 0 && !rank_off) return WSSDL_EINVAL;
  const long long NA = (long long)H * W * A;
  if (A > MAX_ANCHORS || NA > 32768) return WSSDL_ELIMIT;
  TgParams p;
  p.labels_pre = labels_pre; p.argmax_gt = argmax_gt; p.gt_boxes = gt_boxes; p.max_gt = max_gt;
  p.Bs = B_supervised; p.B_total = B_total; p.H = H; p.W = W; p.A = A; p.NA = (int)NA;
  p.feat_stride = feat_stride; p.num_fg = num_fg; p.batchsize = batchsize; p.mode = sample_mode;
  p.ranks = ranks; p.rank_off = rank_off; p.seed = seed;
  for (int i = 0; i < 4; ++i) p.inside_w[i] = inside_weights[i];
  p.positive_weight = positive_weight;
  p.labels_out = labels_out; p.targets = targets; p.inside = inside; p.outside = outside;
  p.final_counts = final_counts;
  for (int i = 0; i < MAX_ANCHORS * 4; ++i) p.base[i] = i < 4 * A ? base_anchors[i] : 0.f;
  const size_t smem = (size_t)(((int)NA + 15) & ~15) + sizeof(unsigned) * (size_t)NA +
                      sizeof(unsigned) * 2 * (size_t)(((int)NA + 31) >> 5);
  if (smem > 48 * 1024)
    WSSDL_RETURN_IF_CUDA(cudaFuncSetAttribute(anchor_targets_kernel,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  anchor_targets_kernel<<<B_total, TG_THREADS, smem, to_cuda(stream)>>>(p);
  WSSDL_CHECK_LAUNCH();
  return WSSDL_OK;
}

extern "C" size_t wssdl_roi_targets_workspace_bytes(int B_supervised, int R, int max_gt) {
  if (B_supervised <= 0 || R < 0 || max_gt < 0) return 256;
  const size_t cap = (size_t)R + (size_t)max_gt;
  return sizeof(int) * ((size_t)B_supervised * cap * 4 + (size_t)B_supervised * 4) + 256;
}

// workspace: cand | assign | fg_list | bg_list, [Bs,cap] each
extern "C" int wssdl_roi_match(const float* rois, int R, const float* gt_boxes, const int* num_gt,
                               int max_gt, int B_supervised, int add_gt, double fg_thresh,
                               double bg_thresh_hi, double bg_thresh_lo, void* workspace,
                               size_t workspace_bytes, int* counts, wssdl_stream_t stream) {
  if (R < 0 || B_supervised < 0 || max_gt <= 0) return WSSDL_EINVAL;
  if (B_supervised == 0) return WSSDL_OK;
  if (!gt_boxes || !num_gt || !workspace || !counts || (R > 0 && !rois)) return WSSDL_EINVAL;
  if (max_gt > MAX_GT_ROWS) return WSSDL_ELIMIT;
  if (workspace_bytes < wssdl_roi_targets_workspace_bytes(B_supervised, R, max_gt)) return WSSDL_EWORKSPACE;
  RmParams p;
  p.rois = rois; p.R = R; p.gt_boxes = gt_boxes; p.num_gt = num_gt; p.max_gt = max_gt;
  p.add_gt = add_gt ? 1 : 0;
  p.cap = R + max_gt;
  p.fg_thresh = fg_thresh; p.bg_hi = bg_thresh_hi; p.bg_lo = bg_thresh_lo;
  int* w = static_cast<int*>(workspace);
  p.cand = w; p.assign = w + (size_t)B_supervised * p.cap;
  p.fg_list = p.assign + (size_t)B_supervised * p.cap;
  p.bg_list = p.fg_list + (size_t)B_supervised * p.cap;
  p.counts = counts;
  roi_match_kernel<<<B_supervised, PT_THREADS, 0, to_cuda(stream)>>>(p);
  WSSDL_CHECK_LAUNCH();
  return WSSDL_OK;
}

extern "C" int wssdl_roi_targets(const float* rois, int R, const float* gt_boxes, int max_gt,
                                 int B_supervised, int num_classes, const void* workspace,
                                 const int* counts, int sample_mode, const int* sel,
                                 const int* sel_off, const int* row_off, int fg_rois_per_image,
                                 int rois_per_image, unsigned long long seed,
                                 const double* normalize_means, const double* normalize_stds,
                                 const float* inside_weights, float* out_rois, float* out_labels,
                                 float* out_targets, float* out_inside, float* out_outside,
                                 int* out_counts, wssdl_stream_t stream) {
  if (R < 0 || B_supervised < 0 || max_gt <= 0 || num_classes <= 0) return WSSDL_EINVAL;
  if (sample_mode != WSSDL_SAMPLE_RANKS && sample_mode != WSSDL_SAMPLE_PHILOX) return WSSDL_EINVAL;
  if (B_supervised == 0) return WSSDL_OK;
  if (!gt_boxes || !workspace || !counts || !inside_weights || !out_rois || !out_labels ||
      !out_targets || !out_inside || !out_outside)
    return WSSDL_EINVAL;
  if (sample_mode == WSSDL_SAMPLE_RANKS && (!sel_off || !row_off)) return WSSDL_EINVAL;
  if (sample_mode == WSSDL_SAMPLE_PHILOX && (rois_per_image <= 0 || fg_rois_per_image < 0)) return WSSDL_EINVAL;
  RtParams p;
  p.rois = rois; p.gt_boxes = gt_boxes; p.max_gt = max_gt; p.cap = R + max_gt; p.K = num_classes;
  const int* w = static_cast<const int*>(workspace);
  p.cand = w; p.assign = w + (size_t)B_supervised * p.cap;
  p.fg_list = p.assign + (size_t)B_supervised * p.cap;
  p.bg_list = p.fg_list + (size_t)B_supervised * p.cap;
  p.counts = counts;
  p.mode = sample_mode; p.sel = sel; p.sel_off = sel_off; p.row_off = row_off;
  p.fg_quota = fg_rois_per_image; p.rois_per_image = rois_per_image; p.seed = seed;
  p.normalize = (normalize_means && normalize_stds) ? 1 : 0;
  for (int k = 0; k < 4; ++k) {
    p.means[k] = p.normalize ? normalize_means[k] : 0.0;
    p.stds[k] = p.normalize ? normalize_stds[k] : 1.0;
    p.inside_w[k] = inside_weights[k];
  }
  p.out_rois = out_rois; p.out_labels = out_labels; p.out_targets = out_targets;
  p.out_inside = out_inside; p.out_outside = out_outside; p.out_counts = out_counts;
  const size_t smem = sample_mode == WSSDL_SAMPLE_PHILOX
                          ? sizeof(int) * ((size_t)rois_per_image + p.cap) : 16;
  if (smem > 200 * 1024) return WSSDL_ELIMIT;
  if (smem > 48 * 1024)
    WSSDL_RETURN_IF_CUDA(cudaFuncSetAttribute(roi_targets_kernel,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  roi_targets_kernel<<<B_supervised, PT_THREADS, smem, to_cuda(stream)>>>(p);
  WSSDL_CHECK_LAUNCH();
  return WSSDL_OK;
}
