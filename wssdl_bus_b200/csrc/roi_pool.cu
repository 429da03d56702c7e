// RoI max pooling, NHWC, forward (+argmax) and backward, for sm_100a.
//
// Semantics: RoiPoolOp / RoiPoolGradOp of the reference
// (roi_pooling_layer/roi_pooling_op.cc:137-196, :383-458); CUDA twin
// (roi_pooling_op_gpu.cu.cc:19-85) selectable as bin_mode GPU_CEIL.
//
// Forward: HBM-write bound by the roofline (8 B written per output element against 4 B read per
// map element), three kernels with identical results, picked by shape in choose_fwd():
//   band   (roi_pool_fwd_band_kernel)  32-channel slice of overlapping row bands of the map in
//          shared memory: conflict-free 128 B cells, full-line stores.  The detector's shapes.
//   direct (roi_pool_fwd_kernel)       one CTA per (roi, ph) output row reading the L2-resident
//          map.  Big bins, grids that the band kernel would split into a few ragged waves.
//   tiled  (roi_pool_fwd_tiled_kernel) 16-channel slice of the whole map in shared memory.
//          C % 32 != 0.
// Common to all: threads run along C (channels, not cells, are the parallel axis in NHWC), each
// thread scans its bin h-major / w-minor with strict '>' exactly like roi_pooling_op.cc:184-191,
// so ties (post-ReLU zeros) resolve to the first cell with no cross-lane reduction; outputs are
// never re-read and leave as streaming stores.
//
// Direct kernel design:
//   - one CTA per (roi, ph) output row; threads run along C in float4 lanes, so every
//     feature-map cell is read as one fully coalesced 16 B/lane segment and every output
//     row (PW*C floats + PW*C ints) is written as contiguous 128-bit streaming stores;
//   - the RoI geometry (4 rounds, 2 IEEE divisions) is computed once per thread and
//     reused for all PW bins instead of once per output element as in the reference;
//   - CPU_TRUNC bins never overlap, so a cell is read once per RoI and nothing is reused
//     inside one RoI (the reuse the shared-memory kernels capture is ACROSS the RoIs of an
//     image); loads go through the read-only path (ld.global.nc) and the outputs use
//     st.global.cs so they do not evict the feature map from L2.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "roi_pool_dev.cuh"

using namespace wssdl_roi;

namespace {

template <int VEC> struct VecT;
template <> struct VecT<4> { using F = float4; using I = int4; };
template <> struct VecT<1> { using F = float;  using I = int; };

__device__ __forceinline__ void upd(float v, int cell, float& m, int& mi) {
  // strict '>' (cc:187): NaN never wins, first maximum wins
  if (v > m) { m = v; mi = cell; }
}
__device__ __forceinline__ void upd4(const float4 v, int cell, float (&m)[4], int (&mi)[4]) {
  upd(v.x, cell, m[0], mi[0]); upd(v.y, cell, m[1], mi[1]);
  upd(v.z, cell, m[2], mi[2]); upd(v.w, cell, m[3], mi[3]);
}

// The kernel is issue-slot bound before it is HBM bound (ncu, profiles/r01_*: the first
// version ran at 82 % issue utilisation and 48 % DRAM), so the loops are written for
// instruction count:
//   - no channel loop: one thread = one channel vector (grid.y covers C > 4*blockDim.x);
//   - one 64-bit running pointer per row, constant strides folded into immediates when C is a
//     compile-time value (CVT = C/4 for the common C = 512 / 1024, 0 = runtime);
//   - the scan tracks cell*C (+= C per cell) and each channel's index register starts at
//     -1 - c, so the flat index (h*W+w)*C + c is a single add per output and "nothing
//     pooled" comes out as -1 without a select;
//   - 3 instructions (FSETP, FSEL, SEL) per channel and cell, which is the floor for a
//     first-maximum argmax.
//   - VPT channel vectors per thread (lanes cv and cv + CV/VPT): the per-bin / per-row
//     bookkeeping is paid once for 4*VPT channels.
template <int VEC, int BIN_MODE, int CVT, int VPT>
__global__ void __launch_bounds__(256)
roi_pool_fwd_kernel(const float* __restrict__ bottom, const float* __restrict__ rois,
                    int B, int H, int W, int C_rt, int PH, int PW, float spatial_scale,
                    float* __restrict__ top, int* __restrict__ argmax) {
  using F = typename VecT<VEC>::F;
  using I = typename VecT<VEC>::I;
  const int C = CVT ? CVT * VEC : C_rt;
  const int CV = CVT ? CVT : C_rt / VEC;           // vector lanes per cell
  const int CVH = CV / VPT;                        // lanes owned by "vector slot" 0
  const int cv = blockIdx.y * blockDim.x + threadIdx.x;
  if (cv >= CVH) return;
  const int n = blockIdx.x / PH;
  const int ph = blockIdx.x - n * PH;

  const RoiCells g = roi_cells(rois + (size_t)n * 5, spatial_scale, PH, PW);
  const bool bad_batch = (g.batch < 0) || (g.batch >= B);

  int hstart = bin_lo<BIN_MODE>(ph, g.bin_h);
  int hend = bin_hi<BIN_MODE>(ph, g.bin_h);
  hstart = min(max(hstart + g.start_h, 0), H);   // cc:173-176
  hend = min(max(hend + g.start_h, 0), H);
  const bool row_empty = (hend <= hstart) || bad_batch;
  const int nh = hend - hstart;

  // this thread's lane in row hstart, column 0 of its image
  const F* __restrict__ img = reinterpret_cast<const F*>(bottom) +
      ((size_t)(bad_batch ? 0 : g.batch) * H + (row_empty ? 0 : hstart)) * (size_t)(W * CV) + cv;
  const int row_stride = W * CV;                   // in vectors
  const int c = cv * VEC;
  const size_t out0 = (((size_t)n * PH + ph) * PW + threadIdx.y) * CV + cv;   // vector units
  F* __restrict__ top_p = reinterpret_cast<F*>(top) + out0;
  I* __restrict__ arg_p = reinterpret_cast<I*>(argmax) + out0;
  const int out_step = blockDim.y * CV;

  for (int pw = threadIdx.y; pw < PW; pw += blockDim.y, top_p += out_step, arg_p += out_step) {
    int wstart = bin_lo<BIN_MODE>(pw, g.bin_w);
    int wend = bin_hi<BIN_MODE>(pw, g.bin_w);
    wstart = min(max(wstart + g.start_w, 0), W);
    wend = min(max(wend + g.start_w, 0), W);
    const bool is_empty = row_empty || (wend <= wstart);
    const int nw = wend - wstart;

    float m[VPT][4];
    int mi[VPT][4];
#pragma unroll
    for (int v = 0; v < VPT; ++v)
#pragma unroll
      for (int k = 0; k < 4; ++k) {                // cc:180-182
        m[v][k] = is_empty ? 0.f : -FLT_MAX;
        mi[v][k] = -1 - (c + v * CVH * VEC) - k;   // + channel at the end => -1 if never updated
      }
    if (!is_empty) {
      const F* __restrict__ prow = img + wstart * CV;
      int cellC0 = (hstart * W + wstart) * C;      // (h*W+w)*C of the row start
      for (int r = nh; r > 0; --r, prow += row_stride, cellC0 += W * C) {
        const F* __restrict__ p = prow;
        int cellC = cellC0;
        int left = nw;
        // two cells per trip: all loads are issued before the dependent compares
        for (; left >= 2; left -= 2, p += 2 * CV, cellC += 2 * C) {
          F v0[VPT], v1[VPT];
#pragma unroll
          for (int v = 0; v < VPT; ++v) { v0[v] = __ldg(p + v * CVH); v1[v] = __ldg(p + CV + v * CVH); }
#pragma unroll
          for (int v = 0; v < VPT; ++v) {
            if constexpr (VEC == 4) {
              upd4(v0[v], cellC, m[v], mi[v]);
              upd4(v1[v], cellC + C, m[v], mi[v]);
            } else {
              upd(v0[v], cellC, m[v][0], mi[v][0]);
              upd(v1[v], cellC + C, m[v][0], mi[v][0]);
            }
          }
        }
        if (left) {
#pragma unroll
          for (int v = 0; v < VPT; ++v) {
            const F v0 = __ldg(p + v * CVH);
            if constexpr (VEC == 4) upd4(v0, cellC, m[v], mi[v]);
            else upd(v0, cellC, m[v][0], mi[v][0]);
          }
        }
      }
    }
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const int cc = c + v * CVH * VEC;
      if constexpr (VEC == 4) {
        __stcs(top_p + v * CVH, make_float4(m[v][0], m[v][1], m[v][2], m[v][3]));
        if (argmax != nullptr)
          __stcs(arg_p + v * CVH, make_int4(mi[v][0] + cc, mi[v][1] + cc + 1, mi[v][2] + cc + 2,
                                            mi[v][3] + cc + 3));
      } else {
        __stcs(top_p + v * CVH, m[v][0]);
        if (argmax != nullptr) __stcs(arg_p + v * CVH, mi[v][0] + cc);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Forward, shared-memory resident channel slice ("tiled").
//
// The direct kernel above re-reads every RoI's cells from L2: with 300 overlapping RoIs on
// a 38x50 map that is ~35x the map per image (ncu, profiles/history/r01_roi_pool_fwd_v2: 34 GB of
// L2->SM reads for 15.4 GB of output, the kernel stalls on long_scoreboard at 57 % of the
// DRAM peak).  Here one CTA owns (image, 16-channel slice): the slice of the whole map
// (H*W*64 B = 121.6 KB for 38x50) is staged ONCE into shared memory with cp.async, then
// every RoI of that image is pooled out of shared memory.  L2 sees the map once per
// (image, RoI chunk) and otherwise only the output stream, so the kernel is bound by the
// HBM write stream as the roofline says it should be.
//   - bins average only ~4.5 cells, so per-bin bookkeeping, not the compare/select work,
//     is what costs issue slots: a work item is (RoI, pw, 8 channels) and its thread walks
//     ph = 0..PH-1 down that column of bins, paying the decode once per PH bins;
//   - two threads per bin column: 32 B of each 64 B cell chunk per thread (2 x LDS.128),
//     outputs leave as full 32 B sectors (the neighbouring slice CTAs, launched back to
//     back, complete the 128 B lines in L2 before they are evicted);
//   - bin edges of a batch of RoIs are computed once per CTA with the reference's float
//     expressions (roi_cells / bin_lo / bin_hi) and kept packed in shared memory;
//   - warps draw 32 items at a time from a shared counter (RoI sizes vary widely);
//   - tried and dropped: routing the outputs through shared-memory staging and
//     cp.async.bulk shared->global copies (64 B per bin column and tensor) to take the stores
//     off the LSU data pipe -- bit-exact but 6.3 ms instead of 3.6 ms on the C4 workload: 64 B
//     transfers are TMA-issue bound (about one per 3.7 cycles per SM) and the warps spin in
//     wait_group.read (profiles/history/r01_roi_fwd_direct_vs_tiled.txt);
//   - RoIs are grouped by image either by an in-CTA scan of the batch column (R <=
//     T_SCAN_MAX_R, no workspace) or by a counting-sort pre-pass (launch_roi_bucket) into
//     the caller's workspace.  Bucket B collects RoIs whose batch index is outside [0,B):
//     they produce (0,-1) without touching the map.


constexpr int T_SLICE = 16;             // channels per CTA
constexpr int T_LANES = T_SLICE / 4;    // float4 lanes per cell
constexpr int T_THREADS = 1024;
constexpr int T_SCAN_MAX_R = 4096;      // in-CTA RoI list capacity (scan mode)
constexpr int T_MAX_RB = 1024;          // RoIs whose bin edges are resident at a time

// Counting sort of RoI indices by image: img_start[B+2] (exclusive offsets, bucket B collects
// the RoIs with no valid image), perm[R] (RoI indices grouped by bucket; the order inside a
// bucket is not specified and no result depends on it).  Three small stream operations
// instead of one CTA walking all R RoIs (the single-CTA version took 61 us for the bench's
// 76 800 RoIs, ncu launch list profiles/history): zero the counters, histogram (warp-aggregated
// atomics; the last CTA to finish turns the counts into offsets), scatter.
// Lanes of `act` that share bucket b: returns the group mask; *rank = this lane's rank in it.
__device__ __forceinline__ unsigned warp_bucket_group(unsigned act, int b, int* rank) {
  const unsigned grp = __match_any_sync(act, b);
  *rank = __popc(grp & ((1u << (threadIdx.x & 31)) - 1u));
  return grp;
}

__global__ void __launch_bounds__(256)
roi_hist_kernel(const float* __restrict__ rois, int R, int B, int* __restrict__ counts,
                int* __restrict__ ticket, int* __restrict__ img_start, int* __restrict__ cursor) {
  const int lane = threadIdx.x & 31;
  // warp-uniform trip count: every lane reaches the ballot
  for (int r0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); r0 < R; r0 += gridDim.x * blockDim.x) {
    const int r = r0 + lane;
    const unsigned act = __ballot_sync(0xffffffffu, r < R);
    if (r < R) {
      const int b = roi_bucket(__ldg(rois + (size_t)r * 5), B);
      int rank;
      const unsigned grp = warp_bucket_group(act, b, &rank);
      if (rank == 0) atomicAdd(counts + b, __popc(grp));
    }
  }
  // last CTA done: exclusive scan of counts[0..B] -> img_start, cursor
  __shared__ int s_last;
  __shared__ int s_warp[8];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int warp = threadIdx.x >> 5;
  int carry = 0;
  for (int i0 = 0; i0 <= B; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    const int v = (i <= B) ? __ldcg(counts + i) : 0;
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    int woff = 0, total = 0;
    for (int k = 0; k < 8; ++k) {
      const int t = s_warp[k];
      if (k < warp) woff += t;
      total += t;
    }
    const int excl = carry + woff + x - v;
    if (i <= B) { img_start[i] = excl; cursor[i] = excl; }
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) img_start[B + 1] = carry;   // == R
}

__global__ void __launch_bounds__(256)
roi_scatter_kernel(const float* __restrict__ rois, int R, int B, int* __restrict__ cursor,
                   int* __restrict__ perm) {
  const int lane = threadIdx.x & 31;
  for (int r0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); r0 < R; r0 += gridDim.x * blockDim.x) {
    const int r = r0 + lane;
    const unsigned act = __ballot_sync(0xffffffffu, r < R);
    if (r < R) {
      const int b = roi_bucket(__ldg(rois + (size_t)r * 5), B);
      int rank;
      const unsigned grp = warp_bucket_group(act, b, &rank);
      int base = 0;
      if (rank == 0) base = atomicAdd(cursor + b, __popc(grp));
      base = __shfl_sync(grp, base, __ffs(grp) - 1);
      perm[base + rank] = r;
    }
  }
}

constexpr int T_CLASSES = 64;           // RoI cost classes: min(rows/bin,7)*8 + min(cells/row,7)

template <int BIN_MODE, bool STREAM_ST>
__global__ void __launch_bounds__(T_THREADS, 1)
roi_pool_fwd_tiled_kernel(const float* __restrict__ bottom, const float* __restrict__ rois,
                          const int* __restrict__ perm, const int* __restrict__ img_start,
                          int B, int H, int W, int C, int R, int PH, int PW,
                          float spatial_scale, int RB, FastDiv divPW, const Ones ones_p,
                          float* __restrict__ top, int* __restrict__ argmax) {
  extern __shared__ __align__(16) unsigned char t_smem[];
  __shared__ int s_count, s_next;
  __shared__ int s_hist[T_CLASSES], s_start[T_CLASSES];
  const int tid = threadIdx.x;
  const int slice = blockIdx.x, chunk = blockIdx.y, nchunks = gridDim.y;
  const int img = blockIdx.z;                       // == B: RoIs with no valid image
  float4* s_map = reinterpret_cast<float4*>(t_smem);
  unsigned* s_he = reinterpret_cast<unsigned*>(s_map + (size_t)H * W * T_LANES);  // hs | he<<16
  unsigned* s_we = s_he + RB * PH;                                                // ws | we<<16
  int* s_n = reinterpret_cast<int*>(s_we + RB * PW);     // RoI index in the caller's array
  int* s_cls = s_n + RB;                                 // cost class
  int* s_order = s_cls + RB;                             // resident RoIs, heaviest class first
  int* s_list = s_order + RB;                            // scan mode only

  // ---- this CTA's RoIs
  const int* list;
  int r_begin, r_end;
  if (perm != nullptr) {
    const int a = img_start[img];
    const int n_img = img_start[img + 1] - a;
    list = perm + a;
    r_begin = (int)((long long)n_img * chunk / nchunks);
    r_end = (int)((long long)n_img * (chunk + 1) / nchunks);
  } else {
    // the chunk of a RoI is a function of its index alone, so every slice CTA of this
    // (image, chunk) collects the same set whatever order the atomics resolve in
    if (tid == 0) s_count = 0;
    __syncthreads();
    for (int r = tid; r < R; r += blockDim.x)
      if (r % nchunks == chunk && roi_bucket(__ldg(rois + (size_t)r * 5), B) == img)
        s_list[atomicAdd(&s_count, 1)] = r;
    __syncthreads();
    list = s_list;
    r_begin = 0;
    r_end = s_count;
  }
  if (r_end <= r_begin) return;

  // ---- stage the channel slice of the whole map (asynchronously; the first batch of bin
  // edges is computed while it lands)
  const bool valid_img = img < B;
  if (valid_img) {
    const int CV = C >> 2;
    const float4* src = reinterpret_cast<const float4*>(bottom + (size_t)img * H * W * C) +
                        slice * T_LANES;
    const int n4 = H * W * T_LANES;
    for (int i = tid; i < n4; i += blockDim.x)
      cp_async16(s_map + i, src + (size_t)(i >> 2) * CV + (i & 3));
  }
  bool staged = false;

  // The unit operands of upd_fma must live in ordinary registers: as kernel-parameter
  // (constant bank) operands ptxas re-loads them with four LDCUs in every trip of the cell
  // loop, and a plain copy is folded away again.  A round trip through shared memory is
  // opaque to it.
  __shared__ Ones s_ones;
  if (tid == 0) s_ones = ones_p;
  __syncthreads();
  Ones ones;
  ones.f = *reinterpret_cast<volatile float*>(&s_ones.f);
#pragma unroll
  for (int k = 0; k < 8; ++k) ones.i[k] = *reinterpret_cast<volatile int*>(&s_ones.i[k]);
  const int lane32 = tid & 31;
  const int WC = W * C;
  const size_t ph_stride = (size_t)PW * C;          // output elements between ph and ph+1

  for (int r0 = r_begin; r0 < r_end; r0 += RB) {
    const int nb = min(RB, r_end - r0);
    __syncthreads();                                // previous batch fully consumed
    if (tid == 0) s_next = 0;
    if (tid < T_CLASSES) s_hist[tid] = 0;
    __syncthreads();
    for (int rl = tid; rl < nb; rl += blockDim.x) {
      const int n = list[r0 + rl];
      s_n[rl] = n;
      const RoiCells g = roi_cells(rois + (size_t)n * 5, spatial_scale, PH, PW);
      int max_nh = 0, max_nw = 0;
      for (int ph = 0; ph < PH; ++ph) {
        int hs = min(max(bin_lo<BIN_MODE>(ph, g.bin_h) + g.start_h, 0), H);   // cc:167-176
        int he = min(max(bin_hi<BIN_MODE>(ph, g.bin_h) + g.start_h, 0), H);
        if (!valid_img) hs = he = 0;
        s_he[rl * PH + ph] = (unsigned)hs | ((unsigned)he << 16);
        max_nh = max(max_nh, he - hs);
      }
      for (int pw = 0; pw < PW; ++pw) {
        int ws = min(max(bin_lo<BIN_MODE>(pw, g.bin_w) + g.start_w, 0), W);
        int we = min(max(bin_hi<BIN_MODE>(pw, g.bin_w) + g.start_w, 0), W);
        if (!valid_img) ws = we = 0;
        s_we[rl * PW + pw] = (unsigned)ws | ((unsigned)we << 16);
        max_nw = max(max_nw, we - ws);
      }
      // cost class: lanes of a warp run in lock step over ph, so RoIs that share a warp
      // should have the same rows-per-bin and cells-per-row
      const int cls = min(max_nh, 7) * 8 + min(max_nw, 7);
      s_cls[rl] = cls;
      atomicAdd(&s_hist[cls], 1);
    }
    if (!staged) { cp_async_wait_all(); staged = true; }
    __syncthreads();
    // counting sort by class, heaviest first (they are handed out first)
    if (tid < T_CLASSES) {
      int before = 0;
      for (int k = tid + 1; k < T_CLASSES; ++k) before += s_hist[k];
      s_start[tid] = before;
    }
    __syncthreads();
    for (int rl = tid; rl < nb; rl += blockDim.x) s_order[atomicAdd(&s_start[s_cls[rl]], 1)] = rl;
    __syncthreads();

    // One work item = (RoI, pw, 8-channel half of the slice): the thread walks ph = 0..PH-1
    // down its column, so the item decode is paid once per PH bins.  Warps draw 32 items
    // at a time from a shared counter.
    const int items = nb * PW * 2;
    for (;;) {
      int base = 0;
      if (lane32 == 0) base = atomicAdd(&s_next, 32);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base >= items) break;
      const int it = base + lane32;
      if (it >= items) continue;
      const int half = it & 1;            // channels [8*half, 8*half+8) of the slice
      const int col = it >> 1;
      const int rs = (int)fastdiv((unsigned)col, divPW);
      const int pw = col - rs * PW;
      const int rl = s_order[rs];
      const unsigned ww = s_we[rl * PW + pw];
      const int ws = ww & 0xffff, nw = (int)(ww >> 16) - ws;
      // The 4 bin columns of a quarter warp mostly sit an even number of cells apart, i.e.
      // in the same 16 banks.  Odd columns therefore read their two 16 B chunks in swapped
      // order (2.3 -> 1.5 wavefronts per LDS.128 on proposal RoIs); registers m[0..3] then
      // hold the UPPER four channels and the roles are swapped back at the store.
      const int sw = col & 1;
      const int c = slice * T_SLICE + half * 8;
      const int c_a = c + 4 * sw, c_b = c + 4 * (1 - sw);   // channels of m[0..3] / m[4..7]
      const unsigned* he_p = s_he + rl * PH;
      size_t o = ((size_t)s_n[rl] * PH * PW + pw) * C + c;
      const float4* pcol = s_map + ws * T_LANES + half * 2;
      const int off_a = sw, off_b = 1 - sw;
      const int cellW = ws * C;
#pragma unroll 1
      for (int ph = 0; ph < PH; ++ph, o += ph_stride) {
        const unsigned hh = he_p[ph];
        const int hs = hh & 0xffff, nh = (int)(hh >> 16) - hs;
        const bool is_empty = (nh <= 0) || (nw <= 0);
        float m[8];
        int mi[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {               // cc:180-182
          m[k] = is_empty ? 0.f : -FLT_MAX;
          mi[k] = -1 - (k < 4 ? c_a + k : c_b + k - 4);   // + channel at the end => -1 if never updated
        }
        if (!is_empty) {
          const float4* prow = pcol + hs * W * T_LANES;
          int cellC0 = hs * WC + cellW;
#pragma unroll 1
          for (int r = nh; r > 0; --r, prow += W * T_LANES, cellC0 += WC) {
            const float4* p = prow;
            int cellC = cellC0;
#define WSSDL_UPD8(VA, VB, CELL)                                  \
  do {                                                            \
    upd_fma(VA.x, CELL, m[0], mi[0], ones.f, ones.i[0]);          \
    upd_fma(VA.y, CELL, m[1], mi[1], ones.f, ones.i[1]);          \
    upd_fma(VA.z, CELL, m[2], mi[2], ones.f, ones.i[2]);          \
    upd_fma(VA.w, CELL, m[3], mi[3], ones.f, ones.i[3]);          \
    upd_fma(VB.x, CELL, m[4], mi[4], ones.f, ones.i[4]);          \
    upd_fma(VB.y, CELL, m[5], mi[5], ones.f, ones.i[5]);          \
    upd_fma(VB.z, CELL, m[6], mi[6], ones.f, ones.i[6]);          \
    upd_fma(VB.w, CELL, m[7], mi[7], ones.f, ones.i[7]);          \
  } while (0)
#pragma unroll 1   // bins are 1-3 cells wide: unrolling only adds branches (measured)
            for (int left = nw; left > 0; --left, p += T_LANES, cellC += C) {
              const float4 v0 = p[off_a], v1 = p[off_b];
              WSSDL_UPD8(v0, v1, cellC);
            }
#undef WSSDL_UPD8
          }
        }
        const float4 ta = make_float4(m[0], m[1], m[2], m[3]);
        const float4 tb = make_float4(m[4], m[5], m[6], m[7]);
        const int4 aa = make_int4(mi[0] + c_a, mi[1] + c_a + 1, mi[2] + c_a + 2, mi[3] + c_a + 3);
        const int4 ab = make_int4(mi[4] + c_b, mi[5] + c_b + 1, mi[6] + c_b + 2, mi[7] + c_b + 3);
        st256<STREAM_ST>(top + o, sw ? tb : ta, sw ? ta : tb);
        if (argmax != nullptr) st256<STREAM_ST>(argmax + o, sw ? ab : aa, sw ? aa : ab);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Forward, shared-memory resident 32-channel slice of a BAND of map rows ("band").
//
// What bounded the tiled kernel above (ncu, profiles/r01_ncu_roi_fwd_tiled.txt: LSU data pipe
// 81 %, issue 75 %) follows from its 64 B cells: the four bin columns of a quarter warp land
// in the same banks 1.5x per LDS.128 and a bin column leaves as a 64 B half line, i.e. one
// LSU wavefront per 64 B.  A 32-channel slice makes a cell exactly one 128 B bank row:
//   - four lanes share a bin column, 32 B (8 channels) each; a quarter warp holds two
//     columns A and B; A reads chunks (2j, 2j+1) and B reads (2j+1, 2j) of their cells, so
//     each LDS.128 phase touches all eight 16 B bank groups once WHATEVER the two cells
//     are: conflict-free by construction, 128 B per wavefront;
//   - the four lanes of a column store 4 x 32 B = one full 128 B line per STG.256 wavefront,
//     half the store wavefronts of the 64 B layout.  The A/B register roles differ, so the
//     store is the one divergent statement (two half-populated STG.256, same wavefronts)
//     instead of 16 selects per bin.
// 38x50 cells x 128 B = 243 KB does not fit 227 KB, so the map is cut into NB bands of Hb
// rows that overlap by `ov` rows (ov >= the tallest bin a RoI inside the map can have): band
// b holds rows [b*step, b*step + Hb), step = Hb - ov, and OWNS the bins whose first row
// falls into [b*step, (b+1)*step) (the last band: all the rest).  An owned bin of height
// <= ov lies inside the band; taller ones (RoIs far larger than the map) take a slow path
// that reads global memory.  Every bin is written by exactly one CTA; RoIs that own no bin
// in a band are dropped from that CTA's list.  For 38x50, 7x7: 2 bands of 23 rows (147 KB),
// the map is staged 1.21x.
constexpr int B_SLICE = 32;             // channels per CTA: one cell = 128 B = all 32 banks
constexpr int B_LANES = B_SLICE / 4;    // float4 chunks per cell
constexpr int B_THREADS = 1024;
constexpr int B_SCAN_MAX_R = 4096;
constexpr int B_MAX_RB = 1024;

template <int BIN_MODE, bool HAS_ARGMAX, bool LINEAR>
__global__ void __launch_bounds__(B_THREADS, 1)
roi_pool_fwd_band_kernel(const float* __restrict__ bottom, const float* __restrict__ rois,
                         const int* __restrict__ perm, const int* __restrict__ img_start,
                         int B, int H, int W, int C, int R, int PH, int PW,
                         float spatial_scale, int RB, int nchunks, BandGeom bg, FastDiv divPW,
                         const Ones ones_p, float* __restrict__ top, int* __restrict__ argmax) {
  extern __shared__ __align__(128) unsigned char b_smem[];
  __shared__ int s_count, s_next, s_active;
  __shared__ int s_hist[T_CLASSES], s_start[T_CLASSES];
  __shared__ Ones s_ones;
  const int tid = threadIdx.x;
  const int slice = blockIdx.x;
  const int band = blockIdx.y % bg.NB, chunk = blockIdx.y / bg.NB;
  const int img = blockIdx.z;                       // == B: RoIs with no valid image
  const bool valid_img = img < B;
  if (!valid_img && band != 0) return;              // their (empty) bins all start at row 0
  const int row0 = band * bg.step;                  // first resident row
  const int row1 = min(row0 + bg.Hb, H);            // one past the last resident row
  const int nrows = row1 - row0;

  // cells must be 128 B aligned (the xor-16 chunk pairing and the conflict-free phases rely
  // on it); the launch reserves 128 spare bytes in case the dynamic segment is not
  const unsigned raw_u32 = (unsigned)__cvta_generic_to_shared(b_smem);
  float4* s_map = reinterpret_cast<float4*>(b_smem + ((128u - (raw_u32 & 127u)) & 127u));
  unsigned* s_he = reinterpret_cast<unsigned*>(s_map + (size_t)bg.Hb * W * B_LANES);  // hs | he<<16
  unsigned* s_we = s_he + RB * PH;                                                   // ws | we<<16
  int* s_n = reinterpret_cast<int*>(s_we + RB * PW);     // RoI index in the caller's array
  int* s_pr = s_n + RB;                                  // owned bins: ph_lo | ph_hi<<8 | class<<16
  int* s_order = s_pr + RB;                              // RoIs with work here, heaviest class first
  int* s_list = s_order + RB;                            // scan mode only

  // ---- this CTA's RoIs (as in the tiled kernel)
  const int* list;
  int r_begin, r_end;
  if (perm != nullptr) {
    const int a = img_start[img];
    const int n_img = img_start[img + 1] - a;
    list = perm + a;
    r_begin = (int)((long long)n_img * chunk / nchunks);
    r_end = (int)((long long)n_img * (chunk + 1) / nchunks);
  } else {
    if (tid == 0) s_count = 0;
    __syncthreads();
    for (int r = tid; r < R; r += blockDim.x)
      if (r % nchunks == chunk && roi_bucket(__ldg(rois + (size_t)r * 5), B) == img)
        s_list[atomicAdd(&s_count, 1)] = r;
    __syncthreads();
    list = s_list;
    r_begin = 0;
    r_end = s_count;
  }
  if (r_end <= r_begin) return;

  // ---- stage this band's rows of the channel slice: 128 contiguous bytes per cell
  if (valid_img) {
    const int CV = C >> 2;
    const float4* src = reinterpret_cast<const float4*>(bottom + ((size_t)img * H + row0) * W * C) +
                        slice * B_LANES;
    const int n4 = nrows * W * B_LANES;
    for (int i = tid; i < n4; i += blockDim.x)
      cp_async16(s_map + i, src + (size_t)(i >> 3) * CV + (i & 7));
  }
  bool staged = false;

  if (tid == 0) s_ones = ones_p;                    // see the tiled kernel: opaque unit operands
  __syncthreads();
  Ones ones;
  ones.f = *reinterpret_cast<volatile float*>(&s_ones.f);
#pragma unroll
  for (int k = 0; k < 8; ++k) ones.i[k] = *reinterpret_cast<volatile int*>(&s_ones.i[k]);
  const int lane32 = tid & 31;
  const int WC = W * C;
  const size_t ph_stride = (size_t)PW * C;          // output elements between ph and ph+1
  const float* img_base = bottom + (size_t)(valid_img ? img : 0) * H * WC;
  // per-thread constants: a warp always holds 8 whole columns, so the quarter j and the
  // column role sw of a lane never change
  const int j = lane32 & 3;                         // channels [8j, 8j+8) of the slice
  const int sw = (lane32 >> 2) & 1;                 // column A (0) or B (1) of its quarter warp
  const int c_thr = slice * B_SLICE + j * 8;
  // registers m[0..3] hold chunk 2j+sw, m[4..7] chunk 2j+1-sw
  const int c_a = c_thr + 4 * sw, c_b = c_thr + 4 * (1 - sw);
  const unsigned row_bytes = (unsigned)W * 128u;    // one resident row
  const unsigned map_u32 = (unsigned)__cvta_generic_to_shared(s_map);
  // shared address of chunk A of cell (h, w): map_a + h*row_bytes + w*128
  const unsigned map_a = map_u32 + (unsigned)(2 * j + sw) * 16u - (unsigned)row0 * row_bytes;
  // LINEAR: flat index of the cell whose chunk A sits at q is q*kmul + lin0 (kmul = C/128)
  const int lin0 = -(int)(map_a * (unsigned)(C >> 7));

  for (int r0 = r_begin; r0 < r_end; r0 += RB) {
    const int nb = min(RB, r_end - r0);
    __syncthreads();                                // previous batch fully consumed
    if (tid == 0) { s_next = 0; s_active = 0; }
    if (tid < T_CLASSES) s_hist[tid] = 0;
    __syncthreads();
    for (int rl = tid; rl < nb; rl += blockDim.x) {
      const int n = list[r0 + rl];
      s_n[rl] = n;
      const RoiCells g = roi_cells(rois + (size_t)n * 5, spatial_scale, PH, PW);
      int max_nw = 0;
      for (int pw = 0; pw < PW; ++pw) {
        int ws = min(max(bin_lo<BIN_MODE>(pw, g.bin_w) + g.start_w, 0), W);   // cc:167-176
        int we = min(max(bin_hi<BIN_MODE>(pw, g.bin_w) + g.start_w, 0), W);
        if (!valid_img) ws = we = 0;
        s_we[rl * PW + pw] = (unsigned)ws | ((unsigned)we << 16);
        max_nw = max(max_nw, we - ws);
      }
      int max_nh = 0, ph_lo = PH, ph_hi = 0;
      for (int ph = 0; ph < PH; ++ph) {
        int hs = min(max(bin_lo<BIN_MODE>(ph, g.bin_h) + g.start_h, 0), H);
        int he = min(max(bin_hi<BIN_MODE>(ph, g.bin_h) + g.start_h, 0), H);
        if (!valid_img) hs = he = 0;
        s_he[rl * PH + ph] = (unsigned)hs | ((unsigned)he << 16);
        // hs is non-decreasing in ph, so the bins a band owns are a contiguous range
        const int owner = min(hs / bg.step, bg.NB - 1);
        if (owner == band) {
          ph_lo = min(ph_lo, ph);
          ph_hi = ph + 1;
          max_nh = max(max_nh, he - hs);
        }
      }
      // cost class: lanes of a warp run in lock step over ph, so RoIs that share a warp
      // should have the same rows-per-bin and cells-per-row
      const int cls = min(max_nh, 7) * 8 + min(max_nw, 7);
      s_pr[rl] = ph_lo | (ph_hi << 8) | (cls << 16);
      if (ph_hi > ph_lo) atomicAdd(&s_hist[cls], 1);
    }
    if (!staged) { cp_async_wait_all(); staged = true; }
    __syncthreads();
    if (tid < T_CLASSES) {                          // counting sort by class, heaviest first
      int before = 0;
      for (int k = tid + 1; k < T_CLASSES; ++k) before += s_hist[k];
      s_start[tid] = before;
      if (tid == 0) s_active = before + s_hist[0];
    }
    __syncthreads();
    for (int rl = tid; rl < nb; rl += blockDim.x) {
      const int pr = s_pr[rl];
      if (((pr >> 8) & 255) > (pr & 255)) s_order[atomicAdd(&s_start[pr >> 16], 1)] = rl;
    }
    __syncthreads();

    // One work item = (RoI, pw, quarter of the slice): the thread walks the owned ph range
    // down its bin column.  Warps draw 32 items (8 columns) at a time from a shared counter.
    // The loop nest is written for instruction count (ncu source counters of the first
    // version: 33 instructions per cell, 16 per bin row, 94 per bin):
    //   - shared memory is addressed with 32-bit window addresses (ld.shared), the chunk of
    //     column B is the chunk of column A xor 16;
    //   - LINEAR (C % 128 == 0): a cell's flat index is linear in its shared-memory address,
    //     flat = q * (C/128) + const, so the conditional move records q * (C/128) (`ones.i`
    //     then hold C/128) and no cell counter is carried through the loops;
    //   - rows advance one pointer by a per-item skip, outputs by two running pointers.
    const int items = s_active * PW * 4;
    for (;;) {
      int base = 0;
      if (lane32 == 0) base = atomicAdd(&s_next, 32);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base >= items) break;
      const int it = base + lane32;
      if (it >= items) continue;
      const int col = it >> 2;
      const int rs = (int)fastdiv((unsigned)col, divPW);
      const int pw = col - rs * PW;
      const int rl = s_order[rs];
      const unsigned ww = s_we[rl * PW + pw];
      const int ws = ww & 0xffff, nw = (int)(ww >> 16) - ws;
      const int pr = s_pr[rl];
      const int ph_lo = pr & 255, ph_hi = (pr >> 8) & 255;
      const unsigned* he_p = s_he + rl * PH + ph_lo;
      const unsigned* he_end = s_he + rl * PH + ph_hi;
      const size_t o = (((size_t)s_n[rl] * PH + ph_lo) * PW + pw) * C + c_thr;
      float* top_p = top + o;
      int* arg_p = argmax + (HAS_ARGMAX ? o : 0);
      const unsigned col_a = map_a + (unsigned)ws * 128u;     // chunk A of (row0, ws)
      const unsigned nw_bytes = (unsigned)nw * 128u;
      const unsigned row_skip = row_bytes - nw_bytes;
      // LINEAR: flat = q*kmul + lin0 + channel; else flat = cellC + channel
      const int fin_a = (LINEAR ? lin0 : 0) + c_a, fin_b = (LINEAR ? lin0 : 0) + c_b;
#pragma unroll 1
      for (; he_p != he_end; ++he_p, top_p += ph_stride, arg_p += ph_stride) {
        const unsigned hh = *he_p;
        const int hs = hh & 0xffff, he = (int)(hh >> 16);
        if (he <= hs || nw <= 0) {                  // empty bin: (0, -1), cc:180-182
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          const int4 n1 = make_int4(-1, -1, -1, -1);
          st256<true>(top_p, z, z);
          if (HAS_ARGMAX) st256<true>(arg_p, n1, n1);
          continue;
        }
        if (he > row1) {                            // not resident: slow path (uniform in j)
          band_slow_bin<HAS_ARGMAX>(img_base, hs, he, ws, nw, W, C, c_thr, top_p, arg_p);
          continue;
        }
        float m[8];
        int mi[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          m[k] = -FLT_MAX;
          mi[k] = -1 - (k < 4 ? fin_a + k : fin_b + k - 4);   // + fin at the end => -1 if never updated
        }
        unsigned q = col_a + (unsigned)hs * row_bytes;
        int cellC = LINEAR ? 0 : hs * WC + ws * C;
        int r = he - hs;
        // Tried and dropped (B200, C4 workload, this loop = 3.49 ms): a flattened cell loop
        // with the next cell's loads issued ahead, two register buffers at 768 threads (short
        // scoreboard stalls 2.6 -> 1.5 per issue, but 2.87 G instead of 2.67 G instructions and
        // 6 instead of 8 warps per scheduler: 3.62 ms); 16 channels per thread at 512 threads
        // (4 warps per scheduler cannot hide the latency, 16 columns per warp diverge more:
        // 5.47 ms).
#pragma unroll 1
        do {
          const unsigned q_end = q + nw_bytes;
#pragma unroll 1
          do {
            float4 v0, v1;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(v0.x), "=f"(v0.y), "=f"(v0.z), "=f"(v0.w) : "r"(q));
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(v1.x), "=f"(v1.y), "=f"(v1.z), "=f"(v1.w) : "r"(q ^ 16u));
            const int key = LINEAR ? (int)q : cellC;
            upd_fma(v0.x, key, m[0], mi[0], ones.f, ones.i[0]);
            upd_fma(v0.y, key, m[1], mi[1], ones.f, ones.i[1]);
            upd_fma(v0.z, key, m[2], mi[2], ones.f, ones.i[2]);
            upd_fma(v0.w, key, m[3], mi[3], ones.f, ones.i[3]);
            upd_fma(v1.x, key, m[4], mi[4], ones.f, ones.i[4]);
            upd_fma(v1.y, key, m[5], mi[5], ones.f, ones.i[5]);
            upd_fma(v1.z, key, m[6], mi[6], ones.f, ones.i[6]);
            upd_fma(v1.w, key, m[7], mi[7], ones.f, ones.i[7]);
            q += 128u;
            if (!LINEAR) cellC += C;
          } while (q != q_end);
          q += row_skip;
          if (!LINEAR) cellC += WC - nw * C;
        } while (--r > 0);
        const float4 ta = make_float4(m[0], m[1], m[2], m[3]);
        const float4 tb = make_float4(m[4], m[5], m[6], m[7]);
        const int4 aa = make_int4(mi[0] + fin_a, mi[1] + fin_a + 1, mi[2] + fin_a + 2, mi[3] + fin_a + 3);
        const int4 ab = make_int4(mi[4] + fin_b, mi[5] + fin_b + 1, mi[6] + fin_b + 2, mi[7] + fin_b + 3);
        if (sw) {
          st256<true>(top_p, tb, ta);
          if (HAS_ARGMAX) st256<true>(arg_p, ab, aa);
        } else {
          st256<true>(top_p, ta, tb);
          if (HAS_ARGMAX) st256<true>(arg_p, aa, ab);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Backward, atomic scatter.  One CTA per (roi, ph-slice).  The reference's gather
// (roi_pooling_op.cc:400-455) adds top_diff[n,ph,pw,c] to bottom_diff[b,h,w,c] iff
//   b == roi batch, (h,w) inside the rounded RoI (:415-419), ph in [phstart(h),phend(h)),
//   pw in [pwstart(w),pwend(w)) (:437-445) and argmax[n,ph,pw,c] == (h*W+w)*C+c (:449).
// The feasible h-interval of every ph and w-interval of every pw are evaluated once per
// CTA with the reference's own float expressions (they are monotone in h / w, so each
// set is an interval) and kept in shared memory; the per-element test is then four
// integer compares.  That makes the scatter equal to the gather for arbitrary argmax
// input, including malformed RoIs whose forward argmax is valid but whose in-RoI test
// fails.
template <int VEC>
__global__ void __launch_bounds__(256)
roi_pool_bwd_atomic_kernel(const float* __restrict__ top_diff, const int* __restrict__ argmax,
                           const float* __restrict__ rois, int B, int H, int W, int C, int PH,
                           int PW, float spatial_scale, FastDiv divC, FastDiv divW,
                           FastDiv divCV, FastDiv divPW, float* __restrict__ bottom_diff) {
  using F = typename VecT<VEC>::F;
  using I = typename VecT<VEC>::I;
  extern __shared__ int s_iv[];  // hlo[PH] hhi[PH] wlo[PW] whi[PW]
  int* hlo = s_iv;
  int* hhi = hlo + PH;
  int* wlo = hhi + PH;
  int* whi = wlo + PW;

  const int n = blockIdx.x;
  const int tid = threadIdx.x;
  const RoiCells g = roi_cells(rois + (size_t)n * 5, spatial_scale, PH, PW);
  if (g.batch < 0 || g.batch >= B) return;  // never matches any image (cc:405)

  for (int i = tid; i < PH; i += blockDim.x) { hlo[i] = INT_MAX; hhi[i] = -1; }
  for (int i = tid; i < PW; i += blockDim.x) { wlo[i] = INT_MAX; whi[i] = -1; }
  __syncthreads();
  // in-RoI cells only (cc:415-416), intersected with the map
  const int h0 = max(g.start_h, 0), h1 = min(g.end_h, H - 1);
  const int w0 = max(g.start_w, 0), w1 = min(g.end_w, W - 1);
  const int nh = max(h1 - h0 + 1, 0), nw = max(w1 - w0 + 1, 0);
  for (int i = tid; i < nh + nw; i += blockDim.x) {
    const bool is_h = i < nh;
    const int x = is_h ? (h0 + i) : (w0 + i - nh);
    const int rel = x - (is_h ? g.start_h : g.start_w);
    const float bin = is_h ? g.bin_h : g.bin_w;
    const int P = is_h ? PH : PW;
    int ps = (int)floorf(__fdiv_rn((float)rel, bin));        // cc:437,439
    int pe = (int)ceilf(__fdiv_rn((float)(rel + 1), bin));   // cc:438,440
    ps = min(max(ps, 0), P);
    pe = min(max(pe, 0), P);
    int* lo = is_h ? hlo : wlo;
    int* hi = is_h ? hhi : whi;
    for (int p = ps; p < pe; ++p) { atomicMin(lo + p, x); atomicMax(hi + p, x); }
  }
  __syncthreads();

  const int CV = C / VEC;
  const int HWC = H * W * C;
  float* __restrict__ img = bottom_diff + (size_t)g.batch * HWC;
  const int ph_per = (PH + gridDim.y - 1) / gridDim.y;
  const int ph_begin = blockIdx.y * ph_per;
  const int ph_end = min(ph_begin + ph_per, PH);
  const int items = (ph_end - ph_begin) * PW * CV;
  const size_t base = ((size_t)n * PH + ph_begin) * PW * C;
  for (int it = tid; it < items; it += blockDim.x) {
    // it = ((ph - ph_begin) * PW + pw) * CV + cv
    const int bin = (int)fastdiv((unsigned)it, divCV);
    const int cv = it - bin * CV;
    const int phr = (int)fastdiv((unsigned)bin, divPW);
    const int pw = bin - phr * PW;
    const int ph = ph_begin + phr;
    const int lo_h = hlo[ph], hi_h = hhi[ph], lo_w = wlo[pw], hi_w = whi[pw];
    if (hi_h < lo_h || hi_w < lo_w) continue;
    const size_t o = base + (size_t)it * VEC;
    I a = __ldcs(reinterpret_cast<const I*>(argmax + o));
    F d = __ldcs(reinterpret_cast<const F*>(top_diff + o));
    int av[VEC]; float dv[VEC];
    if constexpr (VEC == 4) {
      av[0] = a.x; av[1] = a.y; av[2] = a.z; av[3] = a.w;
      dv[0] = d.x; dv[1] = d.y; dv[2] = d.z; dv[3] = d.w;
    } else { av[0] = a; dv[0] = d; }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int idx = av[k];
      if (idx < 0 || idx >= HWC) continue;
      const unsigned cell = fastdiv((unsigned)idx, divC);
      if ((int)(idx - cell * C) != cv * VEC + k) continue;     // cc:449 compares with this c
      const unsigned h = fastdiv(cell, divW);
      const int w = (int)(cell - h * W);
      if ((int)h < lo_h || (int)h > hi_h || w < lo_w || w > hi_w) continue;
      atomicAdd(img + idx, dv[k]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Backward, deterministic gather: one CTA per input cell (b,h,w), threads along C.
// RoIs are scanned in ascending order in chunks of blockDim; a ballot-ordered compaction
// keeps the matching ones in order, then every thread accumulates its channels over
// (roi, ph, pw) ascending -- the reference's order, hence bit-exact.
template <int VEC>
__global__ void __launch_bounds__(256)
roi_pool_bwd_gather_kernel(const float* __restrict__ top_diff, const int* __restrict__ argmax,
                           const float* __restrict__ rois, int B, int H, int W, int C, int R,
                           int PH, int PW, float spatial_scale,
                           float* __restrict__ bottom_diff) {
  using F = typename VecT<VEC>::F;
  using I = typename VecT<VEC>::I;
  __shared__ int s_roi[256];
  __shared__ int s_box[256];     // phstart | phend<<8 | pwstart<<16 | pwend<<24
  __shared__ int s_warp_cnt[8];
  __shared__ int s_total;

  const int cell = blockIdx.x;                // (b*H + h)*W + w
  const int w = cell % W;
  const int h = (cell / W) % H;
  const int b = cell / (W * H);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = blockDim.x >> 5;
  const int CV = C / VEC;
  const int base_index = (h * W + w) * C;

  // each thread owns channels cv = tid, tid + blockDim, ... (at most 4 groups => C <= 4096)
  constexpr int MAXG = 4;
  float acc[MAXG][VEC];
#pragma unroll
  for (int gI = 0; gI < MAXG; ++gI)
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[gI][k] = 0.f;

  for (int r0 = 0; r0 < R; r0 += blockDim.x) {
    const int r = r0 + tid;
    bool hit = false;
    int packed = 0;
    if (r < R) {
      const RoiCells g = roi_cells(rois + (size_t)r * 5, spatial_scale, PH, PW);
      hit = (g.batch == b) && (w >= g.start_w && w <= g.end_w && h >= g.start_h && h <= g.end_h);
      if (hit) {
        int phs = (int)floorf(__fdiv_rn((float)(h - g.start_h), g.bin_h));
        int phe = (int)ceilf(__fdiv_rn((float)(h - g.start_h + 1), g.bin_h));
        int pws = (int)floorf(__fdiv_rn((float)(w - g.start_w), g.bin_w));
        int pwe = (int)ceilf(__fdiv_rn((float)(w - g.start_w + 1), g.bin_w));
        phs = min(max(phs, 0), PH); phe = min(max(phe, 0), PH);
        pws = min(max(pws, 0), PW); pwe = min(max(pwe, 0), PW);
        packed = phs | (phe << 8) | (pws << 16) | (pwe << 24);
        hit = (phe > phs) && (pwe > pws);
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) s_warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int off = 0;
    for (int i = 0; i < warp; ++i) off += s_warp_cnt[i];
    if (hit) {
      const int pos = off + __popc(bal & ((1u << lane) - 1u));
      s_roi[pos] = r;
      s_box[pos] = packed;
    }
    if (tid == 0) {
      int t = 0;
      for (int i = 0; i < nwarps; ++i) t += s_warp_cnt[i];
      s_total = t;
    }
    __syncthreads();
    const int total = s_total;
    for (int i = 0; i < total; ++i) {
      const int rr = s_roi[i];
      const int pk = s_box[i];
      const int phs = pk & 255, phe = (pk >> 8) & 255, pws = (pk >> 16) & 255, pwe = (pk >> 24) & 255;
      for (int ph = phs; ph < phe; ++ph) {
        for (int pw = pws; pw < pwe; ++pw) {
          const size_t o = (((size_t)rr * PH + ph) * PW + pw) * C;
#pragma unroll
          for (int gI = 0; gI < MAXG; ++gI) {
            const int cv = tid + gI * blockDim.x;
            if (cv < CV) {
              const I a = __ldg(reinterpret_cast<const I*>(argmax + o + cv * VEC));
              const F d = __ldg(reinterpret_cast<const F*>(top_diff + o + cv * VEC));
              const int want = base_index + cv * VEC;
              if constexpr (VEC == 4) {
                if (a.x == want) acc[gI][0] = __fadd_rn(acc[gI][0], d.x);
                if (a.y == want + 1) acc[gI][1] = __fadd_rn(acc[gI][1], d.y);
                if (a.z == want + 2) acc[gI][2] = __fadd_rn(acc[gI][2], d.z);
                if (a.w == want + 3) acc[gI][3] = __fadd_rn(acc[gI][3], d.w);
              } else {
                if (a == want) acc[gI][0] = __fadd_rn(acc[gI][0], d);
              }
            }
          }
        }
      }
    }
    __syncthreads();
  }
  float* out = bottom_diff + (size_t)cell * C;
#pragma unroll
  for (int gI = 0; gI < MAXG; ++gI) {
    const int cv = tid + gI * blockDim.x;
    if (cv < CV) {
      if constexpr (VEC == 4)
        *reinterpret_cast<float4*>(out + cv * 4) =
            make_float4(acc[gI][0], acc[gI][1], acc[gI][2], acc[gI][3]);
      else
        out[cv] = acc[gI][0];
    }
  }
}

int pick_block_x(int CV) {
  if (CV >= 256) return 256;
  if (CV >= 32) return ((CV + 31) / 32) * 32;
  int p = 1;
  while (p < CV) p <<= 1;
  return p;
}

}  // namespace

namespace {

// Shape test + launch geometry of the tiled forward kernel.
struct TiledPlan {
  bool ok;
  bool scan;          // RoI lists built in-CTA (no workspace)
  int RB, nchunks;
  size_t smem;
};

}  // namespace

// img_start[B+2] | perm[R] | counts[B+1] | cursor[B+1] | ticket
size_t wssdl_roi::bucket_workspace_bytes(int B, int R) {
  return sizeof(int) * ((size_t)B + 2 + (size_t)R + 2 * ((size_t)B + 1) + 1) + 16;
}

// Groups the RoIs by image into the caller's workspace (see roi_hist_kernel).
cudaError_t wssdl_roi::launch_roi_bucket(const float* rois, int R, int B, void* workspace, cudaStream_t s,
                              int** img_start_out, int** perm_out) {
  int* img_start = static_cast<int*>(workspace);
  int* perm = img_start + (B + 2);
  int* counts = perm + R;
  int* cursor = counts + (B + 1);
  int* ticket = cursor + (B + 1);
  cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(int) * (2 * ((size_t)B + 1) + 1), s);
  if (e != cudaSuccess) return e;
  int grid = (R + 1023) / 1024;                     // ~4 RoIs per thread
  if (grid > 2 * WSSDL_NUM_SMS) grid = 2 * WSSDL_NUM_SMS;
  if (grid < 1) grid = 1;
  roi_hist_kernel<<<grid, 256, 0, s>>>(rois, R, B, counts, ticket, img_start, cursor);
  roi_scatter_kernel<<<grid, 256, 0, s>>>(rois, R, B, cursor, perm);
  *img_start_out = img_start;
  *perm_out = perm;
  return cudaGetLastError();
}

namespace {

TiledPlan plan_tiled(int B, int H, int W, int C, int R, int PH, int PW, bool vec4,
                     size_t workspace_bytes) {
  TiledPlan p = {false, false, 0, 1, 0};
  if (!vec4 || C % T_SLICE != 0 || H > 65535 || W > 65535 || B + 1 > 65535) return p;
  if (PH <= 0 || PW <= 0) return p;
  p.scan = R <= T_SCAN_MAX_R;
  if (!p.scan && workspace_bytes < bucket_workspace_bytes(B, R)) return p;
  const size_t map_bytes = (size_t)H * W * T_SLICE * sizeof(float);
  const size_t list_bytes = p.scan ? sizeof(int) * (size_t)R : 0;
  const size_t budget = T_DYN_SMEM_MAX;
  if (map_bytes + list_bytes >= budget) return p;
  // per resident RoI: PH + PW packed edges, its index, cost class and sorted position
  const size_t per_roi = sizeof(int) * ((size_t)PH + PW + 3);
  const size_t rb = (budget - map_bytes - list_bytes) / per_roi;
  if (rb < 16) return p;
  p.RB = (int)(rb < (size_t)T_MAX_RB ? rb : (size_t)T_MAX_RB);
  // no point keeping more RoIs resident than a CTA will ever see
  if (p.RB > R) p.RB = R < 16 ? 16 : R;          // an image may own every RoI
  p.smem = map_bytes + list_bytes + per_roi * p.RB;
  // split an image's RoIs over CTAs only while the grid stays inside one wave and a chunk
  // still amortises staging the slice (>= 16 RoIs)
  const long long base = (long long)(C / T_SLICE) * (B > 0 ? B : 1);
  const int avg = R / (B > 0 ? B : 1);
  while (base * p.nchunks * 2 <= WSSDL_NUM_SMS && avg / (p.nchunks * 2) >= 16) p.nchunks *= 2;
  p.ok = true;
  return p;
}

// Shape test + launch geometry of the band forward kernel.
struct BandPlan {
  bool ok;
  bool scan;
  int RB, nchunks;
  BandGeom g;
  size_t smem;
};

BandPlan plan_band(int B, int H, int W, int C, int R, int PH, int PW, bool aligned,
                   size_t workspace_bytes) {
  BandPlan p = {false, false, 0, 1, {1, 0, 0}, 0};
  if (!aligned || C % B_SLICE != 0 || H > 65535 || W > 65535 || B + 1 > 65535) return p;
  if (PH <= 0 || PW <= 0 || PH > 255 || H <= 0 || W <= 0) return p;
  p.scan = R <= B_SCAN_MAX_R;
  if (!p.scan && workspace_bytes < bucket_workspace_bytes(B, R)) return p;
  const size_t row_bytes = (size_t)W * B_SLICE * sizeof(float);
  const size_t list_bytes = p.scan ? sizeof(int) * (size_t)R : 0;
  const size_t per_roi = sizeof(int) * ((size_t)PH + PW + 3);
  const size_t budget = T_DYN_SMEM_MAX - 128;       // 128: alignment slack of the map
  // room for the bin edges of a useful number of RoIs first, rows with what is left
  const size_t rb_want = (size_t)(R < 16 ? 16 : (R < 512 ? R : 512));
  if (list_bytes + per_roi * rb_want + row_bytes >= budget) return p;
  const long long hb_max = (long long)((budget - list_bytes - per_roi * rb_want) / row_bytes);
  // the tallest bin of a RoI that lies inside the map, +1 for GPU_CEIL's overlapping edges
  const int ov = (H + 1 + PH - 1) / PH + 2;
  if (hb_max >= H) {
    p.g.NB = 1; p.g.Hb = H; p.g.step = H;
  } else {
    if (hb_max <= ov) return p;
    p.g.NB = (int)((H - ov + (hb_max - ov) - 1) / (hb_max - ov));
    p.g.step = (H - ov + p.g.NB - 1) / p.g.NB;
    p.g.Hb = p.g.step + ov;
  }
  const size_t map_bytes = row_bytes * p.g.Hb;
  size_t rb = (budget - map_bytes - list_bytes) / per_roi;
  if (rb > (size_t)B_MAX_RB) rb = B_MAX_RB;
  if (rb > (size_t)R) rb = R < 16 ? 16 : R;
  p.RB = (int)rb;
  p.smem = map_bytes + list_bytes + per_roi * p.RB + 128;
  const long long base = (long long)(C / B_SLICE) * (B > 0 ? B : 1) * p.g.NB;
  const int avg = R / (B > 0 ? B : 1);
  while (base * p.nchunks * 2 <= WSSDL_NUM_SMS && avg / (p.nchunks * 2) >= 16) p.nchunks *= 2;
  if ((long long)p.g.NB * p.nchunks > 65535) return p;
  p.ok = true;
  return p;
}

}  // namespace

namespace {

// bucket lists of the band / tiled kernels, or everything the sorted-bins kernels need
size_t fwd_workspace_bytes(int B, int R, int PH, int PW) {
  const size_t a = bucket_workspace_bytes(B, R), b = bins_workspace_bytes(B, R, PH, PW);
  return a > b ? a : b;
}

// Which forward kernel a call takes, and its launch geometry.  One function for the launcher
// and for the host-only query wssdl_roi_pool_fwd_plan (CPU tests check the invariants).
enum { FWD_DIRECT = 0, FWD_TILED = 1, FWD_BAND = 2, FWD_BINS = 3 };
struct FwdChoice {
  int kernel;
  TiledPlan tp;
  BandPlan bp;
  BinsPlan np;
};

// force: 0 = by shape, 1 = direct, 2 = tiled, 3 = band, 4 = sorted bins
// (WSSDL_TUNE_ROI_FWD_KERNEL).  aligned = the pointer requirements of the shared-memory kernels
// hold (16 B inputs, 32 B outputs, C % 4).
FwdChoice choose_fwd(int B, int H, int W, int C, int R, int PH, int PW, bool aligned,
                     size_t workspace_bytes, int force) {
  FwdChoice c;
  c.kernel = FWD_DIRECT;
  c.tp = plan_tiled(B, H, W, C, R, PH, PW, aligned, workspace_bytes);
  c.bp = plan_band(B, H, W, C, R, PH, PW, aligned, workspace_bytes);
  c.np = plan_bins(B, H, W, C, R, PH, PW, aligned, workspace_bytes,
                   wssdl_tuning(WSSDL_TUNE_ROI_FWD_THREADS));
  const int sl = wssdl_tuning(WSSDL_TUNE_ROI_FWD_SLICES), ch = wssdl_tuning(WSSDL_TUNE_ROI_FWD_CHUNKS);
  if (sl > 0 && sl <= C / 32) c.np.sg = sl;
  if (ch > 0 && (long long)ch * c.np.g.NB <= 65535) c.np.nchunks = ch;
  // the shared-memory kernels pay where RoIs re-read the map and bins are small (7x7-like);
  // big bins (C3 14x14x1024) run the direct kernel at the HBM roofline already
  const bool reuse = PH * PW <= 64 && (long long)R * PH * PW * 2 >= (long long)B * H * W;
  // The tiled kernel only takes what the 32-channel kernels cannot (C % 32 != 0), while its
  // grid fits one wave (profiles/history/r01_roi_fwd_direct_vs_tiled.txt).
  const bool tiled_pays = c.tp.scan && reuse &&
                          (long long)(C / T_SLICE) * B * c.tp.nchunks <= WSSDL_NUM_SMS;
  // Sorted bins wherever the grid is more than one wave; a problem that fits one wave (one or two
  // images) runs the band kernel: one launch instead of pre-pass + pooling (measured on B200,
  // C1 1x300: 31 us against 37-52 us; 16 images: 0.254 ms sorted against 0.27-0.33 ms band).
  const long long band_ctas =
      (long long)(C / 32) * (B > 0 ? B : 1) * c.bp.g.NB * c.bp.nchunks;
  const bool one_wave = c.bp.ok && c.bp.g.NB <= 4 && band_ctas <= WSSDL_NUM_SMS;
  if (c.np.ok && c.np.g.NB <= 4 && (force == 4 || (force == 0 && reuse && !one_wave))) c.kernel = FWD_BINS;
  else if (c.bp.ok && (force == 3 || (force == 0 && reuse && one_wave))) c.kernel = FWD_BAND;
  else if (c.tp.ok && force != 1 && (force == 2 || (force == 0 && tiled_pays))) c.kernel = FWD_TILED;
  return c;
}

}  // namespace

extern "C" int wssdl_roi_pool_fwd_plan(int B, int H, int W, int C, int R, int PH, int PW,
                                       int with_workspace, int force, int* out) {
  if (!out || B < 0 || H < 0 || W < 0 || C < 0 || R < 0 || PH < 0 || PW < 0) return WSSDL_EINVAL;
  if (force < 0 || force > 4) return WSSDL_EINVAL;
  const size_t ws = with_workspace ? fwd_workspace_bytes(B, R, PH, PW) : 0;
  const FwdChoice c = choose_fwd(B, H, W, C, R, PH, PW, C % 4 == 0, ws, force);
  for (int i = 0; i < 10; ++i) out[i] = 0;
  out[0] = c.kernel;
  if (c.kernel == FWD_BINS) {
    out[1] = c.np.g.NB; out[2] = c.np.g.Hb; out[3] = c.np.g.step; out[4] = c.np.nchunks;
    out[5] = c.np.sort_rch; out[6] = (int)c.np.smem; out[7] = c.np.scan ? 1 : 0;
    out[8] = c.np.sg; out[9] = c.np.threads;
  } else if (c.kernel == FWD_BAND) {
    out[1] = c.bp.g.NB; out[2] = c.bp.g.Hb; out[3] = c.bp.g.step; out[4] = c.bp.nchunks;
    out[5] = c.bp.RB; out[6] = (int)c.bp.smem; out[7] = c.bp.scan ? 1 : 0;
  } else if (c.kernel == FWD_TILED) {
    out[1] = 1; out[2] = H; out[3] = H; out[4] = c.tp.nchunks;
    out[5] = c.tp.RB; out[6] = (int)c.tp.smem; out[7] = c.tp.scan ? 1 : 0;
  }
  return WSSDL_OK;
}

extern "C" size_t wssdl_roi_pool_fwd_workspace_bytes(int B, int R, int PH, int PW) {
  if (B < 0 || R < 0) return 0;
  return fwd_workspace_bytes(B, R, PH, PW);
}

extern "C" int wssdl_roi_pool_fwd(const float* bottom, const float* rois, int B, int H, int W,
                                  int C, int R, int PH, int PW, float spatial_scale,
                                  int bin_mode, float* top, int* argmax, void* workspace,
                                  size_t workspace_bytes, wssdl_stream_t stream) {
  return wssdl_roi_pool_fwd_impl(bottom, rois, B, H, W, C, R, PH, PW, spatial_scale, bin_mode, top,
                                 argmax, workspace, workspace_bytes, stream, 0);
}

extern "C" int wssdl_roi_pool_fwd_grouped(const float* bottom, const float* rois, int roi_stride,
                                          int B, int H, int W, int C, int PH, int PW,
                                          float spatial_scale, int bin_mode, float* top,
                                          int* argmax, void* workspace, size_t workspace_bytes,
                                          wssdl_stream_t stream) {
  if (roi_stride < 0 || B < 0 || (long long)roi_stride * B >= (1ll << 31)) return WSSDL_EINVAL;
  return wssdl_roi_pool_fwd_impl(bottom, rois, B, H, W, C, roi_stride * B, PH, PW, spatial_scale,
                                 bin_mode, top, argmax, workspace, workspace_bytes, stream,
                                 roi_stride);
}

// grouped_stride > 0: the caller guarantees image-major RoIs (row r belongs to image
// r / grouped_stride; rows with another batch index -- the -1 padding rows of the hot-path
// entry -- pool to zeros / -1 in every kernel): the sorted-bins pre-pass then needs no RoI lists.
int wssdl_roi_pool_fwd_impl(const float* bottom, const float* rois, int B, int H, int W, int C,
                            int R, int PH, int PW, float spatial_scale, int bin_mode, float* top,
                            int* argmax, void* workspace, size_t workspace_bytes,
                            wssdl_stream_t stream, int grouped_stride) {
  // attribute checks of the op (roi_pooling_op.cc:73-82) plus pointer sanity
  if (B < 0 || H < 0 || W < 0 || C < 0 || R < 0 || PH < 0 || PW < 0) return WSSDL_EINVAL;
  if (bin_mode != WSSDL_BIN_CPU_TRUNC && bin_mode != WSSDL_BIN_GPU_CEIL) return WSSDL_EINVAL;
  const long long out_elems = (long long)R * PH * PW * C;
  if (out_elems == 0) return WSSDL_OK;
  if (!bottom || !rois || !top) return WSSDL_EINVAL;
  if ((long long)H * W * C >= (1ll << 31) || (long long)R * PH >= (1ll << 31)) return WSSDL_ELIMIT;
  const bool vec4 = (C % 4 == 0) && aligned16(bottom) && aligned16(top) &&
                    (argmax == nullptr || aligned16(argmax));
  cudaStream_t s = to_cuda(stream);

  // Kernel choice (choose_fwd above); WSSDL_TUNE_ROI_FWD_KERNEL forces one (tests, experiments).
  const int force = wssdl_tuning(WSSDL_TUNE_ROI_FWD_KERNEL);
  if (workspace == nullptr) workspace_bytes = 0;
  // the shared-memory kernels store 256 bits per lane: outputs must be 32-byte aligned
  const bool al32 = ((reinterpret_cast<uintptr_t>(top) | reinterpret_cast<uintptr_t>(argmax)) & 31u) == 0;
  const FwdChoice choice = choose_fwd(B, H, W, C, R, PH, PW, vec4 && al32, workspace_bytes, force);
  const TiledPlan& tp = choice.tp;
  const BandPlan& bp = choice.bp;
  if (choice.kernel == FWD_BINS) {
    WSSDL_RETURN_IF_CUDA(launch_fwd_bins(choice.np, bottom, rois, B, H, W, C, R, PH, PW,
                                         spatial_scale, bin_mode, top, argmax, workspace, s,
                                         grouped_stride));
    return WSSDL_OK;
  }
  if (choice.kernel == FWD_BAND) {
    int* img_start = nullptr;
    int* perm = nullptr;
    if (!bp.scan) {
      WSSDL_RETURN_IF_CUDA(launch_roi_bucket(rois, R, B, workspace, s, &img_start, &perm));
    }
    const FastDiv dPW = make_fastdiv((unsigned)PW);
    // C % 128 == 0: the argmax update records (shared address) * (C/128), see the kernel
    const bool linear = (C % 128 == 0);
    const int km = linear ? C / 128 : 1;
    const Ones ones = {1.0f, {km, km, km, km, km, km, km, km}};
    dim3 grid((unsigned)(C / B_SLICE), (unsigned)(bp.g.NB * bp.nchunks), (unsigned)(B + 1));
#define LAUNCH_BAND(M, A, L)                                                                   \
  do {                                                                                         \
    static unsigned long long done_k = 0;                                                      \
    WSSDL_RETURN_IF_CUDA(allow_big_smem(roi_pool_fwd_band_kernel<M, A, L>, &done_k));          \
    roi_pool_fwd_band_kernel<M, A, L><<<grid, B_THREADS, bp.smem, s>>>(                        \
        bottom, rois, perm, img_start, B, H, W, C, R, PH, PW, spatial_scale, bp.RB, bp.nchunks, \
        bp.g, dPW, ones, top, argmax);                                                         \
  } while (0)
#define LAUNCH_BAND_A(M, A)                                                                    \
  do {                                                                                         \
    if (linear) LAUNCH_BAND(M, A, true);                                                       \
    else LAUNCH_BAND(M, A, false);                                                             \
  } while (0)
    if (bin_mode == WSSDL_BIN_CPU_TRUNC) {
      if (argmax) LAUNCH_BAND_A(WSSDL_BIN_CPU_TRUNC, true);
      else LAUNCH_BAND_A(WSSDL_BIN_CPU_TRUNC, false);
    } else {
      if (argmax) LAUNCH_BAND_A(WSSDL_BIN_GPU_CEIL, true);
      else LAUNCH_BAND_A(WSSDL_BIN_GPU_CEIL, false);
    }
#undef LAUNCH_BAND_A
#undef LAUNCH_BAND
    WSSDL_CHECK_LAUNCH();
    return WSSDL_OK;
  }
  if (choice.kernel == FWD_TILED) {
    int* img_start = nullptr;
    int* perm = nullptr;
    if (!tp.scan) {
      WSSDL_RETURN_IF_CUDA(launch_roi_bucket(rois, R, B, workspace, s, &img_start, &perm));
    }
    const FastDiv dPW = make_fastdiv((unsigned)PW);
    const Ones ones = {1.0f, {1, 1, 1, 1, 1, 1, 1, 1}};
    dim3 grid((unsigned)(C / T_SLICE), (unsigned)tp.nchunks, (unsigned)(B + 1));
#define LAUNCH_TILED(M, ST)                                                                    \
  do {                                                                                         \
    static unsigned long long done_t = 0;                                                      \
    WSSDL_RETURN_IF_CUDA(allow_big_smem(roi_pool_fwd_tiled_kernel<M, ST>, &done_t));           \
    roi_pool_fwd_tiled_kernel<M, ST><<<grid, T_THREADS, tp.smem, s>>>(                         \
        bottom, rois, perm, img_start, B, H, W, C, R, PH, PW, spatial_scale, tp.RB, dPW, ones,  \
        top, argmax);                                                                          \
  } while (0)
    if (bin_mode == WSSDL_BIN_CPU_TRUNC) LAUNCH_TILED(WSSDL_BIN_CPU_TRUNC, true);
    else LAUNCH_TILED(WSSDL_BIN_GPU_CEIL, true);
#undef LAUNCH_TILED
    WSSDL_CHECK_LAUNCH();
    return WSSDL_OK;
  }

  const int CV = vec4 ? C / 4 : C;
  // channel vectors per thread: 2 when the channel count allows it
  const int vpt = (vec4 && CV % 64 == 0) ? 2 : 1;
  const int lanes = CV / vpt;
  dim3 block(pick_block_x(lanes), 1);
  block.y = max(1, min(PW, 128 / (int)block.x));
  dim3 grid((unsigned)(R * PH), (unsigned)ceil_div(lanes, block.x));
#define LAUNCH_FWD(V, M, CVT, VPT)                                                            \
  roi_pool_fwd_kernel<V, M, CVT, VPT><<<grid, block, 0, s>>>(bottom, rois, B, H, W, C, PH, PW, \
                                                             spatial_scale, top, argmax)
#define LAUNCH_FWD_MODE(V, CVT, VPT)                                                          \
  do {                                                                                        \
    if (bin_mode == WSSDL_BIN_CPU_TRUNC) LAUNCH_FWD(V, WSSDL_BIN_CPU_TRUNC, CVT, VPT);        \
    else LAUNCH_FWD(V, WSSDL_BIN_GPU_CEIL, CVT, VPT);                                         \
  } while (0)
  if (vec4 && vpt == 2) {
    if (C == 512) LAUNCH_FWD_MODE(4, 128, 2);         // VGG-16 conv5_3
    else if (C == 1024) LAUNCH_FWD_MODE(4, 256, 2);   // ResNet C4
    else LAUNCH_FWD_MODE(4, 0, 2);
  } else if (vec4) {
    if (C == 512) LAUNCH_FWD_MODE(4, 128, 1);
    else if (C == 1024) LAUNCH_FWD_MODE(4, 256, 1);
    else LAUNCH_FWD_MODE(4, 0, 1);
  } else {
    LAUNCH_FWD_MODE(1, 0, 1);
  }
#undef LAUNCH_FWD_MODE
#undef LAUNCH_FWD
  WSSDL_CHECK_LAUNCH();
  return WSSDL_OK;
}

extern "C" int wssdl_roi_pool_bwd(const float* top_diff, const int* argmax, const float* rois,
                                  int B, int H, int W, int C, int R, int PH, int PW,
                                  float spatial_scale, int bwd_mode, float* bottom_diff,
                                  wssdl_stream_t stream) {
  if (B < 0 || H < 0 || W < 0 || C < 0 || R < 0 || PH < 0 || PW < 0) return WSSDL_EINVAL;
  if (bwd_mode != WSSDL_BWD_ATOMIC && bwd_mode != WSSDL_BWD_GATHER) return WSSDL_EINVAL;
  const long long in_elems = (long long)B * H * W * C;
  if (in_elems == 0) return WSSDL_OK;
  if (!bottom_diff) return WSSDL_EINVAL;
  if ((long long)H * W * C >= (1ll << 31)) return WSSDL_ELIMIT;
  cudaStream_t s = to_cuda(stream);
  const long long out_elems = (long long)R * PH * PW * C;
  if (out_elems == 0) {
    WSSDL_RETURN_IF_CUDA(cudaMemsetAsync(bottom_diff, 0, (size_t)in_elems * sizeof(float), s));
    return WSSDL_OK;
  }
  if (!top_diff || !argmax || !rois) return WSSDL_EINVAL;
  const bool vec4 = (C % 4 == 0) && aligned16(top_diff) && aligned16(argmax) &&
                    aligned16(bottom_diff);
  const int CV = vec4 ? C / 4 : C;
  if (bwd_mode == WSSDL_BWD_ATOMIC) {
    if (PH > 4096 || PW > 4096) return WSSDL_ELIMIT;
    WSSDL_RETURN_IF_CUDA(cudaMemsetAsync(bottom_diff, 0, (size_t)in_elems * sizeof(float), s));
    // split PH over blockIdx.y until the grid covers the machine a few times
    int split = 1;
    while (split < PH && (long long)R * split < 4 * WSSDL_NUM_SMS) split *= 2;
    split = min(split, PH);
    dim3 grid((unsigned)R, (unsigned)split);
    const size_t smem = sizeof(int) * 2 * (size_t)(PH + PW);
    const FastDiv dC = make_fastdiv((unsigned)C), dW = make_fastdiv((unsigned)W);
    const FastDiv dCV = make_fastdiv((unsigned)CV), dPW = make_fastdiv((unsigned)PW);
    if (vec4)
      roi_pool_bwd_atomic_kernel<4><<<grid, 256, smem, s>>>(top_diff, argmax, rois, B, H, W, C,
                                                            PH, PW, spatial_scale, dC, dW,
                                                            dCV, dPW, bottom_diff);
    else
      roi_pool_bwd_atomic_kernel<1><<<grid, 256, smem, s>>>(top_diff, argmax, rois, B, H, W, C,
                                                            PH, PW, spatial_scale, dC, dW,
                                                            dCV, dPW, bottom_diff);
  } else {
    if (PH > 255 || PW > 255) return WSSDL_ELIMIT;
    int bx = pick_block_x(CV);
    if (bx < 32) bx = 32;
    if ((long long)bx * 4 < CV) return WSSDL_ELIMIT;  // C <= 4096 (vec4) per cell CTA
    dim3 grid((unsigned)(B * H * W));
    if (vec4)
      roi_pool_bwd_gather_kernel<4><<<grid, bx, 0, s>>>(top_diff, argmax, rois, B, H, W, C, R, PH,
                                                        PW, spatial_scale, bottom_diff);
    else
      roi_pool_bwd_gather_kernel<1><<<grid, bx, 0, s>>>(top_diff, argmax, rois, B, H, W, C, R, PH,
                                                        PW, spatial_scale, bottom_diff);
  }
  WSSDL_CHECK_LAUNCH();
  return WSSDL_OK;
}
