// RoI max pooling, NHWC, forward (+argmax) and backward, for sm_100a.
//
// Semantics: RoiPoolOp / RoiPoolGradOp of the reference
// (roi_pooling_layer/roi_pooling_op.cc:137-196, :383-458); CUDA twin
// (roi_pooling_op_gpu.cu.cc:19-85) selectable as bin_mode GPU_CEIL.
//
// Forward design (HBM-write bound: 8 B written per output element, the feature map is
// L2 resident):
//   - one CTA per (roi, ph) output row; threads run along C in float4 lanes, so every
//     feature-map cell is read as one fully coalesced 16 B/lane segment and every output
//     row (PW*C floats + PW*C ints) is written as contiguous 128-bit streaming stores;
//   - the RoI geometry (4 rounds, 2 IEEE divisions) is computed once per thread and
//     reused for all PW bins instead of once per output element as in the reference;
//   - each thread scans its bin h-major / w-minor with strict '>' exactly like
//     roi_pooling_op.cc:184-191, so ties (post-ReLU zeros) resolve to the first cell with
//     no cross-lane reduction needed: channels, not cells, are the parallel axis in NHWC;
//   - CPU_TRUNC bins never overlap, so a cell is read once per RoI: there is no reuse for
//     shared memory to capture; loads go through the read-only path (ld.global.nc) and
//     the outputs, never re-read by this kernel, use st.global.cs so they do not evict
//     the feature map from L2.
#include <stdlib.h>

#include "common.cuh"

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
// Exact unsigned division by a runtime constant d >= 1 for n < 2^31 (Granlund-Montgomery
// round-up magic: l = ceil(log2 d), m = ceil(2^(31+l)/d) < 2^32, q = (n*m) >> (31+l)).
struct FastDiv {
  unsigned mul, shift;
};
FastDiv make_fastdiv(unsigned d) {
  unsigned l = 0;
  while ((1ull << l) < d) ++l;
  FastDiv f;
  f.mul = (unsigned)(((1ull << (31 + l)) + d - 1) / d);
  f.shift = 31 + l;
  return f;
}
__device__ __forceinline__ unsigned fastdiv(unsigned n, FastDiv f) {
  return (unsigned)(((unsigned long long)n * f.mul) >> f.shift);
}

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

extern "C" int wssdl_roi_pool_fwd(const float* bottom, const float* rois, int B, int H, int W,
                                  int C, int R, int PH, int PW, float spatial_scale,
                                  int bin_mode, float* top, int* argmax,
                                  wssdl_stream_t stream) {
  // attribute checks of the op (roi_pooling_op.cc:73-82) plus pointer sanity
  if (B < 0 || H < 0 || W < 0 || C < 0 || R < 0 || PH < 0 || PW < 0) return WSSDL_EINVAL;
  if (bin_mode != WSSDL_BIN_CPU_TRUNC && bin_mode != WSSDL_BIN_GPU_CEIL) return WSSDL_EINVAL;
  const long long out_elems = (long long)R * PH * PW * C;
  if (out_elems == 0) return WSSDL_OK;
  if (!bottom || !rois || !top) return WSSDL_EINVAL;
  if ((long long)H * W * C >= (1ll << 31) || (long long)R * PH >= (1ll << 31)) return WSSDL_ELIMIT;
  const bool vec4 = (C % 4 == 0) && aligned16(bottom) && aligned16(top) &&
                    (argmax == nullptr || aligned16(argmax));
  const int CV = vec4 ? C / 4 : C;
  // channel vectors per thread: 2 when the channel count allows it (tuning knob for
  // experiments: WSSDL_ROI_FWD_VPT=1|2)
  static const int vpt_env = [] { const char* e = getenv("WSSDL_ROI_FWD_VPT"); return e ? atoi(e) : 0; }();
  int vpt = (vec4 && CV % 64 == 0) ? 2 : 1;
  if (vpt_env == 1) vpt = 1;
  const int lanes = CV / vpt;
  dim3 block(pick_block_x(lanes), 1);
  block.y = max(1, min(PW, 128 / (int)block.x));
  dim3 grid((unsigned)(R * PH), (unsigned)ceil_div(lanes, block.x));
  cudaStream_t s = to_cuda(stream);
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
