// Device / launch helpers shared by the RoI-pool forward kernels (roi_pool.cu, roi_pool_bins.cu).
#pragma once
#include "common.cuh"

namespace wssdl_roi {

// Exact unsigned division by a runtime constant d >= 1 for n < 2^31 (Granlund-Montgomery
// round-up magic: l = ceil(log2 d), m = ceil(2^(31+l)/d) < 2^32, q = (n*m) >> (31+l)).
struct FastDiv {
  unsigned mul, shift;
};
inline FastDiv make_fastdiv(unsigned d) {
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

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

__device__ __forceinline__ int roi_bucket(float batch, int B) {
  const int b = (int)batch;             // same conversion as roi_cells()
  return (b >= 0 && b < B) ? b : B;
}

// First-maximum update with the two conditional moves on the FMA pipe.  FSETP, FSEL and SEL
// all issue to the half-rate ALU pipe (16 lanes/clk per scheduler), which is what bounded
// the tiled kernel (ncu: math_pipe_throttle, ALU 70 %, FMA 20 %).  `@p FMUL m, v, 1.0f` and
// `@p IMAD mi, cell, 1, 0` are exact, run on the full-rate FP32 pipe / the FMA-heavy pipe,
// and leave one ALU instruction (the compare) per element.  `one_f` / `one_i` come from
// kernel parameters so ptxas cannot fold the multiplications back into selects; the eight
// channels of a thread use eight distinct integer ones, otherwise ptxas merges their
// common cell*1 product and falls back to SEL.
struct Ones {
  float f;
  int i[8];
};
__device__ __forceinline__ void upd_fma(float v, int cell, float& m, int& mi, float one_f,
                                        int one_i) {
  asm("{\n\t.reg .pred p;\n\t"
      "setp.gt.f32 p, %2, %0;\n\t"            // strict '>' (cc:187): NaN never wins
      "@p mul.rn.f32 %0, %2, %4;\n\t"
      "@p mad.lo.s32 %1, %3, %5, 0;\n\t}"
      : "+f"(m), "+r"(mi)
      : "f"(v), "r"(cell), "f"(one_f), "r"(one_i));
}

// 256-bit global stores (sm_100a: STG.E.256): one lane writes 8 consecutive channels, the
// lanes of a bin column fill a 64 B half line (tiled) / a full 128 B line (band, bins) in
// one LSU wavefront.
template <bool STREAM_ST>
__device__ __forceinline__ void st256(float* p, const float4 a, const float4 b) {
  if (STREAM_ST)
    asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x),
                 "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
  else
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x),
                 "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}
template <bool STREAM_ST>
__device__ __forceinline__ void st256(int* p, const int4 a, const int4 b) {
  if (STREAM_ST)
    asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x),
                 "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
  else
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x),
                 "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

// dynamic shared memory a CTA may ask for: 227 KB minus the kernels' static arrays
constexpr int T_DYN_SMEM_MAX = 227 * 1024 - 1024;

// Row bands of the map held in shared memory (band and bins kernels): band b holds rows
// [b*step, b*step + Hb) and owns the bins whose first row falls into [b*step, (b+1)*step).
struct BandGeom {
  int NB, Hb, step;   // bands per image, rows per band, first-row distance of two bands
};

// One bin whose rows are not all resident in this band: pooled straight from global memory,
// one channel at a time.  Rare (RoIs several times taller than the map, never the detector's
// own proposals), so it is kept out of line and as small in registers as possible: the hot
// loop's register allocation must not pay for it.  (A vectorised version that stored through
// the st.v8 inline asm of st256 was narrowed to a scalar store by ptxas 12.9 in some clones of
// the out-of-line function; caught by the tall-RoI test.)
template <bool HAS_ARGMAX>
__device__ __noinline__ void band_slow_bin(const float* __restrict__ img_base, int hs, int he,
                                           int ws, int nw, int W, int C, int c_lo,
                                           float* __restrict__ top_o, int* __restrict__ arg_o) {
#pragma unroll 1
  for (int k = 0; k < 8; ++k) {
    float m = -FLT_MAX;
    int mi = -1;
#pragma unroll 1
    for (int h = hs; h < he; ++h) {
      int idx = (h * W + ws) * C + c_lo + k;
#pragma unroll 1
      for (int w = 0; w < nw; ++w, idx += C) {
        const float v = __ldg(img_base + idx);
        if (v > m) { m = v; mi = idx; }            // strict '>' (cc:187)
      }
    }
    top_o[k] = m;
    if (HAS_ARGMAX) arg_o[k] = mi;
  }
}

// Opt a kernel into 227 KB of dynamic shared memory once per device (the attribute is
// per device; one process may drive several).
template <typename K>
cudaError_t allow_big_smem(K kernel, unsigned long long* done_mask) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ull << (dev & 63);
  if (__atomic_load_n(done_mask, __ATOMIC_ACQUIRE) & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T_DYN_SMEM_MAX);
  if (e == cudaSuccess) __atomic_fetch_or(done_mask, bit, __ATOMIC_RELEASE);
  return e;
}

// img_start[B+2] | perm[R] | counts[B+1] | cursor[B+1] | ticket   (roi_pool.cu)
size_t bucket_workspace_bytes(int B, int R);
// Groups the RoIs by image into the caller's workspace (roi_hist_kernel + roi_scatter_kernel).
cudaError_t launch_roi_bucket(const float* rois, int R, int B, void* workspace, cudaStream_t s,
                              int** img_start_out, int** perm_out);

// ---- the class-sorted-bins forward kernels (roi_pool_bins.cu)
struct BinsPlan {
  bool ok;
  bool scan;            // R <= 4096: the sort pre-pass finds each image's RoIs itself
  BandGeom g;
  int nchunks;          // ranges a band's groups of bins are split into (CTAs per band and slice group)
  int sg;               // channel slices (of 32) one CTA pools one after the other
  int threads;          // threads per pooling CTA: 1024 (one CTA per SM) or 512 (two)
  int rec_cap;          // pooling kernel: bin records resident at a time
  size_t smem;          // pooling kernel
  int sort_rch;         // sort pre-pass: RoIs whose geometry is resident at a time
  size_t sort_smem;
};
size_t bins_workspace_bytes(int B, int R, int PH, int PW);
BinsPlan plan_bins(int B, int H, int W, int C, int R, int PH, int PW, bool aligned,
                   size_t workspace_bytes, int threads);
cudaError_t launch_fwd_bins(const BinsPlan& p, const float* bottom, const float* rois, int B,
                            int H, int W, int C, int R, int PH, int PW, float spatial_scale,
                            int bin_mode, float* top, int* argmax, void* workspace,
                            cudaStream_t s, int grouped_stride = 0);

}  // namespace wssdl_roi
