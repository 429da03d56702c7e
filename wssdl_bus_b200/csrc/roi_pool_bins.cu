// RoI max pooling forward, NHWC, "class-sorted bins" kernel for sm_100a.
//
// Semantics: RoiPoolOp of the reference (roi_pooling_layer/roi_pooling_op.cc:137-196; CUDA
// twin roi_pooling_op_gpu.cu.cc:19-85 as bin_mode GPU_CEIL).  Same bytes as the other forward
// kernels of roi_pool.cu.
//
// What bounded the band kernel it replaces (ncu, profiles/r01_ncu_roi_fwd_band.txt): 2.67 G
// warp instructions at 70 % issue utilisation, of which 58 % sit in the cell loop with 21 of 32
// lanes active -- its work item is a COLUMN of bins, and the columns that share a warp are k or
// k+1 cells wide, their RoIs k or k+1 rows tall.  Here the work item is ONE BIN, and the bins a
// CTA owns are counting-sorted by their exact size class (rows x cells per row, 1..8 each) before
// any pooling happens, every class padded to whole warps:
//   - the 8 bins of a warp (4 lanes x 8 channels each) have identical trip counts, so the row and
//     cell loops are warp-uniform: no divergence, and the cell loop is unrolled over the class's
//     cells per row with all loads of a row in flight;
//   - the proposal layer's bins are small (4.6 cells on average, 14 % are a single cell), so the
//     per-bin bookkeeping is what is left: one 32-bit record per bin (first cell, bin id, RoI)
//     read from shared memory, the flat argmax index produced by the conditional move itself
//     (`@p IMAD mi, addr, C/128, const`), outputs addressed by one IMAD.WIDE each.
// Data movement is the band kernel's: a CTA owns (image, row band, RoI chunk, 32-channel slices);
// a cell of the slice is one 128 B bank row; four lanes share a bin (8 channels each); the two
// bins of a quarter warp read their two 16 B chunks in opposite order, so every LDS.128 phase
// touches all eight 16 B bank groups once whatever the two cells are; the four lanes of a bin
// store one full 128 B line per STG.256 wavefront.  The band is staged by ONE TMA tensor copy
// (cp.async.bulk.tensor.4d: box 32 channels x W x Hb rows of the NHWC map, completion on an
// mbarrier) issued by one thread before the sort, so the sort runs while the band lands.
#include <cuda.h>   // CUtensorMap and its enums only: the encoder is fetched through the runtime

#include "common.cuh"
#include "roi_pool_dev.cuh"

using namespace wssdl_roi;

namespace {

#ifndef WSSDL_BINS_INFLIGHT
#define WSSDL_BINS_INFLIGHT 2
#endif
constexpr int N_SLICE = 32;             // channels per slice: one cell = 128 B = all 32 banks
constexpr int N_DIM = 8;                // bins up to 8 x 8 cells have a size class of their own
constexpr int CLS_EMPTY = 64;           // (0, -1) bins (roi_pooling_op.cc:180-182)
constexpr int CLS_SLOW = 65;            // bins with rows outside the resident band
constexpr int N_CLS = 66;
constexpr int N_SCAN_MAX_R = 4096;

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}"
      ::"r"(bar), "r"(parity) : "memory");
}
// box (32 channels, W, Hb, 1) of the [B,H,W,C] map at (c0, 0, h0, b0) -> dst
__device__ __forceinline__ void tma_load_band(unsigned dst, const CUtensorMap* tmap, int c0, int h0,
                                              int b0, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(tmap)), "r"(c0), "r"(0), "r"(h0),
        "r"(b0), "r"(bar) : "memory");
}

// the same box -> L2 only (issued one slice ahead, so the copy above finds the band in L2)
__device__ __forceinline__ void tma_prefetch_band(const CUtensorMap* tmap, int c0, int h0, int b0) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(reinterpret_cast<unsigned long long>(tmap)), "r"(c0), "r"(0), "r"(h0), "r"(b0)
               : "memory");
}

// programmatic dependent launch (no-ops in a kernel launched without the attribute)
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ int edge_lo(int mode, int p, float bin) {
  const float v = __fmul_rn((float)p, bin);
  return mode == WSSDL_BIN_CPU_TRUNC ? (int)v : (int)floorf(v);   // cc:167-170 / gpu.cu.cc:51-58
}
__device__ __forceinline__ int edge_hi(int mode, int p, float bin) {
  const float v = __fmul_rn((float)(p + 1), bin);
  return mode == WSSDL_BIN_CPU_TRUNC ? (int)v : (int)ceilf(v);
}

// First-maximum update (strict '>', cc:187) with both conditional moves on the FMA pipes (see
// upd_fma in roi_pool_dev.cuh: FSETP / FSEL / SEL all issue to the half-rate ALU pipe).  The
// value moves as `@p FMUL m, v, 1.0f`.  The index moves as `@p IMAD mi, bits(v), 0, t` where t is
// the finished flat index of the cell for this lane's chunk (computed once per cell) and the zero
// comes from shared memory, opaque to ptxas: every channel's IMAD has a source of its own (its
// value), so ptxas cannot merge the eight moves of a cell into one product plus eight SELs as it
// does for `mad mi, key, kmul, add` with common operands.
template <bool HAS_ARGMAX>
__device__ __forceinline__ void upd_idx(float v, int t, float& m, int& mi, float one_f, int zero_i) {
  if (HAS_ARGMAX)
    asm("{\n\t.reg .pred p;\n\t"
        "setp.gt.f32 p, %2, %0;\n\t"
        "@p mul.rn.f32 %0, %2, %4;\n\t"
        "@p mad.lo.s32 %1, %3, %5, %6;\n\t}"
        : "+f"(m), "+r"(mi)
        : "f"(v), "r"(__float_as_int(v)), "f"(one_f), "r"(zero_i), "r"(t));
  else
    asm("{\n\t.reg .pred p;\n\t"
        "setp.gt.f32 p, %1, %0;\n\t"
        "@p mul.rn.f32 %0, %1, %2;\n\t}"
        : "+f"(m)
        : "f"(v), "f"(one_f));
}

struct Acc {
  float m[8];
  int mi[8];
};

// channels 0..3 of a lane belong to chunk A, 4..7 to chunk B of a cell; mi[k] carries the flat
// index minus the channel's own offset k & 3 (added back when the bin is stored), so one finished
// index per chunk (t_a, t_b) serves four channels; "nothing pooled" is -1 - (k & 3)
template <int OFF>
__device__ __forceinline__ float4 lds128(unsigned addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+%5];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr), "n"(OFF));
  return v;
}
__device__ __forceinline__ unsigned lds32(unsigned addr) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ unsigned lds8(unsigned addr) {
  unsigned v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

struct IdxK {
  float one_f;
  int zero_i, kmul, add_a, add_b;
};

// one cell: t_a / t_b = flat index (minus the channel's k & 3) of the cell for chunk A / B
template <bool HAS_ARGMAX>
__device__ __forceinline__ void upd8(Acc& a, const float4 v0, const float4 v1, int key, const IdxK& x) {
  const int t_a = HAS_ARGMAX ? key * x.kmul + x.add_a : 0;
  const int t_b = HAS_ARGMAX ? key * x.kmul + x.add_b : 0;
  upd_idx<HAS_ARGMAX>(v0.x, t_a, a.m[0], a.mi[0], x.one_f, x.zero_i);
  upd_idx<HAS_ARGMAX>(v0.y, t_a, a.m[1], a.mi[1], x.one_f, x.zero_i);
  upd_idx<HAS_ARGMAX>(v0.z, t_a, a.m[2], a.mi[2], x.one_f, x.zero_i);
  upd_idx<HAS_ARGMAX>(v0.w, t_a, a.m[3], a.mi[3], x.one_f, x.zero_i);
  upd_idx<HAS_ARGMAX>(v1.x, t_b, a.m[4], a.mi[4], x.one_f, x.zero_i);
  upd_idx<HAS_ARGMAX>(v1.y, t_b, a.m[5], a.mi[5], x.one_f, x.zero_i);
  upd_idx<HAS_ARGMAX>(v1.z, t_b, a.m[6], a.mi[6], x.one_f, x.zero_i);
  upd_idx<HAS_ARGMAX>(v1.w, t_b, a.m[7], a.mi[7], x.one_f, x.zero_i);
}

// The first cell of a bin.  Pooling starts from (-FLT_MAX, -1) (cc:180-182) and a value only
// wins with a strict '>', so the first cell is simply taken over when all of its eight values are
// > -FLT_MAX (no NaN, no -inf): the accumulators are the loaded registers themselves, the indices
// one move each.  Otherwise the general update runs from the initial state.
template <bool HAS_ARGMAX>
__device__ __forceinline__ void first_cell(Acc& a, const float4 v0, const float4 v1, int key,
                                           const IdxK& x) {
  const bool ok = v0.x > -FLT_MAX && v0.y > -FLT_MAX && v0.z > -FLT_MAX && v0.w > -FLT_MAX &&
                  v1.x > -FLT_MAX && v1.y > -FLT_MAX && v1.z > -FLT_MAX && v1.w > -FLT_MAX;
  if (ok) {
    a.m[0] = v0.x; a.m[1] = v0.y; a.m[2] = v0.z; a.m[3] = v0.w;
    a.m[4] = v1.x; a.m[5] = v1.y; a.m[6] = v1.z; a.m[7] = v1.w;
    if (HAS_ARGMAX) {
      const int t_a = key * x.kmul + x.add_a, t_b = key * x.kmul + x.add_b;
#pragma unroll
      for (int k = 0; k < 4; ++k) { a.mi[k] = t_a; a.mi[4 + k] = t_b; }
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) { a.m[k] = -FLT_MAX; a.mi[k] = -1 - (k & 3); }
    upd8<HAS_ARGMAX>(a, v0, v1, key, x);
  }
}

// Cells W0..NW-1 of the row at q (chunk A of the row's first cell; chunk B at qx = q ^ 16), one
// cell (8 registers) in flight at a time: with 1024 threads per SM a thread has 64 registers, and
// a second cell in flight made ptxas rematerialise the thread constants in every bin.
// k0 = the row's first key.
template <int W0, int NW, bool HAS_ARGMAX, int KS>
__device__ __forceinline__ void pool_row(Acc& a, unsigned q, unsigned qx, int k0, const IdxK& x) {
#if WSSDL_BINS_INFLIGHT == 1
  if (W0 <= 0 && NW >= 1) { const float4 v0 = lds128<0>(q), v1 = lds128<0>(qx); upd8<HAS_ARGMAX>(a, v0, v1, k0, x); }
  if (W0 <= 1 && NW >= 2) { const float4 v0 = lds128<128>(q), v1 = lds128<128>(qx); upd8<HAS_ARGMAX>(a, v0, v1, k0 + KS, x); }
  if (W0 <= 2 && NW >= 3) { const float4 v0 = lds128<256>(q), v1 = lds128<256>(qx); upd8<HAS_ARGMAX>(a, v0, v1, k0 + 2 * KS, x); }
  if (W0 <= 3 && NW >= 4) { const float4 v0 = lds128<384>(q), v1 = lds128<384>(qx); upd8<HAS_ARGMAX>(a, v0, v1, k0 + 3 * KS, x); }
  if (W0 <= 4 && NW >= 5) { const float4 v0 = lds128<512>(q), v1 = lds128<512>(qx); upd8<HAS_ARGMAX>(a, v0, v1, k0 + 4 * KS, x); }
  if (W0 <= 5 && NW >= 6) { const float4 v0 = lds128<640>(q), v1 = lds128<640>(qx); upd8<HAS_ARGMAX>(a, v0, v1, k0 + 5 * KS, x); }
  if (W0 <= 6 && NW >= 7) { const float4 v0 = lds128<768>(q), v1 = lds128<768>(qx); upd8<HAS_ARGMAX>(a, v0, v1, k0 + 6 * KS, x); }
#else
  if (W0 <= 0 && NW >= 1) {
    const float4 a0 = lds128<0>(q), b0 = lds128<0>(qx);
    if (W0 <= 1 && NW >= 2) {
      const float4 a1 = lds128<128>(q), b1 = lds128<128>(qx);
      upd8<HAS_ARGMAX>(a, a0, b0, k0, x);
      upd8<HAS_ARGMAX>(a, a1, b1, k0 + KS, x);
    } else {
      upd8<HAS_ARGMAX>(a, a0, b0, k0, x);
    }
  } else if (W0 <= 1 && NW >= 2) {
    const float4 a1 = lds128<128>(q), b1 = lds128<128>(qx);
    upd8<HAS_ARGMAX>(a, a1, b1, k0 + KS, x);
  }
  if (NW >= 3) {
    const float4 a2 = lds128<256>(q), b2 = lds128<256>(qx);
    if (NW >= 4) {
      const float4 a3 = lds128<384>(q), b3 = lds128<384>(qx);
      upd8<HAS_ARGMAX>(a, a2, b2, k0 + 2 * KS, x);
      upd8<HAS_ARGMAX>(a, a3, b3, k0 + 3 * KS, x);
    } else {
      upd8<HAS_ARGMAX>(a, a2, b2, k0 + 2 * KS, x);
    }
  }
  if (NW >= 5) {
    const float4 a4 = lds128<512>(q), b4 = lds128<512>(qx);
    if (NW >= 6) {
      const float4 a5 = lds128<640>(q), b5 = lds128<640>(qx);
      upd8<HAS_ARGMAX>(a, a4, b4, k0 + 4 * KS, x);
      upd8<HAS_ARGMAX>(a, a5, b5, k0 + 5 * KS, x);
    } else {
      upd8<HAS_ARGMAX>(a, a4, b4, k0 + 4 * KS, x);
    }
  }
  if (NW >= 7) {
    const float4 a6 = lds128<768>(q), b6 = lds128<768>(qx);
    upd8<HAS_ARGMAX>(a, a6, b6, k0 + 6 * KS, x);
  }
#endif
}

// nh rows of NW cells starting at shared address q (chunk A of the first cell).  key = the
// value the index is built from: the chunk-A address (LINEAR) or the cell's index in the image
// (h*W+w); it advances by 128 / 1 per cell and by krow per row.
template <int NW, bool HAS_ARGMAX, bool LINEAR>
__device__ __forceinline__ void pool_rows(Acc& a, unsigned q, int key, int nh, unsigned row_bytes,
                                          int krow, const IdxK& x) {
  unsigned qx = q ^ 16u;
  constexpr int KS = LINEAR ? 128 : 1;
  {
    const float4 a0 = lds128<0>(q), b0 = lds128<0>(qx);
    first_cell<HAS_ARGMAX>(a, a0, b0, LINEAR ? (int)q : key, x);
    pool_row<1, NW, HAS_ARGMAX, KS>(a, q, qx, LINEAR ? (int)q : key, x);
  }
#pragma unroll 1
  for (int r = nh - 1; r > 0; --r) {
    q += row_bytes;
    qx += row_bytes;
    if (!LINEAR) key += krow;
    pool_row<0, NW, HAS_ARGMAX, KS>(a, q, qx, LINEAR ? (int)q : key, x);
  }
}

// any nh x nw (per-lane bounds allowed: the loops may diverge); nh or nw may be <= 0
template <bool HAS_ARGMAX, bool LINEAR>
__device__ __forceinline__ void pool_rows_any(Acc& a, unsigned q, int key, int nh, int nw,
                                              unsigned row_bytes, int krow, const IdxK& x) {
#pragma unroll
  for (int k = 0; k < 8; ++k) { a.m[k] = -FLT_MAX; a.mi[k] = -1 - (k & 3); }   // cc:180-182
#pragma unroll 1
  for (int r = nh; r > 0; --r) {
    unsigned qq = q;
    int kk = key;
#pragma unroll 1
    for (int w = nw; w > 0; --w) {
      const float4 v0 = lds128<0>(qq), v1 = lds128<0>(qq ^ 16u);
      upd8<HAS_ARGMAX>(a, v0, v1, LINEAR ? (int)qq : kk, x);
      qq += 128u;
      if (!LINEAR) kk += 1;
    }
    q += row_bytes;
    if (!LINEAR) key += krow;
  }
}

// ------------------------------------------------------------------------------------------
// Pre-pass: one CTA per image counting-sorts the image's bins by (owner band, size class) into
// the workspace.  Record of a bin (8 B): x = its output position n*PH*PW + ph*PW + pw, y = its
// first cell inside the owner band's resident rows (11 bits) | pad flag (bit 11) | size class
// << 12.  Classes are padded to whole warps (8 bins), so the 8 records of a group share a class.
// table[img*NB_MAX + band] = {first group, groups, of which special (leading), of which empty
// (trailing)}.
constexpr int S_THREADS = 1024;
constexpr int NB_MAX = 4;
constexpr int PAD_PER_IMAGE = NB_MAX * N_CLS * 8 + 8;   // per (image, sub-list)
constexpr int S_MAX = 4;                // sub-lists per image (sort CTAs per image)
constexpr unsigned RECY_PAD = 1u << 11;

// special = pooled by the general per-lane code of the pooling kernel
__host__ __device__ __forceinline__ bool class_is_special(int c) {
  return c == CLS_SLOW || (c < CLS_EMPTY && ((c >> 3) == N_DIM - 1 || (c & 7) == N_DIM - 1));
}
// position of a class in a band's list: specials first, uniform classes heaviest first, empty last
__host__ __device__ __forceinline__ int class_rank(int c) {
  if (c == CLS_SLOW) return 0;
  if (c == CLS_EMPTY) return 2 * N_CLS;
  return (class_is_special(c) ? 0 : N_CLS) + (N_CLS - 1 - c);
}

struct SortArgs {
  const float* rois;
  const int* perm;
  const int* img_start;
  uint2* recs;
  int4* table;
  int B, H, W, R, PH, PW;
  float spatial_scale;
  int bin_mode;
  int rch;              // RoIs whose geometry is resident at a time
  int stride;           // > 0: image-major RoIs, image b owns rows [b*stride, (b+1)*stride)
  int nsub;             // CTAs per image: CTA (img, sub) sorts the sub-th part of the image's RoIs into
                        // a list of its own (table entry (img, band, sub))
  BandGeom bg;
  FastDiv divPH, divStep, divPW, divPHPW;
};

__global__ void __launch_bounds__(S_THREADS, 1) roi_bin_sort_kernel(const SortArgs a) {
  extern __shared__ __align__(16) unsigned char k_smem[];
  __shared__ int s_hist[NB_MAX][N_CLS], s_start[NB_MAX][N_CLS], s_cur[NB_MAX][N_CLS];
  __shared__ int s_count, s_before;
  const int tid = threadIdx.x;
  const int H = a.H, W = a.W, PH = a.PH, PW = a.PW, NB = a.bg.NB;
  const int img = blockIdx.x / a.nsub;              // == B: RoIs with no valid image
  const int sub = blockIdx.x - img * a.nsub;
  // the pooling kernel may become resident (and stage its bands) now; this kernel itself may
  // have been launched the same way behind the kernel that writes the RoIs
  griddep_launch_dependents();
  griddep_wait();
  const bool valid_img = img < a.B;
  const int mode = a.bin_mode;

  unsigned long long* s_wf = reinterpret_cast<unsigned long long*>(k_smem);
  int* s_sh = reinterpret_cast<int*>(s_wf + a.rch);            // start_h
  float* s_bh = reinterpret_cast<float*>(s_sh + a.rch);        // bin_h
  int* s_nb = reinterpret_cast<int*>(s_bh + a.rch);            // RoI index x PH*PW
  unsigned short* s_we = reinterpret_cast<unsigned short*>(s_nb + a.rch);   // ws | nw << 8
  unsigned* s_row = reinterpret_cast<unsigned*>(s_we + (((size_t)a.rch * PW + 7) & ~(size_t)7));  // [rch*PH]
  unsigned short* s_list = reinterpret_cast<unsigned short*>(s_row + (size_t)a.rch * PH);   // scan mode only

  // ---- this image's RoIs, and how many RoIs the images before it hold
  const int* list = nullptr;
  int n_img, before;
  if (a.stride > 0) {                               // grouped: no lists at all
    before = img * a.stride;
    n_img = valid_img ? a.stride : 0;
  } else if (a.perm != nullptr) {
    before = a.img_start[img];
    n_img = a.img_start[img + 1] - before;
    list = a.perm + before;
  } else {
    if (tid == 0) { s_count = 0; s_before = 0; }
    __syncthreads();
    int less = 0;
    for (int r = tid; r < a.R; r += S_THREADS) {
      const int b = roi_bucket(__ldg(a.rois + (size_t)r * 5), a.B);
      if (b == img) s_list[atomicAdd(&s_count, 1)] = (unsigned short)r;
      less += b < img;
    }
    if (less) atomicAdd(&s_before, less);
    __syncthreads();
    n_img = s_count;
    before = s_before;
  }
  for (int i = tid; i < NB_MAX * N_CLS; i += S_THREADS) (&s_hist[0][0])[i] = 0;
  {                                                 // this CTA's part of the image's RoIs
    const int lo = (int)((long long)n_img * sub / a.nsub);
    const int hi = (int)((long long)n_img * (sub + 1) / a.nsub);
    before += lo;
    if (list) list += lo;
    n_img = hi - lo;
  }
  if (n_img == 0) {
    if (tid < NB_MAX) a.table[(img * NB_MAX + tid) * S_MAX + sub] = make_int4(0, 0, 0, 0);
    return;
  }
  const size_t rec0 = (((size_t)before * PH * PW + 7) & ~(size_t)7) +
                      (size_t)(img * a.nsub + sub) * PAD_PER_IMAGE;

  // RoI geometry with the reference's float expressions (cc:153-176); the width histogram of a
  // RoI's PW bins (9 fields of 7 bits: 0, 1..7, >= 8 cells) is shared by its PH rows
  auto geometry = [&](int c0, int nb) {
    for (int rl = tid; rl < nb; rl += S_THREADS) {
      const int n = a.stride > 0 ? before + c0 + rl : (list ? list[c0 + rl] : (int)s_list[c0 + rl]);
      s_nb[rl] = n * PH * PW;
      const RoiCells g = roi_cells(a.rois + (size_t)n * 5, a.spatial_scale, PH, PW);
      // grouped: a row whose batch index is not its block's image (the padding rows behind an
      // image's RoIs carry -1) has empty bins only
      const bool mine = a.stride <= 0 || roi_bucket(__ldg(a.rois + (size_t)n * 5), a.B) == img;
      s_sh[rl] = mine ? g.start_h : 0;
      s_bh[rl] = mine ? g.bin_h : 0.f;
      unsigned long long wf = 0;
      for (int pw = 0; pw < PW; ++pw) {
        int ws = min(max(edge_lo(mode, pw, g.bin_w) + g.start_w, 0), W);
        int we = min(max(edge_hi(mode, pw, g.bin_w) + g.start_w, 0), W);
        if (!valid_img || !mine) ws = we = 0;
        const int nw = max(we - ws, 0);
        s_we[rl * PW + pw] = (unsigned short)(ws | (min(nw, 255) << 8));
        wf += 1ull << (7 * min(nw, N_DIM));
      }
      s_wf[rl] = wf;
    }
  };
  // one task = the PW bins of (RoI, ph): owner band (hs is non-decreasing in ph, so a band owns
  // a contiguous ph range of every RoI), rows, and whether all rows are resident in that band
  struct Row { int band, hs, nh; bool slow; };
  auto row_of = [&](int rl, int ph) {
    Row r;
    int hs = min(max(edge_lo(mode, ph, s_bh[rl]) + s_sh[rl], 0), H);
    int he = min(max(edge_hi(mode, ph, s_bh[rl]) + s_sh[rl], 0), H);
    if (!valid_img) hs = he = 0;
    r.band = min((int)fastdiv((unsigned)hs, a.divStep), NB - 1);
    r.hs = hs;
    r.nh = he - hs;
    r.slow = he > min(r.band * a.bg.step + a.bg.Hb, H);
    return r;
  };

  // a (RoI, ph) row as one word for the per-bin pass: hs | min(nh, 8) << 16 | band << 20 | slow << 22
  auto pack_row = [&](const Row& r) {
    return (unsigned)r.hs | ((unsigned)min(max(r.nh, 0), N_DIM) << 16) | ((unsigned)r.band << 20) |
           ((unsigned)(r.slow ? 1 : 0) << 22);
  };

  // ---- pass A: class histogram per band
  for (int c0 = 0; c0 < n_img; c0 += a.rch) {
    const int nb = min(a.rch, n_img - c0);
    __syncthreads();
    geometry(c0, nb);
    __syncthreads();
    for (int t = tid; t < nb * PH; t += S_THREADS) {
      const int rl = (int)fastdiv((unsigned)t, a.divPH);
      const int ph = t - rl * PH;
      const Row r = row_of(rl, ph);
      s_row[t] = pack_row(r);
      if (r.nh <= 0) { atomicAdd(&s_hist[r.band][CLS_EMPTY], PW); continue; }
      const int crow = (min(r.nh, N_DIM) - 1) * N_DIM - 1;
      const unsigned long long wf = s_wf[rl];
#pragma unroll
      for (int k = 0; k <= N_DIM; ++k) {
        const int cnt = (int)((wf >> (7 * k)) & 127u);
        if (cnt) atomicAdd(&s_hist[r.band][k == 0 ? CLS_EMPTY : (r.slow ? CLS_SLOW : crow + k)], cnt);
      }
    }
  }
  __syncthreads();
  // ---- class order inside a band: slow bins, then the oversized classes (8 or more rows or
  // cells per row: exact sizes per lane), then the uniform classes heaviest first, empty bins
  // last; every class padded to whole warps (8 bins); bands one after the other
  if (tid < NB_MAX * N_CLS) {
    const int b = tid / N_CLS, c = tid - b * N_CLS;
    const int my_rank = class_rank(c);
    int before_c = 0, band_total = 0, bands_before = 0, special = 0;
    for (int bb = 0; bb < b; ++bb)
      for (int cc = 0; cc < N_CLS; ++cc) bands_before += (s_hist[bb][cc] + 7) & ~7;
    for (int cc = 0; cc < N_CLS; ++cc) {
      const int p8 = (s_hist[b][cc] + 7) & ~7;
      band_total += p8;
      if (class_rank(cc) < my_rank) before_c += p8;
      if (class_is_special(cc)) special += p8;
    }
    s_start[b][c] = bands_before + before_c;
    s_cur[b][c] = bands_before + before_c;
    if (c == 0)
      a.table[(img * NB_MAX + b) * S_MAX + sub] = make_int4((int)((rec0 + bands_before) >> 3), band_total >> 3,
                                            special >> 3, ((s_hist[b][CLS_EMPTY] + 7) & ~7) >> 3);
  }
  __syncthreads();
  // ---- pass B: scatter the records
  uint2* const recs = a.recs + rec0;
  for (int c0 = 0; c0 < n_img; c0 += a.rch) {
    const int nb = min(a.rch, n_img - c0);
    if (n_img > a.rch) {                            // (one batch: its geometry is still resident)
      __syncthreads();
      geometry(c0, nb);
      __syncthreads();
      for (int t = tid; t < nb * PH; t += S_THREADS) {
        const int rl = (int)fastdiv((unsigned)t, a.divPH);
        s_row[t] = pack_row(row_of(rl, t - rl * PH));
      }
      __syncthreads();
    }
    // one thread per bin (per (RoI, ph) row and class the loops were 60 % of this kernel, most
    // of their iterations skipping): the lanes of a warp that hold bins of the same (band, class)
    // take their slots with ONE shared-memory atomic per run of adjacent lanes, in lane order -- the bins of a row that
    // share a class stay next to each other, so a warp's output rows stay contiguous
    const int nbins = nb * PH * PW;
    for (int t0 = 0; t0 < nbins; t0 += S_THREADS) {
      const int t = t0 + tid;
      const bool act = t < nbins;
      int key = -1 - (tid & 31);                    // idle lanes: a key of their own
      unsigned recx = 0, recy = 0;
      if (act) {
        const int rl = (int)fastdiv((unsigned)t, a.divPHPW);
        const int rem = t - rl * PH * PW;
        const int ph = (int)fastdiv((unsigned)rem, a.divPW);
        const int pw = rem - ph * PW;
        const unsigned rw = s_row[rl * PH + ph];
        const int nh8 = (int)((rw >> 16) & 15u), band = (int)((rw >> 20) & 3u), hs = (int)(rw & 0xffffu);
        const unsigned e = s_we[rl * PW + pw];
        const int nw = (int)(e >> 8);
        int cls;
        if (nh8 == 0 || nw == 0) cls = CLS_EMPTY;
        else if (rw >> 22) cls = CLS_SLOW;
        else cls = (nh8 - 1) * N_DIM - 1 + min(nw, N_DIM);
        key = band * N_CLS + cls;
        const unsigned cell =
            cls < CLS_EMPTY ? (unsigned)((hs - band * a.bg.step) * W + (int)(e & 255u)) : 0u;
        recx = (unsigned)(s_nb[rl] + rem);
        recy = cell | ((unsigned)cls << 12);
      }
      // runs of adjacent lanes with the same key (match.any costs several hundred cycles here)
      const int ln = tid & 31;
      const int prev = __shfl_up_sync(0xffffffffu, key, 1);
      const unsigned heads = __ballot_sync(0xffffffffu, ln == 0 || key != prev);
      const unsigned upto = (2u << ln) - 1u;          // lanes 0..ln (ln = 31: all)
      const int my_head = 31 - __clz((int)(heads & upto));
      const unsigned above = heads & ~upto;
      const int run_end = above ? __ffs((int)above) - 1 : 32;
      int base = 0;
      if (act && ln == my_head) base = atomicAdd(&(&s_cur[0][0])[key], run_end - ln);
      base = __shfl_sync(0xffffffffu, base, my_head);
      if (act) recs[base + (ln - my_head)] = make_uint2(recx, recy);
    }
  }
  if (tid < NB_MAX * N_CLS) {
    const int b = tid / N_CLS, c = tid - b * N_CLS;
    const int n = s_hist[b][c];
    for (int p = s_start[b][c] + n; p < s_start[b][c] + ((n + 7) & ~7); ++p)
      recs[p] = make_uint2(0u, RECY_PAD | ((unsigned)c << 12));
  }
}

// ------------------------------------------------------------------------------------------
// Pooling: one CTA = (image, row band, range of the band's groups, sg 32-channel slices).
struct PoolArgs {
  const float* bottom;
  const float* rois;
  const uint2* recs;
  const int4* table;
  float* top;
  int* argmax;
  int B, H, W, C, PH, PW;
  float spatial_scale;
  int bin_mode;
  int nchunks, sg, n_slices, nsub;
  int rec_cap;          // records resident at a time (multiple of 8)
  BandGeom bg;
  FastDiv divPW, divPHPW;
  Ones ones;            // f = 1.0f, i[0] = C/128 (LINEAR) or C, i[1] = 0
};

// edges of one bin recomputed from its RoI (slow / oversized bins only)
struct BinEdges { int hs, he, ws, nw; };
__device__ __noinline__ BinEdges bin_edges(const float* rois, unsigned out_bin, FastDiv divPHPW,
                                           FastDiv divPW, int PH, int PW, int H, int W, float scale,
                                           int mode) {
  const unsigned n = fastdiv(out_bin, divPHPW);
  const unsigned binid = out_bin - n * (unsigned)(PH * PW);
  const int ph = (int)fastdiv(binid, divPW);
  const int pw = (int)binid - ph * PW;
  const RoiCells g = roi_cells(rois + (size_t)n * 5, scale, PH, PW);
  BinEdges e;
  e.hs = min(max(edge_lo(mode, ph, g.bin_h) + g.start_h, 0), H);
  e.he = min(max(edge_hi(mode, ph, g.bin_h) + g.start_h, 0), H);
  e.ws = min(max(edge_lo(mode, pw, g.bin_w) + g.start_w, 0), W);
  const int we = min(max(edge_hi(mode, pw, g.bin_w) + g.start_w, 0), W);
  e.nw = max(we - e.ws, 0);
  return e;
}

__device__ __forceinline__ void bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint2 lds64(unsigned addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}

template <int NT, bool HAS_ARGMAX, bool LINEAR, bool USE_TMA>
__global__ void __launch_bounds__(NT, 1024 / NT)
roi_pool_fwd_bins_kernel(const __grid_constant__ CUtensorMap tmap, const PoolArgs a) {
  extern __shared__ __align__(128) unsigned char n_smem[];
  __shared__ __align__(8) unsigned long long s_bar[2];     // band, records
  __shared__ Ones s_ones;

  const int tid = threadIdx.x, lane = tid & 31;
  const int H = a.H, W = a.W, C = a.C;
  const int band = blockIdx.y % a.bg.NB, chunk = blockIdx.y / a.bg.NB;
  const int img = blockIdx.z;                       // == B: RoIs with no valid image
  const bool valid_img = img < a.B;

  if (!valid_img && band != 0) return;              // their (empty) bins all sit in band 0
  const int row0 = band * a.bg.step;                // first resident row
  const int row1 = min(row0 + a.bg.Hb, H);          // one past the last resident row
  const int slice_begin = blockIdx.x * a.sg;
  const int slice_end = min(slice_begin + a.sg, a.n_slices);

  // cells 128 B aligned: the xor-16 chunk pairing relies on it
  const unsigned raw_u32 = (unsigned)__cvta_generic_to_shared(n_smem);
  unsigned char* base = n_smem + ((128u - (raw_u32 & 127u)) & 127u);
  float4* s_map = reinterpret_cast<float4*>(base);
  const unsigned band_bytes = (unsigned)a.bg.Hb * (unsigned)W * 128u;
  const unsigned map_u32 = (unsigned)__cvta_generic_to_shared(s_map);
  const unsigned rec_u32 = map_u32 + band_bytes;    // uint2[rec_cap]
  const unsigned bar_band = (unsigned)__cvta_generic_to_shared(&s_bar[0]);
  const unsigned bar_rec = (unsigned)__cvta_generic_to_shared(&s_bar[1]);
  unsigned parity = 0, parity_rec = 0;
  if (tid == 0) {
    s_ones = a.ones;
    mbar_init(bar_band, 1);
    mbar_init(bar_rec, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // staging of one slice of the band: one TMA tensor copy, or cp.async when no tensor map
  auto stage = [&](int slice) {
    if (!valid_img) return;
    if (USE_TMA) {
      if (tid == 0) {
        mbar_expect_tx(bar_band, band_bytes);
        tma_load_band(map_u32, &tmap, slice * N_SLICE, row0, img, bar_band);
      }
    } else {
      const int CV = C >> 2;
      const float4* src = reinterpret_cast<const float4*>(a.bottom + ((size_t)img * H + row0) * W * C) +
                          slice * (N_SLICE / 4);
      const int n4 = (row1 - row0) * W * (N_SLICE / 4);
      for (int i = tid; i < n4; i += NT) cp_async16(s_map + i, src + (size_t)(i >> 3) * CV + (i & 7));
    }
  };
  auto stage_wait = [&]() {
    if (!valid_img) return;
    if (USE_TMA) {
      mbar_wait(bar_band, parity);
      parity ^= 1u;
    } else {
      cp_async_wait_all();
      __syncthreads();
    }
  };
  // The band does not depend on the pre-pass: its copy is in flight before this kernel waits
  // for the bin-sort kernel (programmatic dependent launch: the CTAs of the first wave are
  // resident, barriers initialised and bands staged while the pre-pass still runs).
  stage(slice_begin);
  griddep_wait();
  // chunk = (sub-list, range of its groups): nchunks = nsub * ranges per sub-list
  const int per_sub = a.nchunks / a.nsub;
  const int sub = chunk / per_sub, part = chunk - sub * per_sub;
  const int4 tb = a.table[(img * NB_MAX + band) * S_MAX + sub];
  const int g_begin = (int)((long long)tb.y * part / per_sub);
  const int g_count = (int)((long long)tb.y * (part + 1) / per_sub) - g_begin;
  if (g_count <= 0) {
    stage_wait();                                   // (no exit with a copy into this CTA in flight)
    return;
  }
  // the records of groups [gb, gb + n) of this CTA's range: one bulk copy
  const uint2* const recs_cta = a.recs + (size_t)(tb.x + g_begin) * 8;
  auto stage_recs = [&](int gb, int n) {
    if (tid == 0) {
      mbar_expect_tx(bar_rec, (unsigned)n * 64u);
      bulk_load(rec_u32, recs_cta + (size_t)gb * 8, (unsigned)n * 64u, bar_rec);
    }
  };
  const int cap_g = a.rec_cap >> 3;
  stage_recs(0, min(cap_g, g_count));

  // opaque unit operands of the conditional moves (see the tiled kernel in roi_pool.cu)
  const float one_f = *reinterpret_cast<volatile float*>(&s_ones.f);
  const int kmul = *reinterpret_cast<volatile int*>(&s_ones.i[0]);
  const int zero_i = *reinterpret_cast<volatile int*>(&s_ones.i[1]);

  // per-thread constants: a warp always holds 8 whole bins, so the quarter j and the bin role
  // sw of a lane never change
  const int j = lane & 3;                           // channels [8j, 8j+8) of the slice
  const int sw = (lane >> 2) & 1;                   // bin A (0) or B (1) of its quarter warp
  const unsigned row_bytes = (unsigned)W * 128u;
  // registers m[0..3] hold chunk 2j+sw, m[4..7] chunk 2j+1-sw
  const unsigned map_a = map_u32 + (unsigned)(2 * j + sw) * 16u;
  // this warp's index as a warp-uniform value (the compiler cannot prove tid >> 5 uniform)
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const float* img_base = a.bottom + (size_t)(valid_img ? img : 0) * H * W * C;
  const unsigned rec_lane = rec_u32 + (unsigned)(lane >> 2) * 8u;
  const unsigned C4 = (unsigned)C * 4u;
  constexpr int NWARP = NT / 32;

  // values the pooling loops keep in registers: made opaque so that ptxas does not rebuild them
  // from kernel parameters and special registers in every bin (it did: S2R / LDCU per bin)
  unsigned map_a_r = map_a, rec_lane_r = rec_lane;
  int add_sw = 8 * sw - 4;                          // add_a - add_b
  asm volatile("" : "+r"(map_a_r), "+r"(rec_lane_r), "+r"(add_sw));

  // The groups are handed out to the warps in snake order (round r: warp w takes group
  // NWARP*r + w, the next round NWARP*r + 2*NWARP-1 - w, ...): the list is ordered heaviest class
  // first, so plain round robin would give warp 0 the heaviest group of EVERY round.  A warp
  // walks its groups in three loops -- special (slow / oversized bins, general code), uniform
  // classes (the hot loop: no calls, no per-lane sizes), empty bins -- and reads the record of
  // its next group from shared memory before it pools the current one.
  for (int gb = 0; gb < g_count; gb += cap_g) {
    int ngran = min(cap_g, g_count - gb);
    // segment ends inside this round of records (in groups, relative to gb)
    int end_special = min(max(tb.z - (g_begin + gb), 0), ngran);
    int end_uniform = min(max(tb.y - tb.w - (g_begin + gb), 0), ngran);
    asm volatile("" : "+r"(ngran), "+r"(end_special), "+r"(end_uniform));
    mbar_wait(bar_rec, parity_rec);
    parity_rec ^= 1u;
    for (int slice = slice_begin; slice < slice_end; ++slice) {
      if (USE_TMA && tid == 0 && valid_img && slice + 1 < slice_end)
        tma_prefetch_band(&tmap, (slice + 1) * N_SLICE, row0, img);   // next band -> L2
      stage_wait();
      const int c_thr = slice * N_SLICE + j * 8;
      // flat argmax of the cell whose chunk A sits at shared address q (LINEAR, C % 128 == 0):
      //   ((q - map_a)/128 + row0*W) * C + c = q * (C/128) + [row0*W*C - map_a * (C/128) + c];
      // otherwise key = h*W + w and flat = key * C + c
      const int lin = LINEAR ? row0 * W * C - (int)(map_a * (unsigned)kmul) : 0;
      IdxK x;
      x.one_f = one_f; x.zero_i = zero_i; x.kmul = kmul;
      x.add_a = lin + c_thr + 4 * sw;
      x.add_b = lin + c_thr + 4 * (1 - sw);
      unsigned c4 = (unsigned)c_thr * 4u;
      asm volatile("" : "+r"(x.add_a), "+r"(c4));

      int g = warp_u;
      int step = 2 * NWARP - 1 - 2 * warp_u;        // then 1 + 2 * warp, alternating (sum 2 * NWARP)
      uint2 rec = make_uint2(0u, RECY_PAD);
      if (g < ngran) rec = lds64(rec_lane_r + (unsigned)g * 64u);

      // outputs of a lane: top / argmax + out_bin * C + c_thr
#define WSSDL_OUT_PTRS(REC)                                                                     \
      unsigned long long off_;                                                                  \
      asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(off_) : "r"((REC).x), "r"(C4), "l"((unsigned long long)c4)); \
      float* const top_p = reinterpret_cast<float*>(reinterpret_cast<char*>(a.top) + off_);     \
      int* const arg_p = reinterpret_cast<int*>(reinterpret_cast<char*>(a.argmax) + (HAS_ARGMAX ? off_ : 0))
#define WSSDL_STORE_ACC(ACC)                                                                    \
      do {                                                                                      \
        const bool swapped = add_sw > 0;            /* lanes of bin B: chunks in swapped registers */ \
        const float4 ta = make_float4((ACC).m[0], (ACC).m[1], (ACC).m[2], (ACC).m[3]);          \
        const float4 tb4 = make_float4((ACC).m[4], (ACC).m[5], (ACC).m[6], (ACC).m[7]);         \
        if (swapped) st256<true>(top_p, tb4, ta);                                               \
        else st256<true>(top_p, ta, tb4);                                                       \
        if (HAS_ARGMAX) {                                                                       \
          const int4 aa = make_int4((ACC).mi[0], (ACC).mi[1] + 1, (ACC).mi[2] + 2, (ACC).mi[3] + 3); \
          const int4 ab = make_int4((ACC).mi[4], (ACC).mi[5] + 1, (ACC).mi[6] + 2, (ACC).mi[7] + 3); \
          if (swapped) st256<true>(arg_p, ab, aa);                                              \
          else st256<true>(arg_p, aa, ab);                                                      \
        }                                                                                       \
      } while (0)

      // ---- special groups: slow bins (rows outside the band: global memory) and oversized ones
      while (g < end_special) {
        const bool valid = (rec.y & RECY_PAD) == 0;
        const unsigned cls = rec.y >> 12;
        WSSDL_OUT_PTRS(rec);
        BinEdges e = {0, 0, 0, 0};
        if (valid)
          e = bin_edges(a.rois, rec.x, a.divPHPW, a.divPW, a.PH, a.PW, H, W, a.spatial_scale, a.bin_mode);
        if (cls == CLS_SLOW) {
          if (valid) band_slow_bin<HAS_ARGMAX>(img_base, e.hs, e.he, e.ws, e.nw, W, C, c_thr, top_p, arg_p);
        } else {
          const int cell = (int)(rec.y & 2047u);
          Acc acc;
          pool_rows_any<HAS_ARGMAX, LINEAR>(acc, map_a_r + (unsigned)cell * 128u,
                                            LINEAR ? 0 : cell + row0 * W, e.he - e.hs, e.nw, row_bytes, W, x);
          if (valid) WSSDL_STORE_ACC(acc);
        }
        g += step;
        step = 2 * NWARP - step;
        if (g < ngran) rec = lds64(rec_lane_r + (unsigned)g * 64u);
      }
      // ---- uniform classes: the whole warp runs nhc rows of nwc cells (pad lanes pool cell 0 of
      // the band and store nothing)
      while (g < end_uniform) {
        const int g_nxt = g + step;
        step = 2 * NWARP - step;
        uint2 rec_nxt = rec;
        if (g_nxt < ngran) rec_nxt = lds64(rec_lane_r + (unsigned)g_nxt * 64u);
        const unsigned cls = __shfl_sync(0xffffffffu, rec.y, 0) >> 12;
        const int nhc = (int)(cls >> 3) + 1, nwc = (int)(cls & 7u) + 1;
        const int cell = (int)(rec.y & 2047u);
        const unsigned q = map_a_r + (unsigned)cell * 128u;
        const int key = LINEAR ? 0 : cell + row0 * W;
        Acc acc;
        switch (nwc) {
          case 1: pool_rows<1, HAS_ARGMAX, LINEAR>(acc, q, key, nhc, row_bytes, W, x); break;
          case 2: pool_rows<2, HAS_ARGMAX, LINEAR>(acc, q, key, nhc, row_bytes, W, x); break;
          case 3: pool_rows<3, HAS_ARGMAX, LINEAR>(acc, q, key, nhc, row_bytes, W, x); break;
          case 4: pool_rows<4, HAS_ARGMAX, LINEAR>(acc, q, key, nhc, row_bytes, W, x); break;
          case 5: pool_rows<5, HAS_ARGMAX, LINEAR>(acc, q, key, nhc, row_bytes, W, x); break;
          case 6: pool_rows<6, HAS_ARGMAX, LINEAR>(acc, q, key, nhc, row_bytes, W, x); break;
          case 7: pool_rows<7, HAS_ARGMAX, LINEAR>(acc, q, key, nhc, row_bytes, W, x); break;
          default: pool_rows_any<HAS_ARGMAX, LINEAR>(acc, q, key, nhc, nwc, row_bytes, W, x); break;
        }
        if ((rec.y & RECY_PAD) == 0) {
          WSSDL_OUT_PTRS(rec);
          WSSDL_STORE_ACC(acc);
        }
        g = g_nxt;
        rec = rec_nxt;
      }
      // ---- empty bins: (0, -1), cc:180-182
      while (g < ngran) {
        if ((rec.y & RECY_PAD) == 0) {
          WSSDL_OUT_PTRS(rec);
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          st256<true>(top_p, z, z);
          if (HAS_ARGMAX) {
            const int4 n1 = make_int4(-1, -1, -1, -1);
            st256<true>(arg_p, n1, n1);
          }
        }
        g += step;
        step = 2 * NWARP - step;
        if (g < ngran) rec = lds64(rec_lane_r + (unsigned)g * 64u);
      }
#undef WSSDL_STORE_ACC
#undef WSSDL_OUT_PTRS
      const bool more_slices = slice + 1 < slice_end;
      const bool more_recs = gb + cap_g < g_count;
      if (more_slices || more_recs) {
        __syncthreads();                            // every warp is done with the band (and records)
        stage(more_slices ? slice + 1 : slice_begin);
        if (!more_slices) stage_recs(gb + cap_g, min(cap_g, g_count - gb - cap_g));
      }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

int sort_rch(int PH, int PW) {
  const int m = PH > PW ? PH : PW;
  return 24576 / m < 1024 ? 24576 / m : 1024;
}

// s_wf u64[rch] | s_sh, s_bh, s_nb [rch] | s_we u16[rch*PW rounded to 8] | s_row u32[rch*PH] | s_list u16[R]
size_t sort_smem(int rch, int PH, int PW, int R, bool scan) {
  return (size_t)rch * 8 + (size_t)rch * 12 + ((((size_t)rch * PW + 7) & ~(size_t)7) * 2) +
         (size_t)rch * PH * 4 +
         (scan ? align16((size_t)R * 2) : 0) + 16;
}

// workspace: [bucket lists (R > 4096)] | table int4[(B+1)*NB_MAX*S_MAX] | recs uint2[cap]
struct BinsWs {
  size_t off_table, off_recs, total, cap;
};
BinsWs bins_ws(int B, int R, int PH, int PW) {
  BinsWs w;
  const size_t bucket = R > N_SCAN_MAX_R ? align16(bucket_workspace_bytes(B, R)) : 0;
  w.cap = (((size_t)R * PH * PW + 7) & ~(size_t)7) + (size_t)(B + 1) * S_MAX * PAD_PER_IMAGE;
  w.off_table = bucket;
  w.off_recs = w.off_table + align16(sizeof(int4) * (size_t)(B + 1) * NB_MAX * S_MAX);
  w.total = w.off_recs + w.cap * sizeof(uint2) + 16;
  return w;
}

}  // namespace

size_t wssdl_roi::bins_workspace_bytes(int B, int R, int PH, int PW) {
  if (PH <= 0 || PW <= 0 || R <= 0) return 0;
  return bins_ws(B, R, PH, PW).total;
}

BinsPlan wssdl_roi::plan_bins(int B, int H, int W, int C, int R, int PH, int PW, bool aligned,
                              size_t workspace_bytes, int threads) {
  BinsPlan p = {false, false, {1, 0, 0}, 1, 1, 1024, 0, 0, 0, 0};
  if (!aligned || C % N_SLICE != 0 || H > 65535 || W > 255 || B + 1 > 65535) return p;
  if (PH <= 0 || PW <= 0 || PH * PW > 64 || H <= 0 || W <= 0 || R <= 0) return p;
  if ((long long)R * PH * PW >= (1ll << 31)) return p;
  p.scan = R <= N_SCAN_MAX_R;
  if (workspace_bytes < bins_ws(B, R, PH, PW).total) return p;
  p.threads = threads == 512 ? 512 : 1024;
  // shared memory of a pooling CTA: the band's rows of one 32-channel slice
  const size_t budget = (size_t)(p.threads == 512 ? 113 * 1024 : T_DYN_SMEM_MAX) - 128;
  const size_t row_bytes = (size_t)W * 128;
  const long long hb_max = (long long)(budget / row_bytes);
  // the tallest bin of a RoI that lies inside the map, +1 for GPU_CEIL's overlapping edges
  const int ov = (H + 1 + PH - 1) / PH + 2;
  if (hb_max >= H) {
    p.g.NB = 1; p.g.Hb = H; p.g.step = H;
  } else {
    if (hb_max <= ov) return p;
    p.g.NB = (int)((H - ov + (hb_max - ov) - 1) / (hb_max - ov));
    p.g.step = (H - ov + p.g.NB - 1) / p.g.NB;
    p.g.Hb = p.g.step + ov;
    // A band OWNS the bins whose first row falls into its `step` rows, the last band also those of
    // the ov rows behind: with the shortest bands (above) the last band owns step + ov rows
    // (38 x 50: 15 and 23).  Equal shares (step = H / NB, 19 and 19) need ov more resident rows;
    // taken when they still leave room for the records and the grid is long -- measured on C4:
    // 256 images 3.012 -> 2.966 ms, 128 images 1.505 -> 1.492, but 64 images 0.805 -> 0.822 and
    // 32 images 0.403 -> 0.416 (taller bands wait longer for their copy, and with few waves the
    // equal CTAs end together) -- WSSDL_TUNE_ROI_FWD_BALANCED: -1 by grid size, 0 off, 1 on.
    const int step_eq = (H + p.g.NB - 1) / p.g.NB;
    const int btune = wssdl_tuning(WSSDL_TUNE_ROI_FWD_BALANCED);
    const long long ctas2 = (long long)(C / N_SLICE) * p.g.NB * 2 * (B > 0 ? B : 1);
    const bool balanced = btune > 0 || (btune < 0 && ctas2 >= 32ll * WSSDL_NUM_SMS);
    if (balanced && step_eq + ov <= hb_max &&
        budget - row_bytes * (size_t)(step_eq + ov) >= (size_t)48 * 1024) {
      p.g.step = step_eq;
      p.g.Hb = step_eq + ov;
    }
  }
  if (p.g.NB > NB_MAX || (size_t)p.g.Hb * W > 2047) return p;   // first-cell field: 11 bits
  // the records of a CTA's range of groups take what the band leaves: at most one band's worth
  // of an image of 512 RoIs at a time (longer ranges are pooled in several rounds)
  const size_t left = budget - row_bytes * p.g.Hb;
  size_t cap = (left / 8) & ~(size_t)7;
  const size_t cap_want = (((size_t)512 * PH * PW + 7) & ~(size_t)7) + N_CLS * 8;
  if (cap > cap_want) cap = cap_want;
  if (cap < 8 * 64) return p;
  p.rec_cap = (int)cap;
  p.smem = 128 + row_bytes * p.g.Hb + cap * 8;
  p.sort_rch = sort_rch(PH, PW);
  p.sort_smem = sort_smem(p.sort_rch, PH, PW, R, p.scan);
  if (p.sort_smem > (size_t)T_DYN_SMEM_MAX - 4096) return p;   // (+ the kernel's static arrays)

  // Ranges per band from a makespan model (us): a CTA stages the band (~2.5) and streams its
  // share of the output at ~1/148 of the HBM write rate.  One slice per CTA: the 16 slice CTAs of
  // a band then run side by side and complete each bin's 2 KB output row within microseconds
  // (measured on C4: 3.14 ms against 3.63 ms with two slices one after the other per CTA).
  const int n_slices = C / N_SLICE;
  const int nimg = B > 0 ? B : 1;
  const int per_sm = 1024 / p.threads;
  const double work_us = (double)R / nimg * PH * PW * 256.0 / p.g.NB / 40000.0 * per_sm;
  double best = 1e30;
  p.sg = 1;
  for (int nch = 1; nch <= 16; nch *= 2) {
    if (nch > 1 && (double)R / nimg * PH * PW / p.g.NB / nch < 8.0 * (p.threads / 32)) break;
    if ((long long)p.g.NB * nch > 65535) break;
    const long long ctas = (long long)n_slices * p.g.NB * nch * nimg;
    // (the quadratic term: the 16 slice CTAs of a band drift apart the longer they run, and the
    // 2 KB output rows they share then reach DRAM in pieces -- measured on C4, 256 images:
    // 3.14 / 3.03 / 3.21 ms with 1 / 2 / 3 ranges per band, 32 images: 0.428 / 0.412 / 0.430)
    const double w_cta = work_us / nch;
    const double t_cta = 1.0 + 2.5 / per_sm + w_cta + 0.004 * w_cta * w_cta;
    const double waves = (double)((ctas + WSSDL_NUM_SMS * per_sm - 1) / (WSSDL_NUM_SMS * per_sm));
    const double t = waves * t_cta;
    if (t < best * 0.999) { best = t; p.nchunks = nch; }
  }
  p.ok = true;
  return p;
}

cudaError_t wssdl_roi::launch_fwd_bins(const BinsPlan& p, const float* bottom, const float* rois,
                                       int B, int H, int W, int C, int R, int PH, int PW,
                                       float spatial_scale, int bin_mode, float* top, int* argmax,
                                       void* workspace, cudaStream_t s, int grouped_stride) {
  const bool pdl = wssdl_tuning(WSSDL_TUNE_PDL) != 0;
  int* img_start = nullptr;
  int* perm = nullptr;
  if (grouped_stride > 0 && (long long)grouped_stride * B != R) return cudaErrorInvalidValue;
  if (!p.scan && grouped_stride <= 0) {
    cudaError_t e = launch_roi_bucket(rois, R, B, workspace, s, &img_start, &perm);
    if (e != cudaSuccess) return e;
  }
  const BinsWs w = bins_ws(B, R, PH, PW);
  char* ws = static_cast<char*>(workspace);
  // ---- pre-pass: class-sorted bin records per (image, band)
  SortArgs sa;
  sa.rois = rois; sa.perm = perm; sa.img_start = img_start;
  sa.recs = reinterpret_cast<uint2*>(ws + w.off_recs);
  sa.table = reinterpret_cast<int4*>(ws + w.off_table);
  sa.B = B; sa.H = H; sa.W = W; sa.R = R; sa.PH = PH; sa.PW = PW;
  sa.spatial_scale = spatial_scale;
  sa.bin_mode = bin_mode;
  sa.rch = p.sort_rch;
  sa.stride = grouped_stride > 0 ? grouped_stride : 0;
  // sort CTAs per image: the pre-pass is issue bound on one SM per image, so while the batch
  // leaves SMs idle every image is sorted by 2 or 4 CTAs into as many lists (a pooling CTA's
  // range never spans two lists: nchunks is a multiple).  Not with in-kernel RoI lists (their
  // order differs from CTA to CTA).
  int nsub = 1;
  if (sa.stride > 0 || perm != nullptr) {
    for (int c = S_MAX; c > 1; c >>= 1)
      if (p.nchunks % c == 0 && (long long)(B + 1) * c <= WSSDL_NUM_SMS) { nsub = c; break; }
  }
  sa.nsub = nsub;
  sa.bg = p.g;
  sa.divPH = make_fastdiv((unsigned)PH);
  sa.divStep = make_fastdiv((unsigned)p.g.step);
  sa.divPW = make_fastdiv((unsigned)PW);
  sa.divPHPW = make_fastdiv((unsigned)(PH * PW));
  if (p.sort_smem > 48 * 1024) {                    // beyond the default carve-out: opt in
    cudaError_t e = cudaFuncSetAttribute(roi_bin_sort_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.sort_smem);
    if (e != cudaSuccess) return e;
  }
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((B + 1) * nsub));
    cfg.blockDim = dim3(S_THREADS);
    cfg.dynamicSmemBytes = p.sort_smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, roi_bin_sort_kernel, sa);
    if (e != cudaSuccess) return e;
  }
  // ---- tensor map of the map as (C, W, H, B), box = (32 channels, W, Hb rows, 1 image)
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  bool use_tma = false;
  if (EncodeTiledFn enc = tensor_map_encoder()) {
    const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(B > 0 ? B : 1)};
    const cuuint64_t gstr[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    const cuuint32_t box[4] = {(cuuint32_t)N_SLICE, (cuuint32_t)W, (cuuint32_t)p.g.Hb, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    use_tma = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(bottom), gdim, gstr,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  }
  const bool linear = (C % 128 == 0);
  PoolArgs a;
  a.bottom = bottom; a.rois = rois;
  a.recs = sa.recs; a.table = sa.table;
  a.top = top; a.argmax = argmax;
  a.B = B; a.H = H; a.W = W; a.C = C; a.PH = PH; a.PW = PW;
  a.spatial_scale = spatial_scale;
  a.bin_mode = bin_mode;
  a.nchunks = p.nchunks; a.sg = p.sg;
  a.nsub = nsub;
  a.n_slices = C / N_SLICE;
  a.rec_cap = p.rec_cap;
  a.bg = p.g;
  a.divPW = make_fastdiv((unsigned)PW);
  a.divPHPW = make_fastdiv((unsigned)(PH * PW));
  const int km = linear ? C / 128 : C;
  a.ones = {1.0f, {km, 0, 0, 0, 0, 0, 0, 0}};
  dim3 grid((unsigned)((a.n_slices + p.sg - 1) / p.sg), (unsigned)(p.g.NB * p.nchunks), (unsigned)(B + 1));
#define LAUNCH_BINS(NT, A, L, T)                                                              \
  do {                                                                                        \
    static unsigned long long done_k = 0;                                                     \
    cudaError_t e = allow_big_smem(roi_pool_fwd_bins_kernel<NT, A, L, T>, &done_k);           \
    if (e != cudaSuccess) return e;                                                           \
    cudaLaunchConfig_t cfg = {};                                                              \
    cfg.gridDim = grid;                                                                       \
    cfg.blockDim = dim3(NT);                                                                  \
    cfg.dynamicSmemBytes = p.smem;                                                            \
    cfg.stream = s;                                                                           \
    cudaLaunchAttribute attr[1];                                                              \
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                          \
    attr[0].val.programmaticStreamSerializationAllowed = 1;                                   \
    cfg.attrs = attr;                                                                         \
    cfg.numAttrs = pdl ? 1 : 0;                                                               \
    e = cudaLaunchKernelEx(&cfg, roi_pool_fwd_bins_kernel<NT, A, L, T>, tmap, a);             \
    if (e != cudaSuccess) return e;                                                           \
  } while (0)
#define LAUNCH_BINS_T(NT, A, L)                                                               \
  do {                                                                                        \
    if (use_tma) LAUNCH_BINS(NT, A, L, true);                                                 \
    else LAUNCH_BINS(NT, A, L, false);                                                        \
  } while (0)
#define LAUNCH_BINS_L(NT, A)                                                                  \
  do {                                                                                        \
    if (linear) LAUNCH_BINS_T(NT, A, true);                                                   \
    else LAUNCH_BINS_T(NT, A, false);                                                         \
  } while (0)
  if (p.threads == 512) {
    if (argmax) LAUNCH_BINS_L(512, true);
    else LAUNCH_BINS_L(512, false);
  } else {
    if (argmax) LAUNCH_BINS_L(1024, true);
    else LAUNCH_BINS_L(1024, false);
  }
#undef LAUNCH_BINS_L
#undef LAUNCH_BINS_T
#undef LAUNCH_BINS
  return cudaGetLastError();
}
