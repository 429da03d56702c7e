// Greedy NMS on the device for sm_100a: sort -> 64-bit bitmask over the upper triangle ->
// on-device sweep.  Semantics: cpu_nms (nms/cpu_nms.pyx:17-68), utils.cython_nms.nms /
// nms_new (utils/nms.pyx), gpu_nms/_nms (nms/nms_kernel.cu:34-144), py_cpu_nms.
//
// Stages (all on the caller's stream, scratch in the caller's workspace, no host round
// trip -- the reference copies the N x N/64 mask to the host and sweeps there,
// nms_kernel.cu:118-139):
//   1. key build: 64-bit key = (~orderable(score) << 32) | ~index, so an ascending sort
//      yields (score desc, index desc) = scores.argsort(kind='stable')[::-1];
//   2. bitonic sort of the padded key array (shared-memory tiles of 4096 keys, global
//      compare-exchange passes only for strides >= 4096);
//   3. gather: sorted boxes as float4 + the fp32 area computed like numpy does
//      (cpu_nms.pyx:24: one rounding per operation);
//   4. mask: one thread per sorted row and 64-column block, column boxes staged in shared
//      memory; only blocks on or above the diagonal are computed;
//   5. sweep: one CTA; per 64-row block a single warp resolves the diagonal word set with a
//      ballot-style fixed-point iteration (K <- alive & ~OR_{i in K} row_i, unique fixed
//      point = the greedy solution because the dependency matrix is strictly upper
//      triangular), then all threads OR the kept rows into the removal bitmap.
//
// IoU arithmetic mirrors the generated C of the reference (cpu_nms.c:2442-2495): every
// operation is a single fp32 rounding (no FMA), IEEE division, and the threshold test is
// done in the precision the chosen mode prescribes.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>
#include <mutex>

#include "common.cuh"
#include "nms_common.cuh"

namespace {

constexpr int SORT_TILE = 4096;      // keys per shared-memory tile (32 KB)
constexpr int SORT_THREADS = 1024;

__global__ void nms_build_keys(const float* __restrict__ dets, int N, int stride, int npad,
                               unsigned long long* __restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npad) return;
  unsigned long long k = ~0ull;  // padding sorts last
  if (i < N) {
    const unsigned s = orderable(dets[(size_t)i * stride + 4]);
    k = ((unsigned long long)(~s) << 32) | (unsigned)(~(unsigned)i);
  }
  keys[i] = k;
}

__device__ __forceinline__ void cmpxchg(unsigned long long& a, unsigned long long& b, bool up) {
  if ((a > b) == up) { unsigned long long t = a; a = b; b = t; }
}

// Sort each SORT_TILE-key tile completely (bitonic); direction alternates per tile so the
// tiles form bitonic sequences for the global merge stages.
__global__ void __launch_bounds__(SORT_THREADS)
bitonic_tile_sort(unsigned long long* __restrict__ keys) {
  __shared__ unsigned long long s[SORT_TILE];
  const size_t base = (size_t)blockIdx.x * SORT_TILE;
  for (int i = threadIdx.x; i < SORT_TILE; i += SORT_THREADS) s[i] = keys[base + i];
  __syncthreads();
  for (int k = 2; k <= SORT_TILE; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < SORT_TILE / 2; t += SORT_THREADS) {
        const int i = 2 * t - (t & (j - 1));  // lower index of the pair with stride j
        const size_t gi = base + i;
        const bool up = ((gi & (size_t)k) == 0);
        cmpxchg(s[i], s[i + j], up);
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < SORT_TILE; i += SORT_THREADS) keys[base + i] = s[i];
}

// One global compare-exchange pass: stage size k, stride j >= SORT_TILE.
__global__ void bitonic_global_step(unsigned long long* __restrict__ keys, int npad, int k,
                                    int j) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npad / 2) return;
  const int i = 2 * t - (t & (j - 1));
  const bool up = ((i & k) == 0);
  unsigned long long a = keys[i], b = keys[i + j];
  if ((a > b) == up) { keys[i] = b; keys[i + j] = a; }
}

// Finish stage k inside a tile: all strides j < SORT_TILE.
__global__ void __launch_bounds__(SORT_THREADS)
bitonic_tile_merge(unsigned long long* __restrict__ keys, int k) {
  __shared__ unsigned long long s[SORT_TILE];
  const size_t base = (size_t)blockIdx.x * SORT_TILE;
  for (int i = threadIdx.x; i < SORT_TILE; i += SORT_THREADS) s[i] = keys[base + i];
  __syncthreads();
  const bool up = ((base & (size_t)k) == 0);
  for (int j = SORT_TILE >> 1; j > 0; j >>= 1) {
    for (int t = threadIdx.x; t < SORT_TILE / 2; t += SORT_THREADS) {
      const int i = 2 * t - (t & (j - 1));
      cmpxchg(s[i], s[i + j], up);
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < SORT_TILE; i += SORT_THREADS) keys[base + i] = s[i];
}

// sorted position p -> original index, box (float4) and numpy-style fp32 area
__global__ void nms_gather(const float* __restrict__ dets, int N, int stride,
                           const unsigned long long* __restrict__ keys, int presorted,
                           int* __restrict__ order, float4* __restrict__ boxes,
                           float* __restrict__ areas) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= N) return;
  const int i = presorted ? p : (int)(~(unsigned)(keys[p] & 0xffffffffull));
  const float* d = dets + (size_t)i * stride;
  const float x1 = d[0], y1 = d[1], x2 = d[2], y2 = d[3];
  order[p] = i;
  boxes[p] = make_float4(x1, y1, x2, y2);
  // (x2 - x1 + 1) * (y2 - y1 + 1), each op rounded to fp32 (cpu_nms.pyx:24)
  areas[p] = __fmul_rn(__fadd_rn(__fsub_rn(x2, x1), 1.0f), __fadd_rn(__fsub_rn(y2, y1), 1.0f));
}

constexpr int MASK_ROWS = 256;  // rows (threads) per CTA of the mask kernel

// mask[row * col_blocks + cb] bit c = sorted box row suppresses sorted box 64*cb+c (c > row)
__global__ void __launch_bounds__(MASK_ROWS)
nms_mask_kernel(const float4* __restrict__ boxes, const float* __restrict__ areas, int N,
                int col_blocks, Thresh th, unsigned long long* __restrict__ mask,
                int* __restrict__ status) {
  const int cb = blockIdx.x;
  const int row0 = blockIdx.y * MASK_ROWS;
  if (64 * cb + 63 <= row0) return;  // every column <= every row: below the diagonal
  __shared__ float4 s_box[64];
  __shared__ float s_area[64];
  const int ncol = min(64, N - 64 * cb);
  if (threadIdx.x < ncol) {
    s_box[threadIdx.x] = boxes[64 * cb + threadIdx.x];
    s_area[threadIdx.x] = areas[64 * cb + threadIdx.x];
  }
  __syncthreads();
  const int row = row0 + threadIdx.x;
  if (row >= N || (row >> 6) > cb) return;
  const float4 bi = boxes[row];
  const float ai = areas[row];
  const int start = ((row >> 6) == cb) ? (row & 63) + 1 : 0;
  unsigned long long bits = 0;
  bool zero = false;
  for (int c = start; c < ncol; ++c) {
    if (suppresses(bi, ai, s_box[c], s_area[c], th, &zero)) bits |= 1ull << c;
  }
  mask[(size_t)row * col_blocks + cb] = bits;
  if (zero && status != nullptr) status[0] = 1;
}

constexpr int SWEEP_THREADS = 1024;
constexpr int SWEEP_OR_THREADS = SWEEP_THREADS - 32;   // warps 1..31 run the OR phase

// Single-CTA sweep.  remv (one bit per sorted box) lives in shared memory.
//
// Per 64-row block b: warp 0 RESOLVES the block (fixed point of "alive and not suppressed by
// a kept earlier row of the block" on the diagonal word), then the kept rows are ORed into the
// removal bitmap of the later column blocks.  Those two steps used to run back to back, the
// second one an L2-latency-bound gather on the critical path of every block.  They are now
// software pipelined over the warps: while warp 0 resolves block b, warps 1..31 OR the kept
// rows of block b-1 into columns >= b+1.  Warp 0 needs column b of block b-1 before that OR
// phase has run, so it fetches the 64 words mask[rows of block b-1][b] itself one block ahead
// (they do not depend on what is kept) and reduces the kept ones with two warp ORs
// (`fast`).  One barrier per block; the critical path per block is max(resolve, OR phase)
// instead of their sum.
__global__ void __launch_bounds__(SWEEP_THREADS)
nms_sweep_kernel(const unsigned long long* __restrict__ mask, int N, int col_blocks,
                 const int* __restrict__ order, int max_keep, int* __restrict__ keep,
                 int* __restrict__ num_keep) {
  extern __shared__ unsigned long long remv[];   // col_blocks words
  __shared__ unsigned long long s_kept[2];       // kept mask of block b at [b & 1]
  __shared__ int s_count;
  __shared__ int s_full[2];          // [b & 1]: the keep list is full after block b
  __shared__ int s_rows[2][64];                  // kept rows of block b at [b & 1]
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < col_blocks; i += SWEEP_THREADS) remv[i] = 0;
  if (tid == 0) { s_count = 0; s_kept[0] = 0ull; s_kept[1] = 0ull; }
  __syncthreads();
  const int limit = max_keep > 0 ? min(max_keep, N) : N;

  // warp 0, fetched one block ahead: the diagonal words of rows 64b+lane and 64b+32+lane
  // (nd*) and the words of the same rows in column b+1 (nc*); neither depends on the
  // removal bitmap, only their interpretation does
  unsigned long long nd0 = 0ull, nd1 = 0ull, nc0 = 0ull, nc1 = 0ull;
  if (tid < 32) {
    nd0 = lane < N ? mask[(size_t)lane * col_blocks] : 0ull;
    nd1 = lane + 32 < N ? mask[(size_t)(lane + 32) * col_blocks] : 0ull;
    if (col_blocks > 1) {
      nc0 = lane < N ? mask[(size_t)lane * col_blocks + 1] : 0ull;
      nc1 = lane + 32 < N ? mask[(size_t)(lane + 32) * col_blocks + 1] : 0ull;
    }
  }
  unsigned long long fast = 0ull;                // block b-1's removals in column b (warp 0)
  for (int b = 0; b < col_blocks; ++b) {
    if (tid < 32) {
      // ---- resolve block b
      const int r0 = 64 * b + lane, r1 = r0 + 32;
      const unsigned long long d0 = nd0, d1 = nd1, c0 = nc0, c1 = nc1;
      if (b + 1 < col_blocks) {
        const int q0 = r0 + 64, q1 = r1 + 64;
        nd0 = q0 < N ? mask[(size_t)q0 * col_blocks + b + 1] : 0ull;
        nd1 = q1 < N ? mask[(size_t)q1 * col_blocks + b + 1] : 0ull;
        if (b + 2 < col_blocks) {
          nc0 = q0 < N ? mask[(size_t)q0 * col_blocks + b + 2] : 0ull;
          nc1 = q1 < N ? mask[(size_t)q1 * col_blocks + b + 2] : 0ull;
        }
      }
      const int nrow = min(64, N - 64 * b);
      const unsigned long long valid = nrow == 64 ? ~0ull : ((1ull << nrow) - 1ull);
      const unsigned long long alive = ~(remv[b] | fast) & valid;
      unsigned long long K = alive;
      for (int it = 0; it < 64; ++it) {
        unsigned long long mine = 0;
        if ((K >> lane) & 1ull) mine |= d0;
        if ((K >> (lane + 32)) & 1ull) mine |= d1;
        const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)mine);
        const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(mine >> 32));
        const unsigned long long Knew = alive & ~(((unsigned long long)hi << 32) | lo);
        if (Knew == K) break;
        K = Knew;
      }
      // truncate to the first (limit - count) kept boxes of this block
      int count = s_count;
      int take = __popcll(K);
      if (count + take > limit) {
        int over = count + take - limit;
        while (over-- > 0) K &= ~(1ull << (63 - __clzll(K)));
        take = limit - count;
      }
      // what the kept rows of this block remove in column b+1 (for the next resolve)
      {
        unsigned long long mine = 0;
        if ((K >> lane) & 1ull) mine |= c0;
        if ((K >> (lane + 32)) & 1ull) mine |= c1;
        const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)mine);
        const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(mine >> 32));
        fast = ((unsigned long long)hi << 32) | lo;
      }
      // write kept indices in order
      if ((K >> lane) & 1ull)
        keep[count + __popcll(K & ((1ull << lane) - 1ull))] = order[r0];
      if ((K >> (lane + 32)) & 1ull)
        keep[count + __popcll(K & ((1ull << (lane + 32)) - 1ull))] = order[r1];
      // the OR phase of this block runs in the NEXT iteration (warps 1..31)
      if ((K >> lane) & 1ull) s_rows[b & 1][__popcll(K & ((1ull << lane) - 1ull))] = lane;
      if ((K >> (lane + 32)) & 1ull)
        s_rows[b & 1][__popcll(K & ((1ull << (lane + 32)) - 1ull))] = lane + 32;
      __syncwarp();                       // every lane has read s_count (racecheck: WAR)
      if (lane == 0) {
        s_kept[b & 1] = K;
        s_count = count + take;
        s_full[b & 1] = count + take >= limit;
      }
    } else if (b > 0) {
      // ---- OR the kept rows of block b-1 into the removal bitmap of column blocks >= b+1
      // (column b was handled by warp 0's `fast`).  (kept row, column) pairs are spread over
      // warps 1..31 -- RP row parts x columns -- and every thread issues its loads eight at a
      // time before using them.
      const int pb = b - 1;
      const unsigned long long K = s_kept[pb & 1];
      const int nc = col_blocks - (pb + 2);
      if (nc > 0 && K != 0ull) {
        const int* rows = s_rows[pb & 1];
        const int t = tid - 32;
        const int kc = __popcll(K);
        int rp_log2 = 0;
        while (rp_log2 < 5 && ((nc << (rp_log2 + 1)) <= SWEEP_OR_THREADS)) ++rp_log2;
        const int RP = 1 << rp_log2;
        const int part = t & (RP - 1);
        const int cols_per_pass = SWEEP_OR_THREADS >> rp_log2;
        const unsigned long long* mrow = mask + (size_t)(64 * pb) * col_blocks + (pb + 2);
        for (int c = t >> rp_log2; c < nc; c += cols_per_pass) {
          unsigned long long acc = 0;
          int ri = part;
          for (; ri + 7 * RP < kc; ri += 8 * RP) {
            unsigned long long w[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) w[u] = mrow[(size_t)rows[ri + u * RP] * col_blocks + c];
            acc |= ((w[0] | w[1]) | (w[2] | w[3])) | ((w[4] | w[5]) | (w[6] | w[7]));
          }
          for (; ri + 3 * RP < kc; ri += 4 * RP) {
            const unsigned long long w0 = mrow[(size_t)rows[ri] * col_blocks + c];
            const unsigned long long w1 = mrow[(size_t)rows[ri + RP] * col_blocks + c];
            const unsigned long long w2 = mrow[(size_t)rows[ri + 2 * RP] * col_blocks + c];
            const unsigned long long w3 = mrow[(size_t)rows[ri + 3 * RP] * col_blocks + c];
            acc |= (w0 | w1) | (w2 | w3);
          }
          for (; ri < kc; ri += RP) acc |= mrow[(size_t)rows[ri] * col_blocks + c];
          if (acc != 0ull) {
            if (RP == 1) remv[pb + 2 + c] |= acc;
            else atomicOr(&remv[pb + 2 + c], acc);
          }
        }
      }
    }
    __syncthreads();
    // read a slot warp 0 will not touch again before every warp has passed the NEXT barrier
    // (s_count itself is rewritten by warp 0 as soon as it runs ahead into block b + 1)
    if (s_full[b & 1]) break;
  }
  if (tid == 0) *num_keep = s_count;
}

// Sweep over a thread-block cluster of CS CTAs (large N).  The single-CTA sweep above is a
// dependent chain over 64-row blocks whose OR phase (kept rows x remaining columns) runs on
// one SM; here the column words of the removal bitmap are interleaved over the CTAs of a
// cluster (word w belongs to CTA w % CS).  Block b is resolved by its owner (it holds the
// authoritative remv[b]), which writes the kept mask K of the block into every CTA's shared
// memory (distributed shared memory, two slots alternating by block parity) -- one cluster
// barrier per block -- and then every CTA ORs the kept rows into ITS columns.  The next owner
// needs nothing from its peers but K, so no second barrier is required.  The running keep
// count is replicated (every CTA adds popc(K)).
template <int CS>
__global__ void __launch_bounds__(SWEEP_THREADS)
nms_sweep_cluster_kernel(const unsigned long long* __restrict__ mask, int N, int col_blocks,
                         const int* __restrict__ order, int max_keep, int* __restrict__ keep,
                         int* __restrict__ num_keep) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  extern __shared__ unsigned long long remv[];   // col_blocks words; only words w % CS == crank are kept current
  __shared__ unsigned long long s_kslot[2];      // kept mask of the current block, by parity
  __shared__ int s_rows[64];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int i = tid; i < col_blocks; i += SWEEP_THREADS) remv[i] = 0;
  __syncthreads();
  const int limit = max_keep > 0 ? min(max_keep, N) : N;
  int count = 0;                                 // replicated in every thread of every CTA

  // diagonal words of my next owned block (warp 0), fetched one owned block ahead
  unsigned long long nd0 = 0ull, nd1 = 0ull;
  if (tid < 32 && crank < col_blocks) {
    const int r0 = 64 * crank + lane, r1 = r0 + 32;
    nd0 = r0 < N ? mask[(size_t)r0 * col_blocks + crank] : 0ull;
    nd1 = r1 < N ? mask[(size_t)r1 * col_blocks + crank] : 0ull;
  }
  cluster.sync();                                // every CTA's slots exist before anyone writes
  for (int b = 0; b < col_blocks; ++b) {
    const int set = b & 1;
    if (b % CS == crank && tid < 32) {
      const int r0 = 64 * b + lane, r1 = r0 + 32;
      const unsigned long long d0 = nd0, d1 = nd1;
      if (b + CS < col_blocks) {
        const int q0 = r0 + 64 * CS, q1 = r1 + 64 * CS;
        nd0 = q0 < N ? mask[(size_t)q0 * col_blocks + b + CS] : 0ull;
        nd1 = q1 < N ? mask[(size_t)q1 * col_blocks + b + CS] : 0ull;
      }
      const int nrow = min(64, N - 64 * b);
      const unsigned long long valid = nrow == 64 ? ~0ull : ((1ull << nrow) - 1ull);
      const unsigned long long alive = ~remv[b] & valid;
      unsigned long long K = alive;
      for (int it = 0; it < 64; ++it) {
        unsigned long long mine = 0;
        if ((K >> lane) & 1ull) mine |= d0;
        if ((K >> (lane + 32)) & 1ull) mine |= d1;
        const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)mine);
        const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(mine >> 32));
        const unsigned long long Knew = alive & ~(((unsigned long long)hi << 32) | lo);
        if (Knew == K) break;
        K = Knew;
      }
      int take = __popcll(K);
      if (count + take > limit) {                // truncate to the first (limit - count) kept
        int over = count + take - limit;
        while (over-- > 0) K &= ~(1ull << (63 - __clzll(K)));
      }
      if ((K >> lane) & 1ull)
        keep[count + __popcll(K & ((1ull << lane) - 1ull))] = order[r0];
      if ((K >> (lane + 32)) & 1ull)
        keep[count + __popcll(K & ((1ull << (lane + 32)) - 1ull))] = order[r1];
      if (lane < CS) *cluster.map_shared_rank(&s_kslot[set], lane) = K;
    }
    cluster.sync();
    const unsigned long long K = s_kslot[set];
    count += __popcll(K);
    if (count >= limit) break;
    // OR the kept rows of this block into MY later column words
    const int c0 = b + 1 + ((crank - (b + 1)) % CS + CS) % CS;   // first column > b with c % CS == crank
    const int nc = c0 < col_blocks ? (col_blocks - c0 + CS - 1) / CS : 0;
    if (nc > 0 && K != 0ull) {
      if (tid < 64) {
        if ((K >> tid) & 1ull) s_rows[__popcll(K & ((1ull << tid) - 1ull))] = tid;
      }
      __syncthreads();
      const int kc = __popcll(K);
      int rp_log2 = 0;
      while (rp_log2 < 5 && ((nc << (rp_log2 + 1)) <= SWEEP_THREADS)) ++rp_log2;
      const int RP = 1 << rp_log2;
      const int part = tid & (RP - 1);
      const int cols_per_pass = SWEEP_THREADS >> rp_log2;
      const unsigned long long* mrow = mask + (size_t)(64 * b) * col_blocks;
      for (int ci = tid >> rp_log2; ci < nc; ci += cols_per_pass) {
        const int c = c0 + CS * ci;
        unsigned long long acc = 0;
        int ri = part;
        for (; ri + 7 * RP < kc; ri += 8 * RP) {
          unsigned long long w[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) w[u] = mrow[(size_t)s_rows[ri + u * RP] * col_blocks + c];
          acc |= ((w[0] | w[1]) | (w[2] | w[3])) | ((w[4] | w[5]) | (w[6] | w[7]));
        }
        for (; ri < kc; ri += RP) acc |= mrow[(size_t)s_rows[ri] * col_blocks + c];
        if (acc != 0ull) {
          if (RP == 1) remv[c] |= acc;
          else atomicOr(&remv[c], acc);
        }
      }
    }
    __syncthreads();
  }
  cluster.sync();                                // nobody exits while a peer may still write its slot
  if (crank == 0 && tid == 0) *num_keep = min(count, limit);
}

struct Layout {
  int npad;
  size_t keys, order, boxes, areas, mask, total;
};

Layout layout_for(int N) {
  Layout L;
  int npad = SORT_TILE;
  while (npad < N) npad <<= 1;
  L.npad = npad;
  const size_t cb = (size_t)(N + 63) / 64;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  L.keys = take(sizeof(unsigned long long) * (size_t)npad);
  L.order = take(sizeof(int) * (size_t)N);
  L.boxes = take(sizeof(float4) * (size_t)N);
  L.areas = take(sizeof(float) * (size_t)N);
  L.mask = take(sizeof(unsigned long long) * (size_t)N * cb);
  L.total = off;
  return L;
}

int run_nms(const float* dets, int N, int stride, double thresh, int mode, int max_keep,
            int presorted, int* keep, int* num_keep, int* status, void* workspace,
            size_t workspace_bytes, cudaStream_t s) {
  if (N < 0 || stride < 5 || !num_keep) return WSSDL_EINVAL;
  if ((mode & ~(3 | WSSDL_NMS_CONTAIN)) != 0 || (mode & 3) > 1) return WSSDL_EINVAL;
  if (status) WSSDL_RETURN_IF_CUDA(cudaMemsetAsync(status, 0, 2 * sizeof(int), s));
  if (N == 0) {
    WSSDL_RETURN_IF_CUDA(cudaMemsetAsync(num_keep, 0, sizeof(int), s));
    return WSSDL_OK;
  }
  if (!dets || !keep || !workspace) return WSSDL_EINVAL;
  if (N > (1 << 20)) return WSSDL_ELIMIT;   // sweep bitmap and int indexing
  const Layout L = layout_for(N);
  if (workspace_bytes < L.total) return WSSDL_EWORKSPACE;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return WSSDL_EALIGN;
  char* ws = static_cast<char*>(workspace);
  auto* keys = reinterpret_cast<unsigned long long*>(ws + L.keys);
  int* order = reinterpret_cast<int*>(ws + L.order);
  float4* boxes = reinterpret_cast<float4*>(ws + L.boxes);
  float* areas = reinterpret_cast<float*>(ws + L.areas);
  auto* mask = reinterpret_cast<unsigned long long*>(ws + L.mask);
  const int col_blocks = (N + 63) / 64;

  if (!presorted) {
    nms_build_keys<<<ceil_div(L.npad, 256), 256, 0, s>>>(dets, N, stride, L.npad, keys);
    const int tiles = L.npad / SORT_TILE;
    bitonic_tile_sort<<<tiles, SORT_THREADS, 0, s>>>(keys);
    for (int k = 2 * SORT_TILE; k <= L.npad; k <<= 1) {
      for (int j = k >> 1; j >= SORT_TILE; j >>= 1)
        bitonic_global_step<<<ceil_div(L.npad / 2, 256), 256, 0, s>>>(keys, L.npad, k, j);
      bitonic_tile_merge<<<tiles, SORT_THREADS, 0, s>>>(keys, k);
    }
  }
  nms_gather<<<ceil_div(N, 256), 256, 0, s>>>(dets, N, stride, keys, presorted, order, boxes,
                                             areas);
  const Thresh th = make_thresh(thresh, mode);
  dim3 mgrid((unsigned)col_blocks, (unsigned)ceil_div(N, MASK_ROWS));
  nms_mask_kernel<<<mgrid, MASK_ROWS, 0, s>>>(boxes, areas, N, col_blocks, th, mask, status);
  // Very large N: the sweep runs on a cluster of 8 CTAs.  Measured on B200 (thr 0.7, pipelined
  // single CTA / cluster): boxes that mostly survive 6 000: 0.363 / 0.403 ms, 20 000: 1.39 /
  // 1.44, 50 000: 6.59 / 5.90, 100 000: 21.7 / 16.6; RPN-like clustered boxes (few survivors:
  // the per-block barrier is pure latency) 6 000: 0.272 / 0.370, 20 000: 0.886 / 1.18, 50 000:
  // 3.45 / 4.18, 100 000: 11.8 / 12.6.  Default: cluster from N = 65 536.
  // WSSDL_NMS_SWEEP_CLUSTER=0|1 overrides.
  constexpr int SWEEP_CS = 8;
  const int ctune = wssdl_tuning(WSSDL_TUNE_NMS_SWEEP_CLUSTER);
  const bool clustered = ctune >= 0 ? (ctune == 1) : (col_blocks >= 1024);
  const size_t sweep_smem = sizeof(unsigned long long) * (size_t)col_blocks;
  if (sweep_smem > 48 * 1024) {                     // N > 393 216: opt in to the large carve-out
    if (clustered)
      WSSDL_RETURN_IF_CUDA(cudaFuncSetAttribute(nms_sweep_cluster_kernel<SWEEP_CS>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)sweep_smem));
    else
      WSSDL_RETURN_IF_CUDA(cudaFuncSetAttribute(nms_sweep_kernel,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)sweep_smem));
  }
  if (clustered) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(SWEEP_CS);
    cfg.blockDim = dim3(SWEEP_THREADS);
    cfg.dynamicSmemBytes = sweep_smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = SWEEP_CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    WSSDL_RETURN_IF_CUDA(cudaLaunchKernelEx(&cfg, nms_sweep_cluster_kernel<SWEEP_CS>,
                                            (const unsigned long long*)mask, N, col_blocks,
                                            (const int*)order, max_keep, keep, num_keep));
  } else {
    nms_sweep_kernel<<<1, SWEEP_THREADS, sweep_smem, s>>>(mask, N, col_blocks, order, max_keep,
                                                          keep, num_keep);
  }
  WSSDL_CHECK_LAUNCH();
  return WSSDL_OK;
}

// Grow-only device scratch for the *_host entry points (one per process, mutex-guarded).
std::mutex g_host_mu;
void* g_host_buf = nullptr;
size_t g_host_cap = 0;
int g_host_dev = -1;

int host_scratch(size_t bytes, int device, void** out) {
  if (g_host_buf && (g_host_cap < bytes || g_host_dev != device)) {
    cudaFree(g_host_buf);
    g_host_buf = nullptr;
    g_host_cap = 0;
  }
  if (!g_host_buf) {
    WSSDL_RETURN_IF_CUDA(cudaMalloc(&g_host_buf, bytes));
    g_host_cap = bytes;
    g_host_dev = device;
  }
  *out = g_host_buf;
  return WSSDL_OK;
}

int nms_host_common(int* keep_out, int* num_out, const float* dets_host, int N, int stride,
                    double thresh, int mode, int max_keep, int presorted, int device_id) {
  if (!num_out) return WSSDL_EINVAL;
  *num_out = 0;
  if (N == 0) return WSSDL_OK;
  if (N < 0 || !keep_out || !dets_host || stride < 4) return WSSDL_EINVAL;
  std::lock_guard<std::mutex> lock(g_host_mu);
  WSSDL_RETURN_IF_CUDA(cudaSetDevice(device_id));
  const size_t ws_bytes = layout_for(N).total;
  const size_t dets_bytes = ((sizeof(float) * (size_t)N * 5) + 255) & ~(size_t)255;
  const size_t keep_bytes = ((sizeof(int) * (size_t)N) + 255) & ~(size_t)255;
  void* buf = nullptr;
  int rc = host_scratch(ws_bytes + dets_bytes + keep_bytes + 256, device_id, &buf);
  if (rc != WSSDL_OK) return rc;
  char* p = static_cast<char*>(buf);
  float* d_dets = reinterpret_cast<float*>(p + ws_bytes);
  int* d_keep = reinterpret_cast<int*>(p + ws_bytes + dets_bytes);
  int* d_misc = reinterpret_cast<int*>(p + ws_bytes + dets_bytes + keep_bytes);  // num, status[2]
  cudaStream_t s = 0;
  if (stride == 5) {
    WSSDL_RETURN_IF_CUDA(cudaMemcpyAsync(d_dets, dets_host, sizeof(float) * (size_t)N * 5,
                                         cudaMemcpyHostToDevice, s));
  } else {
    // boxes_dim != 5: copy the first min(stride,5) columns of every row, pad the score with 0
    WSSDL_RETURN_IF_CUDA(cudaMemsetAsync(d_dets, 0, sizeof(float) * (size_t)N * 5, s));
    WSSDL_RETURN_IF_CUDA(cudaMemcpy2DAsync(d_dets, 5 * sizeof(float), dets_host,
                                           (size_t)stride * sizeof(float),
                                           (size_t)(stride < 5 ? stride : 5) * sizeof(float),
                                           (size_t)N, cudaMemcpyHostToDevice, s));
  }
  rc = run_nms(d_dets, N, 5, thresh, mode, max_keep, presorted, d_keep, d_misc, d_misc + 1, buf,
               ws_bytes, s);
  if (rc != WSSDL_OK) return rc;
  int misc[3];
  WSSDL_RETURN_IF_CUDA(cudaMemcpyAsync(misc, d_misc, sizeof(misc), cudaMemcpyDeviceToHost, s));
  WSSDL_RETURN_IF_CUDA(cudaStreamSynchronize(s));
  if (misc[0] > 0)
    WSSDL_RETURN_IF_CUDA(cudaMemcpy(keep_out, d_keep, sizeof(int) * (size_t)misc[0],
                                    cudaMemcpyDeviceToHost));
  *num_out = misc[0];
  return misc[1] ? WSSDL_EZERODIV : WSSDL_OK;
}

}  // namespace

extern "C" size_t wssdl_nms_workspace_bytes(int N) {
  if (N <= 0) return 256;
  return layout_for(N).total;
}

extern "C" int wssdl_nms(const float* dets, int N, int dets_stride, double thresh, int mode,
                         int max_keep, int* keep, int* num_keep, int* status, void* workspace,
                         size_t workspace_bytes, wssdl_stream_t stream) {
  return run_nms(dets, N, dets_stride, thresh, mode, max_keep, /*presorted=*/0, keep, num_keep,
                 status, workspace, workspace_bytes, to_cuda(stream));
}

extern "C" int wssdl_gpu_nms_host(int* keep_out, int* num_out, const float* boxes_host,
                                  int boxes_num, int boxes_dim, float nms_overlap_thresh,
                                  int device_id) {
  // same contract as `_nms` (nms_kernel.cu:91): input sorted by the caller, '>' test in
  // fp32, keep_out holds positions in the sorted array.
  return nms_host_common(keep_out, num_out, boxes_host, boxes_num, boxes_dim,
                         (double)nms_overlap_thresh, WSSDL_NMS_GT_F32, 0, /*presorted=*/1,
                         device_id);
}

extern "C" int wssdl_nms_host(int* keep_out, int* num_out, const float* dets_host, int N,
                              int dets_stride, double thresh, int mode, int max_keep,
                              int device_id) {
  if (dets_stride < 5) return WSSDL_EINVAL;
  return nms_host_common(keep_out, num_out, dets_host, N, dets_stride, thresh, mode, max_keep,
                         /*presorted=*/0, device_id);
}
