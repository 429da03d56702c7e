// Store-pattern microbenchmark (experiment, round 2): how fast does a B200 absorb the RoI-pool
// output stream (two [NBIN, 512] 4-byte arrays, 128 B pieces) depending on the ORDER in which the
// pieces are written?  grid = (16 slices, NCTA): CTA (s, y) writes the 128 B piece s of the 2 KB row
// of every bin of its list, 8 bins per warp iteration, 4 lanes x 32 B per bin (STG.256, .cs), both
// arrays.  mode 0: bins in ascending order; 1: fully random order; 2: random blocks of K
// consecutive bins (K = 7, 49, 392).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024, 1)
store_kernel(const unsigned* __restrict__ order, int bins_per_cta, float* top, int* arg) {
  const int s = blockIdx.x;
  const unsigned* list = order + (size_t)blockIdx.y * bins_per_cta;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = lane & 3, b8 = lane >> 2;
  for (int g = warp; g * 8 < bins_per_cta; g += 32) {
    const unsigned bin = list[g * 8 + b8];
    const size_t off = (size_t)bin * 512 + s * 32 + j * 8;
    asm volatile("st.global.cs.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(top + off), "f"(1.0f) : "memory");
    asm volatile("st.global.cs.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(arg + off), "r"(-1) : "memory");
  }
}

int main() {
  const int NIMG = 256, NCTA = NIMG * 2, PER_IMG = 300 * 49;       // 14700 bins per image
  const int bins_per_cta = PER_IMG / 2 / 8 * 8;                     // 7344
  const size_t nbin = (size_t)NIMG * PER_IMG;
  float* top; int* arg; unsigned* d_order;
  cudaMalloc(&top, nbin * 512 * 4); cudaMalloc(&arg, nbin * 512 * 4);
  cudaMalloc(&d_order, (size_t)NCTA * bins_per_cta * 4);
  std::mt19937 rng(1);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int Ks[] = {0, 1, 2, 4, 7, 16, 49, -49, -196, -392, -784, -1568, -3675};
  for (int K0 : Ks) {
    const bool shuffle_inside = K0 < 0;   // negative: ascending block order, random order inside a block
    const int K = K0 < 0 ? -K0 : K0;
    std::vector<unsigned> order((size_t)NCTA * bins_per_cta);
    for (int c = 0; c < NCTA; ++c) {
      const unsigned base = (unsigned)((c / 2) * PER_IMG + (c % 2) * (PER_IMG / 2));
      std::vector<unsigned> v(bins_per_cta);
      for (int i = 0; i < bins_per_cta; ++i) v[i] = base + i;
      if (K == 1) std::shuffle(v.begin(), v.end(), rng);
      else if (K > 1 && !shuffle_inside) {   // random order of blocks of K consecutive bins
        const int nb = (bins_per_cta + K - 1) / K;
        std::vector<int> blk(nb); for (int i = 0; i < nb; ++i) blk[i] = i;
        std::shuffle(blk.begin(), blk.end(), rng);
        std::vector<unsigned> w; w.reserve(bins_per_cta);
        for (int b : blk) for (int i = b * K; i < std::min((b + 1) * K, bins_per_cta); ++i) w.push_back(base + i);
        v = w;
      } else if (K > 1) {                    // blocks in ascending order, shuffled inside
        for (int b = 0; b * K < bins_per_cta; ++b)
          std::shuffle(v.begin() + b * K, v.begin() + std::min((b + 1) * K, bins_per_cta), rng);
      }
      std::copy(v.begin(), v.end(), order.begin() + (size_t)c * bins_per_cta);
    }
    cudaMemcpy(d_order, order.data(), order.size() * 4, cudaMemcpyHostToDevice);
    dim3 grid(16, NCTA);
    for (int i = 0; i < 3; ++i) store_kernel<<<grid, 1024>>>(d_order, bins_per_cta, top, arg);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) store_kernel<<<grid, 1024>>>(d_order, bins_per_cta, top, arg);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    const double bytes = (double)NCTA * bins_per_cta * 16 * 128 * 2;
    printf("order K=%5d (0 ascending, 1 random, K>1 random blocks of K ascending bins, K<0 ascending blocks shuffled inside): %.3f ms  %.0f GB/s  (%s)\n", K0, ms,
           bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
