// Host emulation of the few CUDA constructs the reference's roi_pooling_op_gpu.cu.cc uses, so
// that its kernel BODIES (the GPU_CEIL bin arithmetic, roi_pooling_op_gpu.cu.cc:19-85) can be
// executed on the CPU without a GPU or nvcc.  TEST INFRASTRUCTURE ONLY; no arithmetic here.
// oracle/build_ref.py rewrites the two `kernel<<<grid, block, smem, stream>>>(args);` launch
// statements into CUDA_EMU_LAUNCH(grid, block, kernel(args)); -- the only source edit -- which
// calls the kernel function once per (block, thread) with fresh by-value parameters, exactly
// like a launch (the kernels advance their pointer parameters inside the grid-stride loop).
//
// nms/nms_kernel.cu (the 64-bit bitmask NMS behind gpu_nms) is built the same way.  Its kernel
// stages 64 boxes in __shared__ memory behind one __syncthreads(); the emulation runs the
// threads of a block one after the other, TWICE: __shared__ is a static array, the first sweep
// fills it (and writes mask words that may have been computed from a half-filled array), the
// second sweep recomputes every mask word from the complete array.  The kernel only reads
// global + shared memory and writes its own mask word, so the second sweep's output is what a
// real launch produces.  cudaMalloc / cudaMemcpy / cudaFree map to malloc / memcpy / free.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

using std::max;
using std::min;

#define __global__
#define __device__
#define __shared__ static
static inline void __syncthreads() {}
struct dim3 {
  int x, y, z;
  dim3(int x_ = 1, int y_ = 1, int z_ = 1) : x(x_), y(y_), z(z_) {}
};
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost };
struct CudaEmuDim { int x, y, z; };
static CudaEmuDim blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, threadIdx = {0, 0, 0}, gridDim = {1, 1, 1};
typedef int cudaError_t;
static const cudaError_t cudaSuccess = 0;
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }

template <typename T>
static inline cudaError_t cudaMalloc(T** p, size_t n) { *p = static_cast<T*>(malloc(n)); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }

// 2-D grid of 1-D blocks whose kernel uses __shared__ + __syncthreads(): two sweeps per block
#define CUDA_EMU_LAUNCH_SHARED(GRID, BLOCK, CALL)                                   \
  do {                                                                              \
    gridDim.x = (GRID).x; gridDim.y = (GRID).y; blockDim.x = (BLOCK).x;             \
    for (blockIdx.y = 0; blockIdx.y < gridDim.y; ++blockIdx.y)                      \
      for (blockIdx.x = 0; blockIdx.x < gridDim.x; ++blockIdx.x)                    \
        for (int cuda_emu_sweep = 0; cuda_emu_sweep < 2; ++cuda_emu_sweep)          \
          for (threadIdx.x = 0; threadIdx.x < blockDim.x; ++threadIdx.x) {          \
            CALL;                                                                   \
          }                                                                         \
  } while (0)

#define CUDA_EMU_LAUNCH(G, T, CALL)                                         \
  do {                                                                      \
    gridDim.x = (G);                                                        \
    blockDim.x = (T);                                                       \
    for (blockIdx.x = 0; blockIdx.x < gridDim.x; ++blockIdx.x)              \
      for (threadIdx.x = 0; threadIdx.x < blockDim.x; ++threadIdx.x) {      \
        CALL;                                                               \
      }                                                                     \
  } while (0)
