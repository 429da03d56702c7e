// Host emulation of the few CUDA constructs the reference's roi_pooling_op_gpu.cu.cc uses, so
// that its kernel BODIES (the GPU_CEIL bin arithmetic, roi_pooling_op_gpu.cu.cc:19-85) can be
// executed on the CPU without a GPU or nvcc.  TEST INFRASTRUCTURE ONLY; no arithmetic here.
// oracle/build_ref.py rewrites the two `kernel<<<grid, block, smem, stream>>>(args);` launch
// statements into CUDA_EMU_LAUNCH(grid, block, kernel(args)); -- the only source edit -- which
// calls the kernel function once per (block, thread) with fresh by-value parameters, exactly
// like a launch (the kernels advance their pointer parameters inside the grid-stride loop).
#pragma once
#include <cstdio>
#include <cstdlib>

#define __global__
struct CudaEmuDim { int x, y, z; };
static CudaEmuDim blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, threadIdx = {0, 0, 0}, gridDim = {1, 1, 1};
typedef int cudaError_t;
static const cudaError_t cudaSuccess = 0;
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }

#define CUDA_EMU_LAUNCH(G, T, CALL)                                         \
  do {                                                                      \
    gridDim.x = (G);                                                        \
    blockDim.x = (T);                                                       \
    for (blockIdx.x = 0; blockIdx.x < gridDim.x; ++blockIdx.x)              \
      for (threadIdx.x = 0; threadIdx.x < blockDim.x; ++threadIdx.x) {      \
        CALL;                                                               \
      }                                                                     \
  } while (0)
