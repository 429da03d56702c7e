// Driver for the host build of the reference's CUDA RoI-pool kernels (see
// oracle/build_ref.py:build_cuda_twin and oracle/tf_stub/cuda_emu.h).  TEST INFRASTRUCTURE.
#include "tf_stub.h"

// defined at global scope by roi_pooling_op_gpu.cu.cc:87-110
bool ROIPoolForwardLaucher(const float* bottom_data, const float spatial_scale, const int num_rois,
                           const int height, const int width, const int channels,
                           const int pooled_height, const int pooled_width,
                           const float* bottom_rois, float* top_data, int* argmax_data,
                           const Eigen::GpuDevice& d);

extern "C" int ref_roi_pool_fwd_cudatwin(const float* bottom, const float* rois, int H, int W, int C,
                                         int R, int PH, int PW, float scale, float* top,
                                         int* argmax) {
  Eigen::GpuDevice d;
  return ROIPoolForwardLaucher(bottom, scale, R, H, W, C, PH, PW, rois, top, argmax, d) ? 0 : -1;
}
