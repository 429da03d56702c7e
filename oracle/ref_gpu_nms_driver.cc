// extern "C" door to the host build of the reference's nms/nms_kernel.cu (see
// oracle/build_ref.py:build_cuda_nms and oracle/tf_stub/cuda_emu.h).  TEST INFRASTRUCTURE.
#include "gpu_nms.hpp"   // the reference's header: declares _nms

extern "C" void ref_gpu_nms_sorted(int* keep_out, int* num_out, const float* boxes_host,
                                   int boxes_num, int boxes_dim, float thresh) {
  _nms(keep_out, num_out, boxes_host, boxes_num, boxes_dim, thresh, 0);
}
