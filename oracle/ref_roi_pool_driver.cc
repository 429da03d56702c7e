// Driver around the reference's own RoiPool / RoiPoolGrad CPU kernels.  TEST INFRASTRUCTURE.
//
// oracle/build_ref.py compiles /root/reference/code/lib/roi_pooling_layer/roi_pooling_op.cc
// UNMODIFIED (against oracle/tf_stub, a stand-in for the TensorFlow op framework that holds
// no arithmetic) and links it with this file into oracle/_ref/ref_roi_pool.so.  The functions
// below build the op through the factory the reference's REGISTER_KERNEL_BUILDER lines
// registered, bind its inputs / outputs to the caller's buffers and call Compute().
#include <thread>

#include "tf_stub.h"
#include "work_sharder.h"   // the reference's header (declares Shard)

namespace tensorflow {
// roi_pooling_op.cc:198-203 hands the op's lambda to Shard(); TensorFlow's splits the range
// over its intra-op pool.  Any partition gives the same result (each output element is
// computed independently), so a plain even split over max_parallelism threads is used.
void Shard(int max_parallelism, thread::ThreadPool*, int64 total, int64,
           std::function<void(int64, int64)> work) {
  int n = max_parallelism < 1 ? 1 : max_parallelism;
  if ((int64)n > total) n = total > 0 ? (int)total : 1;
  if (n == 1) {
    work(0, total);
    return;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < n; ++t) {
    const int64 a = total * t / n, b = total * (t + 1) / n;
    th.emplace_back([=] { work(a, b); });
  }
  for (auto& t : th) t.join();
}
}  // namespace tensorflow

// the CUDA launchers are declared at file scope in roi_pooling_op.cc and referenced by its
// GPU-device specialisations, which are never registered without GOOGLE_CUDA
bool ROIPoolForwardLaucher(const float*, const float, const int, const int, const int, const int,
                           const int, const int, const float*, float*, int*,
                           const Eigen::GpuDevice&) { return false; }
bool ROIPoolBackwardLaucher(const float*, const float, const int, const int, const int, const int,
                            const int, const int, const int, const float*, float*, const int*,
                            const Eigen::GpuDevice&) { return false; }

using namespace tensorflow;

static TensorShape shape_of(std::initializer_list<int64> d) {
  TensorShape s;
  s.d.assign(d.begin(), d.end());
  return s;
}

static OpKernel* make_kernel(const char* name, int ph, int pw, float scale, std::string* err) {
  auto it = kernel_registry().find(std::string(name) + "/CPU");
  if (it == kernel_registry().end()) { *err = "kernel not registered"; return nullptr; }
  OpKernelConstruction c;
  c.attrs["pooled_height"] = ph;
  c.attrs["pooled_width"] = pw;
  c.attrs["spatial_scale"] = scale;
  OpKernel* k = it->second(&c);
  if (!c.status().ok()) { *err = c.status().error_message(); delete k; return nullptr; }
  return k;
}

extern "C" {

// returns 0, or -1 with a message in err (<= 255 chars)
int ref_roi_pool_fwd(const float* bottom, const float* rois, int B, int H, int W, int C, int R,
                     int PH, int PW, float scale, int threads, float* top, int* argmax,
                     char* err) {
  std::string e;
  OpKernel* k = make_kernel("RoiPool", PH, PW, scale, &e);
  if (k) {
    OpKernelContext ctx;
    ctx.dev.threads.num_threads = threads;
    ctx.dev.threads.workers = nullptr;
    ctx.inputs.emplace_back(const_cast<float*>(bottom), shape_of({B, H, W, C}));
    ctx.inputs.emplace_back(const_cast<float*>(rois), shape_of({R, 5}));
    ctx.outputs.emplace_back(top, shape_of({R, PH, PW, C}));
    ctx.outputs.emplace_back(argmax, shape_of({R, PH, PW, C}));
    k->Compute(&ctx);
    if (!ctx.status().ok()) e = ctx.status().error_message();
    delete k;
  }
  if (!e.empty()) { snprintf(err, 256, "%s", e.c_str()); return -1; }
  return 0;
}

int ref_roi_pool_bwd(const float* bottom, const float* rois, const int* argmax, const float* grad,
                     int B, int H, int W, int C, int R, int PH, int PW, float scale, int threads,
                     float* out, char* err) {
  std::string e;
  OpKernel* k = make_kernel("RoiPoolGrad", PH, PW, scale, &e);
  if (k) {
    OpKernelContext ctx;
    ctx.dev.threads.num_threads = threads;
    ctx.dev.threads.workers = nullptr;
    ctx.inputs.emplace_back(const_cast<float*>(bottom), shape_of({B, H, W, C}));
    ctx.inputs.emplace_back(const_cast<float*>(rois), shape_of({R, 5}));
    ctx.inputs.emplace_back(const_cast<int*>(argmax), shape_of({R, PH, PW, C}));
    ctx.inputs.emplace_back(const_cast<float*>(grad), shape_of({R, PH, PW, C}));
    ctx.outputs.emplace_back(out, shape_of({B, H, W, C}));
    k->Compute(&ctx);
    if (!ctx.status().ok()) e = ctx.status().error_message();
    delete k;
  }
  if (!e.empty()) { snprintf(err, 256, "%s", e.c_str()); return -1; }
  return 0;
}

}  // extern "C"
