// Stand-in for the slice of the TensorFlow 1.x op framework that the reference's
// roi_pooling_layer/roi_pooling_op.cc touches.  TEST INFRASTRUCTURE ONLY.
//
// Purpose: compile the reference's UNMODIFIED roi_pooling_op.cc (from where it lies under
// /root/reference) into oracle/_ref/ref_roi_pool.so so that the RoiPool / RoiPoolGrad CPU
// kernels -- the parity target of the product's RoI-pool kernels -- can be run here without
// TensorFlow.  This header contains NO arithmetic of the op: it only provides containers
// (Tensor, TensorShape), the attribute / status plumbing, the registration macros (which
// record a factory per op name) and declarations; every number the kernels produce comes
// from the reference's own source.  Shard() (declared by the reference's work_sharder.h) is
// defined in oracle/ref_roi_pool_driver.cc as a plain multi-threaded range split.
#pragma once
#include <math.h>   // TensorFlow's headers make the C math functions visible at global scope

#include <cstdint>
#include <cstdio>
#include <functional>
#include <initializer_list>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace Eigen {
struct ThreadPoolDevice {};
struct GpuDevice {
  void* stream() const { return nullptr; }
  bool ok() const { return true; }
};
}  // namespace Eigen

namespace tensorflow {

typedef long long int64;
typedef int int32;

class Status {
 public:
  Status() : ok_(true) {}
  explicit Status(const std::string& m) : ok_(false), msg_(m) {}
  static Status OK() { return Status(); }
  bool ok() const { return ok_; }
  const std::string& error_message() const { return msg_; }

 private:
  bool ok_;
  std::string msg_;
};

namespace errors {
template <typename... A>
Status InvalidArgument(const A&... a) {
  std::ostringstream os;
  (void)std::initializer_list<int>{((os << a), 0)...};
  return Status(os.str());
}
}  // namespace errors

namespace thread {
class ThreadPool {};
}  // namespace thread

class TensorShape {
 public:
  std::vector<int64> d;
  int dims() const { return (int)d.size(); }
  int64 dim_size(int i) const { return d[i]; }
  int64 num_elements() const {
    int64 n = 1;
    for (int64 v : d) n *= v;
    return n;
  }
};

struct TensorShapeUtils {
  static Status MakeShape(const int* dims, int n, TensorShape* out) {
    out->d.assign(dims, dims + n);
    return Status::OK();
  }
};

template <typename T>
struct Flat {
  T* p;
  int64 n;
  T* data() const { return p; }
  T& operator()(int64 i) const { return p[i]; }
  int64 size() const { return n; }
};

// A tensor over caller-owned memory (the driver points it at numpy buffers).
class Tensor {
 public:
  Tensor() : ptr_(nullptr) {}
  Tensor(void* p, const TensorShape& s) : ptr_(p), shape_(s) {}
  template <typename T>
  Flat<T> flat() const { return Flat<T>{static_cast<T*>(ptr_), shape_.num_elements()}; }
  int dims() const { return shape_.dims(); }
  int64 dim_size(int i) const { return shape_.dim_size(i); }
  const TensorShape& shape() const { return shape_; }

 private:
  void* ptr_;
  TensorShape shape_;
};

class DeviceBase {
 public:
  struct CpuWorkerThreads {
    int num_threads;
    thread::ThreadPool* workers;
  };
  const CpuWorkerThreads* tensorflow_cpu_worker_threads() const { return &threads; }
  CpuWorkerThreads threads;
};

class OpKernelConstruction {
 public:
  std::map<std::string, double> attrs;
  template <typename T>
  Status GetAttr(const char* name, T* out) const {
    auto it = attrs.find(name);
    if (it == attrs.end()) return Status(std::string("no attr ") + name);
    *out = static_cast<T>(it->second);
    return Status::OK();
  }
  void SetStatus(const Status& s) { status_ = s; }
  const Status& status() const { return status_; }

 private:
  Status status_;
};

class OpKernelContext {
 public:
  std::vector<Tensor> inputs;
  std::vector<Tensor> outputs;        // pre-bound to caller memory by the driver
  std::vector<TensorShape> out_shapes;
  DeviceBase dev;
  const Tensor& input(int i) const { return inputs[i]; }
  Status allocate_output(int i, const TensorShape& shape, Tensor** out) {
    if (i >= (int)outputs.size()) return Status("no such output");
    if (shape.num_elements() != outputs[i].shape().num_elements())
      return Status("output buffer has the wrong size");
    outputs[i] = Tensor(outputs[i].template flat<char>().data(), shape);
    *out = &outputs[i];
    return Status::OK();
  }
  DeviceBase* device() { return &dev; }
  void SetStatus(const Status& s) { status_ = s; }
  const Status& status() const { return status_; }
  template <typename D>
  const D& eigen_device() const {
    static D d;
    return d;
  }

 private:
  Status status_;
};

class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*) {}
  virtual ~OpKernel() {}
  virtual void Compute(OpKernelContext* context) = 0;
};

#define OP_REQUIRES_OK(CTX, ...)                \
  do {                                          \
    ::tensorflow::Status _s(__VA_ARGS__);       \
    if (!_s.ok()) {                             \
      (CTX)->SetStatus(_s);                     \
      return;                                   \
    }                                           \
  } while (0)

#define OP_REQUIRES(CTX, EXP, STATUS)           \
  do {                                          \
    if (!(EXP)) {                               \
      (CTX)->SetStatus(STATUS);                 \
      return;                                   \
    }                                           \
  } while (0)

// ---- shape inference: declarations only (the shape function is registered, never called)
namespace shape_inference {
struct DimensionHandle {};
struct ShapeHandle {};
class InferenceContext {
 public:
  template <typename T>
  Status GetAttr(const char*, T*) const { return Status::OK(); }
  DimensionHandle MakeDim(int64) { return DimensionHandle(); }
  ShapeHandle MakeShape(std::initializer_list<DimensionHandle>) { return ShapeHandle(); }
  DimensionHandle Dim(ShapeHandle, int) { return DimensionHandle(); }
  ShapeHandle input(int) { return ShapeHandle(); }
  void set_output(int, ShapeHandle) {}
};
}  // namespace shape_inference

// ---- registration
struct OpDefBuilderStub {
  explicit OpDefBuilderStub(const char*) {}
  OpDefBuilderStub& Attr(const char*) { return *this; }
  OpDefBuilderStub& Input(const char*) { return *this; }
  OpDefBuilderStub& Output(const char*) { return *this; }
  OpDefBuilderStub& SetShapeFn(std::function<Status(shape_inference::InferenceContext*)>) {
    return *this;
  }
};

typedef OpKernel* (*KernelFactory)(OpKernelConstruction*);
inline std::map<std::string, KernelFactory>& kernel_registry() {
  static std::map<std::string, KernelFactory> r;
  return r;
}

static const char* const DEVICE_CPU = "CPU";
static const char* const DEVICE_GPU = "GPU";

struct Name {
  std::string name, device;
  explicit Name(const char* n) : name(n) {}
  Name& Device(const char* d) { device = d; return *this; }
  template <typename T>
  Name& TypeConstraint(const char*) { return *this; }
};

struct KernelRegistrar {
  KernelRegistrar(const Name& n, KernelFactory f) { kernel_registry()[n.name + "/" + n.device] = f; }
};

#define TF_STUB_CONCAT2(a, b) a##b
#define TF_STUB_CONCAT(a, b) TF_STUB_CONCAT2(a, b)
#define REGISTER_OP(name) \
  static ::tensorflow::OpDefBuilderStub TF_STUB_CONCAT(tf_stub_op_, __COUNTER__) = ::tensorflow::OpDefBuilderStub(name)
#define REGISTER_KERNEL_BUILDER(kb, ...)                                                    \
  static ::tensorflow::KernelRegistrar TF_STUB_CONCAT(tf_stub_kernel_, __COUNTER__)(        \
      kb, [](::tensorflow::OpKernelConstruction* c) -> ::tensorflow::OpKernel* {           \
        return new __VA_ARGS__(c);                                                          \
      })

}  // namespace tensorflow
