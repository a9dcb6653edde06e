// SYNTAX-CHECK STUB, not TensorFlow. A minimal mock of the TF 1.13 C++ op API surface that shim/tf_custom_op.cc
// uses, so that the shim can be compiled (`g++ -fsyntax-only -Ishim/tf_stub -Iinclude -DBSL_TF_STUB`) in an image
// without TensorFlow headers (tests/test_shim_syntax.py). With real headers on the include path this directory is
// not used.
#pragma once
#include <cstdint>
#include <initializer_list>
#include <string>
#include <vector>

namespace Eigen { struct GpuDevice { void* stream() const { return nullptr; } }; struct half {}; }
namespace tensorflow {
typedef long long int64;
typedef unsigned char uint8;
struct bfloat16 { uint16_t value; };
enum DataType { DT_FLOAT, DT_DOUBLE, DT_INT32, DT_UINT8, DT_UINT32, DT_BFLOAT16, DT_INT64, DT_BOOL };
class Status {
 public:
  Status() {}
  static Status OK() { return Status(); }
  bool ok() const { return true; }
};
namespace errors {
template <typename... A> Status Internal(A...) { return Status(); }
template <typename... A> Status InvalidArgument(A...) { return Status(); }
}  // namespace errors
struct StringPiece { const char* data() const { return nullptr; } size_t size() const { return 0; } };
class TensorShape {
 public:
  TensorShape() {}
  TensorShape(std::initializer_list<int64>) {}
  void AddDim(int64) {}
  int dims() const { return 0; }
  int64 dim_size(int) const { return 0; }
  int64 num_elements() const { return 0; }
};
class Tensor {
 public:
  int dims() const { return 0; }
  int64 dim_size(int) const { return 0; }
  int64 NumElements() const { return 0; }
  const TensorShape& shape() const { static TensorShape s; return s; }
  StringPiece tensor_data() const { return StringPiece(); }
  template <typename T> struct Scalar { T operator()() const { return T(); } };
  template <typename T> Scalar<T> scalar() const { return Scalar<T>(); }
  DataType dtype() const { return DT_FLOAT; }
};
class OpKernelConstruction {
 public:
  template <typename T> Status GetAttr(const char*, T*) const { return Status(); }
  void CtxFailure(const Status&) {}
};
class OpKernelContext {
 public:
  const Tensor& input(int) { static Tensor t; return t; }
  int num_inputs() const { return 0; }
  Status allocate_output(int, const TensorShape&, Tensor**) { return Status(); }
  Status allocate_temp(DataType, const TensorShape&, Tensor*) { return Status(); }
  void set_output(int, const Tensor&) {}
  Status forward_input_or_allocate_output(std::initializer_list<int>, int, const TensorShape&, Tensor**) { return Status(); }
  const Eigen::GpuDevice& eigen_gpu_device() const { static Eigen::GpuDevice d; return d; }
  void CtxFailure(const Status&) {}
};
class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*) {}
  virtual ~OpKernel() {}
  virtual void Compute(OpKernelContext*) = 0;
};
namespace shape_inference {
class InferenceContext {};
inline Status UnknownShape(InferenceContext*) { return Status(); }
}  // namespace shape_inference
struct OpDefBuilderStub {
  OpDefBuilderStub& Input(const char*) { return *this; }
  OpDefBuilderStub& Output(const char*) { return *this; }
  OpDefBuilderStub& Attr(const char*) { return *this; }
  OpDefBuilderStub& Doc(const char*) { return *this; }
  template <typename F> OpDefBuilderStub& SetShapeFn(F) { return *this; }
};
struct KernelDefBuilderStub {
  KernelDefBuilderStub& Device(const char*) { return *this; }
  KernelDefBuilderStub& HostMemory(const char*) { return *this; }
  template <typename T> KernelDefBuilderStub& TypeConstraint(const char*) { return *this; }
};
inline KernelDefBuilderStub Name(const char*) { return KernelDefBuilderStub(); }
static const char* const DEVICE_GPU = "GPU";
}  // namespace tensorflow

#define BSL_STUB_CAT2(a, b) a##b
#define BSL_STUB_CAT(a, b) BSL_STUB_CAT2(a, b)
#define REGISTER_OP(name) static ::tensorflow::OpDefBuilderStub BSL_STUB_CAT(bsl_stub_op_, __COUNTER__) = ::tensorflow::OpDefBuilderStub()
#define REGISTER_KERNEL_BUILDER(builder, cls) \
  static ::tensorflow::KernelDefBuilderStub BSL_STUB_CAT(bsl_stub_kb_, __COUNTER__) = (builder); \
  static_assert(sizeof(cls) > 0, "kernel class")
#define OP_REQUIRES(ctx, cond, status) do { if (!(cond)) { (ctx)->CtxFailure(status); return; } } while (0)
#define OP_REQUIRES_OK(ctx, expr) do { ::tensorflow::Status _s = (expr); if (!_s.ok()) { (ctx)->CtxFailure(_s); return; } } while (0)
