// TensorFlow 1.13 custom-op binding of libbsl_b200.so -- the boundary BASELINE.json's north_star names
// ("exposed to the reference's TensorFlow graph as a custom op").
//
//   TF_CFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_compile_flags()))')
//   TF_LFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_link_flags()))')
//   g++ -std=c++14 -shared -fPIC $TF_CFLAGS -Iinclude shim/tf_custom_op.cc -o libbsl_tf_ops.so
//       -Lboxsegliver_b200 -lbsl_b200 $TF_LFLAGS
//
// One REGISTER_OP + GPU OpKernel per op family of INTEGRATION.md's table; the Python side (shim/bsl_tf_ops.py) loads
// the library, attaches the gradients with @ops.RegisterGradient and offers the slim-signature layer functions.
// Rules of the boundary, as SURVEY.md section 8b prescribes for a TF op kernel:
//   * inputs / outputs are TF-owned device tensors (NHWC / NDHWC, bf16 activations, fp32 parameters); outputs and
//     workspaces come from ctx->allocate_output / allocate_temp -- the library never frees a TF buffer;
//   * every Compute only ENQUEUES on ctx->eigen_gpu_device().stream(): no host synchronisation, no default stream;
//   * `is_training` is a host-memory INPUT tensor (NetworksV2/base.py:77 feeds it per sess.run), not an attribute;
//   * errors surface through OP_REQUIRES with bsl_last_error(); nothing aborts;
//   * no mutable global state except the per-device library context (immutable after creation; TF's executor may call
//     different op instances from several inter-op threads, all on the one compute stream of the device).
// TensorFlow cannot be installed in this repo's image (Python 3.12; requirements.txt:2 pins tensorflow-gpu 1.13), so
// this file is compiled there only against shim/tf_stub (a mock of the API surface used, tests/test_shim_syntax.py);
// the same entry points are exercised for real through ctypes by every -m gpu test.
#if defined(BSL_TF_STUB) || __has_include("tensorflow/core/framework/op_kernel.h")

#include <mutex>

#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"

#include "bsl_b200.h"

namespace bsl_tf {
using namespace tensorflow;  // NOLINT

// One library context per process and device, created on first use (thread-safe), never destroyed: op kernels of
// several graphs share it, exactly like a cuDNN handle in TF's stream executor.
inline bsl_ctx* Context() {
  static std::once_flag once;
  static bsl_ctx* ctx = nullptr;
  std::call_once(once, [] {
    int dev = 0;  // TF has already made the op's device current; bsl_init(-1) would also accept "current device"
    if (bsl_init(dev, &ctx) != BSL_OK) ctx = nullptr;
  });
  return ctx;
}

inline const void* In(OpKernelContext* c, int i) { return c->input(i).tensor_data().data(); }
inline void* Out(Tensor* t) { return const_cast<char*>(t->tensor_data().data()); }
inline void* Stream(OpKernelContext* c) { return c->eigen_gpu_device().stream(); }
inline int Dim(const Tensor& t, int i) { return static_cast<int>(t.dim_size(i)); }

#define BSL_TF_CALL(c, expr)                                                             \
  do {                                                                                   \
    bsl_ctx* _ctx = ::bsl_tf::Context();                                                 \
    OP_REQUIRES(c, _ctx != nullptr, errors::Internal("bsl_init failed (no sm_100 GPU?)")); \
    const int _rc = (expr);                                                              \
    OP_REQUIRES(c, _rc == BSL_OK, errors::Internal(#expr, ": ", bsl_last_error(_ctx)));  \
  } while (0)

#define BSL_CTX ::bsl_tf::Context()

// Workspace of `bytes` bytes from TF's allocator, alive until the op's kernels have run (TF keeps temp tensors
// referenced by the stream until the enqueued work is done).
#define BSL_TF_WORKSPACE(c, var, bytes)                                                                  \
  Tensor var;                                                                                            \
  OP_REQUIRES_OK(c, c->allocate_temp(DT_UINT8, TensorShape({static_cast<int64>((bytes) ? (bytes) : 16)}), &var))

// ------------------------------------------------------------------------------------------------ Conv2D family
// slim.conv2d(x, C, 3) -> Conv2D / Conv2DBackpropInput / Conv2DBackpropFilter (NetworksV2/UNet.py:79,85,94).
static bsl_conv2d_desc Conv2dDesc(const Tensor& x, int cout, int kh, int kw) {
  return bsl_conv2d_desc{Dim(x, 0), Dim(x, 1), Dim(x, 2), Dim(x, 3), cout, kh, kw, Dim(x, 3), cout};
}

REGISTER_OP("BslConv2D").Input("x: bfloat16").Input("filter: bfloat16").Output("y: bfloat16")
    .Doc("SAME, stride 1, 3x3 or 1x1; filter HWIO (bf16 shadow of the fp32 variable).");
class BslConv2DOp : public OpKernel {
 public:
  explicit BslConv2DOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* c) override {
    const Tensor &x = c->input(0), &w = c->input(1);
    OP_REQUIRES(c, x.dims() == 4 && w.dims() == 4, errors::InvalidArgument("x NHWC, filter HWIO"));
    bsl_conv2d_desc d = Conv2dDesc(x, Dim(w, 3), Dim(w, 0), Dim(w, 1));
    Tensor* y = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, TensorShape({d.n, d.h, d.w, d.cout}), &y));
    BSL_TF_CALL(c, bsl_conv2d_fprop(BSL_CTX, &d, In(c, 0), In(c, 1), Out(y), Stream(c)));
  }
};
REGISTER_KERNEL_BUILDER(Name("BslConv2D").Device(DEVICE_GPU), BslConv2DOp);

// Conv2D + the reduction half of FusedBatchNorm from the same epilogue (sums fp64 [2][C]); `imgs_per_group` > 0 gives
// instance statistics [N / imgs_per_group][2][C] instead.
REGISTER_OP("BslConv2DStats").Input("x: bfloat16").Input("filter: bfloat16").Output("y: bfloat16").Output("sums: double")
    .Attr("imgs_per_group: int = 0");
class BslConv2DStatsOp : public OpKernel {
 public:
  explicit BslConv2DStatsOp(OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("imgs_per_group", &gi_)); }
  void Compute(OpKernelContext* c) override {
    const Tensor &x = c->input(0), &w = c->input(1);
    bsl_conv2d_desc d = Conv2dDesc(x, Dim(w, 3), Dim(w, 0), Dim(w, 1));
    Tensor *y = nullptr, *s = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, TensorShape({d.n, d.h, d.w, d.cout}), &y));
    const int groups = gi_ > 0 ? d.n / gi_ : 1;
    OP_REQUIRES_OK(c, c->allocate_output(1, TensorShape({groups, 2, d.cout}), &s));
    if (gi_ > 0)
      BSL_TF_CALL(c, bsl_conv2d_fprop_group_stats(BSL_CTX, &d, In(c, 0), In(c, 1), Out(y), gi_,
                                                  static_cast<double*>(Out(s)), Stream(c)));
    else
      BSL_TF_CALL(c, bsl_conv2d_fprop_stats(BSL_CTX, &d, In(c, 0), In(c, 1), Out(y), static_cast<double*>(Out(s)), Stream(c)));
  }
 private:
  int gi_ = 0;
};
REGISTER_KERNEL_BUILDER(Name("BslConv2DStats").Device(DEVICE_GPU), BslConv2DStatsOp);

REGISTER_OP("BslConv2DBackpropInput").Input("dy: bfloat16").Input("filter: bfloat16").Output("dx: bfloat16");
class BslConv2DBackpropInputOp : public OpKernel {
 public:
  explicit BslConv2DBackpropInputOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* c) override {
    const Tensor &dy = c->input(0), &w = c->input(1);
    bsl_conv2d_desc d{Dim(dy, 0), Dim(dy, 1), Dim(dy, 2), Dim(w, 2), Dim(w, 3), Dim(w, 0), Dim(w, 1), Dim(w, 2), Dim(w, 3)};
    Tensor* dx = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, TensorShape({d.n, d.h, d.w, d.cin}), &dx));
    BSL_TF_CALL(c, bsl_conv2d_dgrad(BSL_CTX, &d, In(c, 0), In(c, 1), Out(dx), Stream(c)));
  }
};
REGISTER_KERNEL_BUILDER(Name("BslConv2DBackpropInput").Device(DEVICE_GPU), BslConv2DBackpropInputOp);

REGISTER_OP("BslConv2DBackpropFilter").Input("x: bfloat16").Input("dy: bfloat16").Output("dw: float")
    .Attr("kh: int = 3").Attr("kw: int = 3");
class BslConv2DBackpropFilterOp : public OpKernel {
 public:
  explicit BslConv2DBackpropFilterOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("kh", &kh_));
    OP_REQUIRES_OK(c, c->GetAttr("kw", &kw_));
  }
  void Compute(OpKernelContext* c) override {
    const Tensor &x = c->input(0), &dy = c->input(1);
    bsl_conv2d_desc d = Conv2dDesc(x, Dim(dy, 3), kh_, kw_);
    Tensor* dw = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, TensorShape({kh_, kw_, d.cin, d.cout}), &dw));
    const size_t ws = bsl_conv2d_wgrad_workspace(BSL_CTX, &d);
    BSL_TF_WORKSPACE(c, tmp, ws);
    BSL_TF_CALL(c, bsl_conv2d_wgrad(BSL_CTX, &d, In(c, 0), In(c, 1), static_cast<float*>(Out(dw)), Out(&tmp), ws, Stream(c)));
  }
 private:
  int kh_ = 3, kw_ = 3;
};
REGISTER_KERNEL_BUILDER(Name("BslConv2DBackpropFilter").Device(DEVICE_GPU), BslConv2DBackpropFilterOp);

// The 3-channel stem: im2col rows (bf16) for the K = 64 GEMM (first slim.conv2d of UNet.py:79).
REGISTER_OP("BslStemIm2col").Input("images: float").Output("col: bfloat16").Attr("col_ld: int = 32");
class BslStemIm2colOp : public OpKernel {
 public:
  explicit BslStemIm2colOp(OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("col_ld", &ld_)); }
  void Compute(OpKernelContext* c) override {
    const Tensor& x = c->input(0);
    bsl_conv2d_desc d = Conv2dDesc(x, 64, 3, 3);
    Tensor* col = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, TensorShape({d.n, d.h, d.w, ld_}), &col));
    BSL_TF_CALL(c, bsl_stem_im2col_ld(BSL_CTX, &d, static_cast<const float*>(In(c, 0)), Out(col), ld_, Stream(c)));
  }
 private:
  int ld_ = 32;
};
REGISTER_KERNEL_BUILDER(Name("BslStemIm2col").Device(DEVICE_GPU), BslStemIm2colOp);

// ------------------------------------------------------------------------------------------------ logits layer (1x1 + bias)
REGISTER_OP("BslHeadConv").Input("x: bfloat16").Input("w: float").Input("bias: float").Output("logits: float");
class BslHeadConvOp : public OpKernel {
 public:
  explicit BslHeadConvOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* c) override {
    const Tensor &x = c->input(0), &w = c->input(1);
    bsl_conv2d_desc d = Conv2dDesc(x, Dim(w, w.dims() - 1), 1, 1);
    Tensor* y = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, TensorShape({d.n, d.h, d.w, d.cout}), &y));
    BSL_TF_CALL(c, bsl_conv2d_head_fprop(BSL_CTX, &d, In(c, 0), static_cast<const float*>(In(c, 1)),
                                         static_cast<const float*>(In(c, 2)), static_cast<float*>(Out(y)), Stream(c)));
  }
};
REGISTER_KERNEL_BUILDER(Name("BslHeadConv").Device(DEVICE_GPU), BslHeadConvOp);

REGISTER_OP("BslHeadConvGrad").Input("x: bfloat16").Input("w: float").Input("dlogits: float")
    .Output("dx: bfloat16").Output("dw: float").Output("dbias: float");
class BslHeadConvGradOp : public OpKernel {
 public:
  explicit BslHeadConvGradOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* c) override {
    const Tensor &x = c->input(0), &w = c->input(1);
    bsl_conv2d_desc d = Conv2dDesc(x, Dim(w, w.dims() - 1), 1, 1);
    Tensor *dx = nullptr, *dw = nullptr, *db = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, x.shape(), &dx));
    OP_REQUIRES_OK(c, c->allocate_output(1, w.shape(), &dw));
    OP_REQUIRES_OK(c, c->allocate_output(2, TensorShape({d.cout}), &db));
    const float* dl = static_cast<const float*>(In(c, 2));
    BSL_TF_CALL(c, bsl_conv2d_head_dgrad(BSL_CTX, &d, dl, static_cast<const float*>(In(c, 1)), Out(dx), Stream(c)));
    BSL_TF_CALL(c, bsl_conv2d_head_wgrad(BSL_CTX, &d, In(c, 0), dl, static_cast<float*>(Out(dw)),
                                         static_cast<float*>(Out(db)), Stream(c)));
  }
};
REGISTER_KERNEL_BUILDER(Name("BslHeadConvGrad").Device(DEVICE_GPU), BslHeadConvGradOp);

// ------------------------------------------------------------------------------------------------ conv2d_transpose 2x2 / 2
// slim.conv2d_transpose(x, C/2, 2, 2): Conv2DBackpropInput-as-forward + BiasAdd + Relu (UNet.py:91).
static bsl_convT2d_desc ConvTDesc(const Tensor& x, int cout) {
  return bsl_convT2d_desc{Dim(x, 0), Dim(x, 1), Dim(x, 2), Dim(x, 3), cout, Dim(x, 3), cout, 1};
}
REGISTER_OP("BslConv2DTranspose").Input("x: bfloat16").Input("filter: bfloat16").Input("bias: float").Output("y: bfloat16");
class BslConv2DTransposeOp : public OpKernel {
 public:
  explicit BslConv2DTransposeOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* c) override {
    const Tensor &x = c->input(0), &w = c->input(1);   // filter [2,2,Cout,Cin]
    bsl_convT2d_desc d = ConvTDesc(x, Dim(w, 2));
    Tensor* y = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, TensorShape({d.n, 2 * d.h, 2 * d.w, d.cout}), &y));
    BSL_TF_CALL(c, bsl_convT2d_fwd(BSL_CTX, &d, In(c, 0), In(c, 1), static_cast<const float*>(In(c, 2)), Out(y), Stream(c)));
  }
};
REGISTER_KERNEL_BUILDER(Name("BslConv2DTranspose").Device(DEVICE_GPU), BslConv2DTransposeOp);

// Gradient of the above given y (for the ReLU mask) and dy: ReluGrad + BiasAddGrad in one pass, then data / filter.
REGISTER_OP("BslConv2DTransposeGrad").Input("x: bfloat16").Input("filter: bfloat16").Input("y: bfloat16").Input("dy: bfloat16")
    .Output("dx: bfloat16").Output("dw: float").Output("dbias: float");
class BslConv2DTransposeGradOp : public OpKernel {
 public:
  explicit BslConv2DTransposeGradOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* c) override {
    const Tensor &x = c->input(0), &w = c->input(1), &y = c->input(2);
    bsl_convT2d_desc d = ConvTDesc(x, Dim(w, 2));
    Tensor *dx = nullptr, *dw = nullptr, *db = nullptr, dyr;
    OP_REQUIRES_OK(c, c->allocate_output(0, x.shape(), &dx));
    OP_REQUIRES_OK(c, c->allocate_output(1, w.shape(), &dw));
    OP_REQUIRES_OK(c, c->allocate_output(2, TensorShape({d.cout}), &db));
    OP_REQUIRES_OK(c, c->allocate_temp(DT_BFLOAT16, y.shape(), &dyr));
    const long long px = static_cast<long long>(d.n) * 4 * d.h * d.w;
    BSL_TF_CALL(c, bsl_relu_bwd_bias(BSL_CTX, px, d.cout, In(c, 2), d.cout, In(c, 3), d.cout, Out(&dyr), d.cout,
                                     static_cast<float*>(Out(db)), Stream(c)));
    BSL_TF_CALL(c, bsl_convT2d_bwd_data(BSL_CTX, &d, Out(&dyr), In(c, 1), Out(dx), Stream(c)));
    const size_t ws = bsl_convT2d_bwd_filter_workspace(BSL_CTX, &d);
    BSL_TF_WORKSPACE(c, tmp, ws);
    BSL_TF_CALL(c, bsl_convT2d_bwd_filter(BSL_CTX, &d, In(c, 0), Out(&dyr), static_cast<float*>(Out(dw)), nullptr,
                                          Out(&tmp), ws, Stream(c)));
  }
};
REGISTER_KERNEL_BUILDER(Name("BslConv2DTransposeGrad").Device(DEVICE_GPU), BslConv2DTransposeGradOp);

// ------------------------------------------------------------------------------------------------ Conv3D family (UNet3D)
struct Conv3dAttrs {
  int k[3] = {3, 3, 3}, s[3] = {1, 1, 1};
  void Read(OpKernelConstruction* c) {
    std::vector<int> k_, s_;
    OP_REQUIRES_OK(c, c->GetAttr("ksize", &k_));
    OP_REQUIRES_OK(c, c->GetAttr("strides", &s_));
    for (int i = 0; i < 3 && i < static_cast<int>(k_.size()); ++i) k[i] = k_[i];
    for (int i = 0; i < 3 && i < static_cast<int>(s_.size()); ++i) s[i] = s_[i];
  }
  bsl_conv3d_desc Desc(int n, int d, int h, int w, int cin, int cout) const {
    return bsl_conv3d_desc{n, d, h, w, cin, cout, k[0], k[1], k[2], s[0], s[1], s[2], cin, cout};
  }
};
static int OutDim(int in, int stride) { return (in + stride - 1) / stride; }  // SAME

REGISTER_OP("BslConv3D").Input("x: bfloat16").Input("filter: bfloat16").Output("y: bfloat16")
    .Attr("ksize: list(int)").Attr("strides: list(int)");
class BslConv3DOp : public OpKernel {
 public:
  explicit BslConv3DOp(OpKernelConstruction* c) : OpKernel(c) { a_.Read(c); }
  void Compute(OpKernelContext* c) override {
    const Tensor &x = c->input(0), &w = c->input(1);   // NDHWC, DHWIO
    bsl_conv3d_desc d = a_.Desc(Dim(x, 0), Dim(x, 1), Dim(x, 2), Dim(x, 3), Dim(x, 4), Dim(w, 4));
    Tensor* y = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, TensorShape({d.n, OutDim(d.d, d.sd), OutDim(d.h, d.sh), OutDim(d.w, d.sw), d.cout}), &y));
    BSL_TF_CALL(c, bsl_conv3d_fprop(BSL_CTX, &d, In(c, 0), In(c, 1), Out(y), Stream(c)));
  }
 private:
  Conv3dAttrs a_;
};
REGISTER_KERNEL_BUILDER(Name("BslConv3D").Device(DEVICE_GPU), BslConv3DOp);

REGISTER_OP("BslConv3DGrad").Input("x: bfloat16").Input("filter: bfloat16").Input("dy: bfloat16")
    .Output("dx: bfloat16").Output("dw: float").Attr("ksize: list(int)").Attr("strides: list(int)");
class BslConv3DGradOp : public OpKernel {
 public:
  explicit BslConv3DGradOp(OpKernelConstruction* c) : OpKernel(c) { a_.Read(c); }
  void Compute(OpKernelContext* c) override {
    const Tensor &x = c->input(0), &w = c->input(1);
    bsl_conv3d_desc d = a_.Desc(Dim(x, 0), Dim(x, 1), Dim(x, 2), Dim(x, 3), Dim(x, 4), Dim(w, 4));
    Tensor *dx = nullptr, *dw = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, x.shape(), &dx));
    OP_REQUIRES_OK(c, c->allocate_output(1, w.shape(), &dw));
    BSL_TF_CALL(c, bsl_conv3d_dgrad(BSL_CTX, &d, In(c, 2), In(c, 1), Out(dx), Stream(c)));
    const size_t ws = bsl_conv3d_wgrad_workspace(BSL_CTX, &d);
    BSL_TF_WORKSPACE(c, tmp, ws);
    BSL_TF_CALL(c, bsl_conv3d_wgrad(BSL_CTX, &d, In(c, 0), In(c, 2), static_cast<float*>(Out(dw)), Out(&tmp), ws, Stream(c)));
  }
 private:
  Conv3dAttrs a_;
};
REGISTER_KERNEL_BUILDER(Name("BslConv3DGrad").Device(DEVICE_GPU), BslConv3DGradOp);

// slim.conv3d_transpose(x, c, kernel == stride, biases_initializer=None) + Relu (UNet3D.py:160-162)
REGISTER_OP("BslConv3DTranspose").Input("x: bfloat16").Input("filter: bfloat16").Output("y: bfloat16").Attr("sd: int = 1");
class BslConv3DTransposeOp : public OpKernel {
 public:
  explicit BslConv3DTransposeOp(OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("sd", &sd_)); }
  void Compute(OpKernelContext* c) override {
    const Tensor &x = c->input(0), &w = c->input(1);   // filter [kd,2,2,Cout,Cin]
    bsl_convT3d_desc d{Dim(x, 0), Dim(x, 1), Dim(x, 2), Dim(x, 3), Dim(x, 4), Dim(w, 3), sd_, Dim(x, 4), Dim(w, 3), 1};
    Tensor* y = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, TensorShape({d.n, sd_ * d.d, 2 * d.h, 2 * d.w, d.cout}), &y));
    BSL_TF_CALL(c, bsl_convT3d_fwd(BSL_CTX, &d, In(c, 0), In(c, 1), nullptr, Out(y), Stream(c)));
  }
 private:
  int sd_ = 1;
};
REGISTER_KERNEL_BUILDER(Name("BslConv3DTranspose").Device(DEVICE_GPU), BslConv3DTransposeOp);

REGISTER_OP("BslConv3DTransposeGrad").Input("x: bfloat16").Input("filter: bfloat16").Input("y: bfloat16").Input("dy: bfloat16")
    .Output("dx: bfloat16").Output("dw: float").Attr("sd: int = 1");
class BslConv3DTransposeGradOp : public OpKernel {
 public:
  explicit BslConv3DTransposeGradOp(OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("sd", &sd_)); }
  void Compute(OpKernelContext* c) override {
    const Tensor &x = c->input(0), &w = c->input(1), &y = c->input(2);
    bsl_convT3d_desc d{Dim(x, 0), Dim(x, 1), Dim(x, 2), Dim(x, 3), Dim(x, 4), Dim(w, 3), sd_, Dim(x, 4), Dim(w, 3), 1};
    Tensor *dx = nullptr, *dw = nullptr, dyr;
    OP_REQUIRES_OK(c, c->allocate_output(0, x.shape(), &dx));
    OP_REQUIRES_OK(c, c->allocate_output(1, w.shape(), &dw));
    OP_REQUIRES_OK(c, c->allocate_temp(DT_BFLOAT16, y.shape(), &dyr));
    BSL_TF_CALL(c, bsl_relu_bwd(BSL_CTX, y.NumElements() / d.cout, d.cout, In(c, 2), d.cout, In(c, 3), d.cout, Out(&dyr),
                                d.cout, Stream(c)));
    BSL_TF_CALL(c, bsl_convT3d_bwd_data(BSL_CTX, &d, Out(&dyr), In(c, 1), Out(dx), Stream(c)));
    const size_t ws = bsl_convT3d_bwd_filter_workspace(BSL_CTX, &d);
    BSL_TF_WORKSPACE(c, tmp, ws);
    BSL_TF_CALL(c, bsl_convT3d_bwd_filter(BSL_CTX, &d, In(c, 0), Out(&dyr), static_cast<float*>(Out(dw)), nullptr,
                                          Out(&tmp), ws, Stream(c)));
  }
 private:
  int sd_ = 1;
};
REGISTER_KERNEL_BUILDER(Name("BslConv3DTransposeGrad").Device(DEVICE_GPU), BslConv3DTransposeGradOp);

// ------------------------------------------------------------------------------------------------ normalisation (+ReLU, +pool)
// FusedBatchNorm / instance_norm + Relu (+ MaxPool) -- NetworksV2/base.py:153-169, UNet.py:81. Forward in one op:
// sums (from BslConv2DStats, or computed here when absent) -> finalize (updates the moving statistics in place when
// is_training) -> apply (+pool). Saved for backward: mean, rstd, scale, shift.
struct NormAttrs {
  int mode = 0, relu = 1, center = 1, scale = 1, pool = 0;
  float eps = 1e-3f, decay = 0.999f;
  void Read(OpKernelConstruction* c) {
    OP_REQUIRES_OK(c, c->GetAttr("mode", &mode));
    OP_REQUIRES_OK(c, c->GetAttr("epsilon", &eps));
    OP_REQUIRES_OK(c, c->GetAttr("decay", &decay));
    OP_REQUIRES_OK(c, c->GetAttr("center", &center));
    OP_REQUIRES_OK(c, c->GetAttr("scale", &scale));
  }
  bsl_norm_desc Desc(const Tensor& y) const {
    const int n = Dim(y, 0), ch = Dim(y, y.dims() - 1);
    const int hw = static_cast<int>(y.NumElements() / n / ch);
    return bsl_norm_desc{mode, n, hw, ch, ch, ch, eps, decay, relu, center, scale};
  }
};

REGISTER_OP("BslNormRelu")
    .Input("y: bfloat16").Input("sums: double").Input("gamma: float").Input("beta: float")
    .Input("moving_mean: Ref(float)").Input("moving_variance: Ref(float)").Input("is_training: bool")
    .Output("a: bfloat16").Output("pooled: bfloat16").Output("mean: float").Output("rstd: float")
    .Output("scale_out: float").Output("shift_out: float")
    .Attr("mode: int = 0").Attr("epsilon: float = 0.001").Attr("decay: float = 0.999").Attr("center: int = 1")
    .Attr("scale: int = 1").Attr("pool: int = 0");
class BslNormReluOp : public OpKernel {
 public:
  explicit BslNormReluOp(OpKernelConstruction* c) : OpKernel(c) {
    a_.Read(c);
    OP_REQUIRES_OK(c, c->GetAttr("pool", &a_.pool));
  }
  void Compute(OpKernelContext* c) override {
    const Tensor& y = c->input(0);
    bsl_norm_desc d = a_.Desc(y);
    const bool training = c->input(6).scalar<bool>()();     // host memory
    const int groups = d.mode ? d.n : 1;
    Tensor *a = nullptr, *pooled = nullptr, *stat[4] = {nullptr, nullptr, nullptr, nullptr};
    OP_REQUIRES_OK(c, c->allocate_output(0, y.shape(), &a));
    TensorShape ps = a_.pool ? TensorShape({d.n, Dim(y, 1) / 2, Dim(y, 2) / 2, d.c}) : TensorShape({0});
    OP_REQUIRES_OK(c, c->allocate_output(1, ps, &pooled));
    for (int i = 0; i < 4; ++i) OP_REQUIRES_OK(c, c->allocate_output(2 + i, TensorShape({groups, d.c}), &stat[i]));
    float* f[4];
    for (int i = 0; i < 4; ++i) f[i] = static_cast<float*>(Out(stat[i]));
    BSL_TF_CALL(c, bsl_norm_finalize(BSL_CTX, &d, training ? 1 : 0, static_cast<const double*>(In(c, 1)),
                                     static_cast<const float*>(In(c, 2)), static_cast<const float*>(In(c, 3)),
                                     static_cast<float*>(const_cast<void*>(In(c, 4))),
                                     static_cast<float*>(const_cast<void*>(In(c, 5))), f[0], f[1], f[2], f[3], Stream(c)));
    if (a_.pool)
      BSL_TF_CALL(c, bsl_norm_apply_pool(BSL_CTX, &d, Dim(y, 1), Dim(y, 2), In(c, 0), f[2], f[3], Out(a), Out(pooled),
                                         d.c, Stream(c)));
    else
      BSL_TF_CALL(c, bsl_norm_apply(BSL_CTX, &d, In(c, 0), f[2], f[3], Out(a), Stream(c)));
  }
 private:
  NormAttrs a_;
};
REGISTER_KERNEL_BUILDER(Name("BslNormRelu").Device(DEVICE_GPU).HostMemory("is_training"), BslNormReluOp);

REGISTER_OP("BslNormStats").Input("y: bfloat16").Output("sums: double").Attr("mode: int = 0");
class BslNormStatsOp : public OpKernel {
 public:
  explicit BslNormStatsOp(OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("mode", &a_.mode)); }
  void Compute(OpKernelContext* c) override {
    bsl_norm_desc d = a_.Desc(c->input(0));
    Tensor* s = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, TensorShape({d.mode ? d.n : 1, 2, d.c}), &s));
    BSL_TF_CALL(c, bsl_norm_stats(BSL_CTX, &d, In(c, 0), static_cast<double*>(Out(s)), Stream(c)));
  }
 private:
  NormAttrs a_;
};
REGISTER_KERNEL_BUILDER(Name("BslNormStats").Device(DEVICE_GPU), BslNormStatsOp);

// FusedBatchNormGrad + ReluGrad: da (gradient w.r.t. the activation) -> dy (w.r.t. the conv output), dgamma, dbeta.
REGISTER_OP("BslNormReluGrad")
    .Input("y: bfloat16").Input("da: bfloat16").Input("mean: float").Input("rstd: float").Input("scale_in: float")
    .Input("shift_in: float").Output("dy: bfloat16").Output("dgamma: float").Output("dbeta: float")
    .Attr("mode: int = 0").Attr("epsilon: float = 0.001").Attr("decay: float = 0.999").Attr("center: int = 1")
    .Attr("scale: int = 1");
class BslNormReluGradOp : public OpKernel {
 public:
  explicit BslNormReluGradOp(OpKernelConstruction* c) : OpKernel(c) { a_.Read(c); }
  void Compute(OpKernelContext* c) override {
    const Tensor& y = c->input(0);
    bsl_norm_desc d = a_.Desc(y);
    const int groups = d.mode ? d.n : 1;
    Tensor *dy = nullptr, *dg = nullptr, *db = nullptr, sums, c12;
    OP_REQUIRES_OK(c, c->allocate_output(0, y.shape(), &dy));
    OP_REQUIRES_OK(c, c->allocate_output(1, TensorShape({d.c}), &dg));
    OP_REQUIRES_OK(c, c->allocate_output(2, TensorShape({d.c}), &db));
    OP_REQUIRES_OK(c, c->allocate_temp(DT_DOUBLE, TensorShape({groups, 2, d.c}), &sums));
    OP_REQUIRES_OK(c, c->allocate_temp(DT_FLOAT, TensorShape({2, groups, d.c}), &c12));
    const float *mean = static_cast<const float*>(In(c, 2)), *rstd = static_cast<const float*>(In(c, 3)),
                *sc = static_cast<const float*>(In(c, 4)), *sh = static_cast<const float*>(In(c, 5));
    float* c1 = static_cast<float*>(Out(&c12));
    float* c2 = c1 + static_cast<size_t>(groups) * d.c;
    BSL_TF_CALL(c, bsl_norm_bwd_reduce(BSL_CTX, &d, In(c, 0), In(c, 1), d.c, mean, rstd, sc, sh,
                                       static_cast<double*>(Out(&sums)), Stream(c)));
    BSL_TF_CALL(c, bsl_norm_bwd_finalize(BSL_CTX, &d, static_cast<const double*>(Out(&sums)), c1, c2,
                                         static_cast<float*>(Out(dg)), static_cast<float*>(Out(db)), Stream(c)));
    BSL_TF_CALL(c, bsl_norm_bwd_apply(BSL_CTX, &d, In(c, 0), In(c, 1), d.c, mean, rstd, sc, sh, c1, c2, Out(dy), d.c, Stream(c)));
  }
 private:
  NormAttrs a_;
};
REGISTER_KERNEL_BUILDER(Name("BslNormReluGrad").Device(DEVICE_GPU), BslNormReluGradOp);

// MaxPoolGrad + AddN(skip gradient): UNet.py:81,91-93
REGISTER_OP("BslMaxPoolGradAdd").Input("act: bfloat16").Input("dpool: bfloat16").Input("dskip: bfloat16").Output("dact: bfloat16");
class BslMaxPoolGradAddOp : public OpKernel {
 public:
  explicit BslMaxPoolGradAddOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* c) override {
    const Tensor& a = c->input(0);
    Tensor* out = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, a.shape(), &out));
    const int ch = Dim(a, 3);
    BSL_TF_CALL(c, bsl_maxpool2x2_bwd_add(BSL_CTX, Dim(a, 0), Dim(a, 1), Dim(a, 2), ch, In(c, 0), ch, In(c, 1), ch,
                                          In(c, 2), ch, Out(out), ch, Stream(c)));
  }
};
REGISTER_KERNEL_BUILDER(Name("BslMaxPoolGradAdd").Device(DEVICE_GPU), BslMaxPoolGradAddOp);

// ------------------------------------------------------------------------------------------------ losses, masks, counts
struct LossAttrs {
  int weight_type = 0;
  std::vector<float> numeric_w;
  float proportion_decay = 1000.f, loss_scale = 1.f;
  void Read(OpKernelConstruction* c) {
    OP_REQUIRES_OK(c, c->GetAttr("weight_type", &weight_type));
    OP_REQUIRES_OK(c, c->GetAttr("numeric_w", &numeric_w));
    OP_REQUIRES_OK(c, c->GetAttr("proportion_decay", &proportion_decay));
    OP_REQUIRES_OK(c, c->GetAttr("loss_scale", &loss_scale));
  }
  bsl_loss_desc Desc(const Tensor& logits) const {
    bsl_loss_desc d = {};
    d.n = Dim(logits, 0);
    d.classes = Dim(logits, logits.dims() - 1);
    d.hw = static_cast<int>(logits.NumElements() / d.n / d.classes);
    d.weight_type = weight_type;
    for (size_t i = 0; i < numeric_w.size() && i < 8; ++i) d.numeric_w[i] = numeric_w[i];
    d.proportion_decay = proportion_decay;
    d.loss_scale = loss_scale;
    return d;
  }
};
#define BSL_LOSS_ATTRS \
  .Attr("weight_type: int = 0").Attr("numeric_w: list(float) = []").Attr("proportion_decay: float = 1000.0") \
  .Attr("loss_scale: float = 1.0")

// loss_metrics.weighted_sparse_softmax_cross_entropy (+ its gradient): loss_metrics.py:115-177
REGISTER_OP("BslWeightedXent").Input("logits: float").Input("labels: int32").Output("loss: float").Output("dlogits: float")
    BSL_LOSS_ATTRS;
class BslWeightedXentOp : public OpKernel {
 public:
  explicit BslWeightedXentOp(OpKernelConstruction* c) : OpKernel(c) { a_.Read(c); }
  void Compute(OpKernelContext* c) override {
    const Tensor& lg = c->input(0);
    bsl_loss_desc d = a_.Desc(lg);
    Tensor *loss = nullptr, *dl = nullptr, counts;
    OP_REQUIRES_OK(c, c->allocate_output(0, TensorShape({}), &loss));
    OP_REQUIRES_OK(c, c->allocate_output(1, lg.shape(), &dl));
    OP_REQUIRES_OK(c, c->allocate_temp(DT_INT32, TensorShape({d.n, d.classes}), &counts));
    const size_t ws = bsl_loss_workspace(BSL_CTX, &d);
    BSL_TF_WORKSPACE(c, tmp, ws);
    const int* labels = static_cast<const int*>(In(c, 1));
    BSL_TF_CALL(c, bsl_label_counts(BSL_CTX, &d, labels, static_cast<int*>(Out(&counts)), Stream(c)));
    BSL_TF_CALL(c, bsl_wxent_fwd_bwd(BSL_CTX, &d, static_cast<const float*>(In(c, 0)), labels,
                                     static_cast<const int*>(Out(&counts)), static_cast<float*>(Out(loss)),
                                     static_cast<float*>(Out(dl)), Out(&tmp), ws, Stream(c)));
  }
 private:
  LossAttrs a_;
};
REGISTER_KERNEL_BUILDER(Name("BslWeightedXent").Device(DEVICE_GPU), BslWeightedXentOp);

// loss_metrics.sparse_dice_loss: loss_metrics.py:180-231
REGISTER_OP("BslDiceLoss").Input("logits: float").Input("labels: int32").Output("loss: float").Output("dlogits: float")
    BSL_LOSS_ATTRS;
class BslDiceLossOp : public OpKernel {
 public:
  explicit BslDiceLossOp(OpKernelConstruction* c) : OpKernel(c) { a_.Read(c); }
  void Compute(OpKernelContext* c) override {
    const Tensor& lg = c->input(0);
    bsl_loss_desc d = a_.Desc(lg);
    Tensor *loss = nullptr, *dl = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, TensorShape({}), &loss));
    OP_REQUIRES_OK(c, c->allocate_output(1, lg.shape(), &dl));
    const size_t ws = bsl_loss_workspace(BSL_CTX, &d);
    BSL_TF_WORKSPACE(c, tmp, ws);
    BSL_TF_CALL(c, bsl_dice_fwd_bwd(BSL_CTX, &d, static_cast<const float*>(In(c, 0)), static_cast<const int*>(In(c, 1)),
                                    static_cast<float*>(Out(loss)), static_cast<float*>(Out(dl)), 0, Out(&tmp), ws, Stream(c)));
  }
 private:
  LossAttrs a_;
};
REGISTER_KERNEL_BUILDER(Name("BslDiceLoss").Device(DEVICE_GPU), BslDiceLossOp);

// slim.softmax + `p > 0.5` uint8 masks + argmax + integer Dice sums (UNet.py:107-117, loss_metrics.py:261-339)
REGISTER_OP("BslSoftmaxThreshold").Input("logits: float").Input("labels: int32")
    .Output("prob: float").Output("masks: uint8").Output("argmax: uint8").Output("ilr: uint32");
class BslSoftmaxThresholdOp : public OpKernel {
 public:
  explicit BslSoftmaxThresholdOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* c) override {
    const Tensor& lg = c->input(0);
    LossAttrs a;
    bsl_loss_desc d = a.Desc(lg);
    Tensor *prob = nullptr, *masks = nullptr, *am = nullptr, *ilr = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, lg.shape(), &prob));
    OP_REQUIRES_OK(c, c->allocate_output(1, TensorShape({d.classes - 1, d.n, d.hw}), &masks));
    OP_REQUIRES_OK(c, c->allocate_output(2, TensorShape({d.n, d.hw}), &am));
    OP_REQUIRES_OK(c, c->allocate_output(3, TensorShape({d.n, d.classes - 1, 3}), &ilr));
    BSL_TF_CALL(c, bsl_softmax_threshold(BSL_CTX, &d, static_cast<const float*>(In(c, 0)), static_cast<const int*>(In(c, 1)),
                                         static_cast<float*>(Out(prob)), static_cast<uint8_t*>(Out(masks)),
                                         static_cast<uint8_t*>(Out(am)), static_cast<unsigned int*>(Out(ilr)), Stream(c)));
  }
};
REGISTER_KERNEL_BUILDER(Name("BslSoftmaxThreshold").Device(DEVICE_GPU), BslSoftmaxThresholdOp);

// ------------------------------------------------------------------------------------------------ optimizer, all-reduce
// ApplyAdam / AdamW + L2 + bf16 shadow over a flat arena (core/solver.py:204-219, NetworksV2/base.py:128-135)
REGISTER_OP("BslApplyAdam")
    .Input("var: Ref(float)").Input("m: Ref(float)").Input("v: Ref(float)").Input("shadow: Ref(bfloat16)")
    .Input("grad: float").Input("lr: float").Input("step: int32").Output("sumsq: double")
    .Attr("beta1: float = 0.9").Attr("beta2: float = 0.99").Attr("epsilon: float = 1e-8").Attr("l2_rate: float = 0.0")
    .Attr("decoupled_decay: float = 0.0");
class BslApplyAdamOp : public OpKernel {
 public:
  explicit BslApplyAdamOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("beta1", &b1_));
    OP_REQUIRES_OK(c, c->GetAttr("beta2", &b2_));
    OP_REQUIRES_OK(c, c->GetAttr("epsilon", &eps_));
    OP_REQUIRES_OK(c, c->GetAttr("l2_rate", &l2_));
    OP_REQUIRES_OK(c, c->GetAttr("decoupled_decay", &wd_));
  }
  void Compute(OpKernelContext* c) override {
    const Tensor& var = c->input(0);
    Tensor* sq = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, TensorShape({}), &sq));
    bsl_adam_desc d{c->input(5).scalar<float>()(), b1_, b2_, eps_, l2_, 1.f, c->input(6).scalar<int>()(), wd_};
    BSL_TF_CALL(c, bsl_adam_step(BSL_CTX, &d, static_cast<float*>(const_cast<void*>(In(c, 0))),
                                 static_cast<const float*>(In(c, 4)), static_cast<float*>(const_cast<void*>(In(c, 1))),
                                 static_cast<float*>(const_cast<void*>(In(c, 2))), const_cast<void*>(In(c, 3)),
                                 static_cast<size_t>(var.NumElements()), static_cast<double*>(Out(sq)), Stream(c)));
  }
 private:
  float b1_ = 0.9f, b2_ = 0.99f, eps_ = 1e-8f, l2_ = 0.f, wd_ = 0.f;
};
REGISTER_KERNEL_BUILDER(Name("BslApplyAdam").Device(DEVICE_GPU).HostMemory("lr").HostMemory("step"), BslApplyAdamOp);

REGISTER_OP("BslApplyMomentum")
    .Input("var: Ref(float)").Input("accum: Ref(float)").Input("shadow: Ref(bfloat16)").Input("grad: float")
    .Input("lr: float").Output("sumsq: double")
    .Attr("momentum: float = 0.9").Attr("use_nesterov: bool = false").Attr("l2_rate: float = 0.0");
class BslApplyMomentumOp : public OpKernel {
 public:
  explicit BslApplyMomentumOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("momentum", &mom_));
    OP_REQUIRES_OK(c, c->GetAttr("use_nesterov", &nesterov_));
    OP_REQUIRES_OK(c, c->GetAttr("l2_rate", &l2_));
  }
  void Compute(OpKernelContext* c) override {
    Tensor* sq = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, TensorShape({}), &sq));
    BSL_TF_CALL(c, bsl_momentum_step(BSL_CTX, c->input(4).scalar<float>()(), mom_, nesterov_ ? 1 : 0, l2_, 1.f,
                                     static_cast<float*>(const_cast<void*>(In(c, 0))), static_cast<const float*>(In(c, 3)),
                                     static_cast<float*>(const_cast<void*>(In(c, 1))), const_cast<void*>(In(c, 2)),
                                     static_cast<size_t>(c->input(0).NumElements()), static_cast<double*>(Out(sq)), Stream(c)));
  }
 private:
  float mom_ = 0.9f, l2_ = 0.f;
  bool nesterov_ = false;
};
REGISTER_KERNEL_BUILDER(Name("BslApplyMomentum").Device(DEVICE_GPU).HostMemory("lr"), BslApplyMomentumOp);

// NcclAllReduce under MirroredStrategy (utils/distribution_utils.py:85-98): in-place sum of a flat gradient bucket.
// The communicator is created once per process by BslCommInit (rank / world / 128-byte id from the host side channel).
REGISTER_OP("BslAllReduceSum").Input("bucket: Ref(float)").Output("done: int32");
class BslAllReduceSumOp : public OpKernel {
 public:
  explicit BslAllReduceSumOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* c) override {
    Tensor* done = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, TensorShape({}), &done));
    BSL_TF_CALL(c, bsl_allreduce_sum_f32(BSL_CTX, static_cast<float*>(const_cast<void*>(In(c, 0))),
                                         static_cast<size_t>(c->input(0).NumElements()), Stream(c)));
  }
};
REGISTER_KERNEL_BUILDER(Name("BslAllReduceSum").Device(DEVICE_GPU), BslAllReduceSumOp);

// ------------------------------------------------------------------------------------------------ GUNet side networks
// slim_nets.fc (fully_connected + ReLU + dropout) -- Backbone/slim_nets.py:34-57; slim.avg_pool2d(guide, 2) -- GUNet.py:155
REGISTER_OP("BslFullyConnected").Input("x: float").Input("w: float").Input("bias: float").Input("is_training: bool")
    .Output("y: float").Attr("relu: int = 1").Attr("keep_prob: float = 1.0").Attr("seed: int = 0").Attr("offset: int = 0");
class BslFullyConnectedOp : public OpKernel {
 public:
  explicit BslFullyConnectedOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("relu", &relu_));
    OP_REQUIRES_OK(c, c->GetAttr("keep_prob", &keep_));
    OP_REQUIRES_OK(c, c->GetAttr("seed", &seed_));
    OP_REQUIRES_OK(c, c->GetAttr("offset", &off_));
  }
  void Compute(OpKernelContext* c) override {
    const Tensor &x = c->input(0), &w = c->input(1);
    const bool training = c->input(3).scalar<bool>()();
    bsl_fc_desc d{Dim(x, 0), Dim(w, 0), Dim(w, 1), relu_, (training && keep_ < 1.f) ? 1 : 0,
                  bsl_dropout_desc{keep_, static_cast<unsigned long long>(seed_), static_cast<unsigned long long>(off_)}};
    Tensor* y = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, TensorShape({d.n, d.cout}), &y));
    BSL_TF_CALL(c, bsl_fc_fwd(BSL_CTX, &d, static_cast<const float*>(In(c, 0)), static_cast<const float*>(In(c, 1)),
                              static_cast<const float*>(In(c, 2)), static_cast<float*>(Out(y)), Stream(c)));
  }
 private:
  int relu_ = 1;
  float keep_ = 1.f;
  int64 seed_ = 0, off_ = 0;
};
REGISTER_KERNEL_BUILDER(Name("BslFullyConnected").Device(DEVICE_GPU).HostMemory("is_training"), BslFullyConnectedOp);

REGISTER_OP("BslAvgPool2x2").Input("x: float").Output("y: float");
class BslAvgPool2x2Op : public OpKernel {
 public:
  explicit BslAvgPool2x2Op(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* c) override {
    const Tensor& x = c->input(0);
    Tensor* y = nullptr;
    OP_REQUIRES_OK(c, c->allocate_output(0, TensorShape({Dim(x, 0), Dim(x, 1) / 2, Dim(x, 2) / 2, Dim(x, 3)}), &y));
    BSL_TF_CALL(c, bsl_avgpool2x2_f32(BSL_CTX, Dim(x, 0), Dim(x, 1), Dim(x, 2), Dim(x, 3), static_cast<const float*>(In(c, 0)),
                                      static_cast<float*>(Out(y)), Stream(c)));
  }
};
REGISTER_KERNEL_BUILDER(Name("BslAvgPool2x2").Device(DEVICE_GPU), BslAvgPool2x2Op);

// modulated_conv_block (GUNet.py:162-217): instance norm -> * gamma_mod[n, c] -> + guide . w_sp + b_sp -> ReLU, one pass.
REGISTER_OP("BslModulatedNormRelu")
    .Input("y: bfloat16").Input("sums: double").Input("gamma: float").Input("beta: float").Input("gamma_mod: float")
    .Input("guide: float").Input("w_guide: float").Input("b_guide: float")
    .Output("a: bfloat16").Output("mean: float").Output("rstd: float").Output("scale_out: float").Output("shift_out: float")
    .Attr("epsilon: float = 1e-6").Attr("center: int = 1").Attr("scale: int = 0");
class BslModulatedNormReluOp : public OpKernel {
 public:
  explicit BslModulatedNormReluOp(OpKernelConstruction* c) : OpKernel(c) {
    a_.mode = 1;
    OP_REQUIRES_OK(c, c->GetAttr("epsilon", &a_.eps));
    OP_REQUIRES_OK(c, c->GetAttr("center", &a_.center));
    OP_REQUIRES_OK(c, c->GetAttr("scale", &a_.scale));
  }
  void Compute(OpKernelContext* c) override {
    const Tensor &y = c->input(0), &gm = c->input(4), &gd = c->input(5), &wg = c->input(6);
    bsl_norm_desc d = a_.Desc(y);
    Tensor *a = nullptr, *stat[4] = {nullptr, nullptr, nullptr, nullptr};
    OP_REQUIRES_OK(c, c->allocate_output(0, y.shape(), &a));
    for (int i = 0; i < 4; ++i) OP_REQUIRES_OK(c, c->allocate_output(1 + i, TensorShape({d.n, d.c}), &stat[i]));
    float* f[4];
    for (int i = 0; i < 4; ++i) f[i] = static_cast<float*>(Out(stat[i]));
    BSL_TF_CALL(c, bsl_norm_finalize(BSL_CTX, &d, 1, static_cast<const double*>(In(c, 1)), static_cast<const float*>(In(c, 2)),
                                     static_cast<const float*>(In(c, 3)), nullptr, nullptr, f[0], f[1], f[2], f[3], Stream(c)));
    BSL_TF_CALL(c, bsl_norm_modulate(BSL_CTX, &d, static_cast<const float*>(In(c, 4)), Dim(gm, 1),
                                     static_cast<const float*>(In(c, 7)), f[2], f[3], Stream(c)));
    bsl_guide g{static_cast<const float*>(In(c, 5)), Dim(gd, 3), static_cast<const float*>(In(c, 6)), Dim(wg, wg.dims() - 1)};
    BSL_TF_CALL(c, bsl_norm_apply_mod(BSL_CTX, &d, In(c, 0), f[2], f[3], &g, Out(a), Stream(c)));
  }
 private:
  NormAttrs a_;
};
REGISTER_KERNEL_BUILDER(Name("BslModulatedNormRelu").Device(DEVICE_GPU), BslModulatedNormReluOp);

}  // namespace bsl_tf

#endif  // TensorFlow headers (or the syntax-check stub) available
