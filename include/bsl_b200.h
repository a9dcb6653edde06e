/* bsl_b200.h -- C ABI of the B200-native U-Net hot path for BoxSegLiver.
 *
 * The reference (Jarvis73/BoxSegLiver) has no FFI of its own: its model code calls TF-slim layer
 * functions and TF 1.13 dispatches registered op kernels below them. Each
 * entry point here is the enqueue-only body of the TF op kernel it replaces; the citation on every
 * function names the reference call site whose op it stands in for. Rules of the boundary:
 *
 *   - plain pointers and sizes only; device buffers are owned by the caller (TF allocator, or
 *     bsl_malloc for the ctypes host in boxsegliver_b200/), the library owns only its context;
 *   - every compute call only ENQUEUES on `stream` (a cudaStream_t passed as void*), never
 *     synchronises and never touches the default stream;
 *   - return 0 on success, a negative BSL_E* code otherwise; bsl_last_error(ctx) has the text;
 *   - activations / gradients are NHWC bf16 with an explicit channel stride `*_ld` (elements) so a
 *     skip-concat (UNet.py:93) is two views of one buffer, never a copy;
 *   - conv filters are HWIO (TF variable layout, `weights`), transposed-conv filters are
 *     [kh,kw,Cout,Cin] (slim.conv2d_transpose variable layout); fp32 masters, bf16 shadows.
 */
#ifndef BSL_B200_H_
#define BSL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bsl_ctx bsl_ctx;

enum {
  BSL_OK = 0,
  BSL_EINVAL = -1,       /* malformed descriptor / null pointer */
  BSL_EUNSUPPORTED = -2, /* shape outside what the sm_100a kernels cover */
  BSL_EWORKSPACE = -3,   /* workspace too small */
  BSL_ECUDA = -4,        /* CUDA runtime / driver error */
  BSL_EDEVICE = -5,      /* device-side pipeline watchdog fired (see bsl_device_status) */
  BSL_ENCCL = -6
};

/* ------------------------------------------------------------------ context, memory, streams */
int bsl_init(int device, bsl_ctx** out);
void bsl_destroy(bsl_ctx* ctx);
const char* bsl_last_error(bsl_ctx* ctx);
const char* bsl_version(void);
/* Reads back (synchronously) and clears the device-side watchdog word. 0 = healthy. */
int bsl_device_status(bsl_ctx* ctx, int* block, int* site);

/* Probe hook: overrides a layout constant of the UMMA descriptors (tools/gpu_conv_probe.py). */
int bsl_debug_set(bsl_ctx* ctx, int key, int value);

int bsl_malloc(bsl_ctx* ctx, size_t bytes, void** out);
int bsl_free(bsl_ctx* ctx, void* ptr);
int bsl_memset(bsl_ctx* ctx, void* dst, int value, size_t bytes, void* stream);
int bsl_memcpy_h2d(bsl_ctx* ctx, void* dst, const void* src, size_t bytes, void* stream);
int bsl_memcpy_d2h(bsl_ctx* ctx, void* dst, const void* src, size_t bytes, void* stream);
int bsl_memcpy_d2d(bsl_ctx* ctx, void* dst, const void* src, size_t bytes, void* stream);
int bsl_host_alloc(bsl_ctx* ctx, size_t bytes, void** out); /* pinned */
int bsl_host_free(bsl_ctx* ctx, void* ptr);
int bsl_stream_create(bsl_ctx* ctx, void** out);
int bsl_stream_destroy(bsl_ctx* ctx, void* stream);
int bsl_stream_sync(bsl_ctx* ctx, void* stream);
int bsl_event_create(bsl_ctx* ctx, void** out);
int bsl_event_destroy(bsl_ctx* ctx, void* ev);
int bsl_event_record(bsl_ctx* ctx, void* ev, void* stream);
int bsl_event_sync(bsl_ctx* ctx, void* ev);
int bsl_stream_wait_event(bsl_ctx* ctx, void* stream, void* ev);
int bsl_event_elapsed_ms(bsl_ctx* ctx, void* start, void* stop, float* ms);
/* CUDA-graph capture of an enqueue sequence (the training step is launch-bound at small batch). */
int bsl_graph_begin(bsl_ctx* ctx, void* stream);
int bsl_graph_end(bsl_ctx* ctx, void* stream, void** graph_exec);
int bsl_graph_launch(bsl_ctx* ctx, void* graph_exec, void* stream);
int bsl_graph_destroy(bsl_ctx* ctx, void* graph_exec);

/* ------------------------------------------------------------------ dtype casts (fp32 <-> bf16) */
int bsl_cast_f32_to_bf16(bsl_ctx* ctx, const float* src, void* dst, size_t n, void* stream);
int bsl_cast_bf16_to_f32(bsl_ctx* ctx, const void* src, float* dst, size_t n, void* stream);

/* ------------------------------------------------------------------ conv2d, stride 1, SAME
 * Replaces TF ops Conv2D / Conv2DBackpropInput / Conv2DBackpropFilter behind
 * slim.conv2d(x, C, 3) -- NetworksV2/UNet.py:79,85,94 -- and their gradients created by
 * optimizer.minimize, core/solver.py:239. tcgen05 implicit GEMM when cin % 64 == 0 and
 * cout % 64 == 0; a direct CUDA-core kernel for the stem (cin < 64, UNet.py:79 first call). */
typedef struct {
  int n, h, w;     /* batch and spatial size (output == input size: stride 1, SAME) */
  int cin, cout;
  int kh, kw;      /* 3x3 or 1x1 */
  int x_ld;        /* channel stride of the input buffer (elements), >= cin */
  int y_ld;        /* channel stride of the output buffer (elements), >= cout */
} bsl_conv2d_desc;

int bsl_conv2d_fprop(bsl_ctx* ctx, const bsl_conv2d_desc* d, const void* x_bf16,
                     const void* w_hwio_bf16, void* y_bf16, void* stream);
/* dx = dgrad(dy, w); dy has stride y_ld, dx has stride x_ld. */
int bsl_conv2d_dgrad(bsl_ctx* ctx, const bsl_conv2d_desc* d, const void* dy_bf16,
                     const void* w_hwio_bf16, void* dx_bf16, void* stream);
size_t bsl_conv2d_wgrad_workspace(bsl_ctx* ctx, const bsl_conv2d_desc* d);
/* dw (fp32, HWIO) = wgrad(x, dy). Split-K partials go to `workspace` and are reduced in a fixed
 * order (bit-reproducible). */
int bsl_conv2d_wgrad(bsl_ctx* ctx, const bsl_conv2d_desc* d, const void* x_bf16, const void* dy_bf16,
                     float* dw_hwio_f32, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ conv2d_transpose k2 s2
 * Replaces Conv2DBackpropInput-as-forward + BiasAdd + Relu behind
 * slim.conv2d_transpose(x, C/2, 2, 2) -- NetworksV2/UNet.py:91 -- and its gradients.
 * x: [n,h,w,cin]; y: [n,2h,2w,cout] written with stride y_ld (the upper half of a concat buffer). */
typedef struct {
  int n, h, w;  /* INPUT spatial size */
  int cin, cout;
  int x_ld, y_ld;
  int relu;     /* slim default activation_fn=relu */
} bsl_convT2d_desc;

int bsl_convT2d_fwd(bsl_ctx* ctx, const bsl_convT2d_desc* d, const void* x_bf16,
                    const void* w_kkoi_bf16, const float* bias_f32, void* y_bf16, void* stream);
/* dyr must already carry the ReLU mask (dy * (y > 0)); see bsl_relu_bwd. */
int bsl_convT2d_bwd_data(bsl_ctx* ctx, const bsl_convT2d_desc* d, const void* dyr_bf16,
                         const void* w_kkoi_bf16, void* dx_bf16, void* stream);
size_t bsl_convT2d_bwd_filter_workspace(bsl_ctx* ctx, const bsl_convT2d_desc* d);
int bsl_convT2d_bwd_filter(bsl_ctx* ctx, const bsl_convT2d_desc* d, const void* x_bf16,
                           const void* dyr_bf16, float* dw_kkoi_f32, float* dbias_f32, void* workspace,
                           size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BSL_B200_H_ */
