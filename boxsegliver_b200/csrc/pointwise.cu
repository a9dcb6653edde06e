// HBM-bound passes of the U-Net hot path: casts, reductions over pixels, normalisation, pooling,
// ReLU masks. All are vectorised (16 B per thread per access), coalesced along the NHWC channel
// axis, and reduce with warp shuffles + a fixed-order second level (no float atomics), so results
// are bit-reproducible run to run.
#include <cuda_bf16.h>
#include "internal.h"

namespace {

constexpr int kThreads = 256;

inline unsigned grid_for(long long work, int per_block, int cap) {
  long long b = (work + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (unsigned)b;
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                     size_t n) {
  size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
  for (; i + 3 < n; i += stride) {
    float4 v = *reinterpret_cast<const float4*>(src + i);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    *reinterpret_cast<uint2*>(dst + i) = o;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (size_t j = n & ~(size_t)3; j < n; ++j) dst[j] = __float2bfloat16_rn(src[j]);
}

__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst,
                                     size_t n) {
  size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
  for (; i + 3 < n; i += stride) {
    uint2 v = *reinterpret_cast<const uint2*>(src + i);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&v.x), b = *reinterpret_cast<__nv_bfloat162*>(&v.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    *reinterpret_cast<float4*>(dst + i) = make_float4(fa.x, fa.y, fb.x, fb.y);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (size_t j = n & ~(size_t)3; j < n; ++j) dst[j] = __bfloat162float(src[j]);
}

// Partial per-channel sums: block b covers pixels [b*ppb, (b+1)*ppb); thread layout is
// (pixel lane, 8-channel group) so every access is a 16 B load of 8 consecutive channels.
__global__ void channel_sum_partial_kernel(const __nv_bfloat16* __restrict__ x, long long pixels, int c,
                                           int ld, long long ppb, float* __restrict__ part) {
  extern __shared__ float sm[];  // [rows][c]
  const int groups = c / 8;
  const int rows = blockDim.x / groups;
  const int g = threadIdx.x % groups;
  const int r = threadIdx.x / groups;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (r < rows) {
    const long long p0 = blockIdx.x * ppb;
    const long long p1 = min(pixels, p0 + ppb);
    for (long long p = p0 + r; p < p1; p += rows) {
      uint4 v = *reinterpret_cast<const uint4*>(x + p * ld + g * 8);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __bfloat1622float2(h[j]);
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[r * c + g * 8 + j] = acc[j];
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float s = 0.f;
    for (int rr = 0; rr < rows; ++rr) s += sm[rr * c + ch];
    part[(long long)blockIdx.x * c + ch] = s;
  }
}

__global__ void channel_sum_final_kernel(const float* __restrict__ part, int blocks, int c,
                                         float* __restrict__ out) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  float s = 0.f;
  for (int b = 0; b < blocks; ++b) s += part[(long long)b * c + ch];
  out[ch] = s;
}

}  // namespace

// Scratch for second-level reductions lives in the context (grown on demand, never shrunk).
static int scratch(bsl_ctx* ctx, size_t bytes, float** out);

int bsl_channel_sum_bf16(bsl_ctx* ctx, const void* x, long long pixels, int c, int ld, float* out,
                         cudaStream_t stream) {
  if (c % 8 || c > 2048) return bsl_fail(ctx, BSL_EUNSUPPORTED, "channel_sum: c=%d", c);
  const int groups = c / 8;
  int threads = kThreads;
  if (threads < groups) threads = groups;
  const int rows = threads / groups;
  const int blocks = (int)grid_for(pixels, 2048, 4 * ctx->sm_count);
  const long long ppb = (pixels + blocks - 1) / blocks;
  float* part = nullptr;
  int rc = scratch(ctx, (size_t)blocks * c * sizeof(float), &part);
  if (rc) return rc;
  channel_sum_partial_kernel<<<blocks, threads, (size_t)rows * c * sizeof(float), stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), pixels, c, ld, ppb, part);
  BSL_LAUNCH_CHECK(ctx, "channel_sum_partial_kernel");
  channel_sum_final_kernel<<<(c + 127) / 128, 128, 0, stream>>>(part, blocks, c, out);
  BSL_LAUNCH_CHECK(ctx, "channel_sum_final_kernel");
  return BSL_OK;
}

static float* g_scratch = nullptr;
static size_t g_scratch_bytes = 0;
static int scratch(bsl_ctx* ctx, size_t bytes, float** out) {
  if (bytes > g_scratch_bytes) {
    // Grown only outside graph capture in practice: first (eager) step sizes it for the model.
    if (g_scratch) cudaFree(g_scratch);
    size_t want = bytes < (8u << 20) ? (8u << 20) : bytes;
    BSL_CUDA(ctx, cudaMalloc(&g_scratch, want));
    g_scratch_bytes = want;
  }
  *out = g_scratch;
  return BSL_OK;
}

extern "C" {

int bsl_cast_f32_to_bf16(bsl_ctx* ctx, const float* src, void* dst, size_t n, void* stream) {
  if (!ctx || !src || !dst) return BSL_EINVAL;
  cast_f32_bf16_kernel<<<grid_for((long long)n, kThreads * 4, 8 * ctx->sm_count), kThreads, 0,
                         as_stream(stream)>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), n);
  BSL_LAUNCH_CHECK(ctx, "cast_f32_bf16_kernel");
  return BSL_OK;
}

int bsl_cast_bf16_to_f32(bsl_ctx* ctx, const void* src, float* dst, size_t n, void* stream) {
  if (!ctx || !src || !dst) return BSL_EINVAL;
  cast_bf16_f32_kernel<<<grid_for((long long)n, kThreads * 4, 8 * ctx->sm_count), kThreads, 0,
                         as_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(src), dst, n);
  BSL_LAUNCH_CHECK(ctx, "cast_bf16_f32_kernel");
  return BSL_OK;
}

}  // extern "C"
