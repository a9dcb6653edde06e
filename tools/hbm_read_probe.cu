// Hardware probe (not product code): what does a pure-READ stream reach on this B200, against the measured copy
// bandwidth (read + write) of MEASURED_PEAKS.json? The normalisation reductions (csrc/reduce.cuh) only read.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 tools/hbm_read_probe.cu -o tools/bin/hbm_read_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int UNROLL>
__global__ void read_kernel(const uint4* __restrict__ src, size_t n, unsigned* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned acc = 0;
  for (; i + (UNROLL - 1) * stride < n; i += UNROLL * stride) {
    uint4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) v[u] = src[i + u * stride];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  for (; i < n; i += stride) { uint4 v = src[i]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
  if (acc == 0x12345678u) out[0] = acc;   // keeps the loads alive
}

// Two streams, 8 bytes per thread and stream, WORK dependent FMAs per loaded bf16 pair (the shape of the
// normalisation-backward sums), block-contiguous chunks like csrc/reduce.cuh.
template <int U, int WORK>
__global__ void __launch_bounds__(256, 4) two_stream_kernel(const uint2* __restrict__ a, const uint2* __restrict__ b,
                                                            size_t n, float* __restrict__ out) {
  const size_t per = (n + gridDim.x - 1) / gridDim.x;
  const size_t p0 = blockIdx.x * per, p1 = p0 + per < n ? p0 + per : n;
  float acc0 = 0.f, acc1 = 0.f;
  size_t i = p0 + threadIdx.x;
  for (; i + (U - 1) * 256 < p1; i += U * 256) {
    uint2 va[U], vb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { va[u] = a[i + u * 256]; vb[u] = b[i + u * 256]; }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float x = __uint_as_float(va[u].x << 16), y = __uint_as_float(va[u].y & 0xffff0000u);
      float d = __uint_as_float(vb[u].x << 16), e = __uint_as_float(vb[u].y & 0xffff0000u);
#pragma unroll
      for (int w = 0; w < WORK; ++w) { x = fmaf(x, 1.0001f, d); y = fmaf(y, 0.9999f, e); }
      acc0 += x;
      acc1 = fmaf(y, d, acc1);
    }
  }
  if (acc0 + acc1 == 1234.5f) out[0] = acc0;
}

__global__ void copy_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

int main() {
  const size_t bytes = (size_t)2 << 30, n = bytes / 16;
  uint4 *a, *b;
  unsigned* out;
  cudaMalloc(&a, bytes);
  cudaMalloc(&b, bytes);
  cudaMalloc(&out, 4);
  cudaMemset(a, 1, bytes);
  cudaMemset(b, 2, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto time = [&](auto launch, const char* name, double moved) {
    float best = 1e9f;
    for (int r = 0; r < 6; ++r) {
      cudaEventRecord(e0);
      launch();
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (r && ms < best) best = ms;
    }
    printf("%-44s %.3f ms  %.0f GB/s\n", name, best, moved / best / 1e6);
  };
  for (int bps : {4, 8, 16}) {
    const int grid = 148 * bps;
    char nm[96];
    snprintf(nm, sizeof nm, "read 2 GiB, 256 thr, %2d blocks/SM, unroll 4", bps);
    time([&] { read_kernel<4><<<grid, 256>>>(a, n, out); }, nm, (double)bytes);
    snprintf(nm, sizeof nm, "read 2 GiB, 256 thr, %2d blocks/SM, unroll 8", bps);
    time([&] { read_kernel<8><<<grid, 256>>>(a, n, out); }, nm, (double)bytes);
  }
  float* fo;
  cudaMalloc(&fo, 4);
  const size_t n8 = bytes / 8;
  time([&] { two_stream_kernel<4, 0><<<148 * 8, 256>>>((const uint2*)a, (const uint2*)b, n8, fo); },
       "2 streams x 2 GiB, 8 B loads, chunks, work 0", 2.0 * bytes);
  time([&] { two_stream_kernel<4, 4><<<148 * 8, 256>>>((const uint2*)a, (const uint2*)b, n8, fo); },
       "2 streams x 2 GiB, 8 B loads, chunks, work 4", 2.0 * bytes);
  time([&] { two_stream_kernel<4, 12><<<148 * 8, 256>>>((const uint2*)a, (const uint2*)b, n8, fo); },
       "2 streams x 2 GiB, 8 B loads, chunks, work 12", 2.0 * bytes);
  time([&] { two_stream_kernel<8, 4><<<148 * 8, 256>>>((const uint2*)a, (const uint2*)b, n8, fo); },
       "2 streams x 2 GiB, 8 B loads, unroll 8, work 4", 2.0 * bytes);
  time([&] { copy_kernel<<<148 * 16, 256>>>(a, b, n); }, "copy 2 GiB -> 2 GiB (read + write bytes)", 2.0 * bytes);
  time([&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); }, "cudaMemcpy D2D (read + write bytes)", 2.0 * bytes);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
