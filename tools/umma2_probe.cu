// Hardware probe (not product code): tcgen05.mma.cta_group::2 on B200 -- the building block for the round-2 conv
// kernels of the 64/128-output-channel layers, which are bound by the shared-memory operand fetch of 128xNx16
// UMMAs (profiles/r01_ncu_full_conv_kernels.txt: tensor pipe 48-54 % active, l1tex tc wavefronts 72-82 %).
//   (1) correctness: D[256 x N] = A[256 x K] . B[N x K]^T with a CTA pair: each CTA holds 128 rows of A and N/2 rows
//       of B in its own shared memory; the leader CTA issues the MMAs and commits to both CTAs' barriers.
//   (2) throughput: the same MMA stream over resident operands, cta_group::1 (128 x N x 16 per SM) against
//       cta_group::2 (256 x N x 16 per SM pair), cycles per MMA and TFLOP/s over the whole chip.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I boxsegliver_b200/csrc tools/umma2_probe.cu -o tools/bin/umma2_probe
//   timeout 60 tools/bin/umma2_probe            (on the B200 box)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "ptx.cuh"

using namespace bsl;

constexpr int KB = 4;            // 64-element K blocks resident in shared memory (K = 256)
constexpr uint64_t TIMEOUT_NS = 500000000ull;

// Timing-only variant of the cta_group::1 stream with the operand geometry of the conv kernels: A is a halo tile
// (10 x 18 pixels of 128 B), 8-row groups SBO = 1280 B apart, start address shifted by the tap offset
// (r * 10 + s) * 128 B; B as before. halo = 0 reproduces the aligned stream (SBO 1024, no shift).
struct HaloArgs {
  int halo;
  int reps;
  long long* cycles;
};

template <int N>
__global__ void __launch_bounds__(128) halo_stream_kernel(const HaloArgs p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = base;                      // two sub-tiles of 23552 B
  uint8_t* sB = base + 2 * 23552;          // 9 taps x (N x 128 B)
  for (int i = threadIdx.x; i < (2 * 23552 + 9 * N * 128) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(base)[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
  const uint32_t done = smem_u32(&bar_done);
  if (threadIdx.x == 0) { mbar_init(done, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc<(N < 32 ? 32 : N) * 2>(smem_u32(&tmem_slot)); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);
  if (warp == 1 && elect_one_sync()) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, false, false);
    const uint32_t sbo = p.halo ? 1280 : 1024;
    const long long t0 = clock64();
    for (int rep = 0; rep < p.reps; ++rep) {
      for (int tap = 0; tap < 9; ++tap) {
        const int toff = p.halo ? ((tap / 3) * 10 + tap % 3) * 128 : 0;
        const uint64_t db = make_smem_desc_sw128(smem_u32(sB + tap * N * 128), 16, 1024);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint64_t da = make_smem_desc_sw128(smem_u32(sA + j * 23552) + toff, 16, sbo);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem + j * N, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1u);
        }
      }
    }
    umma_commit(done);
    const uint64_t t1 = globaltimer_ns();
    while (!mbar_try_wait(done, 0)) if (globaltimer_ns() - t1 > TIMEOUT_NS) break;
    p.cycles[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<(N < 32 ? 32 : N) * 2>(tmem); }
}

template <int N>
static void run_halo(int halo, int reps, long long* d_cycles, int grid) {
  auto kern = halo_stream_kernel<N>;
  const int smem = 2 * 23552 + 9 * N * 128 + 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  HaloArgs a{halo, reps, d_cycles};
  kern<<<grid, 128, smem>>>(a);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> cyc(grid);
  cudaMemcpy(cyc.data(), d_cycles, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long cmax = 0;
  for (long long c : cyc) cmax = c > cmax ? c : cmax;
  printf("stream 128x%dx16, %s operand geometry: %.1f cycles/MMA  (%s)\n", N,
         halo ? "halo-tile (SBO 1280, tap-shifted start)" : "aligned (SBO 1024)", cmax / (reps * 72.0), cudaGetErrorString(e));
}

struct Args {
  const __nv_bfloat16* a;   // [256][K]   (cta_group::1: rows [0,128) are used by every CTA)
  const __nv_bfloat16* b;   // [N][K]
  float* out;               // [256][N]   written by cluster 0 only
  int n;
  int reps;                 // MMA stream repetitions for the timing
  long long* cycles;        // [grid] cycles of the MMA stream (leader CTAs)
  int* status;              // 0 ok; otherwise a timeout site
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc2(uint32_t smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}

// rows x 64 bf16 from a row-major [.., ldk] matrix into the canonical SWIZZLE_128B K-major tile at `dst` (1024-aligned):
// row r at r * 128 B, 16-byte chunk c of the row stored at chunk position c ^ (r & 7).
__device__ void fill_tile(uint8_t* dst, const __nv_bfloat16* src, int rows, int ldk) {
  for (int i = threadIdx.x; i < rows * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(src + (size_t)r * ldk + c * 8);
    *reinterpret_cast<uint4*>(dst + r * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
}

__device__ bool wait_bar(uint32_t bar, uint32_t parity, int* status, int site) {
  const uint64_t t0 = globaltimer_ns();
  while (!mbar_try_wait(bar, parity)) {
    if (globaltimer_ns() - t0 > TIMEOUT_NS) {
      atomicCAS(status, 0, site);
      return false;
    }
  }
  return true;
}

template <bool TWO, int N>
__global__ void __launch_bounds__(128) probe_kernel(const Args p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_slot;
  constexpr int BROWS = TWO ? N / 2 : N;          // rows of B this CTA holds
  constexpr int COLS = N < 32 ? 32 : N;           // TMEM columns (power of two >= 32)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = TWO ? cluster_ctarank() : 0;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = base;                              // KB x (128 x 128 B)
  uint8_t* sB = base + KB * 16384;                 // KB x (BROWS x 128 B)
  const int K = KB * 64;
  for (int kb = 0; kb < KB; ++kb) {
    fill_tile(sA + kb * 16384, p.a + (size_t)(rank * 128) * K + kb * 64, 128, K);
    fill_tile(sB + kb * (BROWS * 128), p.b + (size_t)(rank * BROWS) * K + kb * 64, BROWS, K);
  }
  const uint32_t done = smem_u32(&bar_done);
  if (threadIdx.x == 0) {
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    if (TWO) { tmem_alloc2<COLS>(smem_u32(&tmem_slot)); tmem_relinquish2(); }
    else     { tmem_alloc<COLS>(smem_u32(&tmem_slot)); tmem_relinquish(); }
  }
  fence_proxy_async_smem();      // generic-proxy tile writes -> UMMA (async proxy) reads
  tc_fence_before();
  __syncthreads();
  if (TWO) cluster_sync_all();   // the peer's operands and barrier are ready before the leader issues
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);

  bool ok = true;
  if (rank == 0 && warp == 1 && elect_one_sync()) {
    constexpr uint32_t idesc = make_idesc_bf16(TWO ? 256 : 128, N, false, false);
    const long long t0 = clock64();
    for (int rep = 0; rep < p.reps; ++rep) {
      for (int kb = 0; kb < KB; ++kb) {
        const uint64_t da = make_smem_desc_sw128(smem_u32(sA + kb * 16384), 16, 1024);
        const uint64_t db = make_smem_desc_sw128(smem_u32(sB + kb * (BROWS * 128)), 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t acc = (rep == p.reps - 1) ? (uint32_t)((kb | k) != 0) : 1u;   // last repetition = the checked result
          if (TWO) umma2_bf16(tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, acc);
          else     umma_bf16(tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, acc);
        }
      }
    }
    if (TWO) umma2_commit(done, 0x3);
    else     umma_commit(done);
    ok = wait_bar(done, 0, p.status, 1);
    p.cycles[blockIdx.x] = clock64() - t0;
  }
  __syncwarp();
  // every thread of both CTAs: wait for the accumulator, then read it back
  ok = wait_bar(done, 0, p.status, 2 + (int)rank);
  tc_fence_after();
  const bool writer = TWO ? (blockIdx.x < 2) : (blockIdx.x == 0);
  if (ok) {
    const int row = warp * 32 + lane;
    const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    for (int c = 0; c < N; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32(trow + c, v);
      tmem_ld_wait();
      if (writer)
        for (int j = 0; j < 32; ++j) p.out[(size_t)(rank * 128 + row) * N + c + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (TWO) cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if (TWO) tmem_dealloc2<COLS>(tmem);
    else     tmem_dealloc<COLS>(tmem);
  }
}

template <bool TWO, int N>
static int run(const char* name, const Args& a, const std::vector<float>& ref, int grid, float clock_ghz) {
  auto kern = probe_kernel<TWO, N>;
  const int smem = KB * 16384 + KB * (TWO ? N / 2 : N) * 128 + 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaMemset(a.status, 0, sizeof(int));
  cudaMemset(a.out, 0, sizeof(float) * 256 * N);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = TWO ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  cudaError_t err = cudaLaunchKernelEx(&cfg, kern, a);
  cudaEventRecord(e1);
  cudaError_t err2 = cudaDeviceSynchronize();
  if (err != cudaSuccess || err2 != cudaSuccess) {
    printf("%-28s LAUNCH/RUN ERROR: %s / %s\n", name, cudaGetErrorString(err), cudaGetErrorString(err2));
    return 1;
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  int status = 0;
  cudaMemcpy(&status, a.status, sizeof(int), cudaMemcpyDeviceToHost);
  const int rows = TWO ? 256 : 128;
  std::vector<float> out(256 * N);
  cudaMemcpy(out.data(), a.out, sizeof(float) * 256 * N, cudaMemcpyDeviceToHost);
  int bad = 0;
  double maxerr = 0;
  for (int i = 0; i < rows * N; ++i) {
    const double d = fabs((double)out[i] - ref[i]);
    if (d > maxerr) maxerr = d;
    if (d > 1e-2 * (1.0 + fabs(ref[i]))) ++bad;
  }
  std::vector<long long> cyc(grid);
  cudaMemcpy(cyc.data(), a.cycles, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long cmax = 0;
  for (int i = 0; i < grid; i += TWO ? 2 : 1) cmax = cyc[i] > cmax ? cyc[i] : cmax;
  const double mmas = (double)a.reps * KB * 4;
  const double flops = 2.0 * (TWO ? 256 : 128) * N * 16 * mmas * (TWO ? grid / 2 : grid);
  printf("%-28s %s  status=%d  mismatches=%d  max|err|=%.3g  %8.1f cycles/MMA (per %s)  kernel %.3f ms  %.0f TFLOP/s (event time)\n",
         name, (status == 0 && bad == 0) ? "PASS" : "FAIL", status, bad, maxerr, cmax / mmas, TWO ? "SM pair" : "SM", ms,
         flops / (ms * 1e-3) / 1e12);
  (void)clock_ghz;
  return (status == 0 && bad == 0) ? 0 : 1;
}

int main(int argc, char** argv) {
  const int reps = argc > 1 ? atoi(argv[1]) : 2000;
  const int K = KB * 64, NMAX = 256;
  std::vector<__nv_bfloat16> ha(256 * K), hb(NMAX * K);
  std::vector<float> fa(256 * K), fb(NMAX * K);
  srand(7);
  for (size_t i = 0; i < ha.size(); ++i) { fa[i] = (rand() % 17 - 8) / 8.0f; ha[i] = __float2bfloat16(fa[i]); }
  for (size_t i = 0; i < hb.size(); ++i) { fb[i] = (rand() % 13 - 6) / 8.0f; hb[i] = __float2bfloat16(fb[i]); }
  Args a = {};
  cudaMalloc((void**)&a.a, ha.size() * 2);
  cudaMalloc((void**)&a.b, hb.size() * 2);
  cudaMalloc((void**)&a.out, sizeof(float) * 256 * NMAX);
  cudaMalloc((void**)&a.cycles, sizeof(long long) * 256);
  cudaMalloc((void**)&a.status, sizeof(int));
  cudaMemcpy((void*)a.a, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy((void*)a.b, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  a.reps = reps;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, reps %d, K = %d per repetition\n", prop.name, sms, reps, K);
  int fails = 0;
  auto ref_for = [&](int n) {
    std::vector<float> r(256 * n);
    for (int m = 0; m < 256; ++m)
      for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int k = 0; k < K; ++k) s += (double)fa[m * K + k] * fb[j * K + k];
        r[m * n + j] = (float)s;
      }
    return r;
  };
  const int g1 = sms, g2 = sms - sms % 2;
  { a.n = 64;  auto r = ref_for(64);  fails += run<false, 64>("cta_group::1 128x64x16", a, r, g1, 0);  fails += run<true, 64>("cta_group::2 256x64x16", a, r, g2, 0); }
  { a.n = 128; auto r = ref_for(128); fails += run<false, 128>("cta_group::1 128x128x16", a, r, g1, 0); fails += run<true, 128>("cta_group::2 256x128x16", a, r, g2, 0); }
  { a.n = 256; auto r = ref_for(256); fails += run<false, 256>("cta_group::1 128x256x16", a, r, g1, 0); fails += run<true, 256>("cta_group::2 256x256x16", a, r, g2, 0); }
  for (int halo = 0; halo < 2; ++halo) {
    run_halo<64>(halo, reps / 4, a.cycles, g1);
    run_halo<128>(halo, reps / 4, a.cycles, g1);
    run_halo<256>(halo, reps / 4, a.cycles, g1);
  }
  printf("%s\n", fails ? "SOME CASES FAILED" : "ALL PASS");
  return fails ? 1 : 0;
}
