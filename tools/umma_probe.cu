// Hardware probe (not product code): does a SWIZZLE_128B UMMA shared-memory descriptor accept a start
// address that is 128-byte aligned but NOT 1024-byte aligned, and a stride between 8-row groups that is
// not a multiple of 1024 B?  If it does, the 9 taps of a 3x3 convolution can be read as shifted windows
// of ONE halo'd activation tile in shared memory instead of 9 separate TMA loads (6x less L2->SM traffic).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I boxsegliver_b200/csrc tools/umma_probe.cu -o tools/bin/umma_probe -lcuda
//   tools/bin/umma_probe            (on the B200 box; prints one line per case: PASS / FAIL + mismatch count)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "ptx.cuh"

using namespace bsl;

struct ProbeArgs {
  int a_mn;        // 0: A K-major (rows = M, 64 K elements per 128 B row); 1: A MN-major (rows = K, 64 M per row)
  int off_rows;    // descriptor start = tile base + off_rows * 128 B
  int sbo_bytes;   // stride between 8-row groups
  int lbo_bytes;   // MN-major: stride between 64-element M blocks
  int kadv_bytes;  // start-address step per UMMA (K = 16)
  int base_mode;   // 0: base_offset = 0; 1: base_offset = (start >> 7) & 7
  int rows;        // rows of A_src loaded into shared memory (<= 512)
  float* out;      // [128][64]
  int* status;
};

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3ffff) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3fff) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3fff) << 32;
  d |= 1ull << 46;
  d |= static_cast<uint64_t>(base_off & 7) << 49;
  d |= 2ull << 61;
  return d;
}

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmA,
                                                    const __grid_constant__ CUtensorMap tmB, const ProbeArgs p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;                 // up to 512 rows x 128 B = 64 KB
  const uint32_t sB = base + 65536;         // 64 x 128 B
  const uint32_t full = smem_u32(&bars[0]), done = smem_u32(&bars[1]);
  DeviceStatus* st = reinterpret_cast<DeviceStatus*>(p.status);
  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc<64>(smem_u32(&tmem_slot));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&tmem_slot);
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(full, p.rows * 128 + 8192);
    for (int r = 0; r < p.rows; r += 256) tma_load_2d(sA + r * 128, &tmA, full, 0, r);
    tma_load_2d(sB, &tmB, full, 0, 0);
    if (mbar_wait(full, 0, st, 1)) {
      tc_fence_after();
      const uint32_t idesc = make_idesc_bf16(128, 64, p.a_mn != 0, p.a_mn != 0);
#pragma unroll 1
      for (int k = 0; k < 4; ++k) {
        const uint32_t a_addr = sA + p.off_rows * 128 + k * p.kadv_bytes;
        const uint32_t bo = p.base_mode ? ((a_addr >> 7) & 7) : 0;
        const uint64_t da = desc_sw128(a_addr, p.lbo_bytes, p.sbo_bytes, bo);
        const uint64_t db = p.a_mn ? desc_sw128(sB + k * 2048, 8192, 1024, 0) : desc_sw128(sB + k * 32, 16, 1024, 0);
        umma_bf16(tmem, da, db, idesc, k != 0);
      }
      umma_commit(done);
    }
  }
  __syncthreads();
  const bool ok = mbar_wait(done, 0, st, 2);
  tc_fence_after();
  if (ok) {
    const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    for (int c = 0; c < 64; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32(trow + c, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) p.out[(warp * 32 + lane) * 64 + c + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<64>(tmem);
  }
}

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return (uint16_t)(u >> 16);  // values are small integers: exact
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
    printf("no cuTensorMapEncodeTiled\n");
    return 2;
  }
  encode_fn enc = (encode_fn)fn;
  const int R = 512;
  std::vector<float> A(R * 64), B(64 * 64);
  srand(1);
  for (auto& v : A) v = (float)(rand() % 7 - 3);
  for (auto& v : B) v = (float)(rand() % 5 - 2);
  std::vector<uint16_t> Ab(R * 64), Bb(64 * 64);
  for (size_t i = 0; i < A.size(); ++i) Ab[i] = f2bf(A[i]);
  for (size_t i = 0; i < B.size(); ++i) Bb[i] = f2bf(B[i]);
  void *dA, *dB;
  float* dO;
  int* dS;
  cudaMalloc(&dA, Ab.size() * 2);
  cudaMalloc(&dB, Bb.size() * 2);
  cudaMalloc(&dO, 128 * 64 * 4);
  cudaMalloc(&dS, 16);
  cudaMemcpy(dA, Ab.data(), Ab.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bb.data(), Bb.size() * 2, cudaMemcpyHostToDevice);
  alignas(64) CUtensorMap tA, tB;
  {
    cuuint64_t gd[2] = {64, (cuuint64_t)R}, gs[1] = {128};
    cuuint32_t bx[2] = {64, 256}, es[2] = {1, 1};
    CUresult r = enc(&tA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t gd2[2] = {64, 64};
    cuuint32_t bx2[2] = {64, 64};
    CUresult r2 = enc(&tB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, gd2, gs, bx2, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r || r2) {
      printf("encode failed %d %d\n", (int)r, (int)r2);
      return 2;
    }
  }
  const int smem = 65536 + 8192 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  struct Case {
    int a_mn, off, sbo_rows, lbo_rows, kadv_rows, base_mode;
  };
  std::vector<Case> cases;
  // K-major A: rows are GEMM rows; kadv is in BYTES/32 inside the 128 B row (handled below)
  for (int bm = 0; bm < 2; ++bm) {
    cases.push_back({0, 0, 8, 0, 0, bm});
    for (int off : {1, 2, 3, 7, 9, 11}) cases.push_back({0, off, 8, 0, 0, bm});
    cases.push_back({0, 0, 10, 0, 0, bm});   // halo pitch 10 pixels
    cases.push_back({0, 11, 10, 0, 0, bm});  // tap (1,1) of a 10-wide halo tile
    cases.push_back({0, 22, 10, 0, 0, bm});
    cases.push_back({0, 5, 18, 0, 0, bm});
  }
  // MN-major A: rows are K (pixels); M = 2 blocks of 64 (LBO apart); each UMMA covers 16 rows (2 groups)
  for (int bm = 0; bm < 2; ++bm) {
    cases.push_back({1, 0, 8, 64, 16, bm});   // the layout the product kernels use today
    cases.push_back({1, 0, 8, 1, 16, bm});    // second tap = +1 pixel
    cases.push_back({1, 3, 8, 1, 16, bm});
    cases.push_back({1, 5, 8, 16, 18, bm});   // rows of an 18-pixel-pitch halo tile, taps (0,*)->(1,*)
    cases.push_back({1, 19, 8, 17, 18, bm});
    cases.push_back({1, 2, 8, 34, 18, bm});
  }
  int fails = 0;
  for (const Case& c : cases) {
    ProbeArgs p = {};
    p.a_mn = c.a_mn;
    p.off_rows = c.off;
    p.sbo_bytes = c.sbo_rows * 128;
    p.lbo_bytes = c.a_mn ? c.lbo_rows * 128 : 16;
    p.kadv_bytes = c.a_mn ? c.kadv_rows * 128 : 32;
    p.base_mode = c.base_mode;
    p.rows = R;
    p.out = dO;
    p.status = dS;
    cudaMemset(dO, 0xff, 128 * 64 * 4);
    cudaMemset(dS, 0, 16);
    probe_kernel<<<1, 128, smem>>>(tA, tB, p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("CUDA error %s on case a_mn=%d off=%d\n", cudaGetErrorString(e), c.a_mn, c.off);
      return 3;
    }
    std::vector<float> out(128 * 64);
    int stv[4];
    cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(stv, dS, 16, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 64; ++n) {
        double ref = 0;
        if (!c.a_mn) {
          const int row = c.off + (m / 8) * c.sbo_rows + m % 8;
          for (int k = 0; k < 64; ++k) ref += (double)A[row * 64 + k] * B[n * 64 + k];
        } else {
          for (int j = 0; j < 4; ++j)
            for (int kk = 0; kk < 16; ++kk) {
              const int row = c.off + (m / 64) * c.lbo_rows + j * c.kadv_rows + (kk / 8) * c.sbo_rows + kk % 8;
              ref += (double)A[row * 64 + (m % 64)] * B[(j * 16 + kk) * 64 + n];
            }
        }
        if (out[m * 64 + n] != (float)ref) ++bad;
      }
    printf("%s a_mn=%d off=%2d sbo_rows=%2d lbo_rows=%2d kadv_rows=%2d base_offset=%s  mismatches=%d status=%d\n",
           bad ? "FAIL" : "PASS", c.a_mn, c.off, c.sbo_rows, c.lbo_rows, c.kadv_rows,
           c.base_mode ? "(addr>>7)&7" : "0", bad, stv[0]);
    fails += bad != 0;
  }
  printf("%d of %zu cases failed\n", fails, cases.size());
  return 0;
}
