// Hardware probe (not product code): does tcgen05.mma read a K-major SWIZZLE_128B A operand correctly when its shared-memory
// descriptor starts at an arbitrary 128-byte ROW of a TMA-staged window (start address not 1024-byte aligned) and the 8-row
// groups are a non-1024-multiple apart (stride byte offset 1152 = 9 rows)?  That is what lets one staged input window serve
// every filter tap of a stride-2 convolution (DESIGN.md "window operand staging") instead of one TMA box per tap.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I gan-inpainting_b200/csrc -o tools/umma_probe tools/umma_probe.cu -lcuda
//   ./tools/umma_probe
//
// A[r][c] = small integers (exact in bf16), B = 64 x 64 identity -> D[m][n] must equal A[row_off + (m / 8) * (sbo / 128) + m % 8][n].
#include "common.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <vector>

using namespace b200;

constexpr int kRows = 256;

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int row_off, int sbo,
                                                    int bo_mode, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                       // 256 rows x 128 B
  uint8_t* sB = smem + kRows * 128;         // 64 rows x 128 B
  uint64_t* bar = (uint64_t*)(sB + 64 * 128);
  uint32_t* tptr = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tptr, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  if (warp == 0) {
    mbar_arrive_expect_tx_w(&bar[0], (kRows + 64) * 128);
    tma_load_2d_w(sA, &tmA, &bar[0], 0, 0);
    tma_load_2d_w(sB, &tmB, &bar[0], 0, 0);
    mbar_wait(&bar[0], 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    for (int k = 0; k < 4; ++k) {
      const uint32_t a_addr = smem_u32(sA) + row_off * 128 + k * 32;
      uint64_t da = make_smem_desc_sw128(a_addr, 16u, (uint32_t)sbo);
      if (bo_mode == 1) da |= (uint64_t)((a_addr >> 7) & 7) << 49;
      const uint64_t db = make_smem_desc_sw128(smem_u32(sB) + k * 32, 16u, 1024u);
      umma_bf16_w(tmem, da, db, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit_w(&bar[1]);
  }
  mbar_wait(&bar[1], 0);
  tc_fence_after();
  const int r = warp * 32 + lane;
  for (int cc = 0; cc < 64; cc += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + cc, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[r * 64 + cc + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int map2d(EncodeTiledFn fn, CUtensorMap* m, void* base, int rows, int box_rows) {
  cuuint64_t dims[2] = {64, (cuuint64_t)rows};
  cuuint64_t st[1] = {128};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  return (int)fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

int main() {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) {
    printf("UMMA_PROBE no driver\n");
    return 2;
  }
  EncodeTiledFn enc = (EncodeTiledFn)f;
  std::vector<__nv_bfloat16> hA(kRows * 64), hB(64 * 64);
  for (int r = 0; r < kRows; ++r)
    for (int c = 0; c < 64; ++c) hA[r * 64 + c] = __float2bfloat16((float)((r * 7 + c * 3) % 256));
  for (int n = 0; n < 64; ++n)
    for (int k = 0; k < 64; ++k) hB[n * 64 + k] = __float2bfloat16(n == k ? 1.f : 0.f);
  __nv_bfloat16 *dA, *dB;
  float* dOut;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dOut, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tA, tB;
  if (map2d(enc, &tA, dA, kRows, kRows) || map2d(enc, &tB, dB, 64, 64)) {
    printf("UMMA_PROBE encode failed\n");
    return 2;
  }
  const size_t smem = (kRows + 64) * 128 + 1024 + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::vector<float> out(128 * 64);
  const int offs[] = {0, 1, 3, 8, 9, 10, 19};
  const int sbos[] = {1024, 1152, 1280, 2304};
  int all_ok[2] = {1, 1};
  for (int bo = 0; bo < 2; ++bo)
    for (int sbo : sbos)
      for (int off : offs) {
        if (off + 15 * (sbo / 128) + 8 > kRows) continue;
        cudaMemset(dOut, 0xFF, 128 * 64 * 4);
        probe_kernel<<<1, 128, smem>>>(tA, tB, off, sbo, bo, dOut);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("UMMA_PROBE bo=%d sbo=%d off=%d CUDA error %s\n", bo, sbo, off, cudaGetErrorString(e));
          return 3;
        }
        cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0, first = -1;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 64; ++n) {
            const int r = off + (m / 8) * (sbo / 128) + m % 8;
            const float want = (float)((r * 7 + n * 3) % 256);
            if (out[m * 64 + n] != want) {
              if (first < 0) first = m * 64 + n;
              ++bad;
            }
          }
        if (bad) all_ok[bo] = 0;
        printf("UMMA_PROBE base_offset_mode=%d sbo=%4d row_off=%2d mismatches=%4d", bo, sbo, off, bad);
        if (bad) printf("  (first at m=%d n=%d: got %.0f)", first / 64, first % 64, out[first]);
        printf("\n");
      }
  printf("UMMA_PROBE summary: base_offset=0 %s, base_offset=(addr>>7)&7 %s\n", all_ok[0] ? "ALL OK" : "FAILS", all_ok[1] ? "ALL OK" : "FAILS");
  return 0;
}
