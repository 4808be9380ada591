// Hardware probe (not product code): how many SM clocks does one tcgen05.mma (kind::f16, bf16 in / fp32 accumulate, K = 16) take
// when MMAs are issued back to back from shared-memory operands that are already resident — per N, for a single CTA (M = 128)
// and for a CTA pair (cta_group::2, M = 256)?  Separates the tensor pipe's own per-instruction cost from the issue path / the
// operand feed of igemm_kernel (profiles/experiments/README.md).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I gan-inpainting_b200/csrc -o tools/umma_rate tools/umma_rate.cu gan-inpainting_b200/csrc/api.cu -lcuda
//
// mode 0: the same A / B descriptors for every MMA (pure tensor-pipe rate)
// mode 1: descriptors recomputed per MMA like igemm_kernel (4 K slices x 8 "stages" of 16 KB A + B tiles)
// mode 2: as mode 1, A descriptors into a halo window (row offsets, 9-row group stride)
#include "common.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <vector>

using namespace b200;

template <bool PAIR>
__global__ void __launch_bounds__(128) rate_kernel(int N, int mode, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // 4 stages of (A 20 KB window / tile + B up to 32 KB); contents are irrelevant (zeros)
  uint64_t* bar = (uint64_t*)(smem + 4 * (20 + 32) * 1024);
  uint32_t* tptr = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5;
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  for (int i = threadIdx.x; i < 4 * 52 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0u;
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    if (PAIR) {
      tmem_alloc2(tptr, 256);
      tmem_relinquish2();
    } else {
      tmem_alloc(tptr, 256);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  if (warp == 0 && rank == 0) {
    const uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, N, 0, 0);
    const int b_bytes = (PAIR ? N / 2 : N) * 128;
    const long long t0 = clock64();
    int s = 0;
    for (int it = 0; it < iters; ++it) {
      const uint32_t a_base = smem_u32(smem) + (mode ? s * 52 * 1024 : 0);
      const uint32_t b_base = a_base + 20 * 1024;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t a_addr = a_base + (mode ? k * 32 : 0);
        uint32_t sbo = 1024u;
        if (mode == 2) {
          a_addr += (((it >> 1) & 1) * 9 + (it & 1)) * 128;
          sbo = 1152u;
        }
        const uint64_t da = make_smem_desc_sw128(a_addr, 16u, sbo);
        const uint64_t db = make_smem_desc_sw128(b_base + (mode ? k * 32 : 0), 16u, 1024u);
        if (PAIR) umma_bf16_2cta_w(tmem, da, db, idesc, 1u);
        else umma_bf16_w(tmem, da, db, idesc, 1u);
      }
      if (mode && ++s == 4) s = 0;
      (void)b_bytes;
    }
    if (PAIR) umma_commit_2cta_w(&bar[0]);
    else umma_commit_w(&bar[0]);
    mbar_wait(&bar[0], 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x >> (PAIR ? 1 : 0)] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tmem, 256);
    else tmem_dealloc(tmem, 256);
  }
}

template <bool PAIR>
static void run(int N, int mode, int ctas, long long* d_out) {
  const int iters = 2048;
  const size_t smem = 4 * 52 * 1024 + 1024 + 64;
  cudaFuncSetAttribute(rate_kernel<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = PAIR ? 2 : 1;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) cudaLaunchKernelEx(&cfg, rate_kernel<PAIR>, N, mode, iters, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("UMMA_RATE error %s\n", cudaGetErrorString(e));
    exit(3);
  }
  std::vector<long long> h(ctas);
  cudaMemcpy(h.data(), d_out, sizeof(long long) * (PAIR ? ctas / 2 : ctas), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < (PAIR ? ctas / 2 : ctas); ++i) mx = h[i] > mx ? h[i] : mx;
  const double per = (double)mx / (iters * 4.0);
  const double ideal = (double)N * 128 * 16 / 4096.0;
  printf("UMMA_RATE %s M=%d N=%3d mode=%d ctas=%3d: %.1f clk per MMA (%.0f at 4096 MAC/clk/SM -> %.0f %% of the tensor peak)\n", PAIR ? "pair" : "cta ",
         PAIR ? 256 : 128, N, mode, ctas, per, ideal, 100.0 * ideal / per);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 256 * sizeof(long long));
  const int Ns[] = {32, 64, 128, 256};
  for (int mode = 0; mode < 3; ++mode)
    for (int N : Ns) {
      run<false>(N, mode, 1, d_out);
      run<true>(N, mode, 2, d_out);
    }
  // all SMs busy (power / clock effects, shared-memory contention inside a TPC)
  for (int N : Ns) {
    run<false>(N, 1, 148, d_out);
    run<true>(N, 1, 148, d_out);
  }
  return 0;
}
