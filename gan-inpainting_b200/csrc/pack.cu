// Weight layout transforms between the reference's checkpoint layout (fp32, Conv2d (Cout,Cin,4,4),
// ConvTranspose2d (Cin,Cout,4,4) — SURVEY.md §2.1) and the K-major bf16 operand layouts of the
// implicit-GEMM kernel, plus the deterministic split reduction of the weight-gradient partials.
#include "common.cuh"

namespace b200 {

// src fp32 [D0][D1][16]  ->  dst bf16 [D0][16*D1],  dst[n][t*D1 + k] = src[n][k][t]
// (Conv2d forward operand; ConvTranspose2d data-gradient operand)
__global__ void __launch_bounds__(256) pack_gather_kernel(const float* __restrict__ src, int D0, int D1, __nv_bfloat16* __restrict__ dst) {
  // one block per (n, 64-wide k slab): stage 64 x 16 floats through smem so both sides are coalesced
  __shared__ float tile[64][17];
  const int n = blockIdx.y, k0 = blockIdx.x * 64;
  const float* s = src + ((size_t)n * D1 + k0) * 16;
  for (int e = threadIdx.x; e < 64 * 16; e += 256) {
    const int k = e >> 4, t = e & 15;
    tile[k][t] = (k0 + k < D1) ? s[e] : 0.f;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 64 * 16; e += 256) {
    const int t = e >> 6, k = e & 63;
    if (k0 + k < D1) dst[(size_t)n * 16 * D1 + (size_t)t * D1 + k0 + k] = __float2bfloat16(tile[k][t]);
  }
}

// src fp32 [D0][D1][4][4] -> dst bf16 [4 phases][D1][4*D0],
//   dst[p*2+q][n][(a*2+b)*D0 + k] = src[k][n][3-p-2a][3-q-2b]
// (ConvTranspose2d forward operand; Conv2d data-gradient operand)
__global__ void __launch_bounds__(256) pack_phase_kernel(const float* __restrict__ src, int D0, int D1, __nv_bfloat16* __restrict__ dst) {
  // block: 16 k x 16 n; src reads contiguous over (n, tap), dst writes contiguous over k
  __shared__ float tile[16][16 * 16 + 1];
  const int k0 = blockIdx.x * 16, n0 = blockIdx.y * 16;
  for (int e = threadIdx.x; e < 16 * 256; e += 256) {
    const int k = e >> 8, r = e & 255;  // r = n_local*16 + tap
    tile[k][r] = src[((size_t)(k0 + k) * D1 + n0) * 16 + r];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 16 * 256; e += 256) {
    const int k = e & 15, nl = (e >> 4) & 15, ph_tap = e >> 8;  // ph_tap = phase*4 + (a*2+b)
    const int ph = ph_tap >> 2, ab = ph_tap & 3;
    const int p = ph >> 1, q = ph & 1, a = ab >> 1, b = ab & 1;
    const int tap = (3 - p - 2 * a) * 4 + (3 - q - 2 * b);
    dst[((size_t)ph * D1 + n0 + nl) * (4 * (size_t)D0) + (size_t)ab * D0 + k0 + k] = __float2bfloat16(tile[k][nl * 16 + tap]);
  }
}

// partial fp32 [splits][Cs][16*Cl] (column = t*Cl + cl)  ->  grad fp32 [Cs][Cl][16]
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int Cs, int Cl, float* __restrict__ grad,
                                                           int accumulate) {
  // block: one cs, 16 cl:  reads 16 taps x 16 cl (64-byte runs), writes 256 contiguous floats
  __shared__ float tile[16][17];
  const int cs = blockIdx.y, cl0 = blockIdx.x * 16;
  const int t = threadIdx.x >> 4, c = threadIdx.x & 15;
  const size_t split_stride = (size_t)Cs * 16 * Cl;
  const float* src = partial + (size_t)cs * 16 * Cl + (size_t)t * Cl + cl0 + c;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += src[k * split_stride];
  tile[c][t] = s;
  __syncthreads();
  const int oc = threadIdx.x >> 4, ot = threadIdx.x & 15;
  float* dst = grad + ((size_t)cs * Cl + cl0 + oc) * 16 + ot;
  *dst = accumulate ? *dst + tile[oc][ot] : tile[oc][ot];
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_pack_gather(const float* w, int D0, int D1, void* out, cudaStream_t stream) {
  pack_gather_kernel<<<dim3((D1 + 63) / 64, D0), 256, 0, stream>>>(w, D0, D1, (__nv_bfloat16*)out);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_pack_phase(const float* w, int D0, int D1, void* out, cudaStream_t stream) {
  if (D0 % 16 || D1 % 16) {
    b200_set_error("pack_phase: both channel counts must be multiples of 16");
    return B200_ERR_ARG;
  }
  pack_phase_kernel<<<dim3(D0 / 16, D1 / 16), 256, 0, stream>>>(w, D0, D1, (__nv_bfloat16*)out);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_wgrad_reduce(const float* partial, int splits, int Cs, int Cl, float* grad, int accumulate, cudaStream_t stream) {
  if (Cl % 16) {
    b200_set_error("wgrad_reduce: Cl must be a multiple of 16");
    return B200_ERR_ARG;
  }
  wgrad_reduce_kernel<<<dim3(Cl / 16, Cs), 256, 0, stream>>>(partial, splits, Cs, Cl, grad, accumulate);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

}  // extern "C"
