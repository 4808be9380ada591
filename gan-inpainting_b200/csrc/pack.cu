// Weight layout transforms between the reference's checkpoint layout (fp32, Conv2d (Cout,Cin,4,4),
// ConvTranspose2d (Cin,Cout,4,4) — SURVEY.md §2.1) and the K-major bf16 operand layouts of the
// implicit-GEMM kernel, plus the deterministic split reduction of the weight-gradient partials.
#include "common.cuh"

namespace b200 {

// src fp32 [D0][D1][16]  ->  dst bf16 [D0][16*D1],  dst[n][t*D1 + k] = src[n][k][t]
// (Conv2d forward operand; ConvTranspose2d data-gradient operand)
__global__ void __launch_bounds__(256) pack_gather_kernel(const float* __restrict__ src, int D0, int D1, __nv_bfloat16* __restrict__ dst) {
  b200::pdl_enter();
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
  b200::pdl_enter();
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

// Both operand layouts of every 4x4 layer of a network in ONE launch (the fused optimizers rewrite all master weights at once).
// table[l] = {src fp32 [D0][D1][16], gather bf16 [D0][16*D1], phase bf16 [4][D1][4*D0], D0, D1, first tile}; tile = 16 d0 x 32 d1.
struct PackEntry {
  const float* src;
  __nv_bfloat16* gather;
  __nv_bfloat16* phase;
  long long D0, D1, tile0;
};

__global__ void __launch_bounds__(256) pack_all_kernel(const PackEntry* __restrict__ table, int nlayers) {
  b200::pdl_enter();
  constexpr int TS = 33, RS = 16 * TS + 1;  // tile[d0][tap][d1] with padded strides (shared-memory banks)
  __shared__ float tile[16 * RS];
  int l = 0;
  while (l + 1 < nlayers && (long long)blockIdx.x >= table[l + 1].tile0) ++l;
  const PackEntry e = table[l];
  const int D0 = (int)e.D0, D1 = (int)e.D1;
  const int t_in = blockIdx.x - (int)e.tile0;
  const int tiles_d1 = D1 / 32;
  const int d0_0 = (t_in / tiles_d1) * 16, d1_0 = (t_in % tiles_d1) * 32;
  // read: 16 rows of 512 contiguous floats (32 d1 x 16 taps)
  for (int i = threadIdx.x; i < 16 * 128; i += 256) {
    const int r = i >> 7, c4 = i & 127;
    const float4 v = __ldg(reinterpret_cast<const float4*>(e.src + ((size_t)(d0_0 + r) * D1 + d1_0) * 16) + c4);
    const int d1 = c4 >> 2, t0 = (c4 & 3) * 4;
    float* dst = tile + r * RS + t0 * TS + d1;
    dst[0] = v.x, dst[TS] = v.y, dst[2 * TS] = v.z, dst[3 * TS] = v.w;
  }
  __syncthreads();
  // gather layout: dst[d0][t * D1 + d1], 32 consecutive d1 per (d0, t): thread = (d0, t, d1 pair)
  for (int i = threadIdx.x; i < 16 * 16 * 16; i += 256) {
    const int dp = i & 15, t = (i >> 4) & 15, r = i >> 8;
    const float* src = tile + r * RS + t * TS + 2 * dp;
    *reinterpret_cast<uint32_t*>(e.gather + (size_t)(d0_0 + r) * 16 * D1 + (size_t)t * D1 + d1_0 + 2 * dp) = pack_bf16x2(src[0], src[1]);
  }
  // phase layout: dst[ph][d1][ab * D0 + d0], tap(ph, ab) = (3 - p - 2a) * 4 + (3 - q - 2b): thread = (ph*4+ab, d1, d0 pair)
  for (int i = threadIdx.x; i < 16 * 32 * 8; i += 256) {
    const int rp = i & 7, n = (i >> 3) & 31, pt = i >> 8;
    const int ph = pt >> 2, ab = pt & 3;
    const int tap = (3 - (ph >> 1) - 2 * (ab >> 1)) * 4 + (3 - (ph & 1) - 2 * (ab & 1));
    const float* src = tile + (2 * rp) * RS + tap * TS + n;
    *reinterpret_cast<uint32_t*>(e.phase + ((size_t)ph * D1 + d1_0 + n) * (4 * (size_t)D0) + (size_t)ab * D0 + d0_0 + 2 * rp) = pack_bf16x2(src[0], src[RS]);
  }
}

// partial fp32 [splits][Cs][16*Cl] (column = t*Cl + cl)  ->  grad fp32 [Cs][Cl][16]
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int Cs, int Cl, float* __restrict__ grad,
                                                           int accumulate) {
  b200::pdl_enter();
  // block: one cs, 64 cl: thread (tap, cl quad) sums the splits with 128-bit loads (256-byte runs per tap), the 64 x 16 tile is
  // transposed through shared memory and written as 4 KB of contiguous floats
  __shared__ float tile[64][17];
  const int cs = blockIdx.y, cl0 = blockIdx.x * 64;
  const int t = threadIdx.x >> 4, cq = threadIdx.x & 15;
  const size_t split_stride = (size_t)Cs * 16 * Cl;
  const float* src = partial + (size_t)cs * 16 * Cl + (size_t)t * Cl + cl0 + 4 * cq;
  // eight split partials in flight per thread (the layers with the longest pixel reduction have 30-60 splits of a small tile: a
  // serial chain of loads made this kernel latency-bound at 1.3 TB/s); the summation order is fixed, hence deterministic
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  int k = 0;
  for (; k + 8 <= splits; k += 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(src + (size_t)(k + u) * split_stride));
#pragma unroll
    for (int u = 0; u < 8; ++u) s.x += v[u].x, s.y += v[u].y, s.z += v[u].z, s.w += v[u].w;
  }
  for (; k < splits; ++k) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + (size_t)k * split_stride));
    s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
  }
  tile[4 * cq][t] = s.x;
  tile[4 * cq + 1][t] = s.y;
  tile[4 * cq + 2][t] = s.z;
  tile[4 * cq + 3][t] = s.w;
  __syncthreads();
  float* dst = grad + ((size_t)cs * Cl + cl0) * 16;
  for (int e = threadIdx.x; e < 64 * 16; e += 256) {
    const float v = tile[e >> 4][e & 15];
    dst[e] = accumulate ? dst[e] + v : v;
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_pack_gather(const float* w, int D0, int D1, void* out, cudaStream_t stream) {
  B200_LAUNCH((pack_gather_kernel), dim3((D1 + 63) / 64, D0), 256, 0, stream, w, D0, D1, (__nv_bfloat16*)out);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_pack_phase(const float* w, int D0, int D1, void* out, cudaStream_t stream) {
  if (D0 % 16 || D1 % 16) {
    b200_set_error("pack_phase: both channel counts must be multiples of 16");
    return B200_ERR_ARG;
  }
  B200_LAUNCH((pack_phase_kernel), dim3(D0 / 16, D1 / 16), 256, 0, stream, w, D0, D1, (__nv_bfloat16*)out);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

// table: device array of nlayers entries {src, gather, phase, D0, D1, tile0} (int64 each; D0 % 16 == 0, D1 % 32 == 0,
// tile0 = number of 16 x 32 tiles of the layers before); total_tiles = sum over layers of (D0/16)*(D1/32).
int b200_pack_all(const void* table, int nlayers, int total_tiles, cudaStream_t stream) {
  if (nlayers < 1 || total_tiles < 1) {
    b200_set_error("pack_all: empty table");
    return B200_ERR_ARG;
  }
  B200_LAUNCH((pack_all_kernel), total_tiles, 256, 0, stream, (const PackEntry*)table, nlayers);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_wgrad_reduce(const float* partial, int splits, int Cs, int Cl, float* grad, int accumulate, cudaStream_t stream) {
  if (Cl % 64) {
    b200_set_error("wgrad_reduce: Cl must be a multiple of 64");
    return B200_ERR_ARG;
  }
  B200_LAUNCH((wgrad_reduce_kernel), dim3(Cl / 64, Cs), 256, 0, stream, partial, splits, Cs, Cl, grad, accumulate);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

}  // extern "C"
