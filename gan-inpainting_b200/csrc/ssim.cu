// SSIM metric of the reference's training log (lib/pytorch_ssim/__init__.py:20-40, called at
// experiment1_global_local_D.py:209 on detached images): five depthwise 11x11 Gaussian convolutions (mu1, mu2, E[x^2], E[y^2],
// E[xy], zero padding 5) and a pointwise formula, averaged.  The reference runs 5 conv2d + ~15 elementwise kernels over the
// images; here ONE kernel reads each image once (plus the 5-pixel halo), runs the separable filter out of shared memory and
// reduces the SSIM map to one partial sum per block; a second tiny kernel folds the partials in a fixed order (deterministic).
// HBM-bound: 8 B read per pixel.
#include "common.cuh"

namespace b200 {

constexpr int kSsimTile = 32;          // output pixels per block side
constexpr int kSsimWin = 11;
constexpr int kSsimHalo = kSsimWin / 2;
constexpr int kSsimIn = kSsimTile + 2 * kSsimHalo;  // 42

struct SsimWindow {
  float g[kSsimWin];
};

__global__ void __launch_bounds__(256) ssim_tile_kernel(const float* __restrict__ img1, const float* __restrict__ img2, int H, int W, int tiles_w, int tiles_h,
                                                        const SsimWindow win, float* __restrict__ partial) {
  b200::pdl_enter();
  __shared__ float sx[kSsimIn][kSsimIn + 1], sy[kSsimIn][kSsimIn + 1];
  __shared__ float hz[5][kSsimIn][kSsimTile + 1];  // horizontally filtered x, y, xx, yy, xy
  __shared__ float red[8];
  const int tile = blockIdx.x;
  const int tw = tile % tiles_w, th = (tile / tiles_w) % tiles_h;
  const size_t plane = tile / (tiles_w * tiles_h);
  const float* p1 = img1 + plane * (size_t)H * W;
  const float* p2 = img2 + plane * (size_t)H * W;
  const int h0 = th * kSsimTile - kSsimHalo, w0 = tw * kSsimTile - kSsimHalo;
  for (int i = threadIdx.x; i < kSsimIn * kSsimIn; i += 256) {
    const int r = i / kSsimIn, c = i - r * kSsimIn;
    const int h = h0 + r, w = w0 + c;
    const bool in = h >= 0 && h < H && w >= 0 && w < W;  // zero padding (F.conv2d padding = 5)
    sx[r][c] = in ? __ldg(p1 + (size_t)h * W + w) : 0.f;
    sy[r][c] = in ? __ldg(p2 + (size_t)h * W + w) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kSsimIn * kSsimTile; i += 256) {
    const int r = i / kSsimTile, c = i - r * kSsimTile;
    float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
    for (int k = 0; k < kSsimWin; ++k) {
      const float x = sx[r][c + k], y = sy[r][c + k], g = win.g[k];
      a = fmaf(g, x, a);
      b = fmaf(g, y, b);
      aa = fmaf(g, x * x, aa);
      bb = fmaf(g, y * y, bb);
      ab = fmaf(g, x * y, ab);
    }
    hz[0][r][c] = a, hz[1][r][c] = b, hz[2][r][c] = aa, hz[3][r][c] = bb, hz[4][r][c] = ab;
  }
  __syncthreads();
  float acc = 0.f;
  for (int i = threadIdx.x; i < kSsimTile * kSsimTile; i += 256) {
    const int r = i / kSsimTile, c = i - r * kSsimTile;
    if (th * kSsimTile + r >= H || tw * kSsimTile + c >= W) continue;
    float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < kSsimWin; ++k) {
      const float g = win.g[k];
#pragma unroll
      for (int q = 0; q < 5; ++q) m[q] = fmaf(g, hz[q][r + k][c], m[q]);
    }
    const float mu1 = m[0], mu2 = m[1];
    const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
    const float s1 = m[2] - mu1_sq, s2 = m[3] - mu2_sq, s12 = m[4] - mu12;
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    acc += ((2.f * mu12 + C1) * (2.f * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2));
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    partial[tile] = s;
  }
}

// out[g] = sum(partial[g*per_group .. (g+1)*per_group)) / count, fixed summation order, fp64 accumulation
__global__ void __launch_bounds__(256) ssim_finalize_kernel(const float* __restrict__ partial, int per_group, double count, float* __restrict__ out) {
  b200::pdl_enter();
  __shared__ double red[256];
  const float* p = partial + (size_t)blockIdx.x * per_group;
  double s = 0.0;
  for (int i = threadIdx.x; i < per_group; i += 256) s += (double)p[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int h = 128; h > 0; h >>= 1) {
    if (threadIdx.x < h) red[threadIdx.x] += red[threadIdx.x + h];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = (float)(red[0] / count);
}

}  // namespace b200

using namespace b200;

extern "C" {

long long b200_ssim_partials(int N, int C, int H, int W) {
  return (long long)N * C * ((H + kSsimTile - 1) / kSsimTile) * ((W + kSsimTile - 1) / kSsimTile);
}

// img1, img2: fp32 [N, C, H, W]; window11: the 11 taps of the normalised 1-D Gaussian (HOST pointer; pytorch_ssim gaussian(11, 1.5));
// partial: b200_ssim_partials(N, C, H, W) floats of scratch; out: 1 float (size_average) or N floats (per-image mean over C, H, W).
int b200_ssim(const float* img1, const float* img2, int N, int C, int H, int W, const float* window11, int size_average, float* partial, float* out,
              cudaStream_t stream) {
  if (N < 1 || C < 1 || H < 1 || W < 1 || !window11) {
    b200_set_error("ssim: bad sizes");
    return B200_ERR_ARG;
  }
  SsimWindow win;
  for (int i = 0; i < kSsimWin; ++i) win.g[i] = window11[i];
  const int tiles_w = (W + kSsimTile - 1) / kSsimTile, tiles_h = (H + kSsimTile - 1) / kSsimTile;
  const long long tiles = (long long)N * C * tiles_w * tiles_h;
  if (tiles >= (1LL << 31)) {
    b200_set_error("ssim: too many tiles");
    return B200_ERR_ARG;
  }
  B200_LAUNCH((ssim_tile_kernel), (int)tiles, 256, 0, stream, img1, img2, H, W, tiles_w, tiles_h, win, partial);
  B200_CHECK_LAUNCH();
  const int groups = size_average ? 1 : N;
  const long long per_group = tiles / groups;
  B200_LAUNCH((ssim_finalize_kernel), groups, 256, 0, stream, (const float*)partial, (int)per_group, (double)C * H * W * (size_average ? N : 1), out);
  B200_CHECK_LAUNCH();
  return B200_OK;
}
}
