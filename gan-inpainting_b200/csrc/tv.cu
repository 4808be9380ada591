// Total-variation loss of the reference's config 4 (lib/models/loss.py:138-151, called on the inpainted image at
// experiment_list/wgan_perceptual_style_faceparsing.py:219): tv_weight * (mean((x[..., :-1] - x[..., 1:])^2) + mean((x[:, :-1] - x[:, 1:])^2)),
// and its gradient w.r.t. the image (the loss is part of g_loss, :221).  Two HBM-bound stencil kernels over fp32 [planes, H, W]:
// the forward reads each pixel once (right / lower neighbours come from the same rows) and reduces deterministically.
#include "common.cuh"

namespace b200 {

constexpr int kTvThreads = 256;

// one block per (plane, row): partial[2 * row_index + {0, 1}] = sum of squared horizontal / vertical differences of that row
__global__ void __launch_bounds__(kTvThreads) tv_rows_kernel(const float* __restrict__ img, int H, int W, float* __restrict__ partial) {
  b200::pdl_enter();
  __shared__ float red[2][kTvThreads / 32];
  const size_t row = blockIdx.x;
  const int h = (int)(row % H);
  const float* x = img + row * W;
  const bool has_below = h + 1 < H;
  float sw = 0.f, sh = 0.f;
  for (int w = threadIdx.x; w < W; w += kTvThreads) {
    const float v = __ldg(x + w);
    if (w + 1 < W) {
      const float d = v - __ldg(x + w + 1);
      sw = fmaf(d, d, sw);
    }
    if (has_below) {
      const float d = v - __ldg(x + W + w);
      sh = fmaf(d, d, sh);
    }
  }
  sw = warp_sum(sw);
  sh = warp_sum(sh);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = sw;
    red[1][threadIdx.x >> 5] = sh;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    float s = 0.f;
    for (int i = 0; i < kTvThreads / 32; ++i) s += red[threadIdx.x][i];
    partial[2 * row + threadIdx.x] = s;
  }
}

// out[0] = weight * (sum_w / n_w + sum_h / n_h); fixed order, fp64
__global__ void __launch_bounds__(256) tv_finalize_kernel(const float* __restrict__ partial, long long rows, double n_w, double n_h, float weight,
                                                          float* __restrict__ out) {
  b200::pdl_enter();
  __shared__ double red[2][256];
  double a = 0.0, b = 0.0;
  for (long long i = threadIdx.x; i < rows; i += 256) {
    a += (double)partial[2 * i];
    b += (double)partial[2 * i + 1];
  }
  red[0][threadIdx.x] = a;
  red[1][threadIdx.x] = b;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      red[0][threadIdx.x] += red[0][threadIdx.x + s];
      red[1][threadIdx.x] += red[1][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)((double)weight * (red[0][0] / n_w + red[1][0] / n_h));  // W == 1 or H == 1: 0/0 = NaN like torch.mean of an empty tensor
}

// dimg = g * weight * (2/n_w * (d_right - d_left) + 2/n_h * (d_down - d_up)), d_right = x - x[w+1] (0 at the border), ...
__global__ void __launch_bounds__(256) tv_bwd_kernel(const float* __restrict__ img, long long n, int H, int W, float cw, float ch, const float* __restrict__ gscale,
                                                     float* __restrict__ dimg, int accumulate) {
  b200::pdl_enter();
  const float g = gscale ? __ldg(gscale) : 1.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W), h = (int)((i / W) % H);
    const float v = __ldg(img + i);
    float dw = 0.f, dh = 0.f;
    if (w + 1 < W) dw += v - __ldg(img + i + 1);
    if (w > 0) dw -= __ldg(img + i - 1) - v;
    if (h + 1 < H) dh += v - __ldg(img + i + W);
    if (h > 0) dh -= __ldg(img + i - W) - v;
    const float r = g * (cw * dw + ch * dh);
    dimg[i] = accumulate ? dimg[i] + r : r;
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

// img: fp32 [planes, H, W] (planes = N * C); partial: 2 * planes * H floats of scratch; out: 1 float.
int b200_tv_loss(const float* img, long long planes, int H, int W, float weight, float* partial, float* out, cudaStream_t stream) {
  const long long rows = planes * H;
  if (planes < 1 || H < 1 || W < 1 || rows >= (1LL << 31)) {
    b200_set_error("tv_loss: bad sizes");
    return B200_ERR_ARG;
  }
  B200_LAUNCH((tv_rows_kernel), (int)rows, kTvThreads, 0, stream, img, H, W, partial);
  B200_CHECK_LAUNCH();
  B200_LAUNCH((tv_finalize_kernel), 1, 256, 0, stream, (const float*)partial, rows, (double)planes * H * (W - 1), (double)planes * (H - 1) * W, weight, out);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

// dimg (+)= gscale[0] * d tv_loss / d img   (gscale: device scalar = upstream gradient, NULL = 1)
int b200_tv_loss_bwd(const float* img, long long planes, int H, int W, float weight, const float* gscale, float* dimg, int accumulate, cudaStream_t stream) {
  const long long n = planes * H * W;
  if (planes < 1 || H < 1 || W < 1) {
    b200_set_error("tv_loss_bwd: bad sizes");
    return B200_ERR_ARG;
  }
  const double n_w = (double)planes * H * (W - 1), n_h = (double)planes * (H - 1) * W;
  const float cw = n_w > 0 ? (float)(2.0 * weight / n_w) : 0.f, ch = n_h > 0 ? (float)(2.0 * weight / n_h) : 0.f;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  B200_LAUNCH((tv_bwd_kernel), (int)blocks, 256, 0, stream, img, n, H, W, cw, ch, gscale, dimg, accumulate);
  B200_CHECK_LAUNCH();
  return B200_OK;
}
}
