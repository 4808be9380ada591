// GPU side of the reference's input pipeline: what lib/data/dataset.py:6-12 (PIL greyscale load) and
// train.py:69-72 (transforms.Resize + transforms.ToTensor) do to every decoded ground-truth / mask image, for a whole batch:
//
//   uint8 [B,H,W,3] (RGB) or [B,H,W] (L)  --L = (R*19595 + G*38470 + B*7471 + 0x8000) >> 16-->  8-bit greyscale
//   --Pillow's separable triangle-filter resample in 22-bit fixed point, horizontal pass first, 8-bit intermediate-->
//   uint8 [B,outH,outW]  --/ 255.f-->  float32 [B,1,outH,outW]
//
// Byte / integer work, bit-exact with Pillow + torchvision (the per-output-pixel tap tables come from the host, computed with
// Pillow's own double-precision formula: data.py).  HBM-bound: both passes are coalesced along the image row (a warp reads
// neighbouring source bytes / writes 32 consecutive outputs) and run as grid-stride loops over 148 x 8 blocks.
#include "common.cuh"

namespace b200 {

constexpr int kPrecisionBits = 32 - 8 - 2;  // Pillow: PRECISION_BITS

__device__ __forceinline__ int clip8(int acc) {
  const int v = acc >> kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// tmp[b][y][xx] = clip8(2^21 + sum_t L(src[b][y][xmin[xx] + t]) * xk[xx][t])
__global__ void __launch_bounds__(256) preprocess_rows_kernel(const uint8_t* __restrict__ src, long long rows, int W, int channels, int outW,
                                                              const int* __restrict__ xmin, const int* __restrict__ xcnt, const int* __restrict__ xk,
                                                              int ksize, uint8_t* __restrict__ tmp) {
  pdl_enter();
  const long long total = rows * outW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / outW;
    const int xx = (int)(i - row * outW);
    const uint8_t* s = src + (row * W + __ldg(xmin + xx)) * channels;
    const int n = __ldg(xcnt + xx);
    const int* k = xk + (size_t)xx * ksize;
    int acc = 1 << (kPrecisionBits - 1);
    if (channels == 3) {
      for (int t = 0; t < n; ++t) {
        const int l = (s[3 * t] * 19595 + s[3 * t + 1] * 38470 + s[3 * t + 2] * 7471 + 0x8000) >> 16;  // Pillow rgb2l
        acc += l * __ldg(k + t);
      }
    } else {
      for (int t = 0; t < n; ++t) acc += s[t] * __ldg(k + t);
    }
    tmp[i] = (uint8_t)clip8(acc);
  }
}

// out[b][0][yy][x] = clip8(2^21 + sum_t tmp[b][ymin[yy] + t][x] * yk[yy][t]) / 255.f
__global__ void __launch_bounds__(256) preprocess_cols_kernel(const uint8_t* __restrict__ tmp, int B, int H, int outH, int outW, const int* __restrict__ ymin,
                                                              const int* __restrict__ ycnt, const int* __restrict__ yk, int ksize,
                                                              float* __restrict__ out) {
  pdl_enter();
  const long long total = (long long)B * outH * outW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % outW);
    const long long r = i / outW;
    const int yy = (int)(r % outH);
    const long long b = r / outH;
    const uint8_t* s = tmp + (b * H + __ldg(ymin + yy)) * outW + x;
    const int n = __ldg(ycnt + yy);
    const int* k = yk + (size_t)yy * ksize;
    int acc = 1 << (kPrecisionBits - 1);
    for (int t = 0; t < n; ++t) acc += s[(size_t)t * outW] * __ldg(k + t);
    out[i] = __fdiv_rn((float)clip8(acc), 255.f);  // torchvision ToTensor: .to(float32).div(255)
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

// src: uint8 [B,H,W,channels] (channels 3 = RGB, 1 = L); tmp: B*H*outW bytes of scratch; out: float32 [B,1,outH,outW].
// (xmin, xcnt, xk[outW][ksize_x]) / (ymin, ycnt, yk[outH][ksize_y]): first source index, tap count and 22-bit fixed-point weights of
// every output column / row (Pillow's precompute_coeffs + normalize_coeffs_8bpc), device arrays.
int b200_preprocess(const void* src, int B, int H, int W, int channels, int outH, int outW, const int* xmin, const int* xcnt, const int* xk, int ksize_x,
                    const int* ymin, const int* ycnt, const int* yk, int ksize_y, void* tmp, float* out, cudaStream_t stream) {
  if (B < 1 || H < 1 || W < 1 || outH < 1 || outW < 1 || (channels != 1 && channels != 3) || ksize_x < 1 || ksize_y < 1) {
    b200_set_error("preprocess: bad sizes (channels must be 1 or 3)");
    return B200_ERR_ARG;
  }
  const long long rows = (long long)B * H;
  long long g1 = (rows * outW + 255) / 256, g2 = ((long long)B * outH * outW + 255) / 256;
  const long long cap = 148LL * 8;
  if (g1 > cap) g1 = cap;
  if (g2 > cap) g2 = cap;
  B200_LAUNCH((preprocess_rows_kernel), (int)g1, 256, 0, stream, (const uint8_t*)src, rows, W, channels, outW, xmin, xcnt, xk, ksize_x, (uint8_t*)tmp);
  B200_CHECK_LAUNCH();
  B200_LAUNCH((preprocess_cols_kernel), (int)g2, 256, 0, stream, (const uint8_t*)tmp, B, H, outH, outW, ymin, ycnt, yk, ksize_y, out);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

}  // extern "C"
