// The memory-bound pieces of the VGG19 feature losses of config 4 (lib/models/loss.py:49-136, networks.py:367-389): everything
// around the 3x3 tensor-core convolutions (b200_conv3_gemm) and the Gram GEMM (b200_gram_gemm) of csrc/conv_gemm.cu.
//
//   pack_conv3        fp32 [Cout][Cin][3][3] (torchvision layout) -> bf16 [Cout][(kh*3+kw)*Cin + ci] (K-major GEMM operand)
//   conv3_c1_fwd      first VGG layer on a 1-channel image: loss.py feeds `img.repeat(1, 3, 1, 1)`, so Conv2d(3, 64, 3, 1, 1) equals a
//                     1-channel convolution with the filters summed over their 3 input channels (the caller passes w1[c][9]);
//                     + bias + ReLU -> NHWC bf16.  Bound by the 128-byte-per-pixel output stream.
//   maxpool2          MaxPool2d(2, 2) on NHWC bf16
//   sqdiff            sum (a - b)^2 of two bf16 tensors (feature MSE) / of two fp32 Gram stacks (style MSE), two-stage, deterministic
#include "common.cuh"

namespace b200 {

__global__ void __launch_bounds__(256) pack_conv3_kernel(const float* __restrict__ w, int Cout, int Cin, __nv_bfloat16* __restrict__ out) {
  pdl_enter();
  const long long n = (long long)Cout * Cin * 9;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    // destination index i = (co * 9 + t) * Cin + ci  (consecutive threads write consecutive ci)
    const int ci = (int)(i % Cin);
    const long long r = i / Cin;
    const int t = (int)(r % 9);
    const long long co = r / 9;
    out[i] = __float2bfloat16(w[(co * Cin + ci) * 9 + t]);
  }
}

// thread = (8-channel group, run of 8 pixels of one row): the 72 filter taps of its channel group live in registers for the whole
// grid-stride loop (the loop stride is a multiple of 8, so a thread never changes group); per run 12 image loads (3 rows x {left
// neighbour, two float4, right neighbour}, broadcast across the 8 threads that share the run), 576 FMAs and eight 16-byte stores —
// the index arithmetic is paid once per 8 pixels, which is what makes the kernel FMA-bound instead of issue-bound.
// C == 64, W % 8 == 0, x 16-byte aligned, B*H*W < 2^31.
__global__ void __launch_bounds__(256) conv3_c1_fwd_kernel(const float* __restrict__ x, int B, int H, int W, const float* __restrict__ w1 /*[64][9]*/,
                                                           const float* __restrict__ bias, __nv_bfloat16* __restrict__ y) {
  pdl_enter();
  const int cg = (threadIdx.x & 7) * 8;
  float wr[9][8], br[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    br[k] = bias ? __ldg(bias + cg + k) : 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[t][k] = __ldg(w1 + (cg + k) * 9 + t);
  }
  const unsigned Wr = (unsigned)W >> 3, runs = (unsigned)B * H * Wr;
  for (unsigned run = (blockIdx.x * 256u + threadIdx.x) >> 3; run < runs; run += (gridDim.x * 256u) >> 3) {
    const unsigned r = run / Wr;  // b * H + hq
    const int wq = (int)(run - r * Wr) * 8;
    const int hq = (int)(r % (unsigned)H);
    const float* xr = x + (size_t)r * W + wq;
    float v[3][10];
#pragma unroll
    for (int dh = 0; dh < 3; ++dh) {
      const int hh = hq + dh - 1;
      const bool rv = hh >= 0 && hh < H;
      const float* xp = xr + (ptrdiff_t)(dh - 1) * W;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 m0 = rv ? __ldg(reinterpret_cast<const float4*>(xp)) : z, m1 = rv ? __ldg(reinterpret_cast<const float4*>(xp) + 1) : z;
      v[dh][0] = (rv && wq > 0) ? __ldg(xp - 1) : 0.f;
      v[dh][1] = m0.x, v[dh][2] = m0.y, v[dh][3] = m0.z, v[dh][4] = m0.w;
      v[dh][5] = m1.x, v[dh][6] = m1.y, v[dh][7] = m1.z, v[dh][8] = m1.w;
      v[dh][9] = (rv && wq + 8 < W) ? __ldg(xp + 8) : 0.f;
    }
    __nv_bfloat16* yp = y + ((size_t)r * W + wq) * 64 + cg;
#pragma unroll
    for (int px = 0; px < 8; px += 2) {
      float a0[8], a1[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) a0[k] = a1[k] = br[k];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float v0 = v[t / 3][px + t % 3], v1 = v[t / 3][px + t % 3 + 1];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          a0[k] = fmaf(v0, wr[t][k], a0[k]);
          a1[k] = fmaf(v1, wr[t][k], a1[k]);
        }
      }
      *reinterpret_cast<uint4*>(yp + px * 64) = make_uint4(pack_bf16x2(fmaxf(a0[0], 0.f), fmaxf(a0[1], 0.f)), pack_bf16x2(fmaxf(a0[2], 0.f), fmaxf(a0[3], 0.f)),
                                                           pack_bf16x2(fmaxf(a0[4], 0.f), fmaxf(a0[5], 0.f)), pack_bf16x2(fmaxf(a0[6], 0.f), fmaxf(a0[7], 0.f)));
      *reinterpret_cast<uint4*>(yp + px * 64 + 64) = make_uint4(pack_bf16x2(fmaxf(a1[0], 0.f), fmaxf(a1[1], 0.f)), pack_bf16x2(fmaxf(a1[2], 0.f), fmaxf(a1[3], 0.f)),
                                                                pack_bf16x2(fmaxf(a1[4], 0.f), fmaxf(a1[5], 0.f)), pack_bf16x2(fmaxf(a1[6], 0.f), fmaxf(a1[7], 0.f)));
    }
  }
}

__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  const float2 fa = unpack_bf16x2(a), fb = unpack_bf16x2(b);
  return pack_bf16x2(fmaxf(fa.x, fb.x), fmaxf(fa.y, fb.y));
}

__global__ void __launch_bounds__(256) maxpool2_kernel(const __nv_bfloat16* __restrict__ x, int B, int H, int W, int C, __nv_bfloat16* __restrict__ y) {
  pdl_enter();
  const int Ho = H >> 1, Wo = W >> 1, chunks = C >> 3;
  const long long total = (long long)B * Ho * Wo * chunks;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int ch = (int)(i % chunks) * 8;
    const long long p = i / chunks;
    const int wo = (int)(p % Wo);
    const long long r = p / Wo;
    const int ho = (int)(r % Ho);
    const long long b = r / Ho;
    const __nv_bfloat16* s = x + ((b * H + 2 * ho) * W + 2 * wo) * C + ch;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(s)), bq = __ldg(reinterpret_cast<const uint4*>(s + C));
    const uint4 c = __ldg(reinterpret_cast<const uint4*>(s + (size_t)W * C)), d = __ldg(reinterpret_cast<const uint4*>(s + (size_t)W * C + C));
    uint4 o;
    o.x = max_bf16x2(max_bf16x2(a.x, bq.x), max_bf16x2(c.x, d.x));
    o.y = max_bf16x2(max_bf16x2(a.y, bq.y), max_bf16x2(c.y, d.y));
    o.z = max_bf16x2(max_bf16x2(a.z, bq.z), max_bf16x2(c.z, d.z));
    o.w = max_bf16x2(max_bf16x2(a.w, bq.w), max_bf16x2(c.w, d.w));
    *reinterpret_cast<uint4*>(y + p * C + ch) = o;
  }
}

constexpr int kSqBlocks = 148 * 4;

__device__ __forceinline__ void block_sum_store(float s, float* dst) {
  __shared__ float rs[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) rs[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += rs[i];
    *dst = t;
  }
}

__device__ __forceinline__ float sq8(const uint4& va, const uint4& vb, float s) {
  const uint32_t ua[4] = {va.x, va.y, va.z, va.w}, ub[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 fa = unpack_bf16x2(ua[k]), fb = unpack_bf16x2(ub[k]);
    const float d0 = fa.x - fb.x, d1 = fa.y - fb.y;
    s = fmaf(d0, d0, s);
    s = fmaf(d1, d1, s);
  }
  return s;
}

// partial[blk] = sum over this block's elements of (a - b)^2 (bf16 inputs; 2 x 8 per thread and step: four 16-byte loads in flight)
__global__ void __launch_bounds__(256) sqdiff_bf16_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, long long n8,
                                                          float* __restrict__ partial) {
  pdl_enter();
  float s0 = 0.f, s1 = 0.f;
  const long long stride = (long long)gridDim.x * 256;
  const uint4* pa = reinterpret_cast<const uint4*>(a);
  const uint4* pb = reinterpret_cast<const uint4*>(b);
  long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  for (; i + stride < n8; i += 2 * stride) {
    const uint4 va = __ldg(pa + i), vb = __ldg(pb + i), vc = __ldg(pa + i + stride), vd = __ldg(pb + i + stride);
    s0 = sq8(va, vb, s0);
    s1 = sq8(vc, vd, s1);
  }
  if (i < n8) s0 = sq8(__ldg(pa + i), __ldg(pb + i), s0);
  block_sum_store(s0 + s1, partial + blockIdx.x);
}

// Gram stacks [B][S][Cp][C] (S partial sums per image, rows >= C unused): partial[blk] = sum ((sum_s ga - sum_s gb) * gscale)^2.
// The C x C block of every matrix is contiguous (row stride C), so a thread takes 4 consecutive floats of it.
__global__ void __launch_bounds__(256) sqdiff_gram_kernel(const float* __restrict__ ga, const float* __restrict__ gb, int B, int S, int Cp, int C,
                                                          float gscale, float* __restrict__ partial) {
  pdl_enter();
  float s = 0.f;
  const unsigned per = (unsigned)(C * C) >> 2, total = (unsigned)B * per;  // float4 groups per image / in all
  const size_t mat = (size_t)Cp * C;
  for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < total; i += gridDim.x * 256u) {
    const unsigned b = i / per, j = i - b * per;
    size_t off = (size_t)b * S * mat + 4 * (size_t)j;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), g = a;
    for (int k = 0; k < S; ++k, off += mat) {
      const float4 u = __ldg(reinterpret_cast<const float4*>(ga + off)), v = __ldg(reinterpret_cast<const float4*>(gb + off));
      a.x += u.x, a.y += u.y, a.z += u.z, a.w += u.w;
      g.x += v.x, g.y += v.y, g.z += v.z, g.w += v.w;
    }
    const float d0 = (a.x - g.x) * gscale, d1 = (a.y - g.y) * gscale, d2 = (a.z - g.z) * gscale, d3 = (a.w - g.w) * gscale;
    s = fmaf(d0, d0, s);
    s = fmaf(d1, d1, s);
    s = fmaf(d2, d2, s);
    s = fmaf(d3, d3, s);
  }
  block_sum_store(s, partial + blockIdx.x);
}

// out[0] (+)= scale * sum_blk partial[blk]  (fp64 accumulation, fixed order)
__global__ void sqdiff_finalize_kernel(const float* __restrict__ partial, int nblk, double scale, float* __restrict__ out, int accumulate) {
  pdl_enter();
  if (threadIdx.x != 0) return;
  double t = 0.0;
  for (int i = 0; i < nblk; ++i) t += (double)partial[i];
  const float v = (float)(t * scale);
  out[0] = accumulate ? out[0] + v : v;
}

static int grid_for(long long items) {
  long long g = (items + 255) / 256;
  if (g > 148LL * 8) g = 148LL * 8;
  return g < 1 ? 1 : (int)g;
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_pack_conv3(const float* w, int Cout, int Cin, void* out, cudaStream_t stream) {
  if (Cout < 1 || Cin < 1) {
    b200_set_error("pack_conv3: bad sizes");
    return B200_ERR_ARG;
  }
  B200_LAUNCH((pack_conv3_kernel), grid_for((long long)Cout * Cin * 9), 256, 0, stream, w, Cout, Cin, (__nv_bfloat16*)out);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_conv3_c1_fwd(const float* x, int B, int H, int W, const float* w1, const float* bias, int C, void* y, cudaStream_t stream) {
  if (C != 64 || B < 1 || H < 1 || W < 8 || (W & 7) || ((uintptr_t)x & 15) || (long long)B * H * W >= (1LL << 31)) {
    b200_set_error("conv3_c1_fwd: C must be 64 (VGG19 features[0]), W a multiple of 8, x 16-byte aligned, B*H*W < 2^31");
    return B200_ERR_ARG;
  }
  B200_LAUNCH((conv3_c1_fwd_kernel), grid_for((long long)B * H * W), 256, 0, stream, x, B, H, W, w1, bias, (__nv_bfloat16*)y);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_maxpool2(const void* x, int B, int H, int W, int C, void* y, cudaStream_t stream) {
  if ((H & 1) || (W & 1) || C % 8 || B < 1) {
    b200_set_error("maxpool2: H, W must be even and C a multiple of 8");
    return B200_ERR_ARG;
  }
  B200_LAUNCH((maxpool2_kernel), grid_for((long long)B * (H / 2) * (W / 2) * (C / 8)), 256, 0, stream, (const __nv_bfloat16*)x, B, H, W, C, (__nv_bfloat16*)y);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_sqdiff_blocks(void) { return kSqBlocks; }

// out[0] (+)= scale * sum (a - b)^2 over n bf16 elements (n % 8 == 0); partial: b200_sqdiff_blocks() floats of scratch
int b200_sqdiff_bf16(const void* a, const void* b, long long n, double scale, float* partial, float* out, int accumulate, cudaStream_t stream) {
  if (n < 8 || (n & 7)) {
    b200_set_error("sqdiff_bf16: n must be a positive multiple of 8");
    return B200_ERR_ARG;
  }
  B200_LAUNCH((sqdiff_bf16_kernel), kSqBlocks, 256, 0, stream, (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, n / 8, partial);
  B200_CHECK_LAUNCH();
  B200_LAUNCH((sqdiff_finalize_kernel), 1, 32, 0, stream, (const float*)partial, kSqBlocks, scale, out, accumulate);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

// out[0] (+)= scale * sum over b, c < C, c' of ((sum_s ga - sum_s gb)[b][c][c'] * gscale)^2 for two Gram stacks [B][S][Cp][C] from
// b200_gram_gemm (S = b200_gram_splits partial sums per image)
int b200_sqdiff_gram(const float* ga, const float* gb, int B, int S, int Cp, int C, float gscale, double scale, float* partial, float* out, int accumulate,
                     cudaStream_t stream) {
  if (B < 1 || S < 1 || C < 2 || (C & 1) || Cp < C || (long long)B * C * C > (1LL << 32)) {
    b200_set_error("sqdiff_gram: bad sizes");
    return B200_ERR_ARG;
  }
  B200_LAUNCH((sqdiff_gram_kernel), kSqBlocks, 256, 0, stream, ga, gb, B, S, Cp, C, gscale, partial);
  B200_CHECK_LAUNCH();
  B200_LAUNCH((sqdiff_finalize_kernel), 1, 32, 0, stream, (const float*)partial, kSqBlocks, scale, out, accumulate);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

}  // extern "C"
