// Direct (CUDA-core) kernels for the layers with a single channel on one side, which are
// bandwidth/issue bound rather than tensor-core work:
//
//   conv_c1_fwd    x[B,H,W] fp32 (1 channel)      -> y[B,H/2,W/2,C] bf16   Conv2d(1,C,4,2,1)
//                  (generator d1, discriminator c1: networks.py:285-286, 337; also the data
//                  gradient of the generator's last ConvTranspose2d(128,1), networks.py:293-295)
//   convt_c1_fwd   z[B,H,W,C] bf16                -> y[B,2H,2W] fp32       ConvTranspose2d(C,1,4,2,1)
//                  (+bias, tanh and the mask composite of wgan_l1.py:122; also the data gradient of c1)
//   wgrad_c1       z[B,H,W,C] bf16, g[B,2H,2W]    -> dW[C][16] fp32        weight gradient of both
//
// and the discriminator head: Conv2d(512,1,4,1,0) -> Flatten -> Linear(n*n,1) [-> Sigmoid]
// (networks.py:352-357), which is linear end to end and is therefore evaluated as ONE dot
// product per image with the effective weight E = W5 (*) Wl (full correlation of the two).
#include "common.cuh"

namespace b200 {

enum { ACT_NONE = 0, ACT_LEAKY = 1, ACT_RELU = 2, ACT_TANH = 3 };

// ------------------------------------------------------------------------------------------------
// conv_c1_fwd: block = 256 output pixels (TOH x TOW); the (2*TOH+2) x (2*TOW+2) input window is staged in shared
// memory (zero halo, mask multiply applied once), thread = 4 output channels with the 64 weights in registers.
// ------------------------------------------------------------------------------------------------
constexpr int kXsFloats = 34 * 68;  // covers 18 x 68 (8 x 32 tile) and 34 x 36 (16 x 16 tile)

__global__ void __launch_bounds__(256) conv_c1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ xmul, int H, int W, int tow_shift,
                                                          const float* __restrict__ w, int C, int act, __nv_bfloat16* __restrict__ y,
                                                          __nv_bfloat16* __restrict__ y2, int ld2, int off2) {
  __shared__ __align__(16) float xs[kXsFloats];
  const int TOW = 1 << tow_shift, TOH = 256 >> tow_shift;
  const int Ho = H >> 1, Wo = W >> 1;
  const int tiles_w = Wo >> tow_shift, tiles_h = Ho / TOH;
  const int tw_i = blockIdx.x % tiles_w;
  const int th_i = (blockIdx.x / tiles_w) % tiles_h;
  const int b = blockIdx.x / (tiles_w * tiles_h);
  const int oh0 = th_i * TOH, ow0 = tw_i << tow_shift;
  const int rows = 2 * TOH + 2, cols = 2 * TOW + 2, xstride = 2 * TOW + 4;
  const int ih0 = 2 * oh0 - 1, iw0 = 2 * ow0 - 1;
  const float* xb = x + (size_t)b * H * W;
  const float* mb = xmul ? xmul + (size_t)b * H * W : nullptr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < rows; r += 8) {
    const int ih = ih0 + r;
    const bool rok = ih >= 0 && ih < H;
    for (int cI = lane; cI < cols; cI += 32) {
      const int iw = iw0 + cI;
      float v = 0.f;
      if (rok && iw >= 0 && iw < W) {
        v = __ldg(xb + (size_t)ih * W + iw);
        if (mb) v *= __ldg(mb + (size_t)ih * W + iw);
      }
      xs[r * xstride + cI] = v;
    }
  }
  const int lanes = C >> 2;     // threads per output pixel
  const int ppp = 256 / lanes;  // pixels per pass
  const int cg = threadIdx.x % lanes;
  const int pl = threadIdx.x / lanes;
  float wr[16][4];
#pragma unroll
  for (int t = 0; t < 16; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i) wr[t][i] = __ldg(w + (size_t)(cg * 4 + i) * 16 + t);
  __syncthreads();

  for (int pp = pl; pp < 256; pp += ppp) {
    const int ol_h = pp >> tow_shift, ol_w = pp & (TOW - 1);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kh = 0; kh < 4; ++kh) {
      const float* row = xs + (2 * ol_h + kh) * xstride + 2 * ol_w;
      const float2 a = *reinterpret_cast<const float2*>(row);
      const float2 c2 = *reinterpret_cast<const float2*>(row + 2);
      const float v[4] = {a.x, a.y, c2.x, c2.y};
#pragma unroll
      for (int kw = 0; kw < 4; ++kw)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = fmaf(v[kw], wr[kh * 4 + kw][i], acc[i]);
    }
    if (act == ACT_LEAKY) {
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = acc[i] > 0.f ? acc[i] : 0.2f * acc[i];
    }
    const size_t pix = ((size_t)b * Ho + (oh0 + ol_h)) * Wo + (ow0 + ol_w);
    uint2 o = make_uint2(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]));
    *reinterpret_cast<uint2*>(y + pix * C + cg * 4) = o;
    if (y2) {  // relu copy into the decoder's concat buffer (skip connection)
      o = make_uint2(pack_bf16x2(fmaxf(acc[0], 0.f), fmaxf(acc[1], 0.f)), pack_bf16x2(fmaxf(acc[2], 0.f), fmaxf(acc[3], 0.f)));
      *reinterpret_cast<uint2*>(y2 + pix * ld2 + off2 + cg * 4) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// convt_c1_fwd: block = 16x16 input pixels (+1 halo).  Phase 1: per input pixel the 16 tap dot
// products over C (lane = 8 channels x 8 taps, transposing shuffle reduction) -> smem T[pix][16].
// Phase 2: every output pixel sums its 4 contributing (pixel, tap) entries.
//   y[2i+p, 2j+q] = sum_{a,b} T[i+p+a-1, j+q+b-1][(3-p-2a)*4 + (3-q-2b)]
// ------------------------------------------------------------------------------------------------
constexpr int kTT = 16;             // tile side (input pixels)
constexpr int kTH = kTT + 2;        // with halo

template <int LPP>  // lanes per pixel per tap-half = C/8 (8 or 16)
__global__ void __launch_bounds__(256) convt_c1_fwd_kernel(const __nv_bfloat16* __restrict__ z, int B, int H, int W, int ldz, const float* __restrict__ w,
                                                           const float* __restrict__ bias, int act, const float* __restrict__ base,
                                                           const float* __restrict__ mul, float* __restrict__ y, float* __restrict__ comp) {
  __shared__ float T[kTH * kTH][17];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int LP = 2 * LPP;       // lanes per pixel
  constexpr int PPW = 32 / LP;      // pixels per warp pass
  const int chunk = lane & (LPP - 1);
  const int half = (lane / LPP) & 1;
  const int pslot = lane / LP;
  const int tiles_w = W / kTT, tiles_h = H / kTT;
  const int tw_i = blockIdx.x % tiles_w;
  const int th_i = (blockIdx.x / tiles_w) % tiles_h;
  const int b = blockIdx.x / (tiles_w * tiles_h);
  const int i0 = th_i * kTT - 1, j0 = tw_i * kTT - 1;  // image coordinates of the halo origin

  float wr[8][8];  // [tap in half][channel]
#pragma unroll
  for (int t = 0; t < 8; ++t)
#pragma unroll
    for (int c = 0; c < 8; ++c) wr[t][c] = __ldg(w + (size_t)(chunk * 8 + c) * 16 + half * 8 + t);

  const int npix = kTH * kTH;
  constexpr int U = 4;  // pixels in flight per lane (hides the HBM latency of the z loads)
  for (int base0 = warp * PPW; base0 < npix; base0 += U * 8 * PPW) {
    uint4 q[U];
    bool ok[U];
    int pps[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pp = base0 + u * 8 * PPW + pslot;
      const int li = pp / kTH, lj = pp - li * kTH;
      const int i = i0 + li, j = j0 + lj;
      pps[u] = pp;
      ok[u] = (pp < npix) && i >= 0 && i < H && j >= 0 && j < W;
      q[u] = make_uint4(0u, 0u, 0u, 0u);
      if (ok[u]) q[u] = __ldg(reinterpret_cast<const uint4*>(z + (((size_t)b * H + i) * W + j) * ldz + chunk * 8));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (base0 + u * 8 * PPW >= npix) break;  // warp-uniform
      const int pp = pps[u];
      float v[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) v[t] = 0.f;
      if (ok[u]) {
        const uint32_t uu[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
        float f[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 t2 = unpack_bf16x2(uu[k]);
          f[2 * k] = t2.x;
          f[2 * k + 1] = t2.y;
        }
#pragma unroll
        for (int t = 0; t < 8; ++t)
#pragma unroll
          for (int c = 0; c < 8; ++c) v[t] = fmaf(f[c], wr[t][c], v[t]);
      }
      // transposing reduction over the 8 chunk lanes (bits 0..2): lane keeps tap (lane & 7)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool up = lane & 4;
        const float send = up ? v[k] : v[k + 4];
        const float keep = up ? v[k + 4] : v[k];
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const bool up = lane & 2;
        const float send = up ? v[k] : v[k + 2];
        const float keep = up ? v[k + 2] : v[k];
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      }
      {
        const bool up = lane & 1;
        const float send = up ? v[0] : v[1];
        const float keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
      }
      if (LPP == 16) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 8);
      if (pp < npix && (chunk < 8)) T[pp][half * 8 + (lane & 7)] = v[0];
    }
  }
  __syncthreads();

  const float bv = bias ? __ldg(bias) : 0.f;
  const int Ho = 2 * H, Wo = 2 * W;
  for (int o = threadIdx.x; o < 4 * kTT * kTT; o += 256) {
    const int lo_h = o / (2 * kTT), lo_w = o - lo_h * (2 * kTT);
    const int p = lo_h & 1, q = lo_w & 1;
    const int li = (lo_h >> 1) + 1, lj = (lo_w >> 1) + 1;  // halo-local input coordinates of (i, j)
    float s = bv;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int bb = 0; bb < 2; ++bb) s += T[(li + p + a - 1) * kTH + (lj + q + bb - 1)][(3 - p - 2 * a) * 4 + (3 - q - 2 * bb)];
    if (act == ACT_TANH) s = tanhf(s);
    const int oh = th_i * 2 * kTT + lo_h, ow = tw_i * 2 * kTT + lo_w;
    const size_t idx = ((size_t)b * Ho + oh) * Wo + ow;
    if (comp) {
      // y = G(masked); comp = masked + y*mask   (wgan_l1.py:122)
      y[idx] = s;
      comp[idx] = __ldg(base + idx) + s * __ldg(mul + idx);
    } else {
      if (mul) s *= __ldg(mul + idx);
      y[idx] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad_c1: dW[c][kh*4+kw] = sum_{b,i,j} z[b,i,j,c] * g[b,2i+kh-1,2j+kw-1]
// Persistent blocks walk 8 x 32 pixel tiles: the 18 x 66 window of g is staged in shared memory (zero halo, mask
// multiply applied once); lane = 8 channels x 8 taps accumulators, the next pixel's z is prefetched.
// Deterministic per-block partials.
// ------------------------------------------------------------------------------------------------
constexpr int kWgTH = 8, kWgTW = 32, kWgStride = 2 * kWgTW + 4;

template <int LPP>
__global__ void __launch_bounds__(256) wgrad_c1_kernel(const __nv_bfloat16* __restrict__ z, int B, int H, int W, int ldz, const float* __restrict__ g,
                                                       const float* __restrict__ gmul, float* __restrict__ partial) {
  constexpr int LP = 2 * LPP;       // lanes per pixel
  constexpr int PPP = 256 / LP;     // pixels per pass (whole block)
  constexpr int C = 8 * LPP;
  __shared__ __align__(16) float gs[(2 * kWgTH + 2) * kWgStride];
  __shared__ float sum[C * 16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunk = threadIdx.x & (LPP - 1);
  const int half = (threadIdx.x / LPP) & 1;
  const int slot = threadIdx.x / LP;  // pixel slot inside the block
  const int Hg = 2 * H, Wg = 2 * W;
  const int tiles_w = W / kWgTW, tiles_h = H / kWgTH;
  const int ntiles = B * tiles_w * tiles_h;
  float acc[8][8];
#pragma unroll
  for (int t = 0; t < 8; ++t)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[t][c] = 0.f;

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tw_i = tile % tiles_w;
    const int th_i = (tile / tiles_w) % tiles_h;
    const int b = tile / (tiles_w * tiles_h);
    const int i0 = th_i * kWgTH, j0 = tw_i * kWgTW;
    const float* gb = g + (size_t)b * Hg * Wg;
    const float* mb = gmul ? gmul + (size_t)b * Hg * Wg : nullptr;
    __syncthreads();  // previous tile's readers are done
    for (int r = warp; r < 2 * kWgTH + 2; r += 8) {
      const int ih = 2 * i0 - 1 + r;
      const bool rok = ih >= 0 && ih < Hg;
      for (int cI = lane; cI < 2 * kWgTW + 2; cI += 32) {
        const int iw = 2 * j0 - 1 + cI;
        float v = 0.f;
        if (rok && iw >= 0 && iw < Wg) {
          v = __ldg(gb + (size_t)ih * Wg + iw);
          if (mb) v *= __ldg(mb + (size_t)ih * Wg + iw);
        }
        gs[r * kWgStride + cI] = v;
      }
    }
    __syncthreads();
    const __nv_bfloat16* zt = z + ((size_t)b * H * W) * ldz + chunk * 8;
    auto zaddr = [&](int pp) { return zt + ((size_t)(i0 + (pp >> 5)) * W + (j0 + (pp & 31))) * ldz; };
    uint4 q = __ldg(reinterpret_cast<const uint4*>(zaddr(slot)));
#pragma unroll 2
    for (int pp = slot; pp < kWgTH * kWgTW; pp += PPP) {
      const uint4 cur = q;
      if (pp + PPP < kWgTH * kWgTW) q = __ldg(reinterpret_cast<const uint4*>(zaddr(pp + PPP)));
      const uint32_t u[4] = {cur.x, cur.y, cur.z, cur.w};
      float f[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 t2 = unpack_bf16x2(u[k]);
        f[2 * k] = t2.x;
        f[2 * k + 1] = t2.y;
      }
      const int li = pp >> 5, lj = pp & 31;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float* row = gs + (2 * li + half * 2 + r) * kWgStride + 2 * lj;
        const float2 a = *reinterpret_cast<const float2*>(row);
        const float2 c2 = *reinterpret_cast<const float2*>(row + 2);
        const float gv[4] = {a.x, a.y, c2.x, c2.y};
#pragma unroll
        for (int kw = 0; kw < 4; ++kw)
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[r * 4 + kw][c] = fmaf(f[c], gv[kw], acc[r * 4 + kw][c]);
      }
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < C * 16; k += 256) sum[k] = 0.f;
  __syncthreads();
  for (int s = 0; s < PPP; ++s) {
    if (slot == s) {
#pragma unroll
      for (int t = 0; t < 8; ++t)
#pragma unroll
        for (int c = 0; c < 8; ++c) sum[(chunk * 8 + c) * 16 + half * 8 + t] += acc[t][c];
    }
    __syncthreads();
  }
  for (int k = threadIdx.x; k < C * 16; k += 256) partial[(size_t)blockIdx.x * C * 16 + k] = sum[k];
}

__global__ void partial_sum_kernel(const float* __restrict__ partial, int nblk, int n, float* __restrict__ out, int accumulate) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += (double)partial[(size_t)b * n + k];
  out[k] = accumulate ? out[k] + (float)s : (float)s;
}

// ------------------------------------------------------------------------------------------------
// discriminator head
// ------------------------------------------------------------------------------------------------
// E[h][w][c] = sum_{kh,kw} W5[c][kh][kw] * Wl[(h-kh)*n + (w-kw)],  n = S-3
__global__ void head_weight_kernel(const float* __restrict__ w5, const float* __restrict__ wl, int S, int C, float* __restrict__ E) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= S * S * C) return;
  const int c = idx % C, wq = (idx / C) % S, h = idx / (C * S);
  const int n = S - 3;
  float s = 0.f;
#pragma unroll
  for (int kh = 0; kh < 4; ++kh) {
    const int oh = h - kh;
    if (oh < 0 || oh >= n) continue;
#pragma unroll
    for (int kw = 0; kw < 4; ++kw) {
      const int ow = wq - kw;
      if (ow < 0 || ow >= n) continue;
      s = fmaf(__ldg(w5 + (size_t)c * 16 + kh * 4 + kw), __ldg(wl + oh * n + ow), s);
    }
  }
  E[idx] = s;
}

// partial[b][blockIdx.x] = sum over this block's slice of a[b,:] * E[:]
__global__ void __launch_bounds__(256) head_fwd_kernel(const __nv_bfloat16* __restrict__ a, const float* __restrict__ E, int n, float* __restrict__ partial) {
  const int b = blockIdx.y;
  const __nv_bfloat16* ab = a + (size_t)b * n;
  float s = 0.f;
  for (int k = (blockIdx.x * 256 + threadIdx.x) * 8; k < n; k += gridDim.x * 256 * 8) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(ab + k));
    const float4 e0 = __ldg(reinterpret_cast<const float4*>(E + k)), e1 = __ldg(reinterpret_cast<const float4*>(E + k + 4));
    const float2 f0 = unpack_bf16x2(q.x), f1 = unpack_bf16x2(q.y), f2 = unpack_bf16x2(q.z), f3 = unpack_bf16x2(q.w);
    s += f0.x * e0.x + f0.y * e0.y + f1.x * e0.z + f1.y * e0.w + f2.x * e1.x + f2.y * e1.y + f3.x * e1.z + f3.y * e1.w;
  }
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k];
    partial[(size_t)b * gridDim.x + blockIdx.x] = t;
  }
}

__global__ void head_finalize_kernel(const float* __restrict__ partial, int nsplit, int B, const float* __restrict__ bias, int sigmoid,
                                     float* __restrict__ lin, float* __restrict__ pred) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float s = 0.f;
  for (int k = 0; k < nsplit; ++k) s += partial[(size_t)b * nsplit + k];
  s += __ldg(bias);
  lin[b] = s;
  pred[b] = sigmoid ? 1.f / (1.f + expf(-s)) : s;
}

// dlin[b] = dpred[b] * (sigmoid ? p(1-p) : 1)
__global__ void head_dlin_kernel(const float* __restrict__ dpred, const float* __restrict__ pred, int sigmoid, int B, float* __restrict__ dlin) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float p = pred[b];
  dlin[b] = sigmoid ? dpred[b] * p * (1.f - p) : dpred[b];
}

// da[b][k] = dlin[b] * E[k]
__global__ void __launch_bounds__(256) head_bwd_input_kernel(const float* __restrict__ dlin, const float* __restrict__ E, int n, int B,
                                                             __nv_bfloat16* __restrict__ da) {
  const long long total = (long long)B * (n >> 3);
  for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    const int b = (int)(idx / (n >> 3));
    const int k = (int)(idx - (long long)b * (n >> 3)) * 8;
    const float d = __ldg(dlin + b);
    const float4 e0 = __ldg(reinterpret_cast<const float4*>(E + k)), e1 = __ldg(reinterpret_cast<const float4*>(E + k + 4));
    uint4 o = make_uint4(pack_bf16x2(d * e0.x, d * e0.y), pack_bf16x2(d * e0.z, d * e0.w), pack_bf16x2(d * e1.x, d * e1.y), pack_bf16x2(d * e1.z, d * e1.w));
    *reinterpret_cast<uint4*>(da + (size_t)b * n + k) = o;
  }
}

// dE[k] = sum_b dlin[b] * a[b][k]
__global__ void __launch_bounds__(256) head_bwd_dE_kernel(const float* __restrict__ dlin, const __nv_bfloat16* __restrict__ a, int n, int B,
                                                          float* __restrict__ dE) {
  const int k = (blockIdx.x * 256 + threadIdx.x) * 8;
  if (k >= n) return;
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  for (int b = 0; b < B; ++b) {
    const float d = __ldg(dlin + b);
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(a + (size_t)b * n + k));
    const float2 f0 = unpack_bf16x2(q.x), f1 = unpack_bf16x2(q.y), f2 = unpack_bf16x2(q.z), f3 = unpack_bf16x2(q.w);
    s[0] = fmaf(d, f0.x, s[0]);
    s[1] = fmaf(d, f0.y, s[1]);
    s[2] = fmaf(d, f1.x, s[2]);
    s[3] = fmaf(d, f1.y, s[3]);
    s[4] = fmaf(d, f2.x, s[4]);
    s[5] = fmaf(d, f2.y, s[5]);
    s[6] = fmaf(d, f3.x, s[6]);
    s[7] = fmaf(d, f3.y, s[7]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) dE[k + i] = s[i];
}

// dW5[c][kh][kw] = sum_{oh,ow} dE[oh+kh][ow+kw][c] * Wl[oh*n+ow]
__global__ void head_bwd_w5_kernel(const float* __restrict__ dE, const float* __restrict__ wl, int S, int C, float* __restrict__ dw5, int accumulate) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (tap, c) with c fastest for coalesced dE reads
  if (idx >= 16 * C) return;
  const int c = idx % C, tap = idx / C;
  const int kh = tap >> 2, kw = tap & 3;
  const int n = S - 3;
  float s = 0.f;
  for (int oh = 0; oh < n; ++oh)
    for (int ow = 0; ow < n; ++ow) s = fmaf(__ldg(dE + ((size_t)(oh + kh) * S + (ow + kw)) * C + c), __ldg(wl + oh * n + ow), s);
  float* dst = dw5 + (size_t)c * 16 + tap;
  *dst = accumulate ? *dst + s : s;
}

// dWl[oh*n+ow] = sum_{c,kh,kw} dE[oh+kh][ow+kw][c] * W5[c][kh][kw];   block per output.
// blockIdx.x == n*n computes dbias = sum_b dlin[b].
__global__ void __launch_bounds__(256) head_bwd_wl_kernel(const float* __restrict__ dE, const float* __restrict__ w5, const float* __restrict__ dlin, int B,
                                                          int S, int C, float* __restrict__ dwl, float* __restrict__ dbias, int accumulate) {
  const int n = S - 3;
  __shared__ float red[8];
  float s = 0.f;
  if ((int)blockIdx.x == n * n) {
    for (int b = threadIdx.x; b < B; b += 256) s += dlin[b];
  } else {
    const int oh = blockIdx.x / n, ow = blockIdx.x - oh * n;
    for (int k = threadIdx.x; k < 16 * C; k += 256) {
      const int c = k % C, tap = k / C;
      s = fmaf(__ldg(dE + ((size_t)(oh + (tap >> 2)) * S + (ow + (tap & 3))) * C + c), __ldg(w5 + (size_t)c * 16 + tap), s);
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k];
    float* dst = ((int)blockIdx.x == n * n) ? dbias : dwl + blockIdx.x;
    *dst = accumulate ? *dst + t : t;
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_conv_c1_fwd(const float* x, const float* xmul, int B, int H, int W, const float* w, int C, int act, void* y, void* y2, int ld2, int off2,
                     cudaStream_t stream) {
  const int Ho = H / 2, Wo = W / 2;
  int tow_shift = 5;  // 8 x 32 output tile; 16 x 16 when the image is only 16 pixels wide
  if (Wo < 32) tow_shift = 4;
  const int TOW = 1 << tow_shift, TOH = 256 >> tow_shift;
  if ((C != 64 && C != 128 && C != 32 && C != 256) || (H & 1) || (W & 1) || B < 1 || ld2 % 4 || off2 % 4 || Wo % TOW || Ho % TOH) {
    b200_set_error("conv_c1_fwd: C must be 32/64/128/256; output must tile by 8x32 (or 16x16)");
    return B200_ERR_ARG;
  }
  const int grid = B * (Ho / TOH) * (Wo / TOW);
  conv_c1_fwd_kernel<<<grid, 256, 0, stream>>>(x, xmul, H, W, tow_shift, w, C, act, (__nv_bfloat16*)y, (__nv_bfloat16*)y2, ld2, off2);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_convt_c1_fwd(const void* z, int B, int H, int W, int C, int ldz, const float* w, const float* bias, int act, const float* base, const float* mul,
                      float* y, float* comp, cudaStream_t stream) {
  if ((C != 64 && C != 128) || H % kTT || W % kTT || ldz % 8 || B < 1 || (comp && (!base || !mul))) {
    b200_set_error("convt_c1_fwd: C must be 64/128, H and W multiples of 16");
    return B200_ERR_ARG;
  }
  const int grid = B * (H / kTT) * (W / kTT);
  if (C == 64)
    convt_c1_fwd_kernel<8><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)z, B, H, W, ldz, w, bias, act, base, mul, y, comp);
  else
    convt_c1_fwd_kernel<16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)z, B, H, W, ldz, w, bias, act, base, mul, y, comp);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_wgrad_c1_blocks(void) { return 148 * 2; }

// partial: scratch of b200_wgrad_c1_blocks() * C * 16 floats
int b200_wgrad_c1(const void* z, int B, int H, int W, int C, int ldz, const float* g, const float* gmul, float* partial, float* dw, int accumulate,
                  cudaStream_t stream) {
  if ((C != 64 && C != 128) || ldz % 8 || B < 1 || H % kWgTH || W % kWgTW) {
    b200_set_error("wgrad_c1: C must be 64/128, H % 8 == 0, W % 32 == 0");
    return B200_ERR_ARG;
  }
  const int nblk = b200_wgrad_c1_blocks();
  if (C == 64)
    wgrad_c1_kernel<8><<<nblk, 256, 0, stream>>>((const __nv_bfloat16*)z, B, H, W, ldz, g, gmul, partial);
  else
    wgrad_c1_kernel<16><<<nblk, 256, 0, stream>>>((const __nv_bfloat16*)z, B, H, W, ldz, g, gmul, partial);
  B200_CHECK_LAUNCH();
  partial_sum_kernel<<<(C * 16 + 255) / 256, 256, 0, stream>>>(partial, nblk, C * 16, dw, accumulate);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_head_weight(const float* w5, const float* wl, int S, int C, float* E, cudaStream_t stream) {
  if (S < 4) {
    b200_set_error("head_weight: feature map smaller than the 4x4 kernel");
    return B200_ERR_ARG;
  }
  head_weight_kernel<<<(S * S * C + 255) / 256, 256, 0, stream>>>(w5, wl, S, C, E);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

#define B200_HEAD_SPLIT 8
// a: [B][S*S*C] bf16; scratch: B*8 floats; lin / pred: [B]
int b200_head_fwd(const void* a, const float* E, int B, int n, const float* bias, int sigmoid, float* scratch, float* lin, float* pred, cudaStream_t stream) {
  if (n % 8) {
    b200_set_error("head_fwd: n must be a multiple of 8");
    return B200_ERR_ARG;
  }
  head_fwd_kernel<<<dim3(B200_HEAD_SPLIT, B), 256, 0, stream>>>((const __nv_bfloat16*)a, E, n, scratch);
  B200_CHECK_LAUNCH();
  head_finalize_kernel<<<(B + 127) / 128, 128, 0, stream>>>(scratch, B200_HEAD_SPLIT, B, bias, sigmoid, lin, pred);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_head_dlin(const float* dpred, const float* pred, int sigmoid, int B, float* dlin, cudaStream_t stream) {
  head_dlin_kernel<<<(B + 127) / 128, 128, 0, stream>>>(dpred, pred, sigmoid, B, dlin);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_head_bwd_input(const float* dlin, const float* E, int B, int n, void* da, cudaStream_t stream) {
  long long blocks = ((long long)B * (n / 8) + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  head_bwd_input_kernel<<<(int)blocks, 256, 0, stream>>>(dlin, E, n, B, (__nv_bfloat16*)da);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

// dE: scratch [S*S*C] floats
int b200_head_bwd_params(const float* dlin, const void* a, int B, int S, int C, const float* w5, const float* wl, float* dE, float* dw5, float* dwl,
                         float* dbias, int accumulate, cudaStream_t stream) {
  const int n = S * S * C;
  head_bwd_dE_kernel<<<(n / 8 + 255) / 256, 256, 0, stream>>>(dlin, (const __nv_bfloat16*)a, n, B, dE);
  B200_CHECK_LAUNCH();
  head_bwd_w5_kernel<<<(16 * C + 127) / 128, 128, 0, stream>>>(dE, wl, S, C, dw5, accumulate);
  B200_CHECK_LAUNCH();
  const int nn = (S - 3) * (S - 3);
  head_bwd_wl_kernel<<<nn + 1, 256, 0, stream>>>(dE, w5, dlin, B, S, C, dwl, dbias, accumulate);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

}  // extern "C"
