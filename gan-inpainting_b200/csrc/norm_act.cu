// BatchNorm (training / eval) + activation passes over NHWC bf16 activations — the memory-bound
// glue between the tensor-core convolutions.  All kernels are 128-bit vectorised (8 bf16 channels
// per thread), per-channel statistics are reduced deterministically in two stages (per-block fp32
// partials -> fp64 finalize), no atomics.
//
// Reference semantics: nn.BatchNorm2d(affine, track_running_stats, eps 1e-5, momentum 0.1)
// (lib/models/networks.py:38, 341-349), LeakyReLU(0.2) / ReLU prologues applied in place on the
// block input (networks.py:287-289), Dropout(0.5) after the two inner up-norms
// (networks.py:313-314).
#include "common.cuh"

namespace b200 {

enum { ACT_NONE = 0, ACT_LEAKY = 1, ACT_RELU = 2 };

constexpr int kStatThreads = 256;

// ---------------------------------------------------------------- statistics: sum, sum of squares per channel
// partial[blk][2][C].  C is a power of two (chunk index by mask / shift, no integer division in the loop).
__global__ void __launch_bounds__(kStatThreads) bn_stats_kernel(const __nv_bfloat16* __restrict__ y, int P, int C, int cshift, float* __restrict__ partial) {
  b200::pdl_enter();
  const int chunks = C >> 3;                // 8-channel chunks per pixel
  const int lanes = kStatThreads >> cshift;  // pixel lanes per block
  const int chunk = threadIdx.x & (chunks - 1);
  const int pl = threadIdx.x >> cshift;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  if (pl < lanes) {
    const int stride = gridDim.x * lanes;
    const __nv_bfloat16* base = y + chunk * 8;
    int pix = blockIdx.x * lanes + pl;
    // two independent loads in flight per thread
    for (; pix + stride < P; pix += 2 * stride) {
      const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(base + (size_t)pix * C));
      const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(pix + stride) * C));
      const uint32_t u[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 f = unpack_bf16x2(u[i]);
        s[2 * (i & 3)] += f.x;
        q[2 * (i & 3)] += f.x * f.x;
        s[2 * (i & 3) + 1] += f.y;
        q[2 * (i & 3) + 1] += f.y * f.y;
      }
    }
    if (pix < P) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + (size_t)pix * C));
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack_bf16x2(u[i]);
        s[2 * i] += f.x;
        q[2 * i] += f.x * f.x;
        s[2 * i + 1] += f.y;
        q[2 * i + 1] += f.y * f.y;
      }
    }
  }
  __shared__ float red[kStatThreads][17];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[threadIdx.x][i] = s[i];
    red[threadIdx.x][8 + i] = q[i];
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 2 * C; o += kStatThreads) {
    const int k = o >= C, c = o - k * C;
    const int ch = c >> 3, ci = c & 7;
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += red[l * chunks + ch][k * 8 + ci];
    partial[((size_t)blockIdx.x * 2 + k) * C + c] = acc;
  }
}

// Sum the per-block partials of 32 consecutive channels: thread (w, l) walks rows w, w+32, ... of channel c0+l (coalesced
// 128-byte rows, independent loads), then 8 threads per channel fold the 32 walkers (4 adds + 3 shuffle steps) — a fixed
// order, hence deterministic; fp64 throughout.  The result is valid in the threads with (threadIdx.x & 7) == 0 of the first
// 256 threads: channel = blockIdx.x*32 + (threadIdx.x >> 3).
constexpr int kFinThreads = 1024;
// element (blk, k, c) lives at partial[blk * row_stride + k * k_stride + c]
__device__ __forceinline__ void reduce_partials(const float* __restrict__ partial, int nblk, int C, int c_lane, double& s, double& q, int row_stride = 0,
                                                int k_stride = 0) {
  if (row_stride == 0) {
    row_stride = 2 * C;
    k_stride = C;
  }
  __shared__ double sh[2][32][33];
  const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
  double a = 0.0, b = 0.0;
  if (c_lane < C) {
    for (int blk = w; blk < nblk; blk += 32) {
      a += (double)partial[(size_t)blk * row_stride + c_lane];
      b += (double)partial[(size_t)blk * row_stride + k_stride + c_lane];
    }
  }
  sh[0][w][l] = a;
  sh[1][w][l] = b;
  __syncthreads();
  s = 0.0;
  q = 0.0;
  if (threadIdx.x < 256) {
    const int ch = threadIdx.x >> 3, part = threadIdx.x & 7;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s += sh[0][part * 4 + k][ch];
      q += sh[1][part * 4 + k][ch];
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
  }
}

// mean / rstd / scale / shift (+ running-stat update) from the partials.  Block = 32 channels x 32 block walkers.
// stats_out[4][C] = mean, rstd, scale = gamma*rstd, shift = beta - mean*scale
__global__ void __launch_bounds__(kFinThreads) bn_finalize_kernel(const float* __restrict__ partial, int nblk, int C, double count,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  float* __restrict__ running_mean, float* __restrict__ running_var,
                                                                  long long* __restrict__ num_batches_tracked, float momentum, float eps,
                                                                  float* __restrict__ stats_out) {
  b200::pdl_enter();
  double s, q;
  reduce_partials(partial, nblk, C, blockIdx.x * 32 + (threadIdx.x & 31), s, q);
  const int c = blockIdx.x * 32 + (threadIdx.x >> 3);
  if (threadIdx.x >= 256 || (threadIdx.x & 7) || c >= C) return;
  if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;
  const double mean = s / count;
  double var = q / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
  const float scale = g * rstd;
  stats_out[0 * C + c] = (float)mean;
  stats_out[1 * C + c] = rstd;
  stats_out[2 * C + c] = scale;
  stats_out[3 * C + c] = bt - (float)mean * scale;
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// eval-mode: scale/shift from the running statistics.
__global__ void bn_eval_coeffs_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ running_mean,
                                      const float* __restrict__ running_var, float eps, float* __restrict__ stats_out) {
  b200::pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float rstd = 1.f / sqrtf(running_var[c] + eps);
  const float scale = gamma[c] * rstd;
  stats_out[0 * C + c] = running_mean[c];
  stats_out[1 * C + c] = rstd;
  stats_out[2 * C + c] = scale;
  stats_out[3 * C + c] = beta[c] - running_mean[c] * scale;
}

// ---------------------------------------------------------------- apply: out = act(y*scale + shift) [dropout], optional second store relu(out)
struct ApplyParams {
  const __nv_bfloat16* y;
  int P;
  int C;
  int cshift;  // log2(C / 8)
  const float* scale;  // may be null (identity)
  const float* shift;
  int act;
  __nv_bfloat16* out1;  // may be null
  int ld1, off1;
  __nv_bfloat16* out2;  // relu(out) copy, may be null
  int ld2, off2;
  int dropout;  // 1 -> keep-mask from (seed, element index), kept values x2
  uint32_t seed;
  const uint32_t* seed_ctr;  // optional device counter (b200_set_seed_counter)
};

__device__ __forceinline__ void load_coef8(const float* __restrict__ p, float (&o)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
  o[0] = a.x, o[1] = a.y, o[2] = a.z, o[3] = a.w, o[4] = b.x, o[5] = b.y, o[6] = b.z, o[7] = b.w;
}

// sc / sh: this thread's 8 channels of scale / shift.  A thread keeps its channel group for the whole grid-stride loop (the stride,
// gridDim.x * 256, is a multiple of C / 8 <= 256), so they are loaded ONCE per thread instead of four 16-byte loads per 16 bytes of data.
__device__ __forceinline__ void bn_act_one(const ApplyParams& p, uint32_t seed, unsigned idx, int c0, const float (&sc)[8], const float (&sh)[8], const uint4 v) {
  const size_t pix = idx >> p.cshift;
  const uint32_t u[4] = {v.x, v.y, v.z, v.w};
  float f[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = unpack_bf16x2(u[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
  if (p.scale) {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], sc[i], sh[i]);
  }
  if (p.dropout) {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = dropout_keep(seed, (uint64_t)(pix * p.C + c0 + i)) ? 2.f * f[i] : 0.f;
  }
  if (p.act == ACT_LEAKY) {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = f[i] > 0.f ? f[i] : 0.2f * f[i];
  } else if (p.act == ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
  }
  if (p.out1) {
    uint4 o = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    *reinterpret_cast<uint4*>(p.out1 + pix * p.ld1 + p.off1 + c0) = o;
  }
  if (p.out2) {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
    uint4 o = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    *reinterpret_cast<uint4*>(p.out2 + pix * p.ld2 + p.off2 + c0) = o;
  }
}

__global__ void __launch_bounds__(256) bn_act_kernel(const ApplyParams p) {
  b200::pdl_enter();
  const unsigned total = (unsigned)p.P << p.cshift;
  const unsigned cmask = (1u << p.cshift) - 1u;
  const uint32_t seed = p.dropout ? dropout_seed(p.seed, p.seed_ctr) : 0u;
  // software-pipelined, two 16-byte elements per thread and iteration: the next pair's loads are in flight while this pair is
  // processed (up to 4 x 16 B outstanding per thread; 2048 threads x 16 B per SM is less than the bandwidth-latency product of HBM3e)
  const unsigned stride = gridDim.x * blockDim.x;
  unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = (int)(idx & cmask) * 8;  // loop-invariant: stride % (C / 8) == 0
  float sc[8], sh[8];
  if (p.scale) {
    load_coef8(p.scale + c0, sc);
    load_coef8(p.shift + c0, sh);
  }
  auto ld = [&](unsigned i) {
    return i < total ? __ldg(reinterpret_cast<const uint4*>(p.y + (size_t)(i >> p.cshift) * p.C + c0)) : make_uint4(0u, 0u, 0u, 0u);
  };
  uint4 n0 = ld(idx), n1 = ld(idx + stride);
  for (; idx < total; idx += 2 * stride) {
    const uint4 v0 = n0, v1 = n1;
    n0 = ld(idx + 2 * stride);
    n1 = ld(idx + 3 * stride);
    bn_act_one(p, seed, idx, c0, sc, sh, v0);
    if (idx + stride < total) bn_act_one(p, seed, idx + stride, c0, sc, sh, v1);
  }
}

// ---------------------------------------------------------------- backward
// dz (gradient w.r.t. the pre-activation z = y*scale+shift) from up to two incoming gradients:
//   ACT_LEAKY: dz = z>0 ? g1+g2 : 0.2*g1     (g2 = skip-connection gradient that passed a ReLU of leaky(z): only where z>0)
//   ACT_RELU : dz = z>0 ? g1 : 0             (with dropout: dz = keep ? 2*g1*[z>0] : 0)
//   ACT_NONE : dz = g1
struct BwdParams {
  const __nv_bfloat16* y;  // raw conv output
  int P;
  int C;
  int cshift;  // log2(C / 8)
  const float* scale;  // forward scale/shift (null: identity)
  const float* shift;
  int act;
  const __nv_bfloat16* g1;
  int ldg1, offg1;
  const __nv_bfloat16* g2;  // may be null
  int ldg2, offg2;
  int dropout;
  uint32_t seed;
  const uint32_t* seed_ctr;
  // reduce: partial[blk][2][C] = sum dz, sum dz*y
  float* partial;
  // apply: dy = kdz*dz + ky*y + k0   (coef[3][C]; null: dy = dz)
  const float* coef;
  __nv_bfloat16* dy;
};

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = unpack_bf16x2(u[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__device__ __forceinline__ void compute_dz(const BwdParams& p, size_t pix, int c0, float (&yv)[8], float (&dz)[8]) {
  load8(p.y + pix * p.C + c0, yv);
  float g1[8];
  load8(p.g1 + pix * p.ldg1 + p.offg1 + c0, g1);
  float z[8];
  if (p.scale) {
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.scale + c0)), s1 = __ldg(reinterpret_cast<const float4*>(p.scale + c0 + 4));
    const float4 h0 = __ldg(reinterpret_cast<const float4*>(p.shift + c0)), h1 = __ldg(reinterpret_cast<const float4*>(p.shift + c0 + 4));
    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const float sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = fmaf(yv[i], sc[i], sh[i]);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = yv[i];
  }
  if (p.act == ACT_LEAKY) {
    if (p.g2) {
      float g2[8];
      load8(p.g2 + pix * p.ldg2 + p.offg2 + c0, g2);
#pragma unroll
      for (int i = 0; i < 8; ++i) dz[i] = z[i] > 0.f ? g1[i] + g2[i] : 0.2f * g1[i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) dz[i] = z[i] > 0.f ? g1[i] : 0.2f * g1[i];
    }
  } else if (p.act == ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dz[i] = z[i] > 0.f ? g1[i] : 0.f;
    if (p.dropout) {
      const uint32_t seed = dropout_seed(p.seed, p.seed_ctr);
#pragma unroll
      for (int i = 0; i < 8; ++i) dz[i] = dropout_keep(seed, (uint64_t)(pix * p.C + c0 + i)) ? 2.f * dz[i] : 0.f;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) dz[i] = g1[i];
  }
}

__global__ void __launch_bounds__(kStatThreads) bn_bwd_reduce_kernel(const BwdParams p) {
  b200::pdl_enter();
  const int chunks = p.C >> 3;
  const int lanes = kStatThreads >> p.cshift;
  const int chunk = threadIdx.x & (chunks - 1);
  const int pl = threadIdx.x >> p.cshift;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  if (pl < lanes) {
    const int stride = gridDim.x * lanes;
    for (int pix = blockIdx.x * lanes + pl; pix < p.P; pix += stride) {
      float yv[8], dz[8];
      compute_dz(p, (size_t)pix, chunk * 8, yv, dz);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += dz[i];
        q[i] += dz[i] * yv[i];
      }
    }
  }
  __shared__ float red[kStatThreads][17];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[threadIdx.x][i] = s[i];
    red[threadIdx.x][8 + i] = q[i];
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 2 * p.C; o += kStatThreads) {
    const int k = o >= p.C, c = o - k * p.C;
    const int ch = c >> 3, ci = c & 7;
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += red[l * chunks + ch][k * 8 + ci];
    p.partial[((size_t)blockIdx.x * 2 + k) * p.C + c] = acc;
  }
}

// dgamma / dbeta (+= if accumulate) and the apply coefficients.  stats[4][C] from the forward.
__global__ void __launch_bounds__(kFinThreads) bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblk, int C, double count,
                                                                      const float* __restrict__ stats, const float* __restrict__ gamma,
                                                                      float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate,
                                                                      float* __restrict__ coef, int row_stride, int k_stride) {
  b200::pdl_enter();
  double sdz, sdzy;
  reduce_partials(partial, nblk, C, blockIdx.x * 32 + (threadIdx.x & 31), sdz, sdzy, row_stride, k_stride);
  const int c = blockIdx.x * 32 + (threadIdx.x >> 3);
  if (threadIdx.x >= 256 || (threadIdx.x & 7) || c >= C) return;
  const double mean = stats[0 * C + c], rstd = stats[1 * C + c];
  const double dg = rstd * (sdzy - mean * sdz);  // sum dz * xhat
  const double db = sdz;
  if (dgamma) {
    if (accumulate) {
      dgamma[c] += (float)dg;
      dbeta[c] += (float)db;
    } else {
      dgamma[c] = (float)dg;
      dbeta[c] = (float)db;
    }
  }
  const double a = (double)(gamma ? gamma[c] : 1.f) * rstd;
  coef[0 * C + c] = (float)a;
  coef[1 * C + c] = (float)(-a * rstd * dg / count);
  coef[2 * C + c] = (float)(-a * db / count + a * rstd * mean * dg / count);
}

// dy = kdz*dz + ky*y + k0 with dz read as fp32 (the unrounded output of a fused GEMM epilogue).  The 24 coefficients of the
// thread's 8 channels live in registers (loop-invariant channel group); y and dz of the next element are in flight while this one is
// finished (16 + 32 bytes per element and thread, dz as one 256-bit load).
__device__ __forceinline__ void ldg256(const float* p, float (&o)[8]) {
  uint32_t r[8];
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(256, 4) bn_bwd_apply_f32_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ dz, int lddz, int offdz, int P,
                                                               int C, int cshift, const float* __restrict__ coef, __nv_bfloat16* __restrict__ dy) {
  b200::pdl_enter();
  const unsigned total = (unsigned)P << cshift;
  const unsigned cmask = (1u << cshift) - 1u;
  const unsigned stride = gridDim.x * blockDim.x;
  unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = (int)(idx & cmask) * 8;  // loop-invariant: stride % (C / 8) == 0
  float ka[8], kb[8], kd[8];
  load_coef8(coef + c0, ka);
  load_coef8(coef + C + c0, kb);
  load_coef8(coef + 2 * C + c0, kd);
  const __nv_bfloat16* yp = y + c0;
  const float* zp = dz + offdz + c0;
  __nv_bfloat16* op = dy + c0;
  uint4 yn = make_uint4(0u, 0u, 0u, 0u);
  float zn[8];
  if (idx < total) {
    const size_t pix = idx >> cshift;
    yn = __ldg(reinterpret_cast<const uint4*>(yp + pix * C));
    ldg256(zp + pix * lddz, zn);
  }
  for (; idx < total; idx += stride) {
    const size_t pix = idx >> cshift;
    const uint4 yc = yn;
    float zz[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) zz[i] = zn[i];
    const unsigned nx = idx + stride;
    if (nx < total) {
      const size_t pn = nx >> cshift;
      yn = __ldg(reinterpret_cast<const uint4*>(yp + pn * C));
      ldg256(zp + pn * lddz, zn);
    }
    const uint32_t u[4] = {yc.x, yc.y, yc.z, yc.w};
    float o[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = unpack_bf16x2(u[i]);
      o[2 * i] = fmaf(ka[2 * i], zz[2 * i], fmaf(kb[2 * i], t.x, kd[2 * i]));
      o[2 * i + 1] = fmaf(ka[2 * i + 1], zz[2 * i + 1], fmaf(kb[2 * i + 1], t.y, kd[2 * i + 1]));
    }
    *reinterpret_cast<uint4*>(op + pix * C) = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
  }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const BwdParams p) {
  b200::pdl_enter();
  const unsigned total = (unsigned)p.P << p.cshift;
  const unsigned cmask = (1u << p.cshift) - 1u;
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const size_t pix = idx >> p.cshift;
    const int c0 = (int)(idx & cmask) * 8;
    float yv[8], dz[8];
    compute_dz(p, pix, c0, yv, dz);
    if (p.coef) {
      const float* k = p.coef + c0;
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(k)), a1 = __ldg(reinterpret_cast<const float4*>(k + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(k + p.C)), b1 = __ldg(reinterpret_cast<const float4*>(k + p.C + 4));
      const float4 d0 = __ldg(reinterpret_cast<const float4*>(k + 2 * p.C)), d1 = __ldg(reinterpret_cast<const float4*>(k + 2 * p.C + 4));
      const float ka[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float kb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      const float kd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) dz[i] = fmaf(ka[i], dz[i], fmaf(kb[i], yv[i], kd[i]));
    }
    uint4 o = make_uint4(pack_bf16x2(dz[0], dz[1]), pack_bf16x2(dz[2], dz[3]), pack_bf16x2(dz[4], dz[5]), pack_bf16x2(dz[6], dz[7]));
    *reinterpret_cast<uint4*>(p.dy + pix * p.C + c0) = o;
  }
}

static int grid_for(long long work_items, int threads) {
  long long g = (work_items + threads - 1) / threads;
  const long long cap = 148LL * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace b200

using namespace b200;

extern "C" {

static int chunk_shift(int C) {  // log2(C / 8) for power-of-two C in [8, 2048], else -1
  for (int s = 0; s <= 8; ++s)
    if ((8 << s) == C) return s;
  return -1;
}

int b200_bn_stat_blocks(long long P, int C) {
  const int lanes = kStatThreads / (C >> 3);
  long long g = (P + 2 * lanes - 1) / (2 * lanes);
  // two pixels in flight per thread; at most 4 resident blocks on each of the 148 SMs
  long long cap = 148 * 4;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// y[P][C] bf16 -> partial[nblk][2][C]; then finalize -> stats[4][C] (+ running stats).  training mode.
int b200_bn_stats(const void* y, long long P, int C, void* partial, const float* gamma, const float* beta, float* running_mean, float* running_var,
                  long long* num_batches_tracked, float momentum, float eps, float* stats_out, cudaStream_t stream) {
  const int cs = chunk_shift(C);
  if (cs < 0 || P < 1 || P * (C >> 3) >= (1LL << 31)) {
    b200_set_error("bn_stats: C must be a power of two in [8, 2048] and P*C/8 < 2^31");
    return B200_ERR_ARG;
  }
  const int nblk = b200_bn_stat_blocks(P, C);
  B200_LAUNCH((bn_stats_kernel), nblk, kStatThreads, 0, stream, (const __nv_bfloat16*)y, (int)P, C, cs, (float*)partial);
  B200_CHECK_LAUNCH();
  B200_LAUNCH((bn_finalize_kernel), (C + 31) / 32, kFinThreads, 0, stream, (const float*)partial, nblk, C, (double)P, gamma, beta, running_mean, running_var,
                                                                num_batches_tracked, momentum, eps, stats_out);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

// Finalize from per-CTA partial rows written by the convolution epilogue (b200_conv_s2_gemm / b200_convt_s2_gemm `stats`).
int b200_bn_finalize(const float* partial, int rows, long long P, int C, const float* gamma, const float* beta, float* running_mean, float* running_var,
                     long long* num_batches_tracked, float momentum, float eps, float* stats_out, cudaStream_t stream) {
  if (rows < 1 || P < 1 || C < 1) {
    b200_set_error("bn_finalize: bad sizes");
    return B200_ERR_ARG;
  }
  B200_LAUNCH((bn_finalize_kernel), (C + 31) / 32, kFinThreads, 0, stream, partial, rows, C, (double)P, gamma, beta, running_mean, running_var, num_batches_tracked, momentum,
                                                                eps, stats_out);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_bn_eval_coeffs(int C, const float* gamma, const float* beta, const float* running_mean, const float* running_var, float eps, float* stats_out,
                        cudaStream_t stream) {
  B200_LAUNCH((bn_eval_coeffs_kernel), (C + 127) / 128, 128, 0, stream, C, gamma, beta, running_mean, running_var, eps, stats_out);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_bn_act(const void* y, long long P, int C, const float* scale, const float* shift, int act, void* out1, int ld1, int off1, void* out2, int ld2,
                int off2, int dropout, unsigned seed, cudaStream_t stream) {
  const int cs = chunk_shift(C);
  if (cs < 0 || ld1 % 8 || off1 % 8 || ld2 % 8 || off2 % 8 || P * (C >> 3) >= (1LL << 31)) {
    b200_set_error("bn_act: C must be a power of two >= 8; strides / offsets multiples of 8");
    return B200_ERR_ARG;
  }
  ApplyParams p{(const __nv_bfloat16*)y, (int)P, C, cs, scale, shift, act, (__nv_bfloat16*)out1, ld1, off1, (__nv_bfloat16*)out2, ld2, off2, dropout, seed, b200_seed_counter_ptr()};
  B200_LAUNCH((bn_act_kernel), grid_for(P * (C >> 3), 256), 256, 0, stream, p);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

// Backward of (BN ->) [dropout ->] activation.  If `stats` is non-null (training-mode BN) the
// per-channel reductions run, dgamma/dbeta are produced and dy carries the batch-statistics terms;
// otherwise dy = scale * dz (eval BN) or dz (no norm).
int b200_bn_act_bwd(const void* y, long long P, int C, const float* stats /*[4][C] or null*/, const float* gamma, int act, const void* g1, int ldg1,
                    int offg1, const void* g2, int ldg2, int offg2, int dropout, unsigned seed, void* partial, float* coef /*[3][C] scratch*/,
                    float* dgamma, float* dbeta, int accumulate, void* dy, cudaStream_t stream) {
  const int cs = chunk_shift(C);
  if (cs < 0 || ldg1 % 8 || offg1 % 8 || ldg2 % 8 || offg2 % 8 || P * (C >> 3) >= (1LL << 31)) {
    b200_set_error("bn_act_bwd: C must be a power of two in [8, 2048]; strides / offsets multiples of 8");
    return B200_ERR_ARG;
  }
  BwdParams p{};
  p.y = (const __nv_bfloat16*)y;
  p.P = (int)P;
  p.C = C;
  p.cshift = cs;
  p.scale = stats ? stats + 2 * C : nullptr;
  p.shift = stats ? stats + 3 * C : nullptr;
  p.act = act;
  p.g1 = (const __nv_bfloat16*)g1;
  p.ldg1 = ldg1;
  p.offg1 = offg1;
  p.g2 = (const __nv_bfloat16*)g2;
  p.ldg2 = ldg2;
  p.offg2 = offg2;
  p.dropout = dropout;
  p.seed = seed;
  p.seed_ctr = b200_seed_counter_ptr();
  p.partial = (float*)partial;
  p.coef = nullptr;
  p.dy = (__nv_bfloat16*)dy;
  if (stats) {
    const int nblk = b200_bn_stat_blocks(P, C);
    B200_LAUNCH((bn_bwd_reduce_kernel), nblk, kStatThreads, 0, stream, p);
    B200_CHECK_LAUNCH();
    B200_LAUNCH((bn_bwd_finalize_kernel), (C + 31) / 32, kFinThreads, 0, stream, (const float*)partial, nblk, C, (double)P, stats, gamma, dgamma, dbeta, accumulate, coef, 0, 0);
    B200_CHECK_LAUNCH();
    p.coef = coef;
  }
  B200_LAUNCH((bn_bwd_apply_kernel), grid_for(P * (C >> 3), 256), 256, 0, stream, p);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

// BatchNorm backward when the producing data-gradient GEMM already wrote dz (gradient w.r.t. the BatchNorm output) and the
// per-CTA partial rows (sum dz, sum dz*y) — b200_conv_s2_gemm / b200_convt_s2_gemm with a b200_epi_bwd.  Element (row, k, c) of
// the partials is partial[row * row_stride + k * k_stride + c].  Produces dgamma / dbeta and dy = gradient w.r.t. the conv output y.
int b200_bn_bwd_fused(const void* y, long long P, int C, const float* stats, const float* gamma, const void* dz, int lddz, int offdz, int dz_is_f32,
                      const float* partial, int rows, int row_stride, int k_stride, float* coef, float* dgamma, float* dbeta, int accumulate, void* dy,
                      cudaStream_t stream) {
  const int cs = chunk_shift(C);
  if (cs < 0 || !stats || lddz % 8 || offdz % 8 || rows < 1 || P * (C >> 3) >= (1LL << 31) || (dz_is_f32 && ((uintptr_t)dz & 31))) {
    b200_set_error("bn_bwd_fused: C must be a power of two in [8, 2048]; strides / offsets multiples of 8");
    return B200_ERR_ARG;
  }
  B200_LAUNCH((bn_bwd_finalize_kernel), (C + 31) / 32, kFinThreads, 0, stream, partial, rows, C, (double)P, stats, gamma, dgamma, dbeta, accumulate, coef, row_stride,
                                                                    k_stride);
  B200_CHECK_LAUNCH();
  if (dz_is_f32) {
    B200_LAUNCH((bn_bwd_apply_f32_kernel), grid_for(P * (C >> 3), 256), 256, 0, stream, (const __nv_bfloat16*)y, (const float*)dz, lddz, offdz, (int)P, C, cs, coef,
                                                                              (__nv_bfloat16*)dy);
    B200_CHECK_LAUNCH();
    return B200_OK;
  }
  BwdParams p{};
  p.y = (const __nv_bfloat16*)y;
  p.P = (int)P;
  p.C = C;
  p.cshift = cs;
  p.act = ACT_NONE;  // g1 already is dz
  p.g1 = (const __nv_bfloat16*)dz;
  p.ldg1 = lddz;
  p.offg1 = offdz;
  p.coef = coef;
  p.dy = (__nv_bfloat16*)dy;
  B200_LAUNCH((bn_bwd_apply_kernel), grid_for(P * (C >> 3), 256), 256, 0, stream, p);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

}  // extern "C"
