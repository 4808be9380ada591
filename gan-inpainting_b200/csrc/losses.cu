// Memory-bound tail of the train step over the 1-channel fp32 images [B,1,R,R]:
// mask binarisation + masking, compositing, reconstruction losses (L1 / RMSE / masked "LocalLoss"),
// adversarial heads, and the fused seed of the generator backward pass.
// All kernels are float4-vectorised; reductions are deterministic (per-block partials -> fp64 finalize).
//
// Reference: experiment_list/wgan_l1.py:114-122 (ceil, mask-out, composite), :137-143,177-182 (WGAN heads,
// L1), minimaxgan_rmse.py:135-166 (BCE, RMSE), experiment1_global_local_D.py:162-169 (LSGAN, 2x RMSE),
// lib/models/loss.py:11-47 (RMSELoss, LocalLoss).
#include "common.cuh"

namespace b200 {

enum { RECON_NONE = -1, RECON_L1 = 0, RECON_RMSE = 1, RECON_LOCAL_L1 = 2, RECON_LOCAL_MSE = 3, RECON_MASKED_RMSE = 4 };
enum { ADV_WGAN = 0, ADV_BCE = 1, ADV_LSGAN = 2 };

constexpr int kRedBlocks = 148 * 4;

__device__ __forceinline__ float4 ld4(const float* p, long long i) { return __ldg(reinterpret_cast<const float4*>(p) + i); }

// mask_out = ceil ? ceil(mask) : mask ;  masked = ground * (1 - mask_out)
__global__ void __launch_bounds__(256) mask_prepare_kernel(const float* __restrict__ ground, const float* __restrict__ mask, long long n4, int do_ceil,
                                                           float* __restrict__ mask_out, float* __restrict__ masked) {
  b200::pdl_enter();
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    float4 m = ld4(mask, i);
    const float4 g = ld4(ground, i);
    if (do_ceil) {
      m.x = ceilf(m.x);
      m.y = ceilf(m.y);
      m.z = ceilf(m.z);
      m.w = ceilf(m.w);
    }
    if (mask_out) reinterpret_cast<float4*>(mask_out)[i] = m;
    reinterpret_cast<float4*>(masked)[i] = make_float4(g.x * (1.f - m.x), g.y * (1.f - m.y), g.z * (1.f - m.z), g.w * (1.f - m.w));
  }
}

// out = base + g * mask
__global__ void __launch_bounds__(256) composite_kernel(const float* __restrict__ base, const float* __restrict__ g, const float* __restrict__ mask,
                                                        long long n4, float* __restrict__ out) {
  b200::pdl_enter();
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const float4 b = ld4(base, i), v = ld4(g, i), m = ld4(mask, i);
    reinterpret_cast<float4*>(out)[i] = make_float4(fmaf(v.x, m.x, b.x), fmaf(v.y, m.y, b.y), fmaf(v.z, m.z, b.z), fmaf(v.w, m.w, b.w));
  }
}

// out = a * b
__global__ void __launch_bounds__(256) mul_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n4, float* __restrict__ out) {
  b200::pdl_enter();
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const float4 x = ld4(a, i), y = ld4(b, i);
    reinterpret_cast<float4*>(out)[i] = make_float4(x.x * y.x, x.y * y.y, x.z * y.z, x.w * y.w);
  }
}

__device__ __forceinline__ void recon_term(int kind, float y, float yh, float m, float& s, float& c) {
  if (kind == RECON_L1) {
    s += fabsf(y - yh);
  } else if (kind == RECON_RMSE) {
    const float d = y - yh;
    s += d * d;
  } else if (kind == RECON_LOCAL_L1) {
    s += fabsf(y * m - yh * m);
    c += (m != 0.f) ? 1.f : 0.f;
  } else if (kind == RECON_LOCAL_MSE) {
    const float d = y * m - yh * m;
    s += d * d;
    c += (m != 0.f) ? 1.f : 0.f;
  } else {  // RECON_MASKED_RMSE
    const float d = y * m - yh * m;
    s += d * d;
  }
}

// partial[blk][2] = (sum, count)
__global__ void __launch_bounds__(256) recon_reduce_kernel(int kind, const float* __restrict__ y, const float* __restrict__ yh, const float* __restrict__ mask,
                                                           long long n4, float* __restrict__ partial) {
  b200::pdl_enter();
  float s = 0.f, c = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const float4 a = ld4(y, i), b = ld4(yh, i);
    float4 m = make_float4(1.f, 1.f, 1.f, 1.f);
    if (mask) m = ld4(mask, i);
    recon_term(kind, a.x, b.x, m.x, s, c);
    recon_term(kind, a.y, b.y, m.y, s, c);
    recon_term(kind, a.z, b.z, m.z, s, c);
    recon_term(kind, a.w, b.w, m.w, s, c);
  }
  __shared__ float rs[8], rc[8];
  s = warp_sum(s);
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) {
    rs[threadIdx.x >> 5] = s;
    rc[threadIdx.x >> 5] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ts = 0.f, tc = 0.f;
    for (int k = 0; k < 8; ++k) {
      ts += rs[k];
      tc += rc[k];
    }
    partial[2 * blockIdx.x] = ts;
    partial[2 * blockIdx.x + 1] = tc;
  }
}

// out[0] = loss, out[1] = denominator as float, out[2..3] = exact (mask != 0) pixel count (LocalLoss) or N as two int words
__global__ void recon_finalize_kernel(int kind, const float* __restrict__ partial, int nblk, double n, float eps, float* __restrict__ out) {
  b200::pdl_enter();
  // one warp: lane-strided fp64 partial sums in a fixed order, then a fixed shuffle tree (deterministic)
  double s = 0.0, c = 0.0;
  for (int b = threadIdx.x; b < nblk; b += 32) {
    s += (double)partial[2 * b];
    c += (double)partial[2 * b + 1];  // per-block counts are exact integers < 2^24
  }
  s = warp_sum_d(s);
  c = warp_sum_d(c);
  if (threadIdx.x != 0) return;
  double loss;
  if (kind == RECON_L1) {
    loss = s / n;
    c = n;
  } else if (kind == RECON_RMSE || kind == RECON_MASKED_RMSE) {
    loss = sqrt(s / n + (double)eps);
    c = n;
  } else {
    loss = s / c;
  }
  out[0] = (float)loss;
  out[1] = (float)c;
  // exact integer count (bit-exact (mask != 0) parity check): low / high 31-bit words
  const long long ci = (long long)(c + 0.5);
  reinterpret_cast<int*>(out)[2] = (int)(ci & 0x7fffffffLL);
  reinterpret_cast<int*>(out)[3] = (int)(ci >> 31);
}

// ------------------------------------------------------------------------------------------------
// Seed of the generator backward pass.  With inp = inpainted (composite) or G(masked) (no composite):
//   d_inp = dadv1 + dadv2 + lambda1 * d recon1 / d inp + lambda2 * d recon2 / d inp
//   d_g   = composite ? d_inp * mask : d_inp
//   dpre  = d_g * (1 - g^2)                      (Tanh, networks.py:297)
// and per-block partial sums of dpre (the ConvTranspose2d bias gradient).
// ------------------------------------------------------------------------------------------------
struct SeedParams {
  const float* ground;
  const float* inp;
  const float* g;     // tanh output (null: no tanh in the chain, dpre = d_g)
  const float* mask;  // may be null
  const float* dadv1;
  const float* dadv2;
  int kind1, kind2;
  float lambda1, lambda2;
  const float* loss1;  // out[2] of recon_finalize (loss, count)
  const float* loss2;
  const float* gscale;  // optional device scalar multiplying both reconstruction terms (upstream autograd gradient)
  int composite;
  long long n4;
  float* dpre;
  float* dinp;  // optional: d_inp itself
  float* partial;
};

__device__ __forceinline__ float recon_grad(int kind, float y, float yh, float m, float inv_n, float loss, float cnt) {
  // derivative w.r.t. yh (the generated image); y = ground
  switch (kind) {
    case RECON_L1: {
      const float d = yh - y;
      return (d > 0.f ? inv_n : (d < 0.f ? -inv_n : 0.f));
    }
    case RECON_RMSE:
      return (yh - y) * inv_n / loss;
    case RECON_MASKED_RMSE:
      return m * m * (yh - y) * inv_n / loss;
    case RECON_LOCAL_L1: {
      const float d = yh * m - y * m;
      return m * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) / cnt;
    }
    case RECON_LOCAL_MSE:
      return 2.f * m * (yh * m - y * m) / cnt;
    default:
      return 0.f;
  }
}

__global__ void __launch_bounds__(256) gstep_seed_kernel(const SeedParams p, float inv_n) {
  b200::pdl_enter();
  float bsum = 0.f;
  float l1 = 1.f, c1 = 1.f, l2 = 1.f, c2 = 1.f;
  if (p.kind1 >= 0) {
    l1 = __ldg(p.loss1);
    c1 = __ldg(p.loss1 + 1);
  }
  if (p.kind2 >= 0) {
    l2 = __ldg(p.loss2);
    c2 = __ldg(p.loss2 + 1);
  }
  const float gs = p.gscale ? __ldg(p.gscale) : 1.f;
  const float lam1 = p.lambda1 * gs, lam2 = p.lambda2 * gs;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < p.n4; i += (long long)gridDim.x * 256) {
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.dadv1) d = ld4(p.dadv1, i);
    if (p.dadv2) {
      const float4 e = ld4(p.dadv2, i);
      d.x += e.x;
      d.y += e.y;
      d.z += e.z;
      d.w += e.w;
    }
    float4 m = make_float4(1.f, 1.f, 1.f, 1.f);
    if (p.mask) m = ld4(p.mask, i);
    if (p.kind1 >= 0 || p.kind2 >= 0) {
      const float4 y = ld4(p.ground, i), yh = ld4(p.inp, i);
      if (p.kind1 >= 0) {
        d.x += lam1 * recon_grad(p.kind1, y.x, yh.x, m.x, inv_n, l1, c1);
        d.y += lam1 * recon_grad(p.kind1, y.y, yh.y, m.y, inv_n, l1, c1);
        d.z += lam1 * recon_grad(p.kind1, y.z, yh.z, m.z, inv_n, l1, c1);
        d.w += lam1 * recon_grad(p.kind1, y.w, yh.w, m.w, inv_n, l1, c1);
      }
      if (p.kind2 >= 0) {
        d.x += lam2 * recon_grad(p.kind2, y.x, yh.x, m.x, inv_n, l2, c2);
        d.y += lam2 * recon_grad(p.kind2, y.y, yh.y, m.y, inv_n, l2, c2);
        d.z += lam2 * recon_grad(p.kind2, y.z, yh.z, m.z, inv_n, l2, c2);
        d.w += lam2 * recon_grad(p.kind2, y.w, yh.w, m.w, inv_n, l2, c2);
      }
    }
    if (p.dinp) reinterpret_cast<float4*>(p.dinp)[i] = d;
    if (p.composite) {
      d.x *= m.x;
      d.y *= m.y;
      d.z *= m.z;
      d.w *= m.w;
    }
    if (p.g) {
      const float4 t = ld4(p.g, i);
      d.x *= 1.f - t.x * t.x;
      d.y *= 1.f - t.y * t.y;
      d.z *= 1.f - t.z * t.z;
      d.w *= 1.f - t.w * t.w;
    }
    reinterpret_cast<float4*>(p.dpre)[i] = d;
    bsum += (d.x + d.y) + (d.z + d.w);
  }
  __shared__ float rs[8];
  bsum = warp_sum(bsum);
  if ((threadIdx.x & 31) == 0) rs[threadIdx.x >> 5] = bsum;
  __syncthreads();
  if (threadIdx.x == 0 && p.partial) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += rs[k];
    p.partial[blockIdx.x] = t;
  }
}

__global__ void scalar_sum_kernel(const float* __restrict__ partial, int nblk, float* __restrict__ out, int accumulate) {
  b200::pdl_enter();
  double s = 0.0;
  for (int b = threadIdx.x; b < nblk; b += 32) s += (double)partial[b];
  s = warp_sum_d(s);
  if (threadIdx.x != 0) return;
  *out = accumulate ? *out + (float)s : (float)s;
}

// adversarial heads on pred[B]:  loss_out[0] = loss;  dpred[b] = scale * d loss / d pred[b]
__global__ void __launch_bounds__(256) adv_loss_kernel(const float* __restrict__ pred, int B, int mode, float target, float scale, float* __restrict__ loss_out,
                                                       float* __restrict__ dpred) {
  b200::pdl_enter();
  float s = 0.f;
  const float invB = 1.f / (float)B;
  for (int b = threadIdx.x; b < B; b += 256) {
    const float p = pred[b];
    float l, d;
    if (mode == ADV_WGAN) {
      l = p;
      d = 1.f;
    } else if (mode == ADV_BCE) {
      const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(logf(1.f - p), -100.f);
      l = -(target * lp + (1.f - target) * lq);
      d = (p - target) / fmaxf((1.f - p) * p, 1e-12f);
    } else {
      l = (p - target) * (p - target);
      d = 2.f * (p - target);
    }
    s += l;
    if (dpred) dpred[b] = scale * d * invB;
  }
  __shared__ float rs[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) rs[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += rs[k];
    loss_out[0] = t * invB;
  }
}

static int ew_grid(long long n4) {
  long long g = (n4 + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_recon_scratch_floats(void) { return 2 * kRedBlocks; }

#define B200_REQUIRE_N4(n)                                          \
  if ((n) % 4) {                                                    \
    b200_set_error("element count must be a multiple of 4");        \
    return B200_ERR_ARG;                                            \
  }

int b200_mask_prepare(const float* ground, const float* mask, long long n, int do_ceil, float* mask_out, float* masked, cudaStream_t stream) {
  B200_REQUIRE_N4(n);
  B200_LAUNCH((mask_prepare_kernel), ew_grid(n / 4), 256, 0, stream, ground, mask, n / 4, do_ceil, mask_out, masked);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_composite(const float* base, const float* g, const float* mask, long long n, float* out, cudaStream_t stream) {
  B200_REQUIRE_N4(n);
  B200_LAUNCH((composite_kernel), ew_grid(n / 4), 256, 0, stream, base, g, mask, n / 4, out);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_mul(const float* a, const float* b, long long n, float* out, cudaStream_t stream) {
  B200_REQUIRE_N4(n);
  B200_LAUNCH((mul_kernel), ew_grid(n / 4), 256, 0, stream, a, b, n / 4, out);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

// out[4]: loss, denominator (float), exact denominator (2 int words).  scratch: b200_recon_scratch_floats() floats.
int b200_recon_loss(int kind, const float* y, const float* yhat, const float* mask, long long n, float eps, float* scratch, float* out, cudaStream_t stream) {
  B200_REQUIRE_N4(n);
  if (kind < 0 || kind > 4 || (kind >= 2 && !mask)) {
    b200_set_error("recon_loss: bad kind / missing mask");
    return B200_ERR_ARG;
  }
  int nblk = ew_grid(n / 4);
  if (nblk > kRedBlocks) nblk = kRedBlocks;
  B200_LAUNCH((recon_reduce_kernel), nblk, 256, 0, stream, kind, y, yhat, mask, n / 4, scratch);
  B200_CHECK_LAUNCH();
  B200_LAUNCH((recon_finalize_kernel), 1, 32, 0, stream, kind, scratch, nblk, (double)n, eps, out);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_gstep_seed(const float* ground, const float* inp, const float* g, const float* mask, const float* dadv1, const float* dadv2, int kind1, float lambda1,
                    const float* loss1, int kind2, float lambda2, const float* loss2, const float* gscale, int composite, long long n, float* dpre, float* dinp, float* scratch,
                    float* dbias, int accumulate, cudaStream_t stream) {
  B200_REQUIRE_N4(n);
  if ((composite || kind1 >= 2 || kind2 >= 2) && !mask) {
    b200_set_error("gstep_seed: mask required");
    return B200_ERR_ARG;
  }
  SeedParams p{ground, inp, g, mask, dadv1, dadv2, kind1, kind2, lambda1, lambda2, loss1, loss2, gscale, composite, n / 4, dpre, dinp, dbias ? scratch : nullptr};
  int nblk = ew_grid(n / 4);
  if (nblk > kRedBlocks) nblk = kRedBlocks;
  B200_LAUNCH((gstep_seed_kernel), nblk, 256, 0, stream, p, (float)(1.0 / (double)n));
  B200_CHECK_LAUNCH();
  if (dbias) {
    B200_LAUNCH((scalar_sum_kernel), 1, 32, 0, stream, scratch, nblk, dbias, accumulate);
    B200_CHECK_LAUNCH();
  }
  return B200_OK;
}

int b200_adv_loss(const float* pred, int B, int mode, float target, float scale, float* loss_out, float* dpred, cudaStream_t stream) {
  B200_LAUNCH((adv_loss_kernel), 1, 256, 0, stream, pred, B, mode, target, scale, loss_out, dpred);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

}  // extern "C"
