// Optimizer tail: fused RMSprop / Adam updates over flat fp32 parameter storage (one launch per
// network instead of one foreach chain per tensor), with the WGAN weight clip fused into the update,
// and the per-tensor mean |grad| logging reduction as ONE launch + one D2H copy.
//
// Reference: optim.RMSprop(lr=5e-5) (alpha .99, eps 1e-8, no momentum) wgan_l1.py:64-65;
// optim.Adam(lr=2e-4, betas=(.5,.999)) minimaxgan_rmse.py:64-65; p.data.clamp_(-0.01, 0.01) over ALL
// D parameters wgan_l1.py:152-153; p.grad.abs().mean() per non-bias parameter wgan_l1.py:195-197.
#include "common.cuh"

namespace b200 {

__global__ void __launch_bounds__(256) rmsprop_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ sq, long long n, float lr,
                                                      float alpha, float eps, int clamp, float lo, float hi) {
  b200::pdl_enter();
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    float4 w = reinterpret_cast<float4*>(p)[i];
    const float4 d = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 s = reinterpret_cast<float4*>(sq)[i];
    s.x = alpha * s.x + (1.f - alpha) * d.x * d.x;
    s.y = alpha * s.y + (1.f - alpha) * d.y * d.y;
    s.z = alpha * s.z + (1.f - alpha) * d.z * d.z;
    s.w = alpha * s.w + (1.f - alpha) * d.w * d.w;
    w.x -= lr * d.x / (sqrtf(s.x) + eps);
    w.y -= lr * d.y / (sqrtf(s.y) + eps);
    w.z -= lr * d.z / (sqrtf(s.z) + eps);
    w.w -= lr * d.w / (sqrtf(s.w) + eps);
    if (clamp) {
      w.x = fminf(fmaxf(w.x, lo), hi);
      w.y = fminf(fmaxf(w.y, lo), hi);
      w.z = fminf(fmaxf(w.z, lo), hi);
      w.w = fminf(fmaxf(w.w, lo), hi);
    }
    reinterpret_cast<float4*>(sq)[i] = s;
    reinterpret_cast<float4*>(p)[i] = w;
  }
  // tail (n not a multiple of 4)
  for (long long i = (n4 << 2) + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float d = g[i];
    const float s = alpha * sq[i] + (1.f - alpha) * d * d;
    float w = p[i] - lr * d / (sqrtf(s) + eps);
    if (clamp) w = fminf(fmaxf(w, lo), hi);
    sq[i] = s;
    p[i] = w;
  }
}

// step_count: device int, incremented by the single-thread prologue kernel so the whole step is graph-capturable
__global__ void step_inc_kernel(int* step_count) {
  b200::pdl_enter();
  *step_count += 1;
}

__device__ __forceinline__ float adam_one(float w, float d, float& m, float& v, float b1, float b2, float step_size, float inv_bc2_sqrt, float eps) {
  m = b1 * m + (1.f - b1) * d;
  v = b2 * v + (1.f - b2) * d * d;
  const float denom = sqrtf(v) * inv_bc2_sqrt + eps;
  return w - step_size * m / denom;
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                   long long n, float lr, float b1, float b2, float eps, const int* __restrict__ step_count) {
  b200::pdl_enter();
  const float t = (float)(*step_count);
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  const float step_size = lr / bc1, inv_bc2_sqrt = 1.f / sqrtf(bc2);
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    float4 w = reinterpret_cast<float4*>(p)[i];
    const float4 d = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    w.x = adam_one(w.x, d.x, mm.x, vv.x, b1, b2, step_size, inv_bc2_sqrt, eps);
    w.y = adam_one(w.y, d.y, mm.y, vv.y, b1, b2, step_size, inv_bc2_sqrt, eps);
    w.z = adam_one(w.z, d.z, mm.z, vv.z, b1, b2, step_size, inv_bc2_sqrt, eps);
    w.w = adam_one(w.w, d.w, mm.w, vv.w, b1, b2, step_size, inv_bc2_sqrt, eps);
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    reinterpret_cast<float4*>(p)[i] = w;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    float mm = m[i], vv = v[i];
    p[i] = adam_one(p[i], g[i], mm, vv, b1, b2, step_size, inv_bc2_sqrt, eps);
    m[i] = mm;
    v[i] = vv;
  }
}

__global__ void __launch_bounds__(256) clamp_kernel(float* __restrict__ p, long long n, float lo, float hi) {
  b200::pdl_enter();
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) p[i] = fminf(fmaxf(p[i], lo), hi);
}

// Gradient buckets on the wire in bf16 (optional, data-parallel all-reduce): fp32 -> bf16 before the collective, bf16 -> fp32 after.
// n % 4 == 0 (the flat gradient buffer aligns every tensor to 4 elements); 128-bit loads / 64-bit stores and vice versa.
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n4) {
  b200::pdl_enter();
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    reinterpret_cast<uint2*>(dst)[i] = make_uint2(b200::pack_bf16x2(v.x, v.y), b200::pack_bf16x2(v.z, v.w));
  }
}
__global__ void __launch_bounds__(256) cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, long long n4) {
  b200::pdl_enter();
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const uint2 u = reinterpret_cast<const uint2*>(src)[i];
    const float2 a = b200::unpack_bf16x2(u.x), b = b200::unpack_bf16x2(u.y);
    reinterpret_cast<float4*>(dst)[i] = make_float4(a.x, a.y, b.x, b.y);
  }
}

// out[t] = mean |g[off[t] .. off[t+1])|  — one block per tensor (deterministic)
__global__ void __launch_bounds__(256) absmean_kernel(const float* __restrict__ g, const long long* __restrict__ off, float* __restrict__ out) {
  b200::pdl_enter();
  const long long a = off[blockIdx.x], b = off[blockIdx.x + 1];
  float s = 0.f;
  for (long long i = a + threadIdx.x; i < b; i += 256) s += fabsf(g[i]);
  __shared__ float rs[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) rs[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += rs[k];
    out[blockIdx.x] = b > a ? t / (float)(b - a) : 0.f;
  }
}

static int opt_grid(long long n) {
  long long g = (n / 4 + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_rmsprop_step(float* p, const float* g, float* sq, long long n, float lr, float alpha, float eps, int clamp, float lo, float hi, cudaStream_t stream) {
  if (((uintptr_t)p | (uintptr_t)g | (uintptr_t)sq) & 15) {
    b200_set_error("rmsprop_step: buffers must be 16-byte aligned");
    return B200_ERR_ARG;
  }
  B200_LAUNCH((rmsprop_kernel), opt_grid(n), 256, 0, stream, p, g, sq, n, lr, alpha, eps, clamp, lo, hi);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_step_inc(int* step_count, cudaStream_t stream) {
  B200_LAUNCH((step_inc_kernel), 1, 1, 0, stream, step_count);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, const int* step_count,
                   cudaStream_t stream) {
  if (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) {
    b200_set_error("adam_step: buffers must be 16-byte aligned");
    return B200_ERR_ARG;
  }
  B200_LAUNCH((adam_kernel), opt_grid(n), 256, 0, stream, p, g, m, v, n, lr, b1, b2, eps, step_count);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_clamp(float* p, long long n, float lo, float hi, cudaStream_t stream) {
  B200_LAUNCH((clamp_kernel), opt_grid(n * 4), 256, 0, stream, p, n, lo, hi);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

// to_bf16 != 0: dst (bf16) = src (fp32); else dst (fp32) = src (bf16).  n elements, n % 4 == 0, both pointers 16-byte aligned.
int b200_cast_grads(const void* src, void* dst, long long n, int to_bf16, cudaStream_t stream) {
  if (n < 0 || (n & 3)) {
    b200_set_error("cast_grads: n must be a multiple of 4");
    return B200_ERR_ARG;
  }
  if (n == 0) return B200_OK;
  if (to_bf16) B200_LAUNCH((cast_f32_bf16_kernel), opt_grid(n * 4), 256, 0, stream, (const float*)src, (__nv_bfloat16*)dst, n / 4);
  else B200_LAUNCH((cast_bf16_f32_kernel), opt_grid(n * 4), 256, 0, stream, (const __nv_bfloat16*)src, (float*)dst, n / 4);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_absmean(const float* g, const long long* offsets, int ntensors, float* out, cudaStream_t stream) {
  if (ntensors < 1) return B200_OK;
  B200_LAUNCH((absmean_kernel), ntensors, 256, 0, stream, g, offsets, out);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

}  // extern "C"
