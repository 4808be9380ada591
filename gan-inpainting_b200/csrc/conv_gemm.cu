// Implicit-GEMM convolution kernels on tcgen05 tensor cores (sm_100a).
//
// One warp-specialised kernel body serves the three GEMM shapes of the 4x4 / stride-2 / pad-1
// layers of the generator and discriminators (reference: lib/models/networks.py:285-309, 337-352):
//
//   MODE_CONV_S2  y[b,oh,ow,:]  = sum_{kh,kw,ci} x[b,2oh+kh-1,2ow+kw-1,ci] * W[:,kh,kw,ci]
//                 (Conv2d fprop; also ConvTranspose2d dgrad)
//   MODE_CONVT    y[b,2i+p,2j+q,:] = sum_{a,b',ci} z[b,i+p+a-1,j+q+b'-1,ci] * Wt[(p,q),:,(a,b'),ci]
//                 (ConvTranspose2d fprop as 4 sub-pixel phases; also Conv2d dgrad)
//   MODE_WGRAD    G[(kh,kw)][cs][cl] = sum_{b,i,j} S[b,i,j,cs] * L[b,2i+kh-1,2j+kw-1,cl]
//                 (weight gradients of both layer kinds, reduction over pixels, split across CTAs)
//
// Activations are NHWC bf16.  Operand tiles are staged by TMA (tiled mode, SWIZZLE_128B) straight
// from the activation tensors: the stride-2 gather is expressed by viewing [B,H,W,C] as the 5-D
// tensor [B, H/2, 2, W/2, 2C] (row-parity / column-parity split), so every filter tap is a plain box
// whose out-of-image halo TMA zero-fills.  No im2col buffer, no padded copies.
// Accumulation is fp32 in TMEM; one elected thread issues tcgen05.mma (M=128, N<=256, K=16).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> global).
#include "common.cuh"

#include <mutex>
#include <string>

namespace b200 {

enum { MODE_CONV_S2 = 0, MODE_CONVT = 1, MODE_WGRAD = 2 };

struct IGemmParams {
  int mode;
  int B, Hs, Ws;         // batch, and the low-resolution grid the pixel tiles walk over
  int tw, th, nb;        // pixel box (tw*th*nb = 128 for fprop/dgrad, 64 for wgrad)
  int tiles_w, tiles_h;  // pixel tiles per image row / column
  int num_ptiles;        // tiles_w * tiles_h * ceil(B / nb)
  int C_k;               // fprop/dgrad: input channels; wgrad: channels of the high-resolution tensor L
  int cpt;               // C_k / 64
  int kb_per_cta;        // k-blocks each CTA walks
  int kb_total;          // total k-blocks (wgrad: num_ptiles; otherwise == kb_per_cta)
  int BN;                // N tile
  int Cout;              // fprop/dgrad: output channels (row offset of a phase in the packed weights)
  int ldc;               // output row stride in elements
  void* out;
  int stages;
  uint32_t tmem_cols;
};

constexpr int kATileBytes = 128 * 128;  // 128 rows x 64 bf16 (K-major) or 2 x (64 px x 64 ch) (MN-major)

__global__ void __launch_bounds__(192, 2)
igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const IGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = kATileBytes + p.BN * 128;
  uint64_t* full_bar = (uint64_t*)(smem + p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;
  uint32_t* tmem_ptr_smem = (uint32_t*)(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates
  int w0 = 0, h0 = 0, b0 = 0, n0 = blockIdx.y * p.BN, m0 = 0, phase = 0, ph_p = 0, ph_q = 0;
  int kb_begin = 0, kb_count = p.kb_per_cta;
  if (p.mode == MODE_WGRAD) {
    m0 = blockIdx.x * 128;
    kb_begin = blockIdx.z * p.kb_per_cta;
    kb_count = min(p.kb_per_cta, p.kb_total - kb_begin);
    if (kb_count < 0) kb_count = 0;
  } else {
    const int pt = blockIdx.x;
    const int wt = pt % p.tiles_w;
    const int ht = (pt / p.tiles_w) % p.tiles_h;
    const int bt = pt / (p.tiles_w * p.tiles_h);
    w0 = wt * p.tw;
    h0 = ht * p.th;
    b0 = bt * p.nb;
    phase = blockIdx.z;
    ph_p = phase >> 1;
    ph_q = phase & 1;
  }

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t tx_bytes = (uint32_t)stage_bytes;
      for (int i = 0; i < kb_count; ++i) {
        const int s = i % p.stages;
        const uint32_t par = (uint32_t)((i / p.stages) & 1);
        mbar_wait(&empty_bar[s], par ^ 1u);
        uint8_t* sA = smem + s * stage_bytes;
        uint8_t* sB = sA + kATileBytes;
        mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
        const int kb = kb_begin + i;
        if (p.mode == MODE_CONV_S2) {
          const int tap = kb / p.cpt, cc = kb - tap * p.cpt;
          const int kh = tap >> 2, kw = tap & 3;
          tma_load_5d(sA, &tmA, &full_bar[s], ((kw + 1) & 1) * p.C_k + cc * 64, w0 + ((kw + 1) >> 1) - 1, (kh + 1) & 1,
                      h0 + ((kh + 1) >> 1) - 1, b0);
          tma_load_2d(sB, &tmB, &full_bar[s], kb * 64, n0);
        } else if (p.mode == MODE_CONVT) {
          const int tap = kb / p.cpt, cc = kb - tap * p.cpt;
          const int ta = tap >> 1, tb = tap & 1;
          tma_load_4d(sA, &tmA, &full_bar[s], cc * 64, w0 + ph_q + tb - 1, h0 + ph_p + ta - 1, b0);
          tma_load_2d(sB, &tmB, &full_bar[s], kb * 64, phase * p.Cout + n0);
        } else {
          const int wt = kb % p.tiles_w;
          const int ht = (kb / p.tiles_w) % p.tiles_h;
          const int bt = kb / (p.tiles_w * p.tiles_h);
          const int pw0 = wt * p.tw, ph0 = ht * p.th, pb0 = bt * p.nb;
          tma_load_4d(sA, &tmA, &full_bar[s], m0, pw0, ph0, pb0);
          tma_load_4d(sA + 8192, &tmA, &full_bar[s], m0 + 64, pw0, ph0, pb0);
          const int nsub = p.BN >> 6;
          for (int j = 0; j < nsub; ++j) {
            const int nidx = n0 + 64 * j;
            const int tap = nidx / p.C_k, c0 = nidx - tap * p.C_k;
            const int kh = tap >> 2, kw = tap & 3;
            tma_load_5d(sB + j * 8192, &tmB, &full_bar[s], ((kw + 1) & 1) * p.C_k + c0, pw0 + ((kw + 1) >> 1) - 1, (kh + 1) & 1,
                        ph0 + ((kh + 1) >> 1) - 1, pb0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const bool mn_major = (p.mode == MODE_WGRAD);
      const uint32_t idesc = make_idesc_bf16(128, p.BN, mn_major ? 1 : 0, mn_major ? 1 : 0);
      const uint32_t lbo = mn_major ? 8192u : 16u;
      const uint32_t kstep = mn_major ? 2048u : 32u;  // bytes per K=16 slice inside a stage
      for (int i = 0; i < kb_count; ++i) {
        const int s = i % p.stages;
        const uint32_t par = (uint32_t)((i / p.stages) & 1);
        mbar_wait(&full_bar[s], par);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
        const uint32_t b_addr = a_addr + kATileBytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = make_smem_desc_sw128(a_addr + k * kstep, lbo, 1024u);
          const uint64_t db = make_smem_desc_sw128(b_addr + k * kstep, lbo, 1024u);
          umma_bf16(tmem_base, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs retire
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ===================== epilogue =====================
    const int quad = warp & 3;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int r = quad * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    if (p.mode == MODE_WGRAD) {
      float* dst = (float*)p.out + ((size_t)blockIdx.z * gridDim.x * 128 + (size_t)(m0 + r)) * p.ldc + n0;
      for (int c = 0; c < p.BN; c += 32) {
        uint32_t v[32];
        if (kb_count > 0) {
          tmem_ld_32x32(taddr + c, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(dst + c + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    } else {
      const int per_img = p.th * p.tw;
      const int bl = r / per_img, rem = r - bl * per_img;
      const int hl = rem / p.tw, wl = rem - hl * p.tw;
      const int b = b0 + bl;
      size_t pix;
      if (p.mode == MODE_CONV_S2) {
        pix = ((size_t)b * p.Hs + (h0 + hl)) * p.Ws + (w0 + wl);
      } else {
        pix = ((size_t)b * (2 * p.Hs) + (2 * (h0 + hl) + ph_p)) * (size_t)(2 * p.Ws) + (2 * (w0 + wl) + ph_q);
      }
      __nv_bfloat16* dst = (__nv_bfloat16*)p.out + pix * p.ldc + n0;
      const bool valid = (b < p.B);
      if (p.BN >= 32) {
        for (int c = 0; c < p.BN; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c, v);  // .sync.aligned: all lanes participate even if the row is masked
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 o;
              o.x = pack_bf16x2(__uint_as_float(v[j + 0]), __uint_as_float(v[j + 1]));
              o.y = pack_bf16x2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
              o.z = pack_bf16x2(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
              o.w = pack_bf16x2(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
              *reinterpret_cast<uint4*>(dst + c + j) = o;
            }
          }
        }
      } else {
        uint32_t v[16];
        tmem_ld_32x16(taddr, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(v[j + 0]), __uint_as_float(v[j + 1]));
            o.y = pack_bf16x2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            o.z = pack_bf16x2(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
            o.w = pack_bf16x2(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
            *reinterpret_cast<uint4*>(dst + j) = o;
          }
        }
      }
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// Debug aid: load one TMA box into shared memory and copy the raw (swizzled) bytes out.
__global__ void tma_probe_kernel(const __grid_constant__ CUtensorMap tm, int rank, int c0, int c1, int c2, int c3, int c4, int bytes,
                                 uint8_t* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) smem[i] = 0xEE;
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, (uint32_t)bytes);
    if (rank == 2) tma_load_2d(smem, &tm, &bar, c0, c1);
    if (rank == 4) tma_load_4d(smem, &tm, &bar, c0, c1, c2, c3);
    if (rank == 5) tma_load_5d(smem, &tm, &bar, c0, c1, c2, c3, c4);
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
  });
  return fn;
}

// NHWC bf16 tensor map helpers.  `ld` = channel stride (elements) of one pixel.
static int encode(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    b200_set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return B200_ERR_DRIVER;
  }
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    b200_set_error(("cuTensorMapEncodeTiled failed, CUresult=" + std::to_string((int)r)).c_str());
    return B200_ERR_DRIVER;
  }
  return B200_OK;
}

// plain [B,H,W,C] view, box [64 ch, tw, th, nb]
static int map_nhwc(CUtensorMap* m, const void* base, int B, int H, int W, int C, int ld, int tw, int th, int nb) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t st[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)nb};
  return encode(m, base, 4, dims, st, box);
}
// parity-split view of [B,H,W,C] (ld == C required): [2C, W/2, 2, H/2, B], box [64, tw, 1, th, nb]
static int map_nhwc_parity(CUtensorMap* m, const void* base, int B, int H, int W, int C, int tw, int th, int nb) {
  cuuint64_t dims[5] = {(cuuint64_t)2 * C, (cuuint64_t)W / 2, 2, (cuuint64_t)H / 2, (cuuint64_t)B};
  cuuint64_t st[4] = {(cuuint64_t)2 * C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)2 * W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[5] = {64, (cuuint32_t)tw, 1, (cuuint32_t)th, (cuuint32_t)nb};
  return encode(m, base, 5, dims, st, box);
}
// packed weights [rows][K] (K contiguous), box [64, BN]
static int map_weights(CUtensorMap* m, const void* base, int rows, int K, int BN) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t st[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)BN};
  return encode(m, base, 2, dims, st, box);
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

static void pixel_box(int Hs, int Ws, int npix, int* tw, int* th, int* nb) {
  *tw = Ws < npix ? Ws : npix;
  int rest = npix / *tw;
  *th = Hs < rest ? Hs : rest;
  *nb = rest / *th;
}

static int launch(const CUtensorMap& a, const CUtensorMap& b, IGemmParams& p, dim3 grid, cudaStream_t stream) {
  const int stage_bytes = kATileBytes + p.BN * 128;
  if (p.stages <= 0) {
    p.stages = (100 * 1024) / stage_bytes;
    if (p.stages < 2) p.stages = 2;
    if (p.stages > 6) p.stages = 6;
  }
  uint32_t cols = 32;
  while ((int)cols < p.BN) cols <<= 1;
  p.tmem_cols = cols;
  const size_t smem = (size_t)p.stages * stage_bytes + 1024 + 64 + 16 * p.stages;
  static std::once_flag once;
  std::call_once(once, [] { cudaFuncSetAttribute(igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
  igemm_kernel<<<grid, 192, smem, stream>>>(a, b, p);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

static int pick_bn(int Cout, int bn_hint) {
  if (bn_hint > 0) return bn_hint;
  if (Cout >= 256) return 256;
  return Cout;
}

}  // namespace b200

using namespace b200;

extern "C" {

// y[B,H/2,W/2,ldy] (bf16) = conv4x4/s2/p1(x[B,H,W,Cin] bf16, wpk[Cout][16*Cin] bf16 with k = (kh*4+kw)*Cin + ci)
int b200_conv_s2_gemm(const void* x, int B, int H, int W, int Cin, const void* wpk, int Cout, void* y, int ldy, int bn_hint, int stages,
                      cudaStream_t stream) {
  if (!is_pow2(H) || !is_pow2(W) || H < 2 || W < 2 || Cin % 64 || Cout % 16 || ldy % 8 || B < 1) {
    b200_set_error("conv_s2_gemm: unsupported shape");
    return B200_ERR_ARG;
  }
  IGemmParams p{};
  p.mode = MODE_CONV_S2;
  p.B = B;
  p.Hs = H / 2;
  p.Ws = W / 2;
  pixel_box(p.Hs, p.Ws, 128, &p.tw, &p.th, &p.nb);
  p.tiles_w = p.Ws / p.tw;
  p.tiles_h = p.Hs / p.th;
  const int tiles_b = (B + p.nb - 1) / p.nb;
  p.num_ptiles = p.tiles_w * p.tiles_h * tiles_b;
  p.C_k = Cin;
  p.cpt = Cin / 64;
  p.kb_per_cta = p.kb_total = 16 * p.cpt;
  p.BN = pick_bn(Cout, bn_hint);
  if (Cout % p.BN || (p.BN != 16 && p.BN % 32)) {
    b200_set_error("conv_s2_gemm: bad N tile");
    return B200_ERR_ARG;
  }
  p.Cout = Cout;
  p.ldc = ldy;
  p.out = y;
  p.stages = stages;
  CUtensorMap ta, tb;
  int rc;
  if ((rc = map_nhwc_parity(&ta, x, B, H, W, Cin, p.tw, p.th, p.nb))) return rc;
  if ((rc = map_weights(&tb, wpk, Cout, 16 * Cin, p.BN))) return rc;
  return launch(ta, tb, p, dim3(p.num_ptiles, Cout / p.BN, 1), stream);
}

// y[B,2Hi,2Wi,ldy] (bf16) = convT4x4/s2/p1(z[B,Hi,Wi,Cin] (pixel stride ldz), wpk[4 phases][Cout][4*Cin])
int b200_convt_s2_gemm(const void* z, int B, int Hi, int Wi, int Cin, int ldz, const void* wpk, int Cout, void* y, int ldy, int bn_hint, int stages,
                       cudaStream_t stream) {
  if (!is_pow2(Hi) || !is_pow2(Wi) || Cin % 64 || Cout % 16 || ldy % 8 || ldz % 8 || B < 1) {
    b200_set_error("convt_s2_gemm: unsupported shape");
    return B200_ERR_ARG;
  }
  IGemmParams p{};
  p.mode = MODE_CONVT;
  p.B = B;
  p.Hs = Hi;
  p.Ws = Wi;
  pixel_box(p.Hs, p.Ws, 128, &p.tw, &p.th, &p.nb);
  p.tiles_w = p.Ws / p.tw;
  p.tiles_h = p.Hs / p.th;
  const int tiles_b = (B + p.nb - 1) / p.nb;
  p.num_ptiles = p.tiles_w * p.tiles_h * tiles_b;
  p.C_k = Cin;
  p.cpt = Cin / 64;
  p.kb_per_cta = p.kb_total = 4 * p.cpt;
  p.BN = pick_bn(Cout, bn_hint);
  if (Cout % p.BN || (p.BN != 16 && p.BN % 32)) {
    b200_set_error("convt_s2_gemm: bad N tile");
    return B200_ERR_ARG;
  }
  p.Cout = Cout;
  p.ldc = ldy;
  p.out = y;
  p.stages = stages;
  CUtensorMap ta, tb;
  int rc;
  if ((rc = map_nhwc(&ta, z, B, Hi, Wi, Cin, ldz, p.tw, p.th, p.nb))) return rc;
  if ((rc = map_weights(&tb, wpk, 4 * Cout, 4 * Cin, p.BN))) return rc;
  return launch(ta, tb, p, dim3(p.num_ptiles, Cout / p.BN, 4), stream);
}

// partial[splits][Cs][16*Cl] (fp32): per-split sums of S[b,i,j,cs] * L[b,2i+kh-1,2j+kw-1,cl], column = (kh*4+kw)*Cl + cl.
// S: [B,Hs,Ws,Cs] (pixel stride lds), L: [B,2Hs,2Ws,Cl] dense.  Returns the split count actually used in *splits_io.
int b200_wgrad_gemm(const void* S, int B, int Hs, int Ws, int Cs, int lds, const void* L, int Cl, void* partial, int* splits_io, int bn_hint, int stages,
                    cudaStream_t stream) {
  if (!is_pow2(Hs) || !is_pow2(Ws) || Cs % 128 || Cl % 64 || lds % 8 || B < 1 || !splits_io) {
    b200_set_error("wgrad_gemm: unsupported shape");
    return B200_ERR_ARG;
  }
  IGemmParams p{};
  p.mode = MODE_WGRAD;
  p.B = B;
  p.Hs = Hs;
  p.Ws = Ws;
  pixel_box(Hs, Ws, 64, &p.tw, &p.th, &p.nb);
  p.tiles_w = Ws / p.tw;
  p.tiles_h = Hs / p.th;
  const int tiles_b = (B + p.nb - 1) / p.nb;
  p.num_ptiles = p.tiles_w * p.tiles_h * tiles_b;
  p.C_k = Cl;
  p.cpt = Cl / 64;
  p.BN = bn_hint > 0 ? bn_hint : 256;
  if ((16 * Cl) % p.BN || p.BN % 64) {
    b200_set_error("wgrad_gemm: bad N tile");
    return B200_ERR_ARG;
  }
  int splits = *splits_io;
  if (splits < 1) splits = 1;
  if (splits > p.num_ptiles) splits = p.num_ptiles;
  p.kb_total = p.num_ptiles;
  p.kb_per_cta = (p.num_ptiles + splits - 1) / splits;
  splits = (p.num_ptiles + p.kb_per_cta - 1) / p.kb_per_cta;
  *splits_io = splits;
  p.Cout = Cs;
  p.ldc = 16 * Cl;
  p.out = partial;
  p.stages = stages;
  CUtensorMap ta, tb;
  int rc;
  if ((rc = map_nhwc(&ta, S, B, Hs, Ws, Cs, lds, p.tw, p.th, p.nb))) return rc;
  if ((rc = map_nhwc_parity(&tb, L, B, 2 * Hs, 2 * Ws, Cl, p.tw, p.th, p.nb))) return rc;
  return launch(ta, tb, p, dim3(Cs / 128, 16 * Cl / p.BN, splits), stream);
}

// Debug: raw shared-memory image of one TMA box of the parity view (kind 0) or the plain view (kind 1).
int b200_debug_tma_probe(const void* x, int B, int H, int W, int C, int kind, int tw, int th, int nb, int c0, int c1, int c2, int c3, int c4, void* out,
                         cudaStream_t stream) {
  CUtensorMap tm;
  int rc;
  if (kind == 0) {
    if ((rc = map_nhwc_parity(&tm, x, B, H, W, C, tw, th, nb))) return rc;
  } else {
    if ((rc = map_nhwc(&tm, x, B, H, W, C, C, tw, th, nb))) return rc;
  }
  const int bytes = 128 * tw * th * nb;
  cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  tma_probe_kernel<<<1, 128, bytes + 1024, stream>>>(tm, kind == 0 ? 5 : 4, c0, c1, c2, c3, c4, bytes, (uint8_t*)out);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

}  // extern "C"
