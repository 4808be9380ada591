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
  int kb_per_tile;       // k-blocks one output tile accumulates (wgrad: pixel tiles per split)
  int kb_total;          // wgrad: all pixel tiles; otherwise == kb_per_tile
  int BN;                // N tile
  int Cout;              // fprop/dgrad: output channels (row offset of a phase in the packed weights)
  int ldc;               // output row stride in elements
  void* out;
  int stages;
  uint32_t tmem_cols;    // 2 accumulator buffers of BN columns, rounded up to a power of two
  int n_tiles;           // N tiles
  int phases;            // 4 for the sub-pixel phases of the transposed convolution, else 1
  int m_tiles;           // wgrad: Cs / 128
  int num_tiles;         // work items of the persistent tile loop
  // every divisor on the single-thread producer / issuer paths is a power of two: shifts instead of integer division
  int tw_sh, th_sh, tiles_w_sh, tiles_h_sh, ck_sh, cpt_sh;
  int kpb;               // 64-wide k-blocks per pipeline stage (2 for narrow N tiles: halves the per-stage barrier overhead
                         // of the single issuing thread, which otherwise exceeds the MMA time of a 128 x 64 x 64 block)
  float* stats;          // optional (fprop modes): per-CTA column sums of the stored bf16 outputs,
                         // [gridDim.x][2][Cout] (sum, sum of squares) for the BatchNorm that follows
};

constexpr int kNumSMs = 148;
constexpr int kATileBytes = 128 * 128;  // 128 rows x 64 bf16 (K-major) or 2 x (64 px x 64 ch) (MN-major)
constexpr int kStatSmemBytes = 2 * 4 * 2 * 256 * 4;  // [tile parity][epilogue warp][sum|sumsq][column]
constexpr int kStatAccBytes = 2 * 512 * 4;           // per-CTA running [sum|sumsq][Cout <= 512]

struct TileCoord {
  int w0, h0, b0, n0, m0, phase, kb_begin, kb_count, row;
};

// Work item t of the persistent loop.  fprop/dgrad: the pixel tile is the slow index so the CTAs running at the same
// time read the same activations (4 phases x n_tiles) out of L2; wgrad: the pixel range (split) is the slow index.
__device__ __forceinline__ TileCoord decode_tile(const IGemmParams& p, int t) {
  TileCoord c;
  c.w0 = c.h0 = c.b0 = c.m0 = c.phase = 0;
  c.kb_begin = 0;
  c.kb_count = p.kb_per_tile;
  if (p.mode == MODE_WGRAD) {
    const int per_split = p.m_tiles * p.n_tiles;
    const int split = t / per_split, r = t - split * per_split;
    const int mt = r / p.n_tiles, nt = r - mt * p.n_tiles;
    c.m0 = mt * 128;
    c.n0 = nt * p.BN;
    c.kb_begin = split * p.kb_per_tile;
    c.kb_count = min(p.kb_per_tile, p.kb_total - c.kb_begin);
    c.row = split;
  } else {
    const int per_pt = p.phases * p.n_tiles;
    const int pt = t / per_pt, r = t - pt * per_pt;
    c.phase = r / p.n_tiles;
    c.n0 = (r - c.phase * p.n_tiles) * p.BN;
    const int wt = pt & (p.tiles_w - 1);
    const int ht = (pt >> p.tiles_w_sh) & (p.tiles_h - 1);
    const int bt = pt >> (p.tiles_w_sh + p.tiles_h_sh);
    c.w0 = wt << p.tw_sh;
    c.h0 = ht << p.th_sh;
    c.b0 = bt * p.nb;
    c.row = pt * p.phases + c.phase;
  }
  return c;
}

// lane j ends up with sum over the 32 lanes of v[j] (31 shuffles).
__device__ __forceinline__ float transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int k = 0; k < h; ++k) {
      const float send = up ? v[k] : v[k + h];
      const float keep = up ? v[k + h] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  return v[0];
}

__global__ void __launch_bounds__(192, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const IGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int b_bytes = p.BN * 128;
  const int stage_bytes = p.kpb * (kATileBytes + b_bytes);
  uint64_t* full_bar = (uint64_t*)(smem + p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]
  uint32_t* tmem_ptr_smem = (uint32_t*)(tmem_empty_bar + 2);
  float* stat_smem = (float*)(tmem_ptr_smem + 4);   // kStatSmemBytes + kStatAccBytes, only touched when p.stats
  float* stat_acc = stat_smem + kStatSmemBytes / 4;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 4);  // one arrival per epilogue warp
    }
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
      int s = 0;
      uint32_t par = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const TileCoord c = decode_tile(p, t);
        const int ph_p = c.phase >> 1, ph_q = c.phase & 1;
        for (int i = 0; i < c.kb_count; i += p.kpb) {
          mbar_wait(&empty_bar[s], par ^ 1u);
          mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
          for (int u = 0; u < p.kpb; ++u) {
          uint8_t* sA = smem + s * stage_bytes + u * kATileBytes;
          uint8_t* sB = smem + s * stage_bytes + p.kpb * kATileBytes + u * b_bytes;
          const int kb = c.kb_begin + i + u;
          if (p.mode == MODE_CONV_S2) {
            const int tap = kb >> p.cpt_sh, cc = kb & (p.cpt - 1);
            const int kh = tap >> 2, kw = tap & 3;
            tma_load_5d(sA, &tmA, &full_bar[s], ((kw + 1) & 1) * p.C_k + cc * 64, c.w0 + ((kw + 1) >> 1) - 1, (kh + 1) & 1,
                        c.h0 + ((kh + 1) >> 1) - 1, c.b0);
            tma_load_2d(sB, &tmB, &full_bar[s], kb * 64, c.n0);
          } else if (p.mode == MODE_CONVT) {
            const int tap = kb >> p.cpt_sh, cc = kb & (p.cpt - 1);
            const int ta = tap >> 1, tb = tap & 1;
            tma_load_4d(sA, &tmA, &full_bar[s], cc * 64, c.w0 + ph_q + tb - 1, c.h0 + ph_p + ta - 1, c.b0);
            tma_load_2d(sB, &tmB, &full_bar[s], kb * 64, c.phase * p.Cout + c.n0);
          } else {
            const int wt = kb & (p.tiles_w - 1);
            const int ht = (kb >> p.tiles_w_sh) & (p.tiles_h - 1);
            const int bt = kb >> (p.tiles_w_sh + p.tiles_h_sh);
            const int pw0 = wt << p.tw_sh, ph0 = ht << p.th_sh, pb0 = bt * p.nb;
            tma_load_4d(sA, &tmA, &full_bar[s], c.m0, pw0, ph0, pb0);
            tma_load_4d(sA + 8192, &tmA, &full_bar[s], c.m0 + 64, pw0, ph0, pb0);
            const int nsub = p.BN >> 6;
            for (int j = 0; j < nsub; ++j) {
              const int nidx = c.n0 + 64 * j;
              const int tap = nidx >> p.ck_sh, c0 = nidx & (p.C_k - 1);
              const int kh = tap >> 2, kw = tap & 3;
              tma_load_5d(sB + j * 8192, &tmB, &full_bar[s], ((kw + 1) & 1) * p.C_k + c0, pw0 + ((kw + 1) >> 1) - 1, (kh + 1) & 1,
                          ph0 + ((kh + 1) >> 1) - 1, pb0);
            }
          }
          }
          if (++s == p.stages) {
            s = 0;
            par ^= 1u;
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
      int s = 0, lt = 0;
      uint32_t par = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++lt) {
        const TileCoord c = decode_tile(p, t);
        const int buf = lt & 1;
        mbar_wait(&tmem_empty_bar[buf], (uint32_t)(((lt >> 1) & 1) ^ 1));  // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_addr = tmem_base + (uint32_t)(buf * p.BN);
        for (int i = 0; i < c.kb_count; i += p.kpb) {
          mbar_wait(&full_bar[s], par);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
          const uint32_t b_addr = a_addr + p.kpb * kATileBytes;
          for (int u = 0; u < p.kpb; ++u) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = make_smem_desc_sw128(a_addr + u * kATileBytes + k * kstep, lbo, 1024u);
              const uint64_t db = make_smem_desc_sw128(b_addr + u * b_bytes + k * kstep, lbo, 1024u);
              umma_bf16(d_addr, da, db, idesc, (i > 0 || u > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs retire
          if (++s == p.stages) {
            s = 0;
            par ^= 1u;
          }
        }
        umma_commit(&tmem_full_bar[buf]);  // accumulator complete
      }
    }
  } else {
    // ===================== epilogue (overlaps the next tile's main loop) =====================
    const int quad = warp & 3;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int r = quad * 32 + lane;
    const int et = threadIdx.x - 64;  // 0..127
    if (p.stats) {
      for (int e = et; e < 2 * p.Cout; e += 128) stat_acc[e] = 0.f;
    }
    int lt = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++lt) {
      const TileCoord c = decode_tile(p, t);
      const int buf = lt & 1;
      mbar_wait(&tmem_full_bar[buf], (uint32_t)((lt >> 1) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * p.BN);
      if (p.mode == MODE_WGRAD) {
        float* dst = (float*)p.out + ((size_t)c.row * p.m_tiles * 128 + (size_t)(c.m0 + r)) * p.ldc + c.n0;
        for (int cc = 0; cc < p.BN; cc += 32) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + cc, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(dst + cc + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      } else {
        const int bl = r >> (p.th_sh + p.tw_sh), rem = r & ((1 << (p.th_sh + p.tw_sh)) - 1);
        const int hl = rem >> p.tw_sh, wl = rem & (p.tw - 1);
        const int b = c.b0 + bl;
        size_t pix;
        if (p.mode == MODE_CONV_S2) {
          pix = ((size_t)b * p.Hs + (c.h0 + hl)) * p.Ws + (c.w0 + wl);
        } else {
          pix = ((size_t)b * (2 * p.Hs) + (2 * (c.h0 + hl) + (c.phase >> 1))) * (size_t)(2 * p.Ws) + (2 * (c.w0 + wl) + (c.phase & 1));
        }
        __nv_bfloat16* dst = (__nv_bfloat16*)p.out + pix * p.ldc + c.n0;
        const bool valid = (b < p.B);
        float* st = stat_smem + ((lt & 1) * 4 + quad) * 512;
        for (int cc = 0; cc < p.BN; cc += 32) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + cc, v);  // .sync.aligned: all lanes participate even if the row is masked
          tmem_ld_wait();
          uint32_t o[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
          if (valid) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) *reinterpret_cast<uint4*>(dst + cc + 2 * j) = make_uint4(o[j], o[j + 1], o[j + 2], o[j + 3]);
          }
          if (p.stats) {
            float s[32], q[32];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 f = unpack_bf16x2(o[j]);  // statistics of the values as stored
              s[2 * j] = valid ? f.x : 0.f;
              s[2 * j + 1] = valid ? f.y : 0.f;
              q[2 * j] = s[2 * j] * s[2 * j];
              q[2 * j + 1] = s[2 * j + 1] * s[2 * j + 1];
            }
            const float cs = transpose_reduce32(s, lane);
            const float cq = transpose_reduce32(q, lane);
            st[cc + lane] = cs;
            st[256 + cc + lane] = cq;
          }
        }
        if (p.stats) {
          asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps
          const float* sb = stat_smem + (lt & 1) * 4 * 512;
          for (int e = et; e < 2 * p.BN; e += 128) {  // element (k, n0 + col) is always owned by the same thread
            const int k = e >= p.BN, col = e - k * p.BN;
            const float a = (sb[0 * 512 + k * 256 + col] + sb[1 * 512 + k * 256 + col]) + (sb[2 * 512 + k * 256 + col] + sb[3 * 512 + k * 256 + col]);
            stat_acc[k * p.Cout + c.n0 + col] += a;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
    }
    if (p.stats) {  // one deterministic partial row per CTA (static tile schedule)
      float* gdst = p.stats + (size_t)blockIdx.x * 2 * p.Cout;
      for (int e = et; e < 2 * p.Cout; e += 128) gdst[e] = stat_acc[e];
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

static int ilog2(int v) {
  int s = 0;
  while ((1 << s) < v) ++s;
  return s;
}

static int launch(const CUtensorMap& a, const CUtensorMap& b, IGemmParams& p, cudaStream_t stream, int* grid_out) {
  if (!is_pow2(p.tw) || !is_pow2(p.th) || !is_pow2(p.tiles_w) || !is_pow2(p.tiles_h) || !is_pow2(p.C_k) || !is_pow2(p.cpt)) {
    b200_set_error("igemm: image sides and channel counts must be powers of two");
    return B200_ERR_ARG;
  }
  p.tw_sh = ilog2(p.tw);
  p.th_sh = ilog2(p.th);
  p.tiles_w_sh = ilog2(p.tiles_w);
  p.tiles_h_sh = ilog2(p.tiles_h);
  p.ck_sh = ilog2(p.C_k);
  p.cpt_sh = ilog2(p.cpt);
  p.kpb = (p.mode != MODE_WGRAD && p.BN <= 128 && p.kb_per_tile % 2 == 0) ? 2 : 1;
  const int stage_bytes = p.kpb * (kATileBytes + p.BN * 128);
  const int extra = 1024 + 256 + kStatSmemBytes + kStatAccBytes;
  if (p.stages <= 0) {
    p.stages = (227 * 1024 - extra) / stage_bytes;
    if (p.stages > 8) p.stages = 8;
  }
  if (p.stages < 2 || (size_t)p.stages * stage_bytes + extra > 227 * 1024) {
    b200_set_error("igemm: pipeline does not fit shared memory");
    return B200_ERR_ARG;
  }
  uint32_t cols = 32;
  while ((int)cols < 2 * p.BN) cols <<= 1;
  p.tmem_cols = cols;
  const size_t smem = (size_t)p.stages * stage_bytes + extra;
  static std::once_flag once;
  std::call_once(once, [] { cudaFuncSetAttribute(igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
  const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  if (grid_out) *grid_out = grid;
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

int b200_igemm_stat_rows(void) { return kNumSMs; }

// y[B,H/2,W/2,ldy] (bf16) = conv4x4/s2/p1(x[B,H,W,Cin] bf16, wpk[Cout][16*Cin] bf16 with k = (kh*4+kw)*Cin + ci)
// stats (optional): [b200_igemm_stat_rows()][2][Cout] fp32 per-CTA column sums / sums of squares of y; *stat_rows = rows written.
int b200_conv_s2_gemm(const void* x, int B, int H, int W, int Cin, const void* wpk, int Cout, void* y, int ldy, int bn_hint, int stages, float* stats,
                      int* stat_rows, cudaStream_t stream) {
  if (!is_pow2(H) || !is_pow2(W) || H < 2 || W < 2 || Cin % 64 || Cout % 32 || ldy % 8 || B < 1) {
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
  p.kb_per_tile = p.kb_total = 16 * p.cpt;
  p.BN = pick_bn(Cout, bn_hint);
  if (Cout % p.BN || p.BN % 32 || p.BN > 256 || (stats && Cout > 512)) {
    b200_set_error("conv_s2_gemm: bad N tile");
    return B200_ERR_ARG;
  }
  p.Cout = Cout;
  p.ldc = ldy;
  p.out = y;
  p.stages = stages;
  p.n_tiles = Cout / p.BN;
  p.phases = 1;
  p.m_tiles = 1;
  p.num_tiles = p.num_ptiles * p.n_tiles;
  p.stats = stats;
  CUtensorMap ta, tb;
  int rc;
  if ((rc = map_nhwc_parity(&ta, x, B, H, W, Cin, p.tw, p.th, p.nb))) return rc;
  if ((rc = map_weights(&tb, wpk, Cout, 16 * Cin, p.BN))) return rc;
  return launch(ta, tb, p, stream, stat_rows);
}

// y[B,2Hi,2Wi,ldy] (bf16) = convT4x4/s2/p1(z[B,Hi,Wi,Cin] (pixel stride ldz), wpk[4 phases][Cout][4*Cin]); stats as above.
int b200_convt_s2_gemm(const void* z, int B, int Hi, int Wi, int Cin, int ldz, const void* wpk, int Cout, void* y, int ldy, int bn_hint, int stages,
                       float* stats, int* stat_rows, cudaStream_t stream) {
  if (!is_pow2(Hi) || !is_pow2(Wi) || Cin % 64 || Cout % 32 || ldy % 8 || ldz % 8 || B < 1) {
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
  p.kb_per_tile = p.kb_total = 4 * p.cpt;
  p.BN = pick_bn(Cout, bn_hint);
  if (Cout % p.BN || p.BN % 32 || p.BN > 256 || (stats && Cout > 512)) {
    b200_set_error("convt_s2_gemm: bad N tile");
    return B200_ERR_ARG;
  }
  p.Cout = Cout;
  p.ldc = ldy;
  p.out = y;
  p.stages = stages;
  p.n_tiles = Cout / p.BN;
  p.phases = 4;
  p.m_tiles = 1;
  p.num_tiles = p.num_ptiles * p.n_tiles * 4;
  p.stats = stats;
  CUtensorMap ta, tb;
  int rc;
  if ((rc = map_nhwc(&ta, z, B, Hi, Wi, Cin, ldz, p.tw, p.th, p.nb))) return rc;
  if ((rc = map_weights(&tb, wpk, 4 * Cout, 4 * Cin, p.BN))) return rc;
  return launch(ta, tb, p, stream, stat_rows);
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
  if ((16 * Cl) % p.BN || p.BN % 64 || p.BN > 256) {
    b200_set_error("wgrad_gemm: bad N tile");
    return B200_ERR_ARG;
  }
  int splits = *splits_io;
  if (splits < 1) splits = 1;
  if (splits > p.num_ptiles) splits = p.num_ptiles;
  p.kb_total = p.num_ptiles;
  p.kb_per_tile = (p.num_ptiles + splits - 1) / splits;
  splits = (p.num_ptiles + p.kb_per_tile - 1) / p.kb_per_tile;  // every split owns at least one pixel tile
  *splits_io = splits;
  p.Cout = Cs;
  p.ldc = 16 * Cl;
  p.out = partial;
  p.stages = stages;
  p.m_tiles = Cs / 128;
  p.n_tiles = 16 * Cl / p.BN;
  p.phases = 1;
  p.num_tiles = p.m_tiles * p.n_tiles * splits;
  p.stats = nullptr;
  CUtensorMap ta, tb;
  int rc;
  if ((rc = map_nhwc(&ta, S, B, Hs, Ws, Cs, lds, p.tw, p.th, p.nb))) return rc;
  if ((rc = map_nhwc_parity(&tb, L, B, 2 * Hs, 2 * Ws, Cl, p.tw, p.th, p.nb))) return rc;
  return launch(ta, tb, p, stream, nullptr);
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
