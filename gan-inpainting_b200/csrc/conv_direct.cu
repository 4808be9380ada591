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

#include <mutex>

namespace b200 {

enum { ACT_NONE = 0, ACT_LEAKY = 1, ACT_RELU = 2, ACT_TANH = 3 };

// ------------------------------------------------------------------------------------------------
// conv_c1_fwd: y[pix][c] = sum_t patch[pix][t] * w[c][t]  (t = kh*4+kw, K = 16) on the warp-level bf16 MMA
// (m16n8k16), fp32 operands carried as bf16 hi+lo pairs (3 MMAs: hi*hi + lo*hi + hi*lo), fp32 accumulation.
// Persistent blocks walk 256-pixel output tiles (TOH x TOW); the (2*TOH+2) x (2*TOW+2) input window (and the window of the
// optional multiplier, the local discriminator's mask) of the NEXT tile is fetched with zero-filling cp.async while the
// current one is multiplied.  Warp = two 16-pixel row segments; the n-columns of the MMA are permuted so that a lane ends up
// with consecutive channels of a pixel and stores them with 128-bit stores (a quad covers 64 contiguous bytes).  The weight
// fragments live in shared memory.  The kernel is bound by the bf16 output stream (2*C bytes per pixel).
// ------------------------------------------------------------------------------------------------
constexpr int kXsFloats = 34 * 68;  // covers 18 x 68 (8 x 32 tile) and 34 x 36 (16 x 16 tile)

template <int C>
__global__ void __launch_bounds__(256, 3) conv_c1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ xmul, int B, int H, int W, int tow_shift,
                                                             const float* __restrict__ w, int act, __nv_bfloat16* __restrict__ y,
                                                             __nv_bfloat16* __restrict__ y2, int ld2, int off2) {
  b200::pdl_enter();
  constexpr int NT = C / 8;   // n-tiles
  constexpr int CPL = C / 4;  // consecutive channels per lane
  __shared__ __align__(16) float xs[2][kXsFloats];
  __shared__ __align__(16) float xm[2][kXsFloats];
  __shared__ __align__(16) uint32_t bsm[NT * 32 * 4];  // [n-tile][lane][hi b0, hi b1, lo b0, lo b1]: one 128-bit load per n-tile
  const int TOW = 1 << tow_shift, TOH = 256 >> tow_shift;
  const int Ho = H >> 1, Wo = W >> 1;
  const int tiles_w = Wo >> tow_shift, tiles_h = Ho / TOH;
  const int ntiles = B * tiles_w * tiles_h;
  const int rows = 2 * TOH + 2, cols = 2 * TOW + 2, xstride = 2 * TOW + 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;

  auto stage = [&](int tile, int buf) {
    const int tw_i = tile % tiles_w;
    const int th_i = (tile / tiles_w) % tiles_h;
    const int b = tile / (tiles_w * tiles_h);
    const int ih0 = 2 * (th_i * TOH) - 1, iw0 = 2 * (tw_i << tow_shift) - 1;
    const float* xb = x + (size_t)b * H * W;
    const float* mb = xmul ? xmul + (size_t)b * H * W : nullptr;
    // element e = r * cols + cI walked in steps of 256 without a division per step (cols is 66 or 34: 256 = dq * cols + dr)
    const int dq = 256 / cols, dr = 256 - dq * cols;
    int r = (int)threadIdx.x / cols, cI = (int)threadIdx.x - r * cols;
    for (int e = threadIdx.x; e < rows * cols; e += 256) {
      const int ih = ih0 + r, iw = iw0 + cI;
      const bool ok = ih >= 0 && ih < H && iw >= 0 && iw < W;
      const size_t off = ok ? (size_t)ih * W + iw : 0;
      cp_async4_zfill(&xs[buf][r * xstride + cI], xb + off, ok);
      if (mb) cp_async4_zfill(&xm[buf][r * xstride + cI], mb + off, ok);
      r += dq, cI += dr;
      if (cI >= cols) cI -= cols, ++r;
    }
    cp_async_commit();
  };
  if ((int)blockIdx.x < ntiles) stage(blockIdx.x, 0);

  // B fragments -> shared memory: MMA column g of n-tile j carries channel (g>>1)*CPL + 2j + (g&1)
  for (int e = threadIdx.x; e < NT * 32; e += 256) {
    const int j = e >> 5, l = e & 31;
    const int gg = l >> 2, qq = l & 3;
    const float* wc = w + (size_t)((gg >> 1) * CPL + 2 * j + (gg & 1)) * 16 + 2 * qq;
    const float2 w0 = __ldg(reinterpret_cast<const float2*>(wc));
    const float2 w1 = __ldg(reinterpret_cast<const float2*>(wc + 8));
    uint32_t h0, l0, h1, l1;
    split_bf16x2(w0.x, w0.y, h0, l0);
    split_bf16x2(w1.x, w1.y, h1, l1);
    *reinterpret_cast<uint4*>(bsm + (j * 32 + l) * 4) = make_uint4(h0, h1, l0, l1);
  }

  const int kh = q >> 1, kw = (q & 1) * 2;  // taps 2q, 2q+1 (and +8: kh + 2)
  const bool wide2 = y2 && !(ld2 & 15) && !(off2 & 15) && !((uintptr_t)y2 & 31);  // 32-byte aligned rows in the concat buffer
  int buf = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
    const int next = tile + gridDim.x;
    if (next < ntiles) {
      stage(next, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int tw_i = tile % tiles_w;
    const int th_i = (tile / tiles_w) % tiles_h;
    const int b = tile / (tiles_w * tiles_h);
    const int oh0 = th_i * TOH, ow0 = tw_i << tow_shift;
    const float* xw = xs[buf];
    const float* mw = xm[buf];
#pragma unroll 1
    for (int m = 0; m < 2; ++m) {
      const int mt = warp * 2 + m;  // 16 m-tiles of 16 pixels per block
      const int ol_h = (mt * 16) >> tow_shift, ol_w0 = (mt * 16) & (TOW - 1);
      uint32_t ahi[4], alo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int pw = ol_w0 + g + (i & 1) * 8, kk = kh + (i >> 1) * 2;
        const int o = (2 * ol_h + kk) * xstride + 2 * pw + kw;
        float2 v = *reinterpret_cast<const float2*>(xw + o);
        if (xmul) {
          const float2 mm = *reinterpret_cast<const float2*>(mw + o);
          v.x *= mm.x;
          v.y *= mm.y;
        }
        split_bf16x2(v.x, v.y, ahi[i], alo[i]);
      }
      const size_t pix0 = ((size_t)b * Ho + (oh0 + ol_h)) * Wo + (ow0 + ol_w0 + g);
#pragma unroll
      for (int hn = 0; hn < NT; hn += 8) {  // 8 n-tiles (16 consecutive channels per lane) at a time: 32 accumulators
        float acc[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 bq = *reinterpret_cast<const uint4*>(bsm + ((hn + j) * 32 + lane) * 4);
          const uint32_t h0 = bq.x, h1 = bq.y, l0 = bq.z, l1 = bq.w;
          acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
          mma_bf16_16816(acc[j], alo, h0, h1);
          mma_bf16_16816(acc[j], ahi, l0, l1);
          mma_bf16_16816(acc[j], ahi, h0, h1);
        }
        if (act == ACT_LEAKY) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[j][i] = fmaxf(acc[j][i], 0.2f * acc[j][i]);  // == x > 0 ? x : 0.2 x, two instructions
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {  // pixel g (d0,d1) and pixel g+8 (d2,d3)
          const size_t pix = pix0 + half * 8;
          uint32_t o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = pack_bf16x2(acc[j][2 * half], acc[j][2 * half + 1]);
          // a lane's 16 consecutive channels are one 32-byte sector: one 256-bit store (sm_100 STG.256) instead of two half-sector ones
          st_global_256(y + pix * C + q * CPL + 2 * hn, o);
          if (y2) {  // relu copy into the decoder's concat buffer (skip connection)
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = pack_bf16x2(fmaxf(acc[j][2 * half], 0.f), fmaxf(acc[j][2 * half + 1], 0.f));
            __nv_bfloat16* d2 = y2 + pix * ld2 + off2 + q * CPL + 2 * hn;
            if (wide2) {
              st_global_256(d2, o);
            } else {
              uint4* dst2 = reinterpret_cast<uint4*>(d2);
              dst2[0] = make_uint4(o[0], o[1], o[2], o[3]);
              dst2[1] = make_uint4(o[4], o[5], o[6], o[7]);
            }
          }
        }
      }
    }
    __syncthreads();  // everyone is done with `buf` before the next iteration's prefetch overwrites it
  }
}

// ------------------------------------------------------------------------------------------------
// convt_c1_fwd: block = 16x16 input pixels (+1 halo).
// Phase 1: T[pix][t] = sum_c z[pix][c] * w[c][t] (16 taps) for the 18 x 18 halo window on the warp-level bf16 MMA
//   (m16n8k16; z is bf16, the fp32 weights are carried as bf16 hi+lo pairs).  The MMA's k index is permuted so that a
//   lane loads 8 CONSECUTIVE channels of its pixel with one 128-bit load (a quad covers 64 contiguous bytes): for the
//   32-channel group gi and k-step s, logical k = 2q+e <-> channel gi*32 + q*8 + 4s + e, k = 2q+8+e <-> ... + 2 + e.
// Phase 2: every output pixel sums its 4 contributing (pixel, tap) entries
//   y[2i+p, 2j+q] = sum_{a,b} T[i+p+a-1, j+q+b-1][(3-p-2a)*4 + (3-q-2b)]   (+ bias, tanh, composite / mask multiply)
// ------------------------------------------------------------------------------------------------
constexpr int kTT = 16;             // tile side (input pixels)
constexpr int kTH = kTT + 2;        // with halo

template <int C>
__global__ void __launch_bounds__(256, 2) convt_c1_fwd_kernel(const __nv_bfloat16* __restrict__ z, int B, int H, int W, int ldz, const float* __restrict__ w,
                                                           const float* __restrict__ bias, int act, const float* __restrict__ base,
                                                           const float* __restrict__ mul, float* __restrict__ y, float* __restrict__ comp) {
  b200::pdl_enter();
  constexpr int NG = C / 32;  // 32-channel groups
  __shared__ float T[kTH * kTH][17];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int tiles_w = W / kTT, tiles_h = H / kTT;
  const int tw_i = blockIdx.x % tiles_w;
  const int th_i = (blockIdx.x / tiles_w) % tiles_h;
  const int b = blockIdx.x / (tiles_w * tiles_h);
  const int i0 = th_i * kTT - 1, j0 = tw_i * kTT - 1;  // image coordinates of the halo origin

  // B fragments [group][k-step][n-tile] in shared memory (per lane): b0 = {w[ch][tap], w[ch+1][tap]}, b1 = {w[ch+2][tap], w[ch+3][tap]},
  // tap = nt*8 + g; entry ((gi*2 + s)*2 + nt)*4 + {hi b0, hi b1, lo b0, lo b1}
  __shared__ uint32_t bsm[NG * 16 * 32];
  constexpr int npix = kTH * kTH;
  constexpr int nmt = (npix + 15) / 16;
  auto load_rows = [&](int mt, uint4 (&qa)[NG], uint4 (&qb)[NG]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int pp = mt * 16 + g + h * 8;
      const int li = pp / kTH, lj = pp - li * kTH;
      const int i = i0 + li, j = j0 + lj;
      const bool ok = (pp < npix) && i >= 0 && i < H && j >= 0 && j < W;
      const __nv_bfloat16* src = z + (((size_t)b * H + i) * W + j) * ldz + q * 8;
#pragma unroll
      for (int gi = 0; gi < NG; ++gi) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (ok) v = __ldg(reinterpret_cast<const uint4*>(src + gi * 32));
        if (h == 0) qa[gi] = v; else qb[gi] = v;
      }
    }
  };
  uint4 qa[NG], qb[NG];
  if (warp < nmt) load_rows(warp, qa, qb);  // the first activation rows travel while the filter fragments are built
  for (int e = threadIdx.x; e < NG * 4 * 32; e += 256) {
    const int l = e & 31, grp = e >> 5;  // grp = (gi*2 + s)*2 + nt
    const int nt = grp & 1, sk = (grp >> 1) & 1, gi = grp >> 2;
    const int gg = l >> 2, qq = l & 3;
    const float* wc = w + (size_t)(gi * 32 + qq * 8 + 4 * sk) * 16 + nt * 8 + gg;
    uint32_t h0, l0, h1, l1;
    split_bf16x2(__ldg(wc), __ldg(wc + 16), h0, l0);
    split_bf16x2(__ldg(wc + 32), __ldg(wc + 48), h1, l1);
    bsm[(grp * 4 + 0) * 32 + l] = h0;
    bsm[(grp * 4 + 1) * 32 + l] = h1;
    bsm[(grp * 4 + 2) * 32 + l] = l0;
    bsm[(grp * 4 + 3) * 32 + l] = l1;
  }
  __syncthreads();

  for (int mt = warp; mt < nmt; mt += 8) {
    uint4 na[NG], nb[NG];
    if (mt + 8 < nmt) load_rows(mt + 8, na, nb);  // next tile's rows are in flight while this one is multiplied
    float acc[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) {
      const uint32_t pa[4] = {qa[gi].x, qa[gi].y, qa[gi].z, qa[gi].w};
      const uint32_t pb[4] = {qb[gi].x, qb[gi].y, qb[gi].z, qb[gi].w};
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const uint32_t a[4] = {pa[2 * s], pb[2 * s], pa[2 * s + 1], pb[2 * s + 1]};
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const uint32_t* bj = bsm + (((gi * 2 + s) * 2 + nt) * 4) * 32 + lane;
          mma_bf16_16816(acc[nt], a, bj[64], bj[96]);
          mma_bf16_16816(acc[nt], a, bj[0], bj[32]);
        }
      }
    }
    const int pp0 = mt * 16 + g;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      if (pp0 < npix) {
        T[pp0][nt * 8 + 2 * q] = acc[nt][0];
        T[pp0][nt * 8 + 2 * q + 1] = acc[nt][1];
      }
      if (pp0 + 8 < npix) {
        T[pp0 + 8][nt * 8 + 2 * q] = acc[nt][2];
        T[pp0 + 8][nt * 8 + 2 * q + 1] = acc[nt][3];
      }
    }
    if (mt + 8 < nmt) {
#pragma unroll
      for (int gi = 0; gi < NG; ++gi) {
        qa[gi] = na[gi];
        qb[gi] = nb[gi];
      }
    }
  }
  __syncthreads();

  const float bv = bias ? __ldg(bias) : 0.f;
  const int Ho = 2 * H, Wo = 2 * W;
  for (int o = threadIdx.x; o < 4 * kTT * kTT; o += 256) {
    const int lo_h = o / (2 * kTT), lo_w = o - lo_h * (2 * kTT);
    const int p = lo_h & 1, qq = lo_w & 1;
    const int li = (lo_h >> 1) + 1, lj = (lo_w >> 1) + 1;  // halo-local input coordinates of (i, j)
    float s = bv;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int bb = 0; bb < 2; ++bb) s += T[(li + p + a - 1) * kTH + (lj + qq + bb - 1)][(3 - p - 2 * a) * 4 + (3 - qq - 2 * bb)];
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
// wgrad_c1: dW[c][t] = sum_{b,i,j} z[b,i,j,c] * g[b,2i+kh-1,2j+kw-1],  t = kh*4+kw
// as D[t][c] = sum_pix P[t][pix] * z[pix][c] on the warp-level bf16 MMA (m16n8k16): A = the im2col'd fp32 image g as
// bf16 hi+lo pairs (built from a shared-memory window, mask multiply applied once), B = z read with ldmatrix.trans
// from a cp.async-staged shared-memory tile (so the HBM stream of z is fully asynchronous, double buffered).
// Persistent blocks walk 4 x 32 pixel tiles, one 16-pixel k-step per warp and tile; accumulators stay in registers
// across tiles; deterministic per-block partials.
// ------------------------------------------------------------------------------------------------
constexpr int kWgTH = 4, kWgTW = 32, kWgStride = 2 * kWgTW + 4, kWgRows = 2 * kWgTH + 2;

// S: pipeline stages (S - 1 tiles in flight per block).  C = 64: three stages, three blocks per SM (96 KB of z in flight per SM; two
// stages x two blocks = 32 KB was below the bandwidth-latency product of HBM3e: 2.9 TB/s); C = 128: two stages, two blocks.
template <int C, int S>
__global__ void __launch_bounds__(256) wgrad_c1_kernel(const __nv_bfloat16* __restrict__ z, int B, int H, int W, int ldz, const float* __restrict__ g,
                                                       const float* __restrict__ gmul, float* __restrict__ partial) {
  b200::pdl_enter();
  constexpr int NT = C / 8;
  constexpr int ROWB = (C + 8) * 2;              // padded shared-memory row (bytes): conflict-free ldmatrix
  constexpr int ZBYTES = kWgTH * kWgTW * ROWB;   // one stage of z
  constexpr int GFLOATS = kWgRows * kWgStride;
  extern __shared__ __align__(16) uint8_t wsm[];
  uint8_t* zs = wsm;                                   // [S][128][ROWB]
  float* gs = reinterpret_cast<float*>(wsm + S * ZBYTES);  // [S][GFLOATS] image window, then [S][GFLOATS] mask window
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gq = lane >> 2, q = lane & 3;
  const int Hg = 2 * H, Wg = 2 * W;
  const int tiles_w = W / kWgTW, tiles_h = H / kWgTH;
  const int ntiles = B * tiles_w * tiles_h;

  auto stage_tile = [&](int tile, int buf) {
    const int tw_i = tile % tiles_w;
    const int th_i = (tile / tiles_w) % tiles_h;
    const int b = tile / (tiles_w * tiles_h);
    const int i0 = th_i * kWgTH, j0 = tw_i * kWgTW;
    // z: 128 pixels x C channels, 16-byte chunks
    constexpr int CH = C / 8;
    uint8_t* zd = zs + buf * ZBYTES;
    for (int e = threadIdx.x; e < kWgTH * kWgTW * CH; e += 256) {
      const int pp = e / CH, ch = e - pp * CH;
      const int li = pp >> 5, lj = pp & 31;
      cp_async16(zd + pp * ROWB + ch * 16, z + (((size_t)b * H + i0 + li) * W + j0 + lj) * ldz + ch * 8);
    }
    // g (and mask) window: rows 2*i0-1 .. 2*i0+2*TH, cols 2*j0-1 .. 2*j0+2*TW, zero-filled halo
    const float* gb = g + (size_t)b * Hg * Wg;
    const float* mb = gmul ? gmul + (size_t)b * Hg * Wg : nullptr;
    float* gd = gs + buf * GFLOATS;
    float* md = gs + (S + buf) * GFLOATS;
    for (int e = threadIdx.x; e < kWgRows * (2 * kWgTW + 2); e += 256) {
      const int r = e / (2 * kWgTW + 2), cI = e - r * (2 * kWgTW + 2);
      const int ih = 2 * i0 - 1 + r, iw = 2 * j0 - 1 + cI;
      const bool ok = ih >= 0 && ih < Hg && iw >= 0 && iw < Wg;
      const size_t off = ok ? (size_t)ih * Wg + iw : 0;
      cp_async4_zfill(gd + r * kWgStride + cI, gb + off, ok);
      if (mb) cp_async4_zfill(md + r * kWgStride + cI, mb + off, ok);
    }
    cp_async_commit();
  };

  float acc[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;

  // one cp.async group per tile slot, committed even when the slot is empty, so that "all but the newest S - 1 groups" is always
  // exactly "this iteration's tile has landed"
  int buf = 0;
#pragma unroll
  for (int i = 0; i < S - 1; ++i) {
    const int t = blockIdx.x + i * gridDim.x;
    if (t < ntiles) stage_tile(t, i);
    else cp_async_commit();
  }
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf = (buf + 1 == S ? 0 : buf + 1)) {
    const int next = tile + (S - 1) * gridDim.x;
    const int nbuf = buf == 0 ? S - 1 : buf - 1;  // (buf + S - 1) % S: released by the barrier that ended the previous iteration
    if (next < ntiles) stage_tile(next, nbuf);
    else cp_async_commit();
    cp_async_wait<S - 1>();
    __syncthreads();
    // this warp's k-step: pixels (row = warp / 2, cols (warp & 1) * 16 + 0..15)
    const int r = warp >> 1, c0 = (warp & 1) * 16;
    const float* gd = gs + buf * GFLOATS;
    const int moff = S * GFLOATS;  // mask window of the same buffer
    uint32_t ahi[4], alo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = gq + (i & 1) * 8, kp = 2 * q + (i >> 1) * 8;  // tap, first of the two pixels
      const float* src = gd + (2 * r + (t >> 2)) * kWgStride + 2 * (c0 + kp) + (t & 3);
      float v0 = src[0], v1 = src[2];
      if (gmul) {
        v0 *= src[moff];
        v1 *= src[moff + 2];
      }
      split_bf16x2(v0, v1, ahi[i], alo[i]);
    }
    const uint8_t* zrow = zs + buf * ZBYTES + (size_t)((warp * 16) + (lane & 7) + ((lane >> 3) & 1) * 8) * ROWB + ((lane >> 4) * 8) * 2;
#pragma unroll
    for (int j = 0; j < NT; j += 2) {
      uint32_t bf[4];
      ldmatrix_x4_trans(bf, zrow + j * 16);
      mma_bf16_16816(acc[j], alo, bf[0], bf[1]);
      mma_bf16_16816(acc[j], ahi, bf[0], bf[1]);
      mma_bf16_16816(acc[j + 1], alo, bf[2], bf[3]);
      mma_bf16_16816(acc[j + 1], ahi, bf[2], bf[3]);
    }
    __syncthreads();  // everyone is done with `buf` before the next iteration's prefetch overwrites it
  }

  // cross-warp reduction in shared memory (reusing the staging buffers): red[warp][t][c]
  cp_async_wait<0>();
  float* red = reinterpret_cast<float*>(wsm);
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    float* rw = red + (size_t)warp * 16 * C;
    rw[gq * C + j * 8 + 2 * q] = acc[j][0];
    rw[gq * C + j * 8 + 2 * q + 1] = acc[j][1];
    rw[(gq + 8) * C + j * 8 + 2 * q] = acc[j][2];
    rw[(gq + 8) * C + j * 8 + 2 * q + 1] = acc[j][3];
  }
  __syncthreads();
  for (int k = threadIdx.x; k < C * 16; k += 256) {
    const int c = k >> 4, t = k & 15;
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) s += red[((size_t)wv * 16 + t) * C + c];
    partial[(size_t)blockIdx.x * C * 16 + k] = s;
  }
}

// out[k] (+)= sum_b partial[b][k]: block = 32 outputs, warp w sums rows w, w+8, ... (fp64), fixed-order combine (deterministic)
__global__ void __launch_bounds__(256) partial_sum_kernel(const float* __restrict__ partial, int nblk, int n, float* __restrict__ out, int accumulate) {
  b200::pdl_enter();
  __shared__ double red[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + lane;
  double s = 0.0;
  if (k < n)
    for (int b = warp; b < nblk; b += 8) s += (double)partial[(size_t)b * n + k];
  red[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && k < n) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][lane];
    out[k] = accumulate ? out[k] + (float)t : (float)t;
  }
}

// ------------------------------------------------------------------------------------------------
// discriminator head
// ------------------------------------------------------------------------------------------------
// E[h][w][c] = sum_{kh,kw} W5[c][kh][kw] * Wl[(h-kh)*n + (w-kw)],  n = S-3
__global__ void head_weight_kernel(const float* __restrict__ w5, const float* __restrict__ wl, int S, int C, float* __restrict__ E) {
  b200::pdl_enter();
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
  b200::pdl_enter();
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
  b200::pdl_enter();
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
  b200::pdl_enter();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float p = pred[b];
  dlin[b] = sigmoid ? dpred[b] * p * (1.f - p) : dpred[b];
}

// da[b][k] = dlin[b] * E[k]
__global__ void __launch_bounds__(256) head_bwd_input_kernel(const float* __restrict__ dlin, const float* __restrict__ E, int n, int B,
                                                             __nv_bfloat16* __restrict__ da) {
  b200::pdl_enter();
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
  b200::pdl_enter();
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
// block = (tap, 32 channels); warp w walks output rows w, w+8, ...; lane = channel (coalesced 128-byte reads of dE)
__global__ void __launch_bounds__(256) head_bwd_w5_kernel(const float* __restrict__ dE, const float* __restrict__ wl, int S, int C, float* __restrict__ dw5,
                                                          int accumulate) {
  b200::pdl_enter();
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane, tap = blockIdx.y;
  const int kh = tap >> 2, kw = tap & 3;
  const int n = S - 3;
  float s = 0.f;
  if (c < C) {
    for (int oh = warp; oh < n; oh += 8)
      for (int ow = 0; ow < n; ++ow) s = fmaf(__ldg(dE + ((size_t)(oh + kh) * S + (ow + kw)) * C + c), __ldg(wl + oh * n + ow), s);
  }
  red[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][lane];
    float* dst = dw5 + (size_t)c * 16 + tap;
    *dst = accumulate ? *dst + t : t;
  }
}

// dWl[oh*n+ow] = sum_{c,kh,kw} dE[oh+kh][ow+kw][c] * W5[c][kh][kw];   block per output.
// blockIdx.x == n*n computes dbias = sum_b dlin[b].
__global__ void __launch_bounds__(256) head_bwd_wl_kernel(const float* __restrict__ dE, const float* __restrict__ w5, const float* __restrict__ dlin, int B,
                                                          int S, int C, float* __restrict__ dwl, float* __restrict__ dbias, int accumulate) {
  b200::pdl_enter();
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
  if ((C != 64 && C != 128) || (H & 1) || (W & 1) || B < 1 || ld2 % 8 || off2 % 8 || Wo % TOW || Ho % TOH || ((uintptr_t)y & 31)) {
    b200_set_error("conv_c1_fwd: C must be 64/128; output must tile by 8x32 (or 16x16); concat stride / offset multiples of 8; y 32-byte aligned");
    return B200_ERR_ARG;
  }
  const int ntiles = B * (Ho / TOH) * (Wo / TOW);
  const int grid = ntiles < 148 * 3 ? ntiles : 148 * 3;  // persistent: three resident blocks per SM
  if (C == 64)
    B200_LAUNCH((conv_c1_fwd_kernel<64>), grid, 256, 0, stream, x, xmul, B, H, W, tow_shift, w, act, (__nv_bfloat16*)y, (__nv_bfloat16*)y2, ld2, off2);
  else
    B200_LAUNCH((conv_c1_fwd_kernel<128>), grid, 256, 0, stream, x, xmul, B, H, W, tow_shift, w, act, (__nv_bfloat16*)y, (__nv_bfloat16*)y2, ld2, off2);
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
    B200_LAUNCH((convt_c1_fwd_kernel<64>), grid, 256, 0, stream, (const __nv_bfloat16*)z, B, H, W, ldz, w, bias, act, base, mul, y, comp);
  else
    B200_LAUNCH((convt_c1_fwd_kernel<128>), grid, 256, 0, stream, (const __nv_bfloat16*)z, B, H, W, ldz, w, bias, act, base, mul, y, comp);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_wgrad_c1_blocks(void) { return 148 * 3; }  // upper bound of the partial rows b200_wgrad_c1 writes (scratch sizing)

// partial: scratch of b200_wgrad_c1_blocks() * C * 16 floats
int b200_wgrad_c1(const void* z, int B, int H, int W, int C, int ldz, const float* g, const float* gmul, float* partial, float* dw, int accumulate,
                  cudaStream_t stream) {
  if ((C != 64 && C != 128) || ldz % 8 || B < 1 || H % kWgTH || W % kWgTW) {
    b200_set_error("wgrad_c1: C must be 64/128, H % 4 == 0, W % 32 == 0");
    return B200_ERR_ARG;
  }
  const int stages = C == 64 ? 3 : 2;
  const int nblk = C == 64 ? 148 * 3 : 148 * 2;  // resident blocks per SM (registers: 71 / 126 per thread)
  const size_t smem = stages * (size_t)(kWgTH * kWgTW * (C + 8) * 2) + 2 * stages * (size_t)(kWgRows * kWgStride) * sizeof(float);
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(wgrad_c1_kernel<64, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(wgrad_c1_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  });
  if (C == 64)
    B200_LAUNCH((wgrad_c1_kernel<64, 3>), nblk, 256, smem, stream, (const __nv_bfloat16*)z, B, H, W, ldz, g, gmul, partial);
  else
    B200_LAUNCH((wgrad_c1_kernel<128, 2>), nblk, 256, smem, stream, (const __nv_bfloat16*)z, B, H, W, ldz, g, gmul, partial);
  B200_CHECK_LAUNCH();
  B200_LAUNCH((partial_sum_kernel), (C * 16 + 31) / 32, 256, 0, stream, partial, nblk, C * 16, dw, accumulate);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_head_weight(const float* w5, const float* wl, int S, int C, float* E, cudaStream_t stream) {
  if (S < 4) {
    b200_set_error("head_weight: feature map smaller than the 4x4 kernel");
    return B200_ERR_ARG;
  }
  B200_LAUNCH((head_weight_kernel), (S * S * C + 255) / 256, 256, 0, stream, w5, wl, S, C, E);
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
  B200_LAUNCH((head_fwd_kernel), dim3(B200_HEAD_SPLIT, B), 256, 0, stream, (const __nv_bfloat16*)a, E, n, scratch);
  B200_CHECK_LAUNCH();
  B200_LAUNCH((head_finalize_kernel), (B + 127) / 128, 128, 0, stream, scratch, B200_HEAD_SPLIT, B, bias, sigmoid, lin, pred);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_head_dlin(const float* dpred, const float* pred, int sigmoid, int B, float* dlin, cudaStream_t stream) {
  B200_LAUNCH((head_dlin_kernel), (B + 127) / 128, 128, 0, stream, dpred, pred, sigmoid, B, dlin);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

int b200_head_bwd_input(const float* dlin, const float* E, int B, int n, void* da, cudaStream_t stream) {
  long long blocks = ((long long)B * (n / 8) + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  B200_LAUNCH((head_bwd_input_kernel), (int)blocks, 256, 0, stream, dlin, E, n, B, (__nv_bfloat16*)da);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

// dE: scratch [S*S*C] floats
int b200_head_bwd_params(const float* dlin, const void* a, int B, int S, int C, const float* w5, const float* wl, float* dE, float* dw5, float* dwl,
                         float* dbias, int accumulate, cudaStream_t stream) {
  const int n = S * S * C;
  B200_LAUNCH((head_bwd_dE_kernel), (n / 8 + 255) / 256, 256, 0, stream, dlin, (const __nv_bfloat16*)a, n, B, dE);
  B200_CHECK_LAUNCH();
  B200_LAUNCH((head_bwd_w5_kernel), dim3((C + 31) / 32, 16), 256, 0, stream, dE, wl, S, C, dw5, accumulate);
  B200_CHECK_LAUNCH();
  const int nn = (S - 3) * (S - 3);
  B200_LAUNCH((head_bwd_wl_kernel), nn + 1, 256, 0, stream, dE, w5, dlin, B, S, C, dwl, dbias, accumulate);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

}  // extern "C"
