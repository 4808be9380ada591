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
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..9 = epilogue (TMEM -> registers -> global), two per TMEM lane quadrant.
#include "common.cuh"

#include <mutex>
#include <string>

namespace b200 {

//   MODE_CONV3    y[b,h,w,:] = relu(sum_{kh,kw,ci} x[b,h+kh-1,w+kw-1,ci] * W[:,kh,kw,ci] + bias)
//                 (Conv2d k3 s1 p1 + ReLU: the VGG19 feature extractor of the perceptual / style losses, lib/models/loss.py:49-115,
//                 networks.py:367-389; window mode only)
enum { MODE_CONV_S2 = 0, MODE_CONVT = 1, MODE_WGRAD = 2, MODE_CONV3 = 3 };

// Optional backward-of-activation epilogue of the data-gradient GEMMs (mirrors b200_epi_bwd in include/b200_inpaint.h).
// Output columns >= col0 are gradients w.r.t. act(dropout(bn(y)))[:, col - col0]; the epilogue turns them into the gradient
// w.r.t. the BatchNorm output z = y*scale + shift (`dz`), stores dz instead, and (when stats are requested) accumulates the
// per-channel sums  sum dz  and  sum dz*y  that the BatchNorm backward needs — so no separate pass re-reads the gradient.
struct EpiBwd {
  const __nv_bfloat16* y;  // forward conv output (dense, pixel stride ldy); null = plain epilogue
  int ldy;
  const float* scale;      // forward BatchNorm scale / shift per channel of y (null: z = y)
  const float* shift;
  int act;                 // 0 none, 1 LeakyReLU(0.2), 2 ReLU
  const __nv_bfloat16* g2; // optional second gradient (skip connection through ReLU of leaky(z)): added where z > 0
  int ldg2, offg2;
  int dropout;             // Dropout(0.5) between BatchNorm and ReLU: keep-mask regenerated from (seed, element index)
  uint32_t seed;
  const uint32_t* seed_ctr;  // optional device counter advancing the seed (b200_set_seed_counter)
  int col0;
  int Cy;                  // channels of y
  float* dz32;             // optional: dz is ALSO the input of a BatchNorm backward -> stored unrounded, fp32 [pixel][Cy] (then the
                           // bf16 output keeps only the unfused columns); null: dz replaces the fused columns of the bf16 output
};

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
  int ksplit;            // fprop modes: K is split over `ksplit` work items per output tile (small-M layers: few output tiles,
                         // long K); each writes an fp32 partial tile to `partial` [ksplit][pixels][Cout], b200_splitk_reduce sums them
  float* partial;
  long long P_total;     // pixels of the output tensor
  int stat_parity;       // 2: the per-quadrant statistics buffer is double buffered over tiles; 1: single + a second barrier per tile
  int off_acc, off_ss, off_tile, tile_bytes;  // byte offsets of the epilogue's shared-memory regions behind the barriers (host-sized per variant)
  int pair;              // 1: CTA pairs (cta_group::2), num_tiles counts 256-pixel pair tiles
  int kpb;               // 64-wide k-blocks per pipeline stage (2 for narrow N tiles: halves the per-stage barrier overhead
                         // of the single issuing thread, which otherwise exceeds the MMA time of a 128 x 64 x 64 block)
  int win;               // 1: window operand staging (see igemm_kernel): the activation operand of a 16 x 8-pixel tile is staged ONCE per
                         // parity plane (conv) / phase (convT) and 64-channel block as a 17 x 9-pixel halo window; the four filter taps
                         // that read it are four shared-memory descriptors into that window instead of four TMA boxes
  int sa;                // window slots of the activation ring (win)
  int tpb;               // filter taps (weight boxes) per stage of the weight ring (win); `stages` counts the weight stages
  const float* bias;     // MODE_CONV3: per-output-channel bias added before the ReLU of the epilogue
  int bres;              // MODE_CONV3 with a single N tile whose whole weight sequence fits the weight ring: the ring is filled once (first
                         // tile) and stays resident — no weight traffic, no weight-stage hand-shakes for the CTA's other tiles
  int pool;              // MODE_CONV3: 2 x 2 max-pool fused into the epilogue (the output tensor is [B, Hs/2, Ws/2, ldc])
  int gram;              // MODE_WGRAD without taps: the second operand is read at the SAME pixels as the first (tmB is a plain NHWC map),
                         // G[cs][cl] = sum_pixels S[pix][cs] * L[pix][cl]; one split per image gives the per-image Gram matrices of
                         // lib/models/loss.py:117-136
  float* stats;          // optional (fprop modes): per-CTA column sums of the stored bf16 outputs,
                         // [gridDim.x][2][Cout] (sum, sum of squares) for the BatchNorm that follows;
                         // with `bwd`: (sum dz, sum dz*y) of the fused columns instead
  EpiBwd bwd;
};

constexpr int kNumSMs = 148;
constexpr int kEpiWarps = 8;                       // two warps per TMEM lane quadrant, alternating 32-column chunks
constexpr int kThreads = 64 + 32 * kEpiWarps;      // warp 0 = TMA, warp 1 = MMA, warps 2.. = epilogue
constexpr int kATileBytes = 128 * 128;  // 128 rows x 64 bf16 (K-major) or 2 x (64 px x 64 ch) (MN-major)
// window operand staging: a pixel tile is 8 wide x 16 high; its halo window is 9 x 17 pixels x 64 channels = 153 rows of 128 bytes
constexpr int kWinTw = 8, kWinTh = 16;
// WIN = 1: one parity plane (conv) / one phase (convT): halo of 1 pixel on one side per axis -> 17 x 9 window, 4 taps.
// WIN = 4: ALL FOUR phases of a transposed convolution from one window: halo of 1 pixel on both sides -> 18 x 10 window, 16
//          (phase, tap) descriptors, four TMEM accumulators (N <= 64).  0: legacy (one TMA box per tap).
// WIN = 3: the 9 taps of a 3x3 / stride-1 / pad-1 convolution from the same 18 x 10 window (tap (kh, kw) = window row kh * 10 + kw).
__host__ __device__ constexpr bool win_wide(int win) { return win == 4 || win == 3; }                               // halo on both sides
__host__ __device__ constexpr int win_pitch(int win) { return win_wide(win) ? kWinTw + 2 : kWinTw + 1; }            // pixel rows per window row
__host__ __device__ constexpr int win_bytes(int win) { return win_pitch(win) * (win_wide(win) ? kWinTh + 2 : kWinTh + 1) * 128; }  // 23040 / 19584
__host__ __device__ constexpr int win_slot_bytes(int win) { return win_wide(win) ? 24 * 1024 : 20 * 1024; }         // slot pitch (1024-aligned)
__host__ __device__ constexpr int win_sbo(int win) { return win_pitch(win) * 128; }                                 // byte distance of 8-pixel row groups
__host__ __device__ constexpr int win_taps(int win) { return win == 4 ? 16 : (win == 3 ? 9 : 4); }                  // descriptors per window
constexpr int kBarrierBytes = 1024;                    // mbarriers + TMEM pointer behind the operand rings (keeps the epilogue tiles 1024-byte aligned)
constexpr int kStatTileBytes = 4096 + 2048;           // per epilogue warp: 32 x 32 fp32 of dz + 32 x 32 bf16 of y (forward: 32 x 32 bf16 only)

// Epilogue variants (compile-time: the per-element code has no branches and each variant is a small kernel)
enum { ACT_BWD_NONE = 0, ACT_BWD_LEAKY = 1, ACT_BWD_RELU = 2, ACT_FWD_BIAS_RELU = 3 };  // (3: forward epilogue of MODE_CONV3, BWD = false)
template <bool STATS_, bool BWD_, int ACT_, bool G2_, bool BN_>
struct Epi {
  static constexpr bool STATS = STATS_, BWD = BWD_, G2 = G2_, BN = BN_;
  static constexpr int ACT = ACT_;
};

struct TileCoord {
  int w0, h0, b0, n0, m0, phase, kb_begin, kb_count, row, ks;
};

// Work item t of the persistent loop.  fprop/dgrad: the pixel tile is the slow index so the CTAs running at the same
// time read the same activations (4 phases x n_tiles) out of L2; wgrad: the pixel range (split) is the slow index.
template <bool PAIR = false>
__device__ __forceinline__ TileCoord decode_tile(const IGemmParams& p, int t, int rank = 0) {
  TileCoord c;
  c.w0 = c.h0 = c.b0 = c.m0 = c.phase = c.ks = 0;
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
    if (p.ksplit > 1) {  // the K split is the slowest index: concurrently running CTAs still share their activation tiles
      const int per_split = p.num_ptiles * per_pt;
      c.ks = t / per_split;
      t -= c.ks * per_split;
      c.kb_begin = c.ks * p.kb_per_tile;
    }
    int pt = t / per_pt;
    const int r = t - pt * per_pt;
    if (PAIR) pt = 2 * pt + rank;  // a CTA pair owns two adjacent 128-pixel tiles (one 256-row MMA tile)
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

// Column sums of a warp's 32 x 32 bf16 tile whose rows live one per lane, through a shared-memory transpose: lane r stores its
// row (4 chunks of 16 bytes, XOR-swizzled so both the 128-bit stores and loads are bank-conflict free); lane (rg, cg) then reads
// chunk cg (8 columns) of rows rg, rg+8, rg+16, rg+24 and three shuffle stages fold the 8 row groups.
// s[k] / q[k] = sum over the 32 rows of x and x*x (or x*y) for column 8*(lane & 3) + k — identical in every lane with the same cg.
template <bool WITH_Y>
__device__ __forceinline__ void tile_column_sums(uint8_t* wbuf, const uint32_t (&o)[16], const uint32_t (&yw)[16], int lane, float (&s)[8],
                                                 float (&q)[8]) {
  // store: row = lane, 4 chunks; physical chunk = logical ^ ((lane >> 1) & 3) keeps the 8-lane phases of STS.128 / LDS.128 conflict-free
  const int sw = (lane >> 1) & 3;
  uint4* rowp = reinterpret_cast<uint4*>(wbuf + lane * 64);
#pragma unroll
  for (int j = 0; j < 4; ++j) rowp[j ^ sw] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
  if (WITH_Y) {
    uint4* rowy = reinterpret_cast<uint4*>(wbuf + 2048 + lane * 64);
#pragma unroll
    for (int j = 0; j < 4; ++j) rowy[j ^ sw] = make_uint4(yw[4 * j], yw[4 * j + 1], yw[4 * j + 2], yw[4 * j + 3]);
  }
  __syncwarp();
  const int cg = lane & 3, rg = lane >> 2;  // column chunk (8 columns), row group (rows rg + 8 i)
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = q[k] = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = i * 8 + rg;
    const int ph = cg ^ ((row >> 1) & 3);
    const uint4 d = *reinterpret_cast<const uint4*>(wbuf + row * 64 + ph * 16);
    const uint32_t dw[4] = {d.x, d.y, d.z, d.w};
    uint32_t ywd[4] = {0u, 0u, 0u, 0u};
    if (WITH_Y) {
      const uint4 y4 = *reinterpret_cast<const uint4*>(wbuf + 2048 + row * 64 + ph * 16);
      ywd[0] = y4.x, ywd[1] = y4.y, ywd[2] = y4.z, ywd[3] = y4.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack_bf16x2(dw[k]);
      s[2 * k] += f.x;
      s[2 * k + 1] += f.y;
      if (WITH_Y) {
        const float2 g = unpack_bf16x2(ywd[k]);
        q[2 * k] = fmaf(f.x, g.x, q[2 * k]);
        q[2 * k + 1] = fmaf(f.y, g.y, q[2 * k + 1]);
      } else {
        q[2 * k] = fmaf(f.x, f.x, q[2 * k]);
        q[2 * k + 1] = fmaf(f.y, f.y, q[2 * k + 1]);
      }
    }
  }
  __syncwarp();  // the buffer may be overwritten by the next chunk
  // 8 row groups (lane bits 2..4) -> one
#pragma unroll
  for (int h = 4; h <= 16; h <<= 1)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s[k] += __shfl_xor_sync(0xffffffffu, s[k], h);
      q[k] += __shfl_xor_sync(0xffffffffu, q[k], h);
    }
}

// Same for the fused backward: dz rows are fp32 (unrounded, exactly what the BatchNorm apply pass will read), y rows bf16.
// lane (cq, rg) reads 4 columns (chunk cq) of rows rg, rg+4, ..., rg+28; two shuffle stages fold the 4 row groups.
// s[k] / q[k] = sum over the 32 rows of dz and dz*y for column 4*(lane & 7) + k.
__device__ __forceinline__ void tile_stage_f32(uint8_t* wbuf, const uint32_t (&v)[32], const uint32_t (&yw)[16], int lane) {
  uint4* rowz = reinterpret_cast<uint4*>(wbuf + lane * 128);
#pragma unroll
  for (int j = 0; j < 8; ++j) rowz[j ^ (lane & 7)] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  const int sw = (lane >> 1) & 3;
  uint4* rowy = reinterpret_cast<uint4*>(wbuf + 4096 + lane * 64);
#pragma unroll
  for (int j = 0; j < 4; ++j) rowy[j ^ sw] = make_uint4(yw[4 * j], yw[4 * j + 1], yw[4 * j + 2], yw[4 * j + 3]);
  __syncwarp();
}
// Coalesced store of the staged 32 x 32 fp32 tile: one instruction covers 4 pixel rows x 128 contiguous bytes (4 memory
// wavefronts) instead of 32 rows x 16 bytes (32 wavefronts) when every lane stores its own row.  `row_off` = element offset of this
// lane's row in the destination (fetched from the owning lane by shuffle), `ok` = row inside the batch.
__device__ __forceinline__ void tile_store_f32(const uint8_t* wbuf, float* dst, unsigned long long row_off, bool ok, int lane) {
  const int ch = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = i * 4 + (lane >> 3);
    const uint4 d = *reinterpret_cast<const uint4*>(wbuf + rr * 128 + ((ch ^ (rr & 7)) * 16));
    const unsigned long long off = __shfl_sync(0xffffffffu, row_off, rr);
    const int rok = __shfl_sync(0xffffffffu, (int)ok, rr);
    if (rok) *reinterpret_cast<uint4*>(dst + off + ch * 4) = d;
  }
}
__device__ __forceinline__ void tile_column_sums_f32(uint8_t* wbuf, int lane, float (&s)[4], float (&q)[4]) {
  const int cq = lane & 7, rg = lane >> 3;
#pragma unroll
  for (int k = 0; k < 4; ++k) s[k] = q[k] = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = rg + 4 * i;
    const float4 d = *reinterpret_cast<const float4*>(wbuf + row * 128 + ((cq ^ (row & 7)) * 16));
    const uint2 yy = *reinterpret_cast<const uint2*>(wbuf + 4096 + row * 64 + (((cq >> 1) ^ ((row >> 1) & 3)) * 16) + (cq & 1) * 8);
    const float2 y0 = unpack_bf16x2(yy.x), y1 = unpack_bf16x2(yy.y);
    s[0] += d.x, s[1] += d.y, s[2] += d.z, s[3] += d.w;
    q[0] = fmaf(d.x, y0.x, q[0]), q[1] = fmaf(d.y, y0.y, q[1]), q[2] = fmaf(d.z, y1.x, q[2]), q[3] = fmaf(d.w, y1.y, q[3]);
  }
  __syncwarp();
#pragma unroll
  for (int h = 8; h <= 16; h <<= 1)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s[k] += __shfl_xor_sync(0xffffffffu, s[k], h);
      q[k] += __shfl_xor_sync(0xffffffffu, q[k], h);
    }
}

// End of every warp role: all warps of the CTA (pair) meet, then the TMEM owner frees the accumulators.  The LAST statement of each
// role branch of igemm_kernel (see the note at its role dispatch).
template <bool PAIR>
__device__ __forceinline__ void igemm_teardown(bool tmem_owner, uint32_t tmem_base, uint32_t tmem_cols) {
  tc_fence_before();
  asm volatile("bar.sync 0, %0;" ::"n"(kThreads) : "memory");
  if (PAIR) cluster_sync_all();  // the peer may still be reading operands / TMEM this CTA's MMAs touched
  if (tmem_owner) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tmem_base, tmem_cols);
    else tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---- staging tiles of the window-mode epilogue (one 32-row x 32-column chunk per warp).  Both layouts are exactly what TMA's
// SWIZZLE_64B / SWIZZLE_128B boxes look like in shared memory, so the same tile serves the per-lane row access (conflict-free
// 128-bit LDS / STS), the transposed column sums, the cp.async operand loads and the TMA store.
//   bf16 tile: 32 rows x 64 B, 16-byte chunk j of row r at  r * 64 + ((j ^ ((r >> 1) & 3)) * 16)       (SWIZZLE_64B)
//   fp32 tile: 32 rows x 128 B, chunk j of row r at         r * 128 + ((j ^ (r & 7)) * 16)            (SWIZZLE_128B)
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ void stage_bf16_row(uint8_t* tile, const uint32_t (&o)[16], int lane) {
  const int sw = (lane >> 1) & 3;
  uint4* rowp = reinterpret_cast<uint4*>(tile + lane * 64);
#pragma unroll
  for (int j = 0; j < 4; ++j) rowp[j ^ sw] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
}
__device__ __forceinline__ void load_bf16_row(const uint8_t* tile, uint32_t (&o)[16], int lane) {
  const int sw = (lane >> 1) & 3;
  const uint4* rowp = reinterpret_cast<const uint4*>(tile + lane * 64);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 d = rowp[j ^ sw];
    o[4 * j] = d.x, o[4 * j + 1] = d.y, o[4 * j + 2] = d.z, o[4 * j + 3] = d.w;
  }
}
__device__ __forceinline__ void stage_f32_row(uint8_t* tile, const uint32_t (&v)[32], int lane) {
  uint4* rowz = reinterpret_cast<uint4*>(tile + lane * 128);
#pragma unroll
  for (int j = 0; j < 8; ++j) rowz[j ^ (lane & 7)] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
// s[k] / q[k] = sum over the 32 rows of x and x * x for column 8 * (lane & 3) + k of a staged bf16 tile (identical in lanes with equal lane & 3)
__device__ __forceinline__ void column_sums_bf16(const uint8_t* tile, int lane, float (&s)[8], float (&q)[8]) {
  const int cg = lane & 3, rg = lane >> 2;
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = q[k] = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = i * 8 + rg;
    const uint4 d = *reinterpret_cast<const uint4*>(tile + row * 64 + ((cg ^ ((row >> 1) & 3)) * 16));
    const uint32_t dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack_bf16x2(dw[k]);
      s[2 * k] += f.x;
      s[2 * k + 1] += f.y;
      q[2 * k] = fmaf(f.x, f.x, q[2 * k]);
      q[2 * k + 1] = fmaf(f.y, f.y, q[2 * k + 1]);
    }
  }
#pragma unroll
  for (int h = 4; h <= 16; h <<= 1)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s[k] += __shfl_xor_sync(0xffffffffu, s[k], h);
      q[k] += __shfl_xor_sync(0xffffffffu, q[k], h);
    }
}
// s[k] / q[k] = sum over the 32 rows of dz and dz * y for column 4 * (lane & 7) + k: dz from a staged fp32 tile, y from a staged bf16 tile
__device__ __forceinline__ void column_sums_f32(const uint8_t* ztile, const uint8_t* ytile, int lane, float (&s)[4], float (&q)[4]) {
  const int cq = lane & 7, rg = lane >> 3;
#pragma unroll
  for (int k = 0; k < 4; ++k) s[k] = q[k] = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = rg + 4 * i;
    const float4 d = *reinterpret_cast<const float4*>(ztile + row * 128 + ((cq ^ (row & 7)) * 16));
    const uint2 yy = *reinterpret_cast<const uint2*>(ytile + row * 64 + (((cq >> 1) ^ ((row >> 1) & 3)) * 16) + (cq & 1) * 8);
    const float2 y0 = unpack_bf16x2(yy.x), y1 = unpack_bf16x2(yy.y);
    s[0] += d.x, s[1] += d.y, s[2] += d.z, s[3] += d.w;
    q[0] = fmaf(d.x, y0.x, q[0]), q[1] = fmaf(d.y, y0.y, q[1]), q[2] = fmaf(d.z, y1.x, q[2]), q[3] = fmaf(d.w, y1.y, q[3]);
  }
#pragma unroll
  for (int h = 8; h <= 16; h <<= 1)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s[k] += __shfl_xor_sync(0xffffffffu, s[k], h);
      q[k] += __shfl_xor_sync(0xffffffffu, q[k], h);
    }
}

// PAIR: two CTAs of a cluster (the two SMs of a TPC) run ONE tcgen05.mma.cta_group::2 per K step over a 256-pixel x BN tile:
// each CTA stages its own 128 pixel rows of A and HALF of the weight tile (BN/2 rows), so the operand bytes pulled through
// L2 per flop drop by a third at BN = 256 — the GEMMs with 128 x BN single-CTA tiles are bound by exactly that traffic.
// The even CTA issues the MMAs; TMA transaction bytes of both CTAs are counted on its barriers; tcgen05.commit multicasts
// the "slot free" / "accumulator ready" arrivals to both CTAs; both epilogues release the accumulator on the even CTA.
//
// WIN (window operand staging, fprop / dgrad modes with at least 16 x 8 output pixels per image): every input pixel of a
// stride-2 4x4 convolution is read by 4 filter taps (of the 2x2-tap phase of a transposed convolution: by all 4), so staging
// one TMA box per tap pulls each activation through L2 -> shared memory four times — and that stream, not the tensor pipe,
// bounds these GEMMs (~42 B/clk/SM of L2 bandwidth against 128 x N x 16 MACs per 4 KB of A).  Here the 17 x 9-pixel halo
// window of a 16 x 8-pixel tile is staged ONCE per (parity plane | phase, 64-channel block) and tap (eh, ew) is the K-major
// SWIZZLE_128B descriptor that starts at window row eh * 9 + ew with a stride of 9 rows between its 8-row groups: the tensor
// core applies the 128-byte swizzle to absolute shared-memory addresses, so a descriptor may start at any 128-byte row of a
// TMA-written window (tools/umma_probe.cu checks exactly this on the hardware).  Activation bytes per MMA drop 3.3x; the
// weights ride in their own ring (`tpb` taps per stage) because the two operands now arrive at different rates.
template <class E, bool PAIR, int WIN>
__global__ void __launch_bounds__(kThreads, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO,
             const __grid_constant__ CUtensorMap tmZ, const IGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic (no integer round trip): the compiler keeps the shared address space and emits
  // LDS / STS instead of generic loads / stores for everything the epilogue stages here
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int b_bytes = (PAIR ? p.BN / 2 : p.BN) * 128;  // per CTA
  // WIN: [sa window slots][stages weight stages of tpb boxes]; else [stages of kpb (A tile + B tile)]
  const int stage_bytes = WIN ? p.tpb * b_bytes : p.kpb * (kATileBytes + b_bytes);
  uint8_t* smem_b = smem + (WIN ? p.sa * win_slot_bytes(WIN) : 0);
  const int rank = PAIR ? (int)(blockIdx.x & 1u) : 0;  // == %cluster_ctarank for 2 x 1 x 1 clusters; read from blockIdx it is provably warp-uniform (the MMA / TMA loops then compile to the uniform datapath)
  const int tile0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // first work item of this CTA (pair)
  const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  uint64_t* full_bar = (uint64_t*)(smem_b + p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]
  uint64_t* fullA_bar = tmem_empty_bar + 2;         // [sa] (WIN)
  uint64_t* emptyA_bar = fullA_bar + (WIN ? p.sa : 0);
  uint32_t* tmem_ptr_smem = (uint32_t*)(emptyA_bar + (WIN ? p.sa : 0));
  uint8_t* epi_smem = (uint8_t*)full_bar + kBarrierBytes;
  float* stat_smem = (float*)epi_smem;                        // [parity][TMEM quadrant][sum | sumsq][BN]
  float* stat_acc = (float*)(epi_smem + p.off_acc);           // per-CTA running [sum | sumsq][Cout]
  float* bn_ss = (float*)(epi_smem + p.off_ss);               // interleaved (scale, shift)[Cy]
  uint8_t* stat_tile = epi_smem + p.off_tile;                 // [epilogue warp][tile_bytes]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // Every warp role starts with its share of the set-up and then role_begin(): all warps (of both CTAs of a pair) meet once the
  // barriers are initialised and TMEM is allocated.  The role dispatch below is the FIRST thread-dependent branch of the kernel and
  // NOTHING follows it (each role ends with igemm_teardown): ptxas keeps the single-issuer loops on the uniform datapath only then
  // — an `if (threadIdx.x == 0)` set-up block before the dispatch, or a common teardown after it, and every tcgen05.mma / TMA issue
  // pays ELECT + VOTEU + 5-8 R2UR.BROADCAST (found by bisecting the SASS; profiles/experiments/README.md, round 2).
  auto role_begin = [&]() -> uint32_t {
    tc_fence_before();
    asm volatile("bar.sync 0, %0;" ::"n"(kThreads) : "memory");
    if (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them remotely
    tc_fence_after();
    const uint32_t tb = *tmem_ptr_smem;
    // programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) touched no global
    // memory and may overlap the previous kernel's tail; from here on its results are needed
    pdl_enter();
    return tb;
  };

  if (warp == 0) {
    // ---- set-up: descriptor prefetch and barrier initialisation (the whole warp runs it, elect.sync picks the lane)
    tma_prefetch_desc_w(&tmA);
    tma_prefetch_desc_w(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init_w(&full_bar[s], 1);
      mbar_init_w(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init_w(&tmem_full_bar[b], 1);
      mbar_init_w(&tmem_empty_bar[b], PAIR ? 2 * kEpiWarps : kEpiWarps);  // one arrival per epilogue warp (of both CTAs)
    }
    if (WIN) {
      for (int a = 0; a < p.sa; ++a) {
        mbar_init_w(&fullA_bar[a], 1);
        mbar_init_w(&emptyA_bar[a], 1);
      }
    }
    fence_barrier_init();
    const uint32_t tmem_base = role_begin();
    if (WIN) {
    // ===================== TMA producer, window operand staging =====================
    const uint32_t a_tx = (uint32_t)win_bytes(WIN) * (PAIR ? 2u : 1u);
    const uint32_t b_tx = (uint32_t)stage_bytes * (PAIR ? 2u : 1u);
    const int nB = PAIR ? rank * (p.BN / 2) : 0;
    const int planes = (p.mode == MODE_CONV_S2) ? 4 : 1;
    int sa = 0, sb = 0;
    uint32_t para = 0, parb = 0;
    for (int t = tile0; t < p.num_tiles; t += tile_step) {
      const TileCoord c = decode_tile<PAIR>(p, t, rank);
      const int ph_p = c.phase >> 1, ph_q = c.phase & 1;
      for (int pl = 0; pl < planes; ++pl) {
        const int pr = pl >> 1, pc = pl & 1;  // row / column parity of the input plane (conv)
        for (int cc = 0; cc < p.cpt; ++cc) {
          mbar_wait(&emptyA_bar[sa], para ^ 1u);
          if (!PAIR || rank == 0) mbar_arrive_expect_tx_w(&fullA_bar[sa], a_tx);
          uint8_t* sW = smem + sa * win_slot_bytes(WIN);
          if (WIN == 4 || WIN == 3) {
            // all four phases / all nine taps: input pixels (i - 1 .. i + 16, j - 1 .. j + 8) of the tile
            if (PAIR) tma2_load_4d_w(sW, &tmA, &fullA_bar[sa], cc * 64, c.w0 - 1, c.h0 - 1, c.b0);
            else tma_load_4d_w(sW, &tmA, &fullA_bar[sa], cc * 64, c.w0 - 1, c.h0 - 1, c.b0);
          } else if (p.mode == MODE_CONV_S2) {
            // plane (pr, pc) holds input rows 2*h2 + pr: taps kh = 2*eh + 1 - pr read h2 = oh + eh - pr, so the window starts at h0 - pr
            if (PAIR) tma2_load_5d_w(sW, &tmA, &fullA_bar[sa], pc * p.C_k + cc * 64, c.w0 - pc, pr, c.h0 - pr, c.b0);
            else tma_load_5d_w(sW, &tmA, &fullA_bar[sa], pc * p.C_k + cc * 64, c.w0 - pc, pr, c.h0 - pr, c.b0);
          } else {
            // phase (p, q), tap (a, b) reads input pixel (i + p + a - 1, j + q + b - 1): the window starts at (h0 + p - 1, w0 + q - 1)
            if (PAIR) tma2_load_4d_w(sW, &tmA, &fullA_bar[sa], cc * 64, c.w0 + ph_q - 1, c.h0 + ph_p - 1, c.b0);
            else tma_load_4d_w(sW, &tmA, &fullA_bar[sa], cc * 64, c.w0 + ph_q - 1, c.h0 + ph_p - 1, c.b0);
          }
          for (int tg = 0; tg < win_taps(WIN); tg += p.tpb) {
            if ((WIN == 3 || WIN == 1) && p.bres && t != tile0) break;  // resident weights: loaded by the first tile
            mbar_wait(&empty_bar[sb], parb ^ 1u);
            if (!PAIR || rank == 0) mbar_arrive_expect_tx_w(&full_bar[sb], b_tx);
            for (int u = 0; u < p.tpb; ++u) {
              const int j = tg + u, eh = (j >> 1) & 1, ew = j & 1;
              uint8_t* sB = smem_b + sb * stage_bytes + u * b_bytes;
              int kb, row;
              if (WIN == 4) {  // j = phase * 4 + tap
                kb = ((j & 3) << p.cpt_sh) + cc;
                row = (j >> 2) * p.Cout + c.n0 + nB;
              } else if (WIN == 3) {  // j = kh * 3 + kw
                kb = (j << p.cpt_sh) + cc;
                row = c.n0 + nB;
              } else if (p.mode == MODE_CONV_S2) {
                const int kh = 2 * eh + 1 - pr, kw = 2 * ew + 1 - pc;
                kb = ((kh * 4 + kw) << p.cpt_sh) + cc;
                row = c.n0 + nB;
              } else {
                kb = (j << p.cpt_sh) + cc;
                row = c.phase * p.Cout + c.n0 + nB;
              }
              if (PAIR) tma2_load_2d_w(sB, &tmB, &full_bar[sb], kb * 64, row);
              else tma_load_2d_w(sB, &tmB, &full_bar[sb], kb * 64, row);
            }
            if (++sb == p.stages) {
              sb = 0;
              parb ^= 1u;
            }
          }
          if (++sa == p.sa) {
            sa = 0;
            para ^= 1u;
          }
        }
      }
    }
    } else {
    // ===================== TMA producer =====================
    {  // the whole warp runs the loop; elect.sync inside the *_w wrappers picks the issuing lane
      const uint32_t tx_bytes = (uint32_t)stage_bytes * (PAIR ? 2u : 1u);  // PAIR: both CTAs' bytes land on the even CTA's barrier
      const int nB = PAIR ? rank * (p.BN / 2) : 0;                         // this CTA's half of the weight tile
      int s = 0;
      uint32_t par = 0;
      for (int t = tile0; t < p.num_tiles; t += tile_step) {
        const TileCoord c = decode_tile<PAIR>(p, t, rank);
        const int ph_p = c.phase >> 1, ph_q = c.phase & 1;
        for (int i = 0; i < c.kb_count; i += p.kpb) {
          mbar_wait(&empty_bar[s], par ^ 1u);
          if (!PAIR || rank == 0) mbar_arrive_expect_tx_w(&full_bar[s], tx_bytes);
          for (int u = 0; u < p.kpb; ++u) {
          uint8_t* sA = smem + s * stage_bytes + u * kATileBytes;
          uint8_t* sB = smem + s * stage_bytes + p.kpb * kATileBytes + u * b_bytes;
          const int kb = c.kb_begin + i + u;
          if (p.mode == MODE_CONV_S2) {
            const int tap = kb >> p.cpt_sh, cc = kb & (p.cpt - 1);
            const int kh = tap >> 2, kw = tap & 3;
            if (PAIR) {
              tma2_load_5d_w(sA, &tmA, &full_bar[s], ((kw + 1) & 1) * p.C_k + cc * 64, c.w0 + ((kw + 1) >> 1) - 1, (kh + 1) & 1,
                           c.h0 + ((kh + 1) >> 1) - 1, c.b0);
              tma2_load_2d_w(sB, &tmB, &full_bar[s], kb * 64, c.n0 + nB);
            } else {
              tma_load_5d_w(sA, &tmA, &full_bar[s], ((kw + 1) & 1) * p.C_k + cc * 64, c.w0 + ((kw + 1) >> 1) - 1, (kh + 1) & 1,
                          c.h0 + ((kh + 1) >> 1) - 1, c.b0);
              tma_load_2d_w(sB, &tmB, &full_bar[s], kb * 64, c.n0);
            }
          } else if (p.mode == MODE_CONVT) {
            const int tap = kb >> p.cpt_sh, cc = kb & (p.cpt - 1);
            const int ta = tap >> 1, tb = tap & 1;
            if (PAIR) {
              tma2_load_4d_w(sA, &tmA, &full_bar[s], cc * 64, c.w0 + ph_q + tb - 1, c.h0 + ph_p + ta - 1, c.b0);
              tma2_load_2d_w(sB, &tmB, &full_bar[s], kb * 64, c.phase * p.Cout + c.n0 + nB);
            } else {
              tma_load_4d_w(sA, &tmA, &full_bar[s], cc * 64, c.w0 + ph_q + tb - 1, c.h0 + ph_p + ta - 1, c.b0);
              tma_load_2d_w(sB, &tmB, &full_bar[s], kb * 64, c.phase * p.Cout + c.n0);
            }
          } else {
            const int wt = kb & (p.tiles_w - 1);
            const int ht = (kb >> p.tiles_w_sh) & (p.tiles_h - 1);
            const int bt = kb >> (p.tiles_w_sh + p.tiles_h_sh);
            const int pw0 = wt << p.tw_sh, ph0 = ht << p.th_sh, pb0 = bt * p.nb;
            tma_load_4d_w(sA, &tmA, &full_bar[s], c.m0, pw0, ph0, pb0);
            tma_load_4d_w(sA + 8192, &tmA, &full_bar[s], c.m0 + 64, pw0, ph0, pb0);
            const int nsub = p.BN >> 6;
            for (int j = 0; j < nsub; ++j) {
              const int nidx = c.n0 + 64 * j;
              if (p.gram) {
                tma_load_4d_w(sB + j * 8192, &tmB, &full_bar[s], nidx, pw0, ph0, pb0);
                continue;
              }
              const int tap = nidx >> p.ck_sh, c0 = nidx & (p.C_k - 1);
              const int kh = tap >> 2, kw = tap & 3;
              tma_load_5d_w(sB + j * 8192, &tmB, &full_bar[s], ((kw + 1) & 1) * p.C_k + c0, pw0 + ((kw + 1) >> 1) - 1, (kh + 1) & 1,
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
    }
    igemm_teardown<PAIR>(false, tmem_base, p.tmem_cols);
  } else if (warp == 1) {
    // ---- set-up: this warp owns the tensor memory
    if (PAIR) {
      tmem_alloc2(tmem_ptr_smem, p.tmem_cols);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_ptr_smem, p.tmem_cols);
      tmem_relinquish();
    }
    const uint32_t tmem_base = role_begin();
    if (WIN) {
    // ===================== MMA issuer, window operand staging =====================
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, p.BN, 0, 0);
      const int nwin = ((p.mode == MODE_CONV_S2) ? 4 : 1) * p.cpt;
      const int acc_cols = (WIN == 4 ? 4 : 1) * p.BN;  // TMEM columns of one accumulator buffer
      int sa = 0, sb = 0, lt = 0;
      uint32_t para = 0, parb = 0;
      for (int t = tile0; t < p.num_tiles; t += tile_step, ++lt) {
        const int buf = lt & 1;
        mbar_wait(&tmem_empty_bar[buf], (uint32_t)(((lt >> 1) & 1) ^ 1));  // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_addr = tmem_base + (uint32_t)(buf * acc_cols);
        for (int wi = 0; wi < nwin; ++wi) {
          mbar_wait(&fullA_bar[sa], para);
          tc_fence_after();
          const uint32_t w_lo = desc_lo(smem_u32(smem + sa * win_slot_bytes(WIN)), 16u);
          for (int tg = 0; tg < win_taps(WIN); tg += p.tpb) {
            const bool resident = (WIN == 3 || WIN == 1) && p.bres;  // the weight ring holds the whole sequence of a tile and is never refilled
            if (!resident || lt == 0) mbar_wait(&full_bar[sb], parb);
            tc_fence_after();
            const uint32_t b_lo0 = desc_lo(smem_u32(smem_b + sb * stage_bytes), 16u);
            for (int u = 0; u < p.tpb; ++u) {
              const int j = tg + u;
              // WIN 1: tap (eh, ew) = window row eh * 9 + ew.  WIN 4: phase (p, q), tap (a, b) = window row (p + a) * 10 + (q + b), accumulator
              // of phase 2p + q (one window row = 128 bytes = 8 descriptor units)
              const int ph = WIN == 4 ? (j >> 2) : 0, tp = WIN == 4 ? (j & 3) : j;
              const int wr = WIN == 4   ? ((ph >> 1) + (tp >> 1)) * win_pitch(WIN) + (ph & 1) + (tp & 1)
                             : WIN == 3 ? (tp / 3) * win_pitch(WIN) + tp % 3
                                        : (tp >> 1) * win_pitch(WIN) + (tp & 1);
              const uint32_t a_lo = w_lo + (uint32_t)(wr * 8);
              const uint32_t b_lo = b_lo0 + (uint32_t)u * ((uint32_t)b_bytes >> 4);
              const uint32_t d_ph = d_addr + (uint32_t)(ph * p.BN);
#pragma unroll
              for (int k = 0; k < 4; ++k)  // K = 16 slices: 32 bytes = 2 descriptor units apart
                umma_bf16_lo_w<kDescHiSw128 | (win_sbo(WIN) >> 4), kDescHiSw128 | (1024 >> 4), PAIR>(d_ph, a_lo + 2 * k, b_lo + 2 * k, idesc,
                                                                                                        (wi > 0 || tp > 0 || k > 0) ? 1u : 0u);
            }
            if (!resident) {
              if (PAIR) umma_commit_2cta_w(&empty_bar[sb]);
              else umma_commit_w(&empty_bar[sb]);
            }
            if (++sb == p.stages) {
              sb = 0;
              parb ^= 1u;
            }
          }
          if (PAIR) umma_commit_2cta_w(&emptyA_bar[sa]);  // the window is free once the MMAs of its last tap have retired
          else umma_commit_w(&emptyA_bar[sa]);
          if (++sa == p.sa) {
            sa = 0;
            para ^= 1u;
          }
        }
        if (PAIR) umma_commit_2cta_w(&tmem_full_bar[buf]);
        else umma_commit_w(&tmem_full_bar[buf]);
      }
    }
    } else {
    // ===================== MMA issuer =====================
    // the whole warp runs this loop with identical values (waits included); elect.sync inside the *_w wrappers picks the lane that
    // issues, so descriptors and addresses stay in uniform registers (see common.cuh)
    if (rank == 0) {
      const bool mn_major = (p.mode == MODE_WGRAD);
      const uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, p.BN, mn_major ? 1 : 0, mn_major ? 1 : 0);
      const uint32_t lbo = mn_major ? 8192u : 16u;
      const uint32_t kstep = mn_major ? 2048u / 16u : 32u / 16u;  // descriptor units (16 bytes) per K = 16 slice inside a stage
      int s = 0, lt = 0;
      uint32_t par = 0;
      for (int t = tile0; t < p.num_tiles; t += tile_step, ++lt) {
        const TileCoord c = decode_tile<PAIR>(p, t, rank);
        const int buf = lt & 1;
        mbar_wait(&tmem_empty_bar[buf], (uint32_t)(((lt >> 1) & 1) ^ 1));  // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_addr = tmem_base + (uint32_t)(buf * p.BN);
        for (int i = 0; i < c.kb_count; i += p.kpb) {
          mbar_wait(&full_bar[s], par);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
          const uint32_t a_lo0 = desc_lo(a_addr, lbo), b_lo0 = desc_lo(a_addr + p.kpb * kATileBytes, lbo);
          for (int u = 0; u < p.kpb; ++u) {
            const uint32_t a_lo = a_lo0 + (uint32_t)u * (kATileBytes >> 4), b_lo = b_lo0 + (uint32_t)u * ((uint32_t)b_bytes >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_lo_w<kDescHiSw128 | (1024 >> 4), kDescHiSw128 | (1024 >> 4), PAIR>(d_addr, a_lo + k * kstep, b_lo + k * kstep, idesc,
                                                                                             (i > 0 || u > 0 || k > 0) ? 1u : 0u);
          }
          if (PAIR) umma_commit_2cta_w(&empty_bar[s]);  // frees the slot in both CTAs
          else umma_commit_w(&empty_bar[s]);            // frees the smem slot once these MMAs retire
          if (++s == p.stages) {
            s = 0;
            par ^= 1u;
          }
        }
        if (PAIR) umma_commit_2cta_w(&tmem_full_bar[buf]);
        else umma_commit_w(&tmem_full_bar[buf]);  // accumulator complete
      }
    }
    }
    igemm_teardown<PAIR>(true, tmem_base, p.tmem_cols);
  } else {
    // ===================== epilogue (overlaps the next tile's main loop) =====================
    const uint32_t tmem_base = role_begin();
    const int quad = warp & 3;          // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int half = (warp - 2) >> 2;   // which of the alternating 32-column chunks this warp takes
    const int r = quad * 32 + lane;
    const int et = threadIdx.x - 64;    // 0 .. kEpiThreads-1
    constexpr int kEpiThreads = 32 * kEpiWarps;
    if (E::STATS) {
      for (int e = et; e < 2 * p.Cout; e += kEpiThreads) stat_acc[e] = 0.f;
    }
    if (!E::BWD && E::ACT == ACT_FWD_BIAS_RELU) {  // MODE_CONV3: the bias vector (bn_ss doubles as its staging area)
      for (int e = et; e < p.Cout; e += kEpiThreads) bn_ss[e] = p.bias ? __ldg(p.bias + e) : 0.f;
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
    }
    if (E::BWD && E::BN) {
      for (int e = et; e < p.bwd.Cy; e += kEpiThreads) {
        bn_ss[2 * e] = __ldg(p.bwd.scale + e);
        bn_ss[2 * e + 1] = __ldg(p.bwd.shift + e);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
    }
    uint8_t* wbuf = stat_tile + (warp - 2) * p.tile_bytes;
    const uint32_t drop_seed = (E::BWD && p.bwd.dropout) ? dropout_seed(p.bwd.seed, p.bwd.seed_ctr) : 0u;
    int lt = 0;
    if (WIN) {
      // ---- window mode (16 x 8-pixel tiles, never masked): coalesced epilogue I/O.  A warp's 32 accumulator rows are 4 image rows
      // of 8 adjacent pixels, so a 32-row x 32-channel chunk of any NHWC tensor is a 4-D / 5-D TMA box:
      //  * the forward conv output y and the skip gradient the fused backward needs are fetched with cp.async, 4 lanes per pixel
      //    row (8 full 64-byte segments per instruction instead of 32 scattered 16-byte pieces), one chunk ahead, straight into
      //    the swizzled tile layout the column sums read anyway;
      //  * every output chunk (bf16 tile, or the unrounded fp32 dz of a BatchNorm backward) is staged in a swizzled tile and
      //    leaves through ONE TMA store (cp.async.bulk.tensor shared -> global) issued by lane 0.
      // Per-lane 16-byte global accesses made these epilogues LSU-bound (32 L1 wavefronts per instruction): e.g. the data
      // gradient into a 64-channel layer took 141 us against 62 us of MMA time.
      constexpr int kZ = (E::STATS && E::BWD) ? 4096 : 2048;
      uint8_t* zbuf = wbuf;         // fp32 dz tile (SWIZZLE_128B) or bf16 output tile (SWIZZLE_64B)
      uint8_t* ybuf = wbuf + kZ;    // [2][2048] y chunks (double buffered: fetched one chunk ahead)
      uint8_t* gbuf = ybuf + 4096;  // [2][2048] skip-gradient chunks
      const int hq = quad * 4;      // first tile row of this warp's accumulator rows
      const int frow = lane >> 2, fch = lane & 3;  // cp.async mapping: pixel of the image row, 16-byte channel chunk
      const bool convt = (p.mode == MODE_CONVT);
      uint32_t nchunk = 0;
      constexpr int NPH = WIN == 4 ? 4 : 1;  // phases of a transposed convolution handled per tile
      const int acc_cols = NPH * p.BN;       // TMEM columns of one accumulator buffer
      for (int t = tile0; t < p.num_tiles; t += tile_step, ++lt) {
        const TileCoord c = decode_tile<PAIR>(p, t, rank);
        const int buf = lt & 1;
        auto pixel = [&](int phase, int hl, int wl) -> size_t {
          if (!convt) return ((size_t)c.b0 * p.Hs + (c.h0 + hl)) * p.Ws + (c.w0 + wl);
          return ((size_t)c.b0 * (2 * p.Hs) + (2 * (c.h0 + hl) + (phase >> 1))) * (size_t)(2 * p.Ws) + (2 * (c.w0 + wl) + (phase & 1));
        };
        const int sp = (p.stat_parity == 2) ? (lt & 1) : 0;
        float* st = stat_smem + (sp * 4 + quad) * 2 * p.BN;
        // one cp.async group per call (empty when the chunk needs nothing): the y / skip-gradient slices of chunk (phase, cc)
        auto fetch = [&](int phase, int cc, uint32_t slot, bool live) {
          const int cb = c.n0 + cc - p.bwd.col0;
          if (live && cb >= 0 && cc < p.BN) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int row = 8 * i + frow;
              const size_t fp = pixel(phase, hq + i, frow);
              const int off = (int)slot * 2048 + row * 64 + ((fch ^ ((row >> 1) & 3)) * 16);
              cp_async16(ybuf + off, p.bwd.y + fp * p.bwd.ldy + cb + fch * 8);
              if (E::G2) cp_async16(gbuf + off, p.bwd.g2 + fp * p.bwd.ldg2 + p.bwd.offg2 + cb + fch * 8);
            }
          }
          cp_async_commit();
        };
        const int phase0 = WIN == 4 ? 0 : c.phase;
        if (E::BWD) fetch(phase0, half * 32, nchunk & 1u, true);  // the first chunk's operands travel while the tile's MMAs are still running
        mbar_wait(&tmem_full_bar[buf], (uint32_t)((lt >> 1) & 1));
        tc_fence_after();
#pragma unroll 1
        for (int ph = 0; ph < NPH; ++ph) {
        const int phase = WIN == 4 ? ph : c.phase;
        const int ph_p = phase >> 1, ph_q = phase & 1;
        const size_t pix = pixel(phase, r >> 3, r & 7);  // this lane's own accumulator row
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * acc_cols + ph * p.BN);
        const bool first_ph = (ph == 0);
        for (int cc = half * 32; cc < p.BN; cc += 64, ++nchunk) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + cc, v);
          const uint32_t slot = nchunk & 1u;
          if (E::BWD) {
            if (cc + 64 < p.BN) fetch(phase, cc + 64, slot ^ 1u, true);         // next chunk of this phase
            else fetch(phase + 1, half * 32, slot ^ 1u, ph + 1 < NPH);         // first chunk of the next phase of this tile (if any)
            cp_async_wait<1>();  // this chunk's group (the next chunk's may still be in flight)
          }
          if (lane == 0) tma_store_wait_read();  // the previous store has read the staging tile
          __syncwarp();                          // ... and every lane's cp.async data is visible to the whole warp
          tmem_ld_wait();
          const int cb = c.n0 + cc - p.bwd.col0;  // first channel of y this chunk is the gradient of (fused columns only)
          const bool fused = E::BWD && cb >= 0;
          if (fused) {
            // ---- backward of [BatchNorm ->] [dropout ->] activation, evaluated at the forward's stored conv output
            uint32_t yw[16], gw[16];
            load_bf16_row(ybuf + slot * 2048, yw, lane);
            if (E::G2) load_bf16_row(gbuf + slot * 2048, gw, lane);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const float2 yf = unpack_bf16x2(yw[k]);
              float2 gf = make_float2(0.f, 0.f);
              if (E::G2) gf = unpack_bf16x2(gw[k]);
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int jj = 2 * k + e;
                const float yy = e ? yf.y : yf.x;
                float z = yy;
                if (E::BN) {
                  const float2 ss = *reinterpret_cast<const float2*>(bn_ss + 2 * (cb + jj));
                  z = fmaf(yy, ss.x, ss.y);
                }
                float g = __uint_as_float(v[jj]);
                if (E::ACT == ACT_BWD_LEAKY) {
                  g = z > 0.f ? (E::G2 ? g + (e ? gf.y : gf.x) : g) : 0.2f * g;
                } else if (E::ACT == ACT_BWD_RELU) {
                  g = z > 0.f ? g : 0.f;
                  if (p.bwd.dropout) g = dropout_keep(drop_seed, (uint64_t)(pix * p.bwd.Cy + cb + jj)) ? 2.f * g : 0.f;
                }
                v[jj] = __float_as_uint(g);
              }
            }
          }
          if (E::STATS && E::BWD && fused) {
            // dz feeds a BatchNorm backward: stored unrounded (fp32) and summed per channel together with dz * y
            stage_f32_row(zbuf, v, lane);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              if (convt) tma_store_5d(&tmZ, zbuf, ph_q * p.bwd.Cy + cb, c.w0, ph_p, c.h0 + hq, c.b0);
              else tma_store_4d(&tmZ, zbuf, cb, c.w0, c.h0 + hq, c.b0);
              tma_store_commit();
            }
            float cs[4], cq[4];
            column_sums_f32(zbuf, ybuf + slot * 2048, lane, cs, cq);
            if (lane < 8) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                float* s0 = st + cc + lane * 4 + k;
                s0[0] = first_ph ? cs[k] : s0[0] + cs[k];
                s0[p.BN] = first_ph ? cq[k] : s0[p.BN] + cq[k];
              }
            }
          } else {
            if (!E::BWD && E::ACT == ACT_FWD_BIAS_RELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaxf(__uint_as_float(v[j]) + bn_ss[c.n0 + cc + j], 0.f));
            }
            uint32_t o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
            if (!E::BWD && E::ACT == ACT_FWD_BIAS_RELU && p.pool) {
              // MaxPool2d(2, 2) of the layer's output without the output ever reaching memory: a warp's 32 accumulator rows are
              // 4 image rows x 8 pixels (lane = 8 * row + column), so the 2 x 2 window of a pooled pixel is lanes {l, l^1, l^8, l^9};
              // the 2 x 4 pooled pixels of the warp are staged as an 8-row tile and leave through one TMA store
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const uint32_t t = max_bf16x2(o[j], __shfl_xor_sync(0xffffffffu, o[j], 1));
                o[j] = max_bf16x2(t, __shfl_xor_sync(0xffffffffu, t, 8));
              }
              if (!(lane & 9)) stage_bf16_row(zbuf, o, ((lane >> 4) << 2) | ((lane & 7) >> 1));
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                tma_store_4d(&tmO, zbuf, c.n0 + cc, c.w0 >> 1, (c.h0 + hq) >> 1, c.b0);
                tma_store_commit();
              }
            } else {
            stage_bf16_row(zbuf, o, lane);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              if (convt) tma_store_5d(&tmO, zbuf, ph_q * p.ldc + c.n0 + cc, c.w0, ph_p, c.h0 + hq, c.b0);
              else tma_store_4d(&tmO, zbuf, c.n0 + cc, c.w0, c.h0 + hq, c.b0);
              tma_store_commit();
            }
            }
            if (E::STATS && E::BWD) {  // unfused columns of a fused-backward GEMM carry no statistics
              if (first_ph) {
                st[cc + lane] = 0.f;
                st[p.BN + cc + lane] = 0.f;
              }
            } else if (E::STATS) {     // per-channel sums of the values as stored
              float cs[8], cq[8];
              column_sums_bf16(zbuf, lane, cs, cq);
              if (lane < 4) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  float* s0 = st + cc + lane * 8 + k;
                  s0[0] = first_ph ? cs[k] : s0[0] + cs[k];
                  s0[p.BN] = first_ph ? cq[k] : s0[p.BN] + cq[k];
                }
              }
            }
          }
          __syncwarp();  // every lane is done with the staging tiles of this chunk
        }
        }
        tc_fence_before();  // the accumulator has been read: hand the TMEM buffer back before the cross-warp statistics
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_leader(&tmem_empty_bar[buf]);
          else mbar_arrive(&tmem_empty_bar[buf]);
        }
        if (E::STATS) {
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");  // all epilogue warps
          const float* sb = stat_smem + sp * 4 * 2 * p.BN;
          const int qs = 2 * p.BN;  // quadrant stride
          for (int e = et; e < 2 * p.BN; e += kEpiThreads) {  // element (k, n0 + col) is always owned by the same thread
            const int k = e >= p.BN, col = e - k * p.BN;
            const float a = (sb[e] + sb[qs + e]) + (sb[2 * qs + e] + sb[3 * qs + e]);
            stat_acc[k * p.Cout + c.n0 + col] += a;
          }
          if (p.stat_parity == 1) asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");  // before the next tile rewrites the buffer
        }
      }
      if (E::BWD) cp_async_wait<0>();
      if (lane == 0) tma_store_wait_all();
    } else
    for (int t = tile0; t < p.num_tiles; t += tile_step, ++lt) {
      const TileCoord c = decode_tile<PAIR>(p, t, rank);
      const int buf = lt & 1;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * p.BN);
      if (p.mode == MODE_WGRAD) {
        mbar_wait(&tmem_full_bar[buf], (uint32_t)((lt >> 1) & 1));
        tc_fence_after();
        float* dst = (float*)p.out + ((size_t)c.row * p.m_tiles * 128 + (size_t)(c.m0 + r)) * p.ldc + c.n0;
        for (int cc = half * 32; cc < p.BN; cc += 64) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + cc, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(dst + cc + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);  // (weight gradients never run as pairs)
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
        const int sp = (p.stat_parity == 2) ? (lt & 1) : 0;
        float* st = stat_smem + (sp * 4 + quad) * 2 * p.BN;
        // fused activation backward: this row's slice of the forward tensor y (and of the skip gradient) is fetched one
        // 32-channel chunk ahead — the first chunk while the tile's MMAs are still running
        uint4 ynext[4], gnext[4];
        auto fetch = [&](int cc) {
          const int cb = c.n0 + cc - p.bwd.col0;
#pragma unroll
          for (int j = 0; j < 4; ++j) ynext[j] = gnext[j] = make_uint4(0u, 0u, 0u, 0u);
          if (cb >= 0 && cc < p.BN && valid) {
            const uint4* yp = reinterpret_cast<const uint4*>(p.bwd.y + pix * p.bwd.ldy + cb);
#pragma unroll
            for (int j = 0; j < 4; ++j) ynext[j] = __ldg(yp + j);
            if (E::G2) {
              const uint4* gp = reinterpret_cast<const uint4*>(p.bwd.g2 + pix * p.bwd.ldg2 + p.bwd.offg2 + cb);
#pragma unroll
              for (int j = 0; j < 4; ++j) gnext[j] = __ldg(gp + j);
            }
          }
        };
        if (E::BWD) {
          fetch(half * 32);
          if (valid) {  // pull the rest of this row's y / skip-gradient slices towards L2 while the tile's MMAs are still running
            for (int cc = half * 32 + 64; cc < p.BN; cc += 64) {
              const int cb = c.n0 + cc - p.bwd.col0;
              if (cb < 0) continue;
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p.bwd.y + pix * p.bwd.ldy + cb));
              if (E::G2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.bwd.g2 + pix * p.bwd.ldg2 + p.bwd.offg2 + cb));
            }
          }
        }
        mbar_wait(&tmem_full_bar[buf], (uint32_t)((lt >> 1) & 1));
        tc_fence_after();
        for (int cc = half * 32; cc < p.BN; cc += 64) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + cc, v);  // .sync.aligned: all lanes participate even if the row is masked
          tmem_ld_wait();
          const int cb = c.n0 + cc - p.bwd.col0;  // first channel of y this chunk is the gradient of (fused columns only)
          const bool fused = E::BWD && cb >= 0;
          uint32_t yw[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) yw[j] = 0u;
          if (E::BWD) {
            uint32_t gw[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              yw[4 * j] = ynext[j].x, yw[4 * j + 1] = ynext[j].y, yw[4 * j + 2] = ynext[j].z, yw[4 * j + 3] = ynext[j].w;
              gw[4 * j] = gnext[j].x, gw[4 * j + 1] = gnext[j].y, gw[4 * j + 2] = gnext[j].z, gw[4 * j + 3] = gnext[j].w;
            }
            fetch(cc + 64);
            if (fused) {
              // ---- backward of [BatchNorm ->] [dropout ->] activation, evaluated at the forward's stored conv output
#pragma unroll
              for (int k = 0; k < 16; ++k) {
                const float2 yf = unpack_bf16x2(yw[k]), gf = unpack_bf16x2(gw[k]);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int jj = 2 * k + e;
                  const float yy = e ? yf.y : yf.x;
                  float z = yy;
                  if (E::BN) {
                    const float2 ss = *reinterpret_cast<const float2*>(bn_ss + 2 * (cb + jj));
                    z = fmaf(yy, ss.x, ss.y);
                  }
                  float g = __uint_as_float(v[jj]);
                  if (E::ACT == ACT_BWD_LEAKY) {
                    g = z > 0.f ? (E::G2 ? g + (e ? gf.y : gf.x) : g) : 0.2f * g;
                  } else if (E::ACT == ACT_BWD_RELU) {
                    g = z > 0.f ? g : 0.f;
                    if (p.bwd.dropout) g = dropout_keep(drop_seed, (uint64_t)(pix * p.bwd.Cy + cb + jj)) ? 2.f * g : 0.f;
                  }
                  v[jj] = __float_as_uint(g);
                }
              }
            }
          }
          uint32_t o[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
          if (!E::BWD && !E::STATS && p.ksplit > 1) {
            if (valid) {  // split-K: fp32 partial tile
              float* d32 = p.partial + ((size_t)c.ks * p.P_total + pix) * p.Cout + c.n0 + cc;
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(d32 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
          } else if (valid) {
            if (E::BWD && E::STATS && fused && p.bwd.dz32 != nullptr) {
              // stored below, coalesced, from the tile staged for the column sums
            } else {
#pragma unroll
              for (int j = 0; j < 16; j += 4) *reinterpret_cast<uint4*>(dst + cc + 2 * j) = make_uint4(o[j], o[j + 1], o[j + 2], o[j + 3]);
            }
          }
          if (E::STATS && E::BWD) {
            // (sum dz, sum dz*y) of the unrounded dz; masked rows and unfused columns count as zero
            if (!valid || !fused) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
            float cs[4], cq[4];
            tile_stage_f32(wbuf, v, yw, lane);
            if (fused && p.bwd.dz32 != nullptr) tile_store_f32(wbuf, p.bwd.dz32 + cb, (unsigned long long)pix * (unsigned long long)p.bwd.Cy, valid, lane);
            tile_column_sums_f32(wbuf, lane, cs, cq);
            if (lane < 8) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                st[cc + lane * 4 + k] = cs[k];
                st[p.BN + cc + lane * 4 + k] = cq[k];
              }
            }
          } else if (E::STATS) {
            // per-channel sums of the values as stored (masked rows count as zero)
            if (!valid) {
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = 0u;
            }
            float cs[8], cq[8];
            tile_column_sums<false>(wbuf, o, yw, lane, cs, cq);
            if (lane < 4) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                st[cc + lane * 8 + k] = cs[k];
                st[p.BN + cc + lane * 8 + k] = cq[k];
              }
            }
          }
        }
        tc_fence_before();  // the accumulator has been read: hand the TMEM buffer back before the cross-warp statistics
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_leader(&tmem_empty_bar[buf]);
          else mbar_arrive(&tmem_empty_bar[buf]);
        }
        if (E::STATS) {
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");  // all epilogue warps
          const float* sb = stat_smem + sp * 4 * 2 * p.BN;
          const int qs = 2 * p.BN;  // quadrant stride
          for (int e = et; e < 2 * p.BN; e += kEpiThreads) {  // element (k, n0 + col) is always owned by the same thread
            const int k = e >= p.BN, col = e - k * p.BN;
            const float a = (sb[e] + sb[qs + e]) + (sb[2 * qs + e] + sb[3 * qs + e]);
            stat_acc[k * p.Cout + c.n0 + col] += a;
          }
          if (p.stat_parity == 1) asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");  // before the next tile rewrites the buffer
        }
      }
    }
    if (E::STATS) {  // one deterministic partial row per CTA (static tile schedule)
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");  // the last tile's accumulation (other threads' elements) is complete
      float* gdst = p.stats + (size_t)blockIdx.x * 2 * p.Cout;
      for (int e = et; e < 2 * p.Cout; e += kEpiThreads) gdst[e] = stat_acc[e];
    }
    igemm_teardown<PAIR>(false, tmem_base, p.tmem_cols);
  }
}

// out[pix][c] (bf16, pixel stride ldy) = sum_ks partial[ks][pix][c]
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ partial, int ksplit, long long P, int C, __nv_bfloat16* __restrict__ out,
                                                            int ldy) {
  b200::pdl_enter();
  const long long chunks = P * (C >> 3);
  const long long stride = (long long)P * C;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < chunks; idx += (long long)gridDim.x * blockDim.x) {
    const long long pix = idx / (C >> 3);
    const int c0 = (int)(idx - pix * (C >> 3)) * 8;
    const float* src = partial + pix * C + c0;
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < ksplit; ++k) {
      const float4 u = __ldg(reinterpret_cast<const float4*>(src + k * stride)), w = __ldg(reinterpret_cast<const float4*>(src + k * stride + 4));
      a[0] += u.x, a[1] += u.y, a[2] += u.z, a[3] += u.w, a[4] += w.x, a[5] += w.y, a[6] += w.z, a[7] += w.w;
    }
    *reinterpret_cast<uint4*>(out + pix * ldy + c0) = make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]), pack_bf16x2(a[4], a[5]), pack_bf16x2(a[6], a[7]));
  }
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
static int encode(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
                  CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    b200_set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return B200_ERR_DRIVER;
  }
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, dtype, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    b200_set_error(("cuTensorMapEncodeTiled failed, CUresult=" + std::to_string((int)r)).c_str());
    return B200_ERR_DRIVER;
  }
  return B200_OK;
}

// plain [B,H,W,C] view, box [64 ch, tw, th, nb] (window operand staging passes the halo window tw + 1, th + 1)
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
// Store maps of the window-mode epilogue: a warp's 32 rows x 32 channels chunk (4 image rows of 8 pixels) of an NHWC tensor with
// `ld` channels per pixel and elements of `esize` bytes (2: bf16 tile, SWIZZLE_64B; 4: fp32 tile, SWIZZLE_128B).
//   plain:  [ld, W, H, B], box [32, 8, 4, 1]           (stride-2 convolution: output pixel = accumulator row)
//   parity: [2 ld, W/2, 2, H/2, B], box [32, 8, 1, 4, 1] (transposed convolution: accumulator row (i, j) of phase (p, q) is pixel (2i+p, 2j+q))
static int map_store(CUtensorMap* m, const void* base, int B, int H, int W, int ld, int esize, bool parity, bool pooled = false) {
  const CUtensorMapDataType dt = esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUtensorMapSwizzle sw = esize == 4 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  const cuuint64_t e = (cuuint64_t)esize;
  if (!parity) {
    cuuint64_t dims[4] = {(cuuint64_t)ld, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t st[3] = {(cuuint64_t)ld * e, (cuuint64_t)W * ld * e, (cuuint64_t)H * W * ld * e};
    cuuint32_t box[4] = {32, pooled ? 4u : 8u, pooled ? 2u : 4u, 1};  // pooled: the 2 x 4 pooled pixels of a warp's 4 x 8 pixels
    return encode(m, base, 4, dims, st, box, dt, sw);
  }
  cuuint64_t dims[5] = {(cuuint64_t)2 * ld, (cuuint64_t)W / 2, 2, (cuuint64_t)H / 2, (cuuint64_t)B};
  cuuint64_t st[4] = {(cuuint64_t)2 * ld * e, (cuuint64_t)W * ld * e, (cuuint64_t)2 * W * ld * e, (cuuint64_t)H * W * ld * e};
  cuuint32_t box[5] = {32, 8, 1, 4, 1};
  return encode(m, base, 5, dims, st, box, dt, sw);
}
// packed weights [rows][K] (K contiguous), box [64, BN]
static int map_weights(CUtensorMap* m, const void* base, int rows, int K, int BN) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t st[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)BN};
  return encode(m, base, 2, dims, st, box);
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && e[0]) ? atoi(e) : dflt;
}

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

struct StoreMaps {
  CUtensorMap out, dz;  // window-mode epilogue: bf16 output tensor, fp32 dz tensor of a fused BatchNorm backward (zeroed when unused)
};
template <class E, int WIN>
static void launch_variant_w(const CUtensorMap& a, const CUtensorMap& b, const StoreMaps& sm, const IGemmParams& p, int grid, size_t smem,
                             cudaStream_t stream) {
  if (p.pair) {
    static std::once_flag once2;
    std::call_once(once2, [] { cudaFuncSetAttribute(igemm_kernel<E, true, WIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = b200_pdl_gemm_enabled() ? 2 : 1;
    cudaLaunchKernelEx(&cfg, igemm_kernel<E, true, WIN>, a, b, sm.out, sm.dz, p);
    return;
  }
  static std::once_flag once;
  std::call_once(once, [] { cudaFuncSetAttribute(igemm_kernel<E, false, WIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = b200_pdl_gemm_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, igemm_kernel<E, false, WIN>, a, b, sm.out, sm.dz, p);
  }
}
template <class E>
static void launch_variant(const CUtensorMap& a, const CUtensorMap& b, const StoreMaps& sm, const IGemmParams& p, int grid, size_t smem,
                           cudaStream_t stream) {
  if (p.win == 4) launch_variant_w<E, 4>(a, b, sm, p, grid, smem, stream);
  else if (p.win) launch_variant_w<E, 1>(a, b, sm, p, grid, smem, stream);
  else launch_variant_w<E, 0>(a, b, sm, p, grid, smem, stream);
}

// Window operand staging applies to the fprop / dgrad modes when an image holds at least one 16 x 8-pixel tile and K is not
// split; B200_WINDOW=0 falls back to one TMA box per filter tap (A/B comparison).
static bool want_window(int Hs, int Ws, int ksplit) {
  static const bool enabled = env_int("B200_WINDOW", 1) != 0;
  return enabled && ksplit <= 1 && Hs >= kWinTh && Ws >= kWinTw;
}

// CTA pairs pay when the single-CTA tiles are bound by operand traffic: fprop / dgrad with enough pixel tiles to give every
// TPC a pair tile.  B200_IGEMM_2CTA=0 disables them (A/B comparison).
static bool want_pair(const IGemmParams& p, int ksplit) {
  static const bool enabled = [] {
    const char* e = getenv("B200_IGEMM_2CTA");
    return !(e && e[0] == '0');
  }();
  if (!enabled || p.mode == MODE_WGRAD || ksplit > 1 || (p.num_ptiles & 1) || p.BN < 64) return false;
  return (p.num_ptiles / 2) * p.n_tiles * p.phases >= kNumSMs / 2;
}

static int launch(const CUtensorMap& a, const CUtensorMap& b, const StoreMaps& sm, IGemmParams& p, cudaStream_t stream, int* grid_out) {
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
  // 64-wide k-blocks per pipeline stage.  The per-stage hand-shake (full / empty mbarrier round trip through the two single-issuer
  // warps) is a first-order cost of the narrow tiles: measured on B200, N = 64 tiles gain 8 % with four k-blocks per stage (two
  // 80 KB stages), N = 128 tiles are fastest with two (four leave too few stages); wide tiles keep one.  B200_KPB64 / B200_KPB128
  // override the choice for A/B runs.
  static const int kpb64 = env_int("B200_KPB64", 4), kpb128 = env_int("B200_KPB128", 2);
  p.kpb = 1;
  if (p.mode != MODE_WGRAD && p.BN <= 128) {
    p.kpb = p.BN <= 64 ? kpb64 : kpb128;
    while (p.kpb > 1 && p.kb_per_tile % p.kpb) p.kpb >>= 1;
  }
  const bool has_stats = p.stats != nullptr, has_bwd = p.bwd.y != nullptr;
  // shared memory behind the pipeline stages, sized per epilogue variant: barriers, the per-quadrant statistics buffer
  // (double buffered over tiles unless that would cost a pipeline stage), the per-CTA running sums, the BatchNorm
  // coefficients of the fused backward, and the per-warp transpose tiles
  // per-warp staging behind the statistics buffers.  Window mode: output tile (2 KB bf16 / 4 KB fp32 dz) + double-buffered y
  // (+ skip-gradient) chunks of a fused backward; otherwise the transpose tiles of the statistics epilogues
  const int tile_bytes = p.win ? (!has_bwd ? 2048 : ((has_stats ? 4096 : 2048) + 4096 + (p.bwd.g2 ? 4096 : 0)))
                               : (!has_stats ? 0 : (has_bwd ? kStatTileBytes : 2048));
  auto layout = [&](int parity) {
    int off = has_stats ? parity * 4 * 2 * p.BN * 4 : 0;
    p.stat_parity = parity;
    p.off_acc = off;
    off += has_stats ? 2 * p.Cout * 4 : 0;
    p.off_ss = off;
    off += (has_bwd && p.bwd.scale) ? 2 * p.bwd.Cy * 4 : (p.mode == MODE_CONV3 ? p.Cout * 4 : 0);
    off = (off + 15) & ~15;
    if (p.win) off = (off + 1023) & ~1023;  // TMA-swizzled tiles
    p.off_tile = off;
    p.tile_bytes = tile_bytes;
    off += kEpiWarps * tile_bytes;
    return 1024 + kBarrierBytes + off;
  };
  int extra = 0, stage_bytes = 0, ring_a = 0;
  if (p.win) {
    // activation ring: `sa` halo windows; weight ring: stages of `tpb` taps (at most 32 KB per stage).  The statistics buffer is
    // double buffered when that leaves at least 3 window slots and 2 windows' worth of weights in flight.
    const int b_bytes = (p.pair ? p.BN / 2 : p.BN) * 128;
    static const int tpb_env = env_int("B200_WIN_TPB", 0), sa_env = env_int("B200_WIN_SA", 0);
    p.tpb = b_bytes * 4 <= 32 * 1024 ? 4 : (b_bytes * 2 <= 32 * 1024 ? 2 : 1);
    if (tpb_env == 1 || tpb_env == 2 || tpb_env == 4) p.tpb = tpb_env;
    if (p.win == 3) p.tpb = b_bytes * 3 <= 48 * 1024 ? 3 : 1;  // nine taps per window
    stage_bytes = p.tpb * b_bytes;
    const int taps = win_taps(p.win), slot = win_slot_bytes(p.win);
    const int per_window = (taps / p.tpb) * stage_bytes;  // weight bytes one window consumes
    auto fits = [&](int ex, int sa) { return (227 * 1024 - ex - sa * slot) / stage_bytes; };
    extra = layout(2);
    if (has_stats && fits(extra, 3) * stage_bytes < 2 * per_window) extra = layout(1);
    p.sa = sa_env > 0 ? sa_env : ((p.win != 4 && fits(extra, 4) * stage_bytes >= 2 * per_window) ? 4 : 3);
    while (p.sa > 2 && fits(extra, p.sa) < 2) --p.sa;
    ring_a = p.sa * slot;
    p.stages = fits(extra, p.sa);
    if (p.stages > 12) p.stages = 12;
    p.bres = 0;
    static const int bres_env = env_int("B200_BRES", 1);
    if ((p.win == 3 || (p.win == 1 && p.mode == MODE_CONV_S2)) && p.n_tiles == 1 && bres_env) {
      // Single N tile: every tile of a CTA consumes the SAME weight sequence.  With the ring sized to exactly one tile's sequence the
      // stage index sequence repeats tile after tile, so the first tile's loads serve all of them: no weight traffic and no
      // weight-stage hand-shakes afterwards.  conv1_2 of VGG19 (64 -> 64): 74 KB re-read per 128-pixel tile; the 64 -> 128
      // stride-2 convolutions (G d2, D c2): 131 KB per CTA and tile against 78 KB of activation windows.
      const int planes = p.win == 1 ? 4 : 1;
      const int steps = planes * (taps / p.tpb) * p.cpt;
      int sa = p.sa;
      while (sa > 3 && fits(extra, sa) < steps) --sa;  // give window slots to the weight ring if that makes it fit
      if (steps >= 2 && steps <= 12 && fits(extra, sa) >= steps) {
        p.bres = 1, p.stages = steps, p.sa = sa;
        ring_a = p.sa * slot;
      }
    }
    if (2 * p.sa + 2 * p.stages + 4 > (kBarrierBytes - 16) / 8) {
      b200_set_error("igemm: too many pipeline barriers");
      return B200_ERR_ARG;
    }
  } else {
    for (;; p.kpb >>= 1) {  // fewer k-blocks per stage until at least two stages fit
      stage_bytes = p.kpb * (kATileBytes + (p.pair ? p.BN / 2 : p.BN) * 128);  // per CTA
      extra = layout(2);
      if (has_stats && (227 * 1024 - layout(1)) / stage_bytes > (227 * 1024 - extra) / stage_bytes) extra = layout(1);
      else extra = layout(2);
      if (p.kpb == 1 || (227 * 1024 - extra) / stage_bytes >= 2) break;
    }
    if (p.stages <= 0) {
      p.stages = (227 * 1024 - extra) / stage_bytes;
      if (p.stages > 8) p.stages = 8;
    }
  }
  if (p.stages < 2 || (size_t)p.stages * stage_bytes + ring_a + extra > 227 * 1024) {
    b200_set_error("igemm: pipeline does not fit shared memory");
    return B200_ERR_ARG;
  }
  uint32_t cols = 32;
  while ((int)cols < 2 * (p.win == 4 ? 4 : 1) * p.BN) cols <<= 1;  // two accumulator buffers (of four phases each in WIN 4)
  p.tmem_cols = cols;
  const size_t smem = (size_t)p.stages * stage_bytes + ring_a + extra;
  const int sms = b200_sm_limit();  // all 148, or fewer while a trainer keeps SMs free for gradient all-reduces
  const int grid = p.pair ? 2 * (p.num_tiles < sms / 2 ? p.num_tiles : sms / 2) : (p.num_tiles < sms ? p.num_tiles : sms);
  if (grid_out) *grid_out = grid;
  const bool bn = p.bwd.scale != nullptr, g2 = p.bwd.g2 != nullptr;
  const int act = p.bwd.act;
  if (p.mode == MODE_CONV3) {
    launch_variant_w<Epi<false, false, ACT_FWD_BIAS_RELU, false, false>, 3>(a, b, sm, p, grid, smem, stream);
  } else if (!has_bwd) {
    if (has_stats) launch_variant<Epi<true, false, 0, false, false>>(a, b, sm, p, grid, smem, stream);
    else launch_variant<Epi<false, false, 0, false, false>>(a, b, sm, p, grid, smem, stream);
  } else if (act == ACT_BWD_LEAKY && bn && g2 && has_stats) {
    launch_variant<Epi<true, true, ACT_BWD_LEAKY, true, true>>(a, b, sm, p, grid, smem, stream);
  } else if (act == ACT_BWD_LEAKY && bn && !g2 && has_stats) {
    launch_variant<Epi<true, true, ACT_BWD_LEAKY, false, true>>(a, b, sm, p, grid, smem, stream);
  } else if (act == ACT_BWD_LEAKY && !bn && g2 && !has_stats) {
    launch_variant<Epi<false, true, ACT_BWD_LEAKY, true, false>>(a, b, sm, p, grid, smem, stream);
  } else if (act == ACT_BWD_LEAKY && !bn && !g2 && !has_stats) {
    launch_variant<Epi<false, true, ACT_BWD_LEAKY, false, false>>(a, b, sm, p, grid, smem, stream);
  } else if (act == ACT_BWD_RELU && bn && !g2 && has_stats) {
    launch_variant<Epi<true, true, ACT_BWD_RELU, false, true>>(a, b, sm, p, grid, smem, stream);
  } else if (act == ACT_BWD_RELU && !bn && !g2 && !has_stats) {
    launch_variant<Epi<false, true, ACT_BWD_RELU, false, false>>(a, b, sm, p, grid, smem, stream);
  } else {
    b200_set_error("igemm: unsupported fused backward epilogue (supported: LeakyReLU [+skip] with/without BatchNorm statistics, ReLU [+dropout] "
                   "with/without BatchNorm statistics; statistics are required exactly when a BatchNorm scale/shift is given)");
    return B200_ERR_ARG;
  }
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

// C-ABI mirror of EpiBwd (include/b200_inpaint.h)
struct b200_epi_bwd {
  const void* y;
  int ldy;
  const float* scale;
  const float* shift;
  int act;
  const void* g2;
  int ldg2, offg2;
  int dropout;
  unsigned seed;
  int col0;
  int Cy;
  void* dz32;
};

static int set_bwd(IGemmParams& p, const b200_epi_bwd* e, int Cout, const char* who) {
  p.bwd = EpiBwd{};
  if (!e || !e->y) return B200_OK;
  if (e->Cy > 512 || e->Cy % 32 || e->col0 % 32 || e->col0 + e->Cy != Cout || e->ldy % 8 || e->ldg2 % 8 || e->offg2 % 8 || e->act < 0 || e->act > 2 ||
      (p.stats && Cout > 1024)) {
    b200_set_error((std::string(who) + ": bad fused activation-backward epilogue (needs col0 + Cy == Cout, Cy <= 512, 32-channel granularity)").c_str());
    return B200_ERR_ARG;
  }
  p.bwd.y = (const __nv_bfloat16*)e->y;
  p.bwd.ldy = e->ldy;
  p.bwd.scale = e->scale;
  p.bwd.shift = e->shift;
  p.bwd.act = e->act;
  p.bwd.g2 = (const __nv_bfloat16*)e->g2;
  p.bwd.ldg2 = e->ldg2;
  p.bwd.offg2 = e->offg2;
  p.bwd.dropout = e->dropout;
  p.bwd.seed = e->seed;
  p.bwd.seed_ctr = e->dropout ? b200_seed_counter_ptr() : nullptr;
  p.bwd.col0 = e->col0;
  p.bwd.Cy = e->Cy;
  p.bwd.dz32 = (float*)e->dz32;
  return B200_OK;
}

int b200_igemm_stat_rows(void) { return kNumSMs; }

// y[B,H/2,W/2,ldy] (bf16) = conv4x4/s2/p1(x[B,H,W,Cin] bf16, wpk[Cout][16*Cin] bf16 with k = (kh*4+kw)*Cin + ci)
// stats (optional): [b200_igemm_stat_rows()][2][Cout] fp32 per-CTA column sums / sums of squares of y; *stat_rows = rows written.
static int set_split(IGemmParams& p, int ksplit, float* partial, long long P_total, const char* who) {
  p.ksplit = 1;
  p.partial = nullptr;
  p.P_total = P_total;
  if (ksplit <= 1) return B200_OK;
  if (!partial || p.stats || p.bwd.y || p.kb_total % ksplit || ((p.kb_total / ksplit) & 1)) {
    b200_set_error((std::string(who) + ": split-K needs a partial buffer, no fused epilogue, and an even number of k-blocks per split").c_str());
    return B200_ERR_ARG;
  }
  p.ksplit = ksplit;
  p.partial = partial;
  p.kb_per_tile = p.kb_total / ksplit;
  p.num_tiles *= ksplit;
  return B200_OK;
}

// ksplit > 1: K is split; `partial` (ksplit * pixels * Cout floats) receives fp32 partial sums and y is NOT written — finish with b200_splitk_reduce.
int b200_conv_s2_gemm(const void* x, int B, int H, int W, int Cin, const void* wpk, int Cout, void* y, int ldy, int bn_hint, int stages, float* stats,
                      int* stat_rows, const b200_epi_bwd* bwd, int ksplit, float* partial, cudaStream_t stream) {
  if (!is_pow2(H) || !is_pow2(W) || H < 2 || W < 2 || Cin % 64 || Cout % 32 || ldy % 8 || B < 1) {
    b200_set_error("conv_s2_gemm: unsupported shape");
    return B200_ERR_ARG;
  }
  IGemmParams p{};
  p.mode = MODE_CONV_S2;
  p.B = B;
  p.Hs = H / 2;
  p.Ws = W / 2;
  // (a fused BatchNorm backward without the fp32 dz tensor keeps dz in the bf16 output: legacy epilogue only)
  p.win = (want_window(p.Hs, p.Ws, ksplit) && !(stats && bwd && bwd->y && !bwd->dz32)) ? 1 : 0;
  if (p.win) p.tw = kWinTw, p.th = kWinTh, p.nb = 1;
  else pixel_box(p.Hs, p.Ws, 128, &p.tw, &p.th, &p.nb);
  p.tiles_w = p.Ws / p.tw;
  p.tiles_h = p.Hs / p.th;
  const int tiles_b = (B + p.nb - 1) / p.nb;
  p.num_ptiles = p.tiles_w * p.tiles_h * tiles_b;
  p.C_k = Cin;
  p.cpt = Cin / 64;
  p.kb_per_tile = p.kb_total = 16 * p.cpt;
  p.BN = pick_bn(Cout, bn_hint);
  if (Cout % p.BN || p.BN % 32 || p.BN > 256 || (stats && Cout > 1024)) {
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
  int rc;
  if ((rc = set_bwd(p, bwd, Cout, "conv_s2_gemm"))) return rc;
  if ((rc = set_split(p, ksplit, partial, (long long)B * p.Hs * p.Ws, "conv_s2_gemm"))) return rc;
  p.pair = want_pair(p, ksplit) ? 1 : 0;
  if (p.pair) p.num_tiles /= 2;
  CUtensorMap ta, tb;
  if ((rc = map_nhwc_parity(&ta, x, B, H, W, Cin, p.tw + p.win, p.th + p.win, p.nb))) return rc;
  if ((rc = map_weights(&tb, wpk, Cout, 16 * Cin, p.pair ? p.BN / 2 : p.BN))) return rc;
  StoreMaps sm{};
  if (p.win) {
    if ((rc = map_store(&sm.out, y, B, p.Hs, p.Ws, ldy, 2, false))) return rc;
    if (p.bwd.dz32 && (rc = map_store(&sm.dz, p.bwd.dz32, B, p.Hs, p.Ws, p.bwd.Cy, 4, false))) return rc;
  }
  return launch(ta, tb, sm, p, stream, stat_rows);
}

// y[B,2Hi,2Wi,ldy] (bf16) = convT4x4/s2/p1(z[B,Hi,Wi,Cin] (pixel stride ldz), wpk[4 phases][Cout][4*Cin]); stats as above.
int b200_convt_s2_gemm(const void* z, int B, int Hi, int Wi, int Cin, int ldz, const void* wpk, int Cout, void* y, int ldy, int bn_hint, int stages,
                       float* stats, int* stat_rows, const b200_epi_bwd* bwd, int ksplit, float* partial, cudaStream_t stream) {
  if (!is_pow2(Hi) || !is_pow2(Wi) || Cin % 64 || Cout % 32 || ldy % 8 || ldz % 8 || B < 1) {
    b200_set_error("convt_s2_gemm: unsupported shape");
    return B200_ERR_ARG;
  }
  IGemmParams p{};
  p.mode = MODE_CONVT;
  p.B = B;
  p.Hs = Hi;
  p.Ws = Wi;
  p.win = (want_window(p.Hs, p.Ws, ksplit) && !(stats && bwd && bwd->y && !bwd->dz32)) ? 1 : 0;
  if (p.win) p.tw = kWinTw, p.th = kWinTh, p.nb = 1;
  else pixel_box(p.Hs, p.Ws, 128, &p.tw, &p.th, &p.nb);
  p.tiles_w = p.Ws / p.tw;
  p.tiles_h = p.Hs / p.th;
  const int tiles_b = (B + p.nb - 1) / p.nb;
  p.num_ptiles = p.tiles_w * p.tiles_h * tiles_b;
  p.C_k = Cin;
  p.cpt = Cin / 64;
  p.kb_per_tile = p.kb_total = 4 * p.cpt;
  // narrow layers (N <= 64: the last decoder level, the data gradients into 64-channel layers): all four phases per tile from ONE
  // 18 x 10 window into four TMEM accumulators — a 256 x 64 pair tile is otherwise bound by the L2 -> SM operand stream
  // (B200_WIN4=0: one phase per work item)
  static const bool win4_enabled = env_int("B200_WIN4", 1) != 0;
  const bool win4 = p.win && win4_enabled && pick_bn(Cout, bn_hint) <= 64;
  p.BN = pick_bn(Cout, bn_hint);
  if (Cout % p.BN || p.BN % 32 || p.BN > 256 || (stats && Cout > 1024)) {
    b200_set_error("convt_s2_gemm: bad N tile");
    return B200_ERR_ARG;
  }
  p.Cout = Cout;
  p.ldc = ldy;
  p.out = y;
  p.stages = stages;
  p.n_tiles = Cout / p.BN;
  p.phases = win4 ? 1 : 4;
  if (win4) p.win = 4;
  p.m_tiles = 1;
  p.num_tiles = p.num_ptiles * p.n_tiles * p.phases;
  p.stats = stats;
  int rc;
  if ((rc = set_bwd(p, bwd, Cout, "convt_s2_gemm"))) return rc;
  if ((rc = set_split(p, ksplit, partial, (long long)B * 4 * Hi * Wi, "convt_s2_gemm"))) return rc;
  p.pair = want_pair(p, ksplit) ? 1 : 0;
  if (p.pair) p.num_tiles /= 2;
  CUtensorMap ta, tb;
  const int halo = p.win == 4 ? 2 : (p.win ? 1 : 0);
  if ((rc = map_nhwc(&ta, z, B, Hi, Wi, Cin, ldz, p.tw + halo, p.th + halo, p.nb))) return rc;
  if ((rc = map_weights(&tb, wpk, 4 * Cout, 4 * Cin, p.pair ? p.BN / 2 : p.BN))) return rc;
  StoreMaps sm{};
  if (p.win) {
    if ((rc = map_store(&sm.out, y, B, 2 * Hi, 2 * Wi, ldy, 2, true))) return rc;
    if (p.bwd.dz32 && (rc = map_store(&sm.dz, p.bwd.dz32, B, 2 * Hi, 2 * Wi, p.bwd.Cy, 4, true))) return rc;
  }
  return launch(ta, tb, sm, p, stream, stat_rows);
}

// y[B,H,W,ldy] (bf16) = relu(conv3x3/s1/p1(x[B,H,W,Cin] (pixel stride ldx), wpk[Cout][9*Cin] bf16 with k = (kh*3+kw)*Cin + ci) + bias)
// (relu == 0: no activation is not offered — every VGG19 convolution is followed by ReLU(inplace), loss.py:61-66).
// pool != 0: y [B,H/2,W/2,ldy] = MaxPool2d(2, 2) of that result, fused into the epilogue (vgg19.features' conv -> ReLU -> MaxPool runs).
int b200_conv3_gemm(const void* x, int B, int H, int W, int Cin, int ldx, const void* wpk, const float* bias, int Cout, void* y, int ldy, int pool,
                    int bn_hint, cudaStream_t stream) {
  if (!is_pow2(H) || !is_pow2(W) || H < kWinTh || W < kWinTw || Cin % 64 || Cout % 32 || ldy % 8 || ldx % 8 || B < 1) {
    b200_set_error("conv3_gemm: needs power-of-two H >= 16, W >= 8, Cin % 64 == 0, Cout % 32 == 0");
    return B200_ERR_ARG;
  }
  IGemmParams p{};
  p.mode = MODE_CONV3;
  p.B = B;
  p.Hs = H;
  p.Ws = W;
  p.win = 3;
  p.tw = kWinTw, p.th = kWinTh, p.nb = 1;
  p.tiles_w = p.Ws / p.tw;
  p.tiles_h = p.Hs / p.th;
  p.num_ptiles = p.tiles_w * p.tiles_h * B;
  p.C_k = Cin;
  p.cpt = Cin / 64;
  p.kb_per_tile = p.kb_total = 9 * p.cpt;
  p.BN = pick_bn(Cout, bn_hint);
  if (Cout % p.BN || p.BN % 32 || p.BN > 256) {
    b200_set_error("conv3_gemm: bad N tile");
    return B200_ERR_ARG;
  }
  p.Cout = Cout;
  p.ldc = ldy;
  p.out = y;
  p.stages = 0;
  p.n_tiles = Cout / p.BN;
  p.phases = 1;
  p.m_tiles = 1;
  p.num_tiles = p.num_ptiles * p.n_tiles;
  p.stats = nullptr;
  p.bias = bias;
  p.pool = pool ? 1 : 0;
  p.bwd = EpiBwd{};
  p.ksplit = 1;
  p.partial = nullptr;
  p.P_total = (long long)B * H * W;
  p.pair = want_pair(p, 1) ? 1 : 0;
  if (p.pair) p.num_tiles /= 2;
  CUtensorMap ta, tb;
  int rc;
  if ((rc = map_nhwc(&ta, x, B, H, W, Cin, ldx, p.tw + 2, p.th + 2, p.nb))) return rc;
  if ((rc = map_weights(&tb, wpk, Cout, 9 * Cin, p.pair ? p.BN / 2 : p.BN))) return rc;
  StoreMaps sm{};
  if ((rc = pool ? map_store(&sm.out, y, B, H / 2, W / 2, ldy, 2, false, true) : map_store(&sm.out, y, B, H, W, ldy, 2, false))) return rc;
  return launch(ta, tb, sm, p, stream, nullptr);
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
  p.bwd = EpiBwd{};
  p.ksplit = 1;
  p.partial = nullptr;
  p.pair = 0;
  CUtensorMap ta, tb;
  int rc;
  if ((rc = map_nhwc(&ta, S, B, Hs, Ws, Cs, lds, p.tw, p.th, p.nb))) return rc;
  if ((rc = map_nhwc_parity(&tb, L, B, 2 * Hs, 2 * Ws, Cl, p.tw, p.th, p.nb))) return rc;
  return launch(ta, tb, StoreMaps{}, p, stream, nullptr);
}

// Splits per image of b200_gram_gemm.  Two reasons to split the pixel sum: (1) the tensor cores add into the fp32 accumulator with
// truncation, so a sum of non-negative products (post-ReLU features) drifts low by ~0.6e-7 per accumulation step — 4096 steps for a
// 256x256 map (measured 1.4e-4 low); bounding a split at 64 pixel tiles (256 steps) keeps that under 2e-5, the partial sums are then
// added in round-to-nearest fp32 by the consumer; (2) B x (C/128) x (C/BN) tiles alone do not fill 148 SMs for the shallow layers.
int b200_gram_splits(int B, int H, int W, int C) {
  if (B < 1 || H < 1 || W < 1 || C < 64 || H * W < 64) return 1;
  const int tiles = H * W / 64;
  const int mn = ((C < 128 ? 128 : C) / 128) * (C / (C >= 256 ? 256 : C));
  int s = 1;
  while (tiles / s > 64) s <<= 1;
  while (s < tiles && tiles / s > 8 && (long long)B * mn * s < kNumSMs) s <<= 1;
  return s;
}

// Per-image Gram matrices of a feature map (lib/models/loss.py:117-136, `gram_matrix` before its division): F [B,H,W,C] bf16
// (pixel stride ldf) -> G [B][S][Cp][C] fp32, sum_s G[b][s][c][c'] = sum_pixels F[b,pix,c] * F[b,pix,c'], S = b200_gram_splits(...),
// Cp = max(C, 128) rows per matrix (rows >= C are zero).  The weight-gradient GEMM without taps: both operands are the same tensor.
int b200_gram_gemm(const void* F, int B, int H, int W, int C, int ldf, float* G, cudaStream_t stream) {
  if (!is_pow2(H) || !is_pow2(W) || H * W < 64 || C % 64 || ldf % 8 || B < 1) {
    b200_set_error("gram_gemm: needs power-of-two H, W with H*W >= 64 and C % 64 == 0");
    return B200_ERR_ARG;
  }
  IGemmParams p{};
  p.mode = MODE_WGRAD;
  p.gram = 1;
  p.B = B;
  p.Hs = H;
  p.Ws = W;
  pixel_box(H, W, 64, &p.tw, &p.th, &p.nb);
  if (p.nb != 1) {
    b200_set_error("gram_gemm: a 64-pixel tile must lie inside one image");
    return B200_ERR_ARG;
  }
  const int S = b200_gram_splits(B, H, W, C);
  p.tiles_w = W / p.tw;
  p.tiles_h = H / p.th;
  p.num_ptiles = p.tiles_w * p.tiles_h * B;
  p.C_k = C;
  p.cpt = C / 64;
  p.BN = C >= 256 ? 256 : C;
  p.kb_total = p.num_ptiles;
  p.kb_per_tile = p.tiles_w * p.tiles_h / S;  // S splits per image, image-major: split index = b * S + s
  p.Cout = C < 128 ? 128 : C;                 // accumulator rows: 128-row MMA tiles (channels beyond C are TMA zero fill)
  p.ldc = C;
  p.out = G;
  p.stages = 0;
  p.m_tiles = p.Cout / 128;
  p.n_tiles = C / p.BN;
  p.phases = 1;
  p.num_tiles = p.m_tiles * p.n_tiles * B * S;
  CUtensorMap ta, tb;
  int rc;
  if ((rc = map_nhwc(&ta, F, B, H, W, C, ldf, p.tw, p.th, p.nb))) return rc;
  if ((rc = map_nhwc(&tb, F, B, H, W, C, ldf, p.tw, p.th, p.nb))) return rc;
  return launch(ta, tb, StoreMaps{}, p, stream, nullptr);
}

int b200_splitk_reduce(const float* partial, int ksplit, long long P, int C, void* out, int ldy, cudaStream_t stream) {
  if (ksplit < 1 || P < 1 || C % 8 || ldy % 8) {
    b200_set_error("splitk_reduce: bad sizes");
    return B200_ERR_ARG;
  }
  long long blocks = (P * (C >> 3) + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  B200_LAUNCH((splitk_reduce_kernel), (int)blocks, 256, 0, stream, partial, ksplit, P, C, (__nv_bfloat16*)out, ldy);
  B200_CHECK_LAUNCH();
  return B200_OK;
}

}  // extern "C"
