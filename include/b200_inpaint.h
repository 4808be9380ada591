/*
 * b200_inpaint.h — C ABI of libb200inpaint.so: the sm_100a kernels behind the GAN-inpainting train step.
 *
 * The reference (abeytheo/gan-inpainting) has no FFI: its hot path sits behind the Python nn.Module protocol
 * (lib/models/networks.py, lib/models/loss.py) and calls ATen ops (cuDNN / MKL-DNN).  Each entry point below
 * replaces the ATen work of the reference call sites it cites (file:line into the reference tree) and is bound
 * from Python with ctypes (gan-inpainting_b200/_lib.py); INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated; `stream` is a cudaStream_t.
 *   - activations are NHWC bf16 ("bf16*" = const void* / void* here), the 1-channel images and all weights,
 *     gradients of weights, statistics and scalars are fp32.
 *   - every function returning int returns 0 on success, otherwise B200_ERR_* and b200_last_error() describes it.
 *   - nothing allocates device memory; scratch buffers are caller-provided (sizes stated per function).
 */
#ifndef B200_INPAINT_H
#define B200_INPAINT_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* b200_stream_t; /* == cudaStream_t */

#define B200_OK 0
#define B200_ERR_ARG 1
#define B200_ERR_CUDA 2
#define B200_ERR_DRIVER 3

/* activation codes */
#define B200_ACT_NONE 0
#define B200_ACT_LEAKY 1 /* LeakyReLU(0.2)  networks.py:287, 338 */
#define B200_ACT_RELU 2  /* ReLU            networks.py:289 */
#define B200_ACT_TANH 3  /* Tanh            networks.py:297 */

/* reconstruction-loss kinds */
#define B200_RECON_L1 0          /* nn.L1Loss()                       wgan_l1.py:62,180 */
#define B200_RECON_RMSE 1        /* RMSELoss                          loss.py:11-19 */
#define B200_RECON_LOCAL_L1 2    /* LocalLoss(nn.L1Loss)              loss.py:24-47 */
#define B200_RECON_LOCAL_MSE 3   /* LocalLoss(nn.MSELoss)             loss.py:24-47 */
#define B200_RECON_MASKED_RMSE 4 /* RMSELoss(mask*y, mask*yhat)       experiment1_global_local_D.py:167 */

/* adversarial heads */
#define B200_ADV_WGAN 0  /* torch.mean(pred)                  wgan_l1.py:137-143,177 */
#define B200_ADV_BCE 1   /* nn.BCELoss vs constant target     minimaxgan_rmse.py:61,135,141,162 */
#define B200_ADV_LSGAN 2 /* nn.MSELoss vs constant target     experiment1_global_local_D.py:119,162-163 */

int b200_abi_version(void);
const char* b200_last_error(void); /* thread-local, host string */
long long b200_launch_count(void);  /* kernels launched by this library since load */
void b200_launch_count_add(long long n); /* a replayed CUDA graph re-launches n kernels this library enqueued at capture time */
/* Dropout(0.5) masks (networks.py:312-314) are a stateless hash of (seed, element index).  With a device counter installed the
 * effective seed of every kernel enqueued afterwards is seed + *dev_counter * odd constant, read when the kernel RUNS: a captured
 * CUDA graph of the whole train step then draws fresh masks on every replay (bump the counter with b200_step_inc).  NULL = off. */
void b200_set_seed_counter(const unsigned* dev_counter);
/* SMs the persistent GEMM kernels may occupy (default: all 148).  The data-parallel trainer lowers it while bucketed gradient
 * all-reduces are in flight (the reference is single-GPU, wgan_l1.py:50-52), so that the collective's CTAs never hold
 * an SM that a statically scheduled persistent CTA was counting on.  n <= 0 restores all SMs. */
void b200_set_sm_limit(int n);

/* ------------------------------------------------------------------------------------------------------------------
 * tcgen05 implicit-GEMM convolutions (csrc/conv_gemm.cu).  4x4 kernel, stride 2, padding 1; TMA-staged NHWC bf16
 * tiles, fp32 accumulation in TMEM; persistent CTAs (one per SM) walk the output tiles with two TMEM accumulators so
 * the epilogue of one tile overlaps the main loop of the next.  H, W powers of two; Cin % 64 == 0; Cout % 32 == 0;
 * bn_hint / stages 0 = auto.
 * `stats` (optional, fprop entry points): the epilogue also accumulates the per-channel sum / sum of squares of the stored
 * bf16 outputs for the nn.BatchNorm2d that follows (networks.py:310, 341-349): one row [2][Cout] per CTA, *stat_rows (host int)
 * receives the number of rows written (<= b200_igemm_stat_rows()); reduce them with b200_bn_finalize.
 * ------------------------------------------------------------------------------------------------------------------ */

/* nn.Conv2d(Cin, Cout, 4, 2, 1, bias=False) forward — networks.py:285-286 (generator d2..d7), :340-348 (PatchGAN c2..c4);
 * also the data gradient of nn.ConvTranspose2d.  x [B,H,W,Cin]; wpk [Cout][16*Cin] from b200_pack_gather;
 * y [B,H/2,W/2,*] with pixel stride ldy. */
/* Optional fused backward-of-activation epilogue of the two data-gradient uses (autograd of networks.py:287-289, 310-314, 338-349):
 * output columns >= col0 are gradients w.r.t. act(dropout(bn(y)))[:, col - col0]; the epilogue stores dz (gradient w.r.t. the
 * BatchNorm output z = y*scale + shift) instead, and with `stats` accumulates (sum dz, sum dz*y) per channel for b200_bn_bwd_fused.
 * LeakyReLU: dz = z > 0 ? g + g2 : 0.2 g (g2 = optional skip-connection gradient); ReLU: dz = z > 0 ? g : 0, with
 * dropout dz = keep ? 2 dz : 0.  y == NULL disables it. */
typedef struct b200_epi_bwd {
  const void* y;      /* forward conv output, bf16, pixel stride ldy */
  int ldy;
  const float* scale; /* forward BatchNorm scale / shift (stats_out + 2C / + 3C); NULL: no norm */
  const float* shift;
  int act;
  const void* g2;
  int ldg2, offg2;
  int dropout;
  unsigned seed;
  int col0;
  int Cy;             /* channels of y; col0 + Cy == Cout */
  void* dz32;         /* optional (with stats): dz stored unrounded as fp32 [pixel][Cy] for b200_bn_bwd_fused instead of bf16 in place */
} b200_epi_bwd;

int b200_igemm_stat_rows(void);
int b200_conv_s2_gemm(const void* x, int B, int H, int W, int Cin, const void* wpk, int Cout, void* y, int ldy, int bn_hint, int stages, float* stats,
                      int* stat_rows, const b200_epi_bwd* bwd, int ksplit, float* partial, b200_stream_t stream);

/* nn.ConvTranspose2d(Cin, Cout, 4, 2, 1, bias=False) forward as 4 sub-pixel phases — networks.py:300-309 (u7..u2);
 * also the data gradient of nn.Conv2d.  z [B,Hi,Wi,*] with pixel stride ldz; wpk [4][Cout][4*Cin] from b200_pack_phase;
 * y [B,2Hi,2Wi,*] with pixel stride ldy. */
int b200_convt_s2_gemm(const void* z, int B, int Hi, int Wi, int Cin, int ldz, const void* wpk, int Cout, void* y, int ldy, int bn_hint, int stages,
                       float* stats, int* stat_rows, const b200_epi_bwd* bwd, int ksplit, float* partial, b200_stream_t stream);
/* Split-K (the innermost layers: a handful of output tiles, K up to 8192): with ksplit > 1 the two entry points above write fp32
 * partial tiles into `partial` [ksplit][pixels][Cout] (no fused epilogue, y untouched) and this kernel sums them into y (bf16). */
int b200_splitk_reduce(const float* partial, int ksplit, long long P, int C, void* y, int ldy, b200_stream_t stream);

/* Weight gradient of both layer kinds (autograd of the call sites above; wgan_l1.py:138,141,183):
 *   partial[s][cs][(kh*4+kw)*Cl + cl] = sum over the pixels of split s of S[b,i,j,cs] * L[b,2i+kh-1,2j+kw-1,cl]
 * S [B,Hs,Ws,*] (pixel stride lds, Cs % 128 == 0) is the low-resolution tensor (Conv2d: dy; ConvTranspose2d: input),
 * L [B,2Hs,2Ws,Cl] dense the high-resolution one.  *splits_io: requested split count in, used count out.
 * partial: (*splits_io) * Cs * 16 * Cl floats.  Reduce with b200_wgrad_reduce. */
int b200_wgrad_gemm(const void* S, int B, int Hs, int Ws, int Cs, int lds, const void* L, int Cl, void* partial, int* splits_io, int bn_hint, int stages,
                    b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * weight layouts (csrc/pack.cu).  w is the fp32 checkpoint tensor [D0][D1][4][4]
 * (Conv2d: D0 = Cout, D1 = Cin; ConvTranspose2d: D0 = Cin, D1 = Cout — SURVEY.md §2.1 state_dict contract).
 * ------------------------------------------------------------------------------------------------------------------ */
int b200_pack_gather(const float* w, int D0, int D1, void* out /* bf16 [D0][16*D1] */, b200_stream_t stream);
int b200_pack_phase(const float* w, int D0, int D1, void* out /* bf16 [4][D1][4*D0] */, b200_stream_t stream);
/* both layouts of every 4x4 layer of a network in one launch: table = device array of nlayers x {src, gather, phase, D0, D1, tile0}
 * (six 64-bit words each; D0 % 16 == 0, D1 % 32 == 0; tile0 = number of 16x32 tiles of the layers before) */
int b200_pack_all(const void* table, int nlayers, int total_tiles, b200_stream_t stream);
/* grad[cs][cl][kh][kw] (+)= sum_s partial[s][cs][(kh*4+kw)*Cl + cl]   (deterministic split reduction) */
int b200_wgrad_reduce(const float* partial, int splits, int Cs, int Cl, float* grad, int accumulate, b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * BatchNorm2d (training / eval) + activation (csrc/norm_act.cu) — nn.BatchNorm2d networks.py:38, 341-349;
 * LeakyReLU / ReLU prologues and the post-LeakyReLU skip tensor networks.py:287-289, 320-324;
 * Dropout(0.5) networks.py:313-314.
 * ------------------------------------------------------------------------------------------------------------------ */
int b200_bn_stat_blocks(long long P, int C); /* partial buffer holds b200_bn_stat_blocks(P, C) * 2 * C floats */

/* batch statistics of y [P][C] bf16 -> stats_out[4][C] = mean, rstd, scale = gamma*rstd, shift = beta - mean*scale;
 * running_mean / running_var / num_batches_tracked (int64) updated in place when non-null (momentum, unbiased var). */
int b200_bn_stats(const void* y, long long P, int C, void* partial, const float* gamma, const float* beta, float* running_mean, float* running_var,
                  long long* num_batches_tracked, float momentum, float eps, float* stats_out, b200_stream_t stream);
/* same outputs from the partial rows the convolution epilogue wrote (see `stats` above); P = pixels per channel */
int b200_bn_finalize(const float* partial, int rows, long long P, int C, const float* gamma, const float* beta, float* running_mean, float* running_var,
                     long long* num_batches_tracked, float momentum, float eps, float* stats_out, b200_stream_t stream);
int b200_bn_eval_coeffs(int C, const float* gamma, const float* beta, const float* running_mean, const float* running_var, float eps, float* stats_out,
                        b200_stream_t stream);
/* out1[pix*ld1 + off1 + c] = act(dropout(y*scale + shift)); out2 = relu(out1) (the concat half).  scale null = identity. */
int b200_bn_act(const void* y, long long P, int C, const float* scale, const float* shift, int act, void* out1, int ld1, int off1, void* out2, int ld2,
                int off2, int dropout, unsigned seed, b200_stream_t stream);
/* backward of (BN ->) [dropout ->] act with up to two incoming gradients (g2 = skip gradient through ReLU);
 * stats non-null = training-mode BN (dgamma / dbeta written or accumulated); coef: 3*C floats scratch. */
int b200_bn_act_bwd(const void* y, long long P, int C, const float* stats, const float* gamma, int act, const void* g1, int ldg1, int offg1, const void* g2,
                    int ldg2, int offg2, int dropout, unsigned seed, void* partial, float* coef, float* dgamma, float* dbeta, int accumulate, void* dy,
                    b200_stream_t stream);

/* BatchNorm backward after a data-gradient GEMM with a b200_epi_bwd epilogue: dz [P][C] (bf16, or fp32 when dz_is_f32; pixel stride lddz, channel offset offdz)
 * and the GEMM's partial rows (element (row, k, c) at partial[row*row_stride + k*k_stride + c]) -> dgamma, dbeta, dy. */
int b200_bn_bwd_fused(const void* y, long long P, int C, const float* stats, const float* gamma, const void* dz, int lddz, int offdz, int dz_is_f32,
                      const float* partial, int rows, int row_stride, int k_stride, float* coef, float* dgamma, float* dbeta, int accumulate, void* dy,
                      b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * layers with one channel on one side + the discriminator head (csrc/conv_direct.cu)
 * ------------------------------------------------------------------------------------------------------------------ */
/* nn.Conv2d(1, C, 4, 2, 1) forward — networks.py:296 (d1), :337 (c1) — x [B,H,W] fp32 (times xmul if non-null: the
 * local discriminator's mask, experiment1_global_local_D.py:157); y [B,H/2,W/2,C] bf16; y2 = relu copy (concat half).
 * Also the data gradient of the last ConvTranspose2d(128, 1). */
int b200_conv_c1_fwd(const float* x, const float* xmul, int B, int H, int W, const float* w, int C, int act, void* y, void* y2, int ld2, int off2,
                     b200_stream_t stream);
/* nn.ConvTranspose2d(C, 1, 4, 2, 1) (+bias, Tanh) forward — networks.py:293-297 (u1) — z [B,H,W,*] bf16 -> y [B,2H,2W] fp32;
 * comp non-null: comp = base + y*mul, the composite `masked + G(masked)*mask` of wgan_l1.py:122.
 * Also the data gradient of c1 (then y *= mul if mul non-null). */
int b200_convt_c1_fwd(const void* z, int B, int H, int W, int C, int ldz, const float* w, const float* bias, int act, const float* base, const float* mul,
                      float* y, float* comp, b200_stream_t stream);
int b200_wgrad_c1_blocks(void); /* partial: b200_wgrad_c1_blocks() * C * 16 floats */
/* dw[c][kh*4+kw] (+)= sum z[b,i,j,c] * g[b,2i+kh-1,2j+kw-1] (* gmul) — weight gradient of both kinds above */
int b200_wgrad_c1(const void* z, int B, int H, int W, int C, int ldz, const float* g, const float* gmul, float* partial, float* dw, int accumulate,
                  b200_stream_t stream);
/* Conv2d(512,1,4,1,0) -> Flatten -> Linear(n*n,1) [-> Sigmoid] — networks.py:352-357 — as one dot product per image with
 * E[h][w][c] = sum_{kh,kw} w5[c][kh][kw] * wl[(h-kh)*n + (w-kw)], n = S-3. */
int b200_head_weight(const float* w5, const float* wl, int S, int C, float* E, b200_stream_t stream);
int b200_head_fwd(const void* a, const float* E, int B, int n, const float* bias, int sigmoid, float* scratch /* 8*B */, float* lin, float* pred,
                  b200_stream_t stream);
int b200_head_dlin(const float* dpred, const float* pred, int sigmoid, int B, float* dlin, b200_stream_t stream);
int b200_head_bwd_input(const float* dlin, const float* E, int B, int n, void* da, b200_stream_t stream);
int b200_head_bwd_params(const float* dlin, const void* a, int B, int S, int C, const float* w5, const float* wl, float* dE /* S*S*C */, float* dw5,
                         float* dwl, float* dbias, int accumulate, b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * masks, compositing, losses (csrc/losses.cu)
 * ------------------------------------------------------------------------------------------------------------------ */
int b200_recon_scratch_floats(void);
/* mask_out = do_ceil ? ceil(mask) : mask; masked = ground * (1 - mask_out) — wgan_l1.py:114,117 */
int b200_mask_prepare(const float* ground, const float* mask, long long n, int do_ceil, float* mask_out, float* masked, b200_stream_t stream);
/* out = base + g * mask — wgan_l1.py:122 */
int b200_composite(const float* base, const float* g, const float* mask, long long n, float* out, b200_stream_t stream);
int b200_mul(const float* a, const float* b, long long n, float* out, b200_stream_t stream);
/* out[0] = loss, out[1] = denominator (float), out[2..3] = exact denominator as two 31-bit int words */
int b200_recon_loss(int kind, const float* y, const float* yhat, const float* mask, long long n, float eps, float* scratch, float* out,
                    b200_stream_t stream);
/* seed of the generator backward: d_inp = dadv1 + dadv2 + gscale*(lambda1*recon1' + lambda2*recon2');
 * dpre = (composite ? d_inp*mask : d_inp) * (1 - g^2); dbias (+)= sum dpre.  kind < 0 disables a term. */
int b200_gstep_seed(const float* ground, const float* inp, const float* g, const float* mask, const float* dadv1, const float* dadv2, int kind1, float lambda1,
                    const float* loss1, int kind2, float lambda2, const float* loss2, const float* gscale, int composite, long long n, float* dpre, float* dinp,
                    float* scratch, float* dbias, int accumulate, b200_stream_t stream);
/* loss_out[0] = head(pred[B]); dpred[b] = scale * d loss / d pred[b] */
int b200_adv_loss(const float* pred, int B, int mode, float target, float scale, float* loss_out, float* dpred, b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * optimizers (csrc/optim.cu) — optim.RMSprop wgan_l1.py:64-65 (+ clamp :152-153), optim.Adam minimaxgan_rmse.py:64-65,
 * p.grad.abs().mean() logging wgan_l1.py:195-197
 * ------------------------------------------------------------------------------------------------------------------ */
int b200_rmsprop_step(float* p, const float* g, float* sq, long long n, float lr, float alpha, float eps, int clamp, float lo, float hi, b200_stream_t stream);
int b200_step_inc(int* step_count, b200_stream_t stream);
int b200_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, const int* step_count,
                   b200_stream_t stream);
int b200_clamp(float* p, long long n, float lo, float hi, b200_stream_t stream);
/* Gradient buckets on the wire in bf16 (data-parallel all-reduce, optional; the reference is single-GPU):
 * to_bf16 != 0: dst (bf16) = src (fp32), else dst (fp32) = src (bf16).  n % 4 == 0, 16-byte aligned pointers. */
int b200_cast_grads(const void* src, void* dst, long long n, int to_bf16, b200_stream_t stream);
/* out[t] = mean |g[offsets[t] .. offsets[t+1])|, t < ntensors (offsets: int64[ntensors+1]) */
int b200_absmean(const float* g, const long long* offsets, int ntensors, float* out, b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * SSIM metric (csrc/ssim.cu) — lib/pytorch_ssim/__init__.py:20-40 (`_ssim`), logged per batch at
 * experiment1_global_local_D.py:209.  img1, img2: fp32 [N,C,H,W]; window11: HOST pointer to the 11 taps of the normalised 1-D
 * Gaussian (pytorch_ssim gaussian(11, 1.5)); zero padding 5; C1 = 0.01^2, C2 = 0.03^2.  partial: b200_ssim_partials() floats of
 * scratch; out: 1 float (size_average: mean over everything) or N floats (mean over C,H,W per image).
 * ------------------------------------------------------------------------------------------------------------------ */
long long b200_ssim_partials(int N, int C, int H, int W);
int b200_ssim(const float* img1, const float* img2, int N, int C, int H, int W, const float* window11, int size_average, float* partial, float* out,
              b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Total-variation loss (csrc/tv.cu) — lib/models/loss.py:138-151 (`tv_loss`), part of g_loss at
 * experiment_list/wgan_perceptual_style_faceparsing.py:219-221.  img: fp32 [planes = N*C, H, W];
 * out[0] = weight * (mean of squared horizontal differences + mean of squared vertical differences).
 * partial: 2 * planes * H floats of scratch.  _bwd: dimg (+)= gscale[0] * d loss / d img (gscale: device scalar, NULL = 1).
 * ------------------------------------------------------------------------------------------------------------------ */
int b200_tv_loss(const float* img, long long planes, int H, int W, float weight, float* partial, float* out, b200_stream_t stream);
int b200_tv_loss_bwd(const float* img, long long planes, int H, int W, float weight, const float* gscale, float* dimg, int accumulate,
                     b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * VGG19 feature losses of config 4 (csrc/conv_gemm.cu MODE_CONV3 / gram mode, csrc/vgg.cu) — lib/models/loss.py:49-136
 * (`perceptual_loss`, `style_loss`, `perceptual_and_style_loss`, `gram_matrix`) over networks.py:367-389 (`VGG19Wrapper`), called at
 * experiment_list/wgan_perceptual_style_faceparsing.py:216.  Forward only: the reference evaluates them under torch.no_grad() on
 * detached tensors.
 * ------------------------------------------------------------------------------------------------------------------ */
/* nn.Conv2d(Cin, Cout, 3, 1, 1) + ReLU(inplace) of vgg19.features (window operand staging: the nine taps are descriptors into one
 * 18 x 10-pixel halo window per 64-channel block).  x [B,H,W,Cin] bf16 (pixel stride ldx), wpk [Cout][9*Cin] from b200_pack_conv3,
 * bias [Cout] fp32 or NULL; y [B,H,W,*] (pixel stride ldy).  H >= 16, W >= 8, powers of two; Cin % 64 == 0; Cout % 32 == 0.
 * pool != 0: the nn.MaxPool2d(2, 2) that follows the ReLU is fused into the epilogue and y is [B,H/2,W/2,*] — the full-resolution
 * output never reaches memory. */
int b200_conv3_gemm(const void* x, int B, int H, int W, int Cin, int ldx, const void* wpk, const float* bias, int Cout, void* y, int ldy, int pool,
                    int bn_hint, b200_stream_t stream);
/* fp32 [Cout][Cin][3][3] (torch layout) -> bf16 [Cout][(kh*3+kw)*Cin + ci] */
int b200_pack_conv3(const float* w, int Cout, int Cin, void* out, b200_stream_t stream);
/* vgg19.features[0..1] on `img.repeat(1,3,1,1)` (loss.py:55-56): x fp32 [B,H,W], w1 [64][9] = the filters summed over their 3 input
 * channels, bias [64] -> y [B,H,W,64] bf16 = relu(conv + bias). */
int b200_conv3_c1_fwd(const float* x, int B, int H, int W, const float* w1, const float* bias, int C, void* y, b200_stream_t stream);
/* nn.MaxPool2d(2, 2) on NHWC bf16 */
int b200_maxpool2(const void* x, int B, int H, int W, int C, void* y, b200_stream_t stream);
/* gram_matrix (loss.py:117-136) before its division: F [B,H,W,C] bf16 -> G [B][S][max(C,128)][C] fp32 with
 * sum_s G[b][s][c][c'] = sum_pix F[b,pix,c] F[b,pix,c'] and S = b200_gram_splits(B,H,W,C) partial sums per image (short tensor-core
 * accumulation chains: see csrc/conv_gemm.cu; the consumer adds the S partials). */
int b200_gram_splits(int B, int H, int W, int C);
int b200_gram_gemm(const void* F, int B, int H, int W, int C, int ldf, float* G, b200_stream_t stream);
/* out[0] (+)= scale * sum (a - b)^2: two bf16 tensors (torch.mean(torch.pow(f_o - f_t, 2)), loss.py:65) / two Gram stacks
 * [B][S][Cp][C] from b200_gram_gemm, their S partials added and scaled by gscale first (loss.py:85).  partial: b200_sqdiff_blocks()
 * floats. */
int b200_sqdiff_blocks(void);
int b200_sqdiff_bf16(const void* a, const void* b, long long n, double scale, float* partial, float* out, int accumulate, b200_stream_t stream);
int b200_sqdiff_gram(const float* ga, const float* gb, int B, int S, int Cp, int C, float gscale, double scale, float* partial, float* out, int accumulate,
                     b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Input pipeline (csrc/preprocess.cu) — what the reference does to every decoded image on the host, for a batch on the GPU:
 * lib/data/dataset.py:6-12 (`Image.open(f).convert('L')`) and train.py:69-72 (`transforms.Resize(size)` + `transforms.ToTensor()`).
 * src: uint8 [B,H,W,channels] (3 = RGB, 1 = L); out: fp32 [B,1,outH,outW] in [0,1]; tmp: B*H*outW bytes.  Bit-exact with
 * Pillow + torchvision: L = (R*19595 + G*38470 + B*7471 + 0x8000) >> 16; separable triangle-filter resample in 22-bit fixed point,
 * horizontal pass first with an 8-bit intermediate; division by 255.f.  (xmin, xcnt, xk[outW][ksize_x]) and
 * (ymin, ycnt, yk[outH][ksize_y]) are device arrays: first source index, tap count and fixed-point weights per output column / row
 * (Pillow's precompute_coeffs + normalize_coeffs_8bpc; gan-inpainting_b200/data.py computes them).
 * ------------------------------------------------------------------------------------------------------------------ */
int b200_preprocess(const void* src, int B, int H, int W, int channels, int outH, int outW, const int* xmin, const int* xcnt, const int* xk, int ksize_x,
                    const int* ymin, const int* ycnt, const int* yk, int ksize_y, void* tmp, float* out, b200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_INPAINT_H */
