/*
 * stfb200.h -- C ABI of libstfb200.so: hand-written sm_100a kernels for the STF-Unet hot path.
 *
 * The reference (XiangFeng-Wen/STF-Unet) is pure Python: every operator on the hot path is a
 * torch / torchvision library call and there is no FFI to mirror.  Each entry point below
 * therefore cites the reference *call site* whose library operator it replaces
 * (paths are relative to the reference checkout).
 *
 * Conventions
 *   - every function returns 0 on success, a negative STFB_E* code otherwise; it never throws,
 *     never exits, never allocates or frees device memory, never synchronises the device;
 *     stfb_last_error() returns a thread-local message for the last failure.
 *   - all pointers are device pointers owned by the caller unless the name says `host`.
 *   - `stream` is a cudaStream_t passed as void*; kernels are enqueued on it and are safe for
 *     CUDA-graph capture (exception: none).
 *   - activations are NHWC ("pixel rows x channels"); dtype codes: 0 = fp32, 1 = bf16.
 *     Accumulation is always fp32 (fp64 for BatchNorm / loss reductions).
 *   - there is NO CPU fallback: without a CUDA device the launch functions return STFB_ENODEV.
 */
#ifndef STFB200_H_
#define STFB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STFB_OK 0
#define STFB_EINVAL (-1)   /* bad shape / dtype / alignment / null pointer      */
#define STFB_ECUDA (-2)    /* CUDA runtime error at launch                        */
#define STFB_ENODEV (-3)   /* no sm_100 device                                    */
#define STFB_ENOTSUP (-4)  /* shape not supported by the requested kernel family  */

#define STFB_F32 0
#define STFB_BF16 1
/* Split-precision operand ("bf16 x 3"): an fp32 tensor [rows][C] stored as three bf16 planes side by side on the channel
 * axis, [rows][3C] = [hi | mid | lo] with hi = bf16(x), mid = bf16(x - hi), lo = bf16(x - hi - mid) -- 24 mantissa bits,
 * written by stfb_split_bf16x3.  The tcgen05 convolution / weight-gradient families take it as x_dtype / dtype (channel
 * counts in the call stay the LOGICAL ones) and accumulate the six products lo*hi, hi*lo, mid*mid, mid*hi, hi*mid, hi*hi
 * (smallest first) in one fp32 TMEM accumulator at 1/6 of the bf16 tensor rate, still several times the FFMA rate of the
 * fp32 SIMT family.  Accuracy: the dropped terms are <= 2^-25 relative; what remains is the rounding of the tensor core's
 * fp32 accumulator over the hi*hi chain (toward zero, once per K = 16), ~1e-6 relative per layer (tests/test_split_gpu.py). */
#define STFB_BF16X3 2

/* conv gather modes */
#define STFB_CONV_FWD 0        /* iy = oy*stride - pad + ky                       (Conv2d forward, ConvTranspose2d dgrad) */
#define STFB_CONV_TRANSPOSED 1 /* iy = (oy + pad - ky)/stride when divisible      (ConvTranspose2d forward, Conv2d dgrad) */

/* conv kernel families (stfb_conv_params.impl) */
#define STFB_IMPL_AUTO 0       /* == SIMT (the families take different weight packings; ask _supported first) */
#define STFB_IMPL_SIMT 1       /* fp32-FFMA implicit GEMM (fp32-accurate mode and odd shapes) */
#define STFB_IMPL_TCGEN05 2    /* bf16 tcgen05/TMEM implicit GEMM fed by TMA                  */

int stfb_version(void);
const char* stfb_last_error(void);
/* number of kernels launched through this library by the calling process (bench "gpu_launches") */
unsigned long long stfb_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution:  y[m, co] = epilogue( sum_{ky,kx,ci} X[gather(m,ky,kx), ci] * Wp[(ky,kx,ci), co] )
 *   m = (n, oy, ox) output pixel; X is the channel concat of x (C1 channels) and x2 (C2, may be 0/NULL).
 *   epilogue: v += bias[co] + bias2[co]; v = v*scale[co] + shift[co]; v += residual[m, co]; relu.
 * Replaces: nn.Conv2d (src/stf_lstm_unet.py:13,16,46,105,137 ; src/unet.py:12,15,37 ; torchvision
 *   BasicBlock convs behind src/stf_lstm_unet.py:111-114), torch.cat + conv (src/stf_lstm_unet.py:60-63,
 *   src/unet.py:48-54), nn.ConvTranspose2d (src/stf_lstm_unet.py:43,135 ; src/unet.py:28-34) via
 *   STFB_CONV_TRANSPOSED, the nn.LSTM gate GEMMs (src/stf_lstm_unet.py:124-127) as 1x1 convs, the folded
 *   eval-mode nn.BatchNorm2d + ReLU + residual add (src/stf_lstm_unet.py:29-35), and every dgrad of these.
 * ---------------------------------------------------------------------------------------------- */
typedef struct stfb_conv_params {
  const void* x;        /* [N, H, W, C1]  x_dtype                                  */
  const void* x2;       /* [N, H, W, C2]  or NULL                                  */
  const void* w;        /* SIMT: packed [kh*kw*(C1+C2)][ldw]; TCGEN05: packed [Cout][ldw] (stfb_pack_weight_ex n_major=1) */
  void* y;              /* [N, Ho, Wo, Cout] y_dtype                               */
  const float* bias;    /* [Cout] or NULL                                          */
  const float* bias2;   /* [Cout] or NULL                                          */
  const float* scale;   /* [Cout] or NULL (folded BatchNorm)                       */
  const float* shift;   /* [Cout] or NULL                                          */
  const void* residual; /* [N, Ho, Wo, Cout] y_dtype or NULL (may alias y)         */
  int N, H, W, C1, C2;
  int Ho, Wo, Cout;
  int kh, kw, stride, pad;
  int ldw;              /* row stride (elements) of w: >= Cout (SIMT) / >= kh*kw*(C1+C2) (TCGEN05) */
  int mode;             /* STFB_CONV_FWD / STFB_CONV_TRANSPOSED                    */
  int relu;
  int x_dtype, y_dtype; /* (f32,f32) (bf16,bf16) (bf16,f32) (bf16x3,f32: x / x2 are [N,H,W,3*C1] / [..,3*C2],
                           w from stfb_pack_weight_split with ldw >= kh*kw*6*(C1+C2); tcgen05 family only) */
  int impl;             /* STFB_IMPL_*                                             */
  /* Fused train-mode BatchNorm statistics of the OUTPUT (tcgen05 family, bf16 y, no epilogue extras): the epilogue adds
   * per-channel sum / sum of squares of the (bf16-rounded) outputs of image group g = n / (N / stat_groups) into
   * stat_partial[slot][0|1][g][c] (fp32, red.add; slot = CTA % stat_slots), the layout stfb_bn_finalize_train reads with
   * nblk = stat_slots.  The caller zeroes the buffer.  NULL = off.  Use stfb_conv2d_stats_fusable() first. */
  float* stat_partial;
  int stat_slots, stat_groups;
} stfb_conv_params;

int stfb_conv2d(const stfb_conv_params* p, void* stream);
/* 1 if the tcgen05 kernel family supports this problem (host-only check, no launch). */
int stfb_conv2d_tcgen05_supported(const stfb_conv_params* p);
/* 1 if, in addition, the launch can produce the BatchNorm statistics of its output (every tile lies inside one of the
 * `groups` image groups, 128-byte output rows). */
int stfb_conv2d_stats_fusable(const stfb_conv_params* p, int groups);

/* Weight gradient:  dW[cp][cg_off + cg][ky][kx] += sum_pix P[pix, cp] * G[gather(pix,ky,kx), cg]
 *   P = per-pixel tensor [N, Hp, Wp, Cp] (dy of a Conv2d; x of a ConvTranspose2d),
 *   G = gathered tensor  [N, Hg, Wg, Cg] (x of a Conv2d;  dy of a ConvTranspose2d), gather iy = py*stride - pad + ky.
 *   dW is fp32 in the reference parameter layout ([Cout,Cin,kh,kw] for Conv2d, [Cin,Cout,kh,kw] for
 *   ConvTranspose2d, [4C,C] for the LSTM matrices) and is ACCUMULATED into (caller zeroes it).
 * Replaces: autograd of the operators listed above (loss.backward(), train_utils/train_and_eval.py:397-404).
 * dtype STFB_BF16X3: P and G are split-precision operands ([.., 3*Cp] / [.., 3*Cg] bf16 from stfb_split_bf16x3; Cp / Cg stay
 *   the logical counts): fp32-accurate weight gradient on the tensor cores, tcgen05 family only (impl != STFB_IMPL_SIMT).
 * impl: STFB_IMPL_AUTO picks the tcgen05 family when the shape allows (bf16, stride-1 "same" geometry, Cg % 64 == 0,
 * Cp % 64 == 0), else the SIMT family; STFB_IMPL_SIMT / STFB_IMPL_TCGEN05 force one.
 * The tcgen05 family needs a caller-provided fp32 workspace (the [(ky,kx,ci)][co] accumulation buffer the split-K CTAs
 * reduce into); stfb_conv2d_wgrad_workspace_bytes returns its size (0 when the SIMT family will run).
 * Deferred mode (dW == NULL, tcgen05 only): the launch only accumulates into `workspace`, laid out
 * [(ky,kx, ci over cg_total)][Cp] and zeroed by the caller; one stfb_wgrad_scatter_batched at the end of the backward pass
 * folds every such buffer into the flat gradient (instead of a memset + transpose launch per weight). */
size_t stfb_conv2d_wgrad_workspace_bytes(const void* P, const void* G, int N, int Hp, int Wp, int Cp, int Hg, int Wg, int Cg,
                                         int cg_total, int kh, int kw, int stride, int pad, int dtype, int impl);
/* Optional caller-owned scratch for the tcgen05 weight-gradient kernels (per-split partial tiles, summed by a second
 * kernel, instead of fp32 atomics from every CTA).  Register a multiple of stfb_wgrad_scratch_bytes() bytes once per
 * process (one process drives one GPU); NULL unregisters.  The buffer must stay alive.  It is cut into slots of
 * stfb_wgrad_scratch_bytes() and every STREAM that launches such a weight gradient owns one slot (first come, first
 * served, up to 16), so launches on different streams never share partial tiles; a stream without a slot uses atomics. */
size_t stfb_wgrad_scratch_bytes(void);
int stfb_set_wgrad_scratch(void* scratch, size_t bytes);
typedef struct stfb_scatter_job {
  long long start;   /* first tile (32 ci x 32 co, all taps) of this job in [0, total_tiles); khw <= 9 */
  long long off;     /* element offset of the weight in BOTH flat buffers (gradient and accumulation) */
  int Cp, Cg, khw, pad_;
} stfb_scatter_job;
/* grad_flat[off + (co*Cg + ci)*khw + tap] += acc_flat[off + (tap*Cg + ci)*Cp + co] for every job (device job table) */
int stfb_wgrad_scatter_batched(const stfb_scatter_job* jobs_dev, int njobs, long long total_tiles, const float* acc_flat,
                               float* grad_flat, void* stream);
int stfb_conv2d_wgrad(const void* P, const void* G, float* dW, int N, int Hp, int Wp, int Cp, int Hg, int Wg, int Cg,
                      int cg_off, int cg_total, int kh, int kw, int stride, int pad, int dtype, int impl,
                      void* workspace, size_t ws_bytes, void* stream);
int stfb_conv2d_wgrad_tcgen05_supported(const void* P, const void* G, int N, int Hp, int Wp, int Cp, int Hg, int Wg, int Cg,
                                        int kh, int kw, int stride, int pad, int dtype);

/* Parameter layout [D0][D1][kh][kw] fp32 -> GEMM layout [(ky,kx,k)][n] in dtype.
 *   k_is_dim1 = 1: k = D1 index, n = D0 index (Conv2d forward, ConvTranspose2d dgrad, LSTM x @ W^T)
 *   k_is_dim1 = 0: k = D0 index, n = D1 index (Conv2d dgrad, ConvTranspose2d forward, LSTM dgates @ W) */
int stfb_pack_weight(const float* w, void* wp, int D0, int D1, int kh, int kw, int k_is_dim1, int dtype, void* stream);
/* Extended form.  n_major = 1 writes [n][(ky,kx,k)] (each output channel's K contiguous: the K-major B operand of the
 * tcgen05 family, row stride ldw = kh*kw*K).  flip = 1 mirrors the taps (ky,kx) -> (kh-1-ky, kw-1-kx), which turns the
 * dgrad of a stride-1 "same" convolution into a forward convolution over dy. */
int stfb_pack_weight_ex(const float* w, void* wp, int D0, int D1, int kh, int kw, int k_is_dim1, int n_major, int flip,
                        int ld /* n_major row stride, 0 = dense; the caller zero-fills any padding */,
                        int gate_c /* > 0: LSTM [4C][C] matrix, row gate*C + u goes to (u/16)*64 + gate*16 + u%16: each 256-row
                                      block holds gates i,f,g,o of the same 64 hidden units (for stfb_lstm_step_fused) */,
                        int dtype, void* stream);

/* Split-precision forms (STFB_BF16X3).
 * stfb_split_bf16x3: x [rows][C] fp32 -> y [rows][3C] bf16 = [hi | mid | lo]; C % 8 == 0, 16-byte aligned pointers.
 * stfb_pack_weight_split: the B operand of a STFB_BF16X3 convolution, always n_major: wp [n][kh*kw*6*K] bf16.  Per filter tap
 *   the K axis holds six segments of all K source channels (x then x2 for a concat convolution),
 *   [w_hi, w_lo, w_mid, w_hi, w_mid, w_hi] -- the partners of the activation planes [lo, hi, mid, mid, hi, hi] the kernels
 *   walk.  k_is_dim1 / flip as in stfb_pack_weight_ex. */
int stfb_split_bf16x3(const float* x, void* y, long long rows, int C, void* stream);
int stfb_pack_weight_split(const float* w, void* wp, int D0, int D1, int kh, int kw, int k_is_dim1, int flip, void* stream);

/* Batched form: one launch packs every weight of a step.  `jobs_dev` is a DEVICE array (uploaded once per model);
 * job j covers work items [start_j, start_{j+1}) of `total`, one work item = one (d0, d1) position with all its kh*kw
 * taps (so a job has D0*D1 items); fields as in stfb_pack_weight_ex.  flip bit 0 = mirror the taps; flip bit 1 (value 2) = the job
 * writes a split-precision operand (stfb_pack_weight_split layout, bf16, n_major = 1, ld = kh*kw*6*K) whatever `dtype` says;
 * flip bit 2 (value 4) = vector job: one work item covers EIGHT consecutive k of one n (the job has D0*D1/8 items; needs
 * n_major = 1, K % 8 == 0, ld % 8 == 0 and a 16-byte aligned dst). */
typedef struct stfb_pack_job {
  const float* src;
  void* dst;
  long long start;
  int D0, D1, khw, k_is_dim1, n_major, flip, ld, pad_;
} stfb_pack_job;
int stfb_pack_weights_batched(const stfb_pack_job* jobs_dev, int njobs, long long total, int dtype, void* stream);

/* Small-channel convolutions (7x7/2 stem with Cin = 1, src/stf_lstm_unet.py:105,177; UNet enc1.0, src/unet.py:20) on
 * the tensor cores: out[N,Ho,Wo,Kpad] (bf16) = im2col of x with K order (ky,kx,ci), zero-padded to Kpad (multiple of 64),
 * after which the conv is a 1x1 GEMM over `out` and its weight gradient a 1x1 wgrad. */
int stfb_im2col_small(const void* x, void* out, int N, int H, int W, int Cin, int Ho, int Wo, int k, int stride, int pad,
                      int Kpad, void* stream);
/* dW[co][ci][ky][kx] += src[co][(ky,kx,ci)] (src rows ld_src wide): folds the im2col weight gradient back. */
int stfb_unpad_wgrad(float* dW, const float* src, int Cout, int Cin, int kh, int kw, int ld_src, void* stream);

/* ------------------------------------------------------------------------------------------------
 * BatchNorm2d (eps, momentum as nn.BatchNorm2d defaults 1e-5 / 0.1; torchvision BasicBlock bn1/bn2,
 * src/stf_lstm_unet.py:14,17,108 ; src/unet.py:13,16).  Rows are grouped: group g = rows [g*R, (g+1)*R);
 * the STF encoder runs with G = T groups because the reference calls each BN once per time step
 * (src/stf_lstm_unet.py:168-186) -- statistics are per group, running stats get G sequential updates.
 * ---------------------------------------------------------------------------------------------- */
/* Statistics are reduced in two steps without atomics: every CTA of the reduction writes its partial sums
 * partial[blk][0][g][c] = sum x, partial[blk][1][g][c] = sum x^2 (fp32), and the finalize kernel adds the nblk slots in
 * fp64.  nblk = stfb_bn_partial_blocks(G, R) (host-only, deterministic); the caller allocates nblk*2*G*C floats. */
int stfb_bn_partial_blocks(int G, long long R);
int stfb_bn_stats(const void* x, float* partial, int nblk, int G, long long R, int C, int dtype, void* stream);
/* train: batch stats -> scale/shift/mean/invstd [G][C]; running_mean/var updated G times in place (sequentially, one
 * update per group = per reference BN call), num_batches_tracked += G. */
int stfb_bn_finalize_train(const float* partial, int nblk, const float* gamma, const float* beta, float* running_mean,
                           float* running_var, long long* num_batches_tracked, float* scale, float* shift,
                           float* mean, float* invstd, int G, long long R, int C, float eps, float momentum,
                           void* stream);
/* eval: scale = gamma/sqrt(running_var+eps), shift = beta - running_mean*scale  ([C]) */
int stfb_bn_fold_eval(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                      float* scale, float* shift, int C, float eps, void* stream);
/* y = relu?( x*scale[g][c] + shift[g][c] + residual ) */
int stfb_bn_apply(const void* x, const float* scale, const float* shift, const void* residual, void* y, int G,
                  long long R, int C, int relu, int dtype, void* stream);
/* the same pass with scale/shift derived inside the kernel from nblk <= 8 statistics slots (the conv epilogue's), by the
 * exact arithmetic of stfb_bn_finalize_train: the finalize launch then only updates the running statistics and saves
 * mean / invstd for the backward pass, and need not precede this call */
int stfb_bn_apply_from_stats(const void* x, const float* partial, int nblk, const float* gamma, const float* beta,
                             const void* residual, void* y, int G, long long R, int C, float eps, int relu, int dtype,
                             void* stream);
/* the stem of the training step in one pass (src/stf_lstm_unet.py:177-180: bn1 -> relu -> maxpool): y = maxpool_k,stride,pad(
 * relu(x * scale + shift)) with scale/shift derived from the statistics slots as above, plus the uint8 window position of
 * the first maximum (as stfb_maxpool_fwd_idx).  Values and indices equal those of stfb_bn_apply_from_stats followed by
 * stfb_maxpool_fwd_idx (every tap is rounded to bf16 before the comparison); the full-resolution post-BN map is never
 * written.  x [N,H,W,C] bf16, N a multiple of G groups of images, k in {2, 3}. */
int stfb_bn_relu_maxpool_from_stats(const void* x, const float* partial, int nblk, const float* gamma, const float* beta, void* y,
                                    unsigned char* idx, int G, int N, int H, int W, int C, int Ho, int Wo, int k, int stride,
                                    int pad, float eps, int dtype, void* stream);
/* backward, step 1: dz = dy * (y > 0 if relu); partial[blk][0][g][c] = sum dz, partial[blk][1][g][c] = sum dz*xhat */
/* relu: the ReLU mask is y > 0; with y == NULL it is recomputed as fma(x, scale, shift) > 0 from the forward's scale and
 * shift (saves the read of y; only valid when no residual was added before the ReLU). */
int stfb_bn_bwd_reduce(const void* dy, const void* y, const void* x, const float* mean, const float* invstd,
                       const float* scale, const float* shift, float* partial, int nblk, int G, long long R, int C, int relu,
                       int dtype, void* stream);
/* backward, step 2: dgamma[c] += sum_g sum2, dbeta[c] += sum_g sum1; coef[g][c][3] = {gamma*invstd, sum1/R, sum2/R} */
int stfb_bn_bwd_finalize(const float* partial, int nblk, const float* gamma, const float* invstd, float* dgamma,
                         float* dbeta, float* coef, int G, long long R, int C, void* stream);
/* backward, step 3: dx = coef0 * (dz - coef1 - xhat*coef2); if dres != NULL: dres = dz (+ dres when accum_dres:
 * the residual input of a block already carries the gradient of its other consumers) */
int stfb_bn_bwd_apply(const void* dy, const void* y, const void* x, const float* mean, const float* invstd,
                      const float* coef, const float* shift /* with y == NULL: mask = fma(x, coef0, shift) > 0 */, void* dx,
                      void* dres, int accum_dres, int G, long long R, int C, int relu, int dtype, void* stream);
/* The same backward in ONE launch (reduce -> per-group barrier -> finalize -> apply; single-wave grid).  `scratch` holds
 * stfb_bn_bwd_fused_scratch_floats(G, C) ZEROED floats (sums [G][2][C] + one arrival counter per group) and is left dirty.
 * dgamma / dbeta (nullable) are accumulated into.  y may be null when relu is set and `shift` (the forward's [G][C]) is
 * given: the mask is then recomputed from x.  dres (nullable): gradient of the residual input (= masked dy), added to its
 * old content when accum_dres is set. */
size_t stfb_bn_bwd_fused_scratch_floats(int G, int C);
int stfb_bn_bwd_fused(const void* dy, const void* y, const void* x, const float* mean, const float* invstd, const float* gamma,
                      const float* shift, float* scratch, float* dgamma, float* dbeta, void* dx, void* dres, int accum_dres,
                      int G, long long R, int C, int relu, int dtype, void* stream);
/* out[c] += sum_rows x[row][c]  (bias gradients: Conv2d/ConvTranspose2d bias, LSTM bias_ih/bias_hh) */
int stfb_colsum(const void* x, float* out, long long R, int C, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * MaxPool2d(k, stride 2, pad) NHWC, -inf padding, floor mode (src/stf_lstm_unet.py:110,180 k=3 pad=1;
 * src/unet.py:25 k=2 pad=0).  Backward routes to the first maximum in scan order (PyTorch tie rule).
 * ---------------------------------------------------------------------------------------------- */
int stfb_maxpool_fwd(const void* x, void* y, int N, int H, int W, int C, int Ho, int Wo, int k, int stride, int pad,
                     int dtype, void* stream);
int stfb_maxpool_bwd(const void* x, const void* dy, void* dx, int N, int H, int W, int C, int Ho, int Wo, int k,
                     int stride, int pad, int dtype, void* stream);

/* Training variants: forward also stores, per output element, the window position ky*k+kx of its first maximum (uint8);
 * backward gathers through that index instead of re-scanning the input windows. */
int stfb_maxpool_fwd_idx(const void* x, void* y, unsigned char* idx, int N, int H, int W, int C, int Ho, int Wo, int k,
                         int stride, int pad, int dtype, void* stream);
int stfb_maxpool_bwd_idx(const unsigned char* idx, const void* dy, void* dx, int N, int H, int W, int C, int Ho, int Wo, int k,
                         int stride, int pad, int dtype, void* stream);

/* Bilinear resize, align_corners=True (F.interpolate at src/stf_lstm_unet.py:57,191-194), NHWC. */
int stfb_bilinear_fwd(const void* x, void* y, int N, int H, int W, int C, int Ho, int Wo, int dtype, void* stream);
/* dx must be zeroed by the caller (fp32 accumulate): dx[N,H,W,C] fp32 += scatter(dy) */
int stfb_bilinear_bwd(const void* dy, float* dx, int N, int H, int W, int C, int Ho, int Wo, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Per-pixel LSTM cell (nn.LSTM(C, C, batch_first=True) applied to [B*h*w, T, C],
 * src/stf_lstm_unet.py:124-127, :216-242).  gates = pre-activations [R][4C] fp32 in PyTorch order i,f,g,o
 * (already containing W_ih x + b_ih + W_hh h + b_hh, produced by stfb_conv2d as 1x1 GEMMs).
 * ---------------------------------------------------------------------------------------------- */
/* One LSTM step as ONE tcgen05 kernel: implicit GEMM over the K-concatenation [x_t, h_{t-1}] against [W_ih | W_hh]
 * producing all four gates in TMEM, with bias add, gate non-linearities and the cell update fused in the epilogue (the
 * gate pre-activations never leave the SM).  Writes c_out (fp32 [rows][C]), h_out (bf16 [rows][C]) and, for training,
 * the post-activation gates acts (bf16 [rows][4C]) in ACCUMULATOR COLUMN ORDER: (gate, unit u) at (u/16)*64 + gate*16 +
 * u%16 (pass acts_il = 1 to stfb_lstm_cell_bwd).  h_prev = NULL at t = 0 (zero state: only W_ih is walked, c_prev ignored).
 * bf16 only; C % 64 == 0; w_xh_il = [4C][2C]: stfb_pack_weight_ex(W_ih, n_major = 1, ld = 2C, gate_c = C) into columns
 * [0, C) and the same for W_hh into columns [C, 2C); h_out must not alias x_t or h_prev; all pointers 16-byte aligned. */
int stfb_lstm_step_fused(const void* x_t, const void* h_prev, const void* w_xh_il, const float* b_ih, const float* b_hh,
                         const float* c_prev, float* c_out, void* h_out, void* acts, int N, int H, int W, int C, void* stream);
/* ALL T steps of one LSTM level in ONE launch (hidden size 64; stfb_lstm_seq_supported says whether a geometry fits): the CTA
 * that owns a 128-pixel tile keeps [W_ih | W_hh] (64 KB), c (fp32) and h_{t-1} (bf16, the next step's MMA operand) in shared
 * memory for t = 0..T-1; x_seq = [T*B, H, W, C] bf16 time-major.  keep != 0 (training): c_all fp32 [T][rows][C], h_all bf16
 * [T][rows][C], acts_all bf16 [T][rows][4C] (accumulator column order, as above) are written for every step; keep == 0
 * (inference): only h_all = h_T [rows][C].  Bit-identical to T calls of stfb_lstm_step_fused.
 * Replaces: the time loop inside nn.LSTM (src/stf_lstm_unet.py:124-127, :216-221) for the first encoder level. */
int stfb_lstm_seq_supported(int T, int B, int H, int W, int C);
int stfb_lstm_seq_fused(const void* x_seq, const void* w_xh_il, const float* b_ih, const float* b_hh, float* c_all, void* h_all,
                        void* acts_all, int T, int B, int H, int W, int C, int keep, void* stream);
/* One LSTM BACKWARD step as one tcgen05 kernel: dh_{t-1} = dG_t W_hh accumulated in TMEM, and the epilogue differentiates the
 * cell of step t-1 on the spot (the arithmetic of stfb_lstm_cell_bwd): dg_out = dG_{t-1} (bf16 [rows][4C], gate-major i,f,g,o),
 * dc rewritten in place; dh never exists in memory.  dg_next = dG_t [N,H,W,4C] bf16; w_hh_d = W_hh packed for dgrad
 * (stfb_pack_weight_ex(W_hh, k_is_dim1 = 0, n_major = 1): [C][4C]); acts (accumulator column order, from stfb_lstm_step_fused /
 * stfb_lstm_seq_fused), c_cur and c_prev belong to step t-1 (c_prev = NULL when t-1 == 0).  Replaces one
 * stfb_lstm_cell_bwd + one stfb_conv2d launch of the backward time loop (autograd of nn.LSTM, src/stf_lstm_unet.py:216-242). */
int stfb_lstm_bwd_step_fused(const void* dg_next, const void* w_hh_d, const void* acts, const float* c_prev, const float* c_cur,
                             float* dc, void* dg_out, int N, int H, int W, int C, void* stream);
/* c_prev may be NULL (t = 0, zero state).  acts (dtype, [R][4C]) may be NULL in eval mode.
 * c_out fp32 [R][C]; h_out dtype [R][C]. */
int stfb_lstm_cell_fwd(const float* gates, const float* c_prev, void* acts, float* c_out, void* h_out, long long R,
                       int C, int dtype, void* stream);
/* dh fp32 [R][C]; dc fp32 [R][C] in/out (dc_t in, dc_{t-1} out; pass zeros at t = T-1);
 * dgates dtype [R][4C] out. c_prev may be NULL (t = 0). */
int stfb_lstm_cell_bwd(const float* dh, float* dc, const void* acts, const float* c_prev, const float* c_cur,
                       void* dgates, long long R, int C, int acts_il /* acts in stfb_lstm_step_fused order */, int dtype,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * Evaluation metrics on the device (evaluate(), train_utils/train_and_eval.py:322-336; ConfusionMatrix :25-70,
 * DiceCoefficient :73-132): ONE pass over the logits [B,C,HW] fp32 (NCHW) + target [B,HW] int64 does
 *   - argmax over classes (first maximum, as torch.argmax) -> mask [B,HW] uint8 (may be NULL),
 *   - confmat[t*C + pred] += 1 for 0 <= t < C (int64 [C*C], accumulated across calls),
 *   - the Dice update of this batch: with use_ignore, pixels whose target == ignore_index count as class 0 on both
 *     sides (the reference's pred*mask / target*mask); dice_c = 2*inter/(|pred==c| + |target==c|), 1 when the union is
 *     empty; dice_cumulative[c] += dice_c, *dice_updates += 1 (both may be NULL to only collect dice_counts).
 * dice_counts = int64 [3*C] scratch {inter, pred, target}, zeroed by the caller once, cleared by each call that passes
 * dice_cumulative.  No host synchronisation.  C <= 8.
 * ---------------------------------------------------------------------------------------------- */
int stfb_eval_metrics(const float* logits, const long long* target, unsigned char* mask, long long* confmat,
                      long long* dice_counts, float* dice_cumulative, long long* dice_updates, int B, int C, int HW,
                      long long ignore_index, int use_ignore, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer step over flat buffers (replaces torch.optim.AdamW(fused=True),
 * /root/reference/train.py:227-237): decoupled weight decay, lerp first moment, bias-corrected step, in place on
 * param / exp_avg / exp_avg_sq (n fp32 each).  `step` counts from 1; grad is multiplied by grad_scale first
 * (1/world_size after a summing all-reduce, 1 otherwise).  Hyper-parameters are host doubles: 1 - beta and the bias
 * corrections are formed in double and rounded once, as torch does.
 * ---------------------------------------------------------------------------------------------- */
int stfb_adamw_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, double lr, double beta1,
                    double beta2, double eps, double weight_decay, long long step, double grad_scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Layout adapters at the API boundary (reference tensors are NCHW fp32).
 * ---------------------------------------------------------------------------------------------- */
/* x [B, T, C, H, W] fp32  ->  y [T*B, H, W, C] dtype  (time-major image order n = t*B + b) */
int stfb_pack_series(const float* x, void* y, int B, int T, int C, int H, int W, int dtype, void* stream);
/* PK-map branch (src/stf_lstm_unet.py:146-156, :172-174): x [B,T,Cx,H,W] fp32 + maps [B,Cm,H,W] fp32 ->
 * y [T*B, H, W, Cx+Cm] dtype (the torch.cat([x_t, pk_maps], 1) of every time step, time-major) */
int stfb_pack_series_maps(const float* x, const float* maps, void* y, int B, int T, int Cx, int Cm, int H, int W, int dtype,
                          void* stream);
/* Device-side input pipeline: 8-bit series x [B, T, H, W] -> y [T*B, H, W, 1] dtype with
 * the loader's arithmetic fused in, ((x / 255) - mean) / std, operation for operation as ToTensor + Normalize do it
 * (/root/reference/transforms.py:120-134, train.py:147-148: mean 0.709, std 0.127), so the fp32 result is bit-identical. */
int stfb_pack_series_u8(const unsigned char* x, void* y, int B, int T, int H, int W, float mean, float std_, int dtype,
                        void* stream);
/* dst = `times` back-to-back copies of src (bytes each): a per-sample map repeated for every time step */
int stfb_repeat(const void* src, void* dst, size_t bytes, int times, void* stream);
/* y NHWC dtype [N,H,W,C] -> out NCHW fp32 [N,C,H,W] */
int stfb_nhwc_to_nchw(const void* y, float* out, int N, int H, int W, int C, int dtype, void* stream);
/* g NCHW fp32 -> NHWC dtype */
int stfb_nchw_to_nhwc(const float* g, void* out, int N, int H, int W, int C, int dtype, void* stream);
/* dst[i] += src[i]  (gradient accumulation where a tensor has two consumers, e.g. the UNet skip + pool) */
int stfb_add_inplace(void* dst, const void* src, long long n, int dtype, void* stream);
/* dst[i] = (dtype) src[i]  /  dst[i] = (float) src[i] */
int stfb_cast(const void* src, int src_dtype, void* dst, int dst_dtype, long long n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * criterion = cross_entropy(mean) + multiclass softmax-Dice  (train_utils/train_and_eval.py:299-313,
 * train_utils/dice_coefficient_loss.py:5-55).  logits NCHW fp32 [B,C,h,w], target int64 [B,h,w].
 * stats fp64 [B*C*3 + 3] scratch (zeroed inside): per (b,c) {sum p*t, sum p, sum t} over the valid pixels, then
 * {sum w[t]*nll, sum w[t], number of labels outside [0,C) that are not ignore_index}.
 * loss_out fp32 [3] = {total, ce, dice_loss}.
 * _ex: the full signature of the reference's criterion -- class_weight fp32 [C] or NULL (F.cross_entropy's `weight`),
 * ignore_index (pixels with that label enter neither term: cross-entropy is the weighted mean over the others, the
 * per-image Dice sums run over the others; build_target / dice_coeff, dice_coefficient_loss.py:5-39), with_dice = 0 drops
 * the Dice term.  A label outside [0,C) that is not ignore_index makes the reference raise; here every output turns
 * NaN (no host synchronisation).  The plain entry points are the reference's training configuration
 * (no weights, ignore_index = -100, Dice on).
 * ---------------------------------------------------------------------------------------------- */
int stfb_ce_dice_fwd(const float* logits, const long long* target, double* stats, float* loss_out, int B, int C,
                     int HW, float eps, void* stream);
int stfb_ce_dice_fwd_ex(const float* logits, const long long* target, const float* class_weight, double* stats,
                        float* loss_out, int B, int C, int HW, float eps, long long ignore_index, int with_dice,
                        void* stream);
/* dlogits = dloss[0] * d(total)/d(logits); dloss is a device fp32 scalar (NULL = 1.0). */
int stfb_ce_dice_bwd(const float* logits, const long long* target, const double* stats, const float* dloss,
                     float* dlogits, int B, int C, int HW, float eps, void* stream);
int stfb_ce_dice_bwd_ex(const float* logits, const long long* target, const float* class_weight, const double* stats,
                        const float* dloss, float* dlogits, int B, int C, int HW, float eps, long long ignore_index,
                        int with_dice, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Extended Tofts pharmacokinetic model per pixel (offline PK-map preprocessing).
 * Replaces: ToftsModelFitter.extended_tofts_model_batch (pk_fitting.py:193-231) and the Adam fitting loop of
 * ToftsModelFitter.fit_volume_gpu (pk_fitting.py:288-368).
 * Tables (device, fp32): t[T] acquisition times, aif_t[T] = aif(t), t_conv[M] = arange(0, t_max, dt), aif_conv[M] =
 * aif(t_conv), nvalid[T] (int32) = number of grid points with t_conv < t[i].  T <= 32.
 *   stfb_tofts_forward: out[N,T] = vp*aif(t_i) + Ktrans*dt*sum_{tc_j < t_i} aif(tc_j) exp(-Ktrans (t_i - tc_j)/ve)
 *   stfb_tofts_fit:     pixels[N,T] observed curves; ktrans/ve/vp[N] initial values in, fitted values out.  Replays the
 *     reference's optimisation exactly: ONE Adam state over all N pixels stepped once per batch of `batch_size` consecutive
 *     pixels (pixels outside the batch see a zero gradient and still move by their momentum), MSE mean over the batch,
 *     clamp to [clamp_lo, clamp_hi] (host arrays of 3: Ktrans, ve, vp) after every step.  step_size / bc2_sqrt: device
 *     arrays [epochs * ceil(N / batch_size)] with lr / (1 - beta1^s) and sqrt(1 - beta2^s), s = 1, 2, ... (formed in
 *     double on the host like torch).  epoch_loss: optional device fp32 [epochs], mean batch loss per epoch.
 * ---------------------------------------------------------------------------------------------- */
int stfb_tofts_forward(const float* t, const float* aif_t, const float* t_conv, const float* aif_conv, const int* nvalid,
                       int T, int M, float dt, const float* ktrans, const float* ve, const float* vp, float* out,
                       long long N, void* stream);
int stfb_tofts_fit(const float* pixels, const float* t, const float* aif_t, const float* t_conv, const float* aif_conv,
                   const int* nvalid, int T, int M, float dt, float* ktrans, float* ve, float* vp, long long N,
                   int batch_size, int epochs, const float* step_size, const float* bc2_sqrt, float beta1, float beta2,
                   float eps, const float* clamp_lo, const float* clamp_hi, float* epoch_loss, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Paired train-time augmentation on the device.
 * Replaces: get_transform(train=True) (train.py:51-67) = transforms.py RandomResize :18-33, RandomHorizontalFlip :36-45,
 * RandomVerticalFlip :48-57, RandomRotation :136-157, RandomCrop :60-117, ToTensor :120-125, Normalize :127-134, which the
 * reference's loader applies per phase on the CPU through PIL (my_dataset.py:173-179).
 * series uint8 [B,T,H,W], masks uint8 [B,H,W] (values {0,1}; NULL when no target is wanted); x_out fp32 [B,T,1,S,S]
 * normalised crops, target_out int64 [B, ceil(S/stride), ceil(S/stride)] (the mask crop, every stride-th pixel; NULL = skip).
 * samples_dev: one stfb_aug_sample per batch element (host-drawn geometry, shared by all phases and the mask);
 * tables_dev: int32 tables, per sample at tab_off: hb[rw][2], hk[rw][ksize_h], vb[rh][2], vk[rh][ksize_v] (Pillow's bilinear
 * resize: first source index / tap count and 22-bit fixed-point coefficients per output index; ksize == 0: that pass is the
 * identity), xtab[rw], ytab[rh] (source index of Pillow's nearest resize, -1 = outside).
 * Output is bit-identical to the PIL pipeline on 8-bit single-channel images.
 * ---------------------------------------------------------------------------------------------- */
typedef struct stfb_aug_sample {
  int rh, rw;              /* size after RandomResize */
  int hflip, vflip, rot;   /* 0 / 1 */
  int h0, w0;              /* RandomCrop origin in the (zero padded) image */
  int tab_off;             /* offset of this sample's tables in tables_dev, in ints */
  int ksize_h, ksize_v;    /* taps per output index of the horizontal / vertical resize pass (0: pass skipped) */
  int fix[6];              /* 16.16 fixed-point affine of Image.rotate(NEAREST): a0, a1, a2, a3, a4, a5 */
  int pad_;
  double m[6];             /* double affine of Image.rotate(BILINEAR): output pixel centre -> input coordinate */
} stfb_aug_sample;
int stfb_augment_series_u8(const unsigned char* series, const unsigned char* masks, const stfb_aug_sample* samples_dev,
                           const int* tables_dev, float* x_out, long long* target_out, int B, int T, int H, int W, int S,
                           int target_stride, float mean, float stdv, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STFB200_H_ */
