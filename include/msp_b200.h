/*
 * msp_b200.h — C ABI of libmsp_b200.so: the B200 (sm_100a) kernels behind the training / evaluation
 * hot path of aielte-research/MedSegPretrainImageNet.
 *
 * The reference is pure Python on stock PyTorch; "the interface each entry point replaces" is
 * therefore the ATen op(s) the reference's nn.Modules / losses / metrics dispatch to (SURVEY.md
 * §2.3, §8a).  Citations are `reference/src/...:line`.
 *
 * Conventions
 *  - All pointers are DEVICE pointers unless marked host.  The caller owns every buffer
 *    (PyTorch's caching allocator in the shipped host code); the library allocates nothing.
 *  - All calls are asynchronous on `stream` (a cudaStream_t passed as void*), never touch the
 *    default stream implicitly and never synchronise.
 *  - Activations are NHWC bf16 ("pixel-major"): element (n,h,w,c) of a tensor with pixel stride
 *    `cs` lives at ((n*H + h)*W + w)*cs + c.  A pixel stride larger than C lets several
 *    producers write channel slices of one concat buffer (zero-copy torch.cat(dim=1),
 *    reference blocks.py:628,635).  Channel counts and pixel strides are multiples of 8.
 *  - Return value 0 = ok, <0 = error (see MSP_ERR_*); msp_last_error() gives a thread-local text.
 *    There is no CPU fallback: unsupported shapes return MSP_ERR_UNSUPPORTED.
 */
#ifndef MSP_B200_H
#define MSP_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSP_ABI_VERSION 3

const char* msp_last_error(void);
int msp_version(void);
/* number of kernels launched by this library on the calling process so far (bench "gpu_launches") */
long long msp_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Convolution (implicit GEMM on tcgen05 tensor cores, TMEM accumulators, TMA-staged tiles).
 * Replaces nn.Conv2d forward / autograd dgrad / wgrad:
 *   classification/models.py:43-46,161-179,234-253 ; segmentation/models/blocks.py:458,518,590 ;
 *   segmentation/models/unet_models.py:440-445 (cuDNN / mkldnn_convolution in the reference).
 * ------------------------------------------------------------------------------------------ */
typedef struct msp_conv_desc {
  int32_t N, H, W;          /* input batch / height / width                                   */
  int32_t C;                /* input channels as stored (multiple of 8; zero-padded if needed) */
  int32_t x_cs;             /* input pixel stride in elements (>= C, multiple of 8)            */
  int32_t Ho, Wo;           /* output height / width                                           */
  int32_t K;                /* output channels (multiple of 8)                                 */
  int32_t y_cs;             /* output pixel stride in elements (>= K, multiple of 8)           */
  int32_t KH, KW;           /* filter size                                                     */
  int32_t stride;           /* 1 or 2                                                          */
  int32_t pad_t, pad_l;     /* top / left zero padding (bottom/right implied by Ho, Wo)        */
  int32_t relu;             /* fprop: apply ReLU in the epilogue (conv -> ReLU without BN)     */
  /* Row-window mode for the tiny-channel first convolution (7x7/2 stem of DeepResNet,
   * classification/models.py:43-46; 3x3 first_block of UNet_encoder, unet_models.py:440): win_px = 8 or 4
   * pixels per window (0 = off).  x is then the W-padded tensor [N][H][Wp][C] written by
   * msp_nchw_f32_to_rowwin_bf16 with C = 64 / win_px channels per pixel (image column w at column
   * w + pad_l), and the weights come from msp_pack_weights_rowwin.  fprop and wgrad only.          */
  int32_t win_px, Wp;
  /* BatchNorm statistics of fprop (ch_sum / ch_sqsum): stat_rows = 0 -> every CTA ADDS its per-channel partial sums to
   * the [2][K] accumulators with float atomics (the order of the additions varies from run to run).  stat_rows > 0 ->
   * DETERMINISTIC mode (torch.use_deterministic_algorithms, run_experiment.py:65; every downstream YAML sets it): ch_sum
   * points to a zero-initialised [stat_rows][2][K] fp32 workspace (ch_sqsum = ch_sum + K), CTA b STORES its partial sums
   * in row b (stat_rows >= number of SMs) and msp_bn_finalize / msp_reduce_rows add the rows in fixed order. */
  int32_t stat_rows;
} msp_conv_desc;

/* Kernel-variant policy of fprop / dgrad (tuning and tests; -1 restores the default / environment):
 * pair: 0 never use the cta_group::2 CTA-pair tap-GEMM, 1 for >= 128 output channels per tile, 2 also for 64;
 * halo: 0 never use the halo-reuse kernel, 1 when the weights stay resident in shared memory, 2 whenever it applies. */
int msp_conv_set_policy(int pair, int halo);
/* Name of the kernel variant the calling thread's last msp_conv_* call launched ("tapgemm_kernel<256>",
 * "tapgemm_halo_kernel<16>", "wgrad_kernel", ...): lets the benchmark attribute device time per kernel. */
const char* msp_conv_last_kernel(void);

/* OIHW fp32 master weights -> bf16 [K][KH*KW][Cpad] (fprop/wgrad operand) and, if w_dgrad != NULL,
 * bf16 [Cpad][KH*KW][Kpad] (dgrad operand).  Cpad/Kpad >= C/K, multiples of 8, padding zero-filled. */
int msp_pack_weights(const float* w_oihw, int K, int C, int KH, int KW, int Cpad, int Kpad,
                     void* w_fprop, void* w_dgrad, void* stream);

/* OIHW fp32 -> bf16 [K][KH][64] with element q*(64/win_px)+c of filter row r = w[k][c][r][q]. */
/* All convolutions of a model in one launch: `items_dev` = device array of n_items x 10 int64
 * {w_oihw ptr, w_fprop ptr, w_dgrad ptr (0 = not needed), K, C, KH*KW, Cpad, Kpad, first block, blocks} (layouts as in
 * msp_pack_weights); item i owns blocks [first, first + blocks) of the `total_blocks` 256-thread blocks launched. */
int msp_pack_weights_batched(const long long* items_dev, int n_items, int total_blocks, void* stream);
int msp_pack_weights_rowwin(const float* w_oihw, int K, int C, int KH, int KW, int win_px,
                            void* w_rowwin, void* stream);

/* y = conv(x, w) (+ bias) (ReLU optional).  If ch_sum / ch_sqsum are non-NULL the per-output-channel
 * sum and sum of squares of the (bf16-rounded) outputs over N*Ho*Wo are ADDED to them (fp32, [K]) —
 * the BatchNorm batch statistics (blocks.py:462, classification/models.py:47) come for free. */
int msp_conv_fprop(const msp_conv_desc* d, const void* x, const void* w_fprop, const float* bias,
                   void* y, float* ch_sum, float* ch_sqsum, void* stream);

/* dx = conv_transpose(dy, w): gradient w.r.t. the input.  `w_dgrad` is the second output of
 * msp_pack_weights.  dx has d->C channels at pixel stride d->x_cs; dy has d->K channels at d->y_cs. */
int msp_conv_dgrad(const msp_conv_desc* d, const void* dy, const void* w_dgrad, void* dx,
                   int accumulate /* dx += result (residual gradient already in dx) */, void* stream);
/* nn.ConvTranspose2d forward (north star: "conv/transposed-conv layers"; the reference's own up-sampling is nearest x2 +
 * Conv2d, segmentation/models/blocks.py:531-535) = the data gradient of the convolution it transposes, with bias and
 * ReLU in the same epilogue.  `d` describes that convolution: (N, H, W, C) = the transposed conv's OUTPUT,
 * (Ho, Wo, K) = its input x; w_dgrad = the [C][tap][K] packing of the (in_channels = K, out_channels = C, kh, kw)
 * weight.  Its backward passes are msp_conv_fprop (data) and msp_conv_wgrad with the two tensors swapped (weights). */
int msp_conv_transpose_fprop(const msp_conv_desc* d, const void* x, const void* w_dgrad, const float* bias, int relu,
                             void* y, void* stream);

/* Weight gradient, split-K over pixel tiles WITHOUT atomics (deterministic): split s writes its partial
 * sum, packed like w_fprop ([K][KH*KW][C] fp32; [K][KH][64] in row-window mode), at
 * dw_partials + s * K*taps*C.  msp_conv_wgrad_splits(d) >= 1 is the number of partials the caller must
 * provide room for (it depends only on the descriptor and the SM count); every element is written. */
int msp_conv_wgrad_splits(const msp_conv_desc* d);
int msp_conv_wgrad(const msp_conv_desc* d, const void* x, const void* dy, float* dw_partials,
                   void* stream);

/* sum of the partials -> OIHW fp32 [K][C_true][KH][KW] (the layout of nn.Conv2d.weight.grad). */
int msp_unpack_wgrad(const msp_conv_desc* d, const float* dw_partials, int C_true, float* dw_oihw,
                     void* stream);
/* Every weight gradient of a backward pass in ONE launch: item i = the fixed-order sum over `splits` partials
 * (msp_conv_wgrad's output, `split_stride` elements apart) written — or ADDED when `accumulate` (gradient
 * accumulation, all-reduce bucket views) — to `dst` in the OIHW fp32 layout of nn.Conv2d.weight.grad.
 * Generic layers: partials [K][taps][Cpad]; row-window layers (rowwin_KH > 0): partials [K][KH][64] with
 * `rowwin_cpp` channels per window pixel.  `items` is a HOST array of 1..96 entries (it travels in the kernel's
 * parameter space). */
/* ------------------------------------------------------------------------------------------
 * Folded up-convolution: nn.Upsample(scale_factor=2) (nearest) -> Conv2d(k=2, padding='same') of UpConvBlock
 * (segmentation/models/blocks.py:531-535) computed on the LOW-RES input: output pixel (2i+a, 2j+b) only ever reads
 * low-res rows i (a = 0) or i, i+1 (a = 1), so the layer is four output-parity classes of 1x1 / 1x2 / 2x1 / 2x2
 * convolutions with pre-summed weights — 9 instead of 16 taps per 2x2 output block, no x4 tensor, no up-sample kernels.
 * Folded weights: fp32 [K][C][9], tap order class (a,b) = (0,0),(0,1),(1,0),(1,1) at 0,1,3,5, (dr,dq) row-major inside;
 * pack them with msp_pack_weights as a 1x9 filter.  `d` everywhere: (N,H,W,C,x_cs) = the low-res input, (Ho,Wo) = (2H,2W),
 * K, y_cs (relu: fprop epilogue).  The weight gradient runs per class (msp_upconv2x_wgrad_class with the class's own
 * descriptor KH = 1+a, KW = 1+b, pad 0, Ho = H, Wo = W; msp_upconv2x_wgrad_splits sizes its partials), is unpacked per
 * class ([K][C][1+a][1+b]) and mapped back to the 2x2 filter by msp_unfold_upconv_wgrad (the fold is linear).
 * ------------------------------------------------------------------------------------------ */
int msp_fold_upconv_weights(const float* w, int K, int C, float* w_folded, void* stream);
int msp_unfold_upconv_wgrad(const float* g00, const float* g01, const float* g10, const float* g11, int K, int C,
                            float* dw, int accumulate, void* stream);
int msp_upconv2x_fprop(const msp_conv_desc* d, const void* x, const void* w_folded_fprop, const float* bias, void* y,
                       void* stream);
int msp_upconv2x_dgrad(const msp_conv_desc* d, const void* dy, const void* w_folded_dgrad, void* dx, int accumulate,
                       void* stream);
int msp_upconv2x_wgrad_splits(const msp_conv_desc* d);
int msp_upconv2x_wgrad_class(const msp_conv_desc* d, const void* x, const void* dy, int a, int b, float* dw_partials,
                             void* stream);

typedef struct msp_unpack_item {
  const float* partials;
  float* dst;
  long long split_stride;
  int32_t splits, K, C_true, taps, Cpad;
  int32_t rowwin_KH, rowwin_KW, rowwin_cpp;
  int32_t accumulate;
  int32_t reserved;
} msp_unpack_item;
int msp_unpack_wgrad_batched(int n, const msp_unpack_item* items, void* stream);

/* ------------------------------------------------------------------------------------------
 * Layout conversion at the module boundary (reference tensors are NCHW fp32).
 * ------------------------------------------------------------------------------------------ */
int msp_nchw_f32_to_nhwc_bf16(const float* x, int N, int C, int H, int W, int Cpad, void* y,
                              void* stream);
int msp_nhwc_bf16_to_nchw_f32(const void* x, int N, int C, int H, int W, int x_cs, float* y,
                              void* stream);
/* NCHW fp32 -> W-padded [N][H][Wp][cpp] bf16 (cpp = 8 or 16 >= C) for row-window convolutions. */
int msp_nchw_f32_to_rowwin_bf16(const float* x, int N, int C, int H, int W, int cpp, int pad_l, int Wp,
                                void* y, void* stream);
/* same from a bf16 NCHW batch (BASELINE cfg2 feeds bf16 images: half the host -> device bytes of the
 * reference's fp32 `.to(device)`, train_model.py:60). */
int msp_nchw_bf16_to_rowwin_bf16(const void* x, int N, int C, int H, int W, int cpp, int pad_l, int Wp,
                                 void* y, void* stream);
int msp_nchw_f32_grad_to_nhwc_bf16(const float* g, int N, int C, int H, int W, int g_cs_out,
                                   void* y, void* stream);

/* ------------------------------------------------------------------------------------------
 * BatchNorm2d (train + eval) fused with activation / residual.  Replaces nn.BatchNorm2d + nn.ReLU +
 * the zero-fill / stride-2 shortcut + DropPath multiply of the ResNet blocks
 * (classification/models.py:203-212, 277-290) and BN+ReLU of ConvBlock (blocks.py:458-488).
 * ------------------------------------------------------------------------------------------ */
/* sums -> mean / invstd, running-stat update (momentum, unbiased var) like torch (eps 1e-5).
 * reset_sums != 0: the accumulators are zeroed after they have been read (persistent per-layer
 * accumulators need no memset launch per step).  rows > 1: ch_sum is the [rows][2][C] per-CTA workspace of the
 * deterministic mode (msp_conv_desc.stat_rows), summed here row by row in fixed order. */
int msp_bn_finalize(float* ch_sum, float* ch_sqsum, int C, double count, float eps, float momentum,
                    float* mean, float* invstd, float* running_mean, float* running_var, int reset_sums,
                    int rows, void* stream);
/* out[i] = ws[0][i] + ws[1][i] + ... + ws[rows-1][i] (fixed order, i < n); reset != 0 zeroes ws afterwards.  The
 * fixed-order second stage of every deterministic reduction (SyncBN sums before their all-reduce, BatchNorm-backward
 * sums, bias gradients). */
int msp_reduce_rows(float* ws, int rows, int n, float* out, int reset, void* stream);

#define MSP_ACT_NONE 0
#define MSP_ACT_RELU 1
#define MSP_ACT_SIGMOID 2

typedef struct msp_bn_act_desc {
  int32_t N, H, W, C;       /* shape of x / y                                                  */
  int32_t x_cs, y_cs;       /* pixel strides                                                   */
  int32_t act;              /* MSP_ACT_*                                                       */
  /* optional residual r: out = act( scale[n] * bn(x) + r ), r taken from a tensor of shape
   * (N, H*r_stride, W*r_stride, r_C) at pixel stride r_cs, sub-sampled by r_stride, channels
   * >= r_C read as zero (zero-fill shortcut, classification/models.py:257-274).               */
  int32_t r_C, r_cs, r_stride;
} msp_bn_act_desc;

/* y = act( s[n] * (gamma*(x-mean)*invstd + beta) + residual ).  sample_scale may be NULL (=1). */
int msp_bn_act_fwd(const msp_bn_act_desc* d, const void* x, const float* mean, const float* invstd,
                   const float* gamma, const float* beta, const float* sample_scale,
                   const void* residual, void* y, void* stream);

/* In both backward passes `y` may be NULL for BatchNorm -> ReLU without shortcut / sample scale: the ReLU mask is then
 * recomputed from x with the forward's own expression (gamma, beta needed) and the stored activation is not read.
 * Backward pass 1: with g = dy * act'(y) (ReLU mask recomputed from y; sigmoid from y),
 * accumulates sum_g[c] += sum s[n]*g, sum_gx[c] += sum s[n]*g*xhat (fp32 [C]); if dres != NULL also
 * writes the residual-branch gradient (g itself) — strided / channel-truncated like the forward. */
/* rows_ws != NULL (deterministic mode): fp32 workspace of ws_rows x 2C floats; every block stores its partial sums in
 * its own row and a second launch adds the rows in fixed order into sum_g / sum_gx (= sum_g + C required). */
int msp_bn_act_bwd_reduce(const msp_bn_act_desc* d, const void* x, const void* y, const void* dy,
                          const float* mean, const float* invstd, const float* gamma, const float* beta,
                          const float* sample_scale, float* sum_g, float* sum_gx, float* rows_ws, int ws_rows,
                          void* stream);
/* First stage only: block b's partial sums in row b of the [ws_rows][2][C] workspace, *rows_used (host int) = the
 * number of rows written; the caller adds them (msp_p2p_stats_exchange does it inside the SyncBN exchange). */
int msp_bn_act_bwd_reduce_rows(const msp_bn_act_desc* d, const void* x, const void* y, const void* dy,
                               const float* mean, const float* invstd, const float* gamma, const float* beta,
                               const float* sample_scale, float* rows_ws, int ws_rows, int* rows_used, void* stream);
/* Backward pass 2: dx = gamma*invstd*( s*g - sum_g/M - xhat*sum_gx/M ); M = N*H*W (x world size
 * when the sums were all-reduced: pass the global count).  dres (optional, same shape as the
 * residual tensor) receives g added into the sub-sampled positions (others untouched).          */
int msp_bn_act_bwd_apply(const msp_bn_act_desc* d, const void* x, const void* y, const void* dy,
                         const float* mean, const float* invstd, const float* gamma, const float* beta,
                         const float* sample_scale, const float* sum_g, const float* sum_gx,
                         double count, void* dx, void* dres, int dres_accumulate, void* stream);
/* eval-mode BN uses running stats: same forward with mean=running_mean, invstd=rsqrt(var+eps):   */
int msp_bn_eval_prepare(const float* running_var, int C, float eps, float* invstd, void* stream);

/* ------------------------------------------------------------------------------------------
 * Pooling / resampling (nn.MaxPool2d classification/models.py:56, unet_models.py:452;
 * nn.Upsample(scale_factor=2) nearest blocks.py:532,615; AdaptiveAvgPool2d models.py:73).
 * ------------------------------------------------------------------------------------------ */
/* idx (optional, uint8 [N][Ho][Wo][C]) receives the in-window arg-max (first max in scan order,
 * like ATen) for the backward pass. */
int msp_maxpool_fwd(const void* x, int N, int H, int W, int C, int x_cs, int k, int stride, int pad,
                    void* y, void* idx, int Ho, int Wo, int y_cs, void* stream);
/* dx (+)= gather of dy over the windows whose arg-max is this pixel (no atomics). */
int msp_maxpool_bwd(const void* idx, const void* dy, int N, int H, int W, int C, int k, int stride,
                    int pad, int Ho, int Wo, int dy_cs, void* dx, int dx_cs, int accumulate,
                    void* stream);
int msp_upsample2x_fwd(const void* x, int N, int H, int W, int C, int x_cs, void* y, int y_cs,
                       void* stream);
int msp_upsample2x_bwd(const void* dy, int N, int H, int W, int C, int dy_cs, void* dx, int dx_cs,
                       void* stream);
/* nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False) forward / backward on NHWC bf16 (north star
 * "bilinear/nearest upsampling"; the reference's blocks instantiate the nearest mode, blocks.py:532).  x (N,H,W,C) ->
 * y (N,2H,2W,C); the backward is the gather form of the transpose (no atomics). */
int msp_upsample_bilinear2x_fwd(const void* x, int N, int H, int W, int C, int x_cs, void* y, int y_cs,
                                void* stream);
int msp_upsample_bilinear2x_bwd(const void* dy, int N, int H, int W, int C, int dy_cs, void* dx, int dx_cs,
                                void* stream);
int msp_avgpool_fwd(const void* x, int N, int HW, int C, int x_cs, void* y, void* stream);
int msp_avgpool_bwd(const void* dy, int N, int HW, int C, void* dx, int dx_cs, void* stream);
/* y[.., off:off+C] = x (channel-slice copy into a concat buffer) and its inverse for gradients. */
int msp_copy_channels(const void* x, long long P, int C, int x_cs, void* y, int y_cs, void* stream);
/* out[c] = sum over P pixels of x[p][c] (fp32; conv bias gradients). out is zeroed by the call.  rows_ws != NULL:
 * deterministic mode, as in msp_bn_act_bwd_reduce (workspace of ws_rows x C floats). */
int msp_channel_sum(const void* x, long long P, int C, int cs, float* out, float* rows_ws, int ws_rows, void* stream);
/* generic elementwise helpers on NHWC bf16 */
int msp_add_relu_fwd(const void* a, const void* b, long long P, int C, int a_cs, int b_cs, void* y,
                     int y_cs, void* stream);
int msp_relu_bwd(const void* y, const void* dy, long long P, int C, int y_cs, int dy_cs, void* dx,
                 int dx_cs, void* stream);
int msp_add(const void* a, const void* b, long long P, int C, int a_cs, int b_cs, void* y, int y_cs,
            void* stream);
/* attention gate tail (blocks.py:624-628): out[.., off:] = skip * up2(p);  and backward. */
int msp_gate_mul_fwd(const void* skip, const void* p, int N, int H, int W, int C, int skip_cs,
                     int p_cs, void* y, int y_cs, void* stream);
int msp_gate_mul_bwd(const void* skip, const void* p, const void* dy, int N, int H, int W, int C,
                     int skip_cs, int p_cs, int dy_cs, void* dskip, int dskip_cs,
                     int dskip_accumulate, void* dp, int dp_cs, void* stream);

/* ------------------------------------------------------------------------------------------
 * Heads and losses.
 * ------------------------------------------------------------------------------------------ */
/* Final 1x1 conv with <= 8 output channels + activation, NHWC bf16 in -> NCHW fp32 out
 * (unet_models.py:442-445 + :685-686).  act: 0 none, 1 sigmoid, 2 softmax(dim=1). */
int msp_final_conv_act_fwd(const void* x, int N, int H, int W, int C, int x_cs, const float* w,
                           const float* bias, int K, int act, float* logits_nchw, float* prob_nchw,
                           void* stream);
/* given dL/dprob (NCHW fp32) -> dlogits (through act), dx (NHWC bf16), dw [K][C], db [K] (atomics,
 * zeroed by the call).  rows_ws != NULL: deterministic mode — workspace of ws_rows x (K*C + K) floats, the blocks'
 * partial rows are added in fixed order into the contiguous result dw | db (db = dw + K*C required). */
int msp_final_conv_act_bwd(const void* x, int N, int H, int W, int C, int x_cs, const float* w,
                           int K, int act, const float* prob_nchw, const float* dprob_nchw,
                           void* dx, int dx_cs, float* dw, float* db, float* rows_ws, int ws_rows, void* stream);

/* Dice loss (segmentation/losses/losses.py:34-58).  For each (group g, class c):
 *   I = sum y*p, Y = sum y, S = sum p^2   with y = (mask == c + label_offset),
 * groups = 1 (batchwise) or N;  loss = 1 - mean_{g, c >= class_start} (2I+eps)/(Y+S+eps).
 * prob is NCHW fp32 [N][Cp][HW]; mask int64 [N][HW].  If two_class != 0 (requires Cp == 1) the classes
 * are [1-p, p] (losses.py:46-49); the `include_background=False`, one-channel branch (losses.py:50-52)
 * is Cp == 1, two_class = 0, label_offset = 1.
 * msp_dice_sums -> sums fp64 [G][Ceff][3] (zeroed by the call); msp_dice_finalize -> coef fp32
 * [G][Ceff][2] = the per-class gradient coefficients consumed by msp_dice_bwd (dL/dp = a*y + b*p)
 * and loss fp32 [1]. */
int msp_dice_sums(const float* prob, const int64_t* mask, int N, int Cp, long long HW, int two_class,
                  int label_offset, int batchwise, double* sums, void* stream);
/* (sums may have been all-reduced across ranks in between: batchwise Dice is a ratio of GLOBAL sums) */
int msp_dice_finalize(const double* sums, int G, int Ceff, int class_start, float eps, float* coef,
                      float* loss, void* stream);
/* dprob = gscale * (*gscale_dev if non-NULL) * dL/dprob (NCHW fp32, every element written).  The
 * device-side factor is autograd's incoming gradient (loss/loss.py:83-87), read without a host sync. */
int msp_dice_bwd(const float* prob, const int64_t* mask, int N, int Cp, long long HW, int two_class,
                 int label_offset, int batchwise, const float* coef, float gscale,
                 const float* gscale_dev, float* dprob, void* stream);
/* pixel-wise CE on probabilities (classification/losses.py:27-40, apply_softmax=False):
 * loss_sum (fp64 [1], optional, zeroed by the call) = sum over pixels of
 * -sum_c clamp(nan_to_num(log p_c),-100)*t_c with t = one-hot(label) clamped to [smooth/C, 1-smooth/C]
 * when smooth != 0; dprob (optional) = gscale * (*gscale_dev) * d loss_sum / d prob.  label int64 [N][HW]. */
int msp_ce_prob_fwd_bwd(const float* prob, const int64_t* label, int N, int C, long long HW,
                        float smooth, float gscale, const float* gscale_dev, double* loss_sum,
                        float* dprob, void* stream);
/* BCE sum: clamp_log = 0 -> classification/losses.py:4-11 (plain logs, autograd gradient);
 * clamp_log = 1 -> torch.nn.BCELoss (utils/default_dict.py:10; logs clamped at -100,
 * gradient (p-y)/max(p(1-p),1e-12)).  target fp32, same shape as prob. */
int msp_bce_fwd_bwd(const float* prob, const float* target, long long numel, int clamp_log,
                    float gscale, const float* gscale_dev, double* loss_sum, float* dprob,
                    void* stream);
/* classification head loss: F.cross_entropy(logits[N][C], label[N], label_smoothing)
 * (classification/losses.py:24-25): loss_sum fp64 [1] = sum of per-row losses; dlogits (optional) =
 * gscale * (*gscale_dev) * d loss_sum / d logits. */
int msp_softmax_ce_fwd_bwd(const float* logits, const int64_t* label, int N, int C, float smooth,
                           float gscale, const float* gscale_dev, double* loss_sum, float* dlogits,
                           void* stream);
/* torch.nn.CrossEntropyLoss with class-PROBABILITY targets (Mixup / CutMix, config/pretraining/resnet50/advanced.yaml:17,48):
 * target fp32 [N][C]; t' = (1 - smooth) t + smooth / C, loss_sum += sum_n (lse_n sum_c t'_c - sum_c t'_c z_c). */
int msp_softmax_ce_soft_fwd_bwd(const float* logits, const float* target, int N, int C, float smooth, float gscale,
                                const float* gscale_dev, double* loss_sum, float* dlogits, void* stream);
/* F.cross_entropy on spatial logits fp32 [N][C][HW] with class-index targets int64 [N][HW] (classification/losses.py:24-25
 * applied to a segmentation map), label smoothing as in msp_softmax_ce_fwd_bwd; loss_sum accumulates over all N*HW pixels. */
int msp_softmax_ce_spatial_fwd_bwd(const float* logits, const int64_t* label, int N, int C, long long HW, float smooth,
                                   float gscale, const float* gscale_dev, double* loss_sum, float* dlogits, void* stream);
/* out[0] = (float)(in[0] * scale): turns a loss sum into the mean without a host round trip. */
int msp_scale_to_float(const double* in, double scale, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Metrics (metrics/metrics.py:61-95, metrics/multiclass_metrics.py:90-107, 424-446).
 * ------------------------------------------------------------------------------------------ */
/* Binary confusion counts, single pass.  pred fp32 [N][C][HW], target [N][C][HW] (int64 or fp32,
 * target_is_float).  per_channel=0: out int64[6] = {TP,TN,FP,FN,class_count,nan_count} over all
 * elements; per_channel=1: out int64[C][6].  `pred >= thr`, `target == 1`; NaN targets counted
 * (the caller subtracts nan_count*multiplicity from TN like metrics.py:69,76).  out is zeroed. */
int msp_confusion_binary(const float* pred, const void* target, int target_is_float, int N, int C,
                         long long HW, float thr, int per_channel, long long* out, void* stream);
/* Multi-class: cm[t][p] += 1 with p = argmax_c pred[n][c][hw] (first max wins), t = target
 * (int64 [N][HW]) or argmax of a one-hot target (target_is_onehot, fp32 [N][C][HW]).
 * cm int64 [C][C], zeroed by the call. */
int msp_confusion_multiclass(const float* pred, const void* target, int target_is_onehot, int N,
                             int C, long long HW, long long* cm, void* stream);
/* number of rows whose label is among the top-k scores (torch.topk tie order: lower index first)*/
int msp_topk_hits(const float* pred, const int64_t* label, int N, int C, long long HW, int k,
                  long long* hits, void* stream);

/* ------------------------------------------------------------------------------------------
 * Robustness distances (robustness/distance.py:3-10, robustness/eval.py:16-28).
 * q, k: fp32 [N][D] row-major.  For every row i: positive pair (q_i, k_i), negative pair
 * (q_i, k_perm(i)) with perm = [1,0,N-1,N-2,...,2].  out fp32 [6][N]:
 *   0 cos(q,k1) 1 cos(q,k0) 2 l2(q,k1) 3 l2(q,k0) 4 1-pearson(q,k1) 5 1-pearson(q,k0).
 * pooled_hw > 1: inputs are [N][C][pooled_hw] feature maps and the spatial mean is taken first
 * (D = C), fused into the same pass.
 * ------------------------------------------------------------------------------------------ */
int msp_rowpair_distances(const float* q, const float* k, int N, long long D, int pooled_hw,
                          float* out, void* stream);
/* scores[m][d][i] = max(0, pos - neg + margins[m]) from the table above; out fp32 [M][3][N] */
int msp_triplet_hinge(const float* dist, int N, const float* margins, int M, float* out,
                      void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimiser-side helpers (train_model.py:93-107): fused multi-tensor L2 norm is left to the host
 * framework in this round.
 * ------------------------------------------------------------------------------------------ */

/* ------------------------------------------------------------------------------------------
 * Small-vector all-reduce over NVLink peer memory (csrc/msp_p2p.cu): the SyncBN [2C] sums and the global Dice sums
 * of a data-parallel step (replaces ~100 tiny NCCL all-reduces per step; the reference's
 * nn.DataParallel, train_model.py:192-194, has no such exchange).  Set-up, once per process: allocate this rank's
 * communication buffer (msp_p2p_buffer_bytes), exchange the 64-byte cudaIpc handles through any host channel, map the
 * peers' buffers.  These four set-up calls are the only ones in the library that allocate or synchronise.
 * msp_p2p_allreduce_sum_f32: data[0..n) <- sum over ranks, in place, one single-CTA kernel on `stream`, graph
 * capturable; every rank must issue the same sequence of calls.  `bufs` = host array of `world` device pointers
 * (entry `rank` = the local buffer), `seq` = one zero-initialised device uint32 owned by this communicator.
 * ------------------------------------------------------------------------------------------ */
long long msp_p2p_buffer_bytes(int world, int max_n);
int msp_p2p_alloc(long long bytes, void** ptr, void* handle64);
int msp_p2p_open(const void* handle64, void** ptr);
int msp_p2p_close(void* ptr);
int msp_p2p_free(void* ptr);
int msp_p2p_allreduce_sum_f32(float* data, int n, int rank, int world, int max_n, void* const* bufs,
                              unsigned* seq, void* stream);
/* SyncBN statistic exchange in ONE launch (C / 32 blocks): `ws` = this rank's [rows][2][C] per-CTA rows (rows == 1: the
 * plain [2][C] sums), added in fixed order (reset != 0: zeroed afterwards), optionally stored (local_add != 0: ADDED) to
 * `local_a` [C] / `local_b` [C] (the rank-local halves = dbeta / dgamma of the backward pass, e.g. straight into
 * param.grad), exchanged with every peer and added in rank order into
 * `global_out` [2][C] (may be NULL when only the finalize is wanted).  mean != NULL: msp_bn_finalize's arithmetic on the
 * global sums with the GLOBAL element count (mean, invstd, running statistics) — functional._BnAct forward.  Replaces
 * msp_reduce_rows -> msp_p2p_allreduce_sum_f32 -> msp_bn_finalize (and the copies around them).  `ticket` = a second
 * zero-initialised device uint32 of the communicator; 2 * C <= max_n. */
int msp_p2p_stats_exchange(float* ws, int rows, int C, float* local_a, float* local_b, int local_add, float* global_out,
                           int reset, double count,
                           float eps, float momentum, float* mean, float* invstd, float* running_mean,
                           float* running_var, int rank, int world, int max_n, void* const* bufs, unsigned* seq,
                           unsigned* ticket, void* stream);

/* ------------------------------------------------------------------------------------------
 * Multi-tensor gradient norm / clip and optimizer step (csrc/msp_optim.cu).  Replaces
 * torch.nn.utils.clip_grad_norm_ (train_model.py:93-98) and torch.optim.SGD / AdamW .step() (train_model.py:107 via
 * optim/optimizer.py:41-48).  Each call takes 1..32 fp32 contiguous tensors as host arrays of device pointers + element
 * counts (the pointer table travels in the kernel's parameter space); callers chunk longer lists.
 * msp_optim_sqnorm ADDS sum(g^2) of its tensors to the device double `sq_accum` (`zero_first`: cleared before this launch);
 * msp_optim_clip: g *= min(1, max_norm / (sqrt(*sq) + 1e-6)); msp_optim_sgd / msp_optim_adamw: torch's update formulas
 * (`first_step`: momentum buffers are initialised to the gradient; `step_dev`: device float, the 1-based step count).
 * ------------------------------------------------------------------------------------------ */
int msp_optim_sqnorm(int n, void* const* grads, const long long* numel, double* sq_accum, int zero_first,
                     void* stream);
/* *norm_out = (float)sqrt(*sq): the value clip_grad_norm_ returns ('gradient_magnitude', train_model.py:100) */
int msp_optim_norm(const double* sq, float* norm_out, void* stream);
/* *scalars[i] += value for 1..32 device floats: the AdamW step counters (one shared counter per parameter age) */
int msp_optim_add_scalar(int n, void* const* scalars, float value, void* stream);
int msp_optim_clip(int n, void* const* grads, const long long* numel, const double* sq, float max_norm, void* stream);
int msp_optim_sgd(int n, void* const* params, void* const* grads, void* const* momentum_bufs, const long long* numel,
                  float lr, float momentum, float dampening, float weight_decay, int nesterov, int first_step,
                  void* stream);
int msp_optim_adamw(int n, void* const* params, void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                    const long long* numel, float lr, double beta1, double beta2, float eps, float weight_decay,
                    const float* step_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * Input side of the step (csrc/msp_input.cu): what the reference does to a batch on the CPU.
 * msp_u8_to_f32_nchw: uint8 [N][C][hw] -> fp32 [N][C*repeats][hw] = (float)((double)x / divisor) with every channel
 *   repeated `repeats` times in place — `np.load(f) / 255` (classification/datasets.py:47), the float32 cast of
 *   ConvertToType (transform/transforms.py:63-103) and RepeatChannels (transform/transforms.py:134-142).
 * msp_color_jitter: torchvision.transforms.ColorJitter on a float [N][C][hw] batch in [0,1], C = 1 or 3, as
 *   robustness/eval.py:61-66 applies it (one parameter draw for the whole batch).  `order_host` = 4 host ints, the
 *   adjustments in application order (0 brightness, 1 contrast, 2 saturation, 3 hue, -1 none); `one_minus_host` = 3 host
 *   floats (1 - factor) of brightness / contrast / saturation rounded as torch rounds its scalar operand; `gray_sums`
 *   = [N] device doubles of workspace (needed when contrast is in the chain).  y may alias x only when contrast is not
 *   in the chain.
 * ------------------------------------------------------------------------------------------ */
int msp_u8_to_f32_nchw(const void* x_u8, int n, int c, long long hw, int repeats, double divisor, float* y,
                       void* stream);
int msp_color_jitter(const float* x, int n, int c, long long hw, const int* order_host, float brightness,
                     float contrast, float saturation, float hue, const float* one_minus_host, double* gray_sums,
                     float* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MSP_B200_H */
