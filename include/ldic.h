/*
 * ldic.h -- C ABI of libldic_b200.so: the B200 (sm_100a) rate-distortion forward
 * path of xiaobucc/learning-driven-image-compression-algorithm.
 *
 * The reference has no FFI of its own (it is 100% Python/PyTorch);
 * the boundary is its nn.Module surface.  Each entry point below names the
 * reference code it replaces (file:line relative to the reference root).
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference
 * side.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless the
 *    name ends in _host;
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*),
 *    never synchronises, never uses the default stream implicitly, and
 *    allocates no device memory: outputs and workspaces are caller-owned;
 *  - returns 0 on success, a negative LDIC_E* code otherwise; the message is
 *    available from ldic_last_error() (thread local);
 *  - activations inside the transforms are NHWC ("channels-last") bf16 unless
 *    stated; module-surface tensors are NCHW fp32 like the reference's.
 */
#ifndef LDIC_H_
#define LDIC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define LDIC_API __attribute__((visibility("default")))
#else
#define LDIC_API
#endif

#define LDIC_OK 0
#define LDIC_EINVAL (-1)   /* bad argument / unsupported shape            */
#define LDIC_ECUDA (-2)    /* a CUDA runtime / driver call failed         */
#define LDIC_ENOTSUP (-3)  /* device is not sm_100                        */

LDIC_API int ldic_version(void);
LDIC_API const char* ldic_last_error(void);
/* Number of kernels this library has launched in this process (bench.py's
 * gpu_launches).  */
LDIC_API long long ldic_launch_count(void);
/* 0 when device `dev` is compute capability 10.x, LDIC_ENOTSUP otherwise. */
LDIC_API int ldic_check_device(int dev);
/* Tuning / diagnostic switches.  They are read from the environment ONCE, when the library is loaded
 * (LDIC_DEBUG_NOSTORE, LDIC_DEBUG_TIMING, LDIC_GDN_INSERT, LDIC_STAGES, LDIC_TAIL_WIDE, LDIC_LIK_GRID, LDIC_FIRST_EPI,
 * LDIC_FIRST_INSERT, LDIC_FIRST_TMA_STORE: first-layer epilogue warps / ring order of its x^2 slots / TMA stores); no
 * entry point calls getenv on its hot path.  ldic_set_tuning changes one at run time (tests, A/B runs): key is
 * the environment name without the LDIC_ prefix in lower case ("tail_wide", "stages", ...); returns the previous
 * value or LDIC_EINVAL.  Changing a switch invalidates the cached launch plans.                            */
LDIC_API int ldic_set_tuning(const char* key, int value);

/* ---- a3: LowerBound + NonNegativeParametrizer ---------------------------------
 * y = max(x, bound)                                   ops/bound_ops.py:21-22
 * dx = (x >= bound || g < 0) ? g : 0                  ops/bound_ops.py:25-27, model/gdn.py:19-26 */
LDIC_API int ldic_lower_bound(const float* x, float bound, float* y, size_t n, void* stream);
LDIC_API int ldic_lower_bound_bwd(const float* x, float bound, const float* grad_out, float* grad_in, size_t n, void* stream);
/* out = max(p, bound)^2 - pedestal                    ops/parametrizers.py:46-49, model/gdn.py:75-81 */
LDIC_API int ldic_nonneg_reparam(const float* p, float bound, float pedestal, float* out, size_t n, void* stream);

/* beta_eff[C], gamma_eff[C*C] (fp32) from the stored (reparametrised) parameters;
 * optionally also the bf16 operand image of gamma used by the tensor-core GDN
 * epilogue: gamma_bf16 is [Np][Kp] row-major (row i = output channel), built as
 * `groups` block-diagonal copies of the CxC matrix, zero elsewhere; beta_tiled is
 * [Np] fp32 (beta repeated per group, 1.0 in the padding).  Pass NULL to skip.
 * model/gdn.py:75-81,140-146; layers/gdn.py:65-67.                               */
LDIC_API int ldic_gdn_prepare(const float* beta_p, const float* gamma_p, int C,
                     float beta_bound, float gamma_bound, float pedestal,
                     float* beta_eff, float* gamma_eff,
                     void* gamma_bf16, float* beta_tiled, int groups, int Np, int Kp,
                     void* stream);

/* ---- a2: stand-alone GDN / IGDN on the module surface (NCHW fp32) -------------
 * y = x * rsqrt(beta + gamma . x^2)   (inverse=0)   or   x * sqrt(...) (inverse=1)
 * layers/gdn.py:62-75, model/gdn.py:69-92,134-156, model/ops.py:106-136.
 * fp32 CUDA-core kernel (bit-faithful op order except the summation order).     */
LDIC_API int ldic_gdn_nchw_f32(const float* x, const float* beta_eff, const float* gamma_eff, float* y,
                      int B, int C, int H, int W, int inverse, int use_rsqrt, void* stream);

/* ---- a6+a7+a8+a9: quantise, Gaussian likelihood, sum(ln L) --------------------
 * One pass over a logical [rows x cols] problem.  Tensor t is addressed as
 *   t[row * t_rs + t_off + col]            (t_rs = row stride in elements)
 * so NHWC channel slices need no copies.  mu / sigma can be per element (mode 2),
 * per column i.e. per channel (mode 1, pointer to `cols` floats) or absent
 * (mode 0: mu = 0).  NCHW tensors with per-channel sigma use rows = B*C, cols = H*W
 * and mode 3 (per row, index row % sigma_period).
 *
 * quant: 0 = v already quantised; 1 = round(v) half-to-even (BypassRound,
 *        model/net.py:416-426); 2 = round(v - mu) + mu (CompressAI "dequantize",
 *        model/net_unet_ha_hs.py:937); 3 = (round(v) - v) + v (ste_round,
 *        ops/ops.py:34).
 * form:  0 = GaussianModel  Phi((v-mu+.5)/s) - Phi((v-mu-.5)/s), Phi via erf,
 *            clamp(min=lik_bound), NO sigma bound (model/net.py:272-286);
 *        1 = GaussianConditional: |v-mu|, s = max(s, scale_bound), erfc form,
 *            max(L, lik_bound) (model/Net_unet.py:1057).
 * sigma_is_log: sigma tensor holds log-sigma; exp() is applied first
 *            (model/net.py:316).
 * Outputs (any may be NULL): v_hat (fp32, same addressing as v_hat_rs/off),
 * v_hat_bf16 (bf16 copy for the synthesis transform), lik (fp32, dense
 * [rows x cols]).  sum_ln_out[0] receives sum(ln L) over the whole problem
 * (deterministic two-level reduction; NaNs propagate like the reference).
 * `workspace` must hold ldic_likelihood_workspace_bytes() bytes and be zeroed
 * once before the first use (the kernel leaves it zeroed).                      */
typedef struct {
  const float* v;      long long v_rs;      long long v_off;
  const float* mu;     long long mu_rs;     long long mu_off;     int mu_mode;
  const float* sigma;  long long sigma_rs;  long long sigma_off;  int sigma_mode;
  int sigma_period;    /* mode 3: sigma index = row % sigma_period */
  long long rows, cols;
  int quant, form, sigma_is_log;
  float lik_bound, scale_bound;
  float* v_hat;        long long v_hat_rs;  long long v_hat_off;
  void* v_hat_bf16;    long long vb_rs;     long long vb_off;
  float* lik;
  float* sum_ln_out;
  void* workspace;
} LdicLikelihoodArgs;
LDIC_API size_t ldic_likelihood_workspace_bytes(void);
LDIC_API int ldic_round_likelihood_bpp(const LdicLikelihoodArgs* args, void* stream);

/* ---- a11: MSE / PSNR on 8-bit levels --------------------------------------------
 * gt = round((x+1)*127.5); xh = round(clamp((xt+1)*127.5, 0, 255));
 * sq_err[b] = sum((xh-gt)^2) as an exact 64-bit integer; the caller divides by
 * C*H*W and forms 20*log10(255/sqrt(mse)).  model/net.py:864-869;
 * clamp_pm1 != 0 adds clamp(xt,-1,1) first (model/net_unet_ha_hs.py:1006).
 * sq_err must be zeroed by the caller (it is accumulated into).                */
LDIC_API int ldic_mse_sum(const float* x, const float* x_tilde, int B, long long chw, int clamp_pm1,
                 unsigned long long* sq_err, void* stream);
/* Scalar tail of Net.forward (model/net.py:856-869) in two launches instead of ~25 elementwise ones:
 *   ldic_rd_pack_metrics: bits3 = [sum ln L_z, sum ln L_y, sum ln L_syntax] (float, device), sq_err[B] ->
 *     packed5 = [the three sums, sum_i 20 log10(255 / sqrt(sq_err_i / chw)), B] in double (the five numbers the
 *     multi-GPU all-reduce carries) and v_mse[B] = sq_err / chw (float, may be NULL);
 *   ldic_rd_finish_metrics: bpp_psnr[0] = (packed5[0]+[1]+[2]) / (-ln2 * packed5[4] * pixels_per_image)  (:856-861),
 *     bpp_psnr[1] = packed5[3] / packed5[4]  (:869, mean of the per-image PSNR).                              */
LDIC_API int ldic_rd_pack_metrics(const float* bits3, const unsigned long long* sq_err, int B, long long chw,
                         double* packed5, float* v_mse, void* stream);
LDIC_API int ldic_rd_finish_metrics(const double* packed5, double pixels_per_image, float* bpp_psnr, void* stream);
/* Fused tail of Net.forward: per-image 1x1 conv (batch_conv, model/net.py:527-537)
 * of the 16-channel NHWC fp32 synthesis output with weights w[B][3][M], then the
 * a11 arithmetic against the NCHW fp32 input x.  x_tilde_nchw (optional) receives
 * the reconstruction.                                                           */
LDIC_API int ldic_syntax_conv_mse(const float* x_nchw, const float* xt_nhwc, const float* w, int B, int M,
                         int H, int W, int tanh_out, float* x_tilde_nchw, unsigned long long* sq_err, void* stream);

/* ---- layout / dtype glue ------------------------------------------------------------ */
/* NCHW fp32 -> NHWC bf16 (channels padded to Cp with zeros) and back.            */
LDIC_API int ldic_nchw_f32_to_nhwc_bf16(const float* x, void* y, int B, int C, int H, int W, int Cp, int apply_abs, void* stream);
LDIC_API int ldic_nhwc_to_nchw_f32(const void* x, int x_is_bf16, float* y, int B, int C, int H, int W, int Cp, void* stream);
/* uint8 image levels -> fp32 x = (u/255)*2-1: the reference's input map (torchvision ToTensor, then
 * eval_net.py:84), the same correctly rounded fp32 operations.  Only needed where the first layer is not the
 * fused LDIC_CONV_FIRST_5x5S2 kernel (which takes the uint8 image directly).                              */
LDIC_API int ldic_u8_to_f32_pm1(const unsigned char* x, float* y, size_t n, void* stream);
/* y (NHWC fp32, C channels) -> optional outputs: round(y) bf16, |y| bf16, round(y) fp32.
 * model/net.py:197 (abs), :676/:741 (round).                                     */
LDIC_API int ldic_latent_prep(const float* y, size_t n, void* y_round_bf16, void* y_abs_bf16, float* y_round_f32, void* stream);
/* Context-model input image [P][2N] bf16 = round(y) (bf16, N ch) | bf16(h2) (fp32 in, N ch)
 * model/net.py:307-309.                                                          */
LDIC_API int ldic_ctx_pack_input(const void* y_round_bf16, const float* h2, void* x, long long P, int N, void* stream);
/* First-layer patch matrix: x NCHW fp32 (B,3,H,W) -> A[B*Ho*Wo][Kp] bf16 with
 * A[p][(ky*5+kx)*Cin+ci] = x[ci, 2oy+ky-1, 2ox+kx-1] (zero outside), zero for
 * k >= 25*Cin.  model/net.py:97-98 (ZeroPad2d((1,2,1,2)) + Conv2d k5 s2).       */
LDIC_API int ldic_im2col_5x5s2(const float* x, void* a, int B, int Cin, int H, int W, int Kp, void* stream);
/* fp32 patch matrix for the TF32 parity mode (LdicConvDesc.precision = 1) */
LDIC_API int ldic_im2col_5x5s2_f32(const float* x, float* a, int B, int Cin, int H, int W, int Kp, void* stream);

/* ---- a1/a4/a5/a10: tensor-core implicit-GEMM convolutions ---------------------------- */
enum {
  LDIC_CONV_S2_5x5_P12 = 0,   /* ZeroPad2d((1,2,1,2)) + Conv2d(k5,s2,p0)      model/net.py:97-112   */
  LDIC_CONV_S2_5x5_P2 = 1,    /* Conv2d(k5,s2,p2)                             model/net.py:191-193  */
  LDIC_CONV_S1_3x3_P1 = 2,    /* Conv2d(k3,s1,p1)                             model/net.py:189      */
  LDIC_CONV_1x1 = 3,          /* GEMM over pixels (first layer after im2col)                        */
  LDIC_DECONV_GS_5x5 = 4,     /* ZeroPad2d((1,0,1,0)) + ConvTranspose2d(k5,s2,p3,op1) model/net.py:128-142 */
  LDIC_DECONV_HS_5x5 = 5,     /* ConvTranspose2d(k5,s2,p2,op1)                model/net.py:207-209  */
  LDIC_DECONV_S1_3x3 = 6,     /* ConvTranspose2d(k3,s1,p1)                    model/net.py:211      */
  LDIC_DECONV_GS_5x5_MERGED = 7, /* same as 4 with the 4 sub-pixel phases merged into N (small Cout) */
  /* Context model PredictionModel_Context (model/net.py:289-319).  The reference materialises one
   * 4x4 patch per latent position with a one-hot 7x7 conv (BlockSample, :219-242); here the first
   * conv gathers the patch cells straight from the latent images by TMA, 16 jobs = 16 patch cells. */
  LDIC_CTX_CONV1 = 8,  /* x [B,h,w, round(y)(N) | h2(N)] bf16 -> [B*h*w,4,4,N]; aux0=N, aux1=M; Conv2d(2N-M,N,3,1,1) :295.
                          aux1 = M | s << 8 | x_org << 12 | w_in << 16.  s > 0: x is stored SHEARED by s columns per row --
                          patch cell (i,j) of the pixel at (y,x) is read from column x + j - 2 + s (i - 3); w_in > 0: x is
                          w_in columns wide and the W output columns start at its column x_org (the wavefront decoder
                          computes one column of a 10-column band); aux0 = N | pitch << 12 with pitch > 0: that band is a
                          column slice of a wider image with `pitch` pixels per row                                    */
  LDIC_CTX_CONV2 = 9,  /* [P,4,4,N] -> [P,2,2,N]   Conv2d(N,N,3,2,1)  :297                                           */
  LDIC_CTX_CONV3 = 10, /* [P,2,2,N] -> [P,2,2,N]   Conv2d(N,N,3,1,1)  :299                                           */
  LDIC_CTX_FC = 11,    /* [P,2,2,N] -> [P,1,2,Cout_pad] fp32 (mu | log sigma), Linear(4N, 2*Cout) :302; Cout = N-M   */
  /* First analysis layer fused with its GDN: x is the NCHW fp32 IMAGE (B,3,H,W) itself (not NHWC bf16);
   * ZeroPad2d((1,2,1,2)) + Conv2d(3,Cout,5,2) (+GDN) -> NHWC bf16, no patch matrix in HBM.  Cin = 3,
   * Cin_pad = 128 (K = 75 padded), Cout_pad <= 192.  model/net.py:97-99.
   * aux0 = 1: x is the uint8 image (B,3,H,W) of 8-bit levels u; the kernel applies the reference's input map
   * x = (u/255)*2-1 (ToTensor + eval_net.py:84) while it builds the patches, so the host->device copy carries
   * 1 byte per sample instead of 4.  W % 16 == 0.                                                  */
  LDIC_CONV_FIRST_5x5S2 = 12
};
enum { LDIC_ACT_NONE = 0, LDIC_ACT_RELU = 1, LDIC_ACT_LEAKY02 = 2, LDIC_ACT_GDN = 3, LDIC_ACT_IGDN = 4,
       LDIC_ACT_LEAKY001 = 5 /* nn.LeakyReLU() default slope 0.01: CompressAI ResidualBlock, layers/layers.py:87-102 */ };

typedef struct {
  int kind;           /* LDIC_CONV_* */
  int B, H, W;        /* input (NHWC) batch and spatial size                     */
  int Cin, Cout;      /* logical channel counts of the layer                     */
  int Cin_pad;        /* channels of the input tensor in memory (multiple of 64) */
  int Cout_pad;       /* channels of the output tensor in memory                 */
  int act;            /* LDIC_ACT_*                                              */
  int out_f32;        /* 0: bf16 NHWC output, 1: fp32 NHWC output                */
  int aux0, aux1;     /* kind specific (LDIC_CTX_CONV1: N, M; LDIC_CONV_FIRST_5x5S2: aux0 = 1 for a uint8 image), else 0 */
  int sm_limit;       /* SM partition for concurrent streams: 0 = all SMs; n > 0 = at most n SMs; n < 0 = leave |n| SMs
                         free (the kernels are persistent, one CTA per SM, so the grid size IS the SM footprint)  */
  int precision;      /* 0: bf16 operands (product path).  1: TF32 parity mode (kind::tf32, about half the rate): x and y are
                         fp32 NHWC, ldic_conv_pack_weights writes fp32 (tf32-rounded) weights, gamma is the fp32 [Cout][Cout]
                         gamma_eff of ldic_gdn_prepare.  Kinds 0-3 with 128 / 192 output channels (g_a, h_a): SURVEY H4's
                         precision comparison -- how many symbols flip because of bf16 operands.  1 rounds the fp32
                         outputs to tf32 (a layer that feeds another TF32 layer: the tensor core would truncate), 2 keeps
                         them as they are (the layer that feeds the quantiser).                                        */
} LdicConvDesc;

/* Elements (bf16) of the packed weight image for this layer, and the packer:
 * w is the state-dict tensor, (Cout,Cin,k,k) for Conv2d or (Cin,Cout,k,k) for
 * ConvTranspose2d; cin_offset places the Cin logical channels inside Cin_pad
 * (used to feed the full 192-channel latent to the 176-channel g_s input,
 * model/net.py:726).  bias_packed is [Np] fp32.                                */
LDIC_API long long ldic_conv_weight_elems(const LdicConvDesc* d);
LDIC_API int ldic_conv_n_cols(const LdicConvDesc* d);  /* Np: accumulator columns          */
LDIC_API int ldic_conv_bias_elems(const LdicConvDesc* d);  /* floats in bias_packed        */
/* output tensor dims [B', Ho, Wo, C] (NHWC) */
LDIC_API int ldic_conv_out_dims(const LdicConvDesc* d, int* dims4);
LDIC_API int ldic_conv_pack_weights(const LdicConvDesc* d, const float* w, const float* bias, int cin_offset,
                           void* w_packed, float* bias_packed, void* stream);
LDIC_API void ldic_conv_out_shape(const LdicConvDesc* d, int* Ho, int* Wo);
/* y = act(conv(x) + bias); for LDIC_ACT_GDN / IGDN the normalisation
 * x * rsqrt|sqrt(beta + gamma . x^2) runs in the epilogue on the tensor cores
 * (gamma_bf16 / beta_tiled from ldic_gdn_prepare).                              */
LDIC_API int ldic_conv_forward(const LdicConvDesc* d, const void* x, const void* w_packed, const float* bias_packed,
                      const void* gamma_bf16, const float* beta_tiled, void* y, void* stream);
/* y = act(conv(x) + bias) + r: the residual connection of a ResidualBlock (CompressAI, used by layers/layers.py:87-102) or of
 * WinBasedAttention (layers/win_attention.py:204-205) fused into the epilogue.  r is an NHWC bf16 tensor of y's shape; plain
 * layers only (no GDN, one job, at most 192 output channels).                                                        */
LDIC_API int ldic_conv_forward_residual(const LdicConvDesc* d, const void* x, const void* w_packed, const float* bias_packed,
                               const void* residual_bf16, void* y, void* stream);
/* Launch plans (tile / tap tables, TMA descriptors, kernel variant, grid) are cached per (descriptor, parameter
 * tensors, device); the activation addresses of a call are patched into a copy of the cached plan, so a repeated
 * call is a lookup, one cuTensorMapReplaceAddress and one kernel launch.                                      */
LDIC_API int ldic_conv_plan_cache_size(void);
LDIC_API void ldic_conv_plan_cache_clear(void);
/* The merged last synthesis deconv (LDIC_DECONV_GS_5x5_MERGED) with the tail of Net.forward fused into its
 * epilogue: per-image 1x1 conv of the M-channel IGDN output with w[B][3][M] (batch_conv, model/net.py:527-537,
 * :811) and the a11 squared level error against the NCHW input image (model/net.py:864-868), so the
 * M-channel full-resolution tensor never makes a round trip through HBM.  y_or_null (optional) still receives
 * the [B,2H,2W,M] fp32 NHWC tensor; x_tilde_nchw (optional) the reconstruction; sq_err[B] is accumulated
 * into (zero it first).  H, W of the tail are the image size (= 2 x the layer's input size).            */
typedef struct {
  const void* x_nchw;         /* fp32 image in [-1,1], or its uint8 levels when x_is_u8 */
  const float* w;
  float* x_tilde_nchw;
  unsigned long long* sq_err;
  int H, W;
  int x_is_u8;
  int tanh_out;               /* x~ = tanh(batch_conv(...)) (U-Net family, model/net_unet_ha_hs.py:980) */
} LdicConvTail;
LDIC_API int ldic_conv_forward_fused_tail(const LdicConvDesc* d, const void* x, const void* w_packed, const float* bias_packed,
                                 const void* gamma_bf16, const float* beta_tiled, void* y_or_null,
                                 const LdicConvTail* tail, void* stream);
/* Plain CUDA-core fp32 direct convolution of the same layer kinds on NHWC fp32
 * tensors (validation aid for the tensor-core path at sizes the CPU oracle
 * cannot reach; not used by the product forward).                               */
LDIC_API int ldic_conv_forward_f32_reference_kernel(const LdicConvDesc* d, const float* x_nhwc, const float* w,
                                           const float* bias, float* y_nhwc, void* stream);

/* ---- the syntax side branch (SURVEY 8 "items on the forward", f4) -----------------------------------
 * Syntax_Model (model/net.py:349-375), PredictionModel_Syntax (:378-413) and conv_generator
 * (:322-343) in five fp32 launches.  Inputs are the NHWC fp32 latent y [B,h,w,N] (the first M
 * channels are the syntax channels, :712) and the NHWC fp32 h_s output h2 [B,h,w,N].  Weights are the
 * state-dict tensors as they are (Conv2d: [Cout][Cin][3][3], Linear / 1x1 conv: [out][in]).
 * Scratch (caller-owned, fp32): sm_ds1 [B,h1,w1,32], sm_ds2 [B,h2,w2,64], ps_ds0 [B,h1,w1,M],
 * ps_ds1 [B,h2,w2,M] with h1 = (h-1)/2+1, h2 = (h1-1)/2+1 (same for w); pool_part
 * [ldic_syntax_workspace_elems(B,h,w,N,M)] (partial sums of the six mean pools + re-packed 3x3 weights).
 * Outputs: z3 [B,M] (Syntax_Model output), z3_round = round(z3) (:753), mu / sigma [B,M] (the two
 * returns of PredictionModel_Syntax, sigma = exp(.)), conv_w [B,3,M] (per-image 1x1 filters, :805). */
typedef struct {
  int B, h, w, N, M;
  const float* y;
  const float* h2;
  const float *sm_down0_w, *sm_down0_b, *sm_down1_w, *sm_down1_b, *sm_conv_w, *sm_conv_b;
  const float *ps_down0_w, *ps_down0_b, *ps_down1_w, *ps_down1_b, *ps_fc_w, *ps_fc_b;
  const float *cg_w0, *cg_b0, *cg_w1, *cg_b1, *cg_w2, *cg_b2;
  float *sm_ds1, *sm_ds2, *ps_ds0, *ps_ds1, *pool_part;
  float *z3, *z3_round, *mu, *sigma, *conv_w;
  const float* z3_round_in;   /* optional [B,M]: the decoder's path -- conv_generator runs on these (entropy-decoded) symbols
                                 instead of round(Syntax_Model(y)); z3 / z3_round are still written from y */
} LdicSyntaxArgs;
LDIC_API long long ldic_syntax_workspace_elems(int B, int h, int w, int N, int M);
LDIC_API int ldic_syntax_branch(const LdicSyntaxArgs* args, void* stream);

/* ---- a12: progressive trit-plane quantisation + per-plane likelihood (BASELINE configs[4]) ----------
 * Builder-defined extension: the reference's model/Trit_Plane.py:25-57 crashes and holds neither
 * trit planes nor a likelihood, so there is NO reference oracle for this entry point beyond the quantiser /
 * Gaussian mass it shares with a6/a7 ("parity unpinned").
 *   q = clamp(round(v - mu), -H, H), H = (3^L - 1)/2;  q + H = sum_l t_l 3^l (t_l in {0,1,2});
 *   L_l = P(t_l | more significant trits) from the interval thirds of N(mu, max(sigma, scale_bound)), >= lik_bound.
 * planes: int8 [L][n] (plane L-1 most significant) or NULL; q_out: int32 [n] or NULL; sum_ln_per_plane: float [L];
 * workspace: ldic_tritplane_workspace_bytes() bytes, zeroed once before the first use.  mu may be NULL (= 0). */
LDIC_API size_t ldic_tritplane_workspace_bytes(void);
LDIC_API int ldic_tritplane_likelihood(const float* v, const float* mu, const float* sigma, long long n, int L,
                              float scale_bound, float lik_bound, signed char* planes, int* q_out,
                              float* sum_ln_per_plane, void* workspace, void* stream);

/* ---- f4: rANS entropy coder over the quantised latents --------------------------------------
 * Builder-defined extension: the reference only ESTIMATES the rate (model/net.py:856-861) and holds no entropy coder
 * or bitstream, so there is no reference oracle for these entry points ("parity unpinned" against the reference).
 * They are pinned by (1) decode(encode(x)) == x bit for bit, (2) byte-for-byte equality with the CPU restatement
 * oracle/rans_ref.py, (3) 8 * bytes within a fraction of a percent (+ header) of the sum(-log2 L) that
 * ldic_round_likelihood_bpp returns for the same symbols (tests/test_gpu_rans.py).
 * Addressing as in LdicLikelihoodArgs (rows x cols, row strides, broadcast modes 0..3; mode 3 of mu shares
 * sigma_period).  rows_per_segment rows form one segment = one independent bitstream (one image); every segment
 * holds seg_elems = rows_per_segment * cols symbols, interleaved over `streams` rANS states (symbol i -> stream
 * i % streams, at most 65535 symbols per stream; more streams = more parallelism, 6 bytes of header each).
 * quant: 1 = symbols round(v), model N(mu, sigma) (GaussianModel, model/net.py:272-286,:741);
 *        2 = symbols round(v - mu), model N(0, sigma), decoder returns symbol + mu (model/net_unet_ha_hs.py:937).
 * Integer model (16-bit frequencies from a 24-bit normal-CDF table, window of max(15, 2 + ceil(6 sigma)) integers each side
 * of rint(mu), the rest escaped out of band) and byte layout: header of csrc/rans.cu.
 * encode: out = segments x out_stride bytes (out_stride >= ldic_rans_max_bytes for a guaranteed fit, multiple of 4),
 *   sizes[segment] = bytes written (0 if it did not fit), status[segment] = 0 or a bit set of
 *   1 = a symbol was NaN / beyond 2^30 (coded as 0), 2 = out_stride too small, 4 = bad header, 8 = corrupt stream,
 *   16 = decode_ranges called out of order / with a range that crosses streams.
 * decode: v_hat[row * v_hat_rs + v_hat_off + col] receives the symbols (fp32); args->v is ignored.
 * workspace: ldic_rans_workspace_bytes(segments, seg_elems, streams) bytes, 256-byte aligned, no initialisation.  */
typedef struct {
  const float* v;      long long v_rs;      long long v_off;
  const float* mu;     long long mu_rs;     long long mu_off;     int mu_mode;
  const float* sigma;  long long sigma_rs;  long long sigma_off;  int sigma_mode;
  int sigma_period;
  long long rows, cols, rows_per_segment;
  int quant, sigma_is_log;
  float scale_bound;   /* sigma = max(sigma, scale_bound) when > 0 (GaussianConditional: 0.11) */
  int streams;
  int col_groups;      /* G > 1: the segment is coded group by group (all rows' columns [0, cols/G), then the next cols/G, ...),
                          so the symbols of one row lie in G different streams; 0 / 1: row-major order */
} LdicRansArgs;
LDIC_API size_t ldic_rans_max_bytes(long long seg_elems, int streams);
LDIC_API size_t ldic_rans_workspace_bytes(long long segments, long long seg_elems, int streams);
LDIC_API int ldic_rans_encode(const LdicRansArgs* args, unsigned char* out, long long out_stride, unsigned int* sizes,
                     unsigned int* status, void* workspace, void* stream);
LDIC_API int ldic_rans_decode(const LdicRansArgs* args, const unsigned char* in, long long in_stride, const unsigned int* sizes,
                     float* v_hat, long long v_hat_rs, long long v_hat_off, unsigned int* status, void* workspace,
                     void* stream);
/* Incremental decoding, for decoders whose (mu, sigma) depend on symbols decoded earlier (the causal context model of
 * model/net.py:289-319 walked along wavefronts, ldic_b200.Net.decompress).  decode_begin validates the segments and
 * fills `state` (16 bytes per stream and segment: segments * streams * 16 bytes, 16-byte aligned) with the streams'
 * cursors; every decode_ranges call decodes, in every segment, the nranges symbol ranges [ranges[2j], ranges[2j] +
 * ranges[2j+1]) (device array of int pairs; count <= 0 = skip).  A range must lie inside ONE stream and start exactly
 * where that stream stopped (else status bit 16); args->mu / sigma must be valid for the symbols of the ranges at the
 * time of the call.  v_hat_bf16 (optional) receives a bf16 copy with its own row addressing.  Escaped symbols inside a
 * range are written in the same call; a stream is checked against its final state when its last symbol is decoded.
 * param_row_map / bf16_row_map (optional device arrays of `rows` ints): the per-element (mode 2) mu / sigma of row r
 * are read from row param_row_map[r], and the bf16 copy of row r goes to row bf16_row_map[r] -- a wavefront decoder
 * keeps the parameters of the current wavefront only and the rounded latent in a sheared image (Net.decompress).     */
LDIC_API int ldic_rans_decode_begin(const LdicRansArgs* args, const unsigned char* in, long long in_stride,
                           const unsigned int* sizes, void* state, unsigned int* status, void* workspace, void* stream);
LDIC_API int ldic_rans_decode_ranges(const LdicRansArgs* args, const unsigned char* in, long long in_stride, void* state,
                            const int* ranges, int nranges, float* v_hat, long long v_hat_rs, long long v_hat_off,
                            void* v_hat_bf16, long long vb_rs, long long vb_off, const int* param_row_map,
                            const int* bf16_row_map, unsigned int* status, void* stream);
/* the 24-bit normal-CDF table of the format: T[i] = round(Phi(-8 + i/128) * 2^24), *entries = 2049 (host memory) */
LDIC_API const unsigned int* ldic_rans_phi_table(int* entries);

/* ---- f2: window attention (layers/win_attention.py:38-209; blocks of layers/layers.py:56-111) --------------
 * WinBasedAttention.forward = x + proj(W-MSA(qkv(x))) on 8x8 (or 4x4) windows with an optional cyclic shift.  The
 * qkv and proj Linears are 1x1 convs (ldic_conv_forward, LDIC_CONV_1x1: three of them, q pre-scaled by
 * head_dim^-0.5, :108); the entry points below are what lies between and around them.
 *   ldic_window_attention_bias: bias[h][i][j] = table[index[i][j]][h]  (:101-104); table float [(2ws-1)^2][heads],
 *     index int64 [ws^2][ws^2] (the module's relative_position_index buffer), bias float [heads][ws^2][ws^2].
 *   ldic_window_attention_core: q, k, v, out: bf16 NHWC token images [B][H][W][C]; for every window and head
 *     out = softmax(q k^T + bias + shift mask) v (:106-123).  window_partition / window_reverse (:6-36), torch.roll
 *     and its reverse (:181-200) are index arithmetic; the 0 / -100 mask of :160-177 is derived from the token
 *     coordinates.  C % heads == 0, C % 8 == 0, head_dim even and <= 32, heads <= 16, ws in {4, 8} dividing H, W.
 *   ldic_residual_nhwc_to_nchw_f32: y = shortcut + o (:204-205), o NHWC fp32 with Cp >= C channels per pixel.   */
LDIC_API int ldic_window_attention_bias(const float* table, const long long* index, float* bias, int heads, int ws,
                               void* stream);
LDIC_API int ldic_window_attention_core(const void* q, const void* k, const void* v, const float* bias, void* out, int B,
                               int H, int W, int C, int heads, int ws, int shift, void* stream);
/* Same with the module's relative_position_bias_table [(2ws-1)^2][heads] itself (layers/win_attention.py:64): the kernel
 * keeps the table in shared memory and derives bias(i, j) from the two tokens' window coordinates, instead of streaming
 * the gathered [heads][N][N] tensor (128 KB per window at 8 heads, 8x8 windows) through L2.                      */
LDIC_API int ldic_window_attention_core_table(const void* q, const void* k, const void* v, const float* table, void* out, int B,
                                     int H, int W, int C, int heads, int ws, int shift, void* stream);
LDIC_API int ldic_residual_nhwc_to_nchw_f32(const float* o_nhwc, const float* shortcut_nchw, float* y_nchw, int B, int C,
                                   int H, int W, int Cp, void* stream);

/* y (NCHW fp32) = x (NCHW fp32) + a * sigmoid(b) with a, b NHWC bf16 [B,H,W,Cp]: the gate and residual that close
 * Win_noShift_Attention (layers/layers.py:104-111), fused with the return to the module surface's layout.         */
LDIC_API int ldic_gate_residual_nhwc_to_nchw_f32(const void* a_nhwc_bf16, const void* b_nhwc_bf16, const float* x_nchw, float* y_nchw,
                                        int B, int C, int H, int W, int Cp, void* stream);

/* Diagnostics: every in-kernel barrier wait of the conv kernels is bounded; a starved wait records
 * {flag, block, thread, barrier byte offset in dynamic shared memory, parity} in host-mapped memory and
 * traps.  Returns 1 and fills out5 when a timeout has been recorded in this process, else 0 (host call,
 * valid even after the CUDA context reported the error).                                           */
LDIC_API int ldic_debug_last_timeout(unsigned long long* out5);

#ifdef __cplusplus
}
#endif
#endif /* LDIC_H_ */
