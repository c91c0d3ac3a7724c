/* lsps_b200 -- C ABI of the B200 (sm_100a) kernel library behind the LSPS training step.
 *
 * The reference (masabdi/LSPS) has no native layer: its "L0" is torch.nn / ATen called implicitly from
 *   src/trainers/common_net.py:160-181,221-268   (LeakyINSResBlock, LeakyReLUConv2d, LeakyReLUConvTranspose2d, ...)
 *   src/trainers/lsps_nets.py:34-83,86-160,164-272 (poseVAE, SharedDis, SharedResGen)
 *   src/trainers/lsps_trainer.py:55-262           (losses, Adam steps)
 * Each entry point below names the reference op site(s) it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (16-byte aligned); nothing is retained after the call
 *     except cached TMA descriptors keyed by (pointer, shape);
 *   - activations are NHWC; "bf16" tensors are __nv_bfloat16, everything else is float32 unless stated;
 *   - conv weights are "packed": [tap = r*3+s][Cout][Cin] (forward operand) and [tap][Cin][Cout] (dgrad operand),
 *     taps indexed in the FORWARD op's kernel coordinates (for ConvTranspose2d: its own (r,s));
 *   - all work is asynchronous on the given stream (capturable in a CUDA graph); no host synchronisation;
 *   - return value 0 = ok, <0 = LSPS_E_*; text via lsps_last_error().  One ctx per (process, device).
 */
#ifndef LSPS_B200_H
#define LSPS_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct lsps_ctx lsps_ctx;
typedef void* lsps_stream; /* cudaStream_t */

enum { LSPS_OK = 0, LSPS_E_ARG = -1, LSPS_E_SHAPE = -2, LSPS_E_ARCH = -3, LSPS_E_CUDA = -4 };

/* conv kinds: 3x3 stride-1 pad-1 Conv2d | 3x3 stride-2 pad-1 Conv2d | 3x3 stride-2 pad-1 output_padding-1 ConvTranspose2d
   | 4x4 stride-2 pad-1 ConvTranspose2d (Mapping net, lsps_nets.py:17-23; 16 taps, tap = r*4+s) */
enum { LSPS_CONV_S1 = 0, LSPS_CONV_S2 = 1, LSPS_DECONV_S2 = 2, LSPS_DECONV4_S2 = 3,
       /* 1x1 stride-1 Conv2d (LeakyINSResNeXtBlock, common_net.py:116,122): one tap, weights [Cout][Cin] */
       LSPS_CONV1X1 = 4 };
/* epilogue flags of the implicit-GEMM kernels, applied in this order: +bias, LeakyReLU, +add, *lrelu'(mask) */
enum { LSPS_EP_BIAS = 1, LSPS_EP_LRELU = 2, LSPS_EP_MASK = 4, LSPS_EP_ADD = 8,
       /* forward: accumulate per-(image, channel) sum / sum of squares of the fp32 result (after bias) -- the statistics of
          the InstanceNorm2d / BatchNorm2d that follows the conv (common_net.py:168,171,187,190) -- into ext->sums */
       LSPS_EP_STATS = 16,
       /* data gradient landing on a = lrelu(IN(h)): stores g = acc * lrelu'(xhat) and accumulates per-(image, channel)
          sum g, sum g*xhat into ext->bsums (front half of InstanceNorm backward, finished by lsps_norm_bwd_apply);
          xhat is recovered from the stored activation: xhat = a > 0 ? a : a / slope */
       LSPS_EP_INBWD = 32 };

/* n images; h,w = INPUT spatial size of the FORWARD op; cin/cout of the forward op (multiples of 64). */
typedef struct { int kind, n, h, w, cin, cout; } lsps_conv_shape;

/* Optional extras of lsps_conv_{fwd,dgrad}_ex; zero-initialise, set what is used.
     grouped launch : images [0, n_split) use (w, bias), images [n_split, n) use (w2, bias2)   (see lsps_conv_fwd_grouped)
     LSPS_EP_STATS  : sums  f32 [n][2][cout]  (zeroed by the call, then red.add'ed by the kernel)
     LSPS_EP_INBWD  : in_a bf16 [n,h,w,cin] = lrelu(IN(h)), the activation the gradient lands on; bsums f32 [n][2][cin]
                      (zeroed by the call)
     split          : "bf16x3" operands for layers whose bf16 rounding shows in the losses (the discriminator stack,
                      lsps_nets.py:102-126): activations are stored as [n,h,w,2c] = (bf16 hi | bf16 lo) channel halves with
                      hi + lo carrying 16 mantissa bits, weights as two tensors (w = hi, w_lo = lo, same layout); the GEMM
                      accumulates hi*hi + hi*lo + lo*hi in fp32 and writes its result in the same split form.  mask (if
                      any) is a split tensor too (its hi half carries the sign). */
typedef struct {
  const void* w2; const float* bias2; int n_split;
  float* sums;
  const void* in_a; float* bsums;
  const void* w_lo; int split;
  int groups;   /* > 1: grouped 3x3 stride-1 conv (ResNeXt cardinality, common_net.py:118): cin == cout, group width
                   cin/groups in {64, 128}; weights [tap][Cout][cin/groups] (forward) and [tap][Cin][cout/groups]
                   (data gradient).  lsps_conv_wgrad_grouped is the matching weight gradient. */
  /* decoder head fused into the epilogue of the LAST transposed conv (ConvTranspose2d(128,64,3,2,1,1) + LeakyReLU ->
     ConvTranspose2d(64,1,1) + Tanh, lsps_nets.py:222-229): with head_out != NULL the launch also writes
     head_out[pixel] = tanh(sum_c head_w[c] * y[pixel][c] + head_b[0]) (y as stored, i.e. bf16-rounded) and, with
     head_target != NULL, the L1 reconstruction term of lsps_head_fwd_l1 for pixels [head_t0, head_t0 + head_tn).
     Only the fused up-sampling kernel (forward, Cout = 64, Cin <= 128, no second weight set) honours it; any other
     shape returns LSPS_E_ARG instead of silently skipping the head. */
  const float* head_w; const float* head_b; float* head_out;
  const float* head_target; long long head_t0, head_tn; float head_scale; float* head_dout; float* head_acc;
} lsps_conv_ext;

int lsps_ctx_create(lsps_ctx** out, int device);
void lsps_ctx_destroy(lsps_ctx* ctx);
const char* lsps_last_error(lsps_ctx* ctx);
int lsps_abi_version(void);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
long long lsps_launch_count(lsps_ctx* ctx);

/* ---- tcgen05 implicit-GEMM convolutions (nn.Conv2d / nn.ConvTranspose2d forward, common_net.py:246-268,160-175) */
/* y = epilogue(conv(x, w) [+ bias]) ; x bf16 [n,h,w,cin] ; y bf16 [n,ho,wo,cout] */
int lsps_conv_fwd(lsps_ctx*, const lsps_conv_shape*, const void* x, const void* w_fwd, const float* bias, void* y,
                  int flags, float slope, lsps_stream);
/* dx = (conv_backward_data(dy, w) + add) * lrelu'(mask) ; mask/add: bf16 tensors shaped like dx (flags ADD / MASK) */
int lsps_conv_dgrad(lsps_ctx*, const lsps_conv_shape*, const void* dy, const void* w_dgrad, void* dx,
                    const void* mask, const void* add, int flags, float slope, lsps_stream);
/* Two weight sets in ONE launch: images [0, n_split) use (w, bias), images [n_split, n) use (w2, bias2).  The reference
   runs encode_A / encode_B (and decode_B / decode_A in the cycle pass) as separate nn.Sequential calls on half batches
   (lsps_nets.py:245-272); on the concatenated batch the GEMM fills whole waves of CTA pairs.  w2 must lie in the same
   allocation as w, a whole number of GEMM-K rows (cin resp. cout elements) away (either direction). */
int lsps_conv_fwd_grouped(lsps_ctx*, const lsps_conv_shape*, const void* x, const void* w_fwd, const float* bias,
                          const void* w_fwd2, const float* bias2, int n_split, void* y, int flags, float slope, lsps_stream);
int lsps_conv_dgrad_grouped(lsps_ctx*, const lsps_conv_shape*, const void* dy, const void* w_dgrad, const void* w_dgrad2,
                            int n_split, void* dx, const void* mask, const void* add, int flags, float slope, lsps_stream);
int lsps_conv_fwd_ex(lsps_ctx*, const lsps_conv_shape*, const void* x, const void* w_fwd, const float* bias, void* y,
                     int flags, float slope, const lsps_conv_ext* ext, lsps_stream);
int lsps_conv_dgrad_ex(lsps_ctx*, const lsps_conv_shape*, const void* dy, const void* w_dgrad, void* dx, const void* mask,
                       const void* add, int flags, float slope, const lsps_conv_ext* ext, lsps_stream);
/* dw[tap][cout][cin] += conv_backward_weight(x, dy)   (fp32, accumulating; split-K over pixels with red.add) */
int lsps_conv_wgrad(lsps_ctx*, const lsps_conv_shape*, const void* x, const void* dy, float* dw, lsps_stream);
/* grouped 3x3 stride-1 conv: dw[tap][cout][cin/groups] += ... (only the block-diagonal of the dense gradient) */
int lsps_conv_wgrad_grouped(lsps_ctx*, const lsps_conv_shape*, const void* x, const void* dy, float* dw, int groups,
                            lsps_stream);
/* same with split-bf16 operands: x [n,h,w,2cin], dy [n,ho,wo,2cout] as (hi | lo) halves; dy_hi*x_hi + dy_hi*x_lo + dy_lo*x_hi */
int lsps_conv_wgrad_split(lsps_ctx*, const lsps_conv_shape*, const void* x, const void* dy, float* dw, lsps_stream);
/* db[c] += sum over rows of dy[rows][c]   (bias gradients; dy bf16) */
int lsps_colsum_bf16(lsps_ctx*, const void* dy, long long rows, int c, float* db, lsps_stream);

/* ---- Cin=1 7x7 stems (lsps_nets.py:104,186) : img f32 [n,h,w]; y bf16 [n,h/stride,w/stride,64]; w f32 [64][49] */
int lsps_stem_fwd(lsps_ctx*, const float* img, const float* w, const float* bias, void* y, int n, int h, int wd,
                  int stride, float slope, lsps_stream);
/* dy must already carry the LeakyReLU mask.  dw[64][49] += ..., db[64] += ... */
int lsps_stem_wgrad(lsps_ctx*, const float* img, const void* dy, float* dw, float* db, int n, int h, int wd,
                    int stride, lsps_stream);
/* dimg (+)= conv_backward_data(dy, w); accumulate != 0 adds into dimg */
int lsps_stem_dgrad(lsps_ctx*, const void* dy, const float* w, float* dimg, int n, int h, int wd, int stride,
                    int accumulate, lsps_stream);

/* split-bf16 ("bf16x3") variants: y / dy are [n,ho,wo,128] = (bf16 hi | bf16 lo) channel halves, the weights are split
   inside the kernels.  Tensor-core kernels only: output width 64 or 128. */
int lsps_stem_fwd_split(lsps_ctx*, const float* img, const float* w, const float* bias, void* y, int n, int h, int wd,
                        int stride, float slope, lsps_stream);
int lsps_stem_wgrad_split(lsps_ctx*, const float* img, const void* dy, float* dw, float* db, int n, int h, int wd,
                          int stride, lsps_stream);
int lsps_stem_dgrad_split(lsps_ctx*, const void* dy, const float* w, float* dimg, int n, int h, int wd, int stride,
                          int accumulate, lsps_stream);

/* ---- decoder head ConvTranspose2d(64,1,1)+Tanh (lsps_nets.py:226-229): x bf16 [npix,64] -> out f32 [npix] */
int lsps_head_fwd(lsps_ctx*, const void* x, const float* w, const float* bias, float* out, long long npix, lsps_stream);
/* same with the reconstruction loss fused in (nn.L1Loss vs the input image, lsps_trainer.py:118-121): for the pixels
   [t0, t0+tn) of the batch, acc += sum|out - target[p - t0]| and dout[p - t0] = scale*sign(out - target) (dout may be NULL) */
int lsps_head_fwd_l1(lsps_ctx*, const void* x, const float* w, const float* bias, float* out, long long npix,
                     const float* target, long long t0, long long tn, float scale, float* dout, float* acc, lsps_stream);
/* dpre = dout*(1-out^2); dx = dpre*w*lrelu'(x) (bf16); dw[64] += sum dpre*x ; db += sum dpre */
int lsps_head_bwd(lsps_ctx*, const void* x, const float* w, const float* out, const float* dout, void* dx, float* dw,
                  float* db, long long npix, float slope, lsps_stream);

/* ---- InstanceNorm2d(affine=False) fused with what follows it in LeakyINSResBlock (common_net.py:160-181) */
/* mode 0: y = lrelu(IN(h)) ; mode 1: y = res + IN(h).  h,y,res bf16 [n,hw,c]; stats f32 [n,c,2] = (mean, rstd) out */
int lsps_instnorm_fwd(lsps_ctx*, const void* h, const void* res, void* y, float* stats, int n, int hw, int c, int mode,
                      float eps, float slope, lsps_stream);
/* dh = IN_backward(g) with g = dy (mode 1) or dy*lrelu'(IN(h)) (mode 0); db (may be NULL): db[c] += sum over (n, hw) of
   dh -- the bias gradient of the conv that produced h, fused here instead of a separate colsum pass */
int lsps_instnorm_bwd(lsps_ctx*, const void* dy, const void* h, const float* stats, void* dh, int n, int hw, int c,
                      int mode, float slope, float* db, lsps_stream);

/* same, for a batch whose images [n_split, n) went through a second conv: their bias gradient goes to db2 */
int lsps_instnorm_bwd_grouped(lsps_ctx*, const void* dy, const void* h, const float* stats, void* dh, int n, int hw, int c,
                              int mode, float slope, float* db, float* db2, int n_split, lsps_stream);

/* ---- norms whose statistics were taken in the producing conv's epilogue (LSPS_EP_STATS): ONE streaming pass each.
   InstanceNorm2d(affine=False) of LeakyINSResBlock (common_net.py:168,171): per_image = 1, sums/stats/bsums are
   [n][2][c] rows; BatchNorm2d(affine=False) of the BN wrappers (common_net.py:187,190,274,...): per_image = 0, ONE
   [2][c] row (lsps_norm_reduce_images adds the per-image rows up), count = n*hw.
   forward : mode 0 y = lrelu(xhat), 1 y = res + xhat, 2 y = xhat ; xhat = (h - mean)*rstd from sums = (sum, sum of squares);
             stats_out (may be NULL) receives the (mean, rstd) rows the backward pass reads */
int lsps_norm_apply_fwd(lsps_ctx*, const void* h, const void* res, void* y, const float* sums, float* stats_out, int n,
                        int hw, int c, int mode, int per_image, float eps, float slope, const float* gamma,
                        const float* beta, lsps_stream);
/* gamma / beta (f32 [c], either may be NULL = 1 / 0) in the three norm calls: the affine part of BatchNorm2d(affine=True)
   (LeakyReLUBNConv2d, common_net.py:270-281) or the Bias2d that follows BatchNorm2d(affine=False) in the BNNS wrappers
   (common_net.py:294-322): the value handed to the activation is gamma*xhat + beta.  With them the two rows of bsums are
   exactly d beta = sum g and d gamma = sum g*xhat. */
/* bsums (zeroed by the call) += (sum g, sum g*xhat), g = dy (mode 1) or dy*lrelu'(xhat) (mode 0) -- only needed when the
   gradient's producer could not take the sums itself (LSPS_EP_INBWD) */
int lsps_norm_bwd_stats(lsps_ctx*, const void* dy, const void* h, const float* stats, float* bsums, int n, int hw, int c,
                        int mode, int per_image, float slope, const float* gamma, const float* beta, lsps_stream);
/* dh = rstd*(g - mean(g) - xhat*mean(g*xhat)); gmode 0: g is the raw gradient w.r.t. lrelu(xhat) (masked here),
   gmode 1: g is used as is (the gradient w.r.t. res + xhat), gmode 2: g was pre-masked by LSPS_EP_INBWD and `h` is the
   ACTIVATION a = lrelu(xhat) (xhat recovered from it; only the rstd row of stats is read).  The bias of the conv that
   produced h gets NO gradient from here: it is exactly zero (the norm subtracts the mean). */
int lsps_norm_bwd_apply(lsps_ctx*, const void* g, const void* h, const float* stats, const float* bsums, void* dh, int n,
                        int hw, int c, int gmode, int per_image, float slope, const float* gamma, const float* beta,
                        lsps_stream);
/* BatchNorm2d running statistics: momentum update from a batch row of sums (unbiased variance, like torch), and a sums
   row that makes lsps_norm_apply_fwd normalise with the running statistics (eval mode) */
int lsps_bn_running_update(lsps_ctx*, const float* sums, float* running_mean, float* running_var, int c, float count,
                           float momentum, lsps_stream);
int lsps_bn_running_to_sums(lsps_ctx*, const float* running_mean, const float* running_var, float* sums, int c,
                            float count, lsps_stream);
/* out[2][c] = sum over images of sums[n][2][c]  (BatchNorm batch statistics; the row a data-parallel run all-reduces) */
int lsps_norm_reduce_images(lsps_ctx*, const float* sums, float* out, int n, int c, lsps_stream);

/* ---- GaussianNoiseLayer + KL term (common_net.py:36-40; lsps_trainer.py:55-58): z = x + noise, acc[0] += sum z^2 */
int lsps_noise_kl_fwd(lsps_ctx*, const void* x, const float* noise, void* z, float* acc, long long n, lsps_stream);
/* same with the noise drawn inside the kernel (Philox4x32-10, counter = (seed, 16-byte chunk index, offset)): the
   device-RNG mode -- statistically, not bitwise, the reference's layer; no noise tensor touches HBM */
int lsps_noise_kl_philox(lsps_ctx*, const void* x, void* z, float* acc, long long n, unsigned long long seed,
                         unsigned long long offset, lsps_stream);
/* stream-ordered fill / device-to-device copy (cudaMemsetAsync / cudaMemcpyAsync; capturable) */
int lsps_memset(lsps_ctx*, void* dst, int value, long long bytes, lsps_stream);
int lsps_memcpy(lsps_ctx*, void* dst, const void* src, long long bytes, lsps_stream);
/* out = a + alpha * b  (bf16 tensors; a may be NULL) */
int lsps_axpy_bf16(lsps_ctx*, const void* a, const void* b, float alpha, void* out, long long n, lsps_stream);

/* Mapping latent-matching term (lsps_trainer.py:52-53,97 `_compute_l2_loss(shared, z_pose2depth)`), bf16 tensors:
   acc += sum (a-b)^2 ; g = scale*(a-b) */
int lsps_l2_bf16(lsps_ctx*, const void* a, const void* b, void* g, float scale, float* acc, long long n, lsps_stream);

/* ---- losses.  All `acc` arguments are single float accumulators (sum, not mean) in device memory. */
/* L1 (nn.L1Loss, lsps_trainer.py:42,118-121): acc += sum|x-t| ; dx (+)= scale*sign(x-t) */
int lsps_l1_f32(lsps_ctx*, const float* x, const float* t, float* dx, float scale, int accumulate, float* acc,
                long long n, lsps_stream);
/* feature matching L1 on bf16 trunk features (lsps_trainer.py:171-177,241-243):
   acc += sum|a-b| ; da += scale*sign(a-b) ; db -= scale*sign(a-b)  (da/db f32, may be NULL) */
int lsps_l1_feat(lsps_ctx*, const void* a, const void* b, float* da, float* db, float scale, float* acc, long long n,
                 lsps_stream);
/* D head Conv2d(2048,1,1) (lsps_nets.py:124,157): logits[r] = f[r,:] . w + b ; f bf16 [rows,c] */
int lsps_dhead_fwd(lsps_ctx*, const void* f, const float* w, const float* bias, float* logits, long long rows, int c,
                   lsps_stream);
/* sigmoid + binary_cross_entropy vs constant target (lsps_trainer.py:107-112,179-192), torch semantics (log clamp -100):
   acc[0] += sum bce ; acc[1] += #(sigmoid >= .5 if target==1 else <= .5) ; dlogits = scale*(p - target) */
int lsps_bce_logits(lsps_ctx*, const float* logits, float target, float scale, float* dlogits, float* acc,
                    long long rows, lsps_stream);
/* df[r,:] += dlogits[r]*w (f32) ; dw[c] += sum_r dlogits[r]*f[r,c] ; db += sum dlogits  (dw/db may be NULL) */
int lsps_dhead_bwd(lsps_ctx*, const void* f, const float* w, const float* dlogits, float* df, float* dw, float* db,
                   long long rows, int c, lsps_stream);
/* D head + sigmoid + BCE + the backward of all three, one launch (the GAN loss fused into the head it follows).
   f bf16 [rows,c] (split != 0: [rows][hi c | lo c]); rows are grouped by image group: group g = row / rows_per_group has
   the constant target targets[g] (HOST array of ngroups <= 8 floats; < 0: the group takes no part, its df rows are zero)
   and adds (sum bce, #correct) to acc[slots[g]], acc[slots[g]+1] (slots: HOST int array).  dlogit = scale*(sigmoid - t);
   df[r,:] = dlogit*w is WRITTEN (no zero fill needed); dw[c] += sum_r dlogit*f[r,c]; db += sum dlogit.
   logits / df / dw / db may be NULL. */
int lsps_dhead_bce(lsps_ctx*, const void* f, const float* w, const float* bias, long long rows, int c, int split,
                   long long rows_per_group, int ngroups, const float* targets, const int* slots, float scale,
                   float* logits, float* df, float* dw, float* db, float* acc, lsps_stream);
/* trunk-feature gradient f32 -> bf16 with the LeakyReLU mask of the features: out = df * lrelu'(f) */
int lsps_mask_to_bf16(lsps_ctx*, const float* df, const void* f, void* out, float slope, long long n, lsps_stream);
/* the same four on split-bf16 feature tensors f [rows][hi c | lo c] (df / logits stay plain f32 [rows][c]) */
int lsps_l1_feat_split(lsps_ctx*, const void* a, const void* b, float* da, float* db, float scale, float* acc,
                       long long n, int c, lsps_stream);
int lsps_dhead_fwd_split(lsps_ctx*, const void* f, const float* w, const float* bias, float* logits, long long rows,
                         int c, lsps_stream);
int lsps_dhead_bwd_split(lsps_ctx*, const void* f, const float* w, const float* dlogits, float* df, float* dw,
                         float* db, long long rows, int c, lsps_stream);
int lsps_mask_to_bf16_split(lsps_ctx*, const float* df, const void* f, void* out, float slope, long long n, int c,
                            lsps_stream);
/* db[c] += column sums of a split tensor dy [rows][hi c | lo c] */
int lsps_colsum_bf16_split(lsps_ctx*, const void* dy, long long rows, int c, float* db, lsps_stream);

/* ---- small dense layers (Post head = FC 8192->20, lsps_nets.py:123,135-145; poseVAE MLP, lsps_nets.py:34-83) */
/* y[m,n] = act(x[m,k] . w[n,k]^T + b[n]) ; act 0 none, 1 lrelu, 2 softplus.  x_bf16: 0 = f32 rows, 1 = bf16 rows,
   c > 1 = split-bf16 rows laid out [pixels][hi c | lo c] (k a multiple of c; the row holds k/c pixels) */
int lsps_linear_fwd(lsps_ctx*, const void* x, int x_bf16, const float* w, const float* b, float* y, int m, int n, int k,
                    int act, float slope, lsps_stream);
/* dy is the gradient w.r.t. the pre-activation.  dx[m,k] (+)= dy.w ; dw[n,k] += dy^T.x ; db[n] += sum dy. NULL skips. */
int lsps_linear_bwd(lsps_ctx*, const void* x, int x_bf16, const float* w, const float* dy, float* dx, int dx_accumulate,
                    float* dw, float* db, int m, int n, int k, lsps_stream);
/* dy_pre = dy * act'(y) in place on a [n] vector (act as above, y = post-activation output) */
int lsps_act_bwd(lsps_ctx*, float* dy, const float* y, int act, float slope, long long n, lsps_stream);
/* acc += sum (p-e)^2 ; dp = scale*(p-e) */
int lsps_mse(lsps_ctx*, const float* p, const float* e, float* dp, float scale, float* acc, long long n, lsps_stream);
/* poseVAE reparameterisation + KL (lsps_nets.py:72-78; lsps_trainer.py:59-60):
   z = mu + sd*noise ; acc += sum(mu^2 + sd^2 - log sd^2) */
int lsps_vae_reparam(lsps_ctx*, const float* mu, const float* sd, const float* noise, float* z, float* acc, long long n,
                     lsps_stream);
/* dmu = dz + kl_scale*2*mu ; dsd = dz*noise + kl_scale*(2*sd - 2/sd) */
int lsps_vae_reparam_bwd(lsps_ctx*, const float* mu, const float* sd, const float* noise, const float* dz, float* dmu,
                         float* dsd, float kl_scale, long long n, lsps_stream);

/* One launch for the whole pose-VAE training step before the optimiser (vae_update, lsps_trainer.py:62-74 over
   poseVAE.forward lsps_nets.py:68-83): forward, L1 + KL losses, backward.  weights / grads: 10 device pointers in the order
   en_fc1.{weight,bias}, en_mu.{..}, en_sigma.{..}, de_fc1.model.0.{..}, de_fc2.{..} (nn.Linear layouts [out][in]);
   the gradients are ADDED to grads (zero them first); dec [rows][d] = reconstruction; acc[0] += sum(mu^2 + sd^2 - log sd^2),
   acc[1] += sum|dec - y|; ll_scale / kl_scale = d(loss)/d(L1 sum) and d(loss)/d(KL sum). */
int lsps_vae_step(lsps_ctx*, const float* y, const float* noise, const float* const* weights, float* const* grads,
                  float* dec, float* acc, int rows, int d, int h, int z, float ll_scale, float kl_scale, float slope,
                  lsps_stream);

/* ---- optimiser (torch.optim.Adam with L2 weight decay, lsps_trainer.py:26-34) on a flat fp32 segment.
   g += wd*p ; m,v update ; p -= lr * mhat/(sqrt(vhat)+eps) ; optionally refresh the bf16 copy w16 (may be NULL).
   hyper (may be NULL): device pointer to {lr/(1-beta1^step), 1/sqrt(1-beta2^step)}; when given it overrides the values
   derived from lr/step so that a captured CUDA graph can be replayed with advancing step counts. */
int lsps_adam(lsps_ctx*, float* p, const float* g, float* m, float* v, void* w16, long long n, float lr, float beta1,
              float beta2, float eps, float wd, int step, float grad_scale, const float* hyper, lsps_stream);
/* wt[tap][cin][cout] (bf16) = transpose of w[tap][cout][cin] (f32 master) : the dgrad operand */
int lsps_pack_dgrad(lsps_ctx*, const float* w, void* wt, int taps, int cout, int cin, lsps_stream);
/* every conv weight of a flat store in ONE launch.  desc: device array of (count + 1) x 6 int64 = {w_off, wt_off, taps,
   cout, cin, tile0} -- element offsets from w_base / wt_base, tile0 = index of the entry's first 32x32 tile (the extra
   last row carries tile0 = total_tiles).  wt_lo_base (may be NULL): bf16 remainders for the split-bf16 operands. */
int lsps_pack_dgrad_multi(lsps_ctx*, const float* w_base, void* wt_base, void* wt_lo_base, const long long* desc,
                          int count, int total_tiles, lsps_stream);
/* lsps_adam with a second bf16 copy: w16_lo = bf16(p - float(w16)) (split-bf16 forward operands; may be NULL) */
int lsps_adam_ex(lsps_ctx*, float* p, const float* g, float* m, float* v, void* w16, void* w16_lo, long long n, float lr,
                 float beta1, float beta2, float eps, float wd, int step, float grad_scale, const float* hyper,
                 lsps_stream);
/* hi = bf16(x), lo = bf16(x - hi) */
int lsps_f32_split_bf16(lsps_ctx*, const float* x, void* hi, void* lo, long long n, lsps_stream);
int lsps_f32_to_bf16(lsps_ctx*, const float* x, void* y, long long n, lsps_stream);
/* ---- evaluation sweep on device (src/depth_train.py:229-237; src/utils/handpose_evaluation.py:92-97,197-203):
   per frame i: e_j = || (gt[i,j,:] - pred[i,j,:]) * (sx,sy,sz) ||  over joints j in joint_idx (NULL = first nj joints);
   err_mean[i] = mean_j e_j ; err_max[i] = max_j e_j.  pred/gt f32 [n, j3] normalised joints, (sx,sy,sz) = cube/2 in mm. */
int lsps_joint_errors(lsps_ctx*, const float* pred, const float* gt, const int* joint_idx, int nj, int j3, float sx,
                      float sy, float sz, float* err_mean, float* err_max, int n, lsps_stream);
int lsps_bf16_to_f32(lsps_ctx*, const void* x, float* y, long long n, lsps_stream);

/* ---- input pipeline (data/dataset_hand2.py:34-119 augmentCrop, utils/handdetector.py:682-808, cv2
   nearest-neighbour warps): img/out = n normalised 128x128 fp32 crops (distinct buffers), params = n device-resident
   `lsps_aug_sample` records (lsps_b200/csrc/augment_core.h; sizeof via lsps_aug_sample_bytes) computed on the host from
   the reference's random draws, premax = n floats of scratch.  Per pixel: de-normalise, warp (perspective: com / scale
   moves; affine: in-plane rotation), z-threshold, clamp to the new cube, normalise. */
int lsps_augment_crops(lsps_ctx*, const float* img, const void* params, float* premax, float* out, int n, lsps_stream);
int lsps_aug_sample_bytes(void);

#ifdef __cplusplus
}
#endif
#endif
