/*
 * mmbs.h - C ABI of libmmbs.so: the B200 (sm_100a) kernels behind the Cox-loss
 * survival hot path of gevaertlab/MultiModalBrainSurvival.
 *
 * The reference has no FFI of its own (it is pure PyTorch); its boundary for
 * this path is the set of Python callables listed below.  Each entry point
 * cites the reference interface whose arithmetic it replaces.  The host-side
 * mirror in multimodalbrainsurvival_b200/ keeps those Python names/signatures
 * and calls these functions through ctypes.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host;
 *  - the library never allocates, frees or retains device memory: all scratch
 *    is a caller-provided workspace sized by the matching *_workspace_bytes();
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*); no call
 *    synchronises the device;
 *  - return value: 0 = OK, <0 = error (mmbs_last_error() describes it);
 *    nothing throws across this boundary;
 *  - there is NO CPU fallback: on a machine without an sm_100 device every
 *    compute entry point returns MMBS_ERR_DEVICE.
 */
#ifndef MMBS_H_
#define MMBS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMBS_OK 0
#define MMBS_ERR_ARG (-1)
#define MMBS_ERR_WORKSPACE (-2)
#define MMBS_ERR_CUDA (-3)
#define MMBS_ERR_DEVICE (-4)
#define MMBS_ERR_UNSUPPORTED (-5)

/* ------------------------------------------------------------------ misc */
const char* mmbs_last_error(void);
int mmbs_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t mmbs_launch_count(void);
/* 0 when the current device is sm_100 and the driver entry points resolved */
int mmbs_device_check(void);

/* ------------------------------------------------------------------ Cox loss
 * Replaces cox_loss(cox_scores, times, status)
 *   /root/reference/1_HistoPathology/models.py:90-111   (CoxLoss.forward :113-118)
 *   /root/reference/5_JointFusion/models.py:119-140     (CoxLoss.forward :142-147)
 *   /root/reference/2_GeneExpression/models.py:24-45
 *   /root/reference/3_EarlyFusion/models.py:24-45
 *
 * forward:  perm = stable argsort(-times) (u32 radix sort, key-transformed);
 *           s~ = scores[perm]-max(scores); C = cumsum(exp(s~));
 *           loss = -(1/n) sum status[perm]*(s~ - log(C+1e-5)).
 *   perm_out   [n] int32  : sorted order: bits 0-30 = original index (bit-exact vs
 *                           torch.sort(stable)), bit 31 = (status[index] != 0)
 *   saved_e    [n] float  : s~ = scores[perm]-max in sorted order (saved for backward)
 *   saved_w    [n] float  : status[perm]/(C+1e-5)            (saved for backward)
 *   loss_out   [1] float
 *   flags_out  [1] int32  : bit0 = a NaN term was produced (the reference traps
 *                           into pdb at models.py:107-109; we report instead)
 * backward: grad_scores[perm[k]] = -(status_k - e_k * sum_{i>=k} w_i) * grad_loss / n,
 *           minus the gradient through max(scores) shared by all argmax positions.
 *   grad_loss  [1] float (device scalar: upstream gradient)
 */
size_t mmbs_cox_workspace_bytes(int64_t n);
int mmbs_cox_forward(const float* scores, const float* times, const float* status, int64_t n,
                     int32_t* perm_out, float* saved_e, float* saved_w, float* loss_out,
                     int32_t* flags_out, void* workspace, size_t workspace_bytes, void* stream);
int mmbs_cox_backward(const float* scores, const float* status, const int32_t* perm,
                      const float* saved_e, const float* saved_w, const float* grad_loss,
                      int64_t n, float* grad_scores, void* workspace, size_t workspace_bytes,
                      void* stream);
/* stable argsort(-times) only (the permutation/risk-set order), for parity tests and
 * the multi-GPU all-gathered risk set. */
int mmbs_risk_order(const float* times, int64_t n, int32_t* perm_out, void* workspace,
                    size_t workspace_bytes, void* stream);
/* Diagnostics of the last mmbs_cox_forward / mmbs_risk_order that used `workspace` (synchronises the device):
 * out_host[0] = pipeline state (0: bucketed pipeline; bit 0: a bucket overflowed in the partition, bit 1: a sub-bucket
 * was too large for the rank-by-comparison finish - either bit means the LSD-sort pipeline redid the work; -1: n is
 * outside the bucketed pipeline's range), out_host[1] = largest bucket, out_host[2] = buckets, out_host[3] = bucket
 * capacity.  For tests and tools/cox_fallback_scan.py. */
int mmbs_cox_debug_state(const void* workspace, size_t workspace_bytes, int64_t n, int32_t* out_host);

/* ------------------------------------------------------- per-patient aggregation
 * Replaces the aggregation tails of
 *   extract_features   /root/reference/1_HistoPathology/4_HistoPath_extractfeatures.py:75-89
 *                      /root/reference/2_GeneExpression/3_GeneExpress_extractfeatures.py:70-82
 *   get_survival_CI    /root/reference/1_HistoPathology/3_HistoPath_savescore.py:126-152 (+7 copies)
 * values [n,d] float, seg_ids [n] int32 in [0,n_seg): out[g,:] = mean of rows with
 * seg_ids==g (fp32 accumulation in ascending row order), counts[g] = number of rows;
 * last_row[g] = largest row index of segment g ("last row seen", savescore.py:137-138), -1 if empty.
 * A seg_id outside [0,n_seg) poisons the whole result (no host synchronisation): every mean is NaN,
 * every count is -1, every last_row is -1.
 */
size_t mmbs_segmented_mean_workspace_bytes(int64_t n, int64_t n_seg);
int mmbs_segmented_mean(const float* values, const int32_t* seg_ids, int64_t n, int64_t d,
                        int64_t n_seg, float* out, int32_t* counts, int32_t* last_row,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------ implicit-GEMM conv / linear
 * One tcgen05/TMEM kernel, TMA-fed, serves both
 *   nn.Conv2d + BatchNorm2d(eval) + ReLU (+ residual)  of Bottleneck.forward
 *       /root/reference/5_JointFusion/resnet.py:70-90, stem :151-155
 *   nn.Linear (+ ReLU)                                  of the MLPs
 *       /root/reference/2_GeneExpression/1_GeneExpress_train.py:247-257 etc.
 *
 * Activations are NHWC bf16; weights are [Cout][kh][kw][Cin] bf16 (K-major).
 * out[m, n] = act( scale[n] * sum_k A[m,k] W[n,k] + shift[n] (+ residual[m,n]) )
 *
 * A conv is described once by mmbs_conv_plan_create() (host struct holding the
 * TMA descriptors for fixed pointers/shapes) and launched by mmbs_conv_run().
 */
typedef struct mmbs_conv_plan mmbs_conv_plan; /* opaque, host memory */

typedef struct {
  int32_t batch;          /* B */
  int32_t in_h, in_w;     /* input spatial size  */
  int32_t c_in;           /* multiple of 64 (stem: see mmbs_stem_*) */
  int32_t c_out;          /* multiple of 32 */
  int32_t ksize;          /* 1 or 3 (4 = the space-to-depth stem, internal) */
  int32_t stride;         /* 1 or 2 */
  int32_t relu;           /* apply ReLU in the epilogue */
  int32_t out_f32;        /* 1: out is float32 (row stride c_out), else bf16 */
  int32_t flags;          /* bit0: 3x3 weights are [c_out][kw][c_in/64][kh][64] (halo variant)
                           * bit1: data-gradient convolution reading the FORWARD conv's packed weights: `weight` is
                           *       [c_in][ksize][ksize][c_out] (the forward [Cout][kh][kw][Cin] matrix), taps are flipped
                           *       by the kernel, the B operand is fed MN-major - no transposed weight copy (stride 1) */
  const void* in;         /* bf16 NHWC [B, in_h, in_w, c_in] */
  const void* weight;     /* bf16 [c_out, ksize*ksize*c_in] */
  const float* scale;     /* [c_out] or NULL (=1) */
  const float* shift;     /* [c_out] or NULL (=0) */
  const void* residual;   /* bf16 NHWC like out, or NULL */
  void* out;              /* NHWC [B, out_h, out_w, c_out] */
  float* stats;           /* optional [2][c_out] fp32, pre-zeroed: += per-channel sum / sum of squares of the
                           * (raw, bf16) output - the batch statistics of a training-mode BatchNorm2d */
} mmbs_conv_desc;

int mmbs_conv_plan_create(const mmbs_conv_desc* desc, mmbs_conv_plan** plan_out);
void mmbs_conv_plan_destroy(mmbs_conv_plan* plan);
/* Dropout fused into a LINEAR plan's epilogue (mmbs_linear_plan_create): after bias / ReLU the output is multiplied
 * by the keep-mask of probability 1 - p and 1 / (1 - p) - the nn.Dropout in front of the NEXT Linear of the reference
 * MLPs (/root/reference/2_GeneExpression/1_GeneExpress_train.py:247-257).  Mask = Philox-4x32-10 keyed by (seed, tag)
 * at (row, column / 8), the same function mmbs_dropout_cast_bf16 / mmbs_mlp_bwd_elementwise use, so the backward pass
 * regenerates it.  p = 0 switches the dropout off; the setting holds from the next mmbs_conv_run on. */
int mmbs_plan_set_dropout(mmbs_conv_plan* plan, float p, uint64_t seed, uint32_t tag);
int mmbs_conv_run(const mmbs_conv_plan* plan, void* stream);

/* Linear: y[M,N] = act(x[M,K] W[N,K]^T + bias[N]); x/W bf16 with K % 64 == 0 (pad
 * with zeros), N % 32 == 0.  Same kernel as the 1x1 conv. */
int mmbs_linear_plan_create(const void* x_bf16, const void* w_bf16, const float* bias,
                            void* y, int64_t m, int64_t n, int64_t k, int32_t relu,
                            int32_t out_f32, mmbs_conv_plan** plan_out);

/* TN GEMM: y[M,N] (fp32, row stride N) = a[K,M]^T b[K,N]; a / b bf16 ROW-major with K outermost (MN-major tcgen05
 * operands).  The weight gradient dW[co,ci] = sum_p dY[p,co] X[p,ci] of a 1x1 convolution / a linear layer straight
 * from the NHWC / [batch, features] tensors - no transposed copies.  m % 8 == 0, n % 64 == 0, any k (rows beyond k
 * are zero-filled).  Few output tiles + long K: split-K with fp32 reductions into the cleared output. */
int mmbs_linear_tn_plan_create(const void* a_km_bf16, const void* b_kn_bf16, float* y, int64_t m, int64_t n, int64_t k,
                               mmbs_conv_plan** plan_out);

/* NN GEMM: y[M,N] = act(x[M,K] w[K,N] + bias[N]); w bf16 row-major with K outermost (MN-major B operand): the data
 * gradient dh = dz W of a linear layer straight from the forward weights W[N_out, K_in]. k % 64 == 0, n % 64 == 0. */
int mmbs_linear_nn_plan_create(const void* x_bf16, const void* w_kn_bf16, const float* bias, void* y, int64_t m,
                               int64_t n, int64_t k, int32_t relu, int32_t out_f32, mmbs_conv_plan** plan_out);

/* Implicit weight-gradient GEMM of a convolution (ksize 1|3, stride 1|2, pad ksize/2):
 *   dw[co,ci,kh,kw] (fp32, OIHW = nn.Conv2d.weight.grad) = sum_{n,ho,wo} dy[n,ho,wo,co] * x[n, ho*s+kh-pad, wo*s+kw-pad, ci]
 * dy [B,oh,ow,c_out] and x [B,in_h,in_w,c_in] are the bf16 NHWC tensors themselves (MN-major operands, K chunk = 64
 * images at one output pixel, TMA zero fill = padding): no im2col buffer, no transposed copies.  Split-K. */
int mmbs_conv_wgrad_plan_create(const void* dy_bf16, const void* x_bf16, float* dw_oihw, int64_t batch, int64_t in_h,
                                int64_t in_w, int64_t c_in, int64_t c_out, int64_t ksize, int64_t stride,
                                mmbs_conv_plan** plan_out);

/* ------------------------------------------------ ResNet glue kernels (HBM-bound)
 * stem input: NCHW fp32 [B,3,224,224] -> space-to-depth, zero-padded NHWC bf16
 *   [B,116,116,16] (channel = (row parity, col parity, rgb), 12 used) so that the
 *   7x7/2 stem conv (resnet.py:97,152) becomes 4 K=64 taps of the GEMM kernel. */
int mmbs_stem_pack_input(const float* x_nchw, void* out_s2d_bf16, int64_t batch, void* stream);
/* same from raw uint8 pixels: ((x/255 - mean[c]) / std[c]) = ToTensor()+Normalize() of the reference's
 * transforms (1_HistoPathology/4_HistoPath_extractfeatures.py:133-137); mean/std are HOST arrays of 3 */
int mmbs_stem_pack_input_u8(const uint8_t* x_nchw, void* out_s2d_bf16, int64_t batch, const float* mean_host,
                            const float* std_host, void* stream);
/* stem weight: [64,3,7,7] fp32 -> bf16 [64, 4*64] in the matching (a, b, p, q, c) order */
int mmbs_stem_pack_weight(const float* w_oihw, void* out_bf16, void* stream);
/* 1- / 3- / 4-channel stems (RNone / ResNet / RNfour of /root/reference/5_JointFusion/resnet.py:94-337): fp32
 * [B,C,224,224] -> the same space-to-depth buffer (element (p*2+q)*C + c of every 16-slot pixel), and the matching
 * [64,C,7,7] weight pack. */
int mmbs_stem_pack_input_c(const float* x_nchw, void* out_s2d_bf16, int64_t batch, int channels, void* stream);
int mmbs_stem_pack_weight_c(const float* w_oihw, void* out_bf16, int channels, void* stream);
/* conv weight OIHW fp32 -> [O][kh][kw][I] bf16 */
int mmbs_pack_conv_weight(const float* w_oihw, void* out_bf16, int64_t c_out, int64_t c_in,
                          int64_t ksize, void* stream);
/* eval BatchNorm fold: scale = gamma/sqrt(var+eps), shift = beta - mean*scale (resnet.py:60..) */
int mmbs_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var,
                 float eps, int64_t c, float* scale, float* shift, void* stream);
/* MaxPool2d(3,2,1) NHWC bf16 (resnet.py:101,155) */
int mmbs_maxpool_3x3s2(const void* in_bf16, void* out_bf16, int64_t batch, int64_t h, int64_t w,
                       int64_t c, void* stream);
/* AvgPool2d(7)+flatten NHWC bf16 [B,7,7,C] -> fp32 [B,C] (resnet.py:106,162-163) */
int mmbs_avgpool_global(const void* in_bf16, float* out, int64_t batch, int64_t hw, int64_t c,
                        void* stream);
int mmbs_avgpool_global_f32(const float* in, float* out, int64_t batch, int64_t hw, int64_t c,
                            void* stream);
/* fp32 [rows, cols] -> bf16 [rows, cols_padded] (zero pad), optional transposed copy */
int mmbs_cast_pad_bf16(const float* in, void* out_bf16, int64_t rows, int64_t cols,
                       int64_t cols_padded, void* stream);

/* ------------------------------------------------ MLP training glue (HBM-bound)
 * The reference's MLPs are nn.Sequential(Dropout, Linear, ReLU, ...) stacks trained with autograd
 * (/root/reference/2_GeneExpression/1_GeneExpress_train.py:247-257, :161-164).  Forward and backward
 * GEMMs run on mmbs_linear_plan_create plans; these kernels provide the operands:
 *   dropout_cast      x (fp32|bf16) -> bf16, Bernoulli(1-p) mask * 1/(1-p) from Philox(seed, tag)
 *   mlp_bwd_elementwise  dz = g * dropmask/(1-p) * relu'(act); writes dz, dz^T and db += colsum(dz)
 *   transpose_bf16 / cast_transpose_pad_bf16   K-major operands for wgrad / dgrad
 */
int mmbs_dropout_cast_bf16(const void* in, int32_t in_is_bf16, int64_t in_stride, void* out_bf16, int64_t rows,
                           int64_t cols, int64_t cols_padded, float p, uint64_t seed, uint32_t tag, void* stream);
int mmbs_mlp_bwd_elementwise(const void* g, int32_t g_is_bf16, int64_t g_stride, const void* act_bf16,
                             int64_t act_stride, int32_t relu, float p, uint64_t seed, uint32_t tag, int64_t m,
                             int64_t n, int64_t n_padded, int64_t m_padded, void* dz_bf16, void* dzt_bf16, float* db,
                             void* stream);
int mmbs_transpose_bf16(const void* in_bf16, int64_t in_stride, int64_t rows, int64_t cols, int64_t rows_padded,
                        void* out_bf16, void* stream);
int mmbs_cast_transpose_pad_bf16(const float* in, int64_t n, int64_t k, int64_t k_padded, int64_t n_padded,
                                 void* out_bf16, void* stream);

/* ------------------------------------------------ training-mode ResNet trunk (HBM-bound glue)
 * model.train() semantics of Bottleneck.forward (/root/reference/5_JointFusion/resnet.py:70-90): every
 * BatchNorm2d normalises with the batch mean / biased variance and updates its running statistics; autograd
 * runs through layer4 + fc (/root/reference/1_HistoPathology/2_HistoPath_train.py:541-551).
 * Forward:  conv (mmbs_conv_desc.stats: raw bf16 output + per-channel sum / sum of squares)
 *           -> mmbs_bn_finalize -> mmbs_bn_apply (+ residual, ReLU)  |  mmbs_bn_relu_maxpool_3x3s2 (stem).
 * Backward: mmbs_avgpool_global_bwd -> per BatchNorm mmbs_bn_bwd_reduce + mmbs_bn_bwd_apply ->
 *           dgrad = conv plans on mmbs_pack_conv_weight_dgrad weights (stride 2: mmbs_scatter_stride2 first),
 *           wgrad = linear plans on transposed operands (mmbs_im2col_t, mmbs_transpose_bf16),
 *           mmbs_unpack_conv_wgrad -> nn.Conv2d.weight.grad layout; mmbs_add_relu_mask joins the shortcut. */
/* stats [2][c] = per-channel sum | sum of squares over `count` values.  Writes scale = gamma*invstd,
 * shift = beta - mean*scale, mean_out, invstd_out; running_mean/var (may be NULL) updated in place with
 * `momentum` and the unbiased variance, like nn.BatchNorm2d. */
int mmbs_bn_finalize(const float* stats, int64_t c, int64_t count, const float* gamma, const float* beta,
                     float eps, float momentum, float* running_mean, float* running_var, float* scale,
                     float* shift, float* mean_out, float* invstd_out, void* stream);
/* out = [relu]( x*scale + shift + R ), R = 0 | residual | residual*res_scale + res_shift; bf16 [rows, c] */
int mmbs_bn_apply(const void* x_bf16, const float* scale, const float* shift, const void* residual_bf16,
                  const float* res_scale, const float* res_shift, int32_t relu, void* out_bf16, int64_t rows,
                  int64_t c, void* stream);
/* MaxPool2d(3,2,1)(relu(x*scale + shift)), NHWC bf16 (training-mode stem tail, resnet.py:152-155) */
int mmbs_bn_relu_maxpool_3x3s2(const void* in_bf16, const float* scale, const float* shift, void* out_bf16,
                               int64_t batch, int64_t h, int64_t w, int64_t c, void* stream);
/* Fused finalize + apply: the kernel derives scale/shift from the epilogue sums itself (no finalize launch between the
 * convolution and its normalisation), publishes scale/shift/mean/invstd for the backward pass and updates the running
 * statistics (momentum, unbiased variance) exactly once.  Launched with programmatic dependent launch. */
typedef struct {
  const float* stats;     /* [2][c] sum | sum of squares (mmbs_conv_desc.stats) */
  const float* gamma;     /* [c] BatchNorm2d.weight */
  const float* beta;      /* [c] BatchNorm2d.bias */
  float* running_mean;    /* [c] updated in place, or NULL */
  float* running_var;
  float* scale_out;       /* [c] each: gamma*invstd, beta - mean*scale, batch mean, 1/sqrt(var+eps) */
  float* shift_out;
  float* mean_out;
  float* invstd_out;
  float eps, momentum;
  int64_t count;          /* values per channel = B*H*W */
  int64_t c;
} mmbs_bn_train_desc;
/* out = [relu]( bn(x) + R ), R = 0 | residual | res_bn(residual); bf16 [rows, c] */
int mmbs_bn_train_apply(const mmbs_bn_train_desc* bn, const void* x_bf16, const void* residual_bf16,
                        const mmbs_bn_train_desc* res_bn, int32_t relu, void* out_bf16, int64_t rows, void* stream);
/* MaxPool2d(3,2,1)(relu(bn(x))), NHWC bf16 [B,h,w,c] (training-mode stem tail) */
int mmbs_bn_train_relu_maxpool_3x3s2(const mmbs_bn_train_desc* bn, const void* in_bf16, void* out_bf16, int64_t batch,
                                     int64_t h, int64_t w, void* stream);
/* gradient of AvgPool2d(7)+flatten: dfeat fp32 [B,c] -> bf16 [B,hw,c] */
int mmbs_avgpool_global_bwd(const float* dfeat, void* out_bf16, int64_t batch, int64_t hw, int64_t c, void* stream);
/* sums[2][c] (pre-zeroed) += sum dz | sum dz*xhat,  dz = g * (relu_mask > 0), xhat = (raw-mean)*invstd;
 * they are also d(beta) | d(gamma) of the BatchNorm */
int mmbs_bn_bwd_reduce(const void* g_bf16, const void* relu_mask_bf16, const void* raw_bf16, const float* mean,
                       const float* invstd, float* sums, int64_t rows, int64_t c, void* stream);
/* out = scale * (dz - sums[0]/rows - xhat * sums[1]/rows): gradient wrt the conv output */
int mmbs_bn_bwd_apply(const void* g_bf16, const void* relu_mask_bf16, const void* raw_bf16, const float* mean,
                      const float* invstd, const float* scale, const float* sums, void* out_bf16, int64_t rows,
                      int64_t c, void* stream);
/* x NHWC bf16 [B,h,w,c] -> colT bf16 [ksize*ksize*c, p_padded] (pad ksize/2; zero outside and for columns
 * >= B*oh*ow): the K-major B operand of the weight-gradient GEMM; ksize 1 / stride 1 = plain transpose.
 * Row order (tap, channel) when oihw_rows = 0, (channel, tap) when 1: with the latter the GEMM result
 * [c_out, c*ksize*ksize] IS nn.Conv2d.weight.grad (OIHW). */
int mmbs_im2col_t(const void* x_bf16, void* out_bf16, int64_t batch, int64_t h, int64_t w, int64_t c, int64_t ksize,
                  int64_t stride, int64_t p_padded, int32_t oihw_rows, void* stream);
/* OIHW fp32 -> bf16 [I][k][k][O] with flipped taps: the weights of the data-gradient convolution */
int mmbs_pack_conv_weight_dgrad(const float* w_oihw, void* out_bf16, int64_t c_out, int64_t c_in, int64_t ksize,
                                void* stream);
/* wgrad GEMM result [O][kh][kw][I] fp32 -> OIHW fp32 */
int mmbs_unpack_conv_wgrad(const float* g, float* out_oihw, int64_t c_out, int64_t c_in, int64_t ksize, void* stream);
/* u[n,2y,2x,:] = g[n,y,x,:] (u [B,2h,2w,c] bf16, zero elsewhere - zeroed once by the caller) */
int mmbs_scatter_stride2(const void* g_bf16, void* u_bf16, int64_t batch, int64_t h, int64_t w, int64_t c,
                         void* stream);
/* out = a + g * (mask > 0), bf16; a may be NULL */
int mmbs_add_relu_mask(const void* a_bf16, const void* g_bf16, const void* mask_bf16, void* out_bf16, int64_t elems,
                       void* stream);

/* ------------------------------------------------ concordance index (tail of get_survival_CI)
 * Replaces lifelines.utils.concordance_index(survival_months, -score, vital_status)
 *   /root/reference/1_HistoPathology/3_HistoPath_savescore.py:147 (+ 7 copies).
 * counts_out[3] (device, u64) = {admissible pairs, correct, tied}; C = (correct + tied/2) / pairs.
 * Pair rule (restated from lifelines' _concordance_summary_statistics, PARITY UNPINNED - lifelines is not
 * vendored/pinned by the reference): subject i vs every observed death j with t_j < t_i, plus t_j == t_i when
 * i is censored; correct: pred_j < pred_i, tied: pred_j == pred_i.  Exact integer counts, O(n^2). */
int mmbs_concordance_counts(const double* event_times, const double* predicted, const uint8_t* event_observed,
                            int64_t n, unsigned long long* counts_out, void* stream);

/* ------------------------------------------------ fused multi-tensor Adam
 * Replaces the arithmetic of torch.optim.Adam.step() as the reference's scripts build it
 *   /root/reference/2_GeneExpression/1_GeneExpress_train.py:303-305 (two parameter groups),
 *   /root/reference/1_HistoPathology/2_HistoPath_train.py:558, /root/reference/5_JointFusion/1_JointFusion_train.py:413-416:
 *   g += weight_decay * p (L2 form, not AdamW);  m += (1-beta1) (g - m);  v = beta2 v + (1-beta2) g^2;
 *   p -= step_size * m / (sqrt(v) / bias_correction2_sqrt + eps),  step_size = lr / (1 - beta1^t),
 *   bias_correction2_sqrt = sqrt(1 - beta2^t)  (computed by the caller from its step counter, like torch does).
 * One pass over (p, g, m, v), all fp32 device tensors; the descriptor tables are HOST arrays (copied into the
 * kernel's parameter space: nothing to keep alive after the call returns). */
#define MMBS_ADAM_MAX_TENSORS 64   /* per launch; longer lists are split over several launches */
#define MMBS_ADAM_MAX_GROUPS 8
typedef struct {
  float step_size, beta1, beta2, eps, weight_decay, bias_correction2_sqrt;
  float one_minus_beta1, one_minus_beta2;   /* rounded from the caller's double 1 - beta, like torch's scalars
                                               (1.0f - 0.999f differs from float(1 - 0.999) by 1.3e-5 relative) */
} mmbs_adam_group;
typedef struct {
  void* p;          /* parameter, updated in place */
  const void* g;    /* gradient */
  void* m;          /* exp_avg, updated in place */
  void* v;          /* exp_avg_sq, updated in place */
  int64_t n;        /* elements */
  int32_t group;    /* index into the group table */
  int32_t reserved;
} mmbs_adam_tensor;
int mmbs_adam_step(const mmbs_adam_tensor* tensors_host, int32_t n_tensors, const mmbs_adam_group* groups_host,
                   int32_t n_groups, void* stream);

/* ------------------------------------------------ discrete-time survival NLL
 * Replaces nll_loss() / NLLSurvLoss (/root/reference/1_HistoPathology/models.py:120-232; `survival_bin` task):
 * hazards = sigmoid(h), S = cumprod(1 - hazards), loss_i = -(1-c)(log S_padded[y] + log hazards[y])
 * - (1-alpha) c log S_padded[y+1] with every log argument clamped at eps; mean (reduction_mean = 1) or sum over n.
 * h fp32 [n, n_bins], y int64 [n] (bin index in [0, n_bins)), c fp32 [n] (1 = censored).  The forward also writes
 * grad_unit[n, n_bins] = d loss_i / d h_i (saved for backward) and the per-sample losses; flags_out[0] != 0: some
 * y was out of range (those samples are NaN).  backward: grad_h = grad_unit * grad_loss (/ n for the mean). */
int mmbs_nll_surv_forward(const float* h, const int64_t* y, const float* c, int64_t n, int32_t n_bins, float alpha,
                          float eps, int32_t reduction_mean, float* loss_i, float* grad_unit, float* loss_out,
                          int32_t* flags_out, void* stream);
int mmbs_nll_surv_backward(const float* grad_unit, const float* grad_loss, int64_t n, int32_t n_bins,
                           int32_t reduction_mean, float* grad_h, void* stream);

/* ------------------------------------------------ feature-matrix writer (host code, no GPU needed)
 * Replaces np.savetxt(path, features, delimiter=",") of
 *   /root/reference/1_HistoPathology/4_HistoPath_extractfeatures.py:184-192,
 *   /root/reference/2_GeneExpression/3_GeneExpress_extractfeatures.py:143-149
 * byte for byte ('%.18e', ',' between columns, '\n' per row).  data: rows x cols float64, row-major, HOST memory.
 * threads <= 0: one per hardware thread (at most 64). */
int mmbs_write_matrix_csv(const double* data, int64_t rows, int64_t cols, const char* path, int32_t threads);

/* Large cohorts: the same three counts as a 2-D dominance problem (O(n * 2^block_shift) instead of O(n^2)).  The caller
 * sorts the deaths by exit time and by prediction and passes: perm[s] = time position of the s-th smallest prediction, inv =
 * its inverse, table[(nb + 1) x (nb + 1)] = exclusive 2-D prefix of the (s / S, perm[s] / S) histogram (S = 2^block_shift, nb
 * = ceil(n_deaths / S)), and per subject: admissible[i] = number of admissible deaths (a prefix of the time order), lo[i] /
 * hi[i] = deaths with a smaller / smaller-or-equal prediction.  Adds correct to counts_out[1] and tied to counts_out[2]
 * (the caller owns counts_out[0] = sum of admissible). */
int mmbs_concordance_dominance(const int32_t* perm, const int32_t* inv, const int64_t* table, int64_t n_deaths,
                               int block_shift, const int64_t* lo, const int64_t* hi, const int64_t* admissible, int64_t n,
                               unsigned long long* counts_out, void* stream);

/* ------------------------------------------------------- attention aggregation
 * Tail of TanhAttention.forward (/root/reference/1_HistoPathology/models.py:22-33; 5_JointFusion/models.py, same text):
 *   attn[b, p] = softmax_p( tanh(h[b, p, :]) . vector ),  h = x W^T (fp32, from a linear plan)
 *   out[b, p, :] = x[b, p, :] * attn[b, p] * bag          (optional)
 *   pooled[b, :] = sum_p x[b, p, :] * attn[b, p]          (optional; = out.mean(dim = 1))
 * x, h: [batch, bag, dim] fp32; bag <= 8192.  mmbs_tanh_inplace_f32: the F.tanh of AggregationProjectModel.extract. */
int mmbs_attention_pool(const float* x, const float* h, const float* vector, int64_t batch, int bag, int dim,
                        float* attn, float* out, float* pooled, void* stream);
int mmbs_tanh_inplace_f32(float* x, int64_t n, void* stream);

/* ------------------------------------------------------- input pipeline
 * Replaces, for the histopathology loaders, PatchBagDataset.__getitem__'s decode
 *   (/root/reference/1_HistoPathology/models.py:280-286: Image.open(f).convert('RGB'))
 * and the training transforms RandomHorizontalFlip / RandomVerticalFlip / ColorJitter
 *   (/root/reference/1_HistoPathology/2_HistoPath_train.py:474-488);
 * ToTensor + Normalize are fused into mmbs_stem_pack_input_u8.
 *
 * mmbs_png_decode / mmbs_png_decode_files: HOST code (no device needed).  8-bit non-interlaced PNG (grey, RGB, palette,
 * with or without alpha - alpha is dropped) -> uint8 [h, w, 3]; the file list is decoded by `threads` host threads
 * (0 = all cores) into out_hwc[count, h, w, 3] (pinned memory recommended).
 *
 * mmbs_augment_u8: uint8 [batch, h, w, 3] -> uint8 [batch, 3, h, w] on the device, bit-exact with torchvision 0.26 on PIL
 * images (Pillow 12.2): one mmbs_aug_params per image, lsum_ws = batch uint32 words of scratch. */
typedef struct mmbs_aug_params {
  int32_t hflip, vflip;
  int32_t order[4];   /* operations in application order: 0 brightness, 1 contrast, 2 saturation, 3 hue, -1 none */
  float factor[3];    /* brightness, contrast, saturation factors */
  int32_t hue_shift;  /* int(np.int32(hue_factor * 255).astype(np.uint8)) */
} mmbs_aug_params;
int mmbs_png_decode(const uint8_t* file_bytes, size_t nbytes, uint8_t* out_hwc, int h, int w);
int mmbs_png_decode_files(const char* const* paths, int64_t count, uint8_t* out_hwc, int h, int w, int threads);
/* transforms.Resize(img_size) on PIL images = Image.resize(BILINEAR) (/root/reference/1_HistoPathology/
 * 2_HistoPath_train.py:476,484): Pillow's antialiased triangle filter, bit-exact.  mmbs_resample_coeffs (HOST code) fills
 * the 22-bit fixed-point coefficient rows of one axis (call with null pointers for the row length `ksize`); the caller
 * uploads them; mmbs_resize_bilinear_u8 runs the horizontal then the vertical pass on uint8 [batch, h, w, 3]
 * (tmp: batch * in_h * out_w * 3 bytes). */
int mmbs_resample_coeffs(int in_size, int out_size, int32_t* bounds_host, int32_t* kk_host, int kk_capacity);
int mmbs_resize_bilinear_u8(const uint8_t* in_hwc, uint8_t* out_hwc, uint8_t* tmp, int64_t batch, int in_h, int in_w,
                            int out_h, int out_w, const int32_t* bounds_w, const int32_t* kk_w, int ksize_w,
                            const int32_t* bounds_h, const int32_t* kk_h, int ksize_h, void* stream);
int mmbs_augment_u8(const uint8_t* in_hwc, uint8_t* out_chw, int64_t batch, int h, int w, const void* params_dev,
                    uint32_t* lsum_ws, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMBS_H_ */
