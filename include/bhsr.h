/* bhsr.h — C ABI of libbhsr.so: the B200 (sm_100a) kernels behind the RRDBNet x4 feature
 * extractor and the feature-aggregation height head.
 *
 * The reference (lauraset/Super-resolution-building-height-estimation) has no FFI of its own:
 * its hot path is a chain of stock torch.nn calls.  Each entry point below replaces one such
 * call site (cited as reference file:line); the Python nn.Module mirrors in
 * super-resolution-building-height-estimation_b200/ bind these with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller
 *     (the library never allocates, frees or keeps device memory past a call);
 *   - every function enqueues on the CUDA stream passed as `stream` (a cudaStream_t cast to
 *     void*), never synchronises, and is CUDA-graph capturable;
 *   - return 0 on success, a negative BHSR_E* code otherwise; bhsr_last_error() returns a
 *     thread-local human-readable message for the last failure on the calling thread;
 *   - activations between kernels live in "planes": NHWC fp16 arrays [NB][H][W][C] holding the
 *     high half of an fp32 value (`hi`) and, optionally, the residual (value-hi)*2^11 (`lo`).
 *     hi + lo*2^-11 reproduces the fp32 value to ~22 bits; the `exact` numerics mode feeds both
 *     planes to the tensor cores (3 fp16 products), the `fast` mode feeds only `hi`.
 */
#ifndef BHSR_H_
#define BHSR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BHSR_VERSION 100

enum {
  BHSR_OK = 0,
  BHSR_EINVAL = -1,   /* bad argument / unsupported shape */
  BHSR_ECUDA = -2,    /* CUDA runtime or driver error (message has the CUDA string) */
  BHSR_ENOGPU = -3    /* no sm_100 device */
};

/* numerics mode of the tensor-core convolutions */
enum { BHSR_NUMERICS_EXACT_F16X3 = 0, BHSR_NUMERICS_FAST_F16 = 1 };

/* epilogue flags of bhsr_conv_tc (OR-ed) */
enum {
  BHSR_EPI_LRELU = 1,      /* v = v > 0 ? v : 0.2 v            (rrdbnet_arch.py:137-140, 219-220) */
  BHSR_EPI_RES1 = 2,       /* v = v * alpha1 + res1            (rrdbnet_arch.py:143, 217)         */
  BHSR_EPI_RES2 = 4,       /* v = v * alpha2 + res2  (after 1) (rrdbnet_arch.py:167)              */
  BHSR_EPI_OUT_NCHW_F32 = 8, /* write fp32 NCHW instead of fp16 planes (rrdbnet_arch.py:238)     */
  BHSR_EPI_RELU = 16,      /* v = max(v, 0) after the residual adds (HRfuse.py:146-157)          */
  BHSR_EPI_SHUFFLE2 = 32,  /* scatter through nn.PixelShuffle(2) into planes (HRfuse.py:24)      */
  BHSR_EPI_ACCUM = 64      /* with OUT_NCHW_F32: y += v (sums gradient contributions in backward)  */
};

int bhsr_version(void);
const char* bhsr_last_error(void);
/* number of SMs / compute capability of the current device; <0 on error */
int bhsr_device_sm_count(void);
int bhsr_device_cc(void);

/* ------------------------------------------------------------------------------------------
 * One 3x3 (or tap-table) convolution as an im2col-free implicit GEMM on tcgen05 tensor cores.
 * Replaces nn.Conv2d(+LeakyReLU / residual / torch.cat / F.interpolate(nearest x2)) call sites:
 *   ResidualDenseBlock.forward  SR/rrdbnet_arch.py:136-143
 *   RRDB.forward                SR/rrdbnet_arch.py:162-167
 *   RRDBNet.forward_feature     SR/rrdbnet_arch.py:225-240 (conv_body, conv_up1/2, conv_hr)
 * ------------------------------------------------------------------------------------------ */
typedef struct BhsrConvTcDesc {
  /* input planes: NHWC fp16 [nb][h][w][in_ctot]; channels [in_choff, in_choff+cin) are read */
  const void* in_hi;
  const void* in_lo;        /* used only by BHSR_NUMERICS_EXACT_F16X3 */
  int32_t nb, h, w;         /* any size; the kernel walks 64-pixel-wide strips */
  int32_t in_ctot, in_choff, cin; /* cin multiple of 32; in_ctot multiple of the chunk width (64 fast, 32 exact) */
  /* packed weights from bhsr_pack_conv_weights (same numerics mode, same tap table) */
  const void* w_packed;
  int32_t cout;             /* 32 or 64 */
  const float* bias;        /* [cout] or NULL */
  const float* scale;       /* [cout] or NULL: v = acc*scale + bias (eval-mode BatchNorm folded in) */
  int32_t cout_valid;       /* channels actually stored (0 = cout); padded rows of the MMA are dropped */
  /* tap table: output(y,x) = sum_t W_t . in(y+dy[t], x+dx[t]); zero outside the image */
  int32_t ntaps;            /* 1..9 */
  int8_t dy[9], dx[9];
  /* output: pixel (y,x) of the conv grid lands at (y*out_scale+out_oy, x*out_scale+out_ox)
   * of an [nb][oh][ow] image; out_scale is 1, or 2 for the sub-pixel phases of nearest-x2 */
  int32_t oh, ow, out_scale, out_oy, out_ox;
  void* out_hi;             /* NHWC fp16 [nb][oh][ow][out_ctot], channels [out_choff, +cout) */
  void* out_lo;             /* may be NULL */
  int32_t out_ctot, out_choff;
  float* out_f32;           /* with BHSR_EPI_OUT_NCHW_F32: [nb][out_ctot][oh][ow] fp32 */
  int32_t epilogue;         /* BHSR_EPI_* flags */
  float alpha1, alpha2;
  const void* res1_hi; const void* res1_lo; int32_t res1_ctot, res1_choff;
  const void* res2_hi; const void* res2_lo; int32_t res2_ctot, res2_choff;
  int32_t numerics;         /* BHSR_NUMERICS_* */
  int32_t mblocks;          /* 128-pixel accumulator blocks per tile: 1 or 2 (0 = auto) */
  int32_t max_ctas;         /* 0 = one CTA per SM */
  int32_t desc_mode;        /* 0 = default; bit 8 (0x100) forces the per-tap kernel for 32-output layers
                             * (default: the dx-in-N kernel); bit 9 (0x200) forces the single-CTA per-tap
                             * kernel for 64-output exact layers (default: CTA pairs); bit 10 (0x400) opts the
                             * 32-output exact layers into CTA pairs too (measured: no gain), see conv_tc.cu */
} BhsrConvTcDesc;

int bhsr_conv_tc(const BhsrConvTcDesc* desc, void* stream);
/* Debug aid (BHSR_DEBUG_TIMING=1): per-CTA cycle counters of the last bhsr_conv_tc launch's MMA
 * warp — {total, wait accumulator-free, wait activation tile, wait weight slab, tiles, 0,0,0};
 * synchronises. */
int bhsr_debug_timing(long long* host_out, int32_t n_ctas);

/* bytes of the packed weight blob for one conv */
size_t bhsr_packed_conv_weight_bytes(int32_t cout, int32_t cin, int32_t ntaps, int32_t numerics);
/* Pack an OIHW fp32 conv weight [cout][cin][3][3] (device) for bhsr_conv_tc.
 * fold_phase < 0: plain 3x3, 9 taps in (ky,kx) row-major order, dy=ky-1, dx=kx-1.
 * fold_phase = 2*a+b (a,b in {0,1}): the 2x2-tap phase (a,b) of conv3x3(nearest_x2(x)):
 *   rows  a=0: dy=-1 <- w[ky=0], dy=0 <- w[1]+w[2];  a=1: dy=0 <- w[0]+w[1], dy=+1 <- w[2]
 *   (same for columns with b); taps ordered (dy,dx) row-major.  SR/rrdbnet_arch.py:236-237. */
int bhsr_pack_conv_weights(const float* w_oihw, int32_t cout, int32_t cin, int32_t fold_phase,
                           int32_t numerics, void* w_packed, void* stream);

/* ------------------------------------------------------------------------------------------
 * Layout / precision plumbing around the tensor-core convs
 * ------------------------------------------------------------------------------------------ */
/* fp32 NCHW [nb][c][h][w] -> hi/lo planes NHWC [nb][h][w][ctot] at channel offset choff */
int bhsr_nchw_f32_to_planes(const float* x, int32_t nb, int32_t c, int32_t h, int32_t w,
                            void* out_hi, void* out_lo, int32_t ctot, int32_t choff, void* stream);
/* planes -> fp32 NCHW (hi + lo*2^-11; lo may be NULL) */
int bhsr_planes_to_nchw_f32(const void* in_hi, const void* in_lo, int32_t nb, int32_t c, int32_t h,
                            int32_t w, int32_t ctot, int32_t choff, float* y, void* stream);

/* Direct fp32 3x3 convolution on CUDA cores for the thin ends of the net (K=27 or N=3):
 * conv_first (SR/rrdbnet_arch.py:232) reads fp32 NCHW with an arbitrary batch/channel stride
 * (so x[:, :3] views need no copy) and writes planes to one or two destinations (out2_* may be
 * NULL; cout must be 64); conv_last (:222) reads planes and
 * writes fp32 NCHW.  lrelu_in applies LeakyReLU(0.2) to the input (forward(): :221). */
int bhsr_conv3x3_first(const float* x, int64_t x_stride_n, int64_t x_stride_c, int64_t x_stride_h,
                       int64_t x_stride_w, int32_t nb, int32_t cin, int32_t h, int32_t w,
                       const float* weight, const float* bias, int32_t cout, void* out_hi,
                       void* out_lo, int32_t out_ctot, int32_t out_choff, void* out2_hi,
                       void* out2_lo, int32_t out2_ctot, int32_t out2_choff, void* stream);
int bhsr_conv3x3_last(const void* in_hi, const void* in_lo, int32_t in_ctot, int32_t in_choff,
                      int32_t nb, int32_t cin, int32_t h, int32_t w, int32_t lrelu_in,
                      const float* weight, const float* bias, int32_t cout, float* y, void* stream);


/* ------------------------------------------------------------------------------------------
 * Whole-network entry: RRDBNet (SR/rrdbnet_arch.py:170-240; same arithmetic as SR/RRDBNet.py:53-78)
 * num_feat = 64, num_grow_ch = 32 (the only widths the reference instantiates).
 * ------------------------------------------------------------------------------------------ */
typedef struct BhsrRrdbNetDesc {
  int32_t num_in_ch;        /* channels seen by conv_first (after pixel_unshuffle, if any) */
  int32_t num_out_ch;       /* conv_last outputs (<= 8); only used by bhsr_rrdbnet_forward(feature=0) */
  int32_t num_block;
  int32_t numerics;         /* BHSR_NUMERICS_* */
  int32_t nb, h, w;         /* LR batch and tile size seen by conv_first */
  const float* conv_first_w; const float* conv_first_b;   /* raw fp32 OIHW / [64] */
  const float* conv_last_w;  const float* conv_last_b;    /* raw fp32 OIHW / [num_out_ch] */
  const void* packed;       /* blob from bhsr_rrdbnet_pack (same num_block, numerics) */
  const float* biases;      /* blob from bhsr_rrdbnet_pack */
  void* workspace;          /* >= bhsr_rrdbnet_workspace_bytes(nb,h,w, feature) bytes, 1024-aligned */
  size_t workspace_bytes;
  int32_t mblocks;          /* forwarded to bhsr_conv_tc (0 = auto) */
} BhsrRrdbNetDesc;

size_t bhsr_rrdbnet_packed_bytes(int32_t num_block, int32_t numerics);
size_t bhsr_rrdbnet_bias_floats(int32_t num_block);
size_t bhsr_rrdbnet_workspace_bytes(int32_t nb, int32_t h, int32_t w, int32_t feature_only);
/* params: device pointers to the fp32 tensors of the tensor-core convs in state_dict order —
 * for each block b, rdb r, conv c: weight, bias; then conv_body, conv_up1, conv_up2, conv_hr
 * (weight, bias each): 2 * (15*num_block + 4) pointers. */
int bhsr_rrdbnet_pack(const float* const* params, int32_t num_block, int32_t numerics,
                      void* packed, float* biases, void* stream);
/* x: fp32 [nb][num_in_ch][h][w] with element strides (sn, sc, sh, sw) — views such as x[:, :3]
 * need no copy.  feature != 0: y = forward_feature(x), fp32 NCHW [nb][64][4h][4w]
 * (rrdbnet_arch.py:225-240); feature == 0: y = forward(x), [nb][num_out_ch][4h][4w] (:208-223). */
int bhsr_rrdbnet_forward(const BhsrRrdbNetDesc* desc, const float* x, int64_t sn, int64_t sc,
                         int64_t sh, int64_t sw, float* y, int32_t feature, void* stream);

/* ------------------------------------------------------------------------------------------
 * Feature-aggregation head (SR/HRfuse.py:17-190, mymodels.py:259-293, aggregate_utils.py:29-59).
 * fp32 NCHW tensors addressed as (base, channels of the underlying buffer, channel offset) so a
 * producer can write its slice of a concat buffer.  N = 1..64 output channels on 256x256 maps:
 * HBM-bound CUDA-core kernels.
 * ------------------------------------------------------------------------------------------ */
typedef struct BhsrHeadConvDesc {
  const float* x; int32_t x_ctot, x_choff;
  int32_t nb, cin, h, w;       /* conv grid (stride 1, zero "same" padding) */
  int32_t x_unshuffle;         /* read x through the inverse of PixelShuffle(2) (backward of K7) */
  const float* in_scale;       /* optional fused BatchNorm(+ReLU) on the input read:            */
  const float* in_shift;       /*   x' = relu?(x * in_scale[c] + in_shift[c])  (HRfuse.py:146-148) */
  int32_t in_relu;
  const float* weight;         /* [cout][cin][k][k] fp32 */
  const float* bias;           /* [cout] or NULL */
  int32_t cout, ksize;         /* ksize 1 or 3 */
  float* y; int32_t y_ctot, y_choff;
  int32_t y_shuffle;           /* scatter through nn.PixelShuffle(2) (HRfuse.py:24): bit-exact index path */
  double* stats;               /* [2*cout]: += sum, sum of squares of the outputs (BatchNorm batch stats) */
  int32_t accumulate;          /* y += conv (sums gradient contributions) */
} BhsrHeadConvDesc;

int bhsr_head_conv(const BhsrHeadConvDesc* desc, void* stream);
/* dw[cout][cin][k][k] = sum dy * x' (and db = sum dy); desc gives the x side of the forward conv */
int bhsr_head_conv_wgrad(const BhsrHeadConvDesc* desc, const float* dy, int32_t dy_ctot,
                         int32_t dy_choff, int32_t dy_unshuffle, float* dw, float* db, void* stream);
/* training-mode BatchNorm2d from epilogue statistics: scale/shift for the fused apply, saved
 * mean/invstd, running-stat update (momentum, unbiased variance) — HRfuse.py:129-138 */
int bhsr_bn_finalize(const double* stats, int32_t c, double count, const float* gamma,
                     const float* beta, float eps, float momentum, float* running_mean,
                     float* running_var, float* scale, float* shift, float* mean, float* invstd,
                     void* stream);
int bhsr_bn_eval_affine(int32_t c, const float* gamma, const float* beta, const float* running_mean,
                        const float* running_var, float eps, float* scale, float* shift,
                        float* invstd, void* stream);
/* out = relu(a*sa+ta + (b*sb+tb | b))   BasicBlock tail, HRfuse.py:152-157 */
int bhsr_affine_add_relu(const float* a, const float* sa, const float* ta, const float* b,
                         int32_t b_ctot, int32_t b_choff, const float* sb, const float* tb,
                         int32_t nb, int32_t c, int32_t hw, float* out, int32_t o_ctot,
                         int32_t o_choff, void* stream);
/* backward of the fused BN(+add)+ReLU: per-channel sums {g', g'*a, g'*b}, g' = g_out*[out>0]
 * (out == NULL: mask recomputed as a*sa+ta > 0) */
int bhsr_bn_bwd_reduce(const float* g_out, int32_t g_ctot, int32_t g_choff, const float* out,
                       int32_t o_ctot, int32_t o_choff, const float* a, const float* sa,
                       const float* ta, const float* b, int32_t b_ctot, int32_t b_choff, int32_t nb,
                       int32_t c, int32_t hw, double* sums, void* stream);
int bhsr_bn_bwd_coeffs(const double* sums, int32_t which, int32_t c, double count,
                       const float* gamma, const float* mean, const float* invstd,
                       const float* scale_eval, int32_t training, float* k1, float* k2, float* k3,
                       float* dgamma, float* dbeta, void* stream);
int bhsr_bn_bwd_apply(const float* g_out, int32_t g_ctot, int32_t g_choff, const float* out,
                      int32_t o_ctot, int32_t o_choff, const float* a, const float* sa,
                      const float* ta, const float* k1, const float* k2, const float* k3, float* g_a,
                      const float* b, int32_t b_ctot, int32_t b_choff, const float* kb1,
                      const float* kb2, const float* kb3, float* g_b, int32_t gb_ctot,
                      int32_t gb_choff, int32_t gb_accumulate, int32_t nb, int32_t c, int32_t hw,
                      void* stream);
/* ------------------------------------------------------------------------------------------
 * Tensor-core TRAINING path of the head (head_tc.cu): the forward, data-gradient and weight-gradient convs of
 * BasicBlock / Upsampler / conv_last under autograd (SR/HRfuse.py:143-159, 185-190; train.py:246-257) on tcgen05.
 * A BhsrHeadXform describes how a conv reads an fp32 NCHW tensor: channel window, optional BatchNorm-apply + ReLU
 * fused on the read, optional inverse PixelShuffle(2), optional device-resident scalar multiplier (power-of-two
 * gradient scale, divided out again by the consumer).
 * ------------------------------------------------------------------------------------------ */
typedef struct BhsrHeadXform {
  const float* x; int32_t x_ctot, x_choff;
  int32_t c, h, w;            /* LOGICAL channels / grid seen by the conv (after the inverse shuffle) */
  int32_t unshuffle;          /* x is stored [.., c/4, 2h, 2w]: read through the inverse of PixelShuffle(2) */
  const float* in_scale;      /* optional per-channel x' = relu?(x * in_scale + in_shift) */
  const float* in_shift;
  int32_t in_relu;
  const float* premul;        /* optional device scalar */
} BhsrHeadXform;
/* fp32 NCHW -> NHWC hi/lo fp16 planes [nb][h][w][ctot], channels [choff, choff+cpad): c values then zeros */
int bhsr_head_to_planes(const BhsrHeadXform* t, int32_t nb, void* out_hi, void* out_lo, int32_t ctot,
                        int32_t choff, int32_t cpad, void* stream);
/* NHWC hi/lo planes -> fp32 NCHW: y[:, y_choff + j] (=|+=) (hi + lo' * 2^-11) / *unscale for plane channels
 * [choff, choff + c); stats (optional, double[2c]) += sum / sum of squares of the stored values */
int bhsr_head_from_planes(const void* in_hi, const void* in_lo, int32_t nb, int32_t h, int32_t w, int32_t ctot,
                          int32_t choff, int32_t c, const float* unscale, float* y, int32_t y_ctot, int32_t y_choff,
                          int32_t accumulate, double* stats, void* stream);
/* stats[0..c) += sum, stats[c..2c) += sum of squares over (n, pixels) of channels [choff, choff+c) of y */
int bhsr_channel_stats(const float* y, int32_t y_ctot, int32_t y_choff, int32_t nb, int32_t c, int32_t hw,
                       double* stats, void* stream);
size_t bhsr_head_wgrad_workspace_bytes(int32_t nb, int32_t cin, int32_t cout, int32_t ksize, int32_t h, int32_t w);
/* dw[cout][cin][k][k] = sum dy * x' on tcgen05 (deterministic: per-CTA partials reduced in fp64);
 * db_sum (optional, double[cout], zero-initialised by the caller) += sum dy.  xt: the forward conv's input, gt: the
 * output gradient (gt->premul is divided out of dw and db_sum).  workspace: 1024-byte aligned device scratch. */
int bhsr_head_wgrad_tc(const BhsrHeadXform* xt, const BhsrHeadXform* gt, int32_t nb, int32_t ksize, float* dw,
                       double* db_sum, void* workspace, size_t workspace_bytes, void* stream);

/* step x step block aggregation: out = sum(x) / (count(x >= thr | x > thr) + 1e-10)
 * (aggregate_utils.py:29-41 uses >= 0; :44-59 uses > 1.0) on [nimg][h][w] -> [nimg][h/step][w/step] */
int bhsr_aggregate(const float* x, int32_t nimg, int32_t h, int32_t w, int32_t step,
                   float threshold, int32_t strict, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * The element-wise ends of the path (post.cu).
 * bhsr_predict_postproc: predict_realesanet_feature_globe.py:172-177 — height [nb][1][h][w]: negatives -> 0,
 *   round(h * 10) -> uint16; build [nb][k][h][w]: softmax over k, round(p * 255) -> uint16 (round half to even,
 *   like numpy).  Either input may be NULL (with its output).
 * bhsr_weighted_mse: losses_pytorch/selfloss.py:81-90 — loss = mean(weight * (pred - target)^2) * exp(-log_var) +
 *   log_var, forward and backward in one pass: grad_pred[n] (may be NULL), *grad_log_var (may be NULL), *loss;
 *   scratch = one double of device memory.
 * bhsr_ce_dice: losses_pytorch/selfloss.py:145-168 (CE_DICE_adapt_weight; Dice :6-17) — logits [nb][c][h][w] fp32,
 *   labels [nb][h][w] int64 in [0, c), weight [nb][h][w] fp32:
 *   loss = (mean(weight * CE(logits, labels)) + Dice(sum_{k>=1} softmax_k, labels > 0)) * exp(-log_var) + log_var,
 *   forward and backward: grad_logits [nb][c][h][w] (may be NULL), *grad_log_var (may be NULL), *loss;
 *   2 <= c <= 16; scratch = four doubles of device memory.
 * ------------------------------------------------------------------------------------------ */
int bhsr_predict_postproc(const float* height, const float* build, int32_t nb, int32_t k, int32_t h, int32_t w,
                          uint16_t* out_height, uint16_t* out_build, void* stream);
int bhsr_weighted_mse(const float* pred, const float* target, const float* weight, int64_t n, const float* log_var,
                      float* loss, float* grad_pred, float* grad_log_var, double* scratch, void* stream);
int bhsr_ce_dice(const float* logits, const int64_t* labels, const float* weight, int32_t nb, int32_t c, int32_t h,
                 int32_t w, const float* log_var, float* loss, float* grad_logits, float* grad_log_var,
                 double* scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BHSR_H_ */
