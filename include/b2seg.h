/* b2seg — C ABI of the B200-native UNet-family hot path.
 *
 * This is the drop-in boundary for the forward/backward/Adam path that the reference executes inside
 * TensorFlow when `Model.fit` / `Model.predict` run the graphs built by
 *   TensorFlow/2DCNN/models/unet_variants.py:977-1115 (class unet_model_builder) and
 *   TensorFlow/1DCNN/Models/unet_variants.py:222-319 (class UNet), 1DCNN/Models/BCDUNet.py:79-174.
 * The reference has no FFI of its own (it is pure Python on tf.keras); every entry point below replaces the
 * tf.keras layer call(s) named in its comment.  Plain pointers and sizes only: all `uint64_t` "ptr" fields are
 * CUDA device addresses owned by the caller; no torch / TF types cross this boundary.
 *
 * Conventions: every function returns 0 on success or a negative error code; `b2seg_last_error()` returns a
 * thread-local message.  All launches are asynchronous on the `stream` argument (a CUstream / cudaStream_t cast
 * to void*).  There is no CPU fallback: every entry point fails with B2SEG_ERR_DEVICE on a non-sm_100 device.
 *
 * Tensors are channels-last (NHWC; 1D tensors are NHWC with H == 1), bf16 unless stated, addressed through
 * `b2seg_view` (a strided window into a wider NHWC buffer — this is how channel concatenation
 * (2DCNN/models/unet_variants.py:27-32 Concat_Block) is realised without a copy).
 */
#ifndef B2SEG_H_
#define B2SEG_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2SEG_OK 0
#define B2SEG_ERR_ARG -1
#define B2SEG_ERR_DEVICE -2
#define B2SEG_ERR_CUDA -3

#define B2SEG_MAX_SRC 4
#define B2SEG_MAX_TAPS 16
#define B2SEG_MAX_GROUPS 4
#define B2SEG_MAX_GRADSRC 6

/* activation codes (Keras Activation strings: 2DCNN/models/unet_variants.py:7-24,67-82) */
enum { B2SEG_ACT_NONE = 0, B2SEG_ACT_RELU = 1, B2SEG_ACT_LEAKY = 2 /* slope 0.3 */, B2SEG_ACT_SIGMOID = 3, B2SEG_ACT_SOFTMAX = 4,
       B2SEG_ACT_TANH = 5 /* Self-ONN decoders (unet_variants.py:656,663); streaming kernels only, not the conv epilogue */ };

typedef struct b2seg_view {
  uint64_t ptr;          /* address of element (n=0,h=0,w=0,c=0); 16-byte aligned */
  int32_t N, H, W, C;    /* logical extents; channel stride is 1 element */
  int64_t sn, sh, sw;    /* element strides; multiples of 8 for bf16 views */
} b2seg_view;

typedef struct b2seg_tap {
  int32_t src;           /* index into src[] (conv) / pair index (wgrad) */
  int32_t dh, dw;        /* spatial offset added to the output-grid coordinate when reading src */
  int32_t widx;          /* index on the tap axis of the weight tensor */
} b2seg_tap;

/* Implicit-GEMM convolution on tcgen05 tensor cores.  Replaces tf.keras.layers.Conv2D / Conv1D
 * (unet_variants.py:9 / 1DCNN unet_variants.py:55), Conv2DTranspose / Conv1DTranspose forward
 * (:19 / :104; one group per output parity class) and, with b_mn_major = 1, their input-gradient
 * (Conv2DBackpropInput) kernels.
 *   out[g](n,h,w, :) = act( bias + sum_{t in group g} sum_c src[tap.src](n, h+dh, w+dw, c) * Wt )
 * Weights are bf16 [w_cout][w_taps][w_cin] (cin fastest).  b_mn_major = 0: GEMM-N = w_cout, K = w_cin per tap.
 * b_mn_major = 1: GEMM-N = w_cin, K = w_cout per tap (the same buffer read transposed).
 * Reads outside a source view return 0 (SAME padding).  stats (optional): fp32
 * [n_groups * m_tiles][2][out.C] per-tile column sums and sums of squares of the stored bf16 values —
 * the batch statistics of BatchNormalization (unet_variants.py:11).  With mul_view the statistics are taken after the
 * derivative mask (column sums = bias gradient of the layer whose activation derivative was applied, see b2seg_rowsum). */
typedef struct b2seg_conv_desc {
  int32_t n_src;
  b2seg_view src[B2SEG_MAX_SRC];
  uint64_t weights;
  int32_t w_cout, w_taps, w_cin;
  int32_t b_mn_major;
  int32_t n_groups, taps_per_group;
  b2seg_tap taps[B2SEG_MAX_TAPS];
  b2seg_view out[B2SEG_MAX_GROUPS]; /* per group; (N,H,W) is the GEMM-M grid, C the GEMM-N extent */
  uint64_t bias;                    /* fp32 [out.C] or 0 */
  int32_t act;
  uint64_t stats;                   /* fp32 or 0 */
  uint64_t mul_src;                 /* optional bf16 view-compatible tensor: see mul_mode */
  b2seg_view mul_view;              /* dgrad fusion: out *= act'(mul_view) with mul_mode = B2SEG_ACT_*; 0 = off */
  int32_t mul_mode;
  int32_t block_n;                  /* 0 = auto; else 64 / 128 / 256 */
  int32_t stats_atomic;             /* 1: `stats` is ONE caller-zeroed [2][out.C] row every CTA adds its sums into (red.add) instead of one row
                                     * per CTA / tile: consumers read final sums without a finalize launch (b2seg_gate_fwd) */
} b2seg_conv_desc;

/* Weight gradient (Conv2DBackpropFilter) on tcgen05: dW[co][widx][ci] (+)= sum_{n,h,w} dy[tap.src](n,h+dyh,w+dyw,co) * x[tap.src](n,h+dh,w+dw,ci)
 * fp32 output laid out like the weights. */
typedef struct b2seg_wgrad_tap {
  int32_t pair;          /* index into dy[]/x[] */
  int32_t dyh, dyw;      /* offset applied to dy coordinates */
  int32_t dh, dw;        /* offset applied to x coordinates */
  int32_t widx;
} b2seg_wgrad_tap;

typedef struct b2seg_wgrad_desc {
  int32_t n_pair;
  b2seg_view dy[B2SEG_MAX_SRC]; /* C = w_cout */
  b2seg_view x[B2SEG_MAX_SRC];  /* C = w_cin  */
  int32_t gN, gH, gW;           /* pixel iteration grid (GEMM-K) */
  int32_t n_taps;
  b2seg_wgrad_tap taps[B2SEG_MAX_TAPS];
  uint64_t dw;                  /* fp32 [w_cout][w_taps][w_cin] */
  int32_t w_cout, w_taps, w_cin;
  int32_t ksplit;               /* 0 = auto */
  int32_t accumulate;           /* 1: add into dw (must be pre-zeroed when ksplit > 1) */
} b2seg_wgrad_desc;

/* BatchNormalization, training mode (unet_variants.py:11; eps 1e-3, momentum 0.99). */
typedef struct b2seg_bn_finalize_desc {
  uint64_t partials;  /* fp32 [n_partials][2][C] from conv stats */
  int32_t n_partials, C;
  double count;       /* N*H*W */
  uint64_t gamma, beta;            /* fp32 [C] */
  uint64_t moving_mean, moving_var;/* fp32 [C], updated in place if update_moving */
  int32_t update_moving, bessel;   /* bessel = 1 for 4-D inputs (fused Keras path), 0 for 3-D (1D models) */
  float eps, momentum;
  uint64_t scale, shift, mean, rstd; /* fp32 [C] outputs: y = x*scale + shift */
  int32_t inference;               /* 1: use moving stats, ignore partials */
} b2seg_bn_finalize_desc;

/* y = act(x*scale + shift) written to up to 2 destinations, optional fused MaxPooling (p x p, stride p). */
typedef struct b2seg_bn_act_desc {
  b2seg_view x;
  uint64_t scale, shift;  /* fp32 [C] or 0 (identity) */
  int32_t act;
  int32_t n_out;
  b2seg_view out[2];
  int32_t pool_h, pool_w; /* 0/1 = none */
  b2seg_view pooled;
  int32_t c_valid;        /* channels >= c_valid (padding lanes) are written as 0; 0 = all valid */
  /* MultiResBlock / ResPath glue fused into the apply (unet_variants.py:96-99, 108-112), both optional:
   *   add:       y = act(x*scale + shift + add)  -- Add([shortcut, BatchNormalization(concat)]) + Activation('relu') in the BN apply
   *   out_stats: fp32 [2][C], caller-zeroed: per-channel sum of y and y^2 (of the stored bf16 values) are added into it -- the batch
   *              statistics of the BatchNormalization that consumes y, which then needs no statistics pass (b2seg_colstats) */
  b2seg_view add;
  uint64_t out_stats;
  int32_t out_stats_pitch;   /* floats between the row of sums and the row of sums of squares (0 = C); > C when x is one channel window
                              * of a wider tensor whose BatchNormalization owns the accumulator */
} b2seg_bn_act_desc;

typedef struct b2seg_gradsrc {
  b2seg_view g;       /* gradient tensor view (kinds 0, 1) */
  int32_t kind;       /* 0 direct (same grid), 1 max-pool routed (g is on the pooled grid), 2 pointwise head (below) */
  int32_t pool_h, pool_w;
  /* kind 2: the consumer is a Conv 1x1 head with cout <= 2 outputs (the `out` / `level{k}` layers, unet_variants.py:1106,137).
   * Its input gradient g[pix][c] = sum_o dlogits[pix][o] * head_w[c][o] is formed on the fly instead of being written and read
   * back twice, and the head's own parameter gradients dW[c][o] = sum_pix act(BN(x))[pix][c] * dlogits[pix][o], db[o] = sum_pix
   * dlogits[pix][o] are accumulated (red.add, caller-zeroed) by pass 0, which has the activation in registers anyway. */
  uint64_t dlogits;   /* fp32 [N*H*W][cout] */
  uint64_t head_w;    /* fp32 [C][cout] */
  uint64_t head_dw, head_db;
  int32_t cout;
} b2seg_gradsrc;

/* Backward of act(BN(x)): pass 0 reduces dbeta = sum(dy*m) and dgamma = sum(dy*m*xhat), pass 1 writes
 *   dx = scale * (dy*m - dbeta/count - xhat*dgamma/count). With scale == 0 (no BN) dx = dy*m.
 * accumulate = 1: the caller guarantees dgamma / dbeta hold zeros when the op starts (a training plan zeroes the whole
 * gradient arena once per step), which lets pass 0 add its block sums straight into them; 0: the op zeroes them itself. */
typedef struct b2seg_bn_bwd_desc {
  b2seg_view x;                   /* stored pre-BN conv output */
  uint64_t scale, shift, mean, rstd; /* fp32 [C] (0 => no BN) */
  int32_t act;
  int32_t n_src;
  b2seg_gradsrc src[B2SEG_MAX_GRADSRC];
  double count;
  uint64_t partials;              /* fp32 [n_blocks][2][C] scratch */
  int32_t n_blocks;               /* pixel slabs used by pass 0 */
  uint64_t dgamma, dbeta;         /* fp32 [C] outputs */
  b2seg_view dx;                  /* bf16 output */
  int32_t accumulate;
  int32_t x_relu_mask;            /* 1: x is the output of a ReLU whose only reader is this BatchNormalization: dx *= (x > 0), i.e. dx is the
                                   * gradient in front of that ReLU and no separate activation-backward pass runs (MultiResBlock :97-99) */
} b2seg_bn_bwd_desc;

/* Fused Adam (utils/tf_optimizers.py:11; Keras-2 update rule): flat fp32 master weights, fp32 grads,
 * bf16 shadow copy written in the same pass.  grad_scale multiplies the gradient (1/world_size). */
typedef struct b2seg_adam_desc {
  uint64_t w, g, m, v, w_bf16;
  int64_t n;
  float lr, beta1, beta2, eps, grad_scale;
  int64_t step; /* t >= 1 */
} b2seg_adam_desc;

/* Pointwise head: out = act(x . W + b), Cout <= 8 (the `out` / `level{k}` Conv 1x1 layers,
 * unet_variants.py:1106,137). fp32 logits/probabilities; backward produces dx (bf16), dW, db. */
typedef struct b2seg_head_desc {
  b2seg_view x;          /* bf16 activations */
  uint64_t w, b;         /* fp32 [Cin][Cout], [Cout] (Keras kernel layout for 1x1) */
  int32_t cout, act, stride; /* stride 1 or 2 (UNet3+ DS heads) */
  uint64_t y;            /* fp32 [N,H',W',cout] activated output */
  uint64_t dlogits;      /* fp32 [N,H',W',cout] (backward input) */
  b2seg_view dx;         /* bf16 (backward output) */
  uint64_t dw, db;       /* fp32 outputs */
  uint64_t logits;      /* optional fp32 [N,H',W',cout] pre-activation output of the forward */
  /* forward, optional: x is the RAW output of the last Conv_Block and the head consumes bn_act(x * bn_scale + bn_shift) computed on the
   * fly (fp32 [C] each; bn_act NONE / RELU / LEAKY).  With the head's backward folded into that layer's BatchNorm backward
   * (b2seg_gradsrc kind 2) the activated tensor of the last layer is never written or read (unet_variants.py:1104-1106). */
  uint64_t bn_scale, bn_shift;
  int32_t bn_act;
} b2seg_head_desc;

/* Loss value and backward seed dL/dlogits of one model output (tf.keras.losses.* as listed by 2DCNN/utils/tf_losses.py:8-46;
 * reduction SUM_OVER_BATCH_SIZE = mean over every element, or over every pixel for the losses that sum over the channel axis).
 * kind: 0 BinaryCrossentropy, 1 CategoricalCrossentropy, 2 MeanSquaredError, 3 MeanAbsoluteError, 4 MeanSquaredLogarithmicError
 * (the shipped Train_Configs.ini:42), 5 Huber(delta 1), 6 LogCosh, 7 BinaryFocalCrossentropy(gamma 2), 8 Poisson, 9 KLDivergence,
 * 10 Hinge, 11 SquaredHinge, 12 MeanAbsolutePercentageError, 13 CategoricalHinge, 14 CosineSimilarity.
 * act = activation of the head (NONE / SIGMOID / SOFTMAX).  Cross-entropies on their own activation (0 or 7 on sigmoid, 1 on
 * softmax) are evaluated like Keras 2 from the cached logits: seed (p - y) / count; on any other head from the clipped
 * probabilities; every other loss differentiates through the activation. */
typedef struct b2seg_loss_desc {
  uint64_t y_pred;   /* fp32 activated outputs */
  uint64_t y_true;   /* fp32 targets */
  int64_t n_pix; int32_t cout; int32_t kind; int32_t act;
  float weight;      /* loss_weights entry */
  uint64_t dlogits;  /* fp32 out (0 = value only) */
  uint64_t loss;     /* fp32[1] accumulates weight*loss */
  /* optional fp32[5], accumulated (caller-zeroed): {unweighted loss of this output, sum (p-y)^2, sum |p-y|,
   * #elements with (p > .5) == (y > .5), #pixels with argmax p == argmax y} -- the per-output loss and the training metrics of
   * Keras' fit() logs (Train.py:394-419 reads history.history) without a second pass over the outputs */
  uint64_t metrics;
} b2seg_loss_desc;

typedef struct b2seg_eltwise_desc {
  int32_t op;  /* 0: out = act(a + b) ; 1: out = a (copy) ; 2: out = a * leaky'(y=b) ; 3: out = act(a + b + c) ;
                * 4: out = a^p           (tf.math.pow(input, p) of the operational layers, onn_layers.py:19,41) ;
                * 5: out = p * a^(p-1) * b (its backward: a = the forward input, b = the gradient w.r.t. a^p) ; p = `act`, 2..8 */
  b2seg_view a, b, c, out;
  int32_t act; /* B2SEG_ACT_* applied by ops 0 and 3 (Add -> Activation('relu'), unet_variants.py:73-74,96-97,110-111);
                * the exponent p for ops 4 and 5 */
} b2seg_eltwise_desc;

/* UpSampling2D(size, 'bilinear') / UpSampling1D(size) (unet_variants.py:37; 1DCNN :122) with an optional activation
 * (UNet3+ applies sigmoid after the up-sampling, :363).  mode 0 = nearest (repeat), 1 = bilinear with half-pixel
 * centres and edge clamp (tf.image.resize).  Forward: y = act(resize(x)).  Backward: dx = resize^T(dy * act'(yfwd))
 * as a deterministic gather (no atomics). */
typedef struct b2seg_resize_desc {
  b2seg_view x;     /* forward: low-resolution input;   backward: dx (output) */
  b2seg_view y;     /* forward: high-resolution output; backward: dy (input)  */
  b2seg_view yfwd;  /* backward with act != NONE: the stored forward output */
  int32_t fh, fw, mode, act;
  int32_t c_valid;  /* channels >= c_valid are written as 0 (0 = all valid) */
  /* n_vseg > 0: the channel layout is gapped (odd-channel concat slots, MultiResBlock :85-100): only channels inside one of the
   * segments [vseg_off, vseg_off + vseg_cnt) are real; the padding lanes between them are written as 0 (sigmoid(0) != 0) */
  int32_t n_vseg;
  int32_t vseg_off[8], vseg_cnt[8];
} b2seg_resize_desc;

/* out = a * b[...,0]  (skip * resampler, the tf operator overload in Attention_Block, unet_variants.py:81).
 * Backward: da = dout * b[...,0];  db[...,0] = sum_c dout * a, db[...,1:] = 0. */
typedef struct b2seg_mulbc_desc {
  b2seg_view a, b, out;
  b2seg_view dout, da, db;
} b2seg_mulbc_desc;

/* per-channel sum and sum of squares of a bf16 view: batch statistics for a BatchNormalization whose input is not a
 * convolution output (MultiResBlock / ResPath, unet_variants.py:96-99,108-112).  partials: fp32 [n_blocks][2][C]. */
typedef struct b2seg_colstats_desc {
  b2seg_view x; uint64_t partials; int32_t n_blocks;
} b2seg_colstats_desc;

/* ConvLSTM2D/1D on a length-1 sequence with zero initial state (unet_variants.py:145-149): given the input-convolution
 * output z = [z_i | z_g | z_o] (three channel windows of width F; the forget gate and the recurrent kernel are dead),
 *   h = hard_sigmoid(z_o) * tanh(hard_sigmoid(z_i) * tanh(z_g)),  hard_sigmoid(x) = clip(0.2 x + 0.5, 0, 1).
 * Backward writes dz = [dz_i | dz_g | dz_o]. */
typedef struct b2seg_lstm_desc {
  b2seg_view z, h, dh, dz;
  int32_t F;
} b2seg_lstm_desc;

/* MaxPooling backward for an arbitrary window (UNet3+ pools 2..16): dx = dp routed to the first maximum of each window of y */
typedef struct b2seg_poolbwd_desc {
  b2seg_view y, dp, dx;
  int32_t ph, pw;
} b2seg_poolbwd_desc;

typedef struct b2seg_cast_desc { /* fp32 NHWC input -> bf16 view (channel-padded) */
  uint64_t src; int32_t N, H, W, C;
  b2seg_view out;
  /* kh * kw > 1: K-packed im2col, out(n,h,w,(i*kw+j)*C + c) = src(n, h+i-(kh-1)/2, w+j-(kw-1)/2, c) (zero outside): a kh x kw
   * 'same' convolution of the thin network input becomes a 1x1 convolution over out (one tensor-core tap instead of kh*kw) */
  int32_t kh, kw;
} b2seg_cast_desc;

typedef struct b2seg_colsum_desc { /* bias gradient: db[c] = sum over pixels of g (bf16 view) */
  b2seg_view g; uint64_t out; uint64_t scratch; int32_t n_blocks;
} b2seg_colsum_desc;

/* out[c] (+)= sum over rows of an fp32 row list partials[r * pitch + c], c < C.  Turns the per-CTA column sums a convolution
 * wrote into `stats` into a bias gradient: the input gradient of the decoder convolution, masked by LeakyReLU' in its epilogue
 * (mul_view), IS dL/d(pre-activation) of the Conv2DTranspose that feeds it (unet_variants.py:17-24), so its column sums are that
 * layer's bias gradient and no separate pass over the tensor is needed. */
typedef struct b2seg_rowsum_desc {
  uint64_t partials; int32_t n_rows, pitch, C;
  uint64_t out; int32_t accumulate;
} b2seg_rowsum_desc;

/* A model output that is an Activation over a tensor instead of a pointwise convolution with a fused activation: the Self-ONN
 * builders end in Oper2D(output_nums, (1,1), activation=final_activation, q) = activation(sum of q pointwise convolutions)
 * (unet_variants.py:1107-1108; deep-supervision levels :653 are the same without activation).
 * Also used for a pointwise head with more than 8 classes (the b2seg_head_* kernels stop at 8): the convolution runs on the
 * tensor-core kernels and this op applies the activation.
 * Forward: y[pix][o] = act(x[pix][o]), o < cout, fp32 (the layout b2seg_loss reads).  Backward: dx[pix][o] = dlogits[pix][o]
 * as bf16, channels >= cout zero (b2seg_loss has already applied the activation's derivative). */
typedef struct b2seg_outact_desc {
  b2seg_view x;        /* bf16 logits, C % 8 == 0, the first cout channels are real */
  int32_t cout, act;   /* act: NONE, SIGMOID or SOFTMAX */
  uint64_t y;          /* fp32 [N,H,W,cout] */
  uint64_t dlogits;    /* fp32 [N,H,W,cout] (backward input) */
  b2seg_view dx;       /* bf16 (backward output) */
} b2seg_outact_desc;

/* Deep-supervision target pyramid on the device (the reference builds it on the host for every batch:
 * 2DCNN/utils/helper_functions.py:359-380 = MaxPooling2D(2^k) of the mask, called from DataGenerator.py:113;
 * 1DCNN/1D_Segmentation.ipynb cell 31 = window MEAN over 2^k samples).  fp32 [N,H,W,C] -> fp32 [N,H/ph,W/pw,C]. */
typedef struct b2seg_tpool_desc {
  uint64_t src, dst;
  int32_t N, H, W, C, ph, pw;
  int32_t mode;        /* 0 = max, 1 = mean */
} b2seg_tpool_desc;

/* Additive attention gate (2DCNN/models/unet_variants.py:67-82 Attention_Block), everything between the two 1x1 projections and
 * the concat slot, fused (north_star: "the additive attention gate (1x1 convs + sigmoid multiply) fused into one kernel"; in
 * training mode its three BatchNormalizations are grid-wide reductions, i.e. kernel boundaries: two streaming launches forward,
 * four backward, csrc/gate.cu):
 *   a = BN_a(za), b = BN_b(zb)           za = Conv1x1 stride 2 (skip), zb = Conv1x1 (gating signal): b2seg_conv with stats_atomic
 *   z = Conv1x1 -> 1 (ReLU(a + b)),  m = sigmoid(BN_3(z))
 *   out = skip * (UpSampling2D(2, bilinear)(m) + LeakyReLU(Conv2DTranspose(1, 4x4, stride 2)(m)))
 * b2seg_gate_fwd also produces the BatchNorm coefficients of both branches (vec_a / vec_b), updates the three pairs of moving
 * statistics and leaves z for the backward pass; b2seg_gate_bwd produces dskip (through the multiply only: the stride-2
 * projection's own input gradient is a b2seg_conv dgrad), dza, dzb and every parameter gradient of the gate except the two
 * projection kernels. */
typedef struct b2seg_gate_desc {
  b2seg_view za, zb;            /* (N,h,w,C) bf16 raw projection outputs; C = 8 * 2^k */
  uint64_t sums_a, sums_b;      /* fp32 [2][C]: column sums and sums of squares added up by the projection kernels (caller-zeroed per step) */
  uint64_t gamma_a, beta_a, mm_a, mv_a, gamma_b, beta_b, mm_b, mv_b;   /* fp32 [C]; moving statistics updated in place when training */
  uint64_t vec_a, vec_b;        /* fp32 [4][C] out: scale, shift, mean, rstd of each branch */
  uint64_t w3, b3;              /* fp32 [C], [1]: the C -> 1 convolution */
  uint64_t z, m;                /* fp32 [N*h*w] out: the C -> 1 convolution's output and sigmoid(BN_3(z)) (both re-read by the backward) */
  uint64_t sums3;               /* fp32 [2]: sum z, sum z^2 (caller-zeroed per step) */
  uint64_t gamma3, beta3, mm3, mv3;   /* fp32 [1] */
  uint64_t wt, bt;              /* transposed-conv kernel element (ky,kx) at wt[(ky*4+kx)*wt_stride] (fp32), its bias */
  int32_t wt_stride;
  int32_t training, bessel;
  float eps, momentum;
  double count;                 /* N*h*w */
  b2seg_view skip, out;         /* (N,2h,2w,Cs) bf16, Cs = 8 * 2^k */
  /* ---- backward ---- */
  b2seg_view dout, dskip;       /* gradient of out (in), of skip through the multiply (out) */
  uint64_t dr, g3;              /* fp32 scratch [N*2h*2w], [N*h*w] */
  uint64_t bsums3, bsums_ab;    /* fp32 scratch [2], [3][C] (zeroed by the op) */
  b2seg_view dza, dzb;          /* bf16 out */
  uint64_t dgamma_a, dbeta_a, dgamma_b, dbeta_b, dgamma3, dbeta3;   /* fp32 out (stored) */
  uint64_t dw3, db3, dwt, dbt;  /* fp32, accumulated (caller-zeroed): [C], [1], 16 elements of stride wt_stride, [1] */
  /* Two-pass form of the backward (saves a skip-sized memset, a strided write and a three-tensor sum per gate): first call with
   * dskip.ptr == 0 (dza / dzb / parameter gradients only); the caller runs the stride-2 projection's dgrad into the DENSE
   * low-resolution tensor da_low (N,h,w,Cs); a second call with da_low set then writes only
   *   dskip = dout * r + (da_low at the even pixels, 0 elsewhere)  -- the skip tensor's whole gradient through this gate. */
  b2seg_view da_low;
} b2seg_gate_desc;

/* Inference: BatchNormalization (moving statistics) folded into the preceding convolution's kernel and bias, so that the convolution
 * epilogue produces act(BN(conv(x))) directly — no raw tensor, no BatchNorm pass (2DCNN/Test.py:149-164 model.predict).
 *   w_folded[co][:] = bf16(w[co][:] * s[co]),  bias_folded[co] = bias[co] * s[co] + beta[co] - moving_mean[co] * s[co],
 *   s = gamma / sqrt(moving_var + eps).  Replayed whenever the weights change (phase 2 of an inference plan). */
typedef struct b2seg_fold_desc {
  uint64_t w, bias;                 /* fp32 [cout_p][row], [cout_p] (bias may be 0) */
  uint64_t gamma, beta, moving_mean, moving_var;   /* fp32 [cout_p] */
  float eps;
  int32_t cout_p, row;              /* row = taps * cin_p elements per output channel */
  uint64_t w_folded, bias_folded;   /* bf16 [cout_p][row], fp32 [cout_p] */
} b2seg_fold_desc;

const char* b2seg_last_error(void);
int b2seg_version(void);   /* 103; the ctypes binding refuses a library of another version */
int b2seg_device_check(int device);
int b2seg_sizeof_desc(int op);  /* sizeof the descriptor struct of a B2SEG_OP_* code (binding self-check, no GPU needed) */

int b2seg_conv(const b2seg_conv_desc* d, void* stream);
int b2seg_conv_num_mtiles(const b2seg_conv_desc* d);  /* M tiles per group */
/* rows of [2][C] floats the kernel writes into `stats` (one per CTA when a CTA always covers the same output columns,
 * else one per (group, M tile)); pass it to b2seg_bn_finalize as n_partials */
int b2seg_conv_num_stat_rows(const b2seg_conv_desc* d);
int b2seg_wgrad(const b2seg_wgrad_desc* d, void* stream);
int b2seg_bn_finalize(const b2seg_bn_finalize_desc* d, void* stream);
int b2seg_bn_act(const b2seg_bn_act_desc* d, void* stream);
int b2seg_bn_bwd(const b2seg_bn_bwd_desc* d, void* stream);
int b2seg_adam(const b2seg_adam_desc* d, void* stream);
int b2seg_head_fwd(const b2seg_head_desc* d, void* stream);
int b2seg_head_bwd(const b2seg_head_desc* d, void* stream);
int b2seg_loss(const b2seg_loss_desc* d, void* stream);
int b2seg_eltwise(const b2seg_eltwise_desc* d, void* stream);
int b2seg_cast_input(const b2seg_cast_desc* d, void* stream);
int b2seg_colsum(const b2seg_colsum_desc* d, void* stream);
int b2seg_resize_fwd(const b2seg_resize_desc* d, void* stream);
int b2seg_resize_bwd(const b2seg_resize_desc* d, void* stream);
int b2seg_mulbc_fwd(const b2seg_mulbc_desc* d, void* stream);
int b2seg_mulbc_bwd(const b2seg_mulbc_desc* d, void* stream);
int b2seg_colstats(const b2seg_colstats_desc* d, void* stream);
int b2seg_lstm_fwd(const b2seg_lstm_desc* d, void* stream);
int b2seg_lstm_bwd(const b2seg_lstm_desc* d, void* stream);
int b2seg_pool_bwd(const b2seg_poolbwd_desc* d, void* stream);
int b2seg_rowsum(const b2seg_rowsum_desc* d, void* stream);
int b2seg_outact_fwd(const b2seg_outact_desc* d, void* stream);
int b2seg_outact_bwd(const b2seg_outact_desc* d, void* stream);
int b2seg_target_pool(const b2seg_tpool_desc* d, void* stream);
int b2seg_gate_fwd(const b2seg_gate_desc* d, void* stream);
int b2seg_gate_bwd(const b2seg_gate_desc* d, void* stream);
int b2seg_fold_bn(const b2seg_fold_desc* d, void* stream);

/* ---- plan: a recorded sequence of the ops above, replayed per step ---- */
typedef struct b2seg_plan b2seg_plan;
enum { B2SEG_OP_CONV = 1, B2SEG_OP_WGRAD, B2SEG_OP_BN_FINALIZE, B2SEG_OP_BN_ACT, B2SEG_OP_BN_BWD, B2SEG_OP_ADAM,
       B2SEG_OP_HEAD_FWD, B2SEG_OP_HEAD_BWD, B2SEG_OP_LOSS, B2SEG_OP_ELTWISE, B2SEG_OP_CAST, B2SEG_OP_COLSUM,
       B2SEG_OP_MEMSET, B2SEG_OP_RESIZE_FWD, B2SEG_OP_RESIZE_BWD, B2SEG_OP_MULBC_FWD, B2SEG_OP_MULBC_BWD, B2SEG_OP_COLSTATS,
       B2SEG_OP_LSTM_FWD, B2SEG_OP_LSTM_BWD, B2SEG_OP_POOL_BWD, B2SEG_OP_ROWSUM, B2SEG_OP_OUTACT_FWD, B2SEG_OP_OUTACT_BWD, B2SEG_OP_TARGET_POOL, B2SEG_OP_GATE_FWD, B2SEG_OP_GATE_BWD, B2SEG_OP_FOLD_BN };
typedef struct b2seg_memset_desc { uint64_t ptr; int64_t bytes; } b2seg_memset_desc;

/* Data parallel: backward-phase ops added to a plan AFTER this call size their grids for (SMs - sms), leaving room for the
 * CTAs of the gradient all-reduce that runs beside them (process-wide setting; 0 = use every SM). */
int b2seg_set_backward_sm_reserve(int sms);
/* debug aid: with B2SEG_TRACE=1 in the environment, CTA 0 of a halo-tile convolution records clock64() stamps per tile */
int b2seg_debug_read_trace(uint64_t* out, int n);
int b2seg_plan_create(b2seg_plan** out);
/* phase: 0 forward, 1 backward, 2 optimizer (training plans) / weight preparation (inference plans: B2SEG_OP_FOLD_BN). desc is copied. */
int b2seg_plan_add(b2seg_plan* p, int phase, int op, const void* desc, size_t desc_bytes);
int b2seg_plan_run(b2seg_plan* p, int phase, void* stream);
/* replay ops [first_op, first_op + n_ops) of a phase: lets the host interleave the data-parallel gradient exchange
 * (NCCL all-reduce of a finished slice of the gradient arena) with the rest of the backward pass */
int b2seg_plan_run_range(b2seg_plan* p, int phase, int first_op, int n_ops, void* stream);
int b2seg_plan_num_launches(const b2seg_plan* p, int phase);
int b2seg_plan_num_ops(const b2seg_plan* p, int phase);
/* replay one phase with a CUDA event after every op; ms_per_op[i] = device time of op i (profiling aid for bench.py) */
int b2seg_plan_run_timed(b2seg_plan* p, int phase, void* stream, float* ms_per_op, int n_ops);
int b2seg_plan_set_adam(b2seg_plan* p, float lr, int64_t step, float grad_scale);
void b2seg_plan_destroy(b2seg_plan* p);

#ifdef __cplusplus
}
#endif
#endif /* B2SEG_H_ */
