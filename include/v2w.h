/*
 * v2w.h -- C-ABI of the B200-native Vec2Wav generator forward pass.
 *
 * The reference (p1an-lin-jung/WavThruVec_pytorch) has no FFI layer: its boundary for this
 * path is the Python class `Generator` (vec2wav/models.py:77-156) constructed as
 * `Generator(h)` (vec2wav/train.py:65) and called as `generator(wv_feat, spk_emb, noise)`
 * (vec2wav/train.py:167,262).  This header is the boundary a binding would sit on: the
 * Python drop-in (`wavthruvec_pytorch_b200.models.Generator`) binds it with ctypes; see
 * INTEGRATION.md for the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; opaque handles; every function returns V2W_OK (0) or a negative
 *     error code and records a message retrievable with v2w_last_error() (thread-local).
 *   - no exceptions cross the boundary; no torch types appear in any signature.
 *   - "host" pointers are ordinary process memory, "dev" pointers are CUDA device memory
 *     owned by the caller (PyTorch's allocator on the Python side).
 *   - every function that touches the GPU takes the cudaStream_t to work on (as void*).
 *   - a model/plan is used by one host thread and one stream at a time; distinct plans are
 *     independent.  There is NO CPU execution path: calling a device function without a
 *     usable sm_100 GPU fails with V2W_ERR_CUDA.
 *
 * Object model
 *   v2w_model : hyper-parameters (vec2wav/hparams.py:24-57 subset read by
 *               vec2wav/models.py:81-110) -> static table of tap-GEMM layers + the packed
 *               weight format.  Pure host object.
 *   v2w_plan  : model x (batch, frames, precision) -> ordered op list, workspace layout,
 *               TMA descriptors.  Bound to device buffers with v2w_plan_bind().
 */
#ifndef V2W_H_
#define V2W_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define V2W_ABI_VERSION 1

/* ---- status codes ------------------------------------------------------------------ */
#define V2W_OK 0
#define V2W_ERR_INVALID (-1)   /* bad argument / unsupported hyper-parameter combination */
#define V2W_ERR_CUDA (-2)      /* CUDA runtime / driver error, or no usable GPU */
#define V2W_ERR_STATE (-3)     /* call order violated (e.g. forward before bind) */
#define V2W_ERR_NOMEM (-4)

/* ---- limits ------------------------------------------------------------------------ */
#define V2W_MAX_STAGES 8
#define V2W_MAX_KERNELS 4     /* resblock kernel sizes per stage (reference: 3) */
#define V2W_MAX_DILATIONS 4
#define V2W_MAX_TAPS 16
#define V2W_MAX_TILES 16

/* ---- precision of a plan ----------------------------------------------------------- */
#define V2W_PREC_FP32 0  /* FFMA kernels, fp32 everywhere: parity bar max-abs <= 1e-4 */
#define V2W_PREC_BF16 1  /* tcgen05 bf16 x bf16 -> fp32 (TMEM), fp32 residual stream: SNR >= 45 dB */
#define V2W_PREC_FP16 2  /* same kernels with fp16 operands (3 more mantissa bits, same tensor rate;
                            activations are clamped to +-65504).  Needed by ResBlock1 to reach 45 dB. */
#define V2W_PREC_FP32_TC 3 /* fp32 values on the tensor cores: every operand is the pair hi = fp16(v), lo = fp16(v - hi)
                            (22 significant bits) and every product the three-term sum A_hi W_hi + A_lo W_hi + A_hi W_lo
                            in the fp32 accumulator -- one tap-GEMM per convolution over K' = 3 cin, no fused stages.
                            Meets the fp32 bar (max-abs <= 1e-4) at about 1/3 of the 16-bit tensor rate instead of the
                            FFMA rate.  Needs a V2W_HALF_FP16_SPLIT weight buffer. */

/* 16-bit operand format of a packed weight buffer */
#define V2W_HALF_BF16 0
#define V2W_HALF_FP16 1
#define V2W_HALF_FP16_SPLIT 2   /* fp16 hi values in every layer's array, the lo values (fp16(w - hi)) in the second
                                   half of the same section (v2w_layer_desc::w16_off) */
#define V2W_HALF_BF16_SPLIT 3   /* the same with bf16 values: what a BF16 plan of a ResBlock1 generator binds (operands and
                                   weights as bf16 pairs, v2w_op_desc::split == 1; or, $V2W_RB1_BF16=weights, the weights
                                   only as hi + lo, split == 2) */
#define V2W_PACKED_HEADER_BYTES 1024   /* {uint32 magic 'V2WP', abi, half_format} then zeros */

/* Hyper-parameters.  Replaces the attribute bag `h` of Generator.__init__
 * (vec2wav/models.py:78-114; values in vec2wav/hparams.py:26-27,30,40-44,51). */
typedef struct v2w_config {
  int32_t num_wv_feat;                 /* 768 | 1024 ; any multiple of 16 accepted */
  int32_t upsample_initial_channel;    /* must be 512: CBN widths are hard-coded 256>>i (models.py:113) */
  int32_t num_upsamples;
  int32_t upsample_rates[V2W_MAX_STAGES];
  int32_t upsample_kernel_sizes[V2W_MAX_STAGES];
  int32_t resblock_type;               /* 1 = ResBlock1 (h.resblock == '1'), 2 = ResBlock2 (anything else; models.py:84) */
  int32_t num_kernels;
  int32_t resblock_kernel_sizes[V2W_MAX_KERNELS];
  int32_t num_dilations;               /* entries per row below (reference: 3; ResBlock2 reads the first 2) */
  int32_t resblock_dilation_sizes[V2W_MAX_KERNELS][V2W_MAX_DILATIONS];
  int32_t spk_dim;                     /* 192 */
  int32_t noise_dim;                   /* 192 */
} v2w_config;

/* One weight-normalised convolution of the state_dict (models.py:83,89-92,18-33,58-61,100).
 * All pointers are HOST fp32, contiguous, in PyTorch's native layout:
 *   Conv1d          v:[Cout,Cin,K]  g:[Cout] (weight_g[Cout,1,1])   bias:[Cout]
 *   ConvTranspose1d v:[Cin,Cout,K]  g:[Cin]  (weight_g[Cin,1,1])    bias:[Cout]
 * g == NULL means weight norm was removed (Generator.remove_weight_norm, models.py:149-156):
 * v is then the plain `weight`. */
typedef struct v2w_conv_params {
  const float* v;
  const float* g;
  const float* bias;
} v2w_conv_params;

/* Conditioning parameters of one upsample stage (models.py:109-113,131-134; modules.py:14-18). */
typedef struct v2w_cond_params {
  const float* fc_weight;     /* fcs.i.weight [128, spk_dim+noise_dim] */
  const float* fc_bias;       /* fcs.i.bias   [128] */
  const float* sn_weight_orig;/* cbns.i.layer.weight_orig [2C,128] */
  const float* sn_u;          /* cbns.i.layer.weight_u [2C] */
  const float* sn_v;          /* cbns.i.layer.weight_v [128] */
  const float* sn_bias;       /* cbns.i.layer.bias [2C] */
  const float* bn_mean;       /* cbns.i.batch_nrom.running_mean [C] */
  const float* bn_var;        /* cbns.i.batch_nrom.running_var  [C] */
} v2w_cond_params;

/* The whole state_dict, by reference.  `convs` is ordered: conv_pre, ups[0..U), then the
 * resblock convolutions in EXECUTION order (for n in resblocks: ResBlock2 convs[0],convs[1];
 * ResBlock1 convs1[0],convs2[0],convs1[1],convs2[1],convs1[2],convs2[2]), then conv_post. */
typedef struct v2w_params {
  int32_t num_convs;
  const v2w_conv_params* convs;
  int32_t num_stages;
  const v2w_cond_params* cond;
} v2w_params;

/* Role of a layer in the generator. */
#define V2W_ROLE_CONV_PRE 0
#define V2W_ROLE_UPSAMPLE 1
#define V2W_ROLE_RESBLOCK 2
#define V2W_ROLE_CONV_POST 3

/* Static description of one convolution as a "tap-GEMM":
 *   out[b, q, n] = sum_j sum_c in[b, q + tap_off[j], c] * W[j][n][c]      (rows outside [0,L) read 0)
 * Conv1d: n = co, tap_off[j] = (j - (K-1)/2) * dilation.
 * ConvTranspose1d (polyphase): n = p*Cout + co for output phase p in [0,stride); the output
 * row q of width stride*Cout IS rows q*stride .. q*stride+stride-1 of the channels-last
 * upsampled signal.  Taps are the union over phases; (phase,tap) pairs without a kernel
 * element hold zero weights. */
typedef struct v2w_layer_desc {
  int32_t role, stage;
  int32_t conv_index;          /* index into v2w_params.convs */
  int32_t cin, cout;
  int32_t up;                  /* 1 for Conv1d, stride for ConvTranspose1d */
  int32_t n_total;             /* up * cout */
  int32_t kernel_size, dilation, padding;
  int32_t num_taps;            /* size of the union tap set */
  int32_t tap_off[V2W_MAX_TAPS];
  /* tap_mask[j] bit p set  <=>  (tap j, phase p) carries a kernel element */
  uint32_t tap_mask[V2W_MAX_TAPS];
  int64_t w16_off;             /* byte offset in the packed buffer: 16-bit [2][num_taps][n_total][cin]; [0] = the
                                  operands, [1] = their lo halves (V2W_HALF_FP16_SPLIT buffers, else zeros) */
  int64_t w32_off;             /* byte offset: fp32 [num_taps][cin][n_total] */
  int64_t bias_off;            /* byte offset: fp32 [cout] */
  int64_t macs_per_row;        /* real multiply-accumulates per input row (zero blocks excluded) */
} v2w_layer_desc;

/* Where the folded conditioning parameters of one stage live in the packed buffer (byte
 * offsets, all fp32).  snw is weight_orig already divided by sigma = u.(W v). */
typedef struct v2w_cond_desc {
  int32_t channels;            /* C_i */
  int32_t in_dim;              /* spk_dim + noise_dim */
  int64_t fcw_off;             /* [128][in_dim] */
  int64_t fcb_off;             /* [128] */
  int64_t snw_off;             /* [2C][128] */
  int64_t snb_off;             /* [2C] */
  int64_t rstd_off;            /* [C]  1/sqrt(running_var + 1e-5) */
  int64_t mean_off;            /* [C] */
} v2w_cond_desc;

/* Fused multi-receptive-field stage (kernels_fused.cuh, kernels_stacked.cuh).  All resblocks of one
 * stage run in one kernel on 256-row windows: per branch conv1 -> (bias, +x, leaky-relu, 16-bit) ->
 * shared memory -> conv2; all conv2 outputs plus ONE extra convolution `sum` (taps = per-offset
 * sums of the conv1 weights) accumulate in a single TMEM accumulator T2.
 *
 * The stage is described as a program of SEGMENTS; a segment is a run of MMA ENTRIES (one entry =
 * one operand row shift and one weight block) into one accumulator range.
 *
 *  plain program (C = 128): every entry is one C x C tap matrix, segment types say what it feeds.
 *  stacked program (C <= 64): the conv1 taps of all branches that share a time offset, and the
 *    `sum` tap of that offset, are ONE entry whose weight block stacks them along N -- the operand
 *    tile is read once for up to four tap matrices.  TMEM columns of a 128-row sub-tile are
 *    [T2 | T1_{nb-1} | ... | T1_0] (C columns each) with branches sorted by kernel size, so the
 *    blocks sharing an offset are always a contiguous column range [seg_dcol, seg_dcol + seg_n).
 *    Phase-A segments (CONV1 / SUM, read the input tile) come first, then the phase-B segments
 *    (CONV2, read A1) of the same tile. */
#define V2W_FUSED_MAX_SEG 16
#define V2W_FUSED_MAX_TAPS 96
#define V2W_FUSED_SEG_CONV1 0
#define V2W_FUSED_SEG_SUM 1
#define V2W_FUSED_SEG_CONV2 2
/* wait / commit bits of a stacked segment */
#define V2W_FUSED_BIT_T1(j) (1 << (j))       /* accumulator of branch slot j (wait: free, commit: complete) */
#define V2W_FUSED_BIT_T2 (1 << 4)
#define V2W_FUSED_BIT_A0 (1 << 5)            /* wait: input tile landed; commit: input tile no longer read */
#define V2W_FUSED_BIT_A1(j) (1 << (8 + (j))) /* wait only: A1 of slot j written */
typedef struct v2w_fused_desc {
  int32_t ok;                  /* 0: this stage is executed as separate tap-GEMM ops */
  int32_t channels, num_branches;
  int32_t first_layer;         /* layer index of branch 0's first conv (2 layers per branch) */
  int32_t hc, h1, rc;          /* conv2 halo, conv1 halo, central rows per 256-row window */
  int32_t a0_boxes;            /* 64-row TMA boxes per input tile */
  int32_t num_segments, num_taps;      /* num_taps = MMA entries */
  int32_t stacked;
  int32_t pair;                /* 1: one ResBlock1 conv pair c2(lrelu(c1(a) + b1)) + b2 + x (models.py:37-44): no `sum`
                                  convolution, the middle activation does NOT see the residual, bsum = b2 */
  int32_t preload;             /* 1: T1 pre-loaded with x + b1 by the epilogue; conv1 accumulates on top (C = 64) */
  int32_t post;                /* 1: last stage with conv_post folded into its final pass (hc is 3 rows wider:
                                  every central row finds its 7 neighbours inside the window) */
  int32_t order[V2W_MAX_KERNELS];      /* stacked: model branch that runs as slot j */
  int32_t t_bufs, a0_bufs;             /* stacked: accumulator sets / input tiles in flight */
  int32_t seg_type[V2W_FUSED_MAX_SEG], seg_branch[V2W_FUSED_MAX_SEG];   /* branch: slot of a CONV2 */
  int32_t seg_ntaps[V2W_FUSED_MAX_SEG], seg_first[V2W_FUSED_MAX_SEG];
  int32_t seg_flags[V2W_FUSED_MAX_SEG];  /* plain: bit0 last reader of the input tile, bit1 last segment */
  int32_t seg_dcol[V2W_FUSED_MAX_SEG], seg_n[V2W_FUSED_MAX_SEG];    /* accumulator columns / MMA N */
  int32_t seg_acc[V2W_FUSED_MAX_SEG];    /* 0: the first MMA of the segment overwrites */
  int32_t seg_wait[V2W_FUSED_MAX_SEG], seg_commit[V2W_FUSED_MAX_SEG];
  int32_t shift[V2W_FUSED_MAX_TAPS];   /* operand row of the entry relative to its tile / window */
  int32_t offset[V2W_FUSED_MAX_TAPS];  /* time offset of the entry (rows) */
  int32_t w_row[V2W_FUSED_MAX_TAPS];   /* first row of the entry's block in the weight matrix */
  int32_t w_rows;              /* rows of the weight matrix */
  int64_t w_off;               /* packed: 16-bit [w_rows][channels] (n, c), entries in issue order */
  int64_t bsum_off;            /* packed: fp32 [channels] = sum_j (b1_j + b2_j) */
} v2w_fused_desc;

/* Wide multi-receptive-field stage (ResBlock2, C >= 64): the stage as TWO tap-GEMM launches of the
 * generic tcgen05 kernel instead of 2 * num_kernels -- the same algebra as the fused stage kernels,
 *
 *     sum_j out_j = [ sum_j conv2_j(A1_j) + conv_sum(A0) ] + sum_j (b1_j + b2_j) + nk * X0
 *     A1_j = half(lrelu(conv1_j(A0) + b1_j + X0))                  (vec2wav/models.py:65-70,135-141)
 *
 * laid out for the wide stages (C = 256), where no window of the fused kernels fits TMEM:
 *   launch 1 (V2W_OP_MRF_WIDE1): ONE GEMM with N = nk * C, column tile t = branch t with branch t's tap
 *     list, input A0; the epilogue adds b1_t and the SAME residual X0 to every branch's columns and
 *     writes only the 16-bit A1 = [A1_0 | ... | A1_{nk-1}] into a [rows, (nk + 1) * C] tensor whose
 *     last C columns already hold A0 (written there by the stage's transposed conv).  No fp32
 *     intermediate reaches memory.
 *   launch 2 (V2W_OP_MRF_WIDE2): ONE GEMM with N = C and K = (nk + 1) * C in nk + 1 K-groups: group j < nk
 *     = conv2_j over A1_j's channels with conv2_j's own weights and tap list, group nk = the `sum`
 *     convolution over A0 (per-offset sums of the conv1 weights, `wsum_off`); the MRF sum happens in the
 *     accumulator.  Epilogue: + bsum + nk * X0, * 1/nk, leaky-relu, 16-bit (and / or fp32) output.
 * 21 tap matrices become 21 + 21 + 11 (the `sum` convolution is the price), fp32 round trips of six
 * epilogues become one 16-bit tensor written once and read once. */
typedef struct v2w_wide_desc {
  int32_t ok;
  int32_t channels, num_branches;
  int32_t first_layer;         /* layer index of branch 0's first conv (2 layers per branch) */
  int32_t sum_taps;            /* taps of the `sum` convolution: union of the conv1 offsets */
  int32_t sum_off[V2W_MAX_TAPS];
  int64_t wsum_off;            /* packed: 16-bit [sum_taps][channels][channels] (n, c) */
  int64_t b1cat_off;           /* packed: fp32 [num_branches * channels]: b1 of branch 0, 1, ... */
  int64_t bsum_off;            /* packed: fp32 [channels] = sum_j (b1_j + b2_j) */
} v2w_wide_desc;

/* Kinds of op in a plan's execution list. */
#define V2W_OP_TRANSPOSE_IN 0  /* x[B,Cin,T] fp32 -> [B,T,Cin] (fp32 or bf16) */
#define V2W_OP_COND_AFFINE 1   /* fcs + spectral-norm Linear + BN fold -> scale/shift per (b, stage, c) */
#define V2W_OP_TAPGEMM 2
#define V2W_OP_CONV_POST 3     /* lrelu(0.01) -> Conv1d(16->1,k7) -> tanh */
#define V2W_OP_MRF_FUSED 4     /* one whole multi-receptive-field stage (ResBlock2, C <= 128) */
#define V2W_OP_PAIR_FUSED 5    /* one ResBlock1 conv pair (C <= 64): conv, leaky-relu, conv, + residual (+ MRF sum) */
#define V2W_OP_MRF_WIDE1 6     /* wide stage, launch 1: the first convs of all branches N-stacked (v2w_wide_desc) */
#define V2W_OP_MRF_WIDE2 7     /* wide stage, launch 2: the second convs + the `sum` convolution K-grouped */

#define V2W_ENGINE_SIMT 0      /* FFMA tap-GEMM */
#define V2W_ENGINE_TC 1        /* tcgen05 / TMEM / TMA tap-GEMM */

/* One executed op (introspection for tests, DESIGN.md tables and bench accounting).
 * Buffer fields are byte offsets into the bound workspace, or -1 when unused. */
typedef struct v2w_op_desc {
  int32_t kind, engine;
  int32_t layer;               /* index into the model's layer table, -1 if none */
  int32_t batch, rows;         /* GEMM rows per utterance (input time steps of this op) */
  int32_t n_tile, n_tiles;
  int32_t num_taps[V2W_MAX_TILES];
  int32_t tap_index[V2W_MAX_TILES][V2W_MAX_TAPS]; /* indices into the layer's tap set */
  int64_t in_act;              /* input activation: bf16 (TC) / fp32 or bf16 (SIMT) [B,rows,cin] */
  int32_t in_is_bf16;
  float in_slope;              /* leaky-relu slope applied when LOADING the input (1 = none) */
  int32_t cond_stage;          /* >=0: apply scale/shift of that stage after bias (CBN fold) */
  int64_t residual;            /* fp32 [B,rows,n_total] added after the affine */
  int64_t accum_in;            /* fp32 running MRF sum added next */
  float out_scale;             /* multiplies the total (1/num_kernels on the last MRF branch) */
  int64_t out_f32;             /* fp32 result [B,rows,n_total] */
  int64_t out_bf16;            /* bf16 leaky_relu(result, out_slope): the next conv's operand */
  float out_slope;
  int64_t flops;               /* 2 * real MACs of this op (0 for non-GEMM ops) */
  int32_t stage;               /* upsample stage of an MRF_FUSED / PAIR_FUSED op, else -1 */
  int32_t pair;                /* PAIR_FUSED: index for v2w_model_get_pair, else -1 */
  int32_t post;                /* MRF_FUSED: 1 when conv_post + tanh run inside this op (no CONV_POST op follows) */
  int32_t out_bf16_pitch;      /* row pitch of out_bf16 in elements; 0 = dense (n_total).  The transposed conv of a
                                  wide stage writes its 16-bit output into the last C columns of the stage's
                                  [rows, (nk + 1) * C] operand tensor */
  int32_t in_pitch;            /* row pitch of in_act in elements; 0 = dense (cin) */
  int32_t f32_blocked;         /* TAPGEMM: out_f32 is written in the row-blocked layout [batch][ceil(L/32)][cout/4][32][4]
                                  (L = rows * up output rows of cout channels); MRF_FUSED: `residual` is read in it.  The
                                  fused stage kernels read their fp32 residual stream row-per-thread: in this layout a
                                  warp's float4 load of one column quad is 512 contiguous bytes. */
  int32_t out_blocked;         /* PAIR_FUSED: out_f32 is written in the row-blocked layout (f32_blocked of a PAIR_FUSED op says
                                  the same of its `residual` input) */
  int32_t acc_blocked;         /* PAIR_FUSED: accum_in is read in the row-blocked layout */
  int32_t split;               /* 1: in_act / out_bf16 are [rows, hi | lo] tensors of 16-bit pairs (pitches 2 cin / 2 cout),
                                  three-term product: V2W_PREC_FP32_TC plans (fp16 pairs) and BF16 plans of ResBlock1
                                  generators (bf16 pairs: single bf16 values cost that 97-conv chain the 45 dB bar,
                                  DESIGN.md section 5).  2: weights only -- A (W_hi + W_lo), two MMAs per product, plain
                                  16-bit activations ($V2W_RB1_BF16=weights) */
} v2w_op_desc;

/* ---- library ------------------------------------------------------------------------ */
int32_t v2w_abi_version(void);
const char* v2w_last_error(void);
/* Number of CUDA devices usable by this build (compute capability 10.x); 0 on a CPU box. */
int32_t v2w_device_count(void);

/* ---- model (host only) -------------------------------------------------------------- */
int32_t v2w_model_create(const v2w_config* cfg, struct v2w_model** out);
void v2w_model_destroy(struct v2w_model* m);
int32_t v2w_model_num_layers(const struct v2w_model* m);
int32_t v2w_model_get_layer(const struct v2w_model* m, int32_t index, v2w_layer_desc* out);
int32_t v2w_model_get_cond(const struct v2w_model* m, int32_t stage, v2w_cond_desc* out);
int32_t v2w_model_get_fused(const struct v2w_model* m, int32_t stage, v2w_fused_desc* out);
int32_t v2w_model_get_wide(const struct v2w_model* m, int32_t stage, v2w_wide_desc* out);
int32_t v2w_model_num_pairs(const struct v2w_model* m);   /* fusable ResBlock1 conv pairs (0 for ResBlock2) */
int32_t v2w_model_get_pair(const struct v2w_model* m, int32_t index, v2w_fused_desc* out);
int32_t v2w_model_num_convs(const struct v2w_model* m);   /* expected v2w_params.num_convs */
int64_t v2w_model_packed_bytes(const struct v2w_model* m);
/* Fold weight norm (w = g*v/||v||, norm over all dims but 0 -- i.e. per INPUT channel for
 * ConvTranspose1d), spectral norm (sigma = u.(W v), eval semantics of modules.py:16) and the
 * BN running statistics, and lay everything out in the packed format.  `packed_host` must
 * hold v2w_model_packed_bytes() bytes.  Pure host arithmetic, replaces the per-forward
 * weight_norm/spectral_norm pre-forward hooks of the reference. */
int32_t v2w_model_pack(const struct v2w_model* m, const v2w_params* params, void* packed_host);
/* Same with an explicit 16-bit operand format (v2w_model_pack == V2W_HALF_BF16).  A BF16 plan must be
 * bound to a V2W_HALF_BF16 buffer and an FP16 plan to a V2W_HALF_FP16 one (checked in v2w_plan_bind;
 * an FP32 plan accepts either). */
int32_t v2w_model_pack_half(const struct v2w_model* m, const v2w_params* params, void* packed_host,
                            int32_t half_format);
/* Re-fold ONLY the conditioning section (fcs, spectral-norm sigma, BN running statistics) -- what changes
 * when train()-mode forwards moved u, v and the running statistics but no convolution.  The section is
 * the byte range [offset, offset + bytes) of the packed format (v2w_model_cond_region); `region_host`
 * receives exactly those bytes, to be copied over the same range of every device copy of the packed
 * buffer (plans bound to it stay valid). */
int32_t v2w_model_cond_region(const struct v2w_model* m, int64_t* offset, int64_t* bytes);
int32_t v2w_model_pack_cond(const struct v2w_model* m, const v2w_cond_params* cond, int32_t num_stages,
                            void* region_host);
/* Algorithmic FLOPs (2*MAC, zero polyphase blocks excluded) per input frame. */
int64_t v2w_model_flops_per_frame(const struct v2w_model* m);

/* ---- device-side fold --------------------------------------------------------------- */
/* The same fold as v2w_model_pack_half, on the device: what the reference recomputes in its
 * weight_norm / spectral_norm forward-pre hooks on every call (vec2wav/models.py:5,83-100;
 * vec2wav/modules.py:16) straight from the module's own DEVICE parameter storage into a device packed
 * buffer -- no device-to-host copy of any parameter, stream-ordered, four kernel launches.
 * v2w_packer_create uploads the fold tables of `m` to the CURRENT device (once per model and device).
 * v2w_packer_run: every pointer of `params` is a DEVICE pointer (same layouts as v2w_conv_params /
 * v2w_cond_params); `packed_dev` = v2w_model_packed_bytes() bytes.  The result is bit-identical to
 * the host packer's.  Folding into a buffer that plans are bound to is allowed: the plans stay valid
 * (the format's offsets do not move); call v2w_plan_refresh_weights on each of them afterwards.
 * v2w_packer_run_cond re-folds only the conditioning section (cf. v2w_model_pack_cond). */
struct v2w_packer;
int32_t v2w_packer_create(const struct v2w_model* m, struct v2w_packer** out);
int32_t v2w_packer_run(struct v2w_packer* pk, const v2w_params* params_dev, void* packed_dev,
                       int32_t half_format, void* stream, int32_t* gpu_launches);
int32_t v2w_packer_run_cond(struct v2w_packer* pk, const v2w_cond_params* cond_dev, int32_t num_stages,
                            void* packed_dev, void* stream);
void v2w_packer_destroy(struct v2w_packer* pk);

/* ---- plan --------------------------------------------------------------------------- */
int32_t v2w_plan_create(const struct v2w_model* m, int32_t batch, int32_t frames,
                        int32_t precision, struct v2w_plan** out);
void v2w_plan_destroy(struct v2w_plan* p);
/* 16-bit operand format (V2W_HALF_*) of the packed weight buffer this plan must be bound to.  BF16 for an FP32
 * plan (it reads the fp32 arrays every buffer holds). */
int32_t v2w_plan_packed_format(const struct v2w_plan* p);
int64_t v2w_plan_workspace_bytes(const struct v2w_plan* p);
int32_t v2w_plan_num_ops(const struct v2w_plan* p);
int32_t v2w_plan_get_op(const struct v2w_plan* p, int32_t index, v2w_op_desc* out);
/* Launch geometry of op `index` of a BOUND plan (introspection: tests assert that the intended code paths are
 * the ones that run; tools print it).  For ops of the generic tensor-core tap-GEMM (TAPGEMM with V2W_ENGINE_TC,
 * MRF_WIDE1/2) every field is set; for other ops only grid / smem_bytes / cta_pair, the rest 0. */
typedef struct v2w_launch_desc {
  int32_t grid, smem_bytes;
  int32_t cta_pair;      /* 1: clusters of two CTAs (cta_group::2) */
  int32_t k_chunk;       /* K elements per shared-memory chunk (16 / 32 / 64) */
  int32_t msub;          /* 128-row sub-tiles per tile */
  int32_t stages_a, stages_w, tap_group;   /* operand / weight ring depths; taps per weight stage */
  int32_t y_cols, h_cols;                  /* TMA store box widths of the fp32 / 16-bit output */
  int32_t epi_mode;      /* compiled flavour of the epilogue's chunk loop, 0 = run-time tests */
  int32_t epi_wide;      /* 32 columns per epilogue step */
  int32_t epi_phased;    /* phase-interleaved walk of a transposed conv with a row-blocked fp32 output */
  int32_t split;         /* 0 / 1 (operands and weights as hi + lo) / 2 (weights only) */
} v2w_launch_desc;
int32_t v2w_plan_get_launch(const struct v2w_plan* p, int32_t index, v2w_launch_desc* out);
/* Debug: force op `index` onto an engine (V2W_ENGINE_*). Only meaningful for TAPGEMM ops of
 * a BF16 plan; must be called before v2w_plan_bind. */
int32_t v2w_plan_set_engine(struct v2w_plan* p, int32_t index, int32_t engine);
/* Attach device memory: `packed_dev` = device copy of the packed weights, `workspace_dev` =
 * v2w_plan_workspace_bytes() bytes, both 256-byte aligned.  Builds the TMA tensor maps. */
int32_t v2w_plan_bind(struct v2w_plan* p, const void* packed_dev, void* workspace_dev);
/* The packed buffer a plan is bound to was re-folded in place (v2w_packer_run on `stream`): re-read the
 * few weights the plan keeps as kernel parameters (conv_post's 7 x C taps and bias) once the fold has
 * finished, and drop captured CUDA graphs (they hold the old values).  Blocks on `stream`. */
int32_t v2w_plan_refresh_weights(struct v2w_plan* p, void* stream);

/* Generator.forward(x, spk_emb, noise) (vec2wav/models.py:116-147), eval semantics.
 *   x_dev     fp32 [batch, num_wv_feat, frames]  channels-first, as collate_fn_tensor emits
 *             (vec2wav/dataset.py:212-213)
 *   spk_dev   fp32 [batch, spk_dim]     noise_dev fp32 [batch, noise_dim]
 *   out_dev   fp32 [batch, 1, 320*frames]
 * Asynchronous on `stream`. `gpu_launches`, if non-NULL, receives the number of kernels
 * enqueued. */
int32_t v2w_forward(struct v2w_plan* p, const float* x_dev, const float* spk_dev,
                    const float* noise_dev, float* out_dev, void* stream,
                    int32_t* gpu_launches);

/* The same forward replayed from a CUDA graph (the latency path: batch-1 / long-form requests, where
 * the ~18 launches of a forward cost more host time than device time).  The first call with a given
 * (x_dev, spk_dev, noise_dev, out_dev) pointer tuple on a plan that has already run once captures the
 * forward (programmatic-dependent-launch edges and the conditioning fork / join included) into a graph
 * and instantiates it; later calls with the same pointers are one cudaGraphLaunch.  Up to 8 pointer
 * tuples are cached per plan (least recently used evicted); after 16 consecutive misses the plan stops
 * capturing and the call is exactly v2w_forward.  `replayed`, if non-NULL, receives 1 when the call was
 * a graph launch.  Results are bit-identical to v2w_forward (same kernels, same launch parameters). */
int32_t v2w_forward_graph(struct v2w_plan* p, const float* x_dev, const float* spk_dev,
                          const float* noise_dev, float* out_dev, void* stream, int32_t* gpu_launches,
                          int32_t* replayed);
/* Same call with HOST buffers (pinned recommended): H2D of the three inputs, the forward,
 * D2H of the waveform, all on `stream`; returns after the stream is synchronised.  Uses
 * staging areas inside the bound workspace. */
int32_t v2w_forward_host(struct v2w_plan* p, const float* x_host, const float* spk_host,
                         const float* noise_host, float* out_host, void* stream,
                         int32_t* gpu_launches);

/* Variable-length batch.  The reference zero-pads features to the longest
 * utterance of a batch (vec2wav/dataset.py:212, utils.py:60-77) and generates the padded tail too.
 * Here utterance b has row_offsets[b+1] - row_offsets[b] frames (1 .. plan frames); features are ONE
 * packed time-major buffer x_rows[sum frames][num_wv_feat] (a dataset item's layout,
 * dataset.py:181-187, concatenated) and the waveforms come back un-padded in ONE packed buffer:
 * utterance b owns samples [row_offsets[b] * up, row_offsets[b+1] * up), up = total upsampling.
 * out_format: V2W_OUT_F32 (float) or V2W_OUT_PCM16 (int16 = round(clamp(y * 32768)), MAX_WAV_VALUE of
 * dataset.py:13).  `row_offsets_host` is a HOST array of batch + 1 ints (validated; passed to the kernels
 * by value, so the call is fully asynchronous; at most 512 utterances per call).  Values of every valid sample equal those of the padded forward: the caller only has
 * to pick plan frames >= longest utterance + v2w_model_halo_frames() (or the longest utterance of
 * the reference batch, whichever is smaller) -- the receptive field does not reach further. */
#define V2W_OUT_F32 0
#define V2W_OUT_PCM16 1
int32_t v2w_forward_ragged(struct v2w_plan* p, const float* x_rows_dev, const int32_t* row_offsets_host,
                           const float* spk_dev, const float* noise_dev, void* out_packed_dev,
                           int32_t out_format, void* stream, int32_t* gpu_launches);
/* Input frames past an utterance's end that can influence its last valid output sample (right
 * receptive field of conv_pre .. conv_post, in frames, rounded up). */
int32_t v2w_model_halo_frames(const struct v2w_model* m);

/* train()-mode forward under no_grad: what ConditionalBatchNorm1d does when
 * module.training is set (vec2wav/modules.py:14,16,20-30; callers vec2wav/train.py:138,167 and the
 * warm-up loops that converge sigma / fill the running statistics):
 *   - BatchNorm1d normalises with the statistics of THIS batch (biased variance over batch x time,
 *     eps 1e-5) and updates running_mean / running_var in place
 *     (r <- (1 - momentum) r + momentum s, unbiased variance for running_var);
 *   - the spectral-norm Linear runs ONE power iteration first (v <- normalize(W^T u),
 *     u <- normalize(W v), eps 1e-12), stores u and v in place and divides by sigma = u.(W v).
 * Everything else (the convolutions) is mode-independent and comes from the packed buffer.  The power
 * iterations and gamma / beta of all stages run up front (sn_train_kernel, cond_affine_kernel); per stage
 * the transposed conv runs twice: once raw (bias only) into a scratch tensor for the batch statistics
 * (bn_stats_kernel, folded into scale/shift by bn_finalize_kernel), then the plan's normal op with the
 * affine in its epilogue.  All pointers are DEVICE fp32
 * tensors of the module itself (the caller bumps num_batches_tracked, an int64 the library never reads).
 * The conditioning section of the packed buffer is not read by this call, so the packed buffer stays
 * usable across train-mode forwards; an eval-mode forward afterwards needs a re-pack (u, v and the
 * running statistics changed).  Padded batches only. */
typedef struct v2w_train_stage {
  const float* fc_weight;      /* fcs.i.weight [128, spk_dim+noise_dim] */
  const float* fc_bias;        /* fcs.i.bias [128] */
  const float* sn_weight_orig; /* cbns.i.layer.weight_orig [2C,128] */
  const float* sn_bias;        /* cbns.i.layer.bias [2C] */
  float* sn_u;                 /* cbns.i.layer.weight_u [2C]   in/out */
  float* sn_v;                 /* cbns.i.layer.weight_v [128]  in/out */
  float* bn_running_mean;      /* cbns.i.batch_nrom.running_mean [C]  in/out */
  float* bn_running_var;       /* cbns.i.batch_nrom.running_var  [C]  in/out */
} v2w_train_stage;
typedef struct v2w_train_state {
  int32_t num_stages;
  float momentum;              /* nn.BatchNorm1d default 0.1 */
  v2w_train_stage stage[V2W_MAX_STAGES];
} v2w_train_state;
int32_t v2w_forward_train(struct v2w_plan* p, const float* x_dev, const float* spk_dev,
                          const float* noise_dev, float* out_dev, const v2w_train_state* state,
                          void* stream, int32_t* gpu_launches);

/* Log-mel spectrogram of a generated waveform: the consumer the reference runs
 * on every generator output in validation and for the mel loss -- mel_spectrogram(y, n_fft, num_mels,
 * sampling_rate, hop_size, win_size, fmin, fmax, center=False) of vec2wav/dataset.py:53-77, called at
 * vec2wav/train.py:172-173,266-268,282-284 with the values of vec2wav/hparams.py:50-61:
 *   reflect-pad by (n_fft - hop)/2 -> STFT (periodic Hann window, one-sided, no centring)
 *   -> sqrt(re^2 + im^2 + 1e-9) -> Slaney mel filterbank (librosa.filters.mel defaults: slaney scale,
 *   area normalisation) -> log(max(., 1e-5)).
 * The STFT is ONE tap-GEMM on the tcgen05 engine: frames are hop-sample rows, the n_fft/hop taps are
 * row shifts of the same TMA tile, the windowed DFT basis is the weight; fp16 operands split in
 * high + low parts along K (x_hi W_hi + x_lo W_hi + x_hi W_lo, fp32 accumulation in TMEM) keep ~22
 * mantissa bits.  Requires hop % 16 == 0, n_fft % hop == 0, n_fft / hop <= 16, win <= n_fft,
 * samples > (n_fft - hop) / 2. */
typedef struct v2w_mel_config {
  int32_t n_fft, hop_size, win_size, num_mels, sampling_rate;
  float fmin, fmax;            /* fmax <= 0: sampling_rate / 2 (the reference passes fmax_for_loss = None) */
} v2w_mel_config;
struct v2w_mel;
int32_t v2w_mel_create(const v2w_mel_config* cfg, int32_t batch, int32_t samples, struct v2w_mel** out);
void v2w_mel_destroy(struct v2w_mel* m);
int32_t v2w_mel_num_frames(const struct v2w_mel* m);          /* samples / hop_size */
int64_t v2w_mel_workspace_bytes(const struct v2w_mel* m);
/* Host copy of the filterbank [num_mels][n_fft/2 + 1] (tests pin it against librosa's formula). */
int32_t v2w_mel_get_filterbank(const struct v2w_mel* m, float* out_host);
/* Attach 256-byte aligned device memory and upload the DFT basis and the filterbank (synchronous). */
int32_t v2w_mel_bind(struct v2w_mel* m, void* workspace_dev);
/* wav_dev fp32 [batch][samples] -> out_dev fp32 [batch][num_mels][frames].  Asynchronous on `stream`. */
int32_t v2w_mel_forward(struct v2w_mel* m, const float* wav_dev, float* out_dev, void* stream,
                        int32_t* gpu_launches);

/* Feature-file reader: NumPy .npy, little-endian float32, C order, shape [T, C]
 * or [1, T, C] as prepare_data.py:87 writes wav2vec features.  v2w_npy_read_f32 copies the rows into
 * `dst` (room for max_rows x cols floats; ideally pinned) and returns the frame count.  Host only. */
int32_t v2w_npy_shape(const char* path, int64_t* rows, int64_t* cols);
int32_t v2w_npy_read_f32(const char* path, float* dst, int64_t max_rows, int64_t cols, int64_t* rows);

/* Batch forms for the file pipeline: one call per batch of utterances, `threads` workers inside the library,
 * no per-file work on the host language's side.  v2w_npy_shapes reads n headers; v2w_npy_read_batch reads
 * n payloads into ONE row buffer, file i at dst + row_off[i] * cols (the caller's prefix sums of the
 * shapes: the packed x_rows layout of v2w_forward_ragged), failing if a file no longer has
 * rows_expected[i] frames; v2w_wav_write_batch writes n 16-bit mono RIFF/WAVE files from one packed PCM
 * buffer (utterance i = pcm[sample_off[i] .. + samples[i])).  On failure the first failing file's
 * message is the error text.  Host only. */
int32_t v2w_npy_shapes(const char* const* paths, int32_t n, int64_t* rows, int64_t* cols, int32_t threads);
int32_t v2w_npy_read_batch(const char* const* paths, int32_t n, float* dst, const int64_t* row_off,
                           const int64_t* rows_expected, int64_t cols, int32_t threads);
int32_t v2w_wav_write_batch(const char* const* paths, int32_t n, const int16_t* pcm, const int64_t* sample_off,
                            const int64_t* samples, int32_t sampling_rate, int32_t threads);

/* Pipelined generation from HOST buffers (throughput API): up to two steps in flight on three
 * library-owned streams -- H2D of step i+1 and D2H of step i-1 overlap the kernels of step i.
 * Staging buffers are owned by the pipeline (cudaMalloc).  Every submitted step still copies its
 * inputs from the host and its whole waveform back; only the overlap differs from
 * v2w_forward_host.  Host buffers must stay valid (and should be pinned) until v2w_pipeline_wait
 * returns; the same buffers may be passed to successive submits. */
struct v2w_pipeline;
int32_t v2w_pipeline_create(struct v2w_plan* p, struct v2w_pipeline** out);
int32_t v2w_pipeline_submit(struct v2w_pipeline* q, const float* x_host, const float* spk_host,
                            const float* noise_host, float* out_host);
/* Blocks until every submitted step has landed in its out_host; returns kernels launched so far. */
int32_t v2w_pipeline_wait(struct v2w_pipeline* q, int64_t* gpu_launches);
void v2w_pipeline_destroy(struct v2w_pipeline* q);

/* Per-op device timing.  When enabled, v2w_forward brackets every op with CUDA events recorded
 * on the caller's stream; after the stream has been synchronised by the caller,
 * v2w_plan_op_time_ms returns the last forward's duration of op `index` in milliseconds
 * (negative on error).  Used by bench.py for the live roofline numbers. */
int32_t v2w_plan_set_profiling(struct v2w_plan* p, int32_t enabled);
float v2w_plan_op_time_ms(struct v2w_plan* p, int32_t index);

/* ---- kernel-level entry (unit tests; the plan executor goes through the same code) -- */
typedef struct v2w_tapgemm_args {
  int32_t engine;              /* V2W_ENGINE_* */
  int32_t batch, rows, cin, n_total, cmod; /* cmod: period of bias/affine columns (= cout) */
  int32_t n_tile, n_tiles;
  int32_t num_taps[V2W_MAX_TILES];
  int32_t tap_off[V2W_MAX_TILES][V2W_MAX_TAPS];   /* input-row offsets */
  int32_t tap_row[V2W_MAX_TILES][V2W_MAX_TAPS];   /* tap index j: weights at W[j] */
  const void* in_act;          /* dev: bf16 or fp32 [batch,rows,cin] */
  int32_t in_is_bf16;
  float in_slope;
  const void* w16;             /* dev: bf16 [taps][n_total][cin] (TC engine) */
  const float* w32;            /* dev: fp32 [taps][cin][n_total] (SIMT engine) */
  const float* bias;           /* dev: [cmod] or NULL */
  const float* scale;          /* dev: [batch][cmod] or NULL */
  const float* shift;          /* dev: [batch][cmod] or NULL */
  const float* residual;       /* dev fp32 [batch,rows,n_total] or NULL */
  const float* accum_in;       /* dev fp32 or NULL */
  float out_scale;
  float* out_f32;              /* dev or NULL */
  void* out_bf16;              /* dev or NULL */
  float out_slope;
  int32_t tc_flags;            /* TC engine: bit 8 (0x100) = 16-bit operands are fp16 instead of bf16 */
} v2w_tapgemm_args;
int32_t v2w_tapgemm(const v2w_tapgemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* V2W_H_ */
