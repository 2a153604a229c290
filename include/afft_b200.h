/*
 * afft_b200 - C ABI of the B200-native AFFT fusion-and-anticipation forward path.
 *
 * The reference (zeyun-zhong/AFFT) is pure Python/PyTorch and has no FFI layer; the seam this
 * library plugs into is the `models/` module API.  The Python shim classes
 * in afft_b200/models/ keep the reference's constructor/forward signatures and parameter names
 * and call the entry points below through ctypes.  Each entry point cites the reference code it
 * replaces.
 *
 * Conventions
 *  - every pointer named *_dev / in the io structs is DEVICE memory owned by the caller
 *    (PyTorch's allocator); the library never frees it and only retains registered weights
 *    (which it copies/packs into its own storage at afft_set_weight time).
 *  - `stream` is a cudaStream_t passed as void*; work is enqueued on it, there is no hidden
 *    synchronisation.
 *  - every function returns an int status (AFFT_OK == 0); none throws or aborts.  The message for
 *    the last failure is available from afft_last_error() (thread-local for the stateless ops,
 *    per-handle for model calls).
 *  - one handle per (device, model); re-entrant across handles (test.py:130 runs DataParallel
 *    replicas from Python threads), not thread-safe within one handle.
 *  - plain C types only: no torch / C++ types cross the boundary.
 */
#ifndef AFFT_B200_H_
#define AFFT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define AFFT_API __attribute__((visibility("default")))
#else
#define AFFT_API
#endif

#define AFFT_OK 0
#define AFFT_ERR_INVALID 1     /* bad argument / unsupported configuration */
#define AFFT_ERR_CUDA 2        /* a CUDA runtime/driver call failed */
#define AFFT_ERR_MISSING 3     /* forward called before all weights were registered */
#define AFFT_ERR_UNSUPPORTED 4 /* device is not sm_100 */

#define AFFT_MAX_MODS 8
#define AFFT_MAX_CLS 4
#define AFFT_NAME_LEN 32

/* ------------------------------------------------------------------------------------------ */
/* Library / device                                                                            */
/* ------------------------------------------------------------------------------------------ */

/* ABI version of this header (bumped on any struct change). */
AFFT_API int afft_abi_version(void);

/* Last error message of the calling thread (stateless ops) - never NULL. */
AFFT_API const char* afft_last_error(void);

/* ------------------------------------------------------------------------------------------ */
/* Stateless operators (one per library call the reference makes on the path)                  */
/* ------------------------------------------------------------------------------------------ */

/* AFFT_ACT_RELU: nn.ReLU (models/feature_mapping.py:81-88 NonLinear, models/fusion.py:41,44 MATT).
 * AFFT_ACT_GATE: out = residual * sigmoid(A . W^T + bias) - ContextGating's cat + glu (models/feature_mapping.py:21-33);
 *                the residual operand is required and multiplies instead of adding. */
/* Operand arithmetic of the dense contractions (GEMM accumulation, the residual stream, LayerNorm and softmax
 * statistics are fp32 in every mode; the reference computes everything in fp32):
 *  AFFT_PREC_BF16   bf16 operands, one tensor-core pass.
 *  AFFT_PREC_BF16X3 "strict": error-compensated bf16 hi/lo pairs, hi.hi + hi.lo + lo.hi (three passes, fp32-grade products).
 *  AFFT_PREC_FP16   fp16 operands, one pass at the bf16 rate, 11 significand bits (8x less operand rounding than bf16);
 *                   conversions saturate to +-65504. */
enum { AFFT_PREC_BF16 = 0, AFFT_PREC_BF16X3 = 1, AFFT_PREC_FP16 = 2 };

enum { AFFT_ACT_NONE = 0, AFFT_ACT_GELU_ERF = 1, AFFT_ACT_GELU_TANH = 2, AFFT_ACT_RELU = 3, AFFT_ACT_GATE = 4 };

/*
 * C = epilogue(A . W^T): replaces torch.nn.Linear / transformers Conv1D
 * (models/feature_mapping.py:60,74; models/transformerblock.py:21,34,85-87;
 *  models/future_prediction.py:108,149,248,254,267,269; GPT-2 c_attn/c_proj/c_fc).
 * A [M,K] and W [N,K] are bf16 (fp16 with AFFT_PREC_FP16), K contiguous, 16-byte aligned with 16-byte multiple pitches.
 * precision AFFT_PREC_BF16X3: operands are hi/lo bf16 pairs and the product is hi.hi + hi.lo + lo.hi.
 * The 16-bit output (out_hi) has the operand format of the precision (it feeds the next GEMM).
 * Epilogue, in order: + bias[N] -> activation -> + residual -> stores.
 * Output row of GEMM row r: (r / row_group) * row_stride + r % row_group + row_off
 * (row_group == 0: r).  The residual is read at the same mapped row, or at row r % res_mod when
 * res_mod > 0.
 */
typedef struct afft_gemm_desc {
  const void* a_hi;
  const void* a_lo; /* strict only */
  int64_t lda;
  const void* w_hi;
  const void* w_lo; /* strict only */
  int64_t ldw;
  int32_t M, N, K;
  int32_t precision; /* AFFT_PREC_* */
  const float* bias;
  const float* res;
  int64_t ld_res;
  int32_t res_mod;
  int32_t act;
  float* out_f32;
  int64_t ld_f32;
  void* out_hi; /* bf16 (fp16 with AFFT_PREC_FP16) */
  void* out_lo; /* bf16, AFFT_PREC_BF16X3 producers */
  int64_t ld_bf16;
  int32_t row_group, row_stride, row_off;
  int32_t force_block_n; /* 0 = auto, 128 or 256 */
} afft_gemm_desc;

AFFT_API int afft_gemm(const afft_gemm_desc* d, void* stream);

/* Process-wide choice of the large-M GEMM epilogue: 0 (default) = shared-memory transposition + coalesced global
 * accesses; 1 = TMA-staged (fp32 residual fetched by cp.async.bulk.tensor one slab ahead, outputs written with
 * cp.async.bulk.tensor stores).  Same results bit for bit; measured slower or equal on every shape of the path
 * (DESIGN.md section 4.1), kept for A/B runs.  Initial value: environment variable AFFT_GEMM_EPI_V2. */
AFFT_API int afft_set_gemm_epilogue(int32_t v2);

/* Process-wide: GEMMs with at most 96 rows (bf16 / fp16 operands) run the weight-streaming mma.sync kernel
 * (csrc/gemm_skinny.cuh) instead of the tcgen05 kernels: 1 (default) = on, 0 = off (A/B runs, tests of the tcgen05 tail
 * handling).  Initial value: environment variable AFFT_GEMM_SKINNY. */
AFFT_API int afft_set_gemm_skinny(int32_t on);

/* fp32 [rows, cols] (pitch lds) -> bf16 hi (+ lo when lo != NULL), pitch ldd; transpose != 0
 * writes dst[c, r].  Weight packing (Conv1D [in,out] -> K-major) and feature inputs. */
AFFT_API int afft_convert_bf16(const float* src, int64_t lds, int32_t rows, int32_t cols, void* hi, void* lo, int64_t ldd,
                      int32_t transpose, void* stream);
/* The same for any precision: AFFT_PREC_FP16 writes saturated fp16 to hi (lo must be NULL), AFFT_PREC_BF16X3 needs lo. */
AFFT_API int afft_convert_operand(const float* src, int64_t lds, int32_t rows, int32_t cols, void* hi, void* lo, int64_t ldd,
                                  int32_t transpose, int32_t precision, void* stream);

/* LayerNorm over the last dim: replaces nn.LayerNorm (models/fusion.py:281,362;
 * models/transformerblock.py:132,134,157-161; GPT-2 ln_1/ln_2/ln_f).  dim in {512,1024,2048}. */
typedef struct afft_layernorm_desc {
  const float* x;
  int64_t ldx;
  int32_t in_group, in_stride; /* input row of output row r: (r / in_group)*in_stride + r % in_group (0: r) */
  int32_t n_avg, avg_stride;   /* y = mean over s < n_avg of LN(x[in_row + s*avg_stride]) (<= 1: plain LN) */
  const float* gamma; /* may be NULL (elementwise_affine=False) */
  const float* beta;
  float eps;
  int32_t rows, dim;
  float* y_f32; /* each output optional */
  void* y_hi;
  void* y_lo;
  int64_t ldy;
  int32_t aux_mod, aux_stride, aux_rem; /* rows with r % aux_mod == aux_rem are also written to aux row (r/aux_mod)*aux_stride */
  float* aux_f32;
  void* aux_hi;
  void* aux_lo;
  int64_t ld_aux;
  int32_t out_fp16; /* y_hi / aux_hi are fp16 (AFFT_PREC_FP16 consumers) instead of bf16 */
} afft_layernorm_desc;

AFFT_API int afft_layernorm(const afft_layernorm_desc* d, void* stream);

/* Small multi-head attention (L <= 64, head_dim 256 or 512): replaces the two bmm + softmax of
 * models/transformerblock.py:24-33, :64-74 and GPT-2's eager attention.
 * in_dtype: q/k/v element type AFFT_DT_BF16 / AFFT_DT_F32 (strict mode) / AFFT_DT_FP16; the output keeps a 16-bit input
 * format (fp32 inputs: bf16 hi + optional lo).  Element (seq, i, h, d) of q is at
 * q[(seq*L + i)*ldq + h*head_dim + d].  mask: 0 none, 1 causal, 2 block-causal with period T,
 * 3 diagonal masked.  probs (optional, fp32) element (seq,h,i,j) is at
 * probs[(seq / p_inner)*p_outer + (seq % p_inner)*p_inner_stride + (h*L + i)*L + j]. */
enum { AFFT_DT_BF16 = 0, AFFT_DT_F32 = 1, AFFT_DT_FP16 = 2 };
typedef struct afft_attention_desc {
  const void* q;
  const void* k;
  const void* v;
  int64_t ldq, ldk, ldv;
  int32_t in_dtype;
  int32_t n_seq, L, H, head_dim;
  float scale;
  int32_t mask, T;
  void* out_hi;
  void* out_lo;
  int64_t ldo;
  float* probs;
  int64_t p_outer, p_inner_stride;
  int32_t p_inner;
  /* training only - attention-probability dropout (models/transformerblock.py:31,71; GPT-2 attn_pdrop): optional fp32
   * [n_seq, H, L, L] factors (0 or 1/(1-p)) multiplied into the softmax before P.V.  `probs` receives the UNDROPPED
   * softmax (afft_attention_bwd needs it).  NULL in inference. */
  const float* drop_mask;
} afft_attention_desc;

AFFT_API int afft_attention(const afft_attention_desc* d, void* stream);

/* Score fusion: p = softmax(attn_logits[r, :n_mod]); out[r, :C] = sum_i p[i] * logits[i][r, :C].
 * Replaces MATT's softmax (models/fusion.py:57) and the weighted sum of the per-modality logits in
 * CMFPScoreFusion.forward (models/future_prediction.py:341-350).  attn [rows, n_mod] may be NULL; with out == NULL
 * only the softmax is computed (MATT.forward's return value) and logits / C are ignored.
 * n_mod <= 8; the logits / out pitches are multiples of 4 floats and >= ceil4(C), pointers 16-byte aligned. */
AFFT_API int afft_score_fusion(const float* attn_logits, int64_t ld_a, int32_t n_mod, const float* const* logits,
                               int64_t ld_l, int32_t rows, int32_t C, float* attn, float* out, int64_t ld_o, void* stream);

/* Logit post-processing (the step right after the path): softmax over the action logits,
 * verb/noun marginalisation and top-K ranking - replaces challenge.py:196-210 (scipy softmax + two matmuls with the
 * 0/1 class_mappings matrices) and the argsort ranking of common/utils.py:19-42.  verb_of/noun_of [A] int32 give the
 * verb / noun class of every action (the column of the single 1 in each row of the mapping matrices).
 * probs may be NULL; topk [B, 3, K] int32 (action, verb, noun; K <= 16) may be NULL. */
AFFT_API int afft_marginalize_topk(const float* logits, int64_t ld, int32_t B, int32_t A, const int32_t* verb_of,
                                   const int32_t* noun_of, int32_t n_verb, int32_t n_noun, float* probs, float* verb,
                                   float* noun, int32_t* topk, int32_t K, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Training-step operators (BASELINE config 5: forward + backward; reference train.py:234-262 relies on autograd   */
/* through nn.Linear / nn.LayerNorm / nn.GELU / softmax attention).  dgrad and wgrad are afft_gemm calls on       */
/* transposed bf16 operands; these are the remaining backward kernels.  fp32 in / fp32 out.                       */
/* ------------------------------------------------------------------------------------------ */
/* fp32 [rows, cols] -> bf16 copy `hi` (pitch ldh) and bf16 transpose `tr` [cols, rows] (pitch ldt) in one pass, plus
 * colsum[c] += sum_r src[r, c] (each output optional): the three passes a Linear's backward makes over dy. */
AFFT_API int afft_convert_dual(const float* src, int64_t lds, int32_t rows, int32_t cols, void* hi, int64_t ldh, void* tr,
                               int64_t ldt, float* colsum, void* stream);
/* The same pass with the MLP's activation folded in (kind: AFFT_ACT_GELU_ERF / AFFT_ACT_GELU_TANH).  d_act == NULL:
 * v = gelu(src), the second Linear's operands straight from the first Linear's fp32 output.  d_act != NULL (pitch ldd):
 * v = d_act * gelu'(src), the first Linear's backward operands and bias gradient from the gradient w.r.t. the activation. */
AFFT_API int afft_convert_dual_gelu(const float* src, int64_t lds, int32_t rows, int32_t cols, void* hi, int64_t ldh,
                                    void* tr, int64_t ldt, float* colsum, const float* d_act, int64_t ldd, int32_t kind,
                                    void* stream);
AFFT_API int afft_transpose_bf16(const void* src, int64_t lds, int32_t rows, int32_t cols, void* dst, int64_t ldd, void* stream);
/* dx = LayerNorm backward; dgamma / dbeta are ACCUMULATED (+=) and may be NULL (together with gamma). */
AFFT_API int afft_layernorm_bwd(const float* x, int64_t ldx, const float* gamma, float eps, const float* dy, int64_t lddy,
                                int32_t rows, int32_t dim, float* dx, int64_t lddx, float* dgamma, float* dbeta, void* stream);
/* kind: AFFT_ACT_GELU_ERF or AFFT_ACT_GELU_TANH */
AFFT_API int afft_gelu_fwd(const float* x, float* y, int64_t n, int32_t kind, void* stream);
AFFT_API int afft_gelu_bwd(const float* x, const float* dy, float* dx, int64_t n, int32_t kind, void* stream);
/* out[c] += sum over rows of x[r, c] (bias gradient) */
AFFT_API int afft_colsum(const float* x, int64_t ld, int32_t rows, int32_t cols, float* out, void* stream);
/* torch.optim.SGD's update (momentum, weight decay, optional Nesterov; reference train.py / expts/01_SA-Fuser_ek100_train.txt:48-52)
 * over a flat fp32 parameter buffer: g' = g + wd p; m = momentum m + g'; p -= lr (nesterov ? g' + momentum m : m).  p16 (optional)
 * receives the bf16 image of the updated parameters - the GEMM operands of the next step. */
AFFT_API int afft_sgd_nesterov(float* p, const float* g, float* m, void* p16, int64_t n, float lr, float momentum,
                               float weight_decay, int32_t nesterov, void* stream);
/* Backward of afft_attention for fp32 q|k|v (layout as afft_attention_desc with ldq = ldk = ldv = ld, q at column 0,
 * k at H*head_dim, v at 2*H*head_dim); probs [n_seq, H, L, L] as written by the forward (undropped); d_out [n_seq*L, ldo];
 * drop_mask: the forward's dropout factors or NULL. */
AFFT_API int afft_attention_bwd(const float* qkv, int64_t ld, const float* probs, const float* d_out, int64_t ldo, float* dqkv,
                                int32_t n_seq, int32_t L, int32_t H, int32_t head_dim, float scale, const float* drop_mask,
                                void* stream);

/* ------------------------------------------------------------------------------------------ */
/* Model-level API: everything below BaseModel.future_predictor (models/base_model.py:59)      */
/* ------------------------------------------------------------------------------------------ */

enum {
  AFFT_FUSER_SA = 0,       /* models.fusion.ModalTokenCMFuser        (fusion.py:273-365) */
  AFFT_FUSER_SA_NOTOKEN = 1, /* models.fusion.CMFuser                (fusion.py:61-118)  */
  AFFT_FUSER_TSA = 2,      /* models.fusion.TemporalCMFuser          (fusion.py:121-215) */
  AFFT_FUSER_CA = 3,       /* models.fusion.TemporalCrossAttentFuser (fusion.py:218-270) */
  AFFT_FUSER_NONE = 4      /* no fuser: one modality of width `dim` goes straight to dim_encoder -> GPT-2 -> dim_decoder ->
                              classifier (IndividualFuturePrediction / CMFPScoreFusion, future_prediction.py:200-225,
                              315-327); n_mod must be 1, mod_dim[0] == dim, dim a multiple of 8 */
};

/* Which part of the chain a handle runs (afft_config.stages).  AFFT_STAGE_ALL (0) is the whole of CMFPEarly.forward;
 * the other two are the reference's inner seams:
 *  AFFT_STAGE_FUSER  feature mapping + fuser:  fuser(modal_feats, ordered_feature_list) -> (fused, attn)
 *                    (models/fusion.py:319 and the other fusers' forward); needs the mapping.* / fuser.* weights only,
 *                    writes io.orig_past [B, T, dim] (+ io.fuser_attn); io.past_futures / io.logits are ignored.
 *  AFFT_STAGE_GPT    the GPT-2 future predictor alone:  predictor(feats (B, T, fp_inter_dim), output_len) ->
 *                    (B, T + output_len - 1, fp_inter_dim) (models/future_prediction.py:387-415): fuser_kind =
 *                    AFFT_FUSER_NONE, n_mod = 1, mod_dim[0] = dim = gpt_dim, n_cls = 0; needs future_predictor.gpt_model.*;
 *                    io.feat[0] is the input, io.orig_past receives the T prompt positions [B, T, gpt_dim] and
 *                    io.past_futures the generated ones [B, fp_output_len - 1, gpt_dim] (may be NULL when fp_output_len = 1). */
enum { AFFT_STAGE_ALL = 0, AFFT_STAGE_FUSER = 1, AFFT_STAGE_GPT = 4 };

typedef struct afft_config {
  int32_t fuser_kind;
  int32_t T;                              /* timesteps per clip */
  int32_t n_mod;                          /* modalities present, in fusion order (conf/config.yaml:41) */
  char mod_name[AFFT_MAX_MODS][AFFT_NAME_LEN];
  int32_t mod_dim[AFFT_MAX_MODS];         /* input feature width per modality */
  int32_t dim;                            /* model.common_dim */
  int32_t fuser_depth, fuser_heads;
  int32_t modal_encoding, frame_level_token, cross_attn, norm_elementwise;
  int32_t gpt_dim, gpt_layers, gpt_heads; /* fp_inter_dim, fp_layers, fp_heads */
  int32_t n_cls;
  char cls_name[AFFT_MAX_CLS][AFFT_NAME_LEN];
  int32_t cls_dim[AFFT_MAX_CLS];
  int32_t precision;                      /* AFFT_PREC_BF16 / AFFT_PREC_BF16X3 (strict) / AFFT_PREC_FP16 */
  int32_t max_batch;                      /* workspace is sized for this many clips per call */
  int32_t device;                         /* CUDA device ordinal */
  int32_t fp_output_len;                  /* model.common.fp_output_len: future steps rolled out (>= 1) */
  int32_t stages;                         /* AFFT_STAGE_ALL / AFFT_STAGE_FUSER / AFFT_STAGE_GPT */
} afft_config;

typedef struct afft_handle afft_handle;

AFFT_API int afft_create(const afft_config* cfg, afft_handle** out);
/* The same with a CALLER-OWNED workspace ("all tensors are caller-owned device memory (PyTorch
 * allocator) ... no hidden synchronisation"): afft_workspace_bytes_for() is host arithmetic; afft_create_in() carves the
 * handle's activation buffers out of `workspace_dev` (256-byte aligned, at least that many bytes, alive until
 * afft_destroy), clears its split-K counters with a memset enqueued on `stream` and does not synchronise.  Packed
 * weights remain the library's (afft_set_weight allocates them). */
AFFT_API int afft_workspace_bytes_for(const afft_config* cfg, size_t* bytes);
AFFT_API int afft_create_in(const afft_config* cfg, void* workspace_dev, size_t workspace_bytes, void* stream, afft_handle** out);
AFFT_API void afft_destroy(afft_handle* h);
AFFT_API const char* afft_handle_error(const afft_handle* h);

/* Bytes of device workspace + packed weights the handle holds (for capacity planning). */
AFFT_API size_t afft_workspace_bytes(const afft_handle* h);
AFFT_API size_t afft_weight_bytes(const afft_handle* h);

/*
 * Register one tensor of the reference state dict (train.py:55-103 key contract), name relative
 * to "future_predictor." e.g. "fuser.blocks.0.attn.qkv.weight".  src_dev is fp32 device memory,
 * contiguous, with the reference's shape (ndim <= 3).  Unknown names return AFFT_ERR_INVALID
 * (the Python shim filters GPT-2's attn.bias / attn.masked_bias buffers).
 */
AFFT_API int afft_set_weight(afft_handle* h, const char* name, const float* src_dev, int32_t ndim, const int64_t* shape,
                    void* stream);

/* Number of tensors still missing; names are written, newline separated, into buf (may be NULL). */
AFFT_API int afft_missing_weights(const afft_handle* h, char* buf, size_t buf_len);

typedef struct afft_io {
  const float* feat[AFFT_MAX_MODS]; /* [B, T, mod_dim[m]] fp32, fusion order */
  float* orig_past;                 /* [B, T, dim]       fused features z                        */
  float* past_futures;              /* [B, T+O, dim]     slots 0..T-1 = past_futures, T.. = future (O = fp_output_len) */
  float* logits[AFFT_MAX_CLS];      /* [B, T+O, ld_logits] slots 0..T-1 = past_logits, T.. = logits */
  int64_t ld_logits[AFFT_MAX_CLS];  /* row pitch in floats, multiple of 4, >= cls_dim            */
  float* fuser_attn;                /* SA: [B, depth, T, H, n, n]; T-SA: [B, depth, H, nT, nT]; NULL = skip */
  float* gpt_attn;                  /* fp_output_attentions (models/future_prediction.py:403-409, 'gpt2_att_0'): the GPT-2
                                       attention probabilities of the T prompt positions [B, gpt_layers, gpt_heads, T, T];
                                       NULL = skip */
} afft_io;

/* One forward of CMFPEarly.forward (models/future_prediction.py:257-291) for B <= max_batch clips. */
AFFT_API int afft_forward(afft_handle* h, int32_t B, const afft_io* io, void* stream);

/* Kernels launched by the most recent afft_forward on this handle. */
AFFT_API int afft_last_launch_count(const afft_handle* h);

/*
 * Optional per-launch timing.  With profiling on, every kernel of the next forwards records the latest %globaltimer value
 * any of its warps saw on exit into a device slot; afft_profile_read synchronises the forward's stream and attributes to
 * launch i the interval (end of launch i-1, end of launch i].  The slices add up to the device time of the forward
 * exactly and the marks do not disturb the programmatic-dependent-launch overlap of consecutive kernels (CUDA events
 * between the launches would).  Used by bench.py for the roofline of the GEMM kernel; off by default.
 */
enum { AFFT_CAT_GEMM = 0, AFFT_CAT_LAYERNORM = 1, AFFT_CAT_ATTENTION = 2, AFFT_CAT_OTHER = 3 };
#define AFFT_MAX_PROFILE_RECS 256
typedef struct afft_profile_rec {
  int32_t cat;
  int32_t M, N, K; /* GEMM shape (0 for other kernels) */
  float ms;
} afft_profile_rec;
typedef struct afft_profile {
  int32_t n;
  afft_profile_rec recs[AFFT_MAX_PROFILE_RECS];
} afft_profile;
AFFT_API int afft_profile_enable(afft_handle* h, int32_t enable);
/* Upper bound on the number of K splits a GEMM of this handle may use (split-K runs when a GEMM has too few output
 * tiles to occupy the GPU: small batches).  1 switches split-K off; the default is 4 (the last-arriving CTA reduces all
 * partials of its tile, so deeper splits cost more in the reduction than they save in the K loop - measured).  Results are bit-reproducible
 * for a fixed value (partials are summed in split order), and differ between values only by fp32 summation order.
 * There is no reference counterpart (PyTorch picks its cuBLAS algorithm internally). */
AFFT_API int afft_set_max_ksplit(afft_handle* h, int32_t max_split);
/* The split factor the scheduler would choose for a GEMM of `tiles` output tiles and `num_kb` 64-wide K blocks on
 * `slots` persistent CTAs (or CTA pairs) with the given cap - host arithmetic only (exposed for tests and tuning). */
AFFT_API int afft_plan_ksplit(int32_t tiles, int32_t slots, int32_t num_kb, int32_t ctas_per_tile, int32_t max_split);
AFFT_API int afft_profile_read(afft_handle* h, afft_profile* out);

#ifdef __cplusplus
}
#endif
#endif /* AFFT_B200_H_ */
