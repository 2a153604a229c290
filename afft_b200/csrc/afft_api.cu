// C ABI of afft_b200 (see include/afft_b200.h): stateless operators + the model-level forward of
// everything under BaseModel.future_predictor (reference models/base_model.py:59,
// models/future_prediction.py:257-291).
#include "../../include/afft_b200.h"

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "gemm_launch.cuh"
#include "kernels.cuh"
#include "kernels_train.cuh"

using namespace afft;
typedef __nv_bfloat16 bf16;

// ================================================================================================
// errors
// ================================================================================================
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
static int cuda_fail(const char* what, cudaError_t e) {
  return fail(AFFT_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

extern "C" int afft_abi_version(void) { return 8; }
extern "C" const char* afft_last_error(void) { return g_err.c_str(); }

static int device_sm_count(int* out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail("cudaGetDevice", e);
  static int cached[64] = {0};
  if (dev < 64 && cached[dev] > 0) {
    *out = cached[dev];
    return AFFT_OK;
  }
  int major = 0, sms = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return cuda_fail("cudaDeviceGetAttribute", e);
  if (major != 10) return fail(AFFT_ERR_UNSUPPORTED, "afft_b200 kernels are sm_100a only (device is not compute capability 10.x)");
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return cuda_fail("cudaDeviceGetAttribute", e);
  if (dev < 64) cached[dev] = sms;
  *out = sms;
  return AFFT_OK;
}

// ================================================================================================
// stateless operators
// ================================================================================================
// w_static: the W operand is a packed weight of a handle (never written by a kernel of the stream)
static int run_gemm(const afft_gemm_desc& d, int num_sms, cudaStream_t stream, const SplitKScratch* sk = nullptr,
                    unsigned long long* t_end = nullptr, bool w_static = false) {
  GemmOperands g;
  g.a = static_cast<const bf16*>(d.a_hi);
  g.a_lo = static_cast<const bf16*>(d.a_lo);
  g.lda = d.lda;
  g.w = static_cast<const bf16*>(d.w_hi);
  g.w_lo = static_cast<const bf16*>(d.w_lo);
  g.ldw = d.ldw;
  g.M = d.M;
  g.N = d.N;
  g.K = d.K;
  GemmEpilogue ep;
  ep.bias = d.bias;
  ep.res = d.res;
  ep.ld_res = d.ld_res;
  ep.res_mod = d.res_mod;
  ep.out_f32 = d.out_f32;
  ep.ld_f32 = d.ld_f32;
  ep.out_hi = static_cast<bf16*>(d.out_hi);
  ep.out_lo = static_cast<bf16*>(d.out_lo);
  ep.ld_bf16 = d.ld_bf16;
  ep.act = d.act;
  ep.row_group = d.row_group;
  ep.row_stride = d.row_stride;
  ep.row_off = d.row_off;
  if (g.a == nullptr || g.w == nullptr) return fail(AFFT_ERR_INVALID, "gemm: null operand");
  if (ep.out_f32 == nullptr && ep.out_hi == nullptr) return fail(AFFT_ERR_INVALID, "gemm: no output");
  if (ep.act < ACT_NONE || ep.act > ACT_GATE) return fail(AFFT_ERR_INVALID, "gemm: unknown activation");
  if (ep.act == ACT_GATE && ep.res == nullptr) return fail(AFFT_ERR_INVALID, "gemm: AFFT_ACT_GATE needs the gated operand in res");
  std::string err;
  if (d.precision < AFFT_PREC_BF16 || d.precision > AFFT_PREC_FP16) return fail(AFFT_ERR_INVALID, "gemm: unknown precision");
  if (d.precision == AFFT_PREC_FP16 && (ep.out_lo != nullptr || g.a_lo != nullptr || g.w_lo != nullptr))
    return fail(AFFT_ERR_INVALID, "gemm: AFFT_PREC_FP16 takes no lo operands / outputs");
  const int mode = d.precision == AFFT_PREC_BF16X3 ? MODE_BF16X3 : (d.precision == AFFT_PREC_FP16 ? MODE_FP16 : MODE_BF16);
  if (!launch_gemm(g, ep, mode, d.force_block_n, num_sms, stream, &err, sk, t_end, w_static)) return fail(AFFT_ERR_CUDA, err);
  return AFFT_OK;
}

extern "C" int afft_gemm(const afft_gemm_desc* d, void* stream) {
  if (d == nullptr) return fail(AFFT_ERR_INVALID, "gemm: null descriptor");
  int sms = 0;
  int rc = device_sm_count(&sms);
  if (rc != AFFT_OK) return rc;
  return run_gemm(*d, sms, static_cast<cudaStream_t>(stream));
}

static int run_convert(const float* src, long long lds, int rows, int cols, bf16* hi, bf16* lo, long long ldd,
                       int transpose, cudaStream_t stream, int fp16 = 0, unsigned long long* t_end = nullptr) {
  if (src == nullptr || hi == nullptr || rows <= 0 || cols <= 0) return fail(AFFT_ERR_INVALID, "convert: bad argument");
  if (fp16 && lo != nullptr) return fail(AFFT_ERR_INVALID, "convert: fp16 operands have no lo part");
  if (transpose) {
    dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
    convert_transpose_f32_bf16_kernel<<<grid, block, 0, stream>>>(src, lds, rows, cols, hi, lo, ldd, fp16);
  } else {
    const long long total = static_cast<long long>(rows) * cols;
    int blocks = static_cast<int>(std::min<long long>((total / 8 + 255) / 256 + 1, 148 * 16));
    convert_f32_bf16_kernel<<<blocks, 256, 0, stream>>>(src, lds, rows, cols, hi, lo, ldd, fp16, t_end);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail("convert launch", e);
  return AFFT_OK;
}

extern "C" int afft_convert_bf16(const float* src, int64_t lds, int32_t rows, int32_t cols, void* hi, void* lo,
                                 int64_t ldd, int32_t transpose, void* stream) {
  return run_convert(src, lds, rows, cols, static_cast<bf16*>(hi), static_cast<bf16*>(lo), ldd, transpose,
                     static_cast<cudaStream_t>(stream));
}

extern "C" int afft_convert_operand(const float* src, int64_t lds, int32_t rows, int32_t cols, void* hi, void* lo,
                                    int64_t ldd, int32_t transpose, int32_t precision, void* stream) {
  if (precision < AFFT_PREC_BF16 || precision > AFFT_PREC_FP16) return fail(AFFT_ERR_INVALID, "convert: unknown precision");
  if (precision == AFFT_PREC_BF16X3 && lo == nullptr) return fail(AFFT_ERR_INVALID, "convert: AFFT_PREC_BF16X3 needs the lo output");
  return run_convert(src, lds, rows, cols, static_cast<bf16*>(hi), static_cast<bf16*>(lo), ldd, transpose,
                     static_cast<cudaStream_t>(stream), precision == AFFT_PREC_FP16 ? 1 : 0);
}

static int run_layernorm(const LayerNormArgs& a, cudaStream_t stream) {
  if (a.x == nullptr || a.rows <= 0) return fail(AFFT_ERR_INVALID, "layernorm: bad argument");
  if (a.dim % 128 != 0) return fail(AFFT_ERR_INVALID, "layernorm: dim must be a multiple of 128");
  // one warp per row; CTA size is a tuning knob (AFFT_LN_THREADS = 64 / 128 / 256)
  static const int ln_threads = [] { const char* v = getenv("AFFT_LN_THREADS"); const int t = v ? atoi(v) : 256;
                                     return (t == 64 || t == 128 || t == 256) ? t : 256; }();
  const int rows_per_block = ln_threads / 32;
  const int blocks = (a.rows + rows_per_block - 1) / rows_per_block;
  const bool avg = a.n_avg > 1;
  const bool fast = !avg && a.gamma != nullptr && a.beta != nullptr && a.y_hi != nullptr && a.y_f32 == nullptr &&
                    a.y_lo == nullptr && a.aux_mod == 0;  // (an output row map is honoured by every instantiation)
#define AFFT_LN(NV)                                                                                          \
  case NV:                                                                                                   \
    if (fast) launch_pdl(layernorm_kernel<NV, false, true>, dim3(blocks), dim3(ln_threads), 0, stream, a);   \
    else if (avg) launch_pdl(layernorm_kernel<NV, true, false>, dim3(blocks), dim3(ln_threads), 0, stream, a); \
    else launch_pdl(layernorm_kernel<NV, false, false>, dim3(blocks), dim3(ln_threads), 0, stream, a);       \
    break;
  switch (a.dim / 128) {
    AFFT_LN(2)
    AFFT_LN(4)
    AFFT_LN(6)
    AFFT_LN(8)
    AFFT_LN(12)
    AFFT_LN(16)
    default: return fail(AFFT_ERR_INVALID, "layernorm: unsupported dim (256, 512, 768, 1024, 1536, 2048)");
  }
#undef AFFT_LN
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail("layernorm launch", e);
  return AFFT_OK;
}

extern "C" int afft_layernorm(const afft_layernorm_desc* d, void* stream) {
  if (d == nullptr) return fail(AFFT_ERR_INVALID, "layernorm: null descriptor");
  LayerNormArgs a;
  a.x = d->x;
  a.ldx = d->ldx;
  a.in_group = d->in_group;
  a.in_stride = d->in_stride;
  a.n_avg = d->n_avg;
  a.avg_stride = d->avg_stride;
  a.gamma = d->gamma;
  a.beta = d->beta;
  a.eps = d->eps;
  a.rows = d->rows;
  a.dim = d->dim;
  a.y_f32 = d->y_f32;
  a.y_hi = static_cast<bf16*>(d->y_hi);
  a.y_lo = static_cast<bf16*>(d->y_lo);
  a.ldy = d->ldy;
  a.aux_mod = d->aux_mod;
  a.aux_rem = d->aux_rem;
  a.aux_stride = d->aux_stride;
  a.aux_f32 = d->aux_f32;
  a.aux_hi = static_cast<bf16*>(d->aux_hi);
  a.aux_lo = static_cast<bf16*>(d->aux_lo);
  a.ld_aux = d->ld_aux;
  a.out_fp16 = d->out_fp16 != 0;
  a.t_end = nullptr;
  a.out_group = a.out_stride = a.out_off = 0;
  if (a.out_fp16 && (a.y_lo != nullptr || a.aux_lo != nullptr)) return fail(AFFT_ERR_INVALID, "layernorm: fp16 outputs have no lo part");
  return run_layernorm(a, static_cast<cudaStream_t>(stream));
}

template <typename TIn, int HD>
static int launch_attention(const AttentionArgs& a, cudaStream_t stream) {
  auto kern = attention_small_kernel<TIn, HD>;
  const size_t smem = static_cast<size_t>(2) * a.L * HD * sizeof(TIn);
  if (smem > 48 * 1024) {
    static size_t configured[64] = {0};  // per device
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (smem > configured[dev]) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return cuda_fail("attention smem attribute", e);
      configured[dev] = smem;
    }
  }
  int warps = a.L < 8 ? a.L : 8;
  kern<<<a.n_seq * a.H, warps * 32, smem, stream>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail("attention launch", e);
  return AFFT_OK;
}

template <typename TIn>
static int launch_attention_tokens(const AttentionArgs& a, cudaStream_t stream) {
  const int warps = a.n_seq * a.H;
  const int blocks = (warps + 3) / 4;
  switch (a.L) {
    case 2: attention_tokens_kernel<TIn, 2><<<blocks, 128, 0, stream>>>(a); break;
    case 3: attention_tokens_kernel<TIn, 3><<<blocks, 128, 0, stream>>>(a); break;
    case 4: attention_tokens_kernel<TIn, 4><<<blocks, 128, 0, stream>>>(a); break;
    case 5: attention_tokens_kernel<TIn, 5><<<blocks, 128, 0, stream>>>(a); break;
    case 6: attention_tokens_kernel<TIn, 6><<<blocks, 128, 0, stream>>>(a); break;
    default: return fail(AFFT_ERR_INVALID, "attention_tokens: L must be in [2, 6]");
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail("attention_tokens launch", e);
  return AFFT_OK;
}

template <int L, bool FP16>
static int launch_attention_tokens_mma(const AttentionArgs& a, cudaStream_t stream) {
  auto kern = attention_tokens_mma_kernel<L, FP16>;
  const int smem = a.H * 3 * 16 * (256 * 2 + 16);
  static int configured[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (smem > configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return cuda_fail("attention_tokens_mma smem attribute", e);
    configured[dev] = smem;
  }
  constexpr int G = 16 / L;
  cudaError_t e = launch_pdl(kern, dim3((a.n_seq + G - 1) / G), dim3(32 * a.H), smem, stream, a);
  if (e != cudaSuccess) return cuda_fail("attention_tokens_mma launch", e);
  return AFFT_OK;
}

template <int HD, int LP, int NW, bool FP16>
static int launch_attention_mma(const AttentionArgs& a, cudaStream_t stream) {
  auto kern = attention_mma_kernel<HD, LP, NW, FP16>;
  constexpr int smem = AttnMmaSmem<HD, LP>::kBytes;
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return cuda_fail("attention_mma smem attribute", e);
    configured[dev] = true;
  }
  cudaError_t e = launch_pdl(kern, dim3(a.n_seq * a.H), dim3(NW * 32), AttnMmaSmem<HD, LP>::bytes_for(a.L), stream, a);
  if (e != cudaSuccess) return cuda_fail("attention_mma launch", e);
  return AFFT_OK;
}

template <bool FP16>
static int run_attention_mma_tokens(const AttentionArgs& a, cudaStream_t stream) {
  switch (a.L) {
    case 2: return launch_attention_tokens_mma<2, FP16>(a, stream);
    case 3: return launch_attention_tokens_mma<3, FP16>(a, stream);
    case 4: return launch_attention_tokens_mma<4, FP16>(a, stream);
    case 5: return launch_attention_tokens_mma<5, FP16>(a, stream);
    default: return launch_attention_tokens_mma<6, FP16>(a, stream);
  }
}

// in_dtype: AFFT_DT_BF16 / AFFT_DT_F32 / AFFT_DT_FP16
static int run_attention(const AttentionArgs& a, int head_dim, int in_dtype, cudaStream_t stream) {
  if (a.q == nullptr || a.k == nullptr || a.v == nullptr || a.out_hi == nullptr)
    return fail(AFFT_ERR_INVALID, "attention: null pointer");
  if (a.L < 1 || a.L > 64) return fail(AFFT_ERR_INVALID, "attention: sequence length must be in [1, 64]");
  if (a.n_seq <= 0 || a.H <= 0) return fail(AFFT_ERR_INVALID, "attention: empty problem");
  if (in_dtype < AFFT_DT_BF16 || in_dtype > AFFT_DT_FP16) return fail(AFFT_ERR_INVALID, "attention: unknown input dtype");
  const bool in_f32 = in_dtype == AFFT_DT_F32, fp16 = in_dtype == AFFT_DT_FP16;
  if (fp16 && a.out_lo != nullptr) return fail(AFFT_ERR_INVALID, "attention: fp16 outputs have no lo part");
  static const int use_mma = [] { const char* v = getenv("AFFT_ATTN_MMA"); return v == nullptr ? 1 : atoi(v); }();
  // few modality tokens per timestep (SA-Fuser): tensor-core kernel for 16-bit inputs, register-resident
  // warp-per-(timestep, head) kernel for fp32 inputs (strict mode)
  if (head_dim == 256 && a.L >= 2 && a.L <= 6 && (a.mask == 0 || a.mask == 3)) {
    if (use_mma && !in_f32 && a.out_lo == nullptr && a.H <= 8 && a.drop == nullptr)
      return fp16 ? run_attention_mma_tokens<true>(a, stream) : run_attention_mma_tokens<false>(a, stream);
    return in_f32 ? launch_attention_tokens<float>(a, stream)
                  : (fp16 ? launch_attention_tokens<__half>(a, stream) : launch_attention_tokens<bf16>(a, stream));
  }
  // short sequences with 16-bit inputs (GPT-2 predictor, CA-Fuser): tensor-core (mma.sync) kernel
  if (use_mma && !in_f32 && a.out_lo == nullptr && a.L > 6 && a.mask >= 0 && a.mask <= 2 && a.drop == nullptr) {
    if (a.L <= 32) {
      if (head_dim == 256) return fp16 ? launch_attention_mma<256, 32, 4, true>(a, stream) : launch_attention_mma<256, 32, 4, false>(a, stream);
      if (head_dim == 512) return fp16 ? launch_attention_mma<512, 32, 4, true>(a, stream) : launch_attention_mma<512, 32, 4, false>(a, stream);
    } else if (head_dim == 256) {  // T-SA-Fuser: up to 64 tokens, block-causal
      return fp16 ? launch_attention_mma<256, 64, 8, true>(a, stream) : launch_attention_mma<256, 64, 8, false>(a, stream);
    }
  }
  if (head_dim == 256)
    return in_f32 ? launch_attention<float, 256>(a, stream)
                  : (fp16 ? launch_attention<__half, 256>(a, stream) : launch_attention<bf16, 256>(a, stream));
  if (head_dim == 512)
    return in_f32 ? launch_attention<float, 512>(a, stream)
                  : (fp16 ? launch_attention<__half, 512>(a, stream) : launch_attention<bf16, 512>(a, stream));
  return fail(AFFT_ERR_INVALID, "attention: head_dim must be 256 or 512");
}

extern "C" int afft_attention(const afft_attention_desc* d, void* stream) {
  if (d == nullptr) return fail(AFFT_ERR_INVALID, "attention: null descriptor");
  AttentionArgs a;
  a.q = d->q;
  a.k = d->k;
  a.v = d->v;
  a.ldq = d->ldq;
  a.ldk = d->ldk;
  a.ldv = d->ldv;
  a.n_seq = d->n_seq;
  a.L = d->L;
  a.H = d->H;
  a.scale = d->scale;
  a.mask = d->mask;
  a.T = d->T > 0 ? d->T : 1;
  a.out_hi = static_cast<bf16*>(d->out_hi);
  a.out_lo = static_cast<bf16*>(d->out_lo);
  a.ldo = d->ldo;
  a.probs = d->probs;
  a.p_outer = d->p_outer;
  a.p_inner_stride = d->p_inner_stride;
  a.p_inner = d->p_inner > 0 ? d->p_inner : 1;
  a.t_end = nullptr;
  a.drop = d->drop_mask;
  return run_attention(a, d->head_dim, d->in_dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int afft_marginalize_topk(const float* logits, int64_t ld, int32_t B, int32_t A, const int32_t* verb_of,
                                     const int32_t* noun_of, int32_t n_verb, int32_t n_noun, float* probs, float* verb,
                                     float* noun, int32_t* topk, int32_t K, void* stream) {
  if (logits == nullptr || verb_of == nullptr || noun_of == nullptr || verb == nullptr || noun == nullptr)
    return fail(AFFT_ERR_INVALID, "marginalize: null pointer");
  if (B <= 0 || A <= 0 || n_verb <= 0 || n_noun <= 0 || ld < A) return fail(AFFT_ERR_INVALID, "marginalize: bad sizes");
  if (n_verb + n_noun > 3072) return fail(AFFT_ERR_INVALID, "marginalize: too many verb + noun classes (max 3072: 12 B of shared memory each)");
  if (topk != nullptr && (K < 1 || K > 16 || K > n_verb || K > n_noun || K > A))
    return fail(AFFT_ERR_INVALID, "marginalize: K must be in [1, 16] and <= every class count");
  MarginalizeArgs a;
  a.logits = logits;
  a.ld = ld;
  a.B = B;
  a.A = A;
  a.verb_of = verb_of;
  a.noun_of = noun_of;
  a.n_verb = n_verb;
  a.n_noun = n_noun;
  a.probs = probs;
  a.verb = verb;
  a.noun = noun;
  a.topk = topk;
  a.K = K;
  const size_t smem = static_cast<size_t>(n_verb + n_noun) * (8 + 4) + 32 * 8 + 16 * 4;
  marginalize_topk_kernel<<<B, 256, smem, static_cast<cudaStream_t>(stream)>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail("marginalize launch", e);
  return AFFT_OK;
}

extern "C" int afft_score_fusion(const float* attn_logits, int64_t ld_a, int32_t n_mod, const float* const* logits,
                                 int64_t ld_l, int32_t rows, int32_t C, float* attn, float* out, int64_t ld_o, void* stream) {
  if (attn_logits == nullptr) return fail(AFFT_ERR_INVALID, "score_fusion: null attn_logits");
  if (n_mod < 1 || n_mod > 8) return fail(AFFT_ERR_INVALID, "score_fusion: n_mod must be in [1, 8]");
  if (rows < 1 || ld_a < n_mod) return fail(AFFT_ERR_INVALID, "score_fusion: rows must be >= 1 and ld_a >= n_mod");
  const bool softmax_only = (out == nullptr);  // only the modality attention is wanted (MATT.forward)
  if (softmax_only && attn == nullptr) return fail(AFFT_ERR_INVALID, "score_fusion: no output");
  if (softmax_only) C = 0;
  ScoreFusionArgs a;
  a.attn_logits = attn_logits;
  a.ld_a = ld_a;
  a.M = n_mod;
  for (int i = 0; i < 8; ++i) a.logits[i] = nullptr;
  if (!softmax_only) {
    const int64_t c4 = (static_cast<int64_t>(C) + 3) / 4 * 4;
    if (logits == nullptr || C < 1) return fail(AFFT_ERR_INVALID, "score_fusion: null logits / bad class count");
    if (ld_l < c4 || ld_o < c4 || ld_l % 4 != 0 || ld_o % 4 != 0)
      return fail(AFFT_ERR_INVALID, "score_fusion: logits pitches must be multiples of 4 floats and >= ceil4(C)");
    for (int i = 0; i < n_mod; ++i) {
      a.logits[i] = logits[i];
      if (a.logits[i] == nullptr || (reinterpret_cast<uintptr_t>(a.logits[i]) & 15) != 0)
        return fail(AFFT_ERR_INVALID, "score_fusion: logits pointers must be non-null and 16-byte aligned");
    }
    if ((reinterpret_cast<uintptr_t>(out) & 15) != 0) return fail(AFFT_ERR_INVALID, "score_fusion: out must be 16-byte aligned");
  }
  a.ld_l = ld_l;
  a.rows = rows;
  a.C = C;
  a.attn = attn;
  a.out = out;
  a.ld_o = ld_o;
  const int quads = (C + 3) / 4;
  dim3 grid(rows, std::min(std::max(1, (quads + 255) / 256), 65535));
  score_fusion_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail("score_fusion launch", e);
  return AFFT_OK;
}

// ------------------------------------------------------------------------------------------------
// training-step operators
// ------------------------------------------------------------------------------------------------
static int launch_check(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(what, e);
  return AFFT_OK;
}

extern "C" int afft_transpose_bf16(const void* src, int64_t lds, int32_t rows, int32_t cols, void* dst, int64_t ldd,
                                   void* stream) {
  if (src == nullptr || dst == nullptr || rows <= 0 || cols <= 0) return fail(AFFT_ERR_INVALID, "transpose: bad argument");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_bf16_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(src), lds, rows, cols,
                                                                              static_cast<bf16*>(dst), ldd);
  return launch_check("transpose launch");
}

extern "C" int afft_convert_dual(const float* src, int64_t lds, int32_t rows, int32_t cols, void* hi, int64_t ldh, void* tr,
                                 int64_t ldt, float* colsum, void* stream) {
  if (src == nullptr || rows <= 0 || cols <= 0 || (hi == nullptr && tr == nullptr && colsum == nullptr))
    return fail(AFFT_ERR_INVALID, "convert_dual: bad argument");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  convert_dual_kernel<0><<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      src, lds, rows, cols, static_cast<bf16*>(hi), ldh, static_cast<bf16*>(tr), ldt, colsum, nullptr, 0, 0);
  return launch_check("convert_dual launch");
}

extern "C" int afft_convert_dual_gelu(const float* src, int64_t lds, int32_t rows, int32_t cols, void* hi, int64_t ldh,
                                      void* tr, int64_t ldt, float* colsum, const float* d_act, int64_t ldd, int32_t kind,
                                      void* stream) {
  if (src == nullptr || rows <= 0 || cols <= 0 || (hi == nullptr && tr == nullptr && colsum == nullptr) ||
      (kind != 1 && kind != 2))
    return fail(AFFT_ERR_INVALID, "convert_dual_gelu: bad argument");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d_act == nullptr)
    convert_dual_kernel<1><<<grid, block, 0, st>>>(src, lds, rows, cols, static_cast<bf16*>(hi), ldh, static_cast<bf16*>(tr),
                                                  ldt, colsum, nullptr, 0, kind);
  else
    convert_dual_kernel<2><<<grid, block, 0, st>>>(src, lds, rows, cols, static_cast<bf16*>(hi), ldh, static_cast<bf16*>(tr),
                                                  ldt, colsum, d_act, ldd, kind);
  return launch_check("convert_dual_gelu launch");
}

extern "C" int afft_layernorm_bwd(const float* x, int64_t ldx, const float* gamma, float eps, const float* dy, int64_t lddy,
                                  int32_t rows, int32_t dim, float* dx, int64_t lddx, float* dgamma, float* dbeta,
                                  void* stream) {
  if (x == nullptr || dy == nullptr || dx == nullptr || rows <= 0) return fail(AFFT_ERR_INVALID, "layernorm_bwd: bad argument");
  if ((dgamma == nullptr) != (dbeta == nullptr)) return fail(AFFT_ERR_INVALID, "layernorm_bwd: dgamma and dbeta go together");
  LayerNormBwdArgs a{x, ldx, gamma, eps, dy, lddy, rows, dim, dx, lddx, dgamma, dbeta};
  const int blocks = std::min((rows + 7) / 8, 148 * 2);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (dim / 128) {
    case 4: layernorm_bwd_kernel<4><<<blocks, 256, 0, st>>>(a); break;
    case 8: layernorm_bwd_kernel<8><<<blocks, 256, 0, st>>>(a); break;
    case 16: layernorm_bwd_kernel<16><<<blocks, 256, 0, st>>>(a); break;
    default: return fail(AFFT_ERR_INVALID, "layernorm_bwd: dim must be 512, 1024 or 2048");
  }
  if (dim % 128 != 0) return fail(AFFT_ERR_INVALID, "layernorm_bwd: dim must be a multiple of 128");
  return launch_check("layernorm_bwd launch");
}

extern "C" int afft_gelu_fwd(const float* x, float* y, int64_t n, int32_t kind, void* stream) {
  if (x == nullptr || y == nullptr || n <= 0 || (kind != 1 && kind != 2)) return fail(AFFT_ERR_INVALID, "gelu_fwd: bad argument");
  const int blocks = static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 16));
  gelu_fwd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, n, kind);
  return launch_check("gelu_fwd launch");
}

extern "C" int afft_gelu_bwd(const float* x, const float* dy, float* dx, int64_t n, int32_t kind, void* stream) {
  if (x == nullptr || dy == nullptr || dx == nullptr || n <= 0 || (kind != 1 && kind != 2))
    return fail(AFFT_ERR_INVALID, "gelu_bwd: bad argument");
  const int blocks = static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 16));
  gelu_bwd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, dy, dx, n, kind);
  return launch_check("gelu_bwd launch");
}

extern "C" int afft_colsum(const float* x, int64_t ld, int32_t rows, int32_t cols, float* out, void* stream) {
  if (x == nullptr || out == nullptr || rows <= 0 || cols <= 0) return fail(AFFT_ERR_INVALID, "colsum: bad argument");
  dim3 grid((cols + 31) / 32, std::max(1, std::min((rows + 63) / 64, 64)));
  colsum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ld, rows, cols, out);
  return launch_check("colsum launch");
}

extern "C" int afft_sgd_nesterov(float* p, const float* g, float* m, void* p16, int64_t n, float lr, float momentum,
                                 float weight_decay, int32_t nesterov, void* stream) {
  if (p == nullptr || g == nullptr || m == nullptr || n <= 0) return fail(AFFT_ERR_INVALID, "sgd: bad argument");
  if (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m)) & 15u) != 0 ||
      (reinterpret_cast<uintptr_t>(p16) & 7u) != 0)
    return fail(AFFT_ERR_INVALID, "sgd: buffers must be 16-byte aligned (8-byte for the bf16 image)");
  const int blocks = static_cast<int>(std::min<long long>((n / 4 + 255) / 256 + 1, 148 * 16));
  sgd_nesterov_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(p, g, m, static_cast<bf16*>(p16), n, lr, momentum,
                                                                            weight_decay, nesterov);
  return launch_check("sgd launch");
}

extern "C" int afft_attention_bwd(const float* qkv, int64_t ld, const float* probs, const float* d_out, int64_t ldo,
                                  float* dqkv, int32_t n_seq, int32_t L, int32_t H, int32_t head_dim, float scale,
                                  const float* drop_mask, void* stream) {
  if (qkv == nullptr || probs == nullptr || d_out == nullptr || dqkv == nullptr) return fail(AFFT_ERR_INVALID, "attention_bwd: null pointer");
  if (L < 1 || L > 64 || n_seq <= 0 || H <= 0 || head_dim <= 0) return fail(AFFT_ERR_INVALID, "attention_bwd: bad sizes");
  if (head_dim % 4 != 0 || ld % 4 != 0 || ldo % 4 != 0 || ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(d_out) |
                                                             reinterpret_cast<uintptr_t>(dqkv)) & 15u) != 0)
    return fail(AFFT_ERR_INVALID, "attention_bwd: head_dim and pitches must be multiples of 4 floats, pointers 16-byte aligned");
  const size_t smem = (static_cast<size_t>(4) * L * head_dim + 2 * L * L) * sizeof(float);
  if (smem > 227 * 1024) return fail(AFFT_ERR_INVALID, "attention_bwd: sequence too long for shared memory");
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return cuda_fail("attention_bwd smem attribute", e);
  }
  AttentionBwdArgs a{qkv, ld, probs, d_out, ldo, dqkv, n_seq, L, H, head_dim, scale, drop_mask};
  attention_bwd_kernel<<<n_seq * H, 256, smem, static_cast<cudaStream_t>(stream)>>>(a);
  return launch_check("attention_bwd launch");
}

// ================================================================================================
// model-level
// ================================================================================================
namespace {

struct Tensor {
  // exactly one representation is populated
  bf16* hi = nullptr;  // packed GEMM weight [N, K]
  bf16* lo = nullptr;
  float* f32 = nullptr;  // vectors / tables
  long long rows = 0, cols = 0;  // packed: N, K.  f32: flattened [rows, cols]
};

enum class Pack { Gemm, GemmT, F32 };

struct Expect {
  Pack pack;
  long long numel;
  long long d0, d1;  // expected 2-D view (rows, cols) of the source tensor
  long long min_rows = 0;  // > 0: embedding table, any row count >= min_rows is accepted
};

struct PairBuf {  // bf16 activation, hi (+ lo in strict mode)
  bf16* hi = nullptr;
  bf16* lo = nullptr;
};

}  // namespace

struct afft_handle {
  afft_config cfg;
  int num_sms = 0;
  std::string err;
  std::map<std::string, Expect> expected;
  std::map<std::string, Tensor> w;
  size_t weight_bytes = 0;

  // workspace
  char* ws = nullptr;
  size_t ws_bytes = 0;
  bool ws_owned = true;  // false: caller-owned workspace (afft_create_in)
  // true from afft_set_weight until one whole forward has been enqueued behind the packing kernels: only then may a GEMM
  // request its weight bytes ahead of griddepcontrol.wait (gemm_skinny.cuh: w_static) - with a forward's ~90 kernels in
  // between, no chain of early-launched kernels reaches back to the packing kernel
  bool weights_fresh = true;
  SplitKScratch splitk{nullptr, 0, nullptr, 0, 4};  // partial tiles + band counters of the split-K GEMM path
  int n_slots = 0;  // tokens per (b, t) in the fuser stream (CA: 1)
  float* h = nullptr;
  PairBuf y, att, f;
  void* qkv = nullptr;  // bf16, or fp32 in strict mode
  PairBuf xin[AFFT_MAX_MODS];
  float* mem[AFFT_MAX_MODS] = {nullptr};  // CA-Fuser memories (+pos emb)
  PairBuf ykv;
  float* table = nullptr;  // [T, dim] additive embedding rows for projected modalities
  PairBuf zb;
  float* g = nullptr;
  PairBuf y2, att2, f2, pfb;
  void* qkv2 = nullptr;
  // autoregressive roll-out (fp_output_len > 1)
  std::vector<void*> qkv_layer;   // per GPT-2 layer: prompt q|k|v [B*T, 3G]   (the KV cache)
  std::vector<void*> qkv_new;     // per layer: generated positions [B*(O-1), 3G]
  float* gn = nullptr;            // [B, G] residual stream of the position being generated
  float* hid = nullptr;           // [B, G] last hidden state (post ln_f), fed back as the next input
  PairBuf yn, attn_n, fn;         // [B, G], [B, G], [B, 4G]
  int launches = 0;
  int fuser_chunk = 0;
  // optional per-launch timing: slot 0 = start mark, slot 1 + i = latest exit time of launch i (device %globaltimer)
  bool profile = false;
  unsigned long long* prof_slots = nullptr;
  cudaStream_t prof_stream = nullptr;
  afft_profile prof;
};

static int hfail(afft_handle* h, int code, const std::string& msg) {
  h->err = msg;
  g_err = msg;
  return code;
}

static std::string blk_name(const char* prefix, int i, const char* suffix) {
  return std::string(prefix) + std::to_string(i) + suffix;
}

static void build_expected(afft_handle* h) {
  const afft_config& c = h->cfg;
  auto& E = h->expected;
  const long long D = c.dim, G = c.gpt_dim;
  const bool want_fuser = c.stages != AFFT_STAGE_GPT, want_head = c.stages == AFFT_STAGE_ALL, want_gpt = c.stages != AFFT_STAGE_FUSER;
  auto gemm = [&](const std::string& n, long long out, long long in) { E[n] = {Pack::Gemm, out * in, out, in, 0}; };
  auto gemm_t = [&](const std::string& n, long long in, long long out) { E[n] = {Pack::GemmT, out * in, in, out, 0}; };
  auto vec = [&](const std::string& n, long long rows, long long cols) { E[n] = {Pack::F32, rows * cols, rows, cols, 0}; };
  auto table = [&](const std::string& n, long long cols) { E[n] = {Pack::F32, 0, 0, cols, c.T}; };
  auto ln = [&](const std::string& n, long long d, bool affine) {
    if (affine) {
      vec(n + ".weight", 1, d);
      vec(n + ".bias", 1, d);
    }
  };
  for (int m = 0; want_fuser && m < c.n_mod; ++m)
    if (c.mod_dim[m] != c.dim) gemm(std::string("mapping.") + c.mod_name[m] + ".mapping.0.weight", D, c.mod_dim[m]);

  const int n_slots = h->n_slots;
  const bool has_fuser = want_fuser && c.fuser_kind != AFFT_FUSER_NONE;
  const bool affine = (c.fuser_kind == AFFT_FUSER_SA) ? (c.norm_elementwise != 0) : true;
  if (!want_fuser) {
  } else if (c.fuser_kind == AFFT_FUSER_SA) {
    vec("fuser.modal_token", c.frame_level_token ? c.T : 1, D);
    if (c.modal_encoding) vec("fuser.modality_embedding", n_slots, D);
  } else if (c.fuser_kind == AFFT_FUSER_TSA) {
    if (c.frame_level_token) vec("fuser.modal_token", c.T, D);
    if (c.modal_encoding) vec("fuser.modality_embedding", n_slots, D);
    table("fuser.position_embeddings.weight", D);
  } else if (c.fuser_kind == AFFT_FUSER_CA) {
    table("fuser.position_embeddings.weight", D);
  }
  for (int i = 0; has_fuser && i < c.fuser_depth; ++i) {
    const std::string p = blk_name("fuser.blocks.", i, ".");
    if (c.fuser_kind == AFFT_FUSER_CA) {
      ln(p + "norm_self", D, true);
      ln(p + "norm_q", D, true);
      ln(p + "norm_kv", D, true);
      ln(p + "norm_mlp", D, true);
      gemm(p + "cross_attn.w_q.weight", D, D);
      gemm(p + "cross_attn.w_k.weight", D, D);
      gemm(p + "cross_attn.w_v.weight", D, D);
      gemm(p + "cross_attn.proj.weight", D, D);
      vec(p + "cross_attn.proj.bias", 1, D);
    } else {
      ln(p + "norm1", D, affine);
      ln(p + "norm2", D, affine);
    }
    gemm(p + "attn.qkv.weight", 3 * D, D);
    gemm(p + "attn.proj.weight", D, D);
    vec(p + "attn.proj.bias", 1, D);
    gemm(p + "mlp.mlp.0.weight", 4 * D, D);
    vec(p + "mlp.mlp.0.bias", 1, 4 * D);
    gemm(p + "mlp.mlp.2.weight", D, 4 * D);
    vec(p + "mlp.mlp.2.bias", 1, D);
  }
  if (has_fuser) ln("fuser.norm", D, affine);
  if (want_head && c.dim != c.gpt_dim) {
    gemm("dim_encoder.weight", G, D);
    gemm("dim_decoder.weight", D, G);
  }
  const std::string gp = "future_predictor.gpt_model.";
  if (want_gpt) E[gp + "wpe.weight"] = {Pack::F32, 0, 0, G, c.T + c.fp_output_len - 1};
  for (int i = 0; want_gpt && i < c.gpt_layers; ++i) {
    const std::string p = gp + blk_name("h.", i, ".");
    ln(p + "ln_1", G, true);
    ln(p + "ln_2", G, true);
    gemm_t(p + "attn.c_attn.weight", G, 3 * G);
    vec(p + "attn.c_attn.bias", 1, 3 * G);
    gemm_t(p + "attn.c_proj.weight", G, G);
    vec(p + "attn.c_proj.bias", 1, G);
    gemm_t(p + "mlp.c_fc.weight", G, 4 * G);
    vec(p + "mlp.c_fc.bias", 1, 4 * G);
    gemm_t(p + "mlp.c_proj.weight", 4 * G, G);
    vec(p + "mlp.c_proj.bias", 1, G);
  }
  if (want_gpt) ln(gp + "ln_f", G, true);
  for (int k = 0; want_head && k < c.n_cls; ++k) {
    const std::string p = std::string("classifiers.") + c.cls_name[k] + ".all-fused.1.";
    gemm(p + "weight", c.cls_dim[k], D);
    vec(p + "bias", 1, c.cls_dim[k]);
  }
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// One body for afft_create (library-owned workspace), afft_create_in (caller-owned workspace) and
// afft_workspace_bytes_for (size query: `need` set, nothing allocated).
static int create_impl(const afft_config* cfg, void* ws_dev, size_t ws_dev_bytes, bool caller_ws, cudaStream_t stream,
                       size_t* need, afft_handle** out) {
  if (cfg == nullptr || (out == nullptr && need == nullptr)) return fail(AFFT_ERR_INVALID, "create: null argument");
  if (out != nullptr) *out = nullptr;
  const afft_config& c = *cfg;
  if (c.fuser_kind < 0 || c.fuser_kind > AFFT_FUSER_NONE) return fail(AFFT_ERR_INVALID, "create: unknown fuser_kind");
  if (c.precision < AFFT_PREC_BF16 || c.precision > AFFT_PREC_FP16) return fail(AFFT_ERR_INVALID, "create: unknown precision");
  const bool no_fuser = c.fuser_kind == AFFT_FUSER_NONE;
  if (c.stages != AFFT_STAGE_ALL && c.stages != AFFT_STAGE_FUSER && c.stages != AFFT_STAGE_GPT)
    return fail(AFFT_ERR_INVALID, "create: stages must be AFFT_STAGE_ALL, AFFT_STAGE_FUSER or AFFT_STAGE_GPT");
  if (c.stages == AFFT_STAGE_FUSER && no_fuser) return fail(AFFT_ERR_INVALID, "create: AFFT_STAGE_FUSER needs a fuser");
  if (c.stages == AFFT_STAGE_GPT && (!no_fuser || c.n_mod != 1 || c.mod_dim[0] != c.gpt_dim || c.dim != c.gpt_dim))
    return fail(AFFT_ERR_INVALID, "create: AFFT_STAGE_GPT takes fuser_kind NONE and one modality with mod_dim = dim = gpt_dim");
  if (c.n_mod < 1 || c.n_mod > AFFT_MAX_MODS - 1) return fail(AFFT_ERR_INVALID, "create: n_mod out of range");
  if (c.n_cls < (c.stages == AFFT_STAGE_ALL ? 1 : 0) || c.n_cls > AFFT_MAX_CLS) return fail(AFFT_ERR_INVALID, "create: n_cls out of range");
  if (c.T < 1 || c.T > 64) return fail(AFFT_ERR_INVALID, "create: T must be in [1, 64]");
  if (c.max_batch < 1) return fail(AFFT_ERR_INVALID, "create: max_batch must be >= 1");
  if (no_fuser) {
    if (c.n_mod != 1 || c.mod_dim[0] != c.dim) return fail(AFFT_ERR_INVALID, "create: AFFT_FUSER_NONE takes one modality of width dim");
    if (c.dim % 8 != 0 || c.gpt_dim % 128 != 0) return fail(AFFT_ERR_INVALID, "create: dim must be a multiple of 8 and gpt_dim of 128");
  } else if (c.dim % 128 != 0 || c.gpt_dim % 128 != 0) {
    return fail(AFFT_ERR_INVALID, "create: dim and gpt_dim must be multiples of 128");
  }
  if (c.gpt_heads < 1 || c.gpt_dim % c.gpt_heads != 0 || (!no_fuser && (c.fuser_heads < 1 || c.dim % c.fuser_heads != 0)))
    return fail(AFFT_ERR_INVALID, "create: heads must divide dims");
  const int hd1 = no_fuser ? 256 : c.dim / c.fuser_heads, hd2 = c.gpt_dim / c.gpt_heads;
  if ((hd1 != 256 && hd1 != 512) || (hd2 != 256 && hd2 != 512))
    return fail(AFFT_ERR_INVALID, "create: head_dim must be 256 or 512");
  for (int m = 0; m < c.n_mod; ++m)
    if (c.mod_dim[m] < 8 || c.mod_dim[m] % 8 != 0) return fail(AFFT_ERR_INVALID, "create: modality dims must be multiples of 8");
  if (c.fuser_kind == AFFT_FUSER_CA && c.n_mod < 2) return fail(AFFT_ERR_INVALID, "create: CA-Fuser needs >= 2 modalities");
  if (c.fp_output_len < 1 || c.T + c.fp_output_len - 1 > 1024)
    return fail(AFFT_ERR_INVALID, "create: fp_output_len must be >= 1 and T + fp_output_len - 1 <= 1024 (GPT-2 positions)");

  cudaError_t e = cudaSuccess;
  int sms = 148;
  if (need == nullptr) {  // the size query is host arithmetic only
    e = cudaSetDevice(c.device);
    if (e != cudaSuccess) return cuda_fail("cudaSetDevice", e);
    int rc = device_sm_count(&sms);
    if (rc != AFFT_OK) return rc;
  }

  afft_handle* h = new afft_handle();
  h->cfg = c;
  h->num_sms = sms;
  switch (c.fuser_kind) {
    case AFFT_FUSER_SA: h->n_slots = c.n_mod + 1; break;
    case AFFT_FUSER_SA_NOTOKEN: h->n_slots = c.n_mod; break;
    case AFFT_FUSER_TSA: h->n_slots = c.n_mod + (c.frame_level_token ? 1 : 0); break;
    default: h->n_slots = 1; break;
  }
  if (c.fuser_kind == AFFT_FUSER_TSA && h->n_slots * c.T > 64) {
    delete h;
    return fail(AFFT_ERR_INVALID, "create: T-SA-Fuser sequence (slots * T) must be <= 64");
  }
  if (c.fuser_kind == AFFT_FUSER_CA && c.fuser_depth != c.n_mod - 1) h->cfg.fuser_depth = c.n_mod - 1;
  build_expected(h);
  if (const char* env = getenv("AFFT_FUSER_CHUNK")) h->fuser_chunk = atoi(env);

  // ---- workspace ----
  const bool strict = c.precision == AFFT_PREC_BF16X3;
  const size_t B = c.max_batch, T = c.T, D = c.dim, G = c.gpt_dim;
  const size_t OL = c.fp_output_len;
  const size_t R2 = B * T, R1 = R2 * h->n_slots, RP = B * (T + OL);
  struct Req {
    void** dst;
    size_t bytes;
  };
  std::vector<Req> reqs;
  auto want = [&](void** p, size_t bytes) { reqs.push_back({p, align_up(bytes, 256)}); };
  auto want_pair = [&](PairBuf& pb, size_t elems) {
    want(reinterpret_cast<void**>(&pb.hi), elems * 2);
    if (strict) want(reinterpret_cast<void**>(&pb.lo), elems * 2);
  };
  if (!no_fuser) {
    want(reinterpret_cast<void**>(&h->h), R1 * D * 4);
    want_pair(h->y, R1 * D);
    want_pair(h->att, R1 * D);
    want_pair(h->f, R1 * 4 * D);
    want(&h->qkv, R1 * 3 * D * (strict ? 4 : 2));
  }
  for (int m = 0; m < c.n_mod; ++m)
    if (c.mod_dim[m] != c.dim) want_pair(h->xin[m], R2 * c.mod_dim[m]);
  if (c.fuser_kind == AFFT_FUSER_CA) {
    for (int m = 1; m < c.n_mod; ++m) want(reinterpret_cast<void**>(&h->mem[m]), R2 * D * 4);
    want_pair(h->ykv, R2 * D);
  }
  const bool want_gpt = c.stages != AFFT_STAGE_FUSER;
  want(reinterpret_cast<void**>(&h->table), T * D * 4);
  want_pair(h->zb, R2 * D);
  if (want_gpt) {
    want(reinterpret_cast<void**>(&h->g), R2 * G * 4);
    want_pair(h->y2, R2 * G);
    want_pair(h->att2, R2 * G);
    want_pair(h->f2, R2 * 4 * G);
    want(&h->qkv2, R2 * 3 * G * (strict ? 4 : 2));
  }
  if (c.stages == AFFT_STAGE_ALL) want_pair(h->pfb, RP * D);
  want(reinterpret_cast<void**>(&h->splitk.partials), kSplitKPartialFloats * 4);
  want(reinterpret_cast<void**>(&h->splitk.counters), kSplitKCounters * 4);
  if (OL > 1 && want_gpt) {
    h->qkv_layer.assign(c.gpt_layers, nullptr);
    h->qkv_new.assign(c.gpt_layers, nullptr);
    for (int l = 0; l < c.gpt_layers; ++l) {
      want(&h->qkv_layer[l], R2 * 3 * G * (strict ? 4 : 2));
      want(&h->qkv_new[l], B * (OL - 1) * 3 * G * (strict ? 4 : 2));
    }
    want(reinterpret_cast<void**>(&h->gn), B * G * 4);
    want(reinterpret_cast<void**>(&h->hid), B * G * 4);
    want_pair(h->yn, B * G);
    want_pair(h->attn_n, B * G);
    want_pair(h->fn, B * 4 * G);
  }
  size_t total = 0;
  for (auto& r : reqs) total += r.bytes;
  if (need != nullptr) {  // size query only
    *need = total;
    delete h;
    return AFFT_OK;
  }
  if (caller_ws) {
    if (ws_dev == nullptr || ws_dev_bytes < total || (reinterpret_cast<uintptr_t>(ws_dev) & 255u) != 0) {
      delete h;
      return fail(AFFT_ERR_INVALID, "create_in: workspace must be 256-byte aligned device memory of at least afft_workspace_bytes_for() bytes");
    }
    h->ws = static_cast<char*>(ws_dev);
    h->ws_owned = false;
  } else {
    e = cudaMalloc(reinterpret_cast<void**>(&h->ws), total);
    if (e != cudaSuccess) {
      delete h;
      return cuda_fail("cudaMalloc(workspace)", e);
    }
  }
  h->ws_bytes = total;
  size_t off = 0;
  for (auto& r : reqs) {
    *r.dst = h->ws + off;
    off += r.bytes;
  }
  h->splitk.partial_floats = kSplitKPartialFloats;
  h->splitk.n_counters = kSplitKCounters;
  // once: the kernels leave the counters at zero.  Caller-owned workspace: enqueued on the caller's stream, no synchronisation
  e = caller_ws ? cudaMemsetAsync(h->splitk.counters, 0, kSplitKCounters * 4, stream) : cudaMemset(h->splitk.counters, 0, kSplitKCounters * 4);
  if (e == cudaSuccess && !caller_ws) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    if (h->ws_owned) cudaFree(h->ws);
    delete h;
    return cuda_fail("cudaMemset(split-K counters)", e);
  }
  *out = h;
  return AFFT_OK;
}

extern "C" int afft_create(const afft_config* cfg, afft_handle** out) {
  if (out == nullptr) return fail(AFFT_ERR_INVALID, "create: null argument");
  return create_impl(cfg, nullptr, 0, false, nullptr, nullptr, out);
}

extern "C" int afft_workspace_bytes_for(const afft_config* cfg, size_t* bytes) {
  if (bytes == nullptr) return fail(AFFT_ERR_INVALID, "workspace_bytes_for: null argument");
  return create_impl(cfg, nullptr, 0, false, nullptr, bytes, nullptr);
}

extern "C" int afft_create_in(const afft_config* cfg, void* workspace_dev, size_t workspace_bytes, void* stream, afft_handle** out) {
  if (out == nullptr) return fail(AFFT_ERR_INVALID, "create_in: null argument");
  return create_impl(cfg, workspace_dev, workspace_bytes, true, static_cast<cudaStream_t>(stream), nullptr, out);
}

extern "C" void afft_destroy(afft_handle* h) {
  if (h == nullptr) return;
  cudaSetDevice(h->cfg.device);
  for (auto& kv : h->w) {
    if (kv.second.hi) cudaFree(kv.second.hi);
    if (kv.second.lo) cudaFree(kv.second.lo);
    if (kv.second.f32) cudaFree(kv.second.f32);
  }
  if (h->ws && h->ws_owned) cudaFree(h->ws);
  if (h->prof_slots) cudaFree(h->prof_slots);
  delete h;
}

extern "C" const char* afft_handle_error(const afft_handle* h) { return h ? h->err.c_str() : "null handle"; }
extern "C" size_t afft_workspace_bytes(const afft_handle* h) { return h ? h->ws_bytes : 0; }
extern "C" size_t afft_weight_bytes(const afft_handle* h) { return h ? h->weight_bytes : 0; }
extern "C" int afft_last_launch_count(const afft_handle* h) { return h ? h->launches : 0; }

extern "C" int afft_set_weight(afft_handle* h, const char* name, const float* src, int32_t ndim, const int64_t* shape,
                               void* stream_) {
  if (h == nullptr) return fail(AFFT_ERR_INVALID, "set_weight: null handle");
  if (name == nullptr || src == nullptr || shape == nullptr || ndim < 1 || ndim > 3)
    return hfail(h, AFFT_ERR_INVALID, "set_weight: bad argument");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  h->weights_fresh = true;
  auto it = h->expected.find(name);
  if (it == h->expected.end()) return hfail(h, AFFT_ERR_INVALID, std::string("set_weight: unexpected tensor name '") + name + "'");
  const Expect& ex = it->second;
  long long numel = 1;
  for (int i = 0; i < ndim; ++i) numel *= shape[i];
  if (ex.min_rows > 0) {
    if (numel % ex.d1 != 0 || numel / ex.d1 < ex.min_rows)
      return hfail(h, AFFT_ERR_INVALID, std::string("set_weight: table '") + name + "' has fewer rows than T");
  } else if (numel != ex.numel) {
    return hfail(h, AFFT_ERR_INVALID, std::string("set_weight: size mismatch for '") + name + "'");
  }
  if (ex.pack != Pack::F32) {
    if (ndim != 2 || shape[0] != ex.d0 || shape[1] != ex.d1)
      return hfail(h, AFFT_ERR_INVALID, std::string("set_weight: shape mismatch for '") + name + "'");
  }
  cudaError_t e = cudaSetDevice(h->cfg.device);
  if (e != cudaSuccess) return hfail(h, AFFT_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
  Tensor& t = h->w[name];
  const bool strict = h->cfg.precision == AFFT_PREC_BF16X3;
  const int fp16 = h->cfg.precision == AFFT_PREC_FP16 ? 1 : 0;
  if (ex.pack == Pack::F32) {
    if (t.f32 == nullptr) {
      e = cudaMalloc(reinterpret_cast<void**>(&t.f32), align_up(numel * 4, 256));
      if (e != cudaSuccess) return hfail(h, AFFT_ERR_CUDA, std::string("cudaMalloc(weight): ") + cudaGetErrorString(e));
      h->weight_bytes += numel * 4;
    }
    t.rows = numel / ex.d1;
    t.cols = ex.d1;
    e = cudaMemcpyAsync(t.f32, src, numel * 4, cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) return hfail(h, AFFT_ERR_CUDA, std::string("cudaMemcpyAsync(weight): ") + cudaGetErrorString(e));
    return AFFT_OK;
  }
  // GEMM weight -> bf16 [N, K], K contiguous
  const long long N = (ex.pack == Pack::Gemm) ? ex.d0 : ex.d1;
  const long long K = (ex.pack == Pack::Gemm) ? ex.d1 : ex.d0;
  if (t.hi == nullptr) {
    e = cudaMalloc(reinterpret_cast<void**>(&t.hi), align_up(N * K * 2, 256));
    if (e != cudaSuccess) return hfail(h, AFFT_ERR_CUDA, std::string("cudaMalloc(weight): ") + cudaGetErrorString(e));
    h->weight_bytes += N * K * 2;
    if (strict) {
      e = cudaMalloc(reinterpret_cast<void**>(&t.lo), align_up(N * K * 2, 256));
      if (e != cudaSuccess) return hfail(h, AFFT_ERR_CUDA, std::string("cudaMalloc(weight): ") + cudaGetErrorString(e));
      h->weight_bytes += N * K * 2;
    }
  }
  t.rows = N;
  t.cols = K;
  int rc = (ex.pack == Pack::Gemm)
               ? run_convert(src, K, static_cast<int>(N), static_cast<int>(K), t.hi, t.lo, K, 0, stream, fp16)
               : run_convert(src, N, static_cast<int>(K), static_cast<int>(N), t.hi, t.lo, K, 1, stream, fp16);
  if (rc != AFFT_OK) h->err = g_err;
  return rc;
}

extern "C" int afft_missing_weights(const afft_handle* h, char* buf, size_t buf_len) {
  if (h == nullptr) return -1;
  int n = 0;
  std::string names;
  for (auto& kv : h->expected) {
    if (h->w.find(kv.first) == h->w.end()) {
      ++n;
      names += kv.first;
      names += '\n';
    }
  }
  if (buf != nullptr && buf_len > 0) {
    strncpy(buf, names.c_str(), buf_len - 1);
    buf[buf_len - 1] = '\0';
  }
  return n;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
namespace {

struct Fwd {
  afft_handle* h;
  cudaStream_t stream;
  bool strict;  // AFFT_PREC_BF16X3
  bool fp16;    // AFFT_PREC_FP16
  int rc = AFFT_OK;

  const Tensor* W(const std::string& n) {
    auto it = h->w.find(n);
    return it == h->w.end() ? nullptr : &it->second;
  }
  const float* V(const std::string& n) {
    const Tensor* t = W(n);
    return t ? t->f32 : nullptr;
  }
  bool ok() const { return rc == AFFT_OK; }
  void check(int r) {
    if (rc == AFFT_OK && r != AFFT_OK) {
      rc = r;
      h->err = g_err;
    }
    if (r == AFFT_OK) ++h->launches;
  }
  // profiling slot of the next launch (nullptr when profiling is off): the kernel itself records its exit time
  unsigned long long* prof_slot(int cat, int M = 0, int N = 0, int K = 0) {
    if (!h->profile || h->prof.n >= AFFT_MAX_PROFILE_RECS) return nullptr;
    const int i = h->prof.n++;
    h->prof.recs[i] = {cat, M, N, K, 0.f};
    return h->prof_slots + 1 + i;
  }

  // out = epilogue(A . W^T)
  void gemm(const PairBuf& A, long long lda, int M, const std::string& wname, const float* bias, int act,
            const float* res, long long ld_res, int res_mod, float* out_f32, long long ld_f32, const PairBuf* out_b,
            long long ld_bf16, int row_group = 0, int row_stride = 0, int row_off = 0) {
    if (!ok()) return;
    const Tensor* w = W(wname);
    afft_gemm_desc d;
    memset(&d, 0, sizeof(d));
    d.a_hi = A.hi;
    d.a_lo = A.lo;
    d.lda = lda;
    d.w_hi = w->hi;
    d.w_lo = w->lo;
    d.ldw = w->cols;
    d.M = M;
    d.N = static_cast<int>(w->rows);
    d.K = static_cast<int>(w->cols);
    d.precision = h->cfg.precision;
    d.bias = bias;
    d.res = res;
    d.ld_res = ld_res;
    d.res_mod = res_mod;
    d.act = act;
    d.out_f32 = out_f32;
    d.ld_f32 = ld_f32;
    if (out_b != nullptr) {
      d.out_hi = out_b->hi;
      d.out_lo = out_b->lo;
    }
    d.ld_bf16 = ld_bf16;
    d.row_group = row_group;
    d.row_stride = row_stride;
    d.row_off = row_off;
    check(run_gemm(d, h->num_sms, stream, &h->splitk, prof_slot(AFFT_CAT_GEMM, d.M, d.N, d.K), !h->weights_fresh));
  }

  void layernorm(const float* x, long long ldx, int rows, int dim, const std::string& name, float eps, const PairBuf* yb,
                 float* y_f32, long long ldy, int in_group = 0, int in_stride = 0, int n_avg = 0, int avg_stride = 0,
                 int aux_mod = 0, int aux_stride = 0, float* aux_f32 = nullptr, const PairBuf* aux_b = nullptr,
                 long long ld_aux = 0, int aux_rem = 0, int out_group = 0, int out_stride = 0, int out_off = 0) {
    if (!ok()) return;
    LayerNormArgs a;
    a.out_group = out_group;
    a.out_stride = out_stride;
    a.out_off = out_off;
    a.x = x;
    a.ldx = ldx;
    a.in_group = in_group;
    a.in_stride = in_stride;
    a.n_avg = n_avg;
    a.avg_stride = avg_stride;
    a.gamma = V(name + ".weight");
    a.beta = V(name + ".bias");
    a.eps = eps;
    a.rows = rows;
    a.dim = dim;
    a.y_f32 = y_f32;
    a.y_hi = yb ? yb->hi : nullptr;
    a.y_lo = yb ? yb->lo : nullptr;
    a.ldy = ldy;
    a.aux_mod = aux_mod;
    a.aux_rem = aux_rem;
    a.aux_stride = aux_stride;
    a.aux_f32 = aux_f32;
    a.aux_hi = aux_b ? aux_b->hi : nullptr;
    a.aux_lo = aux_b ? aux_b->lo : nullptr;
    a.ld_aux = ld_aux;
    a.out_fp16 = fp16 ? 1 : 0;
    a.t_end = prof_slot(AFFT_CAT_LAYERNORM);
    check(run_layernorm(a, stream));
  }

  // q/k/v live in one buffer of row pitch ld (elements) at column offsets qo/ko/vo
  void attention(const void* buf, long long ld, long long qo, long long ko, long long vo, int n_seq, int L, int H,
                 int hd, int mask, int T, const PairBuf& out, long long ldo, float* probs, long long p_outer,
                 long long p_inner_stride, int p_inner) {
    if (!ok()) return;
    const size_t es = strict ? 4 : 2;
    AttentionArgs a;
    a.q = static_cast<const char*>(buf) + qo * es;
    a.k = static_cast<const char*>(buf) + ko * es;
    a.v = static_cast<const char*>(buf) + vo * es;
    a.ldq = a.ldk = a.ldv = ld;
    a.n_seq = n_seq;
    a.L = L;
    a.H = H;
    a.scale = 1.0f / sqrtf(static_cast<float>(hd));
    a.mask = mask;
    a.T = T > 0 ? T : 1;
    a.out_hi = out.hi;
    a.out_lo = out.lo;
    a.ldo = ldo;
    a.probs = probs;
    a.p_outer = p_outer;
    a.p_inner_stride = p_inner_stride;
    a.p_inner = p_inner > 0 ? p_inner : 1;
    a.drop = nullptr;
    a.t_end = prof_slot(AFFT_CAT_ATTENTION);
    check(run_attention(a, hd, strict ? AFFT_DT_F32 : (fp16 ? AFFT_DT_FP16 : AFFT_DT_BF16), stream));
  }

  void convert(const float* src, long long lds, int rows, int cols, const PairBuf& dst, long long ldd) {
    if (!ok()) return;
    check(run_convert(src, lds, rows, cols, dst.hi, dst.lo, ldd, 0, stream, fp16 ? 1 : 0, prof_slot(AFFT_CAT_OTHER)));
  }

  void assemble(AssembleArgs a) {
    if (!ok()) return;
    const long long rows = static_cast<long long>(a.B) * a.T * a.n_slots;
    const int blocks = static_cast<int>(std::min<long long>(rows, 148 * 16));
    a.t_end = prof_slot(AFFT_CAT_OTHER);
    assemble_tokens_kernel<<<blocks, 256, 0, stream>>>(a);
    cudaError_t e = cudaGetLastError();
    check(e == cudaSuccess ? AFFT_OK : cuda_fail("assemble launch", e));
  }

  void add_row_vector(float* out, const float* in, const float* vec, int rows, int dim) {
    if (!ok()) return;
    add_row_vector_kernel<<<(rows * dim + 255) / 256, 256, 0, stream>>>(out, in, vec, rows, dim, prof_slot(AFFT_CAT_OTHER));
    cudaError_t e = cudaGetLastError();
    check(e == cudaSuccess ? AFFT_OK : cuda_fail("add_row_vector launch", e));
  }

  void decode_attention(const void* cache, const void* fresh, long long ld, int B, int T, int H, int hd, int n_new,
                        int n_new_max, const PairBuf& out) {
    if (!ok()) return;
    DecodeAttnArgs a;
    a.cache = cache;
    a.fresh = fresh;
    a.ld = ld;
    a.B = B;
    a.T = T;
    a.H = H;
    a.n_new = n_new;
    a.n_new_max = n_new_max;
    a.scale = 1.0f / sqrtf(static_cast<float>(hd));
    a.out_hi = out.hi;
    a.out_lo = out.lo;
    a.t_end = prof_slot(AFFT_CAT_ATTENTION);
    const int blocks = (B * H + 3) / 4;
    if (hd == 512) {
      if (strict) attention_decode_kernel<float, 512><<<blocks, 128, 0, stream>>>(a);
      else if (fp16) attention_decode_kernel<__half, 512><<<blocks, 128, 0, stream>>>(a);
      else attention_decode_kernel<bf16, 512><<<blocks, 128, 0, stream>>>(a);
    } else {
      if (strict) attention_decode_kernel<float, 256><<<blocks, 128, 0, stream>>>(a);
      else if (fp16) attention_decode_kernel<__half, 256><<<blocks, 128, 0, stream>>>(a);
      else attention_decode_kernel<bf16, 256><<<blocks, 128, 0, stream>>>(a);
    }
    cudaError_t e = cudaGetLastError();
    check(e == cudaSuccess ? AFFT_OK : cuda_fail("decode attention launch", e));
  }

  void embed_table(float* table, const float* pos, const float* mod, int T, int dim) {
    if (!ok()) return;
    embed_table_kernel<<<(T * dim + 255) / 256, 256, 0, stream>>>(table, pos, mod, T, dim, prof_slot(AFFT_CAT_OTHER));
    cudaError_t e = cudaGetLastError();
    check(e == cudaSuccess ? AFFT_OK : cuda_fail("embed_table launch", e));
  }
};

// qkv buffer views for the GEMM epilogue: bf16 output in fast mode, fp32 in strict mode.
struct QkvOut {
  float* f32;
  PairBuf b;
  const PairBuf* bp;
};
static QkvOut qkv_out(void* buf, bool strict, long long col_off) {
  QkvOut o;
  o.f32 = strict ? static_cast<float*>(buf) + col_off : nullptr;
  o.b.hi = strict ? nullptr : static_cast<bf16*>(buf) + col_off;
  o.b.lo = nullptr;
  o.bp = strict ? nullptr : &o.b;
  return o;
}

}  // namespace

// Self-attention transformer block shared by SA / T-SA / CA fusers
// (reference models/transformerblock.py:118-135).  rows x D stream in h->h.
static void fuser_block(Fwd& F, const std::string& p, const char* n1, const char* n2, int rows, int n_seq, int L,
                        int mask, int T, float* probs, long long p_outer, long long p_inner_stride, int p_inner) {
  afft_handle* h = F.h;
  const afft_config& c = h->cfg;
  const int D = c.dim, H = c.fuser_heads, hd = D / H;
  F.layernorm(h->h, D, rows, D, p + n1, 1e-6f, &h->y, nullptr, D);
  QkvOut q = qkv_out(h->qkv, F.strict, 0);
  F.gemm(h->y, D, rows, p + "attn.qkv.weight", nullptr, ACT_NONE, nullptr, 0, 0, q.f32, 3 * D, q.bp, 3 * D);
  F.attention(h->qkv, 3 * D, 0, D, 2 * D, n_seq, L, H, hd, mask, T, h->att, D, probs, p_outer, p_inner_stride, p_inner);
  F.gemm(h->att, D, rows, p + "attn.proj.weight", F.V(p + "attn.proj.bias"), ACT_NONE, h->h, D, 0, h->h, D, nullptr, 0);
  (void)n2;
}

static void fuser_mlp(Fwd& F, const std::string& p, const char* norm, int rows) {
  afft_handle* h = F.h;
  const int D = h->cfg.dim;
  F.layernorm(h->h, D, rows, D, p + norm, 1e-6f, &h->y, nullptr, D);
  F.gemm(h->y, D, rows, p + "mlp.mlp.0.weight", F.V(p + "mlp.mlp.0.bias"), ACT_GELU_ERF, nullptr, 0, 0, nullptr, 0,
         &h->f, 4 * D);
  F.gemm(h->f, 4 * D, rows, p + "mlp.mlp.2.weight", F.V(p + "mlp.mlp.2.bias"), ACT_NONE, h->h, D, 0, h->h, D, nullptr,
         0);
}

// Fuser over clips [b0, b0 + nb): writes z into orig_past / zb / slot 0 of past_futures & pfb.
static void run_fuser(Fwd& F, const afft_io& io, int b0, int nb) {
  afft_handle* h = F.h;
  const afft_config& c = h->cfg;
  const int T = c.T, D = c.dim, n = h->n_slots, H = c.fuser_heads, depth = c.fuser_depth;
  const int R2 = nb * T;
  const int S = T + c.fp_output_len;                                  // slots per clip in past_futures / logits
  const long long zoff = static_cast<long long>(b0) * T * D;          // into orig_past / zb
  const long long poff = static_cast<long long>(b0) * S * D;          // into past_futures / pfb
  const bool head = c.stages == AFFT_STAGE_ALL;  // AFFT_STAGE_FUSER: only z = orig_past (+ attention) is produced
  PairBuf zb = {h->zb.hi + zoff, h->zb.lo ? h->zb.lo + zoff : nullptr};
  PairBuf pfb_v = {head ? h->pfb.hi + poff : nullptr, (head && h->pfb.lo) ? h->pfb.lo + poff : nullptr};
  const PairBuf* pfbp = head ? &pfb_v : nullptr;
  const int aux_T = head ? T : 0;  // LayerNorm aux scatter (slot 0 of past_futures) off without the head
  float* orig_past = io.orig_past + zoff;
  float* pf = head ? io.past_futures + poff : nullptr;
  const float* feat[AFFT_MAX_MODS];
  for (int m = 0; m < c.n_mod; ++m) feat[m] = io.feat[m] + static_cast<long long>(b0) * T * c.mod_dim[m];

  if (c.fuser_kind == AFFT_FUSER_NONE) {
    // no fusion (future_prediction.py:203-204): z = the modality's features.  bf16 copy for dim_encoder, fp32 copy into
    // orig_past, and frame 0 into slot 0 of past_futures (fp32 + the classifier's bf16 operand) - prepare_output :172-176.
    F.convert(feat[0], D, R2, D, zb, D);
    if (F.ok()) {
      const size_t row4 = static_cast<size_t>(D) * 4, row2 = static_cast<size_t>(D) * 2;
      cudaError_t e = cudaMemcpyAsync(orig_past, feat[0], static_cast<size_t>(R2) * row4, cudaMemcpyDeviceToDevice, F.stream);
      if (e == cudaSuccess)
        e = cudaMemcpy2DAsync(pf, S * row4, feat[0], T * row4, row4, nb, cudaMemcpyDeviceToDevice, F.stream);
      if (e == cudaSuccess)
        e = cudaMemcpy2DAsync(pfb_v.hi, S * row2, zb.hi, T * row2, row2, nb, cudaMemcpyDeviceToDevice, F.stream);
      if (e == cudaSuccess && zb.lo != nullptr)
        e = cudaMemcpy2DAsync(pfb_v.lo, S * row2, zb.lo, T * row2, row2, nb, cudaMemcpyDeviceToDevice, F.stream);
      if (e != cudaSuccess) F.check(cuda_fail("stage features", e));
    }
    return;
  }
  if (c.fuser_kind == AFFT_FUSER_CA) {
    // x = rgb + pos; mems = others + pos  (fusion.py:262-264)
    const float* pos = F.V("fuser.position_embeddings.weight");
    F.embed_table(h->table, pos, nullptr, T, D);
    for (int m = 0; m < c.n_mod; ++m) {
      float* dst = (m == 0) ? h->h : h->mem[m];
      if (c.mod_dim[m] != D) {
        F.convert(feat[m], c.mod_dim[m], R2, c.mod_dim[m], h->xin[m], c.mod_dim[m]);
        F.gemm(h->xin[m], c.mod_dim[m], R2, std::string("mapping.") + c.mod_name[m] + ".mapping.0.weight", nullptr,
               ACT_NONE, h->table, D, T, dst, D, nullptr, 0);
      } else {
        AssembleArgs a;
        memset(&a, 0, sizeof(a));
        a.h = dst;
        a.B = nb;
        a.T = T;
        a.dim = D;
        a.n_slots = 1;
        a.layout = 0;
        a.src[0] = feat[m];
        a.tok_mod = 1;
        a.pos_emb = pos;
        F.assemble(a);
      }
    }
    const int hd = D / H;
    for (int i = 0; i < depth; ++i) {
      const std::string p = blk_name("fuser.blocks.", i, ".");
      fuser_block(F, p, "norm_self", nullptr, R2, nb, T, /*causal*/ 1, T, nullptr, 0, 0, 1);
      // cross attention: q from x, k/v from memory i  (transformerblock.py:60-76,160)
      F.layernorm(h->h, D, R2, D, p + "norm_q", 1e-6f, &h->y, nullptr, D);
      F.layernorm(h->mem[i + 1], D, R2, D, p + "norm_kv", 1e-6f, &h->ykv, nullptr, D);
      QkvOut q = qkv_out(h->qkv, F.strict, 0), k = qkv_out(h->qkv, F.strict, D), v = qkv_out(h->qkv, F.strict, 2 * D);
      F.gemm(h->y, D, R2, p + "cross_attn.w_q.weight", nullptr, ACT_NONE, nullptr, 0, 0, q.f32, 3 * D, q.bp, 3 * D);
      F.gemm(h->ykv, D, R2, p + "cross_attn.w_k.weight", nullptr, ACT_NONE, nullptr, 0, 0, k.f32, 3 * D, k.bp, 3 * D);
      F.gemm(h->ykv, D, R2, p + "cross_attn.w_v.weight", nullptr, ACT_NONE, nullptr, 0, 0, v.f32, 3 * D, v.bp, 3 * D);
      F.attention(h->qkv, 3 * D, 0, D, 2 * D, nb, T, H, hd, 1, T, h->att, D, nullptr, 0, 0, 1);
      F.gemm(h->att, D, R2, p + "cross_attn.proj.weight", F.V(p + "cross_attn.proj.bias"), ACT_NONE, h->h, D, 0, h->h,
             D, nullptr, 0);
      fuser_mlp(F, p, "norm_mlp", R2);
    }
    F.layernorm(h->h, D, R2, D, "fuser.norm", 1e-6f, &zb, orig_past, D, 0, 0, 0, 0, aux_T, S, pf, pfbp, D);
    return;
  }

  const bool tsa = (c.fuser_kind == AFFT_FUSER_TSA);
  const bool has_token = (c.fuser_kind == AFFT_FUSER_SA) || (tsa && c.frame_level_token);
  const int R1 = R2 * n;
  const float* pos = tsa ? F.V("fuser.position_embeddings.weight") : nullptr;
  const float* modemb = c.modal_encoding ? F.V("fuser.modality_embedding") : nullptr;

  // 1. token assembly + projections of modalities whose width differs from dim
  AssembleArgs a;
  memset(&a, 0, sizeof(a));
  a.h = h->h;
  a.B = nb;
  a.T = T;
  a.dim = D;
  a.n_slots = n;
  a.layout = tsa ? 1 : 0;
  a.token = has_token ? F.V("fuser.modal_token") : nullptr;
  a.tok_mod = (has_token && (tsa || c.frame_level_token)) ? T : 1;
  a.pos_emb = pos;
  a.mod_emb = modemb;
  for (int m = 0; m < c.n_mod; ++m) {
    const int s = m + (has_token ? 1 : 0);
    if (c.mod_dim[m] == D) {
      a.src[s] = feat[m];
      continue;
    }
    a.src[s] = nullptr;
    F.convert(feat[m], c.mod_dim[m], R2, c.mod_dim[m], h->xin[m], c.mod_dim[m]);
    const std::string wn = std::string("mapping.") + c.mod_name[m] + ".mapping.0.weight";
    const float* res = nullptr;
    int res_mod = 0;
    if (tsa) {
      F.embed_table(h->table, pos, modemb ? modemb + static_cast<long long>(s) * D : nullptr, T, D);
      res = h->table;
      res_mod = T;
    } else if (modemb != nullptr) {
      res = modemb + static_cast<long long>(s) * D;
      res_mod = 1;
    }
    if (tsa)
      F.gemm(h->xin[m], c.mod_dim[m], R2, wn, nullptr, ACT_NONE, res, D, res_mod, h->h, D, nullptr, 0, T, n * T, s * T);
    else
      F.gemm(h->xin[m], c.mod_dim[m], R2, wn, nullptr, ACT_NONE, res, D, res_mod, h->h + static_cast<long long>(s) * D,
             static_cast<long long>(n) * D, nullptr, 0);
  }
  if (has_token) a.is_token[0] = 1;
  F.assemble(a);

  // 2. blocks
  const int L = tsa ? n * T : n;
  const int n_seq = tsa ? nb : R2;
  const int mask = tsa ? 2 : (c.cross_attn ? 3 : 0);
  for (int i = 0; i < depth; ++i) {
    const std::string p = blk_name("fuser.blocks.", i, ".");
    float* probs = nullptr;
    long long p_outer = 0, p_inner_stride = 0;
    int p_inner = 1;
    if (io.fuser_attn != nullptr) {
      const long long per_seq = static_cast<long long>(H) * L * L;
      if (tsa) {
        p_outer = depth * per_seq;
        probs = io.fuser_attn + b0 * p_outer + i * per_seq;
      } else {
        p_outer = static_cast<long long>(depth) * T * per_seq;
        p_inner = T;
        p_inner_stride = per_seq;
        probs = io.fuser_attn + b0 * p_outer + static_cast<long long>(i) * T * per_seq;
      }
    }
    fuser_block(F, p, "norm1", nullptr, R1, n_seq, L, mask, T, probs, p_outer, p_inner_stride, p_inner);
    fuser_mlp(F, p, "norm2", R1);
  }

  // 3. final norm + token selection / averaging
  if (c.fuser_kind == AFFT_FUSER_SA) {
    // token 0 of every (b, t): fusion.py:362-364
    F.layernorm(h->h, static_cast<long long>(n) * D, R2, D, "fuser.norm", 1e-6f, &zb, orig_past, D, 0, 0, 0, 0, aux_T, S,
                pf, pfbp, D);
  } else if (c.fuser_kind == AFFT_FUSER_SA_NOTOKEN) {
    // mean over the modality tokens of LN(x): fusion.py:114-116
    F.layernorm(h->h, D, R2, D, "fuser.norm", 1e-6f, &zb, orig_past, D, 1, n, n, 1, aux_T, S, pf, pfbp, D);
  } else if (c.frame_level_token) {
    // first T tokens of every clip: fusion.py:207-209
    F.layernorm(h->h, D, R2, D, "fuser.norm", 1e-6f, &zb, orig_past, D, T, n * T, 0, 0, aux_T, S, pf, pfbp, D);
  } else {
    // mean over modalities per timestep: fusion.py:211-214
    F.layernorm(h->h, D, R2, D, "fuser.norm", 1e-6f, &zb, orig_past, D, T, n * T, n, T, aux_T, S, pf, pfbp, D);
  }
}

// The GPT-2 blocks over the T prompt positions of B clips: residual stream in h->g [B*T, G] (transformers GPT2Block;
// SURVEY.md Appendix A step 7).  With a roll-out the per-layer q|k|v buffers are kept as the KV cache.
static void gpt_prompt_layers(Fwd& F, int B, float* gpt_attn) {
  afft_handle* h = F.h;
  const afft_config& c = h->cfg;
  const int T = c.T, G = c.gpt_dim, H = c.gpt_heads, hd = G / H, OL = c.fp_output_len;
  const int R2 = B * T;
  const std::string gp = "future_predictor.gpt_model.";
  for (int i = 0; i < c.gpt_layers; ++i) {
    const std::string p = gp + blk_name("h.", i, ".");
    void* qkv = (OL > 1) ? h->qkv_layer[i] : h->qkv2;  // kept per layer when it is the KV cache of a roll-out
    F.layernorm(h->g, G, R2, G, p + "ln_1", 1e-5f, &h->y2, nullptr, G);
    QkvOut q = qkv_out(qkv, F.strict, 0);
    F.gemm(h->y2, G, R2, p + "attn.c_attn.weight", F.V(p + "attn.c_attn.bias"), ACT_NONE, nullptr, 0, 0, q.f32, 3 * G,
           q.bp, 3 * G);
    const long long per_clip = static_cast<long long>(c.gpt_layers) * H * T * T;  // probabilities [B, layers, H, T, T]
    F.attention(qkv, 3 * G, 0, G, 2 * G, B, T, H, hd, 1, T, h->att2, G,
                gpt_attn != nullptr ? gpt_attn + static_cast<long long>(i) * H * T * T : nullptr, per_clip, 0, 1);
    F.gemm(h->att2, G, R2, p + "attn.c_proj.weight", F.V(p + "attn.c_proj.bias"), ACT_NONE, h->g, G, 0, h->g, G, nullptr,
           0);
    F.layernorm(h->g, G, R2, G, p + "ln_2", 1e-5f, &h->y2, nullptr, G);
    F.gemm(h->y2, G, R2, p + "mlp.c_fc.weight", F.V(p + "mlp.c_fc.bias"), ACT_GELU_TANH, nullptr, 0, 0, nullptr, 0,
           &h->f2, 4 * G);
    F.gemm(h->f2, 4 * G, R2, p + "mlp.c_proj.weight", F.V(p + "mlp.c_proj.bias"), ACT_NONE, h->g, G, 0, h->g, G, nullptr,
           0);
  }
}

// One generated position (k-th, k >= 1) of every clip through the GPT-2 blocks with the KV cache: the fed-back hidden
// state h->hid + wpe[T + k - 1] enters h->gn (future_prediction.py:395-412).
static void gpt_decode_layers(Fwd& F, int B, int k) {
  afft_handle* h = F.h;
  const afft_config& c = h->cfg;
  const int T = c.T, G = c.gpt_dim, H = c.gpt_heads, hd = G / H, OL = c.fp_output_len;
  const std::string gp = "future_predictor.gpt_model.";
  const float* wpe = F.V(gp + "wpe.weight");
  const int pos = T + k - 1;  // position id of the token being generated (future_prediction.py:398-399)
  F.add_row_vector(h->gn, h->hid, wpe + static_cast<long long>(pos) * G, B, G);
  for (int i = 0; i < c.gpt_layers; ++i) {
    const std::string p = gp + blk_name("h.", i, ".");
    F.layernorm(h->gn, G, B, G, p + "ln_1", 1e-5f, &h->yn, nullptr, G);
    QkvOut q = qkv_out(h->qkv_new[i], F.strict, 0);
    F.gemm(h->yn, G, B, p + "attn.c_attn.weight", F.V(p + "attn.c_attn.bias"), ACT_NONE, nullptr, 0, 0, q.f32, 3 * G,
           q.bp, 3 * G, 1, OL - 1, k - 1);
    F.decode_attention(h->qkv_layer[i], h->qkv_new[i], 3 * G, B, T, H, hd, k, OL - 1, h->attn_n);
    F.gemm(h->attn_n, G, B, p + "attn.c_proj.weight", F.V(p + "attn.c_proj.bias"), ACT_NONE, h->gn, G, 0, h->gn, G,
           nullptr, 0);
    F.layernorm(h->gn, G, B, G, p + "ln_2", 1e-5f, &h->yn, nullptr, G);
    F.gemm(h->yn, G, B, p + "mlp.c_fc.weight", F.V(p + "mlp.c_fc.bias"), ACT_GELU_TANH, nullptr, 0, 0, nullptr, 0, &h->fn,
           4 * G);
    F.gemm(h->fn, 4 * G, B, p + "mlp.c_proj.weight", F.V(p + "mlp.c_proj.bias"), ACT_NONE, h->gn, G, 0, h->gn, G, nullptr,
           0);
  }
}

// dim_encoder -> GPT-2 -> dim_decoder -> classifiers over all B clips
// (future_prediction.py:267-269,282-288; transformers GPT2Model).
// With fp_output_len = O > 1 the predictor is rolled out O - 1 more positions with its KV cache
// (future_prediction.py:395-412): z_hat has T + O - 1 rows per clip.
static void run_predictor(Fwd& F, const afft_io& io, int B) {
  afft_handle* h = F.h;
  const afft_config& c = h->cfg;
  const int T = c.T, D = c.dim, G = c.gpt_dim;
  const int OL = c.fp_output_len, S = T + OL;
  const int R2 = B * T;
  const std::string gp = "future_predictor.gpt_model.";
  const float* wpe = F.V(gp + "wpe.weight");
  // common_dim == fp_inter_dim: dim_encoder / dim_decoder are nn.Identity (future_prediction.py:245-255) - z enters
  // GPT-2 as it is and ln_f's output IS z_hat
  const bool identity = (D == G);
  if (identity) {
    AssembleArgs a;  // g = z + wpe[t]
    memset(&a, 0, sizeof(a));
    a.h = h->g;
    a.B = B;
    a.T = T;
    a.dim = G;
    a.n_slots = 1;
    a.layout = 0;
    a.src[0] = io.orig_past;
    a.tok_mod = 1;
    a.pos_emb = wpe;
    F.assemble(a);
  } else {
    // g = z . Wenc^T + wpe[t]
    F.gemm(h->zb, D, R2, "dim_encoder.weight", nullptr, ACT_NONE, wpe, G, T, h->g, G, nullptr, 0);
  }
  gpt_prompt_layers(F, B, io.gpt_attn);
  if (identity) {
    // z_hat[b, t] = ln_f(g)[b, t] -> slot t + 1 of past_futures (fp32) and of the classifier's 16-bit operand buffer;
    // roll-out: the last position's hidden state is also kept in h->hid
    F.layernorm(h->g, G, R2, G, gp + "ln_f", 1e-5f, &h->pfb, io.past_futures, D, 0, 0, 0, 0, OL > 1 ? T : 0, 1,
                OL > 1 ? h->hid : nullptr, nullptr, G, T - 1, T, S, 1);
  } else {
    if (OL > 1)  // also keep the last position's hidden state (fp32): it is the next input embedding
      F.layernorm(h->g, G, R2, G, gp + "ln_f", 1e-5f, &h->y2, nullptr, G, 0, 0, 0, 0, T, 1, h->hid, nullptr, G, T - 1);
    else
      F.layernorm(h->g, G, R2, G, gp + "ln_f", 1e-5f, &h->y2, nullptr, G);
    // z_hat[b, t] -> slot t + 1 of the [B, S, D] past_futures buffer (fp32 output + bf16 classifier input)
    F.gemm(h->y2, G, R2, "dim_decoder.weight", nullptr, ACT_NONE, nullptr, 0, 0, io.past_futures, D, &h->pfb, D, T, S, 1);
  }

  for (int k = 1; k < OL; ++k) {
    gpt_decode_layers(F, B, k);
    if (identity) {
      F.layernorm(h->gn, G, B, G, gp + "ln_f", 1e-5f, &h->pfb, io.past_futures, D, 0, 0, 0, 0, 1, 1, h->hid, nullptr, G, 0,
                  1, S, T + k);
    } else {
      F.layernorm(h->gn, G, B, G, gp + "ln_f", 1e-5f, &h->yn, h->hid, G);
      F.gemm(h->yn, G, B, "dim_decoder.weight", nullptr, ACT_NONE, nullptr, 0, 0, io.past_futures, D, &h->pfb, D, 1, S, T + k);
    }
  }
  for (int k = 0; k < c.n_cls; ++k) {
    const std::string p = std::string("classifiers.") + c.cls_name[k] + ".all-fused.1.";
    F.gemm(h->pfb, D, B * S, p + "weight", F.V(p + "bias"), ACT_NONE, nullptr, 0, 0, io.logits[k], io.ld_logits[k],
           nullptr, 0);
  }
}

// AFFT_STAGE_GPT: BaseFuturePredictor.forward (models/future_prediction.py:387-415 with Identity encoder / decoder):
// io.feat[0] [B, T, G] -> last hidden states of the T prompt positions (io.orig_past) and of the fp_output_len - 1
// generated ones (io.past_futures [B, O - 1, G]).
static void run_gpt_only(Fwd& F, const afft_io& io, int B) {
  afft_handle* h = F.h;
  const afft_config& c = h->cfg;
  const int T = c.T, G = c.gpt_dim, OL = c.fp_output_len;
  const int R2 = B * T;
  const std::string gp = "future_predictor.gpt_model.";
  AssembleArgs a;  // g = inputs_embeds + wpe[t]
  memset(&a, 0, sizeof(a));
  a.h = h->g;
  a.B = B;
  a.T = T;
  a.dim = G;
  a.n_slots = 1;
  a.layout = 0;
  a.src[0] = io.feat[0];
  a.tok_mod = 1;
  a.pos_emb = F.V(gp + "wpe.weight");
  F.assemble(a);
  gpt_prompt_layers(F, B, io.gpt_attn);
  if (OL > 1)
    F.layernorm(h->g, G, R2, G, gp + "ln_f", 1e-5f, nullptr, io.orig_past, G, 0, 0, 0, 0, T, 1, h->hid, nullptr, G, T - 1);
  else
    F.layernorm(h->g, G, R2, G, gp + "ln_f", 1e-5f, nullptr, io.orig_past, G);
  for (int k = 1; k < OL; ++k) {
    gpt_decode_layers(F, B, k);
    // hidden state of the generated position: fed back (h->hid) and written to io.past_futures[b, k - 1]
    F.layernorm(h->gn, G, B, G, gp + "ln_f", 1e-5f, nullptr, h->hid, G, 0, 0, 0, 0, 1, OL - 1,
                io.past_futures + static_cast<long long>(k - 1) * G, nullptr, G);
  }
}

extern "C" int afft_forward(afft_handle* h, int32_t B, const afft_io* io, void* stream_) {
  if (h == nullptr) return fail(AFFT_ERR_INVALID, "forward: null handle");
  if (io == nullptr) return hfail(h, AFFT_ERR_INVALID, "forward: null io");
  const afft_config& c = h->cfg;
  if (B < 1 || B > c.max_batch) return hfail(h, AFFT_ERR_INVALID, "forward: B out of range [1, max_batch]");
  for (int m = 0; m < c.n_mod; ++m)
    if (io->feat[m] == nullptr) return hfail(h, AFFT_ERR_INVALID, "forward: null feature pointer");
  if (io->orig_past == nullptr) return hfail(h, AFFT_ERR_INVALID, "forward: null output pointer");
  if (io->past_futures == nullptr && (c.stages == AFFT_STAGE_ALL || (c.stages == AFFT_STAGE_GPT && c.fp_output_len > 1)))
    return hfail(h, AFFT_ERR_INVALID, "forward: null past_futures pointer");
  for (int k = 0; c.stages == AFFT_STAGE_ALL && k < c.n_cls; ++k) {
    if (io->logits[k] == nullptr) return hfail(h, AFFT_ERR_INVALID, "forward: null logits pointer");
    if (io->ld_logits[k] < c.cls_dim[k] || io->ld_logits[k] % 4 != 0)
      return hfail(h, AFFT_ERR_INVALID, "forward: ld_logits must be a multiple of 4 and >= cls_dim");
  }
  if (h->w.size() != h->expected.size()) {
    char buf[512];
    int n = afft_missing_weights(h, buf, sizeof(buf));
    return hfail(h, AFFT_ERR_MISSING, "forward: " + std::to_string(n) + " weights not registered, e.g. " + std::string(buf));
  }
  cudaError_t e = cudaSetDevice(c.device);
  if (e != cudaSuccess) return hfail(h, AFFT_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));

  Fwd F;
  F.h = h;
  F.stream = static_cast<cudaStream_t>(stream_);
  F.strict = c.precision == AFFT_PREC_BF16X3;
  F.fp16 = c.precision == AFFT_PREC_FP16;
  h->launches = 0;
  h->prof.n = 0;
  if (h->profile) {  // clear the slots and record the start mark (stream-ordered ahead of the first kernel)
    h->prof_stream = F.stream;
    e = cudaMemsetAsync(h->prof_slots, 0, (AFFT_MAX_PROFILE_RECS + 1) * sizeof(unsigned long long), F.stream);
    if (e != cudaSuccess) return hfail(h, AFFT_ERR_CUDA, std::string("profile reset: ") + cudaGetErrorString(e));
    prof_mark_kernel<<<1, 32, 0, F.stream>>>(h->prof_slots);
  }
  if (c.stages == AFFT_STAGE_GPT) {
    run_gpt_only(F, *io, B);
    if (F.ok()) h->weights_fresh = false;
    return F.rc;
  }
  const int chunk = (h->fuser_chunk > 0) ? h->fuser_chunk : B;
  for (int b0 = 0; b0 < B && F.ok(); b0 += chunk) run_fuser(F, *io, b0, std::min(chunk, B - b0));
  if (F.ok() && c.stages == AFFT_STAGE_ALL) run_predictor(F, *io, B);
  if (F.ok()) h->weights_fresh = false;
  return F.rc;
}

extern "C" int afft_set_gemm_epilogue(int32_t v2) {
  gemm_epilogue_v2_flag().store(v2 != 0 ? 1 : 0, std::memory_order_relaxed);
  return AFFT_OK;
}

extern "C" int afft_set_gemm_skinny(int32_t on) {
  gemm_skinny_flag().store(on != 0 ? 1 : 0, std::memory_order_relaxed);
  return AFFT_OK;
}

extern "C" int afft_set_max_ksplit(afft_handle* h, int32_t max_split) {
  if (h == nullptr) return fail(AFFT_ERR_INVALID, "set_max_ksplit: null handle");
  if (max_split < 1 || max_split > 64) return hfail(h, AFFT_ERR_INVALID, "set_max_ksplit: value must be in [1, 64]");
  h->splitk.max_split = max_split;
  return AFFT_OK;
}

extern "C" int afft_plan_ksplit(int32_t tiles, int32_t slots, int32_t num_kb, int32_t ctas_per_tile, int32_t max_split) {
  if (tiles < 1 || slots < 1 || num_kb < 1 || ctas_per_tile < 1) return 1;
  static float dummy_partials;
  static unsigned dummy_counters;
  SplitKScratch sk{&dummy_partials, kSplitKPartialFloats, &dummy_counters, kSplitKCounters, max_split};  // never dereferenced
  return pick_ksplit(tiles, slots, num_kb, ctas_per_tile, static_cast<size_t>(kBlockM) * (ctas_per_tile == 2 ? 256 : 128), &sk);
}

extern "C" int afft_profile_enable(afft_handle* h, int32_t enable) {
  if (h == nullptr) return fail(AFFT_ERR_INVALID, "profile_enable: null handle");
  if (enable && h->prof_slots == nullptr) {
    cudaError_t e = cudaSetDevice(h->cfg.device);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->prof_slots), (AFFT_MAX_PROFILE_RECS + 1) * sizeof(unsigned long long));
    if (e != cudaSuccess) return hfail(h, AFFT_ERR_CUDA, std::string("profile_enable: ") + cudaGetErrorString(e));
  }
  h->profile = enable != 0;
  h->prof.n = 0;
  return AFFT_OK;
}

extern "C" int afft_profile_read(afft_handle* h, afft_profile* out) {
  if (h == nullptr || out == nullptr) return fail(AFFT_ERR_INVALID, "profile_read: null argument");
  if (h->prof.n > 0) {
    static thread_local unsigned long long host[AFFT_MAX_PROFILE_RECS + 1];
    cudaError_t e = cudaStreamSynchronize(h->prof_stream);
    if (e == cudaSuccess)
      e = cudaMemcpy(host, h->prof_slots, (h->prof.n + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return hfail(h, AFFT_ERR_CUDA, std::string("profile_read: ") + cudaGetErrorString(e));
    unsigned long long prev = host[0];
    for (int i = 0; i < h->prof.n; ++i) {  // launch i owns (end of launch i - 1, end of launch i]
      const unsigned long long t = host[1 + i] > prev ? host[1 + i] : prev;
      h->prof.recs[i].ms = static_cast<float>(static_cast<double>(t - prev) * 1e-6);
      prev = t;
    }
  }
  *out = h->prof;
  return AFFT_OK;
}
