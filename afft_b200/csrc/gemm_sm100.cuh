// The one dense-contraction kernel of the AFFT hot path:
//
//     C[M,N] = epilogue( A[M,K] . W[N,K]^T )         A, W bf16 (K-major), fp32 accumulation
//
// It replaces every nn.Linear / Conv1D the reference executes on the path
// (reference: models/feature_mapping.py:60, models/transformerblock.py:21,34,85-87,
//  models/future_prediction.py:108,248,254 and transformers' GPT-2 c_attn/c_proj/c_fc).
//
// sm_100a design: persistent CTAs (one per SM), warp-specialised:
//   warp 0      TMA producer   cp.async.bulk.tensor 2-D tiles, SWIZZLE_128B, 64-wide K slabs,
//                               multi-stage mbarrier ring
//   warp 1      MMA issuer     one thread issues tcgen05.mma (128 x BLOCK_N x 16, kind::f16) from
//                               shared-memory descriptors; accumulators live in TMEM, double
//                               buffered (2 x BLOCK_N columns) so the epilogue of tile i overlaps
//                               the main loop of tile i+1
//   warp 2      TMEM allocator
//   warps 4..11 epilogue       tcgen05.ld 32x32b -> registers -> swizzled smem transposition -> bias / GELU /
//                               residual / casts -> coalesced 16-B global accesses.  Two warps per 32-lane
//                               TMEM quadrant (= 32 output rows), interleaved over 32-column slabs.
//
// MODE selects the operand arithmetic:
//   MODE_BF16   (1)  bf16 operands, one tensor-core pass
//   MODE_FP16   (2)  fp16 operands, one pass at the same tensor rate: 11 significand bits instead of 8, i.e. 8x less
//                    operand rounding than bf16 (the mode that keeps the reference's top-5 at speed); the host side
//                    saturates on conversion, accumulation / residual / LayerNorm / softmax stay fp32
//   MODE_BF16X3 (3)  "strict" (SURVEY.md Appendix D): both operands are bf16 hi/lo pairs (x = hi + lo to ~16
//                    mantissa bits) and the kernel accumulates hi.hi + hi.lo + lo.hi into the same TMEM tile,
//                    i.e. three tensor-core passes per K slab for ~fp32-grade products.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ptx_sm100.cuh"

namespace afft {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int kUmmaK = 16;
constexpr int kNumEpilogueWarps = 8;   // two per TMEM lane quadrant
constexpr int kGemmThreads = 128 + 32 * kNumEpilogueWarps;

enum : int { MODE_BF16 = 1, MODE_FP16 = 2, MODE_BF16X3 = 3 };
enum : int { ACT_NONE = 0, ACT_GELU_ERF = 1, ACT_GELU_TANH = 2, ACT_RELU = 3, ACT_GATE = 4 };

// Everything the epilogue may do with an accumulator tile.  Pointers may be null (= skip).
// Row r of the GEMM is written to row  orow(r) = (r / row_group) * row_stride + r % row_group + row_off
// of every output (row_group == 0: orow = r).  The residual is read from the same mapped row of
// `res`, or from row r % res_mod when res_mod > 0 (position-embedding add).
struct GemmEpilogue {
  const float* bias;  // [N]
  const float* res;   // fp32, pitch ld_res
  long long ld_res;
  int res_mod;
  float* out_f32;  // pitch ld_f32
  long long ld_f32;
  __nv_bfloat16* out_hi;  // bf16(x), pitch ld_bf16
  __nv_bfloat16* out_lo;  // bf16(x - float(bf16(x))), pitch ld_bf16 (strict mode)
  long long ld_bf16;
  int act;
  int row_group, row_stride, row_off;
};

// Tile order and L2 policy.  n_fastest: consecutive tiles share the A (activation) row block and sweep the
// N tiles of the (small, L2-resident) weight, so every activation tile is fetched from DRAM once.
//
// Split-K (skinny problems: fewer tiles than SMs, e.g. the GPT-2 projections at the reference's eval batch of 32,
// or any GEMM at batch 1): a work unit is (tile, split); split s accumulates K blocks [s * kb_per, (s+1) * kb_per)
// into TMEM and its epilogue warps store the raw fp32 partial tile to `partials` (unit-major, tile-local layout,
// so the scratch is bounded by one tile per CTA).  Each epilogue warp then bumps the counter of its (tile, warp)
// band; the warp that arrives last sums the `ksplit` partials of its band IN SPLIT ORDER (bit-reproducible
// whatever the arrival order) and runs the normal epilogue on the sum.  Counters reset themselves.
struct GemmSched {
  int n_fastest;
  unsigned long long policy_a, policy_b;
  int ksplit;           // 1 = off
  float* partials;      // >= grid CTAs * 128 * BLOCK_N floats
  unsigned* counters;   // >= tiles * (CTAs per tile) * kNumEpilogueWarps, zero before the first launch
  unsigned long long* t_end;  // profiling slot (ptx::prof_mark_end) or nullptr
  // 2-CTA kernel only: depth of the TMA -> MMA ring (runtime: the TMA-staged epilogue trades ring stages for slab buffers)
  int stages;
  // TMA-staged epilogue (epilogue_v2 below): 0 = off, else the bytes of shared memory each epilogue warp owns
  int v2_warp_bytes;
};

// Tensor maps of the TMA-staged epilogue (fp32 boxes 32 rows x 32 columns = 128-byte rows, SWIZZLE_128B; 16-bit boxes
// 32 x 32 = 64-byte rows, SWIZZLE_64B).
struct GemmTmaEpi {
  CUtensorMap res;  // fp32 residual operand
  CUtensorMap f32;  // fp32 output
  CUtensorMap b16;  // bf16 / fp16 output
};

template <int BLOCK_N, int MODE>
struct GemmTraits {
  static_assert(BLOCK_N == 128 || BLOCK_N == 256, "BLOCK_N must be 128 or 256");
  static_assert(MODE >= 1 && MODE <= 3, "MODE must be MODE_BF16, MODE_FP16 or MODE_BF16X3");
  static constexpr int kPairs = (MODE == MODE_BF16X3) ? 2 : 1;
  static constexpr uint32_t kABytes = kBlockM * kBlockK * 2;
  static constexpr uint32_t kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr uint32_t kStageBytes = kPairs * (kABytes + kBBytes);
  static constexpr int kStages = (MODE == MODE_BF16X3) ? (BLOCK_N == 256 ? 2 : 3) : (BLOCK_N == 256 ? 4 : 6);
  static constexpr uint32_t kTmemCols = 2 * BLOCK_N;
  static constexpr uint32_t kBarrierBytes = 256;
  // per epilogue warp: one 32 x 32 fp32 transposition tile (4 KB, 16-B chunks XOR-swizzled by row)
  static constexpr uint32_t kStagingBytes = kNumEpilogueWarps * 4096;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kStagingBytes + kBarrierBytes + 1024;  // +align slack
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
};

// --------------------------------------------------------------------------------------------
// Activation functions (fp32 in, fp32 out), written against the MUFU approximations directly so that an
// element costs ~16 (erf) / ~8 (tanh) issue slots: at K = 1024 the tensor core finishes a 128x256 tile
// in 8192 cycles = 0.25 cycle per element, and the epilogue has to keep up with that.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ float mufu_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float gelu_erf(float x) {
  // x * Phi(x), Phi(x) = 0.5 erfc(-x / sqrt 2)   (torch.nn.GELU default; reference models/transformerblock.py:79)
  // erfc(z) = t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) exp(-z^2), t = 1 / (1 + p z)   Abramowitz-Stegun 7.1.26,
  // |error| < 1.5e-7.  q = 0.5 erfc(|x| / sqrt 2);  Phi = x < 0 ? q : 1 - q.
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = mufu_rcp(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  const float e = mufu_ex2((z * -1.4426950408889634f) * z);  // exp(-z^2)
  const float q = (p * t) * e;
  // x >= 0: x (1 - q);  x < 0: x q = -|x| q   =>   relu(x) - |x| q for both signs
  return fmaf(-fabsf(x), q, fmaxf(x, 0.f));
}

__device__ __forceinline__ float gelu_tanh(float x) {
  // 0.5 x (1 + tanh(u)), u = sqrt(2/pi) (x + 0.044715 x^3)   (HF "gelu_new", GPT-2 MLP)
  // = x * sigmoid(2u) = x - x / (1 + exp(2u));  exp -> inf gives x, exp -> 0 gives 0: both limits exact.
  const float k0 = 2.0f * 0.7978845608028654f * 1.4426950408889634f;  // 2 sqrt(2/pi) log2(e)
  const float w = fmaf(k0 * 0.044715f, x * x, k0);
  const float e = mufu_ex2(x * w);
  const float r = mufu_rcp(1.0f + e);
  return fmaf(-x, r, x);
}

__device__ __forceinline__ float gelu_erf_bf16out(float x) {
  // erf-GELU for results that are rounded to bf16 straight away (FC1 of the fuser MLP in bf16 mode).
  // Phi(x) = 0.5 (1 + erf(x / sqrt 2)) = sigmoid(2 g(x)) with g = atanh(erf(x / sqrt 2)), an odd function; g(x) / x is
  // fitted by c0 + c1 x^2 + c2 x^4 (minimax on x Phi(x) over |x| <= 9: max |error| 2.6e-5, i.e. <= 1/4 of a bf16
  // half-ulp wherever |x Phi(x)| >= 0.05, and the saturated limits are exact).  8 issue slots instead of 15.
  const float l2e2 = 2.0f * 1.4426950408889634f;
  const float x2 = fminf(x * x, 81.0f);  // the fit (and w > 0) holds for |x| <= 9; beyond that Phi is 0 or 1 in fp32
  float w = fmaf(l2e2 * -0.0003515187397213507f, x2, l2e2 * 0.03700565900935642f);
  w = fmaf(w, x2, l2e2 * 0.7975078687760178f);
  const float e = mufu_ex2(x * w);
  const float r = mufu_rcp(1.0f + e);
  return fmaf(-x, r, x);
}

// --------------------------------------------------------------------------------------------
// Epilogue variants.  The combinations the forward path uses are compiled with their flags as
// constants (the 8-row unrolled epilogue of the all-runtime version is ~60 KB of SASS and thrashes
// the instruction cache - ncu: stall_no_inst dominated); EPI_GENERIC keeps every flag at run time
// for the stateless afft_gemm() entry point.
//   bit 0-1 activation, bit 2 residual, bit 3 fp32 output, bit 4 bf16 output (+ lo when MODE == MODE_BF16X3)
// ACT_RELU / ACT_GATE (ablation mappings, MATT) exist in the generic variant only.
// --------------------------------------------------------------------------------------------
constexpr int EPI_GENERIC = -1;
constexpr int epi_code(int act, bool res, bool f32, bool bf16) {
  return act | (res ? 4 : 0) | (f32 ? 8 : 0) | (bf16 ? 16 : 0);
}

template <int EPI, int MODE>
struct EpiFlags {
  static constexpr bool kGeneric = EPI < 0;
  __device__ static __forceinline__ int act(const GemmEpilogue& ep) { return kGeneric ? ep.act : (EPI & 3); }
  __device__ static __forceinline__ bool res(const GemmEpilogue& ep) { return kGeneric ? ep.res != nullptr : (EPI & 4) != 0; }
  __device__ static __forceinline__ bool f32(const GemmEpilogue& ep) { return kGeneric ? ep.out_f32 != nullptr : (EPI & 8) != 0; }
  __device__ static __forceinline__ bool bf16(const GemmEpilogue& ep) { return kGeneric ? ep.out_hi != nullptr : (EPI & 16) != 0; }
  __device__ static __forceinline__ bool lo(const GemmEpilogue& ep) { return kGeneric ? ep.out_lo != nullptr : (MODE == MODE_BF16X3); }
};

__device__ __forceinline__ uint32_t pack2_bf16_rn(float a, float b) {  // a -> low half
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// 16-bit operand packing of the epilogue: bf16 (round to nearest even), or fp16 saturated to the finite range
__device__ __forceinline__ uint32_t pack2_f16_sat(float a, float b) {  // a -> low half; one F2FP.SATFINITE instruction
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
template <int MODE>
__device__ __forceinline__ uint32_t pack2_operand(float a, float b) {
  return MODE == MODE_FP16 ? pack2_f16_sat(a, b) : pack2_bf16_rn(a, b);
}
__device__ __forceinline__ float bf16_lo_f32(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16_hi_f32(uint32_t packed) { return __uint_as_float(packed & 0xffff0000u); }

template <int EPI, int MODE>
__device__ __forceinline__ float4 epilogue_math(const GemmEpilogue& ep, float4 x, const float4& bias4) {
  using F = EpiFlags<EPI, MODE>;
  x.x += bias4.x;
  x.y += bias4.y;
  x.z += bias4.z;
  x.w += bias4.w;
  const int act = F::act(ep);
  if (act == ACT_GELU_ERF) {
    if (!F::kGeneric && MODE != MODE_BF16X3 && !F::f32(ep)) {
      // result only ever stored as a 16-bit operand.  fp16 as well: the fit's absolute error (2.5e-5) is 1/10 of the fp16
      // rounding of an O(1) activation, and a degree-4 fit (3.2e-6) cost FC1 10 % (measured, 220 vs 200 us) for no change
      // in the 1024-clip parity statistics
      x.x = gelu_erf_bf16out(x.x);
      x.y = gelu_erf_bf16out(x.y);
      x.z = gelu_erf_bf16out(x.z);
      x.w = gelu_erf_bf16out(x.w);
    } else {
      x.x = gelu_erf(x.x);
      x.y = gelu_erf(x.y);
      x.z = gelu_erf(x.z);
      x.w = gelu_erf(x.w);
    }
  } else if (act == ACT_GELU_TANH) {
    x.x = gelu_tanh(x.x);
    x.y = gelu_tanh(x.y);
    x.z = gelu_tanh(x.z);
    x.w = gelu_tanh(x.w);
  } else if (F::kGeneric && act == ACT_RELU) {
    x.x = fmaxf(x.x, 0.f);
    x.y = fmaxf(x.y, 0.f);
    x.z = fmaxf(x.z, 0.f);
    x.w = fmaxf(x.w, 0.f);
  } else if (F::kGeneric && act == ACT_GATE) {
    x.x = 1.0f / (1.0f + __expf(-x.x));
    x.y = 1.0f / (1.0f + __expf(-x.y));
    x.z = 1.0f / (1.0f + __expf(-x.z));
    x.w = 1.0f / (1.0f + __expf(-x.w));
  }
  return x;
}

// residual combine: + for the residual stream, * for ACT_GATE (the "residual" is the gated value)
template <int EPI, int MODE>
__device__ __forceinline__ float epilogue_combine(const GemmEpilogue& ep, float x, float r) {
  if (EpiFlags<EPI, MODE>::kGeneric && ep.act == ACT_GATE) return x * r;
  return x + r;
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 x;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(addr));
  return x;
}

__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Per-tile row pointers of the tensors the epilogue touches (column 0 of each of this lane's 8 rows), so that the
// per-slab address is one 64-bit add per row and tensor.
struct EpiRowPtrs {
  float* f32[8];
  const float* res[8];
  __nv_bfloat16* hi[8];
  __nv_bfloat16* lo[8];
};

template <int EPI, int MODE>
__device__ __forceinline__ void epilogue_row_ptrs(const GemmEpilogue& ep, int row0, int sub_row, EpiRowPtrs& P) {
  using F = EpiFlags<EPI, MODE>;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = row0 + i * 4 + sub_row;
    long long orow = r;
    if (ep.row_group > 0)
      orow = static_cast<long long>(r / ep.row_group) * ep.row_stride + (r % ep.row_group) + ep.row_off;
    if (F::f32(ep)) P.f32[i] = ep.out_f32 + orow * ep.ld_f32;
    if (F::res(ep)) P.res[i] = ep.res + ((ep.res_mod > 0) ? static_cast<long long>(r % ep.res_mod) : orow) * ep.ld_res;
    if (F::bf16(ep)) P.hi[i] = ep.out_hi + orow * ep.ld_bf16;
    if (F::bf16(ep) && F::lo(ep)) P.lo[i] = ep.out_lo + orow * ep.ld_bf16;
  }
}

// Optional L2 prefetch of the residual block of the NEXT tile (AFFT_RES_PREFETCH, off).  Together with the TMEM-load
// software pipeline (AFFT_TMEM_PIPE, off) and the one-slab-ahead residual load (AFFT_RES_AHEAD, off) it was built on
// the hypothesis that the residual GEMMs (50 % tensor-active at K = 1024) are bound by the epilogue's dependent
// chain.  A/B builds timed on the same box say otherwise: none of the three helps.  The limiter is the L2 -> SM
// operand feed the epilogue's residual reads share with the TMA loads; the switches stay for
// re-testing once the main loop's feed changes.  Lane (sub_row, chunk) covers row chunk * 4 + sub_row: 4 x 128 B.
template <int EPI, int MODE, int BLOCK_N>
__device__ __forceinline__ void epilogue_prefetch_residual(const GemmEpilogue& ep, int row0, int n_tile0, int egrp, int lane,
                                                           int M, int N) {
  using F = EpiFlags<EPI, MODE>;
#ifndef AFFT_RES_PREFETCH
#define AFFT_RES_PREFETCH 0  // A/B on one box (tools/gemm_time.py): no gain (60.7 us off vs 61.9 us on, K = 1024 projection)
#endif
  if (!AFFT_RES_PREFETCH || EPI < 0 || !F::res(ep)) return;
  if (row0 + 32 > M) return;
  const int r = row0 + (lane & 7) * 4 + (lane >> 3);
  long long rrow = r;
  if (ep.res_mod > 0) rrow = r % ep.res_mod;
  else if (ep.row_group > 0) rrow = static_cast<long long>(r / ep.row_group) * ep.row_stride + (r % ep.row_group) + ep.row_off;
  const float* p = ep.res + rrow * ep.ld_res;
#pragma unroll
  for (int s = 0; s < BLOCK_N / 32 / (kNumEpilogueWarps / 4); ++s) {
    const int col = n_tile0 + (egrp + s * (kNumEpilogueWarps / 4)) * 32;
    if (col + 32 <= N) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + col));
  }
}

// One 32-row x 32-column slab, interior case: every row < M and every column < N, so no predicates.
// Lane (sub_row, chunk) handles rows sub_row + 4 i (i < 8), columns col .. col + 3.
//
// Residual operand: `r` may arrive PRELOADED (issued one slab earlier by this function, see next_col), so that the
// HBM / L2 latency of the residual overlaps the stores of the previous slab and the TMEM wait / transposition of this
// one instead of sitting between the shared-memory read and the add.  All residual loads of a slab are issued
// before any of its stores (the residual may alias the output: in-place stream); loads of the NEXT slab touch other
// columns, so they may be issued ahead of this slab's stores.
template <int EPI, int MODE>
__device__ __forceinline__ void epilogue_slab_full(const GemmEpilogue& ep, uint32_t stage, int sub_row, int chunk, int col,
                                                   const EpiRowPtrs& P, const float4& bias4, float4 (&r)[8], bool preloaded,
                                                   int next_col) {
  using F = EpiFlags<EPI, MODE>;
  float4 x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rl = i * 4 + sub_row;
    x[i] = lds128(stage + static_cast<uint32_t>(rl) * 128u + static_cast<uint32_t>((chunk ^ (rl & 7)) * 16));
  }
  if (F::res(ep)) {
    if (!preloaded) {
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = *reinterpret_cast<const float4*>(P.res[i] + col);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      x[i] = epilogue_math<EPI, MODE>(ep, x[i], bias4);
      x[i].x = epilogue_combine<EPI, MODE>(ep, x[i].x, r[i].x);
      x[i].y = epilogue_combine<EPI, MODE>(ep, x[i].y, r[i].y);
      x[i].z = epilogue_combine<EPI, MODE>(ep, x[i].z, r[i].z);
      x[i].w = epilogue_combine<EPI, MODE>(ep, x[i].w, r[i].w);
    }
    if (next_col >= 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = *reinterpret_cast<const float4*>(P.res[i] + next_col);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = epilogue_math<EPI, MODE>(ep, x[i], bias4);
  }
  if (F::f32(ep)) {
#pragma unroll
    for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(P.f32[i] + col) = x[i];
  }
  if (F::bf16(ep)) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t h01 = pack2_operand<MODE>(x[i].x, x[i].y), h23 = pack2_operand<MODE>(x[i].z, x[i].w);
      *reinterpret_cast<uint2*>(P.hi[i] + col) = make_uint2(h01, h23);
      if (F::lo(ep)) {
        const uint32_t l01 = pack2_bf16_rn(x[i].x - bf16_lo_f32(h01), x[i].y - bf16_hi_f32(h01));
        const uint32_t l23 = pack2_bf16_rn(x[i].z - bf16_lo_f32(h23), x[i].w - bf16_hi_f32(h23));
        *reinterpret_cast<uint2*>(P.lo[i] + col) = make_uint2(l01, l23);
      }
    }
  }
}

// Edge slab (partial rows and/or columns): element-wise predicates; kept out of line so the interior path
// stays small in the instruction cache.
template <int EPI, int MODE>
__device__ __noinline__ void epilogue_slab_edge(const GemmEpilogue& ep, uint32_t stage, int sub_row, int chunk, int row0,
                                                int col, int M, int N) {
  using F = EpiFlags<EPI, MODE>;
  if (col >= N) return;
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (ep.bias != nullptr) {
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (col + e < N) bias[e] = __ldg(ep.bias + col + e);
  }
  const float4 bias4 = make_float4(bias[0], bias[1], bias[2], bias[3]);
#pragma unroll 1
  for (int i = 0; i < 8; ++i) {
    const int rl = i * 4 + sub_row;
    const int row = row0 + rl;
    if (row >= M) break;
    long long orow = row;
    if (ep.row_group > 0)
      orow = static_cast<long long>(row / ep.row_group) * ep.row_stride + (row % ep.row_group) + ep.row_off;
    float4 x4 = lds128(stage + static_cast<uint32_t>(rl) * 128u + static_cast<uint32_t>((chunk ^ (rl & 7)) * 16));
    x4 = epilogue_math<EPI, MODE>(ep, x4, bias4);
    float x[4] = {x4.x, x4.y, x4.z, x4.w};
    const long long rrow = (ep.res_mod > 0) ? static_cast<long long>(row % ep.res_mod) : orow;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (col + e >= N) break;
      float v = x[e];
      if (F::res(ep)) v = epilogue_combine<EPI, MODE>(ep, v, ep.res[rrow * ep.ld_res + col + e]);
      if (F::f32(ep)) ep.out_f32[orow * ep.ld_f32 + col + e] = v;
      if (F::bf16(ep)) {
        if (MODE == MODE_FP16) {
          reinterpret_cast<__half*>(ep.out_hi)[orow * ep.ld_bf16 + col + e] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
        } else {
          const __nv_bfloat16 h = __float2bfloat16_rn(v);
          ep.out_hi[orow * ep.ld_bf16 + col + e] = h;
          if (F::lo(ep)) ep.out_lo[orow * ep.ld_bf16 + col + e] = __float2bfloat16_rn(v - __bfloat162float(h));
        }
      }
    }
  }
}

// --------------------------------------------------------------------------------------------
// Split-K helpers (see GemmSched).  `band` = this warp's 32 rows x BLOCK_N columns of the CTA's 128-row half tile;
// the warp owns slabs c = egrp, egrp + 2, ... of it, in the partial buffer as in the accumulator.
// --------------------------------------------------------------------------------------------
template <int BLOCK_N>
__device__ __forceinline__ void splitk_store_partial(float* cta_partial, uint32_t stage, int quad, int sub_row, int chunk, int c) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rl = i * 4 + sub_row;
    const float4 x = lds128(stage + static_cast<uint32_t>(rl) * 128u + static_cast<uint32_t>((chunk ^ (rl & 7)) * 16));
    __stcg(reinterpret_cast<float4*>(cta_partial + static_cast<size_t>(quad * 32 + rl) * BLOCK_N + c * 32 + chunk * 4), x);
  }
}

// Sum of the ksplit partials of one slab, in split order, written back into the warp's staging tile in exactly the
// layout the TMEM transposition produces (each lane later re-reads the positions it wrote itself).
template <int BLOCK_N>
__device__ __forceinline__ void splitk_sum_to_stage(const float* tile_partial0, size_t split_stride, int ksplit, uint32_t stage,
                                                    int quad, int sub_row, int chunk, int c) {
  float4 x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int sp = 0; sp < ksplit; ++sp) {
    const float* p = tile_partial0 + sp * split_stride;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rl = i * 4 + sub_row;
      const float4 v = __ldcg(reinterpret_cast<const float4*>(p + static_cast<size_t>(quad * 32 + rl) * BLOCK_N + c * 32 + chunk * 4));
      x[i].x += v.x;
      x[i].y += v.y;
      x[i].z += v.z;
      x[i].w += v.w;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rl = i * 4 + sub_row;
    const uint32_t addr = stage + static_cast<uint32_t>(rl) * 128u + static_cast<uint32_t>((chunk ^ (rl & 7)) * 16);
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(x[i].x), "f"(x[i].y), "f"(x[i].z), "f"(x[i].w) : "memory");
  }
}

// The epilogue of one work unit for one warp (both kernels).  Pass 0 drains the warp's slabs from TMEM (optionally
// software pipelined, AFFT_TMEM_PIPE: the TMEM load of the next slab in flight while the current one is
// transposed, transformed and stored) and either finishes them (ksplit == 1) or stores them as split-K partials;
// `release_tmem()` then hands the accumulator back to the MMA warp.  With split-K the warp that arrives last on
// the band counter runs pass 1: same slab loop, but the staging tile is filled with the in-order sum of the
// partials instead of from TMEM.
template <int EPI, int MODE, int BLOCK_N, typename ReleaseFn>
__device__ __forceinline__ void epilogue_unit(const GemmEpilogue& ep, int ksplit, uint32_t stage, uint32_t t_row, int quad, int egrp,
                                              int lane, int row0, int n_tile0, int M, int N, float* cta_partial,
                                              const float* tile_partial0, size_t split_stride, unsigned* counter,
                                              int next_row0, int next_n_tile0, ReleaseFn release_tmem) {
  constexpr int kCStep = kNumEpilogueWarps / 4;
#ifndef AFFT_TMEM_PIPE
#define AFFT_TMEM_PIPE 0  // A/B on one box: no gain (59.1 us off vs 61.9 us on; FC1 160.4 vs 162.1 us)
#endif
  constexpr bool kPipe = AFFT_TMEM_PIPE && EPI >= 0;  // the all-runtime epilogue has no registers to spare for the overlap
  const int sub_row = lane >> 3;    // row within a group of 4
  const int chunk = lane & 7;       // 16-byte chunk = 4 fp32 columns
  const bool rows_full = (row0 + 32 <= M);
  EpiRowPtrs rp;
  epilogue_row_ptrs<EPI, MODE>(ep, row0, sub_row, rp);
  if (ksplit == 1 && next_row0 >= 0) epilogue_prefetch_residual<EPI, MODE, BLOCK_N>(ep, next_row0, next_n_tile0, egrp, lane, M, N);
  using F = EpiFlags<EPI, MODE>;
#ifndef AFFT_RES_AHEAD
#define AFFT_RES_AHEAD 0  // measured (A/B on one box): 61.6 us without vs 67.4 us with, on the K = 1024 projection
#endif
  constexpr bool kResAhead = AFFT_RES_AHEAD && EPI >= 0 && (EPI & 4) != 0;  // compiled-in residual variants: load one slab ahead
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    uint32_t v[32];
    float4 r[8];
    bool r_loaded = false;
    int c = egrp;
    bool have = n_tile0 + c * 32 < N;  // warp-uniform
    if (pass == 0 && have && kPipe) ptx::tmem_ld_32x32(t_row + c * 32, v);
    if (kResAhead && pass == 0 && ksplit == 1 && have && rows_full && n_tile0 + c * 32 + 32 <= N) {
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = *reinterpret_cast<const float4*>(rp.res[i] + n_tile0 + c * 32 + chunk * 4);
      r_loaded = true;
    }
#pragma unroll 1
    while (have) {
      const int n0 = n_tile0 + c * 32;
      const int cn = c + kCStep;
      have = cn < BLOCK_N / 32 && n_tile0 + cn * 32 < N;
      if (pass == 0) {
        if (!kPipe) ptx::tmem_ld_32x32(t_row + c * 32, v);
        ptx::tmem_ld_wait();
        // own row (= lane) -> staging, 16-B chunk j at position j ^ (lane & 7): conflict-free
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t addr = stage + static_cast<uint32_t>(lane) * 128u + static_cast<uint32_t>((j ^ (lane & 7)) * 16);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[4 * j]), "r"(v[4 * j + 1]),
                       "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                       : "memory");
        }
        if (have && kPipe) ptx::tmem_ld_32x32(t_row + cn * 32, v);
      } else {
        splitk_sum_to_stage<BLOCK_N>(tile_partial0, split_stride, ksplit, stage, quad, sub_row, chunk, c);
      }
      __syncwarp();
      const int col = n0 + chunk * 4;
      if (pass == 0 && ksplit > 1) {
        splitk_store_partial<BLOCK_N>(cta_partial, stage, quad, sub_row, chunk, c);
      } else if (row0 < M) {
        if (rows_full && n0 + 32 <= N) {
          float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ep.bias != nullptr) bias4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
          // the next slab's residual is requested now if that slab is an interior one too
          const bool next_full = kResAhead && r_loaded && have && n_tile0 + cn * 32 + 32 <= N;
          epilogue_slab_full<EPI, MODE>(ep, stage, sub_row, chunk, col, rp, bias4, r, r_loaded,
                                         next_full ? n_tile0 + cn * 32 + chunk * 4 : -1);
          r_loaded = next_full;
        } else {
          epilogue_slab_edge<EPI, MODE>(ep, stage, sub_row, chunk, row0, col, M, N);
        }
      }
      __syncwarp();  // staging tile is rewritten by the next slab
      c = cn;
    }
    if (pass == 1) break;
    ptx::tcgen05_fence_before();
    __syncwarp();
    release_tmem();
    if (ksplit == 1) break;
    __threadfence();  // this lane's partial stores are visible device-wide before the band counter moves
    __syncwarp();
    unsigned old = 0;
    if (lane == 0) old = atomicAdd(counter, 1u);
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old != static_cast<unsigned>(ksplit - 1)) break;  // another split's warp will finish this band
    __threadfence();
    if (lane == 0) *counter = 0u;  // ready for the next launch (stream-ordered)
  }
}

// --------------------------------------------------------------------------------------------
// The kernel
// --------------------------------------------------------------------------------------------
template <int BLOCK_N, int MODE, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                         const __grid_constant__ CUtensorMap tm_a_lo,
                         const __grid_constant__ CUtensorMap tm_b_lo, const __grid_constant__ GemmEpilogue ep, const int M,
                         const int N, const int K, const __grid_constant__ GemmSched sched) {
  using T = GemmTraits<BLOCK_N, MODE>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_raw_u32 = ptx::smem_u32(smem_raw);
  const uint32_t smem_base = (smem_raw_u32 + 1023u) & ~1023u;  // SWIZZLE_128B tiles: 1024-B aligned
  const uint32_t staging_base = smem_base + T::kStages * T::kStageBytes;
  const uint32_t bar_base = staging_base + T::kStagingBytes;
  auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (T::kStages + s); };
  auto tmem_full_bar = [&](uint32_t a) { return bar_base + 8u * (2 * T::kStages + a); };
  auto tmem_empty_bar = [&](uint32_t a) { return bar_base + 8u * (2 * T::kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * T::kStages + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_raw_u32));

  ptx::griddep_launch();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (M + kBlockM - 1) / kBlockM;
  const int num_n = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + kBlockK - 1) / kBlockK;
  const int ksplit = sched.ksplit;  // work unit = (tile, split); ksplit == 1: unit = tile
  const int num_units = num_tiles * ksplit;
  const int kb_per = (num_kb + ksplit - 1) / ksplit;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_a);
    ptx::prefetch_tensormap(&tm_b);
    if (MODE == MODE_BF16X3) {
      ptx::prefetch_tensormap(&tm_a_lo);
      ptx::prefetch_tensormap(&tm_b_lo);
    }
  }
  if (warp == 1 && lane == 0) {
    for (uint32_t s = 0; s < T::kStages; ++s) {
      ptx::mbar_init(full_bar(s), 1);   // producer's arrive.expect_tx
      ptx::mbar_init(empty_bar(s), 1);  // tcgen05.commit
    }
    for (uint32_t a = 0; a < 2; ++a) {
      ptx::mbar_init(tmem_full_bar(a), 1);   // tcgen05.commit
      ptx::mbar_init(tmem_empty_bar(a), kNumEpilogueWarps);  // one arrive per epilogue warp
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc<T::kTmemCols>(tmem_slot);
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  ptx::griddep_wait();  // operands / residual written by the previous kernel are visible from here on

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        const int tile = unit / ksplit, split = unit - tile * ksplit;
        const int m_idx = sched.n_fastest ? tile / num_n : tile % num_m;
        const int n_idx = sched.n_fastest ? tile % num_n : tile / num_m;
        const int kb0 = split * kb_per, kb1 = min(num_kb, kb0 + kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t a_dst = smem_base + stage * T::kStageBytes;
          const uint32_t b_dst = a_dst + T::kABytes;
          const uint32_t fb = full_bar(stage);
          ptx::mbar_arrive_expect_tx(fb, T::kStageBytes);
          ptx::tma_load_2d(a_dst, &tm_a, fb, kb * kBlockK, m_idx * kBlockM, sched.policy_a);
          ptx::tma_load_2d(b_dst, &tm_b, fb, kb * kBlockK, n_idx * BLOCK_N, sched.policy_b);
          if (MODE == MODE_BF16X3) {
            const uint32_t a_lo_dst = b_dst + T::kBBytes;
            const uint32_t b_lo_dst = a_lo_dst + T::kABytes;
            ptx::tma_load_2d(a_lo_dst, &tm_a_lo, fb, kb * kBlockK, m_idx * kBlockM, sched.policy_a);
            ptx::tma_load_2d(b_lo_dst, &tm_b_lo, fb, kb * kBlockK, n_idx * BLOCK_N, sched.policy_b);
          }
          if (++stage == T::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16_f32(kBlockM, BLOCK_N, MODE == MODE_FP16);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        const int split = unit % ksplit;
        const int kb0 = split * kb_per, kb1 = min(num_kb, kb0 + kb_per);
        ptx::mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1u);  // epilogue has drained this buffer
        ptx::tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tcgen05_fence_after();
          const uint32_t a_src = smem_base + stage * T::kStageBytes;
          const uint32_t b_src = a_src + T::kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            ptx::umma_bf16_ss(d_tmem, ptx::make_smem_desc_sw128(a_src + k * (kUmmaK * 2)),
                              ptx::make_smem_desc_sw128(b_src + k * (kUmmaK * 2)), idesc,
                              ((kb - kb0) | k) != 0 ? 1u : 0u);
          }
          if (MODE == MODE_BF16X3) {
            const uint32_t a_lo_src = b_src + T::kBBytes;
            const uint32_t b_lo_src = a_lo_src + T::kABytes;
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              ptx::umma_bf16_ss(d_tmem, ptx::make_smem_desc_sw128(a_src + k * (kUmmaK * 2)),
                                ptx::make_smem_desc_sw128(b_lo_src + k * (kUmmaK * 2)), idesc, 1u);
            }
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              ptx::umma_bf16_ss(d_tmem, ptx::make_smem_desc_sw128(a_lo_src + k * (kUmmaK * 2)),
                                ptx::make_smem_desc_sw128(b_src + k * (kUmmaK * 2)), idesc, 1u);
            }
          }
          ptx::umma_commit(empty_bar(stage));  // frees the smem slot when these MMAs retire
          if (++stage == T::kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        ptx::umma_commit(tmem_full_bar(acc));  // accumulator complete -> epilogue
        acc ^= 1u;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ======================= epilogue =======================
    // Two warps per TMEM lane quadrant (warp % 4), interleaved over the 32-column slabs of the tile.
    // A slab is read from TMEM row-per-thread, transposed through a private swizzled smem tile, and
    // leaves as fully coalesced 16-byte accesses: each warp instruction covers 4 rows x 128 B.
    const int quad = warp & 3;
    const int egrp = (warp - 4) >> 2;
    const uint32_t stage = staging_base + static_cast<uint32_t>(warp - 4) * 4096u;
    constexpr size_t kTileFloats = static_cast<size_t>(kBlockM) * BLOCK_N;
    uint32_t acc = 0, acc_phase = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      const int tile = unit / ksplit;
      const int m_idx = sched.n_fastest ? tile / num_n : tile % num_m;
      const int n_idx = sched.n_fastest ? tile % num_n : tile / num_m;
      const int row0 = m_idx * kBlockM + quad * 32;
      // residual of the NEXT unit of this CTA -> L2 (and of this one, for the CTA's first unit)
      if (unit == static_cast<int>(blockIdx.x) && ksplit == 1)
        epilogue_prefetch_residual<EPI, MODE, BLOCK_N>(ep, row0, n_idx * BLOCK_N, egrp, lane, M, N);
      int next_row0 = -1, next_n0 = 0;
      if (unit + static_cast<int>(gridDim.x) < num_units) {
        const int nt = (unit + static_cast<int>(gridDim.x)) / ksplit;
        next_row0 = (sched.n_fastest ? nt / num_n : nt % num_m) * kBlockM + quad * 32;
        next_n0 = (sched.n_fastest ? nt % num_n : nt / num_m) * BLOCK_N;
      }
      ptx::mbar_wait(tmem_full_bar(acc), acc_phase);
      ptx::tcgen05_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N;
      const uint32_t empty_addr = tmem_empty_bar(acc);
      epilogue_unit<EPI, MODE, BLOCK_N>(ep, ksplit, stage, t_row, quad, egrp, lane, row0, n_idx * BLOCK_N, M, N,
                                         sched.partials + static_cast<size_t>(unit) * kTileFloats,
                                         sched.partials + static_cast<size_t>(tile) * ksplit * kTileFloats, kTileFloats,
                                         sched.counters + tile * kNumEpilogueWarps + (warp - 4), next_row0, next_n0,
                                         [&] { if (lane == 0) ptx::mbar_arrive(empty_addr); });
      acc ^= 1u;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc<T::kTmemCols>(tmem_base);
  }
  if (warp == 0) ptx::prof_mark_end(sched.t_end);
}


// --------------------------------------------------------------------------------------------
// 2-CTA variant: a CTA pair (cluster of 2 = the two SMs of a TPC) computes one 256 x 256 tile with
// tcgen05.mma.cta_group::2 (M = 256).  Each CTA loads its own 128 rows of A and HALF of the B tile
// (128 of the 256 weight rows), so a K slab costs 32 KB of L2->SM traffic per SM instead of 48 KB:
// the main loop of the 1-CTA kernel is bound by that traffic (96 B/clk/SM at full tensor rate).
//   - both CTAs run a TMA producer; all transaction bytes are signalled on the LEADER's full barrier
//   - the leader's MMA thread issues for the pair; tcgen05.commit multicasts the slot-free / accumulator-ready
//     arrivals to both CTAs
//   - each CTA's epilogue warps drain their own 128 x 256 TMEM accumulator and arrive (remotely for the peer)
//     on the leader's tmem-empty barrier
// --------------------------------------------------------------------------------------------
// --------------------------------------------------------------------------------------------
// TMA-staged epilogue of the 2-CTA kernel ("v2").  The v1 epilogue above transposes every accumulator slab through
// shared memory so that the warp's global accesses coalesce, and fetches the fp32 residual with LDGs the adds then wait
// for (ncu: long_scoreboard on the FADD; 50 % tensor-active on the K = 1024 projection).  Here a thread keeps the TMEM
// layout - lane = output row, 32 consecutive columns in registers - and the TMA engine moves the data:
//   * the fp32 residual slab (32 rows x 128 B) is fetched into shared memory by cp.async.bulk.tensor one slab ahead
//     (the slab after the current one, across tile boundaries) behind a per-warp mbarrier pair, SWIZZLE_128B so that
//     the row-per-lane 16-byte shared-memory accesses are conflict free;
//   * results are written back into the same swizzled slab (fp32, in place over the residual) and / or a 64-byte-row
//     SWIZZLE_64B slab (16-bit operand for the next GEMM) and leave through cp.async.bulk.tensor stores - no STG, no
//     transposition;
//   * the two slab buffers of a warp alternate: a buffer is refilled only after the bulk-group of the store that read
//     it has finished reading (cp.async.bulk.wait_group.read).
// Per epilogue warp: quadrant = warp % 4 (TMEM lanes = rows), column half = (warp - 4) / 4 (128 of the 256 columns:
// four 32-column slabs).
// Conditions (checked by the launcher): ksplit == 1, no output row map, residual at the output row (res_mod == 0),
// MODE_BF16 / MODE_FP16, N a multiple of 32, one of the compiled epilogue variants.
//
// STATUS: opt-in (afft_set_gemm_epilogue(1) / AFFT_GEMM_EPI_V2=1), NOT the default.  A/B on one B200 (isolated launches,
// profiles/r02_epilogue_v2_ab.txt): the K = 1024 residual projection 57.1 vs 57.8 us (v1), i.e. no gain - that GEMM moves
// 235 MB per launch (47 MB A + 94 MB residual in + 94 MB out) = 36 us of HBM time next to 37 us of tensor time, it sits on
// the roofline ridge, not in the epilogue's dependent chain; FC1 + GELU 171 vs 160 us and the GPT-2 projections 4 - 14 %
// slower (the slab buffers cost ring stages, and a fence.proxy.async + bulk-group wait per 4 KB slab is more overhead than
// the transposition it replaces).  Kept as the measured answer to "TMA-load the residual, TMA-store the outputs".
// --------------------------------------------------------------------------------------------
template <int EPI, int MODE>
__device__ __forceinline__ void epilogue_v2(const GemmEpilogue& ep, const GemmTmaEpi& tme, const GemmSched& sched,
                                            const uint32_t wbase, const uint32_t rbar, const uint32_t tmem_base,
                                            const uint32_t tmem_full_bar0, const uint32_t tmem_empty_leader0, const int quad,
                                            const int half, const int lane, const uint32_t rank, const int cluster_id,
                                            const int num_clusters, const int num_tiles, const int num_m, const int num_n,
                                            const int M, const int N) {
  constexpr bool kRes = EPI >= 0 && (EPI & 4) != 0, kF32 = EPI >= 0 && (EPI & 8) != 0, kB16 = EPI >= 0 && (EPI & 16) != 0;
  constexpr uint32_t kB16Off = (kRes || kF32) ? 8192u : 0u;  // 16-bit slabs sit behind the two fp32 slabs when both exist
  auto tile_rc = [&](int tile, int& row0, int& col0) {
    const int m_idx = sched.n_fastest ? tile / num_n : tile % num_m;
    const int n_idx = sched.n_fastest ? tile % num_n : tile / num_m;
    row0 = m_idx * 256 + static_cast<int>(rank) * 128 + quad * 32;
    col0 = n_idx * 256 + half * 128;
  };
  // advance (tile, s) to the next slab that has rows < M and columns < N; false when this warp has none left
  auto next_live = [&](int& tile, int& sl, int& row, int& col) -> bool {
    for (;;) {
      if (++sl == 4) {
        sl = 0;
        tile += num_clusters;
      }
      if (tile >= num_tiles) return false;
      int r0, c0;
      tile_rc(tile, r0, c0);
      if (r0 < M && c0 + sl * 32 < N) {
        row = r0;
        col = c0 + sl * 32;
        return true;
      }
    }
  };
  auto issue_res = [&](uint32_t slab_no, int row, int col) {  // one lane: residual slab -> buffer slab_no & 1
    const uint32_t bar = rbar + (slab_no & 1u) * 8u;
    ptx::mbar_arrive_expect_tx(bar, 4096u);
    ptx::tma_load_2d(wbase + (slab_no & 1u) * 4096u, &tme.res, bar, col, row, ptx::kEvictNormal);
  };

  uint32_t slab = 0;  // live slabs processed so far by this warp: buffer = slab & 1, mbarrier parity = (slab >> 1) & 1
  int nt = cluster_id, ns = -1, nrow = 0, ncol = 0;  // the next live slab (prefetch cursor)
  bool have_next = next_live(nt, ns, nrow, ncol);
  if (kRes && have_next && lane == 0) issue_res(0, nrow, ncol);
  uint32_t acc = 0, acc_phase = 0;
  for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
    int row0, col0;
    tile_rc(tile, row0, col0);
    ptx::mbar_wait(tmem_full_bar0 + 8u * acc, acc_phase);
    ptx::tcgen05_fence_after();
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 256u + static_cast<uint32_t>(half * 128);

#pragma unroll 1
    for (int sl = 0; sl < 4; ++sl) {
      const int col = col0 + sl * 32;
      if (row0 >= M || col >= N) break;  // warp-uniform; later slabs of this tile are dead as well
      // this slab is the prefetch cursor's slab: move the cursor to the one after it
      have_next = next_live(nt, ns, nrow, ncol);
      const uint32_t fb = wbase + (slab & 1u) * 4096u;
      const uint32_t hb = wbase + kB16Off + (slab & 1u) * 2048u;
      uint32_t v[32];
      ptx::tmem_ld_32x32(t_row + static_cast<uint32_t>(sl * 32), v);
      if (!kRes) {
        if (lane == 0) ptx::bulk_wait_group_read<1>();  // the stores of slab - 2 have left this slab's buffers
        __syncwarp();
      }
      ptx::tmem_ld_wait();
      float x[32];
      const bool full_cols = (col + 32 <= N);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ep.bias != nullptr) {
          if (full_cols) {
            b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col) + j);
          } else {
            if (col + 4 * j + 0 < N) b4.x = __ldg(ep.bias + col + 4 * j + 0);
            if (col + 4 * j + 1 < N) b4.y = __ldg(ep.bias + col + 4 * j + 1);
            if (col + 4 * j + 2 < N) b4.z = __ldg(ep.bias + col + 4 * j + 2);
            if (col + 4 * j + 3 < N) b4.w = __ldg(ep.bias + col + 4 * j + 3);
          }
        }
        const float4 r4 = epilogue_math<EPI, MODE>(
            ep, make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                            __uint_as_float(v[4 * j + 3])), b4);
        x[4 * j] = r4.x;
        x[4 * j + 1] = r4.y;
        x[4 * j + 2] = r4.z;
        x[4 * j + 3] = r4.w;
      }
      if (kRes) {
        if (lane == 0) {
          // Every earlier store has finished reading shared memory: the other buffer (read by the store of slab - 1)
          // can take the residual of slab + 1, which then has this slab's whole processing time to arrive.
          ptx::bulk_wait_group_read<0>();
          if (have_next) issue_res(slab + 1u, nrow, ncol);
        }
        ptx::mbar_wait(rbar + (slab & 1u) * 8u, (slab >> 1) & 1u);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 r = lds128(fb + static_cast<uint32_t>(lane) * 128u + static_cast<uint32_t>((j ^ (lane & 7)) << 4));
          x[4 * j] += r.x;
          x[4 * j + 1] += r.y;
          x[4 * j + 2] += r.z;
          x[4 * j + 3] += r.w;
        }
      }
      if (kF32) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(fb + static_cast<uint32_t>(lane) * 128u + static_cast<uint32_t>((j ^ (lane & 7)) << 4), x[4 * j], x[4 * j + 1],
                 x[4 * j + 2], x[4 * j + 3]);
      }
      if (kB16) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          sts128u(hb + static_cast<uint32_t>(lane) * 64u + static_cast<uint32_t>((j ^ ((lane >> 1) & 3)) << 4),
                  pack2_operand<MODE>(x[8 * j], x[8 * j + 1]), pack2_operand<MODE>(x[8 * j + 2], x[8 * j + 3]),
                  pack2_operand<MODE>(x[8 * j + 4], x[8 * j + 5]), pack2_operand<MODE>(x[8 * j + 6], x[8 * j + 7]));
      }
      ptx::fence_proxy_async_smem();  // this lane's slab rows -> visible to the TMA engine
      __syncwarp();
      if (lane == 0) {
        if (kF32) ptx::tma_store_2d(&tme.f32, fb, col, row0);
        if (kB16) ptx::tma_store_2d(&tme.b16, hb, col, row0);
        ptx::bulk_commit_group();
      }
      ++slab;
    }
    // every tcgen05.ld of this tile has completed: hand the accumulator back to the MMA warp
    ptx::tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive_cluster(tmem_empty_leader0 + 8u * acc);
    acc ^= 1u;
    if (acc == 0) acc_phase ^= 1u;
  }
  if (lane == 0) ptx::bulk_wait_group<0>();  // all stores performed before the grid can be considered complete
  __syncwarp();
}

template <int MODE>
struct Gemm2Traits {
  static constexpr int kPairs = (MODE == MODE_BF16X3) ? 2 : 1;
  static constexpr uint32_t kABytes = kBlockM * kBlockK * 2;       // 128 rows of A
  static constexpr uint32_t kBBytes = 128 * kBlockK * 2;           // this CTA's half of the 256-row B tile
  static constexpr uint32_t kStageBytes = kPairs * (kABytes + kBBytes);
#ifndef AFFT_2CTA_STAGES
#define AFFT_2CTA_STAGES 6  // A/B knob (tools/gemm_time.py with AFFT_B200_LIB): depth of the TMA -> MMA ring
#endif
  static constexpr int kStages = (MODE == MODE_BF16X3) ? 3 : AFFT_2CTA_STAGES;  // default ring depth (v1 epilogue)
  static constexpr int kMaxStages = 8;
  static constexpr uint32_t kTmemCols = 512;                       // 2 accumulators x 256 columns
  static constexpr uint32_t kBarrierBytes = 1024;                  // barriers first: the ring behind them stays 1024-B aligned
  static constexpr uint32_t kStagingBytes = kNumEpilogueWarps * 4096;  // v1 epilogue: one transposition tile per warp
  // shared memory of a launch with `stages` ring slots and `epi_bytes` of epilogue staging (+ alignment slack)
  static constexpr uint32_t smem_bytes(int stages, uint32_t epi_bytes) {
    return kBarrierBytes + static_cast<uint32_t>(stages) * kStageBytes + epi_bytes + 1024;
  }
  static constexpr uint32_t kSmemBytes = smem_bytes(kStages, kStagingBytes);
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
};

template <int MODE, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                              const __grid_constant__ CUtensorMap tm_a_lo,
                              const __grid_constant__ CUtensorMap tm_b_lo, const __grid_constant__ GemmEpilogue ep, const int M,
                              const int N, const int K, const __grid_constant__ GemmSched sched,
                              const __grid_constant__ GemmTmaEpi tme) {
  using T = Gemm2Traits<MODE>;
  constexpr int BLOCK_N = 256;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_raw_u32 = ptx::smem_u32(smem_raw);
  const uint32_t bar_base = (smem_raw_u32 + 1023u) & ~1023u;   // [barriers 1 KB][ring: stages x kStageBytes][epilogue staging]
  const uint32_t smem_base = bar_base + T::kBarrierBytes;
  const uint32_t num_stages = static_cast<uint32_t>(sched.stages);
  const uint32_t staging_base = smem_base + num_stages * T::kStageBytes;
  auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (T::kMaxStages + s); };
  auto tmem_full_bar = [&](uint32_t a) { return bar_base + 8u * (2 * T::kMaxStages + a); };
  auto tmem_empty_bar = [&](uint32_t a) { return bar_base + 8u * (2 * T::kMaxStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * T::kMaxStages + 4);
  auto res_bar = [&](uint32_t w) { return bar_base + 8u * (2 * T::kMaxStages + 6 + 2 * w); };  // v2 epilogue: 2 per warp
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_raw_u32));

  ptx::griddep_launch();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();  // 0 = leader
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int num_m = (M + 255) / 256;
  const int num_n = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + kBlockK - 1) / kBlockK;
  const int ksplit = sched.ksplit;  // work unit = (tile, split), see GemmSched
  const int num_units = num_tiles * ksplit;
  const int kb_per = (num_kb + ksplit - 1) / ksplit;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_a);
    ptx::prefetch_tensormap(&tm_b);
    if (MODE == MODE_BF16X3) {
      ptx::prefetch_tensormap(&tm_a_lo);
      ptx::prefetch_tensormap(&tm_b_lo);
    }
  }
  if (warp == 1 && lane == 0) {
    for (uint32_t s = 0; s < num_stages; ++s) {
      ptx::mbar_init(full_bar(s), 2);   // leader's arrive.expect_tx + the peer producer's remote arrive
      ptx::mbar_init(empty_bar(s), 1);  // tcgen05.commit (multicast from the leader)
    }
    for (uint32_t a = 0; a < 2; ++a) {
      ptx::mbar_init(tmem_full_bar(a), 1);                        // tcgen05.commit (multicast)
      ptx::mbar_init(tmem_empty_bar(a), 2 * kNumEpilogueWarps);   // epilogue warps of both CTAs
    }
    for (uint32_t w = 0; w < 2 * kNumEpilogueWarps; ++w) ptx::mbar_init(res_bar(0) + 8u * w, 1);  // residual slabs (v2)
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc_2cta<T::kTmemCols>(tmem_slot);
  ptx::tcgen05_fence_before();
  ptx::cluster_sync();  // barriers of both CTAs initialised before any remote arrive / multicast commit
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  ptx::griddep_wait();  // operands / residual written by the previous kernel are visible from here on

  if (warp == 0) {
    // ======================= TMA producer (both CTAs) =======================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int unit = cluster_id; unit < num_units; unit += num_clusters) {
        const int tile = unit / ksplit, split = unit - tile * ksplit;
        const int m_idx = sched.n_fastest ? tile / num_n : tile % num_m;
        const int n_idx = sched.n_fastest ? tile % num_n : tile / num_m;
        const int a_row = m_idx * 256 + static_cast<int>(rank) * 128;
        const int b_row = n_idx * BLOCK_N + static_cast<int>(rank) * 128;
        const int kb0 = split * kb_per, kb1 = min(num_kb, kb0 + kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t a_dst = smem_base + stage * T::kStageBytes;
          const uint32_t b_dst = a_dst + T::kABytes;
          const uint32_t fb_leader = ptx::mapa(full_bar(stage), 0);
          if (rank == 0) ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * T::kStageBytes);
          else ptx::mbar_arrive_cluster(fb_leader);
          ptx::tma_load_2d_2cta(a_dst, &tm_a, fb_leader, kb * kBlockK, a_row, sched.policy_a);
          ptx::tma_load_2d_2cta(b_dst, &tm_b, fb_leader, kb * kBlockK, b_row, sched.policy_b);
          if (MODE == MODE_BF16X3) {
            const uint32_t a_lo_dst = b_dst + T::kBBytes;
            const uint32_t b_lo_dst = a_lo_dst + T::kABytes;
            ptx::tma_load_2d_2cta(a_lo_dst, &tm_a_lo, fb_leader, kb * kBlockK, a_row, sched.policy_a);
            ptx::tma_load_2d_2cta(b_lo_dst, &tm_b_lo, fb_leader, kb * kBlockK, b_row, sched.policy_b);
          }
          if (++stage == num_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (leader CTA only) =======================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16_f32(256, BLOCK_N, MODE == MODE_FP16);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int unit = cluster_id; unit < num_units; unit += num_clusters) {
        const int split = unit % ksplit;
        const int kb0 = split * kb_per, kb1 = min(num_kb, kb0 + kb_per);
        ptx::mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1u);
        ptx::tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tcgen05_fence_after();
          const uint32_t a_src = smem_base + stage * T::kStageBytes;
          const uint32_t b_src = a_src + T::kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            ptx::umma_bf16_ss_2cta(d_tmem, ptx::make_smem_desc_sw128(a_src + k * (kUmmaK * 2)),
                                   ptx::make_smem_desc_sw128(b_src + k * (kUmmaK * 2)), idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
          }
          if (MODE == MODE_BF16X3) {
            const uint32_t a_lo_src = b_src + T::kBBytes;
            const uint32_t b_lo_src = a_lo_src + T::kABytes;
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              ptx::umma_bf16_ss_2cta(d_tmem, ptx::make_smem_desc_sw128(a_src + k * (kUmmaK * 2)),
                                     ptx::make_smem_desc_sw128(b_lo_src + k * (kUmmaK * 2)), idesc, 1u);
            }
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              ptx::umma_bf16_ss_2cta(d_tmem, ptx::make_smem_desc_sw128(a_lo_src + k * (kUmmaK * 2)),
                                     ptx::make_smem_desc_sw128(b_src + k * (kUmmaK * 2)), idesc, 1u);
            }
          }
          ptx::umma_commit_2cta(empty_bar(stage), 0b11);  // frees the slot in both CTAs
          if (++stage == num_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        ptx::umma_commit_2cta(tmem_full_bar(acc), 0b11);  // accumulators of both CTAs complete
        acc ^= 1u;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ======================= epilogue (both CTAs, own 128 rows) =======================
    const int quad = warp & 3;
    const int egrp = (warp - 4) >> 2;
    bool v2 = false;
    if constexpr (EPI >= 0 && MODE != MODE_BF16X3) {  // TMA-staged epilogue (the launcher guarantees ksplit == 1)
      v2 = sched.v2_warp_bytes > 0;
      if (v2)
        epilogue_v2<EPI, MODE>(ep, tme, sched,
                               staging_base + static_cast<uint32_t>(warp - 4) * static_cast<uint32_t>(sched.v2_warp_bytes),
                               res_bar(static_cast<uint32_t>(warp - 4)), tmem_base, tmem_full_bar(0), ptx::mapa(tmem_empty_bar(0), 0),
                               quad, egrp, lane, rank, cluster_id, num_clusters, num_tiles, num_m, num_n, M, N);
    }
    if (!v2) {
    const uint32_t stage = staging_base + static_cast<uint32_t>(warp - 4) * 4096u;
    constexpr size_t kHalfTile = static_cast<size_t>(kBlockM) * BLOCK_N;  // one CTA's 128 x 256 accumulator
    uint32_t acc = 0, acc_phase = 0;
    for (int unit = cluster_id; unit < num_units; unit += num_clusters) {
      const int tile = unit / ksplit;
      const int m_idx = sched.n_fastest ? tile / num_n : tile % num_m;
      const int n_idx = sched.n_fastest ? tile % num_n : tile / num_m;
      const int row0 = m_idx * 256 + static_cast<int>(rank) * 128 + quad * 32;
      // residual of the NEXT unit of this CTA pair -> L2 (and of this one, for the pair's first unit)
      if (unit == cluster_id && ksplit == 1)
        epilogue_prefetch_residual<EPI, MODE, BLOCK_N>(ep, row0, n_idx * BLOCK_N, egrp, lane, M, N);
      int next_row0 = -1, next_n0 = 0;
      if (unit + num_clusters < num_units) {
        const int nt = (unit + num_clusters) / ksplit;
        next_row0 = (sched.n_fastest ? nt / num_n : nt % num_m) * 256 + static_cast<int>(rank) * 128 + quad * 32;
        next_n0 = (sched.n_fastest ? nt % num_n : nt / num_m) * BLOCK_N;
      }
      ptx::mbar_wait(tmem_full_bar(acc), acc_phase);
      ptx::tcgen05_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N;
      const uint32_t empty_leader = ptx::mapa(tmem_empty_bar(acc), 0);
      epilogue_unit<EPI, MODE, BLOCK_N>(ep, ksplit, stage, t_row, quad, egrp, lane, row0, n_idx * BLOCK_N, M, N,
                                         sched.partials + (static_cast<size_t>(unit) * 2 + rank) * kHalfTile,
                                         sched.partials + (static_cast<size_t>(tile) * ksplit * 2 + rank) * kHalfTile, 2 * kHalfTile,
                                         sched.counters + (tile * 2 + rank) * kNumEpilogueWarps + (warp - 4), next_row0, next_n0,
                                         [&] { if (lane == 0) ptx::mbar_arrive_cluster(empty_leader); });
      acc ^= 1u;
      if (acc == 0) acc_phase ^= 1u;
    }
    }
  }

  ptx::tcgen05_fence_before();
  ptx::cluster_sync();  // the peer may still be arriving on / reading this CTA's shared memory and TMEM
  if (warp == 2) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc_2cta<T::kTmemCols>(tmem_base);
  }
  if (warp == 0) ptx::prof_mark_end(sched.t_end);
}

}  // namespace afft
