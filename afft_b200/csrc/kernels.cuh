// The non-GEMM kernels of the AFFT hot path.  All of them are HBM/L2-bandwidth or latency bound
// (SURVEY.md section 8d), so they are plain SIMT kernels: 16-byte coalesced accesses, fp32 math,
// warp-shuffle reductions, operands staged in shared memory where they are re-read.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "ptx_sm100.cuh"

namespace afft {

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {  // a -> low half; one F2FP instruction
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// fp16 operand pair, saturated to the finite range (MODE_FP16: activations that feed the next GEMM)
__device__ __forceinline__ uint32_t pack_f16x2_sat(float a, float b) {  // a -> low half; one F2FP.SATFINITE instruction
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// the 16-bit GEMM operand format of this launch: bf16, or fp16 when `fp16` (warp-uniform)
__device__ __forceinline__ uint32_t pack_op16x2(float a, float b, bool fp16) {
  return fp16 ? pack_f16x2_sat(a, b) : pack_bf16x2(a, b);
}
__device__ __forceinline__ void unpack_f16x8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
// residual part of the hi/lo split: bf16(a - float(bf16(a)))
__device__ __forceinline__ float bf16_residual(float a) {
  return a - __bfloat162float(__float2bfloat16_rn(a));
}
__device__ __forceinline__ void unpack_bf16x8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 -> bf16 (hi [+ lo]) conversion, optionally transposing.  Used once per weight at load time
// (Conv1D [in,out] -> K-major [out,in]) and per forward for feature inputs that feed a projection.
// ------------------------------------------------------------------------------------------------
__global__ void convert_f32_bf16_kernel(const float* __restrict__ src, long long lds, int rows, int cols,
                                        __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                        long long ldd, int fp16, unsigned long long* t_end = nullptr) {
  // fast path: 8 elements per thread (two 16-byte loads, one 16-byte store per output) when rows are 8-element
  // multiples and every pitch / base is 16-byte aligned - weights (per training step) and most activations
  const bool vec = (cols % 8 == 0) && (lds % 4 == 0) && (ldd % 8 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) &&
                   ((reinterpret_cast<uintptr_t>(hi) & 15u) == 0) && (lo == nullptr || (reinterpret_cast<uintptr_t>(lo) & 15u) == 0);
  if (vec) {
    const int c8 = cols / 8;
    const long long total8 = static_cast<long long>(rows) * c8;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total8;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const long long r = i / c8;
      const int c = static_cast<int>(i - r * c8) * 8;
      const float4 v0 = *reinterpret_cast<const float4*>(src + r * lds + c);
      const float4 v1 = *reinterpret_cast<const float4*>(src + r * lds + c + 4);
      const uint4 h = make_uint4(pack_op16x2(v0.x, v0.y, fp16), pack_op16x2(v0.z, v0.w, fp16), pack_op16x2(v1.x, v1.y, fp16),
                                 pack_op16x2(v1.z, v1.w, fp16));
      *reinterpret_cast<uint4*>(hi + r * ldd + c) = h;
      if (lo != nullptr)
        *reinterpret_cast<uint4*>(lo + r * ldd + c) =
            make_uint4(pack_bf16x2(bf16_residual(v0.x), bf16_residual(v0.y)), pack_bf16x2(bf16_residual(v0.z), bf16_residual(v0.w)),
                       pack_bf16x2(bf16_residual(v1.x), bf16_residual(v1.y)), pack_bf16x2(bf16_residual(v1.z), bf16_residual(v1.w)));
    }
    ptx::prof_mark_end(t_end);
    return;
  }
  const long long total = static_cast<long long>(rows) * cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / cols), c = static_cast<int>(i % cols);
    const float v = src[r * lds + c];
    if (fp16) {
      reinterpret_cast<__half*>(hi)[r * ldd + c] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
      continue;
    }
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[r * ldd + c] = h;
    if (lo != nullptr) lo[r * ldd + c] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
  ptx::prof_mark_end(t_end);
}

// dst[c, r] = src[r, c]   (src [rows, cols] pitch lds; dst [cols, rows] pitch ldd)
__global__ void convert_transpose_f32_bf16_kernel(const float* __restrict__ src, long long lds, int rows,
                                                  int cols, __nv_bfloat16* __restrict__ hi,
                                                  __nv_bfloat16* __restrict__ lo, long long ldd, int fp16) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? src[r * lds + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (c < cols && r < rows) {
      const float v = tile[threadIdx.x][j];
      if (fp16) {
        reinterpret_cast<__half*>(hi)[c * ldd + r] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
        continue;
      }
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      hi[c * ldd + r] = h;
      if (lo != nullptr) lo[c * ldd + r] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the last dim (biased variance, affine), one warp per row, row in registers.
// (reference: nn.LayerNorm eps=1e-6 in models/fusion.py:281 / transformerblock.py:132,134,
//  eps=1e-5 in GPT-2 ln_1/ln_2/ln_f)
// Outputs (each optional): y_f32, y_hi (bf16), y_lo (bf16 residual) at row r, pitch ldy; and an
// "aux" copy of rows with r % aux_mod == 0 to row (r / aux_mod) * aux_stride of aux_f32 / aux_hi /
// aux_lo (pitch ld_aux) - used to drop z[:, 0] into slot 0 of the past_futures buffer.
// ------------------------------------------------------------------------------------------------
struct LayerNormArgs {
  const float* x;
  long long ldx;
  int in_group, in_stride;  // input row of output row r: (r / in_group) * in_stride + r % in_group (0: r)
  int n_avg, avg_stride;    // output = mean over s < n_avg of LN(x[in_row + s * avg_stride])  (n_avg <= 1: plain LN)
  const float* gamma;
  const float* beta;
  float eps;
  int rows, dim;
  float* y_f32;
  __nv_bfloat16* y_hi;
  __nv_bfloat16* y_lo;
  long long ldy;
  int aux_mod, aux_stride, aux_rem;  // rows with r % aux_mod == aux_rem are also written to aux row (r / aux_mod) * aux_stride
  float* aux_f32;
  __nv_bfloat16* aux_hi;
  __nv_bfloat16* aux_lo;
  long long ld_aux;
  int out_fp16;  // y_hi / aux_hi are fp16 (MODE_FP16) instead of bf16
  unsigned long long* t_end;  // profiling slot or nullptr
  int out_group, out_stride, out_off;  // output row of row r: (r / out_group) * out_stride + r % out_group + out_off (0: r)
};

__device__ __forceinline__ void ln_store(const LayerNormArgs& a, long long row, bool aux, long long arow, int c4, const float4& y) {
  const uint2 hi = make_uint2(pack_op16x2(y.x, y.y, a.out_fp16), pack_op16x2(y.z, y.w, a.out_fp16));
  if (a.y_f32 != nullptr) reinterpret_cast<float4*>(a.y_f32 + row * a.ldy)[c4] = y;
  if (a.y_hi != nullptr) reinterpret_cast<uint2*>(a.y_hi + row * a.ldy)[c4] = hi;
  if (a.y_lo != nullptr)
    reinterpret_cast<uint2*>(a.y_lo + row * a.ldy)[c4] =
        make_uint2(pack_bf16x2(bf16_residual(y.x), bf16_residual(y.y)), pack_bf16x2(bf16_residual(y.z), bf16_residual(y.w)));
  if (aux) {
    if (a.aux_f32 != nullptr) reinterpret_cast<float4*>(a.aux_f32 + arow * a.ld_aux)[c4] = y;
    if (a.aux_hi != nullptr) reinterpret_cast<uint2*>(a.aux_hi + arow * a.ld_aux)[c4] = hi;
    if (a.aux_lo != nullptr)
      reinterpret_cast<uint2*>(a.aux_lo + arow * a.ld_aux)[c4] =
          make_uint2(pack_bf16x2(bf16_residual(y.x), bf16_residual(y.y)), pack_bf16x2(bf16_residual(y.z), bf16_residual(y.w)));
  }
}

// NV float4 per lane: dim = 128 * NV.  AVG = false: plain LayerNorm of one row per warp (the hot case: the row
// lives in NV float4 registers, one 16-B load per lane per 512 B, all loads in flight before the reductions).
// AVG = true: mean over n_avg normalised rows (CMFuser / T-SA-Fuser without frame-level token).
// FAST = true is the hot instantiation (bf16 output only, affine, no aux scatter): no per-vector null checks.
template <int NV, bool AVG, bool FAST = false>
__global__ void __launch_bounds__(256) layernorm_kernel(const LayerNormArgs a) {
  ptx::griddep_launch();
  ptx::griddep_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= a.rows) return;
  long long irow = warp;
  if (a.in_group > 0) irow = static_cast<long long>(warp / a.in_group) * a.in_stride + (warp % a.in_group);
  const float inv_d = 1.0f / static_cast<float>(a.dim);
  const float4* g4 = reinterpret_cast<const float4*>(a.gamma);
  const float4* b4 = reinterpret_cast<const float4*>(a.beta);
  const bool aux = a.aux_mod > 0 && (warp % a.aux_mod) == a.aux_rem;
  const long long arow = aux ? static_cast<long long>(warp / a.aux_mod) * a.aux_stride : 0;
  long long orow = warp;  // output row (row map: the Identity dim_decoder writes z_hat[b, t] to slot t + 1 of past_futures)
  if (a.out_group > 0) orow = static_cast<long long>(warp / a.out_group) * a.out_stride + (warp % a.out_group) + a.out_off;
  const int n_avg = (AVG && a.n_avg > 1) ? a.n_avg : 1;
  float4 acc[AVG ? NV : 1];
  if (AVG) {
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[AVG ? i : 0] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int sl = 0; sl < n_avg; ++sl) {
    const float4* xr = reinterpret_cast<const float4*>(a.x + (irow + static_cast<long long>(sl) * a.avg_stride) * a.ldx);
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = __ldg(xr + lane + 32 * i);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
      q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_d + a.eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = lane + 32 * i;
      float4 g = make_float4(1.f, 1.f, 1.f, 1.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (FAST || a.gamma != nullptr) g = __ldg(g4 + c4);
      if (FAST || a.beta != nullptr) b = __ldg(b4 + c4);
      float4 y;
      y.x = fmaf((v[i].x - mean) * rstd, g.x, b.x);
      y.y = fmaf((v[i].y - mean) * rstd, g.y, b.y);
      y.z = fmaf((v[i].z - mean) * rstd, g.z, b.z);
      y.w = fmaf((v[i].w - mean) * rstd, g.w, b.w);
      if (FAST) {
        reinterpret_cast<uint2*>(a.y_hi + orow * a.ldy)[c4] =
            make_uint2(pack_op16x2(y.x, y.y, a.out_fp16), pack_op16x2(y.z, y.w, a.out_fp16));
      } else if (AVG) {
        float4& t = acc[AVG ? i : 0];
        t.x += y.x; t.y += y.y; t.z += y.z; t.w += y.w;
      } else {
        ln_store(a, orow, aux, arow, c4, y);
      }
    }
  }
  if (AVG) {
    const float inv_avg = 1.0f / static_cast<float>(n_avg);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 y = acc[AVG ? i : 0];
      y.x *= inv_avg; y.y *= inv_avg; y.z *= inv_avg; y.w *= inv_avg;
      ln_store(a, orow, aux, arow, lane + 32 * i, y);
    }
  }
  ptx::prof_mark_end(a.t_end);
}

// ------------------------------------------------------------------------------------------------
// Token assembly: builds the fp32 residual stream of a fuser from the per-modality feature
// sequences (reference: models/fusion.py:338-353 SA-Fuser, :177-196 T-SA-Fuser, :262 CA-Fuser).
//   slot s of (b, t) goes to row  b*n_slots*T + (layout == 0 ? t*n_slots + s : s*T + t)
//   value = src_s[b, t, :]            (src_s == nullptr: slot is filled by a projection GEMM, skip)
//         | token[(t % tok_mod), :]   (learned modality-agnostic token; tok_mod = 1 or T)
//         + pos_emb[t, :] (optional) + mod_emb[s, :] (optional)
// ------------------------------------------------------------------------------------------------
constexpr int kMaxSlots = 8;
struct AssembleArgs {
  float* h;           // [B * n_slots * T, dim]
  int B, T, dim, n_slots, layout;
  const float* src[kMaxSlots];  // [B*T, dim] contiguous, or nullptr
  int is_token[kMaxSlots];      // slot reads `token` instead of src
  const float* token;           // [tok_mod, dim]
  int tok_mod;
  const float* pos_emb;  // [>=T, dim] or nullptr
  const float* mod_emb;  // [n_slots, dim] or nullptr
  unsigned long long* t_end;  // profiling slot or nullptr
};

__global__ void __launch_bounds__(256) assemble_tokens_kernel(const AssembleArgs a) {
  // One CTA per token row (grid-stride over rows), threads across the row's 16-byte vectors: the (b, t, slot)
  // decomposition is three 32-bit divisions per ROW (the first version did 64-bit divisions per element and ran at
  // 2 TB/s).
  const int d4 = a.dim / 4;
  const int rows = a.B * a.T * a.n_slots;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const int s = r % a.n_slots;
    const int bt = r / a.n_slots;
    const int t = bt % a.T;
    const int b = bt / a.T;
    const float4* src;
    if (a.is_token[s]) {
      src = reinterpret_cast<const float4*>(a.token + static_cast<long long>(t % a.tok_mod) * a.dim);
    } else if (a.src[s] != nullptr) {
      src = reinterpret_cast<const float4*>(a.src[s] + static_cast<long long>(bt) * a.dim);
    } else {
      continue;  // slot written by a projection GEMM
    }
    const float4* pe = a.pos_emb != nullptr ? reinterpret_cast<const float4*>(a.pos_emb + static_cast<long long>(t) * a.dim) : nullptr;
    const float4* me = a.mod_emb != nullptr ? reinterpret_cast<const float4*>(a.mod_emb + static_cast<long long>(s) * a.dim) : nullptr;
    const long long orow = static_cast<long long>(b) * a.n_slots * a.T +
                           (a.layout == 0 ? static_cast<long long>(t) * a.n_slots + s
                                          : static_cast<long long>(s) * a.T + t);
    float4* dst = reinterpret_cast<float4*>(a.h + orow * a.dim);
    for (int c4 = threadIdx.x; c4 < d4; c4 += blockDim.x) {
      float4 v = __ldg(src + c4);
      if (pe != nullptr) {
        const float4 p = __ldg(pe + c4);
        v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
      }
      if (me != nullptr) {
        const float4 p = __ldg(me + c4);
        v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
      }
      dst[c4] = v;
    }
  }
  ptx::prof_mark_end(a.t_end);
}

// table[t, :] = pos_emb[t, :] + (mod_emb ? mod_emb[:] : 0)   for t < T - the additive term a projected
// modality receives through the GEMM residual input (T-SA-Fuser / CA-Fuser embeddings).
__global__ void embed_table_kernel(float* __restrict__ table, const float* __restrict__ pos_emb,
                                   const float* __restrict__ mod_emb, int T, int dim, unsigned long long* t_end = nullptr) {
  const int total = T * dim;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    float v = (pos_emb != nullptr) ? pos_emb[i] : 0.f;
    if (mod_emb != nullptr) v += mod_emb[i % dim];
    table[i] = v;
  }
  ptx::prof_mark_end(t_end);
}

// ------------------------------------------------------------------------------------------------
// Small multi-head attention: one CTA per (sequence, head); L <= 64 keys, head_dim 256 or 512.
// K and V of the head are staged in shared memory once and re-read by every query row; each warp
// owns query rows i = warp, warp + NW, ...; lanes split head_dim in 16-B chunks; scores are
// reduced with warp shuffles; softmax in fp32.
// (reference: models/transformerblock.py:24-33 and :64-74; transformers GPT-2 eager attention)
//   mask 0: none | 1: causal (j <= i) | 2: block-causal, period T ((j % T) <= (i % T))
//        | 3: diagonal masked (j != i)   [ModalTokenCMFuser cross_attn=True]
// q/k/v element (seq, i, h, d) is at  base[(seq*L + i) * ld + h*HD + d].
// probs (optional, fp32): probs[(seq / p_inner) * p_outer + (seq % p_inner) * p_inner_stride
//                               + h*L*L + i*L + j]
// ------------------------------------------------------------------------------------------------
struct AttentionArgs {
  const void* q;
  const void* k;
  const void* v;
  long long ldq, ldk, ldv;
  int n_seq, L, H;
  float scale;
  int mask, T;
  __nv_bfloat16* out_hi;
  __nv_bfloat16* out_lo;
  long long ldo;
  float* probs;
  long long p_outer, p_inner_stride;
  int p_inner;
  unsigned long long* t_end;  // profiling slot or nullptr
  // training (attention-probability dropout, reference models/transformerblock.py:31,71 and GPT-2 attn_pdrop): optional
  // [n_seq, H, L, L] fp32 factors (0 or 1 / (1 - p)) applied to the softmax before P.V; `probs` keeps the undropped
  // softmax (the backward needs it).  SIMT kernels only.
  const float* drop;
};

template <typename TIn, int HD>
struct AttnTraits {
  static constexpr int kVecElems = 16 / sizeof(TIn);        // elements per 16-B vector
  static constexpr int kChunkElems = 32 * kVecElems;        // elements covered by one warp-wide vector load
  static constexpr int kNV = HD / kChunkElems;              // vectors per lane
  static constexpr int kPerLane = HD / 32;                  // elements per lane
};

template <typename TIn, int HD>
__device__ __forceinline__ void load_row_regs(const TIn* row, int lane, float (&f)[HD / 32]) {
  using A = AttnTraits<TIn, HD>;
#pragma unroll
  for (int c = 0; c < A::kNV; ++c) {
    const uint4 u = *reinterpret_cast<const uint4*>(row + c * A::kChunkElems + lane * A::kVecElems);
    if constexpr (sizeof(TIn) == 2) {
      float t[8];
      if constexpr (std::is_same<TIn, __half>::value) unpack_f16x8(u, t);
      else unpack_bf16x8(u, t);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[c * 8 + e] = t[e];
    } else {
      f[c * 4 + 0] = __uint_as_float(u.x);
      f[c * 4 + 1] = __uint_as_float(u.y);
      f[c * 4 + 2] = __uint_as_float(u.z);
      f[c * 4 + 3] = __uint_as_float(u.w);
    }
  }
}

template <typename TIn, int HD>
__global__ void __launch_bounds__(256) attention_small_kernel(const AttentionArgs a) {
  using A = AttnTraits<TIn, HD>;
  extern __shared__ uint4 smem_attn[];
  TIn* ks = reinterpret_cast<TIn*>(smem_attn);
  TIn* vs = ks + static_cast<size_t>(a.L) * HD;
  const int seq = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int L = a.L;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const TIn* qg = reinterpret_cast<const TIn*>(a.q) + static_cast<long long>(seq) * L * a.ldq + h * HD;
  const TIn* kg = reinterpret_cast<const TIn*>(a.k) + static_cast<long long>(seq) * L * a.ldk + h * HD;
  const TIn* vg = reinterpret_cast<const TIn*>(a.v) + static_cast<long long>(seq) * L * a.ldv + h * HD;

  // stage K, V (16-B vectors, coalesced)
  constexpr int kVecPerRow = HD / A::kVecElems;
  for (int i = threadIdx.x; i < L * kVecPerRow; i += blockDim.x) {
    const int r = i / kVecPerRow, c = i % kVecPerRow;
    reinterpret_cast<uint4*>(ks)[i] = *reinterpret_cast<const uint4*>(kg + r * a.ldk + c * A::kVecElems);
    reinterpret_cast<uint4*>(vs)[i] = *reinterpret_cast<const uint4*>(vg + r * a.ldv + c * A::kVecElems);
  }
  __syncthreads();

  for (int i = warp; i < L; i += nw) {
    float q[A::kPerLane];
    load_row_regs<TIn, HD>(qg + static_cast<long long>(i) * a.ldq, lane, q);
    // scores: lane (j % 32) keeps s_j in slot j / 32
    float s0 = -INFINITY, s1 = -INFINITY;
    for (int j = 0; j < L; ++j) {
      bool ok = true;
      if (a.mask == 1) ok = (j <= i);
      else if (a.mask == 2) ok = ((j % a.T) <= (i % a.T));
      else if (a.mask == 3) ok = (j != i);
      if (!ok) continue;  // warp-uniform
      float kr[A::kPerLane];
      load_row_regs<TIn, HD>(ks + static_cast<size_t>(j) * HD, lane, kr);
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < A::kPerLane; ++e) d = fmaf(q[e], kr[e], d);
      d = warp_sum(d) * a.scale;
      if (lane == (j & 31)) {
        if (j < 32) s0 = d; else s1 = d;
      }
    }
    const float mx = warp_max(fmaxf(s0, s1));
    float p0 = (s0 == -INFINITY) ? 0.f : expf(s0 - mx);
    float p1 = (s1 == -INFINITY) ? 0.f : expf(s1 - mx);
    const float inv = 1.0f / warp_sum(p0 + p1);
    p0 *= inv;
    p1 *= inv;
    if (a.probs != nullptr) {
      float* pr = a.probs + static_cast<long long>(seq / a.p_inner) * a.p_outer +
                  static_cast<long long>(seq % a.p_inner) * a.p_inner_stride +
                  (static_cast<long long>(h) * L + i) * L;
      if (lane < L) pr[lane] = p0;
      if (lane + 32 < L) pr[lane + 32] = p1;
    }
    if (a.drop != nullptr) {
      const float* dm = a.drop + ((static_cast<long long>(seq) * a.H + h) * L + i) * L;
      if (lane < L) p0 *= dm[lane];
      if (lane + 32 < L) p1 *= dm[lane + 32];
    }
    float o[A::kPerLane];
#pragma unroll
    for (int e = 0; e < A::kPerLane; ++e) o[e] = 0.f;
    for (int j = 0; j < L; ++j) {
      const float pj = __shfl_sync(0xffffffffu, (j < 32) ? p0 : p1, j & 31);
      if (pj == 0.f) continue;  // warp-uniform (masked keys)
      float vr[A::kPerLane];
      load_row_regs<TIn, HD>(vs + static_cast<size_t>(j) * HD, lane, vr);
#pragma unroll
      for (int e = 0; e < A::kPerLane; ++e) o[e] = fmaf(pj, vr[e], o[e]);
    }
    // write merged-head output row: element d of lane = chunk c, lane*kVecElems + e
    const long long orow = (static_cast<long long>(seq) * L + i) * a.ldo + h * HD;
#pragma unroll
    for (int c = 0; c < A::kNV; ++c) {
      constexpr int VE = A::kVecElems;
      const int d0 = c * A::kChunkElems + lane * VE;
      if constexpr (VE == 8) {
        const float* oo = &o[c * 8];
        constexpr bool kF16 = std::is_same<TIn, __half>::value;  // 16-bit inputs: the output keeps their format
        uint4 hv = make_uint4(pack_op16x2(oo[0], oo[1], kF16), pack_op16x2(oo[2], oo[3], kF16), pack_op16x2(oo[4], oo[5], kF16),
                              pack_op16x2(oo[6], oo[7], kF16));
        *reinterpret_cast<uint4*>(a.out_hi + orow + d0) = hv;
        if (a.out_lo != nullptr) {
          uint4 lv = make_uint4(pack_bf16x2(bf16_residual(oo[0]), bf16_residual(oo[1])),
                                pack_bf16x2(bf16_residual(oo[2]), bf16_residual(oo[3])),
                                pack_bf16x2(bf16_residual(oo[4]), bf16_residual(oo[5])),
                                pack_bf16x2(bf16_residual(oo[6]), bf16_residual(oo[7])));
          *reinterpret_cast<uint4*>(a.out_lo + orow + d0) = lv;
        }
      } else {
        const float* oo = &o[c * 4];
        uint2 hv = make_uint2(pack_bf16x2(oo[0], oo[1]), pack_bf16x2(oo[2], oo[3]));
        *reinterpret_cast<uint2*>(a.out_hi + orow + d0) = hv;
        if (a.out_lo != nullptr) {
          uint2 lv = make_uint2(pack_bf16x2(bf16_residual(oo[0]), bf16_residual(oo[1])),
                                pack_bf16x2(bf16_residual(oo[2]), bf16_residual(oo[3])));
          *reinterpret_cast<uint2*>(a.out_lo + orow + d0) = lv;
        }
      }
    }
  }
  ptx::prof_mark_end(a.t_end);
}


// ------------------------------------------------------------------------------------------------
// SA-Fuser attention: L <= 6 modality tokens per timestep, head_dim 256 (reference models/fusion.py:358-360 ->
// transformerblock.py:24-33).  One warp per (timestep, head), everything in registers: the 3 L rows of the head
// (q, k, v) are fetched with 16-byte loads that are all in flight before the first use, the L x L scores are
// reduced with warp shuffles, softmax in fp32, P.V accumulated in registers.  No shared memory, no block sync.
// ------------------------------------------------------------------------------------------------
template <typename TIn, int HD>
__device__ __forceinline__ void store_row_regs(__nv_bfloat16* out_hi, __nv_bfloat16* out_lo, long long off, int lane,
                                               const float (&o)[HD / 32]) {
  using A = AttnTraits<TIn, HD>;
#pragma unroll
  for (int c = 0; c < A::kNV; ++c) {
    constexpr int VE = A::kVecElems;
    const int d0 = c * A::kChunkElems + lane * VE;
    if constexpr (VE == 8) {
      const float* oo = &o[c * 8];
      constexpr bool kF16 = std::is_same<TIn, __half>::value;  // 16-bit inputs: the output keeps their format
      *reinterpret_cast<uint4*>(out_hi + off + d0) = make_uint4(pack_op16x2(oo[0], oo[1], kF16), pack_op16x2(oo[2], oo[3], kF16),
                                                                 pack_op16x2(oo[4], oo[5], kF16), pack_op16x2(oo[6], oo[7], kF16));
      if (out_lo != nullptr)
        *reinterpret_cast<uint4*>(out_lo + off + d0) =
            make_uint4(pack_bf16x2(bf16_residual(oo[0]), bf16_residual(oo[1])), pack_bf16x2(bf16_residual(oo[2]), bf16_residual(oo[3])),
                       pack_bf16x2(bf16_residual(oo[4]), bf16_residual(oo[5])), pack_bf16x2(bf16_residual(oo[6]), bf16_residual(oo[7])));
    } else {
      const float* oo = &o[c * 4];
      *reinterpret_cast<uint2*>(out_hi + off + d0) = make_uint2(pack_bf16x2(oo[0], oo[1]), pack_bf16x2(oo[2], oo[3]));
      if (out_lo != nullptr)
        *reinterpret_cast<uint2*>(out_lo + off + d0) =
            make_uint2(pack_bf16x2(bf16_residual(oo[0]), bf16_residual(oo[1])), pack_bf16x2(bf16_residual(oo[2]), bf16_residual(oo[3])));
    }
  }
}

template <typename TIn, int L>
__global__ void __launch_bounds__(128) attention_tokens_kernel(const AttentionArgs a) {
  constexpr int HD = 256;
  constexpr int PL = HD / 32;
  const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= a.n_seq * a.H) return;
  const int lane = threadIdx.x & 31;
  const int seq = gw / a.H, h = gw % a.H;
  const TIn* qg = reinterpret_cast<const TIn*>(a.q) + static_cast<long long>(seq) * L * a.ldq + h * HD;
  const TIn* kg = reinterpret_cast<const TIn*>(a.k) + static_cast<long long>(seq) * L * a.ldk + h * HD;
  const TIn* vg = reinterpret_cast<const TIn*>(a.v) + static_cast<long long>(seq) * L * a.ldv + h * HD;
  float q[L][PL], k[L][PL], v[L][PL];
#pragma unroll
  for (int i = 0; i < L; ++i) load_row_regs<TIn, HD>(qg + static_cast<long long>(i) * a.ldq, lane, q[i]);
#pragma unroll
  for (int i = 0; i < L; ++i) load_row_regs<TIn, HD>(kg + static_cast<long long>(i) * a.ldk, lane, k[i]);
#pragma unroll
  for (int i = 0; i < L; ++i) load_row_regs<TIn, HD>(vg + static_cast<long long>(i) * a.ldv, lane, v[i]);
  float* pr = nullptr;
  if (a.probs != nullptr)
    pr = a.probs + static_cast<long long>(seq / a.p_inner) * a.p_outer +
         static_cast<long long>(seq % a.p_inner) * a.p_inner_stride + static_cast<long long>(h) * L * L;
#pragma unroll
  for (int i = 0; i < L; ++i) {
    float s[L];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < L; ++j) {
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < PL; ++e) d = fmaf(q[i][e], k[j][e], d);
      d = warp_sum(d) * a.scale;
      if (a.mask == 3 && j == i) d = -INFINITY;  // diagonal masked (cross_attn=True)
      s[j] = d;
      mx = fmaxf(mx, d);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < L; ++j) {
      s[j] = (s[j] == -INFINITY) ? 0.f : expf(s[j] - mx);
      sum += s[j];
    }
    const float inv = 1.0f / sum;
    float o[PL];
#pragma unroll
    for (int e = 0; e < PL; ++e) o[e] = 0.f;
    float mine = 0.f;
    const float* dm = a.drop != nullptr ? a.drop + ((static_cast<long long>(seq) * a.H + h) * L + i) * L : nullptr;
#pragma unroll
    for (int j = 0; j < L; ++j) {
      float p = s[j] * inv;
      if (lane == j) mine = p;  // the undropped softmax is what `probs` keeps
      if (dm != nullptr) p *= dm[j];
#pragma unroll
      for (int e = 0; e < PL; ++e) o[e] = fmaf(p, v[j][e], o[e]);
    }
    if (pr != nullptr && lane < L) pr[i * L + lane] = mine;
    store_row_regs<TIn, HD>(a.out_hi, a.out_lo, (static_cast<long long>(seq) * L + i) * a.ldo + h * HD, lane, o);
  }
  ptx::prof_mark_end(a.t_end);
}


// ------------------------------------------------------------------------------------------------
// Causal / full attention over a short sequence (L <= 32) with head_dim 256 or 512, bf16 inputs: the GPT-2
// future predictor (18 x 18 causal, head_dim 512; transformers eager attention) and the CA-Fuser's causal
// self/cross attention (10 x 10, head_dim 256; reference models/transformerblock.py:24-33,64-74).
// CUDA cores cannot do 2 * 18 * 18 * 512 MACs per (clip, head) inside the kernel's HBM time (~12 us per launch
// at B = 256), so Q.K^T and P.V run on the tensor cores with warp-level mma.sync.m16n8k16 (bf16 in, fp32
// accumulate): one CTA (4 warps) per (sequence, head); Q, K, V rows staged once in padded shared memory
// (ldmatrix conflict-free), scores and softmax in fp32, P re-quantised to bf16 for the second product, the
// output tile staged through shared memory and written with coalesced 16-byte stores.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void mma_f16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// operand format of the warp-level attention kernels: bf16 (MODE_BF16) or fp16 (MODE_FP16)
template <bool FP16>
__device__ __forceinline__ void mma_op16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (FP16) mma_f16_16816(d, a, b0, b1);
  else mma_bf16_16816(d, a, b0, b1);
}

template <int HD, int LP>
struct AttnMmaSmem {
  static constexpr int kLP = LP;                  // padded sequence length (LP / 16 m16 tiles): 32 or 64
  static constexpr int kRowBytes = HD * 2 + 16;   // odd multiple of 16 B: ldmatrix rows hit distinct banks
  static constexpr int kTileBytes = kLP * kRowBytes;
  static constexpr int kSStride = LP + 1;         // fp32 scores [LP][LP + 1]
  static constexpr int kPRowBytes = LP * 2 + 16;  // bf16 probabilities [LP][LP] + pad
  static constexpr int kBytes = 3 * kTileBytes + kLP * kSStride * 4 + kLP * kPRowBytes;  // upper bound (L = LP)
  // Only the L real rows of Q, K and V are staged (ldmatrix row addresses are clamped to row L - 1: the duplicated
  // rows only feed scores / probabilities that are masked or zero), which is what lets 3 CTAs share an SM for the
  // GPT-2 shape (L = 18, head_dim 512: 61.5 KB instead of 104 KB).
  static constexpr int bytes_for(int L) { return 3 * L * kRowBytes + kLP * kSStride * 4 + kLP * kPRowBytes; }
};

// NW warps per CTA: 4 for sequences up to 32 tokens, 8 for the 50-token T-SA-Fuser sequences (16 score blocks and
// 64 softmax rows per CTA: twice the warps halve the CTA's serial compute between its load and its store).
template <int HD, int LP, int NW = 4, bool FP16 = false>
__global__ void __launch_bounds__(NW * 32) attention_mma_kernel(const AttentionArgs a) {
  using S = AttnMmaSmem<HD, LP>;
  constexpr int MT = LP / 16;  // 16-row tiles
  extern __shared__ uint4 smem_attn[];
  uint8_t* base = reinterpret_cast<uint8_t*>(smem_attn);
  const int L = a.L;
  const int tile_bytes = L * S::kRowBytes;  // L real rows per operand
  const uint32_t sq = static_cast<uint32_t>(__cvta_generic_to_shared(base));
  const uint32_t sk = sq + tile_bytes, sv = sk + tile_bytes;
  float* ss = reinterpret_cast<float*>(base + 3 * tile_bytes);
  uint8_t* sp_ptr = base + 3 * tile_bytes + S::kLP * S::kSStride * 4;
  const uint32_t sp = sv + tile_bytes + S::kLP * S::kSStride * 4;

  ptx::griddep_launch();
  ptx::griddep_wait();
  const int seq = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const __nv_bfloat16* qg = reinterpret_cast<const __nv_bfloat16*>(a.q) + static_cast<long long>(seq) * L * a.ldq + h * HD;
  const __nv_bfloat16* kg = reinterpret_cast<const __nv_bfloat16*>(a.k) + static_cast<long long>(seq) * L * a.ldk + h * HD;
  const __nv_bfloat16* vg = reinterpret_cast<const __nv_bfloat16*>(a.v) + static_cast<long long>(seq) * L * a.ldv + h * HD;

  // ---- stage the L rows of Q, K, V with cp.async (every 16-byte copy in flight at once) ----
  // Tile rows >= L are never stored: the ldmatrix addresses below clamp to row L - 1.  Duplicated Q / K rows only
  // produce scores that are masked or never read; duplicated V rows meet zero probabilities and are finite.
  constexpr int kVecPerRow = HD / 8;             // 16-byte vectors per row
  constexpr int kRowsPerPass = NW * 32 / kVecPerRow;  // rows covered by the CTA per pass
  {
    const int c = tid % kVecPerRow;
    const int r0 = tid / kVecPerRow;
    const __nv_bfloat16* qp = qg + static_cast<long long>(r0) * a.ldq + c * 8;
    const __nv_bfloat16* kp = kg + static_cast<long long>(r0) * a.ldk + c * 8;
    const __nv_bfloat16* vp = vg + static_cast<long long>(r0) * a.ldv + c * 8;
    uint32_t dst = sq + r0 * S::kRowBytes + c * 16;
    // two groups: Q and K first (scores and softmax run while V is still in flight), then V
    for (int r = r0; r < L; r += kRowsPerPass) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(qp) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + tile_bytes), "l"(kp) : "memory");
      qp += kRowsPerPass * a.ldq;
      kp += kRowsPerPass * a.ldk;
      dst += kRowsPerPass * S::kRowBytes;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    dst = sv + r0 * S::kRowBytes + c * 16;
    for (int r = r0; r < L; r += kRowsPerPass) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(vp) : "memory");
      vp += kRowsPerPass * a.ldv;
      dst += kRowsPerPass * S::kRowBytes;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");  // Q and K have landed
  }
  __syncthreads();

  const int g8 = lane >> 2, t4 = lane & 3;
  // ---- S = Q K^T: 16 x 16 blocks (query tile mt, key tile nb) over the full head_dim, dealt round-robin to warps ----
  for (int blk = warp; blk < MT * MT; blk += NW) {
    const int mt = blk / MT, nb = blk % MT;
    const bool skip = (nb * 16 >= L) || (mt * 16 >= L) || (a.mask == 1 && nb * 16 > mt * 16 + 15);
    if (skip) continue;
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    const uint32_t a_addr = sq + min(mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, L - 1) * S::kRowBytes + (lane >> 4) * 16;
    const uint32_t b_addr = sk + min(nb * 16 + (lane & 7) + (lane >> 4) * 8, L - 1) * S::kRowBytes + ((lane >> 3) & 1) * 16;
#pragma unroll 4
    for (int ks = 0; ks < HD / 16; ++ks) {
      uint32_t af[4], bf[4];
      ldmatrix_x4(a_addr + ks * 32, af);
      ldmatrix_x4(b_addr + ks * 32, bf);
      mma_op16_16816<FP16>(acc[0], af, bf[0], bf[1]);
      mma_op16_16816<FP16>(acc[1], af, bf[2], bf[3]);
    }
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      float* r0 = ss + (mt * 16 + g8) * S::kSStride + nb * 16 + nt * 8 + t4 * 2;
      r0[0] = acc[nt][0];
      r0[1] = acc[nt][1];
      r0[8 * S::kSStride] = acc[nt][2];
      r0[8 * S::kSStride + 1] = acc[nt][3];
    }
  }
  __syncthreads();

  // ---- softmax rows (fp32), P -> bf16.  mask 1: j <= i; mask 2: (j % T) <= (i % T); mask 0: none ----
  for (int i = warp; i < S::kLP; i += NW) {
    float p[LP / 32];
#pragma unroll
    for (int c = 0; c < LP / 32; ++c) p[c] = 0.f;
    if (i < L) {
      float sc[LP / 32];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < LP / 32; ++c) {
        const int j = lane + 32 * c;
        const bool ok = (j < L) && (a.mask == 0 || (a.mask == 1 && j <= i) || (a.mask == 2 && (j % a.T) <= (i % a.T)));
        sc[c] = ok ? ss[i * S::kSStride + j] * a.scale : -INFINITY;
        mx = fmaxf(mx, sc[c]);
      }
      mx = warp_max(mx);
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < LP / 32; ++c) {
        p[c] = (sc[c] == -INFINITY) ? 0.f : expf(sc[c] - mx);
        sum += p[c];
      }
      const float inv = 1.0f / warp_sum(sum);
#pragma unroll
      for (int c = 0; c < LP / 32; ++c) {
        p[c] *= inv;
        const int j = lane + 32 * c;
        if (a.probs != nullptr && j < L)
          a.probs[static_cast<long long>(seq / a.p_inner) * a.p_outer + static_cast<long long>(seq % a.p_inner) * a.p_inner_stride +
                  (static_cast<long long>(h) * L + i) * L + j] = p[c];
      }
    }
#pragma unroll
    for (int c = 0; c < LP / 32; ++c)
    {
      if constexpr (FP16) reinterpret_cast<__half*>(sp_ptr + i * S::kPRowBytes)[lane + 32 * c] = __float2half_rn(p[c]);
      else reinterpret_cast<__nv_bfloat16*>(sp_ptr + i * S::kPRowBytes)[lane + 32 * c] = __float2bfloat16_rn(p[c]);
    }
  }
  __syncthreads();

  // ---- O = P V: warp -> HD / NW output dims; output tile staged in the (now free) Q region ----
  asm volatile("cp.async.wait_group 0;" ::: "memory");  // V has landed (the barrier above published the probabilities;
  __syncthreads();                                       //  this one publishes every thread's share of V)
  {
    const int ksteps = (L + 15) / 16;
#pragma unroll 1
    for (int np = 0; np < HD / NW / 16; ++np) {
      const int n0 = warp * (HD / NW) + np * 16;
      float acc[MT][2][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < MT; ++ks) {
        if (ks < ksteps) {
          uint32_t vf[4];
          ldmatrix_x4_trans(sv + min(ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, L - 1) * S::kRowBytes + (n0 + (lane >> 4) * 8) * 2, vf);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            uint32_t pa[4];
            ldmatrix_x4(sp + (mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * S::kPRowBytes + ks * 32 + (lane >> 4) * 16, pa);
            mma_op16_16816<FP16>(acc[mt][0], pa, vf[0], vf[1]);
            mma_op16_16816<FP16>(acc[mt][1], pa, vf[2], vf[3]);
          }
        }
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          // output rows < L only: the Q region holds L rows and the K / V regions behind it are still being read
          uint8_t* o0 = base + (mt * 16 + g8) * S::kRowBytes + (n0 + nt * 8 + t4 * 2) * 2;
          if (mt * 16 + g8 < L) *reinterpret_cast<uint32_t*>(o0) = pack_op16x2(acc[mt][nt][0], acc[mt][nt][1], FP16);
          if (mt * 16 + g8 + 8 < L) *reinterpret_cast<uint32_t*>(o0 + 8 * S::kRowBytes) = pack_op16x2(acc[mt][nt][2], acc[mt][nt][3], FP16);
        }
    }
  }
  __syncthreads();
  for (int i = tid; i < L * kVecPerRow; i += NW * 32) {
    const int r = i / kVecPerRow, c = i % kVecPerRow;
    *reinterpret_cast<uint4*>(a.out_hi + (static_cast<long long>(seq) * L + r) * a.ldo + h * HD + c * 8) =
        *reinterpret_cast<const uint4*>(base + r * S::kRowBytes + c * 16);
  }
  ptx::prof_mark_end(a.t_end);
}

// ------------------------------------------------------------------------------------------------
// SA-Fuser attention on the tensor cores (bf16 inputs): the register kernel above needs ~1700 instructions per
// (timestep, head) and is issue-bound at twice the kernel's HBM time (ncu).  Here one warp handles G = 16 / L
// consecutive timesteps of one head as ONE 16 x 16 mma tile: S = Q K^T over the G*L token rows (16 k-steps of
// m16n8k16), block-diagonal mask (a token only sees the tokens of its own timestep), softmax on the accumulator
// fragments, P.V with one k-step per 8 output dims.  Rows are staged per warp with cp.async; no block barrier.
// ------------------------------------------------------------------------------------------------
template <int L, bool FP16 = false>
__global__ void __launch_bounds__(256) attention_tokens_mma_kernel(const AttentionArgs a) {
  constexpr int HD = 256;
  constexpr int G = 16 / L;
  constexpr int RB = HD * 2 + 16;  // padded row: odd multiple of 16 B
  constexpr int TILE = 16 * RB;
  extern __shared__ uint4 smem_attn[];
  ptx::griddep_launch();
  ptx::griddep_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = warp;  // one warp per head
  const int seq0 = blockIdx.x * G;
  const int n_t = (a.n_seq - seq0) < G ? (a.n_seq - seq0) : G;
  const int nrows = n_t * L;
  uint8_t* wbase = reinterpret_cast<uint8_t*>(smem_attn) + warp * 3 * TILE;
  const uint32_t sq = static_cast<uint32_t>(__cvta_generic_to_shared(wbase));
  const uint32_t sk = sq + TILE, sv = sk + TILE;
  const long long row0 = static_cast<long long>(seq0) * L;
  const __nv_bfloat16* qg = reinterpret_cast<const __nv_bfloat16*>(a.q) + row0 * a.ldq + h * HD + lane * 8;
  const __nv_bfloat16* kg = reinterpret_cast<const __nv_bfloat16*>(a.k) + row0 * a.ldk + h * HD + lane * 8;
  const __nv_bfloat16* vg = reinterpret_cast<const __nv_bfloat16*>(a.v) + row0 * a.ldv + h * HD + lane * 8;
  {
    // two groups: Q and K first (the scores and the softmax run while V is still in flight), then V
    uint32_t dst = sq + lane * 16;
    for (int r = 0; r < nrows; ++r) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(qg) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + TILE), "l"(kg) : "memory");
      qg += a.ldq;
      kg += a.ldk;
      dst += RB;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    dst = sv + lane * 16;
    for (int r = 0; r < nrows; ++r) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(vg) : "memory");
      vg += a.ldv;
      dst += RB;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // pad rows: V must be finite (it meets zero probabilities); Q / K pad rows only feed masked scores, but keep
    // them finite as well so no NaN ever enters the accumulators
    for (int r = nrows; r < 16; ++r) {
      *reinterpret_cast<uint4*>(wbase + r * RB + lane * 16) = make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(wbase + TILE + r * RB + lane * 16) = make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(wbase + 2 * TILE + r * RB + lane * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    asm volatile("cp.async.wait_group 1;" ::: "memory");  // Q and K have landed
  }
  __syncwarp();

  const int g8 = lane >> 2, t4 = lane & 3;
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  {
    const uint32_t a_addr = sq + ((lane & 7) + ((lane >> 3) & 1) * 8) * RB + (lane >> 4) * 16;
    const uint32_t b_addr = sk + ((lane & 7) + (lane >> 4) * 8) * RB + ((lane >> 3) & 1) * 16;
#pragma unroll 4
    for (int ks = 0; ks < HD / 16; ++ks) {
      uint32_t af[4], bf[4];
      ldmatrix_x4(a_addr + ks * 32, af);
      ldmatrix_x4(b_addr + ks * 32, bf);
      mma_op16_16816<FP16>(acc[0], af, bf[0], bf[1]);
      mma_op16_16816<FP16>(acc[1], af, bf[2], bf[3]);
    }
  }
  // softmax on the fragments: this thread holds rows g8 and g8 + 8, columns nt * 8 + 2 * t4 + {0, 1}
  float p[2][4];  // [row half][nt * 2 + e]
#pragma unroll
  for (int rh = 0; rh < 2; ++rh) {
    const int i = g8 + rh * 8;
    float sc[4];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = (c >> 1) * 8 + t4 * 2 + (c & 1);
      const bool ok = (i < nrows) && (j < nrows) && (i / L == j / L) && !(a.mask == 3 && i == j);
      sc[c] = ok ? acc[c >> 1][rh * 2 + (c & 1)] * a.scale : -INFINITY;
      mx = fmaxf(mx, sc[c]);
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      p[rh][c] = (sc[c] == -INFINITY) ? 0.f : expf(sc[c] - mx);
      sum += p[rh][c];
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float inv = sum > 0.f ? 1.0f / sum : 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) p[rh][c] *= inv;
    if (a.probs != nullptr && i < nrows) {
      const int seq = seq0 + i / L, ii = i % L;
      float* pr = a.probs + static_cast<long long>(seq / a.p_inner) * a.p_outer +
                  static_cast<long long>(seq % a.p_inner) * a.p_inner_stride + (static_cast<long long>(h) * L + ii) * L;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = (c >> 1) * 8 + t4 * 2 + (c & 1);
        if (j < nrows && j / L == i / L) pr[j % L] = p[rh][c];
      }
    }
  }
  uint32_t pa[4];
  pa[0] = pack_op16x2(p[0][0], p[0][1], FP16);
  pa[1] = pack_op16x2(p[1][0], p[1][1], FP16);
  pa[2] = pack_op16x2(p[0][2], p[0][3], FP16);
  pa[3] = pack_op16x2(p[1][2], p[1][3], FP16);
  asm volatile("cp.async.wait_group 0;" ::: "memory");  // V has landed
  __syncwarp();  // ... for every lane; and every lane is done reading Q before the region is reused for the output tile
#pragma unroll 4
  for (int np = 0; np < HD / 16; ++np) {
    uint32_t vf[4];
    ldmatrix_x4_trans(sv + ((lane & 7) + ((lane >> 3) & 1) * 8) * RB + (np * 16 + (lane >> 4) * 8) * 2, vf);
    float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    mma_op16_16816<FP16>(o[0], pa, vf[0], vf[1]);
    mma_op16_16816<FP16>(o[1], pa, vf[2], vf[3]);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      uint8_t* o0 = wbase + g8 * RB + (np * 16 + nt * 8 + t4 * 2) * 2;
      *reinterpret_cast<uint32_t*>(o0) = pack_op16x2(o[nt][0], o[nt][1], FP16);
      *reinterpret_cast<uint32_t*>(o0 + 8 * RB) = pack_op16x2(o[nt][2], o[nt][3], FP16);
    }
  }
  __syncwarp();
  __nv_bfloat16* og = a.out_hi + row0 * a.ldo + h * HD + lane * 8;
  for (int r = 0; r < nrows; ++r) {
    *reinterpret_cast<uint4*>(og) = *reinterpret_cast<const uint4*>(wbase + r * RB + lane * 16);
    og += a.ldo;
  }
  ptx::prof_mark_end(a.t_end);
}


// ------------------------------------------------------------------------------------------------
// Logit post-processing (reference challenge.py:196-210 marginalize_verb_noun and
// common/utils.py:19-42 top-k ranking): per clip  p = softmax(action logits);  verb[v] = sum_{a: verb(a)=v} p[a],
// noun[n] likewise (the reference multiplies by the 0/1 matrices class_mappings[('verb','action')] /
// [('noun','action')], which have exactly one 1 per action row);  top-K indices of action / verb / noun scores.
// One CTA per clip; verb/noun accumulators in shared memory; K rounds of block arg-max (ties -> lower index).
// The accumulators are 64-bit fixed point (p * 2^40, integer atomics): integer addition is associative, so the verb /
// noun scores are bit-reproducible run to run whatever order the threads arrive in (float atomics are not); the
// quantisation is 2^-41 per action, < 2e-9 on a sum of 3806 terms.  NaN logits rank last instead of poisoning the
// comparisons.
// ------------------------------------------------------------------------------------------------
struct MarginalizeArgs {
  const float* logits;  // [B, ld]
  long long ld;
  int B, A;
  const int* verb_of;  // [A]
  const int* noun_of;  // [A]
  int n_verb, n_noun;
  float* probs;  // [B, A] or nullptr
  float* verb;   // [B, n_verb]
  float* noun;   // [B, n_noun]
  int* topk;     // [B, 3, K]: action, verb, noun
  int K;
};

__device__ __forceinline__ void block_argmax(float& v, int& idx, float* red_v, int* red_i) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  __syncthreads();
  if (lane == 0) { red_v[warp] = v; red_i[warp] = idx; }
  __syncthreads();
  v = (lane < nw) ? red_v[lane] : -INFINITY;
  idx = (lane < nw) ? red_i[lane] : 0x7fffffff;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
}

__device__ __forceinline__ void block_topk(const float* vals, int n, int K, int* out, float* red_v, int* red_i, int* chosen) {
  for (int k = 0; k < K; ++k) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      bool taken = false;
      for (int j = 0; j < k; ++j) taken |= (chosen[j] == i);
      float v = vals[i];
      if (v != v) v = -INFINITY;  // NaN ranks last
      if (!taken && (v > best || (v == best && i < bi))) { best = v; bi = i; }
    }
    block_argmax(best, bi, red_v, red_i);
    if (threadIdx.x == 0) { chosen[k] = bi; out[k] = bi; }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) marginalize_topk_kernel(const MarginalizeArgs a) {
  extern __shared__ unsigned long long smem_mg64[];
  unsigned long long* acc = smem_mg64;                     // [n_verb + n_noun] fixed-point sums
  float* sv = reinterpret_cast<float*>(acc + a.n_verb + a.n_noun);  // [n_verb]
  float* sn = sv + a.n_verb;      // [n_noun]
  float* red_v = sn + a.n_noun;   // [32]
  int* red_i = reinterpret_cast<int*>(red_v + 32);
  int* chosen = red_i + 32;       // [K]
  const int b = blockIdx.x;
  const float* lg = a.logits + b * a.ld;
  for (int i = threadIdx.x; i < a.n_verb + a.n_noun; i += blockDim.x) acc[i] = 0ull;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < a.A; i += blockDim.x) mx = fmaxf(mx, lg[i]);  // fmaxf drops NaN operands
  int dummy = 0;
  block_argmax(mx, dummy, red_v, red_i);
  float sum = 0.f;
  for (int i = threadIdx.x; i < a.A; i += blockDim.x) sum += expf(lg[i] - mx);
  sum = warp_sum(sum);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red_v[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = (threadIdx.x & 31) < (blockDim.x >> 5) ? red_v[threadIdx.x & 31] : 0.f;
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  __syncthreads();
  for (int i = threadIdx.x; i < a.A; i += blockDim.x) {
    const float p = expf(lg[i] - mx) * inv;
    if (a.probs != nullptr) a.probs[static_cast<long long>(b) * a.A + i] = p;
    const unsigned long long q = (p == p) ? static_cast<unsigned long long>(static_cast<double>(p) * 1099511627776.0) : 0ull;  // 2^40
    atomicAdd(&acc[a.verb_of[i]], q);
    atomicAdd(&acc[a.n_verb + a.noun_of[i]], q);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < a.n_verb + a.n_noun; i += blockDim.x)
    sv[i] = static_cast<float>(static_cast<double>(acc[i]) * (1.0 / 1099511627776.0));
  __syncthreads();
  for (int i = threadIdx.x; i < a.n_verb; i += blockDim.x) a.verb[static_cast<long long>(b) * a.n_verb + i] = sv[i];
  for (int i = threadIdx.x; i < a.n_noun; i += blockDim.x) a.noun[static_cast<long long>(b) * a.n_noun + i] = sn[i];
  if (a.topk != nullptr) {
    int* out = a.topk + static_cast<long long>(b) * 3 * a.K;
    block_topk(lg, a.A, a.K, out, red_v, red_i, chosen);  // softmax is monotone: rank the logits
    block_topk(sv, a.n_verb, a.K, out + a.K, red_v, red_i, chosen);
    block_topk(sn, a.n_noun, a.K, out + 2 * a.K, red_v, red_i, chosen);
  }
}


// ------------------------------------------------------------------------------------------------
// Autoregressive roll-out of the future predictor (fp_output_len > 1; reference models/future_prediction.py:
// 395-412: the last hidden state is fed back as the next input embedding, GPT-2 runs one position with its KV
// cache).  Two small kernels: the position-embedding add for the fed-back row, and single-query attention over
// the cached keys/values of the T prompt positions plus the positions generated so far (online softmax).
// ------------------------------------------------------------------------------------------------
__global__ void add_row_vector_kernel(float* __restrict__ out, const float* __restrict__ in, const float* __restrict__ vec,
                                      int rows, int dim, unsigned long long* t_end = nullptr) {
  const int total = rows * dim;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) out[i] = in[i] + vec[i % dim];
  ptx::prof_mark_end(t_end);
}

// first kernel of a profiled forward: the reference point of launch 0's interval
__global__ void prof_mark_kernel(unsigned long long* slot) { ptx::prof_mark_end(slot); }

struct DecodeAttnArgs {
  const void* cache;   // [B * T, ld] rows (b, t): q | k | v of the prompt positions
  const void* fresh;   // [B * n_new_max, ld] rows (b, s): q | k | v of generated positions; the query is row (b, n_new - 1)
  long long ld;        // 3 * H * HD
  int B, T, H, n_new, n_new_max;
  float scale;
  __nv_bfloat16* out_hi;  // [B, H * HD]
  __nv_bfloat16* out_lo;
  unsigned long long* t_end;  // profiling slot or nullptr
};

template <typename TIn, int HD>
__global__ void __launch_bounds__(128) attention_decode_kernel(const DecodeAttnArgs a) {
  constexpr int PL = HD / 32;
  const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= a.B * a.H) return;
  const int lane = threadIdx.x & 31;
  const int b = gw / a.H, h = gw % a.H;
  const long long D = static_cast<long long>(a.H) * HD;
  const TIn* cache = reinterpret_cast<const TIn*>(a.cache);
  const TIn* fresh = reinterpret_cast<const TIn*>(a.fresh);
  float q[PL];
  load_row_regs<TIn, HD>(fresh + (static_cast<long long>(b) * a.n_new_max + a.n_new - 1) * a.ld + h * HD, lane, q);
  float m = -INFINITY, l = 0.f, o[PL];
#pragma unroll
  for (int e = 0; e < PL; ++e) o[e] = 0.f;
  const int n_keys = a.T + a.n_new;
  for (int j = 0; j < n_keys; ++j) {
    const TIn* row = (j < a.T) ? cache + (static_cast<long long>(b) * a.T + j) * a.ld
                               : fresh + (static_cast<long long>(b) * a.n_new_max + (j - a.T)) * a.ld;
    float kr[PL], vr[PL];
    load_row_regs<TIn, HD>(row + D + h * HD, lane, kr);
    load_row_regs<TIn, HD>(row + 2 * D + h * HD, lane, vr);
    float d = 0.f;
#pragma unroll
    for (int e = 0; e < PL; ++e) d = fmaf(q[e], kr[e], d);
    const float sc = warp_sum(d) * a.scale;
    const float m_new = fmaxf(m, sc);
    const float corr = expf(m - m_new);  // exp(-inf) = 0 on the first key
    const float p = expf(sc - m_new);
    l = l * corr + p;
#pragma unroll
    for (int e = 0; e < PL; ++e) o[e] = fmaf(p, vr[e], o[e] * corr);
    m = m_new;
  }
  const float inv = 1.0f / l;
#pragma unroll
  for (int e = 0; e < PL; ++e) o[e] *= inv;
  store_row_regs<TIn, HD>(a.out_hi, a.out_lo, static_cast<long long>(b) * D + h * HD, lane, o);
  ptx::prof_mark_end(a.t_end);
}

// ------------------------------------------------------------------------------------------------
// Score fusion (reference models/fusion.py:57 softmax over MATT's outputs; models/future_prediction.py:341-350):
//   p[r, :] = softmax(attn_logits[r, :M]);  out[r, c] = sum_i p[r, i] * logits_i[r, c]
// HBM-bound: one pass over the M logit tensors, float4 accesses (pitches are multiples of 4 floats).
// ------------------------------------------------------------------------------------------------
struct ScoreFusionArgs {
  const float* attn_logits;
  long long ld_a;
  int M;
  const float* logits[8];
  long long ld_l;
  int rows, C;
  float* attn;  // [rows, M] or nullptr
  float* out;
  long long ld_o;
};

__global__ void __launch_bounds__(256) score_fusion_kernel(const ScoreFusionArgs a) {
  const int r = blockIdx.x;  // rows on gridDim.x (2^31 - 1 limit), column blocks on gridDim.y
  float p[8];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    p[i] = (i < a.M) ? a.attn_logits[r * a.ld_a + i] : -INFINITY;
    mx = fmaxf(mx, p[i]);
  }
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    p[i] = (i < a.M) ? __expf(p[i] - mx) : 0.f;
    sum += p[i];
  }
  const float inv = 1.0f / sum;
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] *= inv;
  if (a.attn != nullptr && blockIdx.y == 0 && threadIdx.x < a.M) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i == static_cast<int>(threadIdx.x)) v = p[i];
    a.attn[static_cast<long long>(r) * a.M + threadIdx.x] = v;
  }
  const int quads = (a.C + 3) / 4;
  for (int q = blockIdx.y * blockDim.x + threadIdx.x; q < quads; q += gridDim.y * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < a.M) {
        const float4 v = *reinterpret_cast<const float4*>(a.logits[i] + r * a.ld_l + q * 4);
        acc.x = fmaf(p[i], v.x, acc.x);
        acc.y = fmaf(p[i], v.y, acc.y);
        acc.z = fmaf(p[i], v.z, acc.z);
        acc.w = fmaf(p[i], v.w, acc.w);
      }
    }
    *reinterpret_cast<float4*>(a.out + r * a.ld_o + q * 4) = acc;
  }
}

}  // namespace afft
