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
//   warps 4..7  epilogue       tcgen05.ld 32x32b -> registers -> bias / GELU / residual / casts ->
//                               global.  Each warp owns one 32-lane TMEM quadrant = 32 output rows.
//
// SPLIT == 3 is the "strict" mode: both operands are given as bf16 hi/lo
// pairs (x = hi + lo to ~16 mantissa bits) and the kernel accumulates hi.hi + hi.lo + lo.hi into
// the same TMEM tile, i.e. three tensor-core passes per K slab for ~fp32-grade products.
#pragma once

#include <cuda_bf16.h>

#include "ptx_sm100.cuh"

namespace afft {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 256;

enum : int { ACT_NONE = 0, ACT_GELU_ERF = 1, ACT_GELU_TANH = 2 };

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

template <int BLOCK_N, int SPLIT>
struct GemmTraits {
  static_assert(BLOCK_N == 128 || BLOCK_N == 256, "BLOCK_N must be 128 or 256");
  static_assert(SPLIT == 1 || SPLIT == 3, "SPLIT must be 1 (bf16) or 3 (bf16x3)");
  static constexpr int kPairs = (SPLIT == 3) ? 2 : 1;
  static constexpr uint32_t kABytes = kBlockM * kBlockK * 2;
  static constexpr uint32_t kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr uint32_t kStageBytes = kPairs * (kABytes + kBBytes);
  static constexpr int kStages = (SPLIT == 3) ? (BLOCK_N == 256 ? 2 : 3) : (BLOCK_N == 256 ? 4 : 6);
  static constexpr uint32_t kTmemCols = 2 * BLOCK_N;
  static constexpr uint32_t kBarrierBytes = 256;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kBarrierBytes + 1024;  // +align slack
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
};

// --------------------------------------------------------------------------------------------
// Activation functions (fp32).  erf via Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7) so the epilogue
// stays well under the tensor-core time of a K=1024 tile; tanh via exp.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) {
  // 0.5 x (1 + erf(x / sqrt 2))   (torch.nn.GELU default; reference models/transformerblock.py:79)
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = __expf(-z * z);
  const float erf_abs = fmaf(-p, e, 1.0f);  // erf(|x|/sqrt2)
  const float erf_v = copysignf(erf_abs, x);
  return 0.5f * x * (1.0f + erf_v);
}

__device__ __forceinline__ float gelu_tanh(float x) {
  // 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))   (HF "gelu_new", GPT-2 MLP)
  const float u = 0.7978845608028654f * fmaf(0.044715f * x * x, x, x);
  // tanh(u) = 1 - 2 / (1 + exp(2u)); exp overflow -> inf -> tanh = 1, underflow -> -1: both exact.
  const float e = __expf(2.0f * u);
  const float th = 1.0f - __fdividef(2.0f, 1.0f + e);
  return 0.5f * x * (1.0f + th);
}

// --------------------------------------------------------------------------------------------
// Epilogue for one thread: one output row, 32 consecutive columns starting at n0.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ void epilogue_row32(const GemmEpilogue& ep, float (&x)[32], int row, long long orow,
                                               int n0, int N) {
  const bool full = (n0 + 32 <= N);
  if (ep.bias != nullptr) {
    if (full) {
      const float4* b4 = reinterpret_cast<const float4*>(ep.bias + n0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b = __ldg(b4 + j);
        x[4 * j + 0] += b.x;
        x[4 * j + 1] += b.y;
        x[4 * j + 2] += b.z;
        x[4 * j + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + j < N) x[j] += __ldg(ep.bias + n0 + j);
    }
  }
  if (ep.act == ACT_GELU_ERF) {
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = gelu_erf(x[j]);
  } else if (ep.act == ACT_GELU_TANH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = gelu_tanh(x[j]);
  }
  if (ep.res != nullptr) {
    const long long rrow = (ep.res_mod > 0) ? static_cast<long long>(row % ep.res_mod) : orow;
    const float* rp = ep.res + rrow * ep.ld_res + n0;
    if (full) {
      const float4* r4 = reinterpret_cast<const float4*>(rp);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 r = r4[j];
        x[4 * j + 0] += r.x;
        x[4 * j + 1] += r.y;
        x[4 * j + 2] += r.z;
        x[4 * j + 3] += r.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + j < N) x[j] += rp[j];
    }
  }
  if (ep.out_f32 != nullptr) {
    float* op = ep.out_f32 + orow * ep.ld_f32 + n0;
    if (full) {
      float4* o4 = reinterpret_cast<float4*>(op);
#pragma unroll
      for (int j = 0; j < 8; ++j) o4[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n0 + j < N) op[j] = x[j];
    }
  }
  if (ep.out_hi != nullptr) {
    __nv_bfloat16* hp = ep.out_hi + orow * ep.ld_bf16 + n0;
    __nv_bfloat16* lp = (ep.out_lo != nullptr) ? ep.out_lo + orow * ep.ld_bf16 + n0 : nullptr;
    if (full) {
      uint32_t hw[16], lw[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * j]);
        const __nv_bfloat16 h1 = __float2bfloat16_rn(x[2 * j + 1]);
        hw[j] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) |
                (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(x[2 * j] - __bfloat162float(h0));
        const __nv_bfloat16 l1 = __float2bfloat16_rn(x[2 * j + 1] - __bfloat162float(h1));
        lw[j] = static_cast<uint32_t>(__bfloat16_as_ushort(l0)) |
                (static_cast<uint32_t>(__bfloat16_as_ushort(l1)) << 16);
      }
      uint4* h4 = reinterpret_cast<uint4*>(hp);
#pragma unroll
      for (int j = 0; j < 4; ++j) h4[j] = make_uint4(hw[4 * j], hw[4 * j + 1], hw[4 * j + 2], hw[4 * j + 3]);
      if (lp != nullptr) {
        uint4* l4 = reinterpret_cast<uint4*>(lp);
#pragma unroll
        for (int j = 0; j < 4; ++j) l4[j] = make_uint4(lw[4 * j], lw[4 * j + 1], lw[4 * j + 2], lw[4 * j + 3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (n0 + j < N) {
          const __nv_bfloat16 h = __float2bfloat16_rn(x[j]);
          hp[j] = h;
          if (lp != nullptr) lp[j] = __float2bfloat16_rn(x[j] - __bfloat162float(h));
        }
      }
    }
  }
}

// --------------------------------------------------------------------------------------------
// The kernel
// --------------------------------------------------------------------------------------------
template <int BLOCK_N, int SPLIT>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                         const __grid_constant__ CUtensorMap tm_a_lo,
                         const __grid_constant__ CUtensorMap tm_b_lo, const GemmEpilogue ep, const int M,
                         const int N, const int K) {
  using T = GemmTraits<BLOCK_N, SPLIT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_raw_u32 = ptx::smem_u32(smem_raw);
  const uint32_t smem_base = (smem_raw_u32 + 1023u) & ~1023u;  // SWIZZLE_128B tiles: 1024-B aligned
  const uint32_t bar_base = smem_base + T::kStages * T::kStageBytes;
  auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (T::kStages + s); };
  auto tmem_full_bar = [&](uint32_t a) { return bar_base + 8u * (2 * T::kStages + a); };
  auto tmem_empty_bar = [&](uint32_t a) { return bar_base + 8u * (2 * T::kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * T::kStages + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_raw_u32));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (M + kBlockM - 1) / kBlockM;
  const int num_n = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + kBlockK - 1) / kBlockK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_a);
    ptx::prefetch_tensormap(&tm_b);
    if (SPLIT == 3) {
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
      ptx::mbar_init(tmem_empty_bar(a), 4);  // one arrive per epilogue warp
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc<T::kTmemCols>(tmem_slot);
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_idx = tile % num_m;
        const int n_idx = tile / num_m;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t a_dst = smem_base + stage * T::kStageBytes;
          const uint32_t b_dst = a_dst + T::kABytes;
          const uint32_t fb = full_bar(stage);
          ptx::mbar_arrive_expect_tx(fb, T::kStageBytes);
          // Activations stream through once per n-tile column; weights are re-read by every
          // m-tile: keep weights in L2 preferentially.
          ptx::tma_load_2d(a_dst, &tm_a, fb, kb * kBlockK, m_idx * kBlockM, ptx::kEvictNormal);
          ptx::tma_load_2d(b_dst, &tm_b, fb, kb * kBlockK, n_idx * BLOCK_N, ptx::kEvictLast);
          if (SPLIT == 3) {
            const uint32_t a_lo_dst = b_dst + T::kBBytes;
            const uint32_t b_lo_dst = a_lo_dst + T::kABytes;
            ptx::tma_load_2d(a_lo_dst, &tm_a_lo, fb, kb * kBlockK, m_idx * kBlockM, ptx::kEvictNormal);
            ptx::tma_load_2d(b_lo_dst, &tm_b_lo, fb, kb * kBlockK, n_idx * BLOCK_N, ptx::kEvictLast);
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
      constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(kBlockM, BLOCK_N);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1u);  // epilogue has drained this buffer
        ptx::tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tcgen05_fence_after();
          const uint32_t a_src = smem_base + stage * T::kStageBytes;
          const uint32_t b_src = a_src + T::kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            ptx::umma_bf16_ss(d_tmem, ptx::make_smem_desc_sw128(a_src + k * (kUmmaK * 2)),
                              ptx::make_smem_desc_sw128(b_src + k * (kUmmaK * 2)), idesc,
                              (kb | k) != 0 ? 1u : 0u);
          }
          if (SPLIT == 3) {
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
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_idx = tile % num_m;
      const int n_idx = tile / num_m;
      ptx::mbar_wait(tmem_full_bar(acc), acc_phase);
      ptx::tcgen05_fence_after();
      const int row = m_idx * kBlockM + quad * 32 + lane;
      const bool row_ok = row < M;
      long long orow = row;
      if (ep.row_group > 0)
        orow = static_cast<long long>(row / ep.row_group) * ep.row_stride + (row % ep.row_group) + ep.row_off;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        const int n0 = n_idx * BLOCK_N + c * 32;
        if (n0 >= N) break;  // warp-uniform
        uint32_t v[32];
        ptx::tmem_ld_32x32(t_row + c * 32, v);
        ptx::tmem_ld_wait();
        if (row_ok) {
          float x[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
          epilogue_row32(ep, x, row, orow, n0, N);
        }
      }
      ptx::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tmem_empty_bar(acc));
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
}

}  // namespace afft
