// Skinny GEMM for the small-batch end of the path: y [M, N] = A [M, K] . W [N, K]^T with M <= 96 rows (batch 1 of the headline
// config has 90 fuser rows and 18 GPT-2 rows; reference test.py evaluates with small batches, SURVEY.md section 8 row e).
//
// At these sizes a GEMM is a weight stream: 2 N K bytes from HBM against 2 M N K flops, i.e. <= 96 flop per byte - far below
// the tensor-core ridge - and the tcgen05 kernel's fixed cost (TMEM allocation, tensor-map fetch, TMA ring fill, split-K
// reduction through global memory: ~13 us per launch, profiles/r01_small_batch_launches.txt) is what a forward at batch 1
// pays 52 times.  This kernel has no fixed machinery: warp-level mma.sync (m16n8k16, fp32 accumulate) fed straight from a
// per-thread cp.async FIFO.
//
//   * a CTA owns 8 * NT weight rows (one or two n8 tiles) and the whole K; its 8 warps split K in 32-element chunks
//     (chunk c -> warp c % 8, so the CTA reads 512 contiguous bytes of every row per round) and are summed through shared
//     memory in warp order (bit-reproducible);
//   * every lane copies 16 contiguous bytes (8 k values) of "its" weight row (n = lane / 4) and of its two A rows per m16
//     tile (lane / 4 and lane / 4 + 8) per chunk: a warp-level copy is 8 rows x 64 B - whole 32-byte sectors, no shared-memory
//     transposition.  The 8 values feed TWO mma k-steps: the mma contracts over a k index whose order is irrelevant as long
//     as A and B agree, so the lane's elements (0,1),(2,3) serve as the k pairs (2t, 2t+1), (2t+8, 2t+9) of the first mma and
//     (4,5),(6,7) of the second - for the A fragment and the B fragment alike;
//   * each thread reads back only the bytes it copied itself, so the FIFO needs no barrier at all: cp.async.wait_group is
//     per-thread.  Depth 3 - 8 stages (shared-memory budget);
//   * rows >= M / >= N and k >= K are zero-filled by cp.async's src-size operand: no tails in the main loop;
//   * the epilogue is the generic one of the tcgen05 kernel (bias, GELU / ReLU / gate, residual with row maps, fp32 and
//     16-bit outputs), applied to (row, column pair) items after the cross-warp sum.
//
// Cost model and limits (measured, profiles/r02_skinny_gemm.txt): every CTA re-reads A from L2, M K 2 bytes per 8 NT weight
// rows, i.e. M / (8 NT) bytes of L2 traffic per weight byte.  That is why the launcher uses this kernel for M <= 32 rows, and
// up to 96 rows only for weights of at most 4 M elements (the fuser's at batch 1); the big GPT-2 weights at 36 - 90 rows
// (batches 2 - 5) stay on the tcgen05 split-K path, which is faster there.  Two restructurings were built and measured
// slower than this version: staging A once per CTA in shared memory with K spans across CTAs and a last-arriver reduction
// (shorter per-CTA streams, an extra wave and the reduction chain cost more than the A re-reads), and loading the A fragments
// straight from L2 into registers with a weights-only FIFO (one L2 round trip per 32-element chunk on every warp's chain).
#pragma once

#include "gemm_sm100.cuh"

namespace afft {

struct SkinnyArgs {
  const void* a;  // 16-bit [M, K], row pitch lda elements
  long long lda;
  const void* w;  // 16-bit [N, K], row pitch ldw elements
  long long ldw;
  int M, N, K;
  GemmEpilogue ep;
  unsigned long long* t_end;
  int w_static;  // the weight is not written by the preceding kernels of the stream: its prefetch may start before
                 // griddepcontrol.wait, i.e. under the tail of the previous kernel (engine calls; not the stateless entry)
};

constexpr int kSkinnyThreads = 256;
constexpr int kSkinnyWarps = kSkinnyThreads / 32;
constexpr int kSkinnyMaxM = 96;

template <int MT, int NT>
struct SkinnyTraits {
  static constexpr int kSlots = NT + 2 * MT;  // 16-byte slots per thread and stage: NT weight rows, 2 A rows per m16 tile
  static constexpr int kSlotBytes = 16 * kSkinnyThreads;
  static constexpr int kStageBytes = kSlots * kSlotBytes;
  static constexpr int kCtasPerSm = MT <= 2 ? 2 : 1;
  static constexpr int kBudget = MT <= 2 ? 108 * 1024 : 200 * 1024;
  static constexpr int kStagesRaw = kBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : (kStagesRaw < 2 ? 2 : kStagesRaw);
  static constexpr int kRedBytes = kSkinnyWarps * MT * NT * 32 * 16;  // one float4 per lane, tile and warp
  static constexpr int kSmemBytes = kStages * kStageBytes > kRedBytes ? kStages * kStageBytes : kRedBytes;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
};

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
template <bool FP16>
__device__ __forceinline__ void skinny_mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                           uint32_t b1) {
  if constexpr (FP16) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
}

__device__ __forceinline__ float skinny_act(float x, int act) {
  switch (act) {
    case ACT_GELU_ERF: return gelu_erf(x);
    case ACT_GELU_TANH: return gelu_tanh(x);
    case ACT_RELU: return fmaxf(x, 0.f);
    case ACT_GATE: return 1.0f / (1.0f + __expf(-x));
    default: return x;
  }
}

// one output element: bias, activation, residual (same conventions as epilogue_slab_edge), stores
template <bool FP16>
__device__ __forceinline__ void skinny_store(const GemmEpilogue& ep, int row, int col, float v) {
  if (ep.bias != nullptr) v += __ldg(ep.bias + col);
  v = skinny_act(v, ep.act);
  long long orow = row;
  if (ep.row_group > 0) orow = static_cast<long long>(row / ep.row_group) * ep.row_stride + (row % ep.row_group) + ep.row_off;
  if (ep.res != nullptr) {
    const long long rrow = (ep.res_mod > 0) ? static_cast<long long>(row % ep.res_mod) : orow;
    const float r = ep.res[rrow * ep.ld_res + col];
    v = (ep.act == ACT_GATE) ? v * r : v + r;
  }
  if (ep.out_f32 != nullptr) ep.out_f32[orow * ep.ld_f32 + col] = v;
  if (ep.out_hi != nullptr) {
    if (FP16) {
      reinterpret_cast<__half*>(ep.out_hi)[orow * ep.ld_bf16 + col] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    } else {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      ep.out_hi[orow * ep.ld_bf16 + col] = h;
      if (ep.out_lo != nullptr) ep.out_lo[orow * ep.ld_bf16 + col] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

template <int MT, int NT, bool FP16>
__global__ void __launch_bounds__(kSkinnyThreads, SkinnyTraits<MT, NT>::kCtasPerSm) gemm_skinny_kernel(const SkinnyArgs p) {
  using T = SkinnyTraits<MT, NT>;
  constexpr int S = T::kStages;
  extern __shared__ __align__(16) uint8_t skinny_smem[];
  ptx::griddep_launch();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * (8 * NT);
  const int n_chunks = (p.K + 31) >> 5;
  const int my_chunks = n_chunks > warp ? (n_chunks - warp + kSkinnyWarps - 1) / kSkinnyWarps : 0;
  const uint32_t fifo = ptx::smem_u32(skinny_smem) + static_cast<uint32_t>(tid) * 16u;

  const uint8_t* wrow[NT];
  bool wok[NT];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int n = n0 + nt * 8 + g;
    wok[nt] = n < p.N;
    wrow[nt] = static_cast<const uint8_t*>(p.w) + static_cast<long long>(wok[nt] ? n : 0) * p.ldw * 2;
  }
  const uint8_t* arow[MT][2];
  bool aok[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = mt * 16 + g + h * 8;
      aok[mt][h] = m < p.M;
      arow[mt][h] = static_cast<const uint8_t*>(p.a) + static_cast<long long>(aok[mt][h] ? m : 0) * p.lda * 2;
    }

  // the i-th chunk of this warp -> stage i % S; K is a multiple of 8: a lane's 8 values are inside or outside K together
  auto issue_w = [&](int i) {
    const int kc = (warp + kSkinnyWarps * i) * 32 + t * 8;
    const bool kok = kc < p.K;
    const long long koff = kok ? static_cast<long long>(kc) * 2 : 0;
    const uint32_t st = fifo + static_cast<uint32_t>(i % S) * T::kStageBytes;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) cp_async16_zfill(st + nt * T::kSlotBytes, wrow[nt] + koff, (wok[nt] && kok) ? 16u : 0u);
  };
  auto issue_a = [&](int i) {
    const int kc = (warp + kSkinnyWarps * i) * 32 + t * 8;
    const bool kok = kc < p.K;
    const long long koff = kok ? static_cast<long long>(kc) * 2 : 0;
    const uint32_t st = fifo + static_cast<uint32_t>(i % S) * T::kStageBytes;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h)
        cp_async16_zfill(st + (NT + mt * 2 + h) * T::kSlotBytes, arow[mt][h] + koff, (aok[mt][h] && kok) ? 16u : 0u);
  };
  // static weights: the first S - 1 chunks are requested while the previous kernel of the stream is still draining
  const bool early_w = p.w_static != 0;
  if (early_w) {
#pragma unroll
    for (int i = 0; i < S - 1; ++i)
      if (i < my_chunks) issue_w(i);
  }
  ptx::griddep_wait();  // A, the residual operand and the outputs belong to the predecessors

  float acc[MT][NT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;

#pragma unroll
  for (int i = 0; i < S - 1; ++i) {
    if (i < my_chunks) {
      if (!early_w) issue_w(i);
      issue_a(i);
    }
    cp_async_commit();  // with early_w the first group also carries the weight copies of all S - 1 chunks
  }
  for (int i = 0; i < my_chunks; ++i) {
    if (i + S - 1 < my_chunks) {
      issue_w(i + S - 1);
      issue_a(i + S - 1);
    }
    cp_async_commit();
    cp_async_wait<S - 1>();  // groups 0 .. i are complete
    const uint32_t st = fifo + static_cast<uint32_t>(i % S) * T::kStageBytes;
    uint4 wf[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) wf[nt] = lds128u(st + nt * T::kSlotBytes);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const uint4 a0 = lds128u(st + (NT + mt * 2) * T::kSlotBytes);      // row g
      const uint4 a1 = lds128u(st + (NT + mt * 2 + 1) * T::kSlotBytes);  // row g + 8
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        skinny_mma<FP16>(acc[mt][nt], a0.x, a1.x, a0.y, a1.y, wf[nt].x, wf[nt].y);
        skinny_mma<FP16>(acc[mt][nt], a0.z, a1.z, a0.w, a1.w, wf[nt].z, wf[nt].w);
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();  // every warp is done with the FIFO: its memory becomes the reduction buffer

  float4* red = reinterpret_cast<float4*>(skinny_smem);
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
      red[((warp * MT + mt) * NT + nt) * 32 + lane] = make_float4(acc[mt][nt][0], acc[mt][nt][1], acc[mt][nt][2], acc[mt][nt][3]);
  __syncthreads();
  // item = (m tile, n tile, lane): the four accumulator values of that lane, summed over the warps in warp order
  for (int item = tid; item < MT * NT * 32; item += kSkinnyThreads) {
    float4 s = red[item];
#pragma unroll
    for (int w = 1; w < kSkinnyWarps; ++w) {
      const float4 v = red[w * MT * NT * 32 + item];
      s.x += v.x;
      s.y += v.y;
      s.z += v.z;
      s.w += v.w;
    }
    const int il = item & 31, tile = item >> 5, nt = tile % NT, mt = tile / NT;
    const int row = mt * 16 + (il >> 2), col = n0 + nt * 8 + (il & 3) * 2;
    if (row < p.M) {
      if (col < p.N) skinny_store<FP16>(p.ep, row, col, s.x);
      if (col + 1 < p.N) skinny_store<FP16>(p.ep, row, col + 1, s.y);
    }
    if (row + 8 < p.M) {
      if (col < p.N) skinny_store<FP16>(p.ep, row + 8, col, s.z);
      if (col + 1 < p.N) skinny_store<FP16>(p.ep, row + 8, col + 1, s.w);
    }
  }
  ptx::prof_mark_end(p.t_end);
}

}  // namespace afft
