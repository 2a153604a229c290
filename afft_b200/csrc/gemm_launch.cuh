// Host-side launcher for gemm_bf16_tcgen05_kernel: builds the TMA tensor maps and picks the
// tile shape.  No libcuda link dependency: cuTensorMapEncodeTiled is resolved at run time
// through cudaGetDriverEntryPoint.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>

#include "gemm_sm100.cuh"
#include "gemm_skinny.cuh"

namespace afft {

struct GemmOperands {
  const __nv_bfloat16* a;     // [M,K] row pitch lda (elements)
  const __nv_bfloat16* a_lo;  // strict mode only
  long long lda;
  const __nv_bfloat16* w;     // [N,K] row pitch ldw
  const __nv_bfloat16* w_lo;  // strict mode only
  long long ldw;
  int M, N, K;
};

// Scratch of the split-K path (owned by a handle; nullptr = split-K off, e.g. the stateless afft_gemm entry).
struct SplitKScratch {
  float* partials;        // partial_floats fp32
  size_t partial_floats;
  unsigned* counters;     // n_counters, zeroed once at allocation (the kernel leaves them zero)
  int n_counters;
  int max_split;          // <= 1: off
};
constexpr size_t kSplitKPartialFloats = static_cast<size_t>(160) * kBlockM * 256;  // one 128 x 256 tile per CTA, 160 >= SMs
constexpr int kSplitKCounters = 160 * kNumEpilogueWarps;

// Number of K splits for a problem of `tiles` output tiles on `slots` persistent CTAs (or CTA pairs).  Cost model in
// units of one K block of main-loop time: a CTA runs ceil(tiles * S / slots) units back to back, each
// ceil(num_kb / S) K blocks plus a fixed pipeline fill / drain / epilogue (~6 K blocks, measured from the per-kernel
// floor of ~13 us at K = 1024); the last-arriving CTA of a tile re-reads S partial tiles (~1 K block each).  The
// smallest cost wins, ties go to fewer splits; S = 1 unless splitting saves at least 10 %.
inline int pick_ksplit(int tiles, int slots, int num_kb, int ctas_per_tile, size_t tile_floats, const SplitKScratch* sk) {
  static const int env = [] { const char* v = getenv("AFFT_GEMM_KSPLIT"); return v == nullptr ? 64 : atoi(v); }();
  if (sk == nullptr || sk->partials == nullptr || sk->counters == nullptr || env <= 1 || sk->max_split <= 1) return 1;
  if (tiles * ctas_per_tile * kNumEpilogueWarps > sk->n_counters) return 1;
  const double kFixed = 6.0, kReduce = 1.0;
  auto cost = [&](int s) {
    const int waves = (tiles * s + slots - 1) / slots;
    return waves * ((num_kb + s - 1) / s + kFixed) + (s > 1 ? kReduce * s : 0.0);
  };
  const double c1 = cost(1);
  int best = 1;
  double best_cost = c1;
  const int smax = std::min(std::min(env, sk->max_split), num_kb / 2);
  for (int s = 2; s <= smax; ++s) {
    if ((s - 1) * ((num_kb + s - 1) / s) >= num_kb) continue;  // would leave an empty split
    if (static_cast<size_t>(tiles) * s * ctas_per_tile * tile_floats > sk->partial_floats) break;
    const double c = cost(s);
    if (c < best_cost - 1e-9) {
      best_cost = c;
      best = s;
    }
  }
  return (best > 1 && best_cost <= 0.9 * c1) ? best : 1;
}

inline PFN_cuTensorMapEncodeTiled_v12000 get_tensormap_encoder(std::string* err) {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  static std::string init_err;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) {
      init_err = std::string("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: ") +
                 cudaGetErrorString(e);
      (void)cudaGetLastError();
    } else {
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
  });
  if (fn == nullptr && err != nullptr) *err = init_err;
  return fn;
}

// A tensor map is a pure function of (address, extents, pitch, box, type): the same operands come back forward after
// forward (weights and workspace buffers have stable addresses), so the encoded descriptors are memoised per thread -
// ~200 driver calls per forward otherwise.
struct TmapKey {
  const void* ptr;
  long long rows, cols, ld;
  int box_rows, kind;  // kind: 0 bf16 operand, 1 fp16 operand, 2.. epilogue maps
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && kind == o.kind;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    h ^= std::hash<long long>()(k.rows * 1000003LL + k.cols) + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
    h ^= std::hash<long long>()(k.ld * 131 + k.box_rows * 7 + k.kind) + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
    return h;
  }
};
inline std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash>& tmap_cache() {
  static thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  if (cache.size() > 8192) cache.clear();
  return cache;
}

// bf16 matrix [rows, cols] with row pitch ld (elements) -> 2-D tiled map, box = 64 x box_rows,
// SWIZZLE_128B, out-of-bounds elements read as zero.
inline bool make_tmap_bf16(CUtensorMap* out, const void* ptr, long long rows, long long cols, long long ld,
                           int box_rows, std::string* err, bool fp16 = false) {
  const TmapKey key{ptr, rows, cols, ld, box_rows, fp16 ? 1 : 0};
  auto& cache = tmap_cache();
  auto hit = cache.find(key);
  if (hit != cache.end()) {
    *out = hit->second;
    return true;
  }
  auto enc = get_tensormap_encoder(err);
  if (enc == nullptr) return false;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0 || (ld * 2) % 16 != 0) {
    if (err) *err = "TMA operand must be 16-B aligned with a 16-B multiple row pitch";
    return false;
  }
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = enc(out, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box,
                   estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r));
    return false;
  }
  cache.emplace(key, *out);
  return true;
}

// Generic 2-D tiled map (row-major [rows, cols], row pitch ld elements) for the TMA-staged epilogue: fp32 boxes of
// 32 x 32 elements (128-byte rows, SWIZZLE_128B) and 16-bit boxes of 32 x 32 (64-byte rows, SWIZZLE_64B).
inline bool make_tmap_epi(CUtensorMap* out, const void* ptr, long long rows, long long cols, long long ld, int elem_bytes,
                          bool fp16, std::string* err) {
  auto enc = get_tensormap_encoder(err);
  if (enc == nullptr) return false;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0 || (ld * elem_bytes) % 16 != 0) {
    if (err) *err = "TMA epilogue tensor must be 16-B aligned with a 16-B multiple row pitch";
    return false;
  }
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * elem_bytes};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estride[2] = {1, 1};
  const CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                                 : (fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  CUresult r = enc(out, dt, 2, const_cast<void*>(ptr), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   elem_bytes == 4 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled (epilogue) failed with CUresult " + std::to_string(static_cast<int>(r));
    return false;
  }
  return true;
}

// Which epilogue the 2-CTA kernel runs: 0 = v1 (smem transposition + coalesced LDG / STG; the default), 1 = v2 (TMA-staged,
// see gemm_sm100.cuh).  Process-wide; initial value from AFFT_GEMM_EPI_V2.
inline std::atomic<int>& gemm_epilogue_v2_flag() {
  static std::atomic<int> flag([] { const char* v = getenv("AFFT_GEMM_EPI_V2"); return v == nullptr ? 0 : atoi(v); }());
  return flag;
}

// Skinny problems (M <= kSkinnyMaxM rows, bf16 / fp16 operands) run gemm_skinny_kernel (gemm_skinny.cuh) instead of the
// tcgen05 kernels: 1 (default) = on, 0 = off.  Process-wide; initial value from AFFT_GEMM_SKINNY.
inline std::atomic<int>& gemm_skinny_flag() {
  static std::atomic<int> flag([] { const char* v = getenv("AFFT_GEMM_SKINNY"); return v == nullptr ? 1 : atoi(v); }());
  return flag;
}

inline bool pdl_enabled() {
  static const bool on = [] { const char* v = getenv("AFFT_PDL"); return v == nullptr || atoi(v) != 0; }();
  return on;
}

// Launch with the programmatic-stream-serialization attribute (the kernel must call griddepcontrol.wait before
// touching data produced by its predecessor).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

template <int BLOCK_N, int MODE, int EPI>
inline cudaError_t launch_gemm_variant(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tal,
                                       const CUtensorMap& tbl, const GemmEpilogue& ep, int M, int N, int K,
                                       int num_sms, GemmSched sched, const SplitKScratch* sk, cudaStream_t stream) {
  using T = GemmTraits<BLOCK_N, MODE>;
  auto kern = gemm_bf16_tcgen05_kernel<BLOCK_N, MODE, EPI>;
  static bool attr_set[64] = {false};  // per variant and device; benign race (idempotent)
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  const int num_tiles = ((M + kBlockM - 1) / kBlockM) * ((N + BLOCK_N - 1) / BLOCK_N);
  sched.ksplit = pick_ksplit(num_tiles, num_sms, (K + kBlockK - 1) / kBlockK, 1, static_cast<size_t>(kBlockM) * BLOCK_N, sk);
  sched.partials = sk ? sk->partials : nullptr;
  sched.counters = sk ? sk->counters : nullptr;
  const int num_units = num_tiles * sched.ksplit;
  const int grid = num_units < num_sms ? num_units : num_sms;
  return launch_pdl(kern, dim3(grid), dim3(kGemmThreads), T::kSmemBytes, stream, ta, tb, tal, tbl, ep, M, N, K, sched);
}

template <int MODE, int EPI>
inline cudaError_t launch_gemm_2cta_variant(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tal,
                                            const CUtensorMap& tbl, const GemmEpilogue& ep, int M, int N, int K,
                                            int num_sms, GemmSched sched, const SplitKScratch* sk, cudaStream_t stream,
                                            const GemmTmaEpi& tme) {
  using T = Gemm2Traits<MODE>;
  auto kern = gemm_bf16_tcgen05_2cta_kernel<MODE, EPI>;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  const int num_tiles = ((M + 255) / 256) * ((N + 255) / 256);
  int clusters = num_sms / 2;
  sched.ksplit = pick_ksplit(num_tiles, clusters, (K + kBlockK - 1) / kBlockK, 2, static_cast<size_t>(kBlockM) * 256, sk);
  sched.partials = sk ? sk->partials : nullptr;
  sched.counters = sk ? sk->counters : nullptr;
  if (num_tiles * sched.ksplit < clusters) clusters = num_tiles * sched.ksplit;
  // The TMA-staged epilogue needs whole tiles per unit; with split-K the v1 epilogue (partials + in-order sum) runs.
  // It trades ring depth for slab buffers: 5 stages + 8 KB per warp without a residual operand, 4 stages + 12 KB with.
  if (sched.ksplit > 1 || EPI < 0 || MODE == MODE_BF16X3) sched.v2_warp_bytes = 0;
  uint32_t smem = T::kSmemBytes;
  sched.stages = T::kStages;
  if (sched.v2_warp_bytes > 0) {
    sched.stages = (EPI & 4) != 0 ? 4 : 5;
    sched.v2_warp_bytes = (EPI & 4) != 0 ? 12288 : 8192;
    smem = T::smem_bytes(sched.stages, kNumEpilogueWarps * sched.v2_warp_bytes);
  }
  // __cluster_dims__(2) is compiled into the kernel
  return launch_pdl(kern, dim3(2 * clusters), dim3(kGemmThreads), smem, stream, ta, tb, tal, tbl, ep, M, N, K, sched, tme);
}

template <int MT, int NT, bool FP16>
inline cudaError_t launch_skinny_variant(const SkinnyArgs& p, cudaStream_t stream) {
  using T = SkinnyTraits<MT, NT>;
  auto kern = gemm_skinny_kernel<MT, NT, FP16>;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  const int grid = (p.N + 8 * NT - 1) / (8 * NT);
  return launch_pdl(kern, dim3(grid), dim3(kSkinnyThreads), T::kSmemBytes, stream, p);
}

// M <= 16 / 32 / 64 / 96 rows -> 1 / 2 / 4 / 6 m16 tiles; two n8 tiles per CTA once that still fills the SMs' CTA slots
// (two per SM up to 32 rows, one above: ncu showed the 6-tile variant at one CTA per SM running N = 4096 in 3.5 waves).
inline cudaError_t launch_skinny(const SkinnyArgs& p, bool fp16, int num_sms, cudaStream_t stream) {
  const int mt = p.M <= 16 ? 1 : (p.M <= 32 ? 2 : (p.M <= 64 ? 4 : 6));
  const int slots = (mt <= 2 ? 2 : 1) * num_sms;
  const int nt = ((p.N + 15) / 16 >= slots) ? 2 : 1;
#define AFFT_SKINNY(MT_, NT_)                                                      \
  if (mt == MT_ && nt == NT_)                                                      \
    return fp16 ? launch_skinny_variant<MT_, NT_, true>(p, stream) : launch_skinny_variant<MT_, NT_, false>(p, stream)
  AFFT_SKINNY(1, 1); AFFT_SKINNY(1, 2); AFFT_SKINNY(2, 1); AFFT_SKINNY(2, 2);
  AFFT_SKINNY(4, 1); AFFT_SKINNY(4, 2); AFFT_SKINNY(6, 1); AFFT_SKINNY(6, 2);
#undef AFFT_SKINNY
  return cudaErrorInvalidValue;
}

// Which problems take the skinny kernel (cost model in gemm_skinny.cuh): up to 32 rows always; up to 96 rows when the
// weight is small (<= 4 M elements: the fuser's at batch 1), where the launch floor of the tcgen05 kernel dominates.
inline bool skinny_eligible(int M, int N, int K) {
  static const int max_m = [] { const char* v = getenv("AFFT_GEMM_SKINNY_M"); return v == nullptr ? 32 : atoi(v); }();
  if (M > kSkinnyMaxM || K % 8 != 0) return false;
  return M <= max_m || static_cast<long long>(N) * K <= (4ll << 20);
}

inline int pick_block_n(int M, int N, int num_sms, int forced) {
  if (forced == 128 || forced == 256) return forced;
  // Fewest waves wins; ties go to the 256-wide tile (less A re-streaming through shared memory).
  const long long m_tiles = (M + kBlockM - 1) / kBlockM;
  const long long t256 = m_tiles * ((N + 255) / 256), t128 = m_tiles * ((N + 127) / 128);
  const long long w256 = (t256 + num_sms - 1) / num_sms, w128 = (t128 + num_sms - 1) / num_sms;
  // A 128-wide tile costs ~0.62 of a 256-wide one (measured: its main loop re-streams A through shared
  // memory twice as often per flop), so it only pays when it removes a mostly empty last wave.
  return (w128 * 0.62 < w256 * 1.0) ? 128 : 256;
}

inline unsigned long long policy_from_env(const char* name, unsigned long long dflt) {
  const char* v = getenv(name);
  if (v == nullptr) return dflt;
  if (v[0] == 'f') return ptx::kEvictFirst;
  if (v[0] == 'l') return ptx::kEvictLast;
  return ptx::kEvictNormal;
}

inline const GemmSched& default_sched() {
  static const GemmSched s = [] {
    GemmSched g;
    const char* o = getenv("AFFT_GEMM_ORDER");  // "m" = m-fastest, default n-fastest
    g.n_fastest = (o != nullptr && o[0] == 'm') ? 0 : 1;
    g.policy_a = policy_from_env("AFFT_GEMM_HINT_A", ptx::kEvictFirst);   // activations: streamed once
    g.policy_b = policy_from_env("AFFT_GEMM_HINT_B", ptx::kEvictNormal);  // weights: hot for the whole launch
    g.ksplit = 1;
    g.partials = nullptr;
    g.counters = nullptr;
    g.t_end = nullptr;
    g.stages = 0;
    g.v2_warp_bytes = 0;
    return g;
  }();
  return s;
}

// Returns false and fills *err on failure.  mode: MODE_BF16 / MODE_FP16 / MODE_BF16X3.  force_block_n: 0 = auto.
inline bool launch_gemm(const GemmOperands& g, const GemmEpilogue& ep, int mode, int force_block_n,
                        int num_sms, cudaStream_t stream, std::string* err, const SplitKScratch* sk = nullptr,
                        unsigned long long* t_end = nullptr, bool w_static = false) {
  if (mode != MODE_BF16 && mode != MODE_FP16 && mode != MODE_BF16X3) {
    if (err) *err = "gemm: unknown precision mode";
    return false;
  }
  const bool strict = (mode == MODE_BF16X3), fp16 = (mode == MODE_FP16);
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) {
    if (err) *err = "gemm: empty problem";
    return false;
  }
  if (strict && (g.a_lo == nullptr || g.w_lo == nullptr)) {
    if (err) *err = "gemm: strict mode needs hi/lo operands";
    return false;
  }
  auto misaligned = [](const void* p, long long ld, int elem) {
    return p != nullptr && (((reinterpret_cast<uintptr_t>(p) & 15u) != 0) || ((ld * elem) % 16 != 0));
  };
  if (misaligned(ep.out_f32, ep.ld_f32, 4) || misaligned(ep.res, ep.ld_res, 4) ||
      misaligned(ep.out_hi, ep.ld_bf16, 2) || misaligned(ep.out_lo, ep.ld_bf16, 2) ||
      (ep.bias != nullptr && (reinterpret_cast<uintptr_t>(ep.bias) & 15u) != 0)) {
    if (err) *err = "gemm: epilogue pointers must be 16-B aligned with 16-B multiple pitches";
    return false;
  }
  // skinny problems: the weight-streaming kernel (no TMEM / TMA / split-K machinery)
  if (!strict && force_block_n == 0 && skinny_eligible(g.M, g.N, g.K) && !misaligned(g.a, g.lda, 2) &&
      !misaligned(g.w, g.ldw, 2) && gemm_skinny_flag().load(std::memory_order_relaxed) != 0) {
    SkinnyArgs p{g.a, g.lda, g.w, g.ldw, g.M, g.N, g.K, ep, t_end, w_static ? 1 : 0};
    const cudaError_t se = launch_skinny(p, fp16, num_sms, stream);
    if (se != cudaSuccess) {
      if (err) *err = std::string("skinny gemm launch failed: ") + cudaGetErrorString(se);
      return false;
    }
    return true;
  }
  int bn = pick_block_n(g.M, g.N, num_sms, force_block_n);
  // CTA pairs (cta_group::2, 256 x 256 tiles) whenever the 256-wide tile is chosen and there is enough work
  static const int use_2cta_env = [] { const char* v = getenv("AFFT_GEMM_2CTA"); return v == nullptr ? 1 : atoi(v); }();
  const bool use_2cta = (force_block_n == 512) || (use_2cta_env != 0 && force_block_n == 0 && bn == 256 && g.M > 128);
  if (force_block_n == 512) bn = 256;
  const int b_box_rows = use_2cta ? 128 : bn;  // each CTA of a pair loads half of the 256-row weight tile
  CUtensorMap ta, tb, tal, tbl;
  if (!make_tmap_bf16(&ta, g.a, g.M, g.K, g.lda, kBlockM, err, fp16)) return false;
  if (!make_tmap_bf16(&tb, g.w, g.N, g.K, g.ldw, b_box_rows, err, fp16)) return false;
  if (strict) {
    if (!make_tmap_bf16(&tal, g.a_lo, g.M, g.K, g.lda, kBlockM, err)) return false;
    if (!make_tmap_bf16(&tbl, g.w_lo, g.N, g.K, g.ldw, b_box_rows, err)) return false;
  } else {
    tal = ta;
    tbl = tb;
  }
  // epilogue variant: the combinations the forward path uses are compiled with constant flags
  const int code = epi_code(ep.act, ep.res != nullptr, ep.out_f32 != nullptr, ep.out_hi != nullptr);
  const bool lo_ok = (ep.out_hi == nullptr) || ((ep.out_lo != nullptr) == strict);  // specialised kernels tie lo to MODE
  cudaError_t e = cudaErrorInvalidValue;
  GemmSched sched0 = default_sched();
  sched0.t_end = t_end;
  // TMA-staged epilogue (2-CTA kernel, compiled epilogue variants, plain row mapping): build its tensor maps
  const int v2_env = gemm_epilogue_v2_flag().load(std::memory_order_relaxed);
  GemmTmaEpi tme;
  memset(&tme, 0, sizeof(tme));
  const bool both_out = ep.out_f32 != nullptr && ep.out_hi != nullptr;
  bool v2 = v2_env != 0 && use_2cta && !strict && lo_ok && ep.act <= ACT_GELU_TANH && ep.row_group == 0 && ep.res_mod == 0 &&
            ep.out_lo == nullptr && (!both_out || ep.res != nullptr) && g.N % 32 == 0;
  if (v2) {
    switch (code) {  // the variants compiled below
      case epi_code(ACT_NONE, false, false, true): case epi_code(ACT_NONE, false, true, false):
      case epi_code(ACT_NONE, true, true, false): case epi_code(ACT_GELU_ERF, false, false, true):
      case epi_code(ACT_GELU_TANH, false, false, true): case epi_code(ACT_NONE, true, true, true): break;
      default: v2 = false;
    }
  }
  if (v2) {
    if (ep.res != nullptr && !make_tmap_epi(&tme.res, ep.res, g.M, g.N, ep.ld_res, 4, false, err)) return false;
    if (ep.out_f32 != nullptr && !make_tmap_epi(&tme.f32, ep.out_f32, g.M, g.N, ep.ld_f32, 4, false, err)) return false;
    if (ep.out_hi != nullptr && !make_tmap_epi(&tme.b16, ep.out_hi, g.M, g.N, ep.ld_bf16, 2, fp16, err)) return false;
    sched0.v2_warp_bytes = 1;  // the variant launcher sizes it
  }
#define AFFT_LAUNCH(BN, SP, EP)                                                                                       \
  e = (BN == 512) ? launch_gemm_2cta_variant<SP, EP>(ta, tb, tal, tbl, ep, g.M, g.N, g.K, num_sms, sched0, sk, stream, tme) \
                  : launch_gemm_variant<(BN == 512 ? 256 : BN), SP, EP>(ta, tb, tal, tbl, ep, g.M, g.N, g.K, num_sms, sched0, sk, stream)
#define AFFT_DISPATCH_EPI(BN, SP)                                                            \
  do {                                                                                       \
    if (!lo_ok || ep.act > ACT_GELU_TANH) { AFFT_LAUNCH(BN, SP, EPI_GENERIC); break; }                               \
    switch (code) {                                                                          \
      case epi_code(ACT_NONE, false, false, true): AFFT_LAUNCH(BN, SP, epi_code(ACT_NONE, false, false, true)); break;          \
      case epi_code(ACT_NONE, false, true, false): AFFT_LAUNCH(BN, SP, epi_code(ACT_NONE, false, true, false)); break;          \
      case epi_code(ACT_NONE, true, true, false): AFFT_LAUNCH(BN, SP, epi_code(ACT_NONE, true, true, false)); break;            \
      case epi_code(ACT_GELU_ERF, false, false, true): AFFT_LAUNCH(BN, SP, epi_code(ACT_GELU_ERF, false, false, true)); break;  \
      case epi_code(ACT_GELU_TANH, false, false, true): AFFT_LAUNCH(BN, SP, epi_code(ACT_GELU_TANH, false, false, true)); break; \
      case epi_code(ACT_NONE, false, true, true): AFFT_LAUNCH(BN, SP, epi_code(ACT_NONE, false, true, true)); break;            \
      case epi_code(ACT_NONE, true, true, true): AFFT_LAUNCH(BN, SP, epi_code(ACT_NONE, true, true, true)); break;              \
      default: AFFT_LAUNCH(BN, SP, EPI_GENERIC); break;                                      \
    }                                                                                        \
  } while (0)
  if (strict) {
    if (use_2cta) AFFT_DISPATCH_EPI(512, MODE_BF16X3); else if (bn == 256) AFFT_DISPATCH_EPI(256, MODE_BF16X3); else AFFT_DISPATCH_EPI(128, MODE_BF16X3);
  } else if (fp16) {
    if (use_2cta) AFFT_DISPATCH_EPI(512, MODE_FP16); else if (bn == 256) AFFT_DISPATCH_EPI(256, MODE_FP16); else AFFT_DISPATCH_EPI(128, MODE_FP16);
  } else {
    if (use_2cta) AFFT_DISPATCH_EPI(512, MODE_BF16); else if (bn == 256) AFFT_DISPATCH_EPI(256, MODE_BF16); else AFFT_DISPATCH_EPI(128, MODE_BF16);
  }
#undef AFFT_DISPATCH_EPI
#undef AFFT_LAUNCH
  if (e != cudaSuccess) {
    if (err) *err = std::string("gemm launch failed: ") + cudaGetErrorString(e);
    return false;
  }
  return true;
}

}  // namespace afft
