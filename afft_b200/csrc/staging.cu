// Input side of the hot path (include/afft_staging.h): the batch PLAN on the host (integer / float64 arithmetic of
// the reference's sampler and LMDB reader, no feature bytes touched) and the device GATHER that builds the
// (B, T, C_m) fp32 batch straight from the row tables (pinned host memory read over PCIe, or HBM).
//
// Reference: datasets/base_video_dataset.py:225-337 (_sample), datasets/reader_fns.py:65-157
// (read_representations, _read_rulstm_features, _convert_to_orig_video_fps), test.py:81 (H2D).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/afft_staging.h"

namespace {

constexpr int kSearchRadius = 10;  // reader_fns.py:76  `for search_radius in range(10)`

struct VideoIndex {
  std::vector<int32_t> frames;  // ascending
  int64_t first_row = 0;
};

struct Modality {
  int width = 0;
  bool orig_fps = false;
  const float* rows = nullptr;
  int64_t n_rows = 0;
  bool rows_on_device = false;
  std::unordered_map<std::string, VideoIndex> videos;
};

// Python's round() / numpy's rint: to nearest, ties to even.  nearbyint follows the current rounding mode, which is
// round-to-nearest-even unless someone changed it; the library never does.
inline long long round_half_even(double x) { return static_cast<long long>(std::nearbyint(x)); }

// reader_fns.py:148-157
inline bool orig_video_fps(const std::string& name, double* out) {
  const size_t p = name.rfind('_');
  const size_t n = (p == std::string::npos) ? name.size() : name.size() - p - 1;
  if (n == 3) { *out = 50.0; return true; }
  if (n == 2) { *out = 59.94005994005994; return true; }
  return false;
}

}  // namespace

struct afft_feature_store {
  std::vector<Modality> mods;
  std::string err;
  bool allow_empty_clips = false;  // false: refuse clips without any stored frame in their window (reference behaviour)
};

static thread_local std::string g_stage_err;

static int sfail(afft_feature_store* s, int code, const std::string& msg) {
  if (s != nullptr) s->err = msg;
  g_stage_err = msg;
  return code;
}

extern "C" const char* afft_store_error(const afft_feature_store* s) { return s ? s->err.c_str() : g_stage_err.c_str(); }

extern "C" int afft_store_create(int32_t n_mod, const int32_t* widths, const int32_t* orig_fps_index, afft_feature_store** out) {
  if (out == nullptr || widths == nullptr || n_mod < 1 || n_mod > AFFT_MAX_MODS) return sfail(nullptr, AFFT_ERR_INVALID, "store_create: bad arguments");
  auto* s = new afft_feature_store;
  s->mods.resize(n_mod);
  for (int m = 0; m < n_mod; ++m) {
    if (widths[m] < 4 || widths[m] % 4 != 0) {
      delete s;
      return sfail(nullptr, AFFT_ERR_INVALID, "store_create: row widths must be positive multiples of 4 (16-byte rows)");
    }
    s->mods[m].width = widths[m];
    s->mods[m].orig_fps = orig_fps_index != nullptr && orig_fps_index[m] != 0;
  }
  *out = s;
  return AFFT_OK;
}

extern "C" void afft_store_destroy(afft_feature_store* s) { delete s; }

extern "C" int afft_store_add_video(afft_feature_store* s, int32_t mod, const char* video_name, const int32_t* frame_ids,
                                    int64_t n, int64_t first_row) {
  if (s == nullptr) return sfail(nullptr, AFFT_ERR_INVALID, "store_add_video: null store");
  if (mod < 0 || mod >= static_cast<int>(s->mods.size()) || video_name == nullptr || n < 0 || first_row < 0 || (n > 0 && frame_ids == nullptr))
    return sfail(s, AFFT_ERR_INVALID, "store_add_video: bad arguments");
  for (int64_t i = 1; i < n; ++i)
    if (frame_ids[i] <= frame_ids[i - 1]) return sfail(s, AFFT_ERR_INVALID, std::string("store_add_video: frame ids of ") + video_name + " are not strictly ascending");
  VideoIndex& v = s->mods[mod].videos[video_name];
  v.frames.assign(frame_ids, frame_ids + n);
  v.first_row = first_row;
  return AFFT_OK;
}

extern "C" int afft_store_set_rows(afft_feature_store* s, int32_t mod, const void* rows, int64_t n_rows) {
  if (s == nullptr) return sfail(nullptr, AFFT_ERR_INVALID, "store_set_rows: null store");
  if (mod < 0 || mod >= static_cast<int>(s->mods.size()) || rows == nullptr || n_rows < 0) return sfail(s, AFFT_ERR_INVALID, "store_set_rows: bad arguments");
  if ((reinterpret_cast<uintptr_t>(rows) & 15u) != 0) return sfail(s, AFFT_ERR_INVALID, "store_set_rows: the row table must be 16-byte aligned");
  s->mods[mod].rows = static_cast<const float*>(rows);
  s->mods[mod].n_rows = n_rows;
  // where the table lives decides the gather's grid (no device needed to build a plan-only store: errors are ignored)
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, rows) == cudaSuccess) s->mods[mod].rows_on_device = (attr.type == cudaMemoryTypeDevice);
  (void)cudaGetLastError();
  return AFFT_OK;
}

namespace {

// The T frame ids (30-fps numbering) the model sees for one clip, in model order; false = reader_fns.py:122 assertion.
// Every operation mirrors the reference's float64 / Python-int arithmetic in the same order.
// AFFT_SAMPLE_RANDOM: rand_start_frame / rand_offset are the two numbers the reference draws per (clip, modality) call
// (base_video_dataset.py:246-248 rng.integers, :283-286 random.random()); the caller supplies them.
bool clip_frames(double start, double end, double fps, int T, double frame_rate, int strategy, std::vector<long long>* out,
                 std::vector<long long>* window_scratch, long long rand_start_frame = 0, long long rand_offset = 0) {
  start = std::max(start, 0.0);                                         // base_video_dataset.py:236-237
  end = std::max(end, 0.0);
  const double req_fps = frame_rate > 0 ? frame_rate : fps;            // :238-240
  const long long nframes = static_cast<long long>(fps * (end - start));                 // int(): truncation, argument >= 0 or tiny
  const long long frames_to_ext = round_half_even(T * (fps / req_fps));                  // :242
  long long start_frame;
  if (strategy == AFFT_SAMPLE_CENTER) {
    const long long d = nframes - frames_to_ext;                       // Python // floors; max(.., 0) hides the sign difference
    start_frame = d > 0 ? d / 2 : 0;
  } else if (strategy == AFFT_SAMPLE_LAST) {
    start_frame = std::max(nframes - frames_to_ext, 0LL);
  } else if (strategy == AFFT_SAMPLE_RANDOM) {
    start_frame = rand_start_frame;                                     // :246-248 (drawn in [0, max(nframes - frames_to_ext, 0)))
  } else {
    start_frame = 0;
  }
  double new_start = start + std::max(static_cast<double>(start_frame) / fps, 0.0);       // :258
  double new_end = start + std::max(static_cast<double>(start_frame + frames_to_ext) / fps, 0.0);
  new_end = std::max(std::min(end, new_end), 0.0);
  new_start = std::min(std::max(new_start, 0.0), new_end);
  // reader_fns.py:116-124
  const long long sf = static_cast<long long>(std::floor(new_start * fps));
  const long long ef = static_cast<long long>(std::floor(new_end * fps));
  if (ef <= sf || ef < 1) return false;  // empty window, or no id >= 1
  std::vector<long long>& w = *window_scratch;
  w.clear();
  const long long min_valid = std::max(sf + 1, 1LL);
  for (long long f = sf + 1; f <= ef; ++f) w.push_back(f < 1 ? min_valid : f);
  // base_video_dataset.py:279-292 then :309-335
  const long long n = static_cast<long long>(w.size());
  const long long step = std::max(round_half_even(fps / req_fps), 1LL);
  std::vector<long long> keep;
  const bool from_back = (strategy == AFFT_SAMPLE_LAST || strategy == AFFT_SAMPLE_RANDOM);
  if (from_back) {
    for (long long i = n - 1; i >= 0; i -= step) keep.push_back(i);
    std::reverse(keep.begin(), keep.end());
    if (strategy == AFFT_SAMPLE_RANDOM)                                  // :283-287
      for (auto& i : keep)
        if (i - rand_offset > 0) i -= rand_offset;
  } else {
    for (long long i = 0; i < n; i += step) keep.push_back(i);
  }
  out->clear();
  const long long have = static_cast<long long>(keep.size());
  if (from_back) {
    for (long long i = 0; i < T - have; ++i) out->push_back(w[keep.front()]);
    for (long long i = std::max(have - T, 0LL); i < have; ++i) out->push_back(w[keep[i]]);
  } else {
    for (long long i = 0; i < std::min<long long>(have, T); ++i) out->push_back(w[keep[i]]);
    for (long long i = have; i < T; ++i) out->push_back(w[keep.back()]);
  }
  return true;
}

// reader_fns.py:71-91: row of `frame` or of the closest earlier stored frame within the radius; -1 = zero row.
inline int64_t lookup_row(const VideoIndex& v, long long frame) {
  auto it = std::upper_bound(v.frames.begin(), v.frames.end(), static_cast<int32_t>(std::min<long long>(frame, INT32_MAX)));
  if (it == v.frames.begin()) return -1;
  --it;
  if (frame - *it >= kSearchRadius) return -1;
  return v.first_row + (it - v.frames.begin());
}

}  // namespace

extern "C" int afft_store_allow_empty_clips(afft_feature_store* s, int32_t allow) {
  if (s == nullptr) return sfail(nullptr, AFFT_ERR_INVALID, "store_allow_empty_clips: null store");
  s->allow_empty_clips = allow != 0;
  return AFFT_OK;
}

static int store_plan_impl(afft_feature_store* s, int32_t B, const char* const* video_names, const double* start_sec,
                           const double* end_sec, double fps, int32_t T, double frame_rate, int32_t strategy,
                           const int64_t* rand_start_frame, const int32_t* rand_offset, int32_t* row_idx, int32_t* frame_ids_out) {
  if (s == nullptr) return sfail(nullptr, AFFT_ERR_INVALID, "store_plan: null store");
  if (B < 0 || T < 1 || video_names == nullptr || start_sec == nullptr || end_sec == nullptr || row_idx == nullptr || !(fps > 0))
    return sfail(s, AFFT_ERR_INVALID, "store_plan: bad arguments");
  const bool random = (strategy == AFFT_SAMPLE_RANDOM);
  if (strategy != AFFT_SAMPLE_LAST && strategy != AFFT_SAMPLE_CENTER && strategy != AFFT_SAMPLE_FIRST && !random)
    return sfail(s, AFFT_ERR_INVALID, "store_plan: unknown sampling strategy");
  if (random && (rand_start_frame == nullptr || rand_offset == nullptr))
    return sfail(s, AFFT_ERR_INVALID, "store_plan: AFFT_SAMPLE_RANDOM needs the drawn start frames and offsets (afft_store_plan_random)");
  const int n_mod = static_cast<int>(s->mods.size());
  const int n_threads = B >= 64 ? static_cast<int>(std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency()))) : 1;
  std::vector<std::string> errs(n_threads);
  auto work = [&](int tid) {
    std::vector<long long> frames, window;
    for (int b = tid; b < B; b += n_threads) {
      if (video_names[b] == nullptr) { errs[tid] = "store_plan: null video name"; return; }
      const std::string name(video_names[b]);
      if (!random && !clip_frames(start_sec[b], end_sec[b], fps, T, frame_rate, strategy, &frames, &window)) {
        errs[tid] = "store_plan: clip " + std::to_string(b) + " (" + name + ") has no frame id >= 1 in its window (reader_fns.py:122)";
        return;
      }
      for (int m = 0; m < n_mod; ++m) {
        // random_clip: the reference calls _sample once per modality, each call with its own two draws
        if (random) {
          const size_t di = static_cast<size_t>(m) * B + b;
          if (rand_start_frame[di] < 0 || rand_offset[di] < 0) { errs[tid] = "store_plan: negative random draw"; return; }
          if (!clip_frames(start_sec[b], end_sec[b], fps, T, frame_rate, strategy, &frames, &window, rand_start_frame[di],
                           rand_offset[di])) {
            errs[tid] = "store_plan: clip " + std::to_string(b) + " (" + name + ") has no frame id >= 1 in its window (reader_fns.py:122)";
            return;
          }
        }
        const Modality& mod = s->mods[m];
        auto vit = mod.videos.find(name);
        if (vit == mod.videos.end()) { errs[tid] = "store_plan: video " + name + " is not in modality " + std::to_string(m); return; }
        double ofps = 0.0;
        if (mod.orig_fps && !orig_video_fps(name, &ofps)) { errs[tid] = "Unkown video name format: " + name; return; }
        int32_t* ri = row_idx + (static_cast<size_t>(m) * B + b) * T;
        int32_t* fo = frame_ids_out ? frame_ids_out + (static_cast<size_t>(m) * B + b) * T : nullptr;
        for (int t = 0; t < T; ++t) {
          long long f = frames[t];
          if (mod.orig_fps) f = round_half_even(static_cast<double>(f) / fps * ofps);  // reader_fns.py:143-145
          const int64_t row = lookup_row(vit->second, f);
          if (row >= mod.n_rows && mod.rows != nullptr) { errs[tid] = "store_plan: index of " + name + " points past the row table"; return; }
          ri[t] = static_cast<int32_t>(row);
          if (fo) fo[t] = static_cast<int32_t>(f);
        }
        // The reference asserts that at least one frame of the window is stored (`assert len(features_not_none) > 0`,
        // reader_fns.py:97): a clip with no stored frame near any of its T kept frames is a data problem, not zeros.
        bool any_row = false;
        for (int t = 0; t < T; ++t) any_row |= (ri[t] >= 0);
        if (!any_row && !s->allow_empty_clips) {  // rare: look at every frame of the window, as the reference's reader does before it subsamples
          for (size_t wi = 0; wi < window.size() && !any_row; ++wi) {
            long long f = window[wi];
            if (mod.orig_fps) f = round_half_even(static_cast<double>(f) / fps * ofps);
            any_row = lookup_row(vit->second, f) >= 0;
          }
        }
        if (!any_row && !s->allow_empty_clips) {
          errs[tid] = "store_plan: clip " + std::to_string(b) + " (" + name + ") has no stored frame in modality " +
                      std::to_string(m) + " within its window (reader_fns.py:97 asserts)";
          return;
        }
      }
    }
  };
  if (n_threads == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int i = 0; i < n_threads; ++i) th.emplace_back(work, i);
    for (auto& t : th) t.join();
  }
  for (auto& e : errs)
    if (!e.empty()) return sfail(s, AFFT_ERR_INVALID, e);
  return AFFT_OK;
}

extern "C" int afft_store_plan(afft_feature_store* s, int32_t B, const char* const* video_names, const double* start_sec,
                               const double* end_sec, double fps, int32_t T, double frame_rate, int32_t strategy,
                               int32_t* row_idx, int32_t* frame_ids_out) {
  if (strategy == AFFT_SAMPLE_RANDOM)
    return sfail(s, AFFT_ERR_INVALID, "store_plan: AFFT_SAMPLE_RANDOM needs the drawn numbers: use afft_store_plan_random");
  return store_plan_impl(s, B, video_names, start_sec, end_sec, fps, T, frame_rate, strategy, nullptr, nullptr, row_idx, frame_ids_out);
}

extern "C" int afft_store_plan_random(afft_feature_store* s, int32_t B, const char* const* video_names, const double* start_sec,
                                      const double* end_sec, double fps, int32_t T, double frame_rate,
                                      const int64_t* rand_start_frame, const int32_t* rand_offset, int32_t* row_idx,
                                      int32_t* frame_ids_out) {
  return store_plan_impl(s, B, video_names, start_sec, end_sec, fps, T, frame_rate, AFFT_SAMPLE_RANDOM, rand_start_frame,
                         rand_offset, row_idx, frame_ids_out);
}

// ------------------------------------------------------------------------------------------------
// Device gather.  One warp per output row (clip b, step t) of one modality; lanes stride the row in 16-byte
// vectors, all loads of a row in flight before the first store.  Source rows are 1.4 - 4 KB contiguous, so reads from
// pinned host memory are full PCIe bursts; -1 rows are written as zeros (reader_fns.py:95).
// Bound: PCIe (pinned store) or HBM (device store); algorithmic bytes = 2 * B * T * C * 4.
// Grid: the gather runs on a side stream next to the forward pass, whose persistent GEMM CTAs fill every SM's
// register file, so gather CTAs only get SMs at kernel boundaries.  HBM tables: one short CTA per 8 rows (the whole
// gather is ~0.1 ms of SM time).  Pinned tables: the transfer lasts ~1.2 ms per 256 clips at PCIe speed whatever
// the grid; 32 looping CTAs keep ~1 MB in flight (PCIe needs ~100 KB) and hold 32 SMs for that long, which costs the
// concurrent GEMMs ~0.6 ms per 6.3 ms step (measured).  Short CTAs on a low-priority stream were measured too: they
// starve behind the back-to-back forward kernels and the transfer ends up serialised after the step (+1.0 ms).
// ------------------------------------------------------------------------------------------------
struct GatherArgs {
  const float* rows[AFFT_MAX_MODS];
  float* out[AFFT_MAX_MODS];
  int width4[AFFT_MAX_MODS];   // row width in float4
  int n_mod;
  int rows_per_mod;            // B * T
  const int32_t* idx;          // [n_mod][B*T]
};

__global__ void __launch_bounds__(256) gather_rows_kernel(const GatherArgs a) {
  const int warps_per_cta = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = static_cast<long long>(a.n_mod) * a.rows_per_mod;
  for (long long r = static_cast<long long>(blockIdx.x) * warps_per_cta + (threadIdx.x >> 5); r < total;
       r += static_cast<long long>(gridDim.x) * warps_per_cta) {
    const int m = static_cast<int>(r / a.rows_per_mod);
    const int o = static_cast<int>(r - static_cast<long long>(m) * a.rows_per_mod);
    const int src = __ldg(a.idx + r);
    const int w4 = a.width4[m];
    float4* dst = reinterpret_cast<float4*>(a.out[m]) + static_cast<long long>(o) * w4;
    if (src < 0) {
      for (int c = lane; c < w4; c += 32) dst[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    const float4* sp = reinterpret_cast<const float4*>(a.rows[m]) + static_cast<long long>(src) * w4;
    float4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (lane + 32 * i < w4) v[i] = __ldcs(sp + lane + 32 * i);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (lane + 32 * i < w4) dst[lane + 32 * i] = v[i];
    for (int c = lane + 256; c < w4; c += 32) dst[c] = __ldcs(sp + c);  // rows wider than 1024 floats
  }
}

extern "C" int afft_store_gather(afft_feature_store* s, int32_t B, int32_t T, const int32_t* row_idx_dev, void* const* out_dev,
                                 void* stream) {
  if (s == nullptr) return sfail(nullptr, AFFT_ERR_INVALID, "store_gather: null store");
  if (B < 0 || T < 1 || row_idx_dev == nullptr || out_dev == nullptr) return sfail(s, AFFT_ERR_INVALID, "store_gather: bad arguments");
  if (B == 0) return AFFT_OK;
  GatherArgs a;
  a.n_mod = static_cast<int>(s->mods.size());
  a.rows_per_mod = B * T;
  a.idx = row_idx_dev;
  for (int m = 0; m < a.n_mod; ++m) {
    if (s->mods[m].rows == nullptr) return sfail(s, AFFT_ERR_MISSING, "store_gather: modality " + std::to_string(m) + " has no row table");
    if (out_dev[m] == nullptr || (reinterpret_cast<uintptr_t>(out_dev[m]) & 15u) != 0) return sfail(s, AFFT_ERR_INVALID, "store_gather: outputs must be 16-byte aligned device tensors");
    a.rows[m] = s->mods[m].rows;
    a.out[m] = static_cast<float*>(out_dev[m]);
    a.width4[m] = s->mods[m].width / 4;
  }
  const long long total = static_cast<long long>(a.n_mod) * a.rows_per_mod;
  bool all_device = true;
  for (int m = 0; m < a.n_mod; ++m) all_device = all_device && s->mods[m].rows_on_device;
  const int ctas = static_cast<int>(all_device ? (total + 7) / 8 : std::min<long long>((total + 7) / 8, 32LL));
  gather_rows_kernel<<<ctas, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return sfail(s, AFFT_ERR_CUDA, std::string("gather launch failed: ") + cudaGetErrorString(e));
  return AFFT_OK;
}
