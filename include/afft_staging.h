/*
 * afft_staging - C ABI of the INPUT side of the AFFT hot path: from "clip = (video,
 * start, end)" to the (B, T, C_m) fp32 feature tensors afft_forward consumes.
 *
 * Reference being replaced (pure Python, run by 32 DataLoader worker processes per GPU):
 *   datasets/base_video_dataset.py:225-337  BaseVideoDataset._sample        window, every-k-th-frame subsampling, padding
 *   datasets/reader_fns.py:108-157          EpicRULSTMFeatsReader._read_rulstm_features / _convert_to_orig_video_fps
 *   datasets/reader_fns.py:65-106           read_representations: LMDB get per frame, closest-earlier-frame fallback
 *                                           (radius 9), zero rows for missing frames
 *   test.py:81                              feature_dict = {mod: tensor.to(device)}   (the H2D copy)
 *
 * B200 design: the feature rows of a modality live in ONE row table [n_rows, C_m] fp32 that the GPU can address -
 * pinned host memory (read by the gather kernel over PCIe, zero-copy) or HBM when it fits (180 GB).  The host only
 * computes the PLAN: T row numbers per (modality, clip) - the reference's window / subsampling / fallback arithmetic
 * on integers, no feature bytes touched by the CPU.  afft_store_gather then builds the batch on the device: one warp
 * per (clip, step) row, 16-byte loads from the table, zero rows for row number -1.
 *
 * Conventions as in afft_b200.h: int status returns, caller-owned memory, explicit stream, plain C types.
 */
#ifndef AFFT_STAGING_H_
#define AFFT_STAGING_H_

#include <stddef.h>
#include <stdint.h>

#include "afft_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct afft_feature_store afft_feature_store;

/* base_video_dataset.py:28-31 (SAMPLE_STRAT_*).  'random_clip' draws from process-global RNGs and is not offered. */
enum { AFFT_SAMPLE_LAST = 0, AFFT_SAMPLE_CENTER = 1, AFFT_SAMPLE_FIRST = 2, AFFT_SAMPLE_RANDOM = 3 };

/* n_mod modalities of row width widths[m]; orig_fps_index[m] != 0: the modality's frames are numbered in the ORIGINAL
 * video's frame rate (audio / poses LMDBs, reader_fns.py:131-133), 50 fps for EK100 names (3-digit suffix) and
 * 59.94 for EK55 (2-digit), reader_fns.py:148-157. */
AFFT_API int afft_store_create(int32_t n_mod, const int32_t* widths, const int32_t* orig_fps_index, afft_feature_store** out);
AFFT_API void afft_store_destroy(afft_feature_store* s);
AFFT_API const char* afft_store_error(const afft_feature_store* s);

/* Index of one video in one modality: frame_ids[i] (strictly ascending, the numbers in the LMDB keys
 * "{video}_frame_{id:010d}.jpg", reader_fns.py:133) is stored in row first_row + i of the modality's row table.
 * The arrays are copied. */
AFFT_API int afft_store_add_video(afft_feature_store* s, int32_t mod, const char* video_name, const int32_t* frame_ids,
                                  int64_t n, int64_t first_row);

/* The row table of a modality: [n_rows, widths[mod]] fp32, contiguous, caller-owned and alive until destroy.
 * Either pinned host memory (addressable from the device under UVA) or device memory; the library only reads it,
 * and only from afft_store_gather's kernel. */
AFFT_API int afft_store_set_rows(afft_feature_store* s, int32_t mod, const void* rows, int64_t n_rows);

/* A clip with no stored frame anywhere in its window makes the reference reader assert
 * (`assert len(features_not_none) > 0`, datasets/reader_fns.py:97); afft_store_plan refuses such a clip with
 * AFFT_ERR_INVALID.  allow != 0 turns the check off (the clip is then planned as T zero rows). */
AFFT_API int afft_store_allow_empty_clips(afft_feature_store* s, int32_t allow);

/* The plan of a batch: for every modality m, clip b and step t the row number (or -1 = zero row) into
 * row_idx[(m * B + b) * T + t], and - when frame_ids_out is not NULL - the frame id the reference would have asked
 * the LMDB for into the same position.  Pure host arithmetic, bit-identical to the reference's float64/int
 * arithmetic.  Returns AFFT_ERR_INVALID for an unknown video, a window without a frame id >= 1
 * (reader_fns.py:122), or T < 1. */
AFFT_API int afft_store_plan(afft_feature_store* s, int32_t B, const char* const* video_names, const double* start_sec,
                             const double* end_sec, double fps, int32_t T, double frame_rate /* <= 0: the video's fps */,
                             int32_t strategy, int32_t* row_idx, int32_t* frame_ids_out);

/* sample_strategy = random_clip (datasets/base_video_dataset.py:245-248,282-287): the reference calls _sample once per
 * (clip, modality) and draws two numbers per call - rng.integers(max(nframes - frames_to_ext, 0)) from its numpy Generator
 * (only when that bound is positive) and random.random() from Python's global generator, turned into
 * offset = round(random * max(round(fps / frame_rate / 3), 1)).  The caller makes those draws (afft_b200.staging does it in
 * the reference's order) and passes them as rand_start_frame / rand_offset [n_mod, B]; everything else is afft_store_plan. */
AFFT_API int afft_store_plan_random(afft_feature_store* s, int32_t B, const char* const* video_names, const double* start_sec,
                                    const double* end_sec, double fps, int32_t T, double frame_rate,
                                    const int64_t* rand_start_frame, const int32_t* rand_offset, int32_t* row_idx,
                                    int32_t* frame_ids_out);

/* Device gather: out[m] (B, T, widths[m]) fp32 device tensors, row_idx_dev the plan on the device (same layout).
 * One kernel launch per call (all modalities), enqueued on `stream`. */
AFFT_API int afft_store_gather(afft_feature_store* s, int32_t B, int32_t T, const int32_t* row_idx_dev, void* const* out_dev,
                               void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AFFT_STAGING_H_ */
