/*
 * lc_abi.h - C ABI of liblc_b200.so: brute-force 256-bit Hamming matching for loop closing on B200.
 *
 * This is the drop-in boundary for the one data-parallel hot path of F-Fer/SLAM-Loop-Closing.
 * The reference has no plugin/FFI mechanism: `matcher_` is a concrete cv::Ptr<cv::BFMatcher>
 * (reference include/loop_closing.hpp:73).  The seam is therefore the bodies of
 *     LoopClosingSystem::matchFeatures   (include/loop_closing.hpp:40, README.md:116-117)
 *     LoopClosingSystem::detectLoops     (include/loop_closing.hpp:48, README.md:121-126)
 * plus the `frames_.push_back` in processFrame (include/loop_closing.hpp:34,69), which must also
 * append the frame's descriptors to the device-resident store.  include/loop_closing_b200.hpp shows
 * the C++ adapter a maintainer would compile against OpenCV; INTEGRATION.md shows the bindings.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch/OpenCV types cross this boundary;
 *   - every function returns an lc_status (0 = ok) and never throws or aborts; lc_last_error()
 *     gives the message of the last failure on that context;
 *   - descriptors are 32-byte rows (cv::Mat CV_8UC1, n x 32, Frame::descriptors,
 *     include/loop_closing.hpp:16); `row_stride` is in bytes so non-continuous Mats are accepted;
 *     the library copies them before returning, the caller keeps ownership;
 *   - empty inputs are not errors: they produce empty results, as cv::BFMatcher does;
 *   - there is NO CPU fallback: if no CUDA device is usable lc_create fails.
 *   - a context is driven by one application thread at a time (the reference is single-threaded).
 */
#ifndef LC_ABI_H
#define LC_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LC_ABI_VERSION 2 /* 2: lc_db_append_many_device takes the producer stream; LC_MAX_STORE_ROWS; lc_process_frame */
#define LC_DESC_BYTES 32       /* 256-bit ORB descriptor, README.md:115 */
#define LC_MAX_ROWS_PER_KF ((1 << 20) - 1) /* rows of one frame; index shares a u32 key with the distance */
/* rows of ALL keyframes of ALL shards together: best_packed = distance << 32 | global_row keeps the global row id in
 * 32 bits (137 GB of descriptors).  An append that would cross it fails with LC_ERR_CAPACITY and changes nothing. */
#define LC_MAX_STORE_ROWS (1ll << 32)

typedef struct lc_ctx lc_ctx;

typedef enum lc_status {
  LC_OK = 0,
  LC_ERR_INVALID = 1,  /* bad argument */
  LC_ERR_CUDA = 2,     /* CUDA runtime/driver failure (message has the CUDA error string) */
  LC_ERR_NOMEM = 3,    /* host or device allocation failed */
  LC_ERR_CAPACITY = 4, /* caller-provided output array too small, or the store is full (LC_MAX_STORE_ROWS) */
  LC_ERR_IO = 5        /* lc_db_save / lc_db_load */
} lc_status;

typedef enum lc_kernel {
  LC_KERNEL_AUTO = 0, /* library default: the measured-fastest bit-exact path (LC_KERNEL_MXF4 when the problem fills
                         the GPU, else the carry-save integer kernels) */
  LC_KERNEL_POPC = 1, /* 8 x (XOR, POPC) per pair: the naive integer-popc formulation */
  LC_KERNEL_CSA = 2,  /* carry-save adder tree: 16 LOP3 + 4 POPC per pair, same result */
  LC_KERNEL_TENSOR = 3, /* tcgen05.mma kind::i8 on +1/-1 expanded bits (dot = 256 - 2*hamming) + arg-max epilogue, same
                           result; problems too small to fill the GPU use the CSA kernels */
  LC_KERNEL_MXF4 = 4    /* tcgen05.mma kind::mxf4 on E2M1 +1/-1 expanded bits (unit block scales, exact f32 sums): twice
                           the int8 rate; the kernel keeps distances only and the per-query best row is recovered on
                           the winning keyframe afterwards; same result */
} lc_kernel;

/* CUDA-event timing of host-buffer calls (lc_get_last_timing).  A pair match is a ~50 us call in which each event record
 * costs ~3 us of stream time, so by default it is not timed (its lc_timing reports path and launches, times read 0). */
typedef enum lc_timing_mode {
  LC_TIMING_AUTO = 0, /* every host-buffer call except lc_match_pair / lc_match_vs_keyframe */
  LC_TIMING_ON = 1    /* every host-buffer call */
} lc_timing_mode;

/* Parameters of the path.  Defaults (lc_default_config) follow the reference header and README;
 * the extra knobs exist because src/loop_closing.cpp is absent and README prose is ambiguous
 * (SURVEY.md Appendix B). */
typedef struct lc_config {
  uint32_t struct_size;  /* = sizeof(lc_config); set by lc_default_config */
  int32_t device;        /* CUDA device ordinal */
  double loop_threshold; /* ctor default 0.7 (include/loop_closing.hpp:31); README.md:108 passes 0.15 */
  int32_t min_loop_gap;  /* 30 (include/loop_closing.hpp:31, README.md:109) */
  int32_t min_good;      /* 50 (README.md:124) */
  int32_t filter_floor;  /* good iff d <= max(2*min_dist, filter_floor); README-literal 0 */
  int32_t filter_strict; /* 0: "<=" ; 1: "<" */
  int32_t min_init;      /* <0: min_dist is the true minimum; else it starts at this value */
  int32_t gap_strict;    /* 0: cur - id >= gap ; 1: ">" */
  int32_t thr_inclusive; /* 0: similarity > threshold ; 1: ">=" */
  int32_t shard_rank;    /* keyframe k is stored on the context with k % shard_world == shard_rank */
  int32_t shard_world;   /* 1 = unsharded */
  int32_t kernel;        /* lc_kernel */
  int32_t timing;        /* lc_timing_mode */
  int64_t slab_rows;     /* descriptor rows per device slab; 0 = default (8 Mi rows = 256 MiB) */
} lc_config;

/* One record per stored keyframe, the integer facts detectLoops needs (README.md:121-126).
 * similarity = (double)n_good / (double)min(nq, n_kf) is formed on the host from these ints so it
 * is bit-identical to the reference's double division.  min_dist = n_good = -1 when the keyframe
 * was not compared (closer than min_loop_gap, or an empty side). */
typedef struct lc_kf_result {
  int32_t frame_id;
  int32_t n_kf;
  int32_t min_dist;
  int32_t n_good;
} lc_kf_result;

/* Same members, order and layout as the reference's LoopCandidate (include/loop_closing.hpp:22-27). */
typedef struct lc_loop_candidate {
  int32_t current_frame_id;
  int32_t matched_frame_id;
  int32_t num_matches;
  double similarity_score;
} lc_loop_candidate;

/* CUDA-event timing of the last host-buffer call on the context (milliseconds). */
typedef struct lc_timing {
  float h2d_ms;    /* pinned staging -> device */
  float kernel_ms; /* all kernels of the call */
  float d2h_ms;    /* device -> pinned results */
  float total_ms;  /* first H2D to last D2H on the stream */
  int32_t kernel_launches;
  int32_t path;    /* 1 = whole-keyframe scan kernel, 2 = split kernel + finalize, 3 = int8 tensor-core kernel + finalize, 4 = FP4 tensor-core kernel + finalize + resolve, 5 = single-launch pair kernel (lc_match_pair and friends), 6 = SIFT/L2 tensor-core kernel (lc_l2_*) */
  float match_ms;  /* the dominant (matching) kernel alone, summed over its launches of the call */
  int32_t match_launches;
} lc_timing;

int32_t lc_abi_version(void);
void lc_default_config(lc_config* cfg);

/* Create / destroy a context on cfg->device (cfg == NULL: defaults). */
int32_t lc_create(const lc_config* cfg, lc_ctx** out);
void lc_destroy(lc_ctx* ctx);
/* Message of the last failed call on ctx (ctx == NULL: last lc_create failure of this thread). */
const char* lc_last_error(const lc_ctx* ctx);
int32_t lc_get_config(const lc_ctx* ctx, lc_config* out);
/* loop_threshold, min_loop_gap, min_good and the filter/gap knobs may change between calls;
 * device and sharding may not. */
int32_t lc_set_params(lc_ctx* ctx, const lc_config* cfg);

/* ---- keyframe-descriptor store: replaces frames_[i].descriptors as the matcher's train side
 *      (include/loop_closing.hpp:16,69).  Device-resident, append-only, ragged. ---- */

/* Append one keyframe (pinned staging -> device).  With sharding every rank appends every keyframe
 * in the same order; a rank that does not own keyframe k records only (frame_id, n) and may pass
 * desc == NULL.  *kf_index (optional) receives the global insertion index. */
int32_t lc_db_append(lc_ctx* ctx, int32_t frame_id, const uint8_t* desc, int32_t n, size_t row_stride,
                     int64_t* kf_index);
/* Bulk append of `count` keyframes with contiguous 32-byte rows each; desc[i] may be NULL for
 * keyframes this shard does not own (or when n_rows[i] == 0). */
int32_t lc_db_append_many(lc_ctx* ctx, int32_t count, const int32_t* frame_ids, const int32_t* n_rows,
                          const uint8_t* const* desc);
/* Same with device-resident source rows: d_rows holds the rows of the keyframes of this call that THIS shard owns, back
 * to back (device-to-device copies).  `cuda_stream` is the stream whose work produced d_rows (a cudaStream_t;
 * cudaStreamLegacy (0x1) for the legacy default stream): the copies are ordered behind everything queued on it so far,
 * and it is made to wait for them, so the caller may reuse or free d_rows in stream order right after the call.  NULL:
 * the rows are already complete (the caller has synchronised) and must stay untouched until lc_synchronize(). */
int32_t lc_db_append_many_device(lc_ctx* ctx, int32_t count, const int32_t* frame_ids, const int32_t* n_rows,
                                 const void* d_rows, void* cuda_stream);
/* Global and local (this shard) keyframe / row counts; any pointer may be NULL. */
int32_t lc_db_size(const lc_ctx* ctx, int64_t* n_keyframes, int64_t* n_rows, int64_t* n_local_keyframes,
                   int64_t* n_local_rows);
int32_t lc_db_clear(lc_ctx* ctx);
/* Copy keyframe `kf_index` (global index, must be owned) back to the host: desc_out has room for
 * cap_rows rows; *n and *frame_id receive the stored values. */
int32_t lc_db_read(lc_ctx* ctx, int64_t kf_index, uint8_t* desc_out, int32_t cap_rows, int32_t* n,
                   int32_t* frame_id);
/* Persist / restore this shard of the store (raw rows + frame ids + counts). */
int32_t lc_db_save(lc_ctx* ctx, const char* path);
int32_t lc_db_load(lc_ctx* ctx, const char* path);

/* ---- matchFeatures (include/loop_closing.hpp:40): BFMatcher(NORM_HAMMING).match(q, t) followed by
 *      README.md:117's "2x minimum distance" filter.  Outputs have nq entries (any may be NULL):
 *      train_idx[i]/dist[i] = DMatch(i, train_idx[i], 0, (float)dist[i]); good[i] = survives the
 *      filter.  *n_matches = nq, or 0 when either side is empty (then nothing else is written
 *      except *min_dist = -1, *n_good = 0). ---- */
int32_t lc_match_pair(lc_ctx* ctx, const uint8_t* q, int32_t nq, size_t q_stride, const uint8_t* t,
                      int32_t nt, size_t t_stride, int32_t* train_idx, int32_t* dist, uint8_t* good,
                      int32_t* n_matches, int32_t* min_dist, int32_t* n_good);
/* Same, the train side being stored keyframe `kf_index` (no re-upload of the resident frame). */
int32_t lc_match_vs_keyframe(lc_ctx* ctx, const uint8_t* q, int32_t nq, size_t q_stride, int64_t kf_index,
                             int32_t* train_idx, int32_t* dist, uint8_t* good, int32_t* n_matches,
                             int32_t* min_dist, int32_t* n_good);

/* One surviving match of matchFeatures: cv::DMatch(query_idx, train_idx, 0, (float)distance). */
typedef struct lc_match {
  int32_t query_idx;
  int32_t train_idx;
  int32_t distance;
} lc_match;
/* matchFeatures(current frame, frames_[k]) for a LIST of stored keyframes in one call, good matches compacted: what
 * the loop candidates hand to estimatePose / triangulatePoints (include/loop_closing.hpp:43-53, README.md:101-102)
 * without re-matching on the CPU.  matches[offsets[i] .. offsets[i + 1]) are the surviving matches against keyframe
 * kf_index[i] in ascending query order (offsets has n_sel + 1 entries); min_dist[i] (optional) is that match's
 * minimum distance, -1 when either side is empty.  All keyframes must be owned by this shard.  If matches_cap is too
 * small the call fails with LC_ERR_CAPACITY after filling offsets[] with the sizes needed. */
int32_t lc_match_keyframes(lc_ctx* ctx, const uint8_t* q, int32_t nq, size_t q_stride, int32_t n_sel,
                           const int64_t* kf_index, lc_match* matches, int64_t matches_cap, int64_t* offsets,
                           int32_t* min_dist);

/* ---- detectLoops' data-parallel part (include/loop_closing.hpp:48): the current frame against
 *      every stored keyframe of this shard.  results[] receives one record per LOCAL keyframe in
 *      insertion order (local slot s is global keyframe shard_rank + s*shard_world); *n_results
 *      their number.  best_packed (optional, nq entries) = min over compared rows of
 *      (distance << 32 | global_row), global_row counting rows of ALL keyframes in insertion order
 *      (= OpenCV's multi-image tie-break); UINT64_MAX if nothing was compared. ---- */
int32_t lc_loop_check(lc_ctx* ctx, const uint8_t* q, int32_t nq, size_t q_stride, int32_t current_frame_id,
                      lc_kf_result* results, int64_t results_cap, int64_t* n_results, uint64_t* best_packed);
/* The same in two halves, so that the host can run ORB detection of the next frame (or anything else) while the GPU
 * scans the store: _begin copies the query, queues upload + kernels + result download on the context's stream and
 * returns at once; _end waits and hands out what lc_loop_check would have returned (*n_results records).  The pending
 * check keeps its own pinned result buffer, so between the two every call except another loop check (host-buffer
 * lc_loop_check*, lc_detect_loops, lc_process_frame) and lc_db_load is allowed - the next frame's lc_match_pair, an
 * lc_db_append (the pending check does not see the new keyframe), lc_match_keyframes.  Such a call runs on the same
 * stream BEHIND the check: it returns only after the check has finished on the GPU, so make it before _begin or after
 * _end when its latency matters; it is not event-timed, and lc_get_last_timing describes the check again after _end. */
int32_t lc_loop_check_begin(lc_ctx* ctx, const uint8_t* q, int32_t nq, size_t q_stride, int32_t current_frame_id,
                            int32_t want_best);
int32_t lc_loop_check_end(lc_ctx* ctx, lc_kf_result* results, int64_t results_cap, int64_t* n_results,
                          uint64_t* best_packed);
/* Throughput mode: n_frames query frames in one call.  results is [n_frames][n_local_keyframes];
 * best_packed (optional) is the concatenation of the frames' nq[i]-entry arrays. */
int32_t lc_loop_check_batch(lc_ctx* ctx, int32_t n_frames, const uint8_t* const* q, const int32_t* nq,
                            const int32_t* current_frame_ids, lc_kf_result* results, int64_t results_cap,
                            uint64_t* best_packed);
/* Host-side decision rule of README.md:121-126 over records (from one shard or gathered from all):
 * candidate iff similarity > loop_threshold and n_good >= min_good; ascending input order. */
int32_t lc_select_candidates(const lc_ctx* ctx, const lc_kf_result* results, int64_t n_results, int32_t nq,
                             int32_t current_frame_id, lc_loop_candidate* out, int64_t out_cap, int64_t* n_out);
/* detectLoops in one call (unsharded contexts): lc_loop_check + lc_select_candidates. */
int32_t lc_detect_loops(lc_ctx* ctx, const uint8_t* q, int32_t nq, size_t q_stride, int32_t current_frame_id,
                        lc_loop_candidate* out, int64_t out_cap, int64_t* n_out);

/* ---- processFrame's matcher side in ONE call (include/loop_closing.hpp:34, workflow README.md:94-102): for the frame
 *      that ORB has just described,
 *        matchFeatures(previous frame, this frame)   (query = the frame appended last, train = this one),
 *        frames_.push_back(this frame)               (= lc_db_append), and
 *        detectLoops(frame_id)                       (= lc_loop_check + lc_select_candidates)
 *      behind ONE upload of the frame's rows - they land in the store, and both matches read them there - with the
 *      table entries sent as kernel arguments, every launch queued back to back and one wait at the end.  Same results
 *      as the three separate calls.  Unsharded contexts only.  train_idx / dist / good have `cap` entries, at least the
 *      row count of the PREVIOUS frame (the query side); *n_matches = that count, or 0 for the first frame or an empty
 *      side (then *min_dist = -1, *n_good = 0).  out[] receives the loop candidates in ascending keyframe order.
 *      If the call fails after the append, *kf_index >= 0 says the frame is in the store. ---- */
int32_t lc_process_frame(lc_ctx* ctx, int32_t frame_id, const uint8_t* desc, int32_t n, size_t row_stride,
                         int32_t* train_idx, int32_t* dist, uint8_t* good, int32_t cap, int32_t* n_matches,
                         int32_t* min_dist, int32_t* n_good, lc_loop_candidate* out, int64_t out_cap, int64_t* n_out,
                         int64_t* kf_index);
/* The same in three steps, so that the host's estimatePose / triangulatePoints (include/loop_closing.hpp:43-53) run on
 * the pair matches WHILE the GPU is still scanning the store: _begin queues everything and returns; _matches (optional)
 * waits for the pair match alone; _end waits for the loop check, hands out the candidates and - if results != NULL -
 * the per-keyframe records (*n_results of them, one per stored keyframe including this frame).  Between _begin and
 * _end the rules of lc_loop_check_begin apply. */
int32_t lc_process_frame_begin(lc_ctx* ctx, int32_t frame_id, const uint8_t* desc, int32_t n, size_t row_stride,
                               int64_t* kf_index);
int32_t lc_process_frame_matches(lc_ctx* ctx, int32_t* train_idx, int32_t* dist, uint8_t* good, int32_t cap,
                                 int32_t* n_matches, int32_t* min_dist, int32_t* n_good);
int32_t lc_process_frame_end(lc_ctx* ctx, lc_loop_candidate* out, int64_t out_cap, int64_t* n_out,
                             lc_kf_result* results, int64_t results_cap, int64_t* n_results);

/* Merge step of a keyframe-sharded check (one process per GPU, store created with shard_rank / shard_world): every rank
 * lets lc_loop_check_batch_device write its records [n_frames][slots] and bests [rows] into ONE contribution buffer
 * [records | bests] of contrib_bytes (a multiple of 16), all-gathers the contributions (ncclAllGather) and calls this:
 * records_out[f][slot * world + r] = rank r's record (keyframe k lives on rank k % world, slot k / world) and
 * best_out[i] = the unsigned minimum over the ranks - one kernel, asynchronous on `cuda_stream`. */
int32_t lc_merge_shards_device(lc_ctx* ctx, const void* d_gathered, int64_t contrib_bytes, int32_t world, int32_t n_frames,
                               int64_t slots, int64_t rows, lc_kf_result* d_records_out, uint64_t* d_best_out,
                               void* cuda_stream);
/* The same merge WITHOUT a collective, over peer memory (NVLink / NVSwitch): every rank owns an exchange buffer
 * [flags | slot 0 | slot 1] that its peers map; a check of sequence number `seq` (1, 2, 3, ... - the same on every rank)
 * lets lc_loop_check_batch_device write its [records | bests] contribution into slot seq & 1 of its OWN buffer
 * (lc_exchange_slot), and ONE kernel signals the peers, waits for their signals and reads their slots directly while it
 * merges (lc_kernels.cuh: merge_peers_kernel).  Setup, once: lc_exchange_create (slot_bytes >= the largest contribution;
 * returns a 64-byte CUDA IPC handle and/or the buffer's device address), exchange the handles by any means (MPI,
 * torch.distributed, a socket), lc_exchange_connect with all `world` handles in rank order - or, for ranks of ONE process
 * with peer access enabled, with their buffers' device addresses.  phase: 3 = signal + wait + merge (the normal call);
 * 1 = signal only, 2 = wait + merge only (for drivers that emulate several ranks on one GPU, where kernels must not wait
 * for each other).  A peer that never signals makes the kernel trap after 5 s (an error at the next synchronisation). */
#define LC_IPC_HANDLE_BYTES 64
int32_t lc_exchange_create(lc_ctx* ctx, int64_t slot_bytes, void* ipc_handle_out, void** d_base_out);
int32_t lc_exchange_connect(lc_ctx* ctx, int32_t world, int32_t rank, const void* ipc_handles, void* const* d_bases);
int32_t lc_exchange_slot(lc_ctx* ctx, int64_t seq, void** d_slot, int64_t* slot_bytes);
int32_t lc_exchange_close(lc_ctx* ctx);
int32_t lc_merge_shards_peer_device(lc_ctx* ctx, int64_t seq, int32_t n_frames, int64_t slots, int64_t rows,
                                    lc_kf_result* d_records_out, uint64_t* d_best_out, int32_t phase, void* cuda_stream);
/* ---- device-pointer variants: everything already in HBM, asynchronous on `cuda_stream`
 *      (a cudaStream_t; NULL = the context's own non-blocking stream - pass cudaStreamLegacy (0x1) to
 *      mean the legacy default stream).  d_q: contiguous 32-byte rows, 16-byte
 *      aligned.  Used for multi-GPU merging (all-gather of [d_results | d_best_packed], lc_merge_shards_device)
 *      and for kernel-only timing.  d_best_packed may be NULL.
 *      A context's calls share its scratch buffers, so the library serialises them across streams itself: a call
 *      on `cuda_stream` waits (on the device, by events) for everything the context queued before on any stream,
 *      and later calls wait for it.  A stream handed in must stay alive until the call returns; the caller still
 *      synchronises with ITS stream before it reads d_results. ---- */
int32_t lc_loop_check_device(lc_ctx* ctx, const void* d_q, int32_t nq, int32_t current_frame_id,
                             lc_kf_result* d_results, uint64_t* d_best_packed, void* cuda_stream);
int32_t lc_loop_check_batch_device(lc_ctx* ctx, int32_t n_frames, const void* d_q, const int32_t* nq,
                                   const int32_t* current_frame_ids, lc_kf_result* d_results,
                                   uint64_t* d_best_packed, void* cuda_stream);
/* d_summary[0] = min_dist, d_summary[1] = n_good (both -1/0 when a side is empty). */
int32_t lc_match_pair_device(lc_ctx* ctx, const void* d_q, int32_t nq, const void* d_t, int32_t nt,
                             int32_t* d_train_idx, int32_t* d_dist, uint8_t* d_good, int32_t* d_summary,
                             void* cuda_stream);

/* ---- the SHIPPED pipeline's matcher (src/main.cpp:509-534): SIFT descriptors, cv::BFMatcher(NORM_L2).knnMatch(k = 2)
 *      and Lowe's ratio test  best.distance < ratio * second.distance;  its loop search (src/main.cpp:1375-1392) matches
 *      the current view against every view at least loop_gap back, skipping pairs where a side has < min_rows rows.
 *      Descriptors are OpenCV SIFT rows of 128 values: desc_type LC_DESC_F32 (CV_32F, the default of cv::SIFT; the
 *      values must be integers in [0, 255], which is what cv::SIFT produces) or LC_DESC_U8 (CV_8U).  Squared distances
 *      are exact integers, distances are sqrtf of them: bit-identical to OpenCV's.  The views live in their own
 *      device-resident store (sharded like the ORB store). ---- */
#define LC_L2_DIMS 128
typedef enum lc_desc_type { LC_DESC_F32 = 0, LC_DESC_U8 = 1 } lc_desc_type;
typedef struct lc_l2_result { /* one per stored view */
  int32_t frame_id;
  int32_t n_kf;
  int32_t n_good; /* matches surviving the ratio test (matches.size() of main.cpp:1389); -1 when the pair was skipped */
  int32_t reserved;
} lc_l2_result;
int32_t lc_l2_db_append(lc_ctx* ctx, int32_t frame_id, const void* desc, int32_t n, size_t row_stride, int32_t desc_type,
                        int64_t* kf_index);
int32_t lc_l2_db_size(const lc_ctx* ctx, int64_t* n_keyframes, int64_t* n_local_keyframes);
int32_t lc_l2_db_clear(lc_ctx* ctx);
/* matchFeatures(desc1 = q, desc2 = t, ratio): per query the best train row, its distance, the second distance (-1 when
 * the train side has fewer than two rows) and whether it survives the ratio test; *n_good = goodMatches.size(). */
int32_t lc_l2_match_pair(lc_ctx* ctx, const void* q, int32_t nq, size_t q_stride, const void* t, int32_t nt,
                         size_t t_stride, int32_t desc_type, double ratio, int32_t* train_idx, float* distance,
                         float* second_distance, uint8_t* good, int32_t* n_good);
/* The loop search's matching step: results[] receives one record per LOCAL view in insertion order. */
int32_t lc_l2_loop_check(lc_ctx* ctx, const void* q, int32_t nq, size_t q_stride, int32_t desc_type,
                         int32_t current_frame_id, int32_t loop_gap, double ratio, int32_t min_rows,
                         lc_l2_result* results, int64_t results_cap, int64_t* n_results);

/* The whole matching step of the shipped loop search (src/main.cpp:1375-1389: for curr, for past <= curr - loopGap:
 * matchFeatures(desc[curr], desc[past], 0.7)) in one call, the query side being the RESIDENT store itself: stored views
 * first_view .. first_view + n_query_views - 1, each against every stored view; results is [n_query_views][n_views]
 * (*n_views = number of stored views), n_good = -1 where main.cpp skips the pair.  Nothing is uploaded.  Unsharded contexts. */
int32_t lc_l2_loop_search(lc_ctx* ctx, int64_t first_view, int32_t n_query_views, int32_t loop_gap, double ratio,
                          int32_t min_rows, lc_l2_result* results, int64_t results_cap, int64_t* n_views);

/* ---- measurement aid: the peak of the pipe the default loop-check kernel runs on.  For about `milliseconds` every SM
 *      issues nothing but tcgen05.mma kind::mxf4 (M 128 x N 256 x K 64, operands in shared memory, the match kernel's
 *      value distribution); *mac_per_s = multiply-accumulates per second of the whole GPU (a descriptor pair is 256 of
 *      them), *sm_mhz = the SM clock the pipe really held (cycle counter / elapsed time).  bench.py's roofline divides
 *      by this figure, measured in the same run. ---- */
int32_t lc_probe_mxf4_peak(lc_ctx* ctx, int32_t milliseconds, double* mac_per_s, double* sm_mhz);

/* ---- introspection ---- */
int32_t lc_get_last_timing(const lc_ctx* ctx, lc_timing* out);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
int64_t lc_launch_count(const lc_ctx* ctx);
/* Block until everything queued on the context's own stream has finished. */
int32_t lc_synchronize(lc_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* LC_ABI_H */
