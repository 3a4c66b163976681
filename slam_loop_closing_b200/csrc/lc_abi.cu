// lc_abi.cu - host side of liblc_b200.so: the device-resident keyframe-descriptor store, kernel
// dispatch and the extern "C" boundary declared in include/lc_abi.h.
//
// Replaces the bodies of LoopClosingSystem::matchFeatures / detectLoops and the descriptor side of
// frames_ (reference include/loop_closing.hpp:40,48,69).  No CPU fallback: every compute entry
// point launches the sm_100a kernels of lc_kernels.cuh or fails with LC_ERR_CUDA.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only; the ranges cost a null-pointer check when no profiler is attached

#include <algorithm>
#include <cstdarg>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <vector>

#include "lc_kernels.cuh"
#include "lc_tensor.cuh"
#include "lc_mxf4.cuh"
#include "lc_l2.cuh"

using namespace lc;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

constexpr size_t kPinHalfBytes = 8u << 20;  // two pinned upload halves of 8 MiB
constexpr int64_t kDefaultSlabRows = 8ll << 20;
constexpr size_t kScratchBudgetBytes = 256u << 20;
constexpr int kFusedMaxNq = kMaxThreads * 8;
constexpr size_t kPairDirectOutBytes = 256u << 10;  // pair results up to this size are stored straight into pinned host memory
constexpr long long kTensorMinPairs = 30ll << 20;  // loop checks of fewer pairs AND fewer (query block, keyframe) units than SMs stay on the integer kernels
constexpr double kTimedMinPairs = 1e9;  // LC_TIMING_AUTO: loop checks of fewer pairs are not event-timed
constexpr int kPairMinSegRows = 8;  // fewest train rows a pair-kernel CTA scans

struct Shape {  // launch shape for nq queries
  int rq, threads, chunks;
};

}  // namespace

struct lc_ctx {
  lc_config cfg;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t pin_ev[2] = {nullptr, nullptr};
  // cross-stream ordering of the *_device entry points (StreamScope below): everything a context does shares its scratch
  // buffers, so work on a caller's stream is chained behind the context's own stream and the other way round
  cudaEvent_t own_ev = nullptr, foreign_ev = nullptr;
  bool foreign_pending = false;
  bool pair_state_dirty = false;  // a CUDA failure may have left the pair kernel's keys / ticket state off their idle values
  std::string err;

  // ---- store
  std::vector<void*> slabs;
  int64_t slab_rows = kDefaultSlabRows;
  int64_t last_slab_cap = 0, last_slab_used = 0;
  std::vector<KfEntry> h_kf;  // local keyframes
  std::vector<int64_t> h_kf_global;
  DevBuf d_kf;
  size_t d_kf_synced = 0;
  int64_t g_kf = 0, g_rows = 0, l_rows = 0;
  int32_t max_n_local = 0;
  int64_t n_partial_kf = 0;  // local keyframes whose last FP4 tile is not full width (lc_mxf4.cuh: full_tiles_only)

  // ---- staging / work buffers
  uint8_t* pin_up = nullptr;  // 2 x kPinHalfBytes
  int pin_half = 0;
  uint8_t* pin_down = nullptr;
  uint8_t* pin_loop = nullptr;      // loop-check results (may be pending while other calls use pin_down)
  uint8_t* pin_loop_dev = nullptr;
  size_t pin_loop_cap = 0;
  uint8_t* pin_down_dev = nullptr;  // the same buffer as the device addresses it (results stored by kernels directly)
  size_t pin_down_cap = 0;
  DevBuf d_q, d_t, d_results, d_scratch, d_frames, d_pairkeys, d_paircnt, d_out, d_counter, d_qimg, d_ablocks;

  lc_timing timing{};
  int64_t launches = 0;
  // host-buffer loop check split in two (lc_loop_check_begin / _end): what the queued half left for the other
  struct Pending {
    bool active = false, queued = false;
    size_t res_bytes = 0, best_bytes = 0;
    int64_t n_res = 0, rows = 0;
    // another call (a pair match, an append) ran between _begin and _end: it ran untimed and the check's timing state
    // waits here for lc_loop_check_end
    bool stashed = false, stashed_on = false;
    lc_timing stashed_timing{};
  } pending;
  // lc_process_frame_begin / _matches / _end: the frame being processed (its loop check is `pending`)
  struct FramePending {
    bool active = false, pair_done = false, pair_direct = true;
    int64_t kf_index = -1;
    int32_t frame_id = 0, nq = 0, n_prev = 0;
    bool matched = false;  // a pair kernel ran (both frames non-empty)
  } frame;
  uint8_t* pin_frame = nullptr;      // matchFeatures(prev, cur) results of the frame in flight
  uint8_t* pin_frame_dev = nullptr;
  size_t pin_frame_cap = 0;
  cudaEvent_t pair_ev = nullptr;     // after the pair kernel of lc_process_frame_begin
  // peer-memory exchange of a keyframe-sharded check (lc_exchange_* / lc_merge_shards_peer_device)
  struct Exchange {
    uint8_t* base = nullptr;  // this rank's buffer: header (flags) + two slots
    size_t slot_bytes = 0;
    int world = 0, rank = 0;
    bool connected = false;
    uint8_t* peer[kMaxPeers] = {};
    bool ipc[kMaxPeers] = {};  // opened with cudaIpcOpenMemHandle (to be closed)
  } xchg;
  bool smem_attr_set = false;  // the tensor kernels' dynamic shared-memory limit has been raised on this device
  // ---- the SIFT / L2 store of the shipped pipeline's matcher (lc_l2.cuh): 128-byte rows + their squared norms
  struct L2Store {
    std::vector<void*> slabs;
    size_t slab_cap = 0, slab_used = 0;
    std::vector<l2::L2Kf> h_kf;  // local views
    DevBuf d_kf;
    size_t synced = 0;
    int64_t g_kf = 0;
  } l2s;
  DevBuf d_l2q, d_l2t, d_l2norm, d_l2res, d_l2frames, d_l2out, d_l2img, d_l2tables;
  bool l2_attr_set = false;
  // what d_ablocks holds (query-block + CTA tables of the tensor paths), to skip identical uploads
  const void* tables_dev = nullptr;
  std::vector<uint8_t> tables_host;
  // last CTA plan of the tensor paths (run_match): valid for the same tile counts per query block and batch size
  long long plan_kc = -1;
  std::vector<int> plan_tiles, plan_ns;
  // event pairs around the launches of the matching kernel of the current host-buffer call (lc_timing.match_ms)
  std::vector<cudaEvent_t> match_ev;
  size_t match_ev_used = 0;
  bool timing_on = false;
};

namespace {

int fail(lc_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf; else g_create_error = buf;
  if (c && code == LC_ERR_CUDA) c->pair_state_dirty = true;  // pair_device_impl re-initialises keys[] and the ticket state
  return code;
}

#define LC_CUDA(c, expr)                                                                        \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess)                                                                      \
      return fail((c), e_ == cudaErrorMemoryAllocation ? LC_ERR_NOMEM : LC_ERR_CUDA, "%s: %s", #expr, \
                  cudaGetErrorString(e_));                                                      \
  } while (0)

#define LC_TRY(expr)          \
  do {                        \
    int rc_ = (expr);         \
    if (rc_ != LC_OK) return rc_; \
  } while (0)

// Every entry point that takes a `cuda_stream` runs inside one of these.  The context's scratch (pair keys, distance
// scratch, query images, CTA tables, frame table) is shared by all its calls, so a call on a caller's stream first waits
// for what the context's own stream has queued, and whoever comes next - on any stream - waits for that call's end.  The
// events are recorded while the caller's stream is certainly alive (inside its own call).
struct StreamScope {
  lc_ctx* c;
  cudaStream_t st;
  StreamScope(lc_ctx* ctx, void* cuda_stream) : c(ctx), st(cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream) {}
  int enter() {
    if (st != c->stream) {
      LC_CUDA(c, cudaEventRecord(c->own_ev, c->stream));
      LC_CUDA(c, cudaStreamWaitEvent(st, c->own_ev, 0));
    }
    if (c->foreign_pending) {
      LC_CUDA(c, cudaStreamWaitEvent(st, c->foreign_ev, 0));
      c->foreign_pending = false;
    }
    return LC_OK;
  }
  ~StreamScope() {
    if (st != c->stream && cudaEventRecord(c->foreign_ev, st) == cudaSuccess) c->foreign_pending = true;
  }
};
// host-buffer calls run on the context's own stream: wait for the last *_device call on a caller's stream, if any
int enter_host(lc_ctx* c) {
  LC_CUDA(c, cudaSetDevice(c->cfg.device));
  if (c->foreign_pending) {
    LC_CUDA(c, cudaStreamWaitEvent(c->stream, c->foreign_ev, 0));
    c->foreign_pending = false;
  }
  return LC_OK;
}

int ensure(lc_ctx* c, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return LC_OK;
  size_t ncap = std::max(bytes, b.cap * 2);
  ncap = (ncap + 255) & ~size_t(255);
  // the old buffer may still be read by queued work on the context's stream
  if (b.p) {
    LC_CUDA(c, cudaStreamSynchronize(c->stream));
    LC_CUDA(c, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
  }
  LC_CUDA(c, cudaMalloc(&b.p, ncap));
  b.cap = ncap;
  return LC_OK;
}

int ensure_pinned(lc_ctx* c, uint8_t*& host, uint8_t*& dev, size_t& cap, size_t bytes) {
  if (bytes <= cap) return LC_OK;
  if (host) {
    LC_CUDA(c, cudaStreamSynchronize(c->stream));
    LC_CUDA(c, cudaFreeHost(host));
    host = dev = nullptr;
    cap = 0;
  }
  size_t ncap = std::max(bytes, cap * 2);
  LC_CUDA(c, cudaHostAlloc((void**)&host, ncap, cudaHostAllocMapped));
  LC_CUDA(c, cudaHostGetDevicePointer((void**)&dev, host, 0));
  cap = ncap;
  return LC_OK;
}
// results of the synchronous host-buffer calls
int ensure_pin_down(lc_ctx* c, size_t bytes) { return ensure_pinned(c, c->pin_down, c->pin_down_dev, c->pin_down_cap, bytes); }
// results of a loop check: a buffer of its own, because between lc_loop_check_begin and lc_loop_check_end the caller may
// make other calls (the next frame's pair match) that land their results in pin_down
int ensure_pin_loop(lc_ctx* c, size_t bytes) { return ensure_pinned(c, c->pin_loop, c->pin_loop_dev, c->pin_loop_cap, bytes); }

// bracket the launches of the matching kernel with events while a host-buffer call is being timed
struct MatchTimer {
  lc_ctx* c;
  cudaStream_t st;
  bool on;
  MatchTimer(lc_ctx* ctx, cudaStream_t s) : c(ctx), st(s), on(ctx->timing_on && s == ctx->stream) {
    if (!on) return;
    if (c->match_ev_used + 2 > c->match_ev.size()) {
      cudaEvent_t a = nullptr, b = nullptr;
      if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { on = false; return; }
      c->match_ev.push_back(a);
      c->match_ev.push_back(b);
    }
    cudaEventRecord(c->match_ev[c->match_ev_used], st);
  }
  ~MatchTimer() {
    if (!on) return;
    cudaEventRecord(c->match_ev[c->match_ev_used + 1], st);
    c->match_ev_used += 2;
  }
};

Knobs knobs_of(const lc_config& g) {
  return Knobs{g.min_loop_gap, g.gap_strict, g.filter_floor, g.filter_strict, g.min_init, 1u << kKeyShift};
}

int mode_of(const lc_config& g) { return g.kernel == LC_KERNEL_POPC ? MODE_POPC : MODE_CSA; }  // the tensor paths fall back to CSA for tiny problems

int validate(lc_ctx* c, const lc_config& g) {
  if (g.shard_world < 1 || g.shard_rank < 0 || g.shard_rank >= g.shard_world)
    return fail(c, LC_ERR_INVALID, "shard_rank %d / shard_world %d invalid", g.shard_rank, g.shard_world);
  if (g.kernel < LC_KERNEL_AUTO || g.kernel > LC_KERNEL_MXF4) return fail(c, LC_ERR_INVALID, "unknown kernel %d", g.kernel);
  if (g.slab_rows < 0) return fail(c, LC_ERR_INVALID, "slab_rows < 0");
  if (g.timing != LC_TIMING_AUTO && g.timing != LC_TIMING_ON) return fail(c, LC_ERR_INVALID, "timing = %d is not an lc_timing_mode", g.timing);
  if (g.min_loop_gap < -(1 << 30) || g.min_loop_gap > (1 << 30)) return fail(c, LC_ERR_INVALID, "min_loop_gap out of range");
  return LC_OK;
}

// ---------------------------------------------------------------- store
// rows for a new keyframe: contiguous, inside one slab
int slab_alloc(lc_ctx* c, int64_t n, uint4** out) {
  if (n == 0) {
    *out = nullptr;
    return LC_OK;
  }
  if (c->slabs.empty() || c->last_slab_used + n > c->last_slab_cap) {
    const int64_t cap = std::max<int64_t>(c->slab_rows, n);
    void* p = nullptr;
    LC_CUDA(c, cudaMalloc(&p, (size_t)cap * LC_DESC_BYTES));
    c->slabs.push_back(p);
    c->last_slab_cap = cap;
    c->last_slab_used = 0;
  }
  *out = (uint4*)((uint8_t*)c->slabs.back() + (size_t)c->last_slab_used * LC_DESC_BYTES);
  c->last_slab_used += n;
  return LC_OK;
}

// host rows (any stride) -> device `dst` through the two pinned halves: the CPU packs one half while
// the copy engine drains the other (every user of a half waits for that half's event first)
int upload_rows(lc_ctx* c, uint8_t* dst, const uint8_t* src, int64_t n, size_t stride) {
  const int64_t half_rows = (int64_t)(kPinHalfBytes / LC_DESC_BYTES);
  for (int64_t done = 0; done < n;) {
    uint8_t* half = c->pin_up + (size_t)c->pin_half * kPinHalfBytes;
    const int64_t take = std::min(half_rows, n - done);
    LC_CUDA(c, cudaEventSynchronize(c->pin_ev[c->pin_half]));  // the last copy out of this half has drained
    if (stride == LC_DESC_BYTES) {
      memcpy(half, src + (size_t)done * LC_DESC_BYTES, (size_t)take * LC_DESC_BYTES);
    } else {
      for (int64_t r = 0; r < take; ++r)
        memcpy(half + (size_t)r * LC_DESC_BYTES, src + (size_t)(done + r) * stride, LC_DESC_BYTES);
    }
    LC_CUDA(c, cudaMemcpyAsync(dst + (size_t)done * LC_DESC_BYTES, half, (size_t)take * LC_DESC_BYTES,
                               cudaMemcpyHostToDevice, c->stream));
    LC_CUDA(c, cudaEventRecord(c->pin_ev[c->pin_half], c->stream));
    c->pin_half ^= 1;
    done += take;
  }
  return LC_OK;
}

int store_append(lc_ctx* c, int32_t frame_id, const uint8_t* desc, int32_t n, size_t stride, bool device_src,
                 int64_t* kf_index) {
  if (n < 0 || n > LC_MAX_ROWS_PER_KF) return fail(c, LC_ERR_INVALID, "n = %d outside [0, %d]", n, LC_MAX_ROWS_PER_KF);
  // best_packed = distance << 32 | global_row: the store (ALL shards together) ends at LC_MAX_STORE_ROWS rows
  if (c->g_rows + n > LC_MAX_STORE_ROWS)
    return fail(c, LC_ERR_CAPACITY, "the store holds %lld rows and cannot take %d more: global row ids are 32 bits (LC_MAX_STORE_ROWS)",
                (long long)c->g_rows, n);
  const int64_t k = c->g_kf;
  const bool own = (k % c->cfg.shard_world) == c->cfg.shard_rank;
  if (own) {
    if (n > 0 && !desc) return fail(c, LC_ERR_INVALID, "keyframe %lld is owned by this shard but desc is NULL", (long long)k);
    if (!device_src && n > 0 && stride < LC_DESC_BYTES) return fail(c, LC_ERR_INVALID, "row_stride %zu < 32", stride);
    uint4* rows = nullptr;
    LC_TRY(slab_alloc(c, n, &rows));
    if (n > 0) {
      if (device_src)
        LC_CUDA(c, cudaMemcpyAsync(rows, desc, (size_t)n * LC_DESC_BYTES, cudaMemcpyDeviceToDevice, c->stream));
      else
        LC_TRY(upload_rows(c, (uint8_t*)rows, desc, n, stride));
    }
    c->h_kf.push_back(KfEntry{rows, n, frame_id, c->g_rows});
    c->h_kf_global.push_back(k);
    c->l_rows += n;
    c->max_n_local = std::max(c->max_n_local, n);
    c->n_partial_kf += (n > 0 && !fx::full_tiles_only(n)) ? 1 : 0;
  }
  if (kf_index) *kf_index = k;
  c->g_kf += 1;
  c->g_rows += n;
  return LC_OK;
}

int sync_kf_table(lc_ctx* c) {
  const size_t n = c->h_kf.size();
  if (n == c->d_kf_synced) return LC_OK;
  if (n * sizeof(KfEntry) > c->d_kf.cap) {
    LC_TRY(ensure(c, c->d_kf, std::max<size_t>(n * 2, 1024) * sizeof(KfEntry)));
    c->d_kf_synced = 0;
  }
  LC_CUDA(c, cudaMemcpyAsync((KfEntry*)c->d_kf.p + c->d_kf_synced, c->h_kf.data() + c->d_kf_synced,
                             (n - c->d_kf_synced) * sizeof(KfEntry), cudaMemcpyHostToDevice, c->stream));
  // h_kf may reallocate on the next append: the copy above must have left the host vector
  LC_CUDA(c, cudaStreamSynchronize(c->stream));
  c->d_kf_synced = n;
  return LC_OK;
}

// ---------------------------------------------------------------- dispatch
Shape scan_shape(int nq) {  // whole-keyframe kernel: one chunk must hold all queries
  Shape s;
  s.rq = nq > 1024 ? 8 : (nq > 256 ? 4 : 2);
  s.threads = std::max(32, ((nq + 32 * s.rq - 1) / (32 * s.rq)) * 32);
  s.chunks = 1;
  return s;
}

template <int MODE>
cudaError_t launch_scan_m(int rq, dim3 grid, int threads, cudaStream_t st, const KfEntry* kfs, int n_kf,
                          const QFrame* frames, int n_frames, lc_kf_result* res, int* counter, Knobs kb) {
  switch (rq) {
    case 2: kf_scan_kernel<2, MODE><<<grid, threads, 0, st>>>(kfs, n_kf, frames, n_frames, res, counter, kb); break;
    case 4: kf_scan_kernel<4, MODE><<<grid, threads, 0, st>>>(kfs, n_kf, frames, n_frames, res, counter, kb); break;
    default: kf_scan_kernel<8, MODE><<<grid, threads, 0, st>>>(kfs, n_kf, frames, n_frames, res, counter, kb); break;
  }
  return cudaGetLastError();
}

template <int MODE>
int scan_occupancy(int rq, int threads) {
  int occ = 0;
  switch (rq) {
    case 2: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kf_scan_kernel<2, MODE>, threads, 0); break;
    case 4: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kf_scan_kernel<4, MODE>, threads, 0); break;
    default: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kf_scan_kernel<8, MODE>, threads, 0); break;
  }
  return std::max(occ, 1);
}

template <int MODE>
cudaError_t launch_split_m(int rq, dim3 grid, int threads, cudaStream_t st, const KfEntry* kfs, int k0, int kc,
                           const QFrame* frames, uint32_t* scratch, int stride, int seg_rows, Knobs kb) {
  switch (rq) {
    case 2: split_kernel<2, MODE><<<grid, threads, 0, st>>>(kfs, k0, kc, frames, scratch, stride, seg_rows, kb); break;
    case 4: split_kernel<4, MODE><<<grid, threads, 0, st>>>(kfs, k0, kc, frames, scratch, stride, seg_rows, kb); break;
    default: split_kernel<8, MODE><<<grid, threads, 0, st>>>(kfs, k0, kc, frames, scratch, stride, seg_rows, kb); break;
  }
  return cudaGetLastError();
}

struct PairOut {  // optional per-query outputs of the pair API (device pointers)
  int32_t* train_idx = nullptr;
  int32_t* dist = nullptr;
  uint8_t* good = nullptr;
  int32_t* summary = nullptr;
};

int fill_best(lc_ctx* c, cudaStream_t st, uint64_t* d_best, size_t n) {
  if (!d_best || n == 0) return LC_OK;
  const int blocks = (int)std::min<size_t>((n + 255) / 256, 1024);
  fill_u64_kernel<<<blocks, 256, 0, st>>>(d_best, n, ~0ull);
  LC_CUDA(c, cudaGetLastError());
  c->launches += 1;
  c->timing.kernel_launches += 1;
  return LC_OK;
}

// Launch with programmatic stream serialization (lc_mxf4.cuh: pdl_trigger / pdl_wait): the kernel may be scheduled while its
// predecessor in the stream is still running and waits inside for its completion.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// Match n_frames query frames against n_kf keyframes (device tables) on `st`.
// h_frames mirrors d_frames (needed for nq).  max_rows = max n over the keyframes.
int run_match(lc_ctx* c, cudaStream_t st, const KfEntry* d_kfs, int n_kf, int max_rows, const QFrame* d_frames,
              const QFrame* h_frames, int n_frames, lc_kf_result* d_results, Knobs kb, bool full_tiles, uint64_t* d_best, size_t best_n) {
  // d_best[best_n]: the frames' per-query global bests, to be put to all ones before the first kernel folds into them
  const long long n_items = (long long)n_kf * n_frames;
  if (n_items == 0) return fill_best(c, st, d_best, best_n);
  if (n_items > 0x7fffffffll) return fail(c, LC_ERR_INVALID, "too many (frame, keyframe) items: %lld", n_items);
  int max_nq = 0;
  for (int f = 0; f < n_frames; ++f) max_nq = std::max(max_nq, h_frames[f].nq);
  const int mode = mode_of(c->cfg);

  // ---- tensor cores when there is a machine's worth of work: FP4 GEMM + max (lc_mxf4.cuh, the default) or int8 GEMM
  //      + arg-max (lc_tensor.cuh)
  const int kern = c->cfg.kernel;
  const bool want_fp4 = kern == LC_KERNEL_MXF4 || kern == LC_KERNEL_AUTO, want_i8 = kern == LC_KERNEL_TENSOR;
  if ((want_fp4 || want_i8) && max_nq > 0 && max_rows > 0) {
    // Queries per block.  int8 path: always 512.  FP4 path: 128 x T, T = 8 resident tiles when there are plenty of
    // keyframes; for small stores the T in {8, 4, 2} whose (block, keyframe) units fill the SMs in the fewest tile-times
    // (waves x T: 30 keyframes x 2 blocks of 1024 queries keep 60 SMs busy for 8 tile-times each, 4 blocks of 512 keep 120 busy
    // for 4).  T = 1 is left out: the four expander warps cannot deliver a train tile per MMA tile.
    int q_per_cta = want_fp4 ? fx::kQPerCta : tc::kQPerCta;
    if (want_fp4) {
      long long best_cost = -1;
      for (int T = fx::kQTilesPerCta; T >= 2; T >>= 1) {
        long long n_blocks = 0;
        for (int f = 0; f < n_frames; ++f) n_blocks += (h_frames[f].nq + T * fx::kQTile - 1) / (T * fx::kQTile);
        const long long waves = (n_blocks * n_kf + c->num_sms - 1) / c->num_sms;
        const long long cost = waves * T;
        // a smaller T re-expands every train tile for fewer queries: it has to win by 10 % to be taken
        if (best_cost < 0 || cost * 10 < best_cost * 9) best_cost = cost, q_per_cta = T * fx::kQTile;
      }
    }
    const size_t img_bytes = want_fp4 ? (size_t)fx::kAImageBytes : (size_t)tc::kABytes;
    std::vector<tc::ABlock> blocks;  // fx::ABlock has the same layout
    static_assert(sizeof(tc::ABlock) == sizeof(fx::ABlock), "ABlock layouts differ");
    std::vector<int> tiles;
    long long total_tiles = 0;
    for (int f = 0; f < n_frames; ++f)
      for (int q0 = 0; q0 < h_frames[f].nq; q0 += q_per_cta) {
        const int nq_blk = std::min(q_per_cta, h_frames[f].nq - q0);
        blocks.push_back(tc::ABlock{nullptr, f, q0, nq_blk, 0});
        tiles.push_back((nq_blk + tc::kQTile - 1) / tc::kQTile);
        total_tiles += tiles.back();
      }
    const int n_ab = (int)blocks.size();
    // Worth it from ~8 full keyframes on (measured, tools/loop_latency.py: at 67 keyframes x 2000 rows the tensor path takes
    // a third of the integer kernels' end-to-end time although its CTAs - one per (query block, keyframe) - fill half the SMs).
    if ((long long)n_ab * n_kf >= c->num_sms || (long long)total_tiles * tc::kQTile * n_kf * max_rows >= kTensorMinPairs) {
      // scratch: one distance (mxf4: u16) or key (int8: u32) per (frame, keyframe, query); rows start 16-byte aligned.
      // The keyframes go through in batches of kc_max bounded by the scratch budget.
      const int stride = (max_nq + 7) & ~7;
      const size_t elem = want_fp4 ? 2 : 4;
      long long kc_max = std::max<long long>(1, (long long)(kScratchBudgetBytes / ((size_t)stride * elem)) / n_frames);
      kc_max = std::min<long long>(kc_max, n_kf);
      // CTAs per query block (each takes every ns-th keyframe of a batch): proportional to the block's number of 128-row
      // tiles, their total a whole number m of waves of one CTA per SM.  m is searched: more waves quantise better
      // (128 query blocks on 148 SMs need 37 CTAs each = 32 waves), fewer reload the resident queries less often.
      std::vector<int> ns_best((size_t)n_ab, 1);
      double cost_best = 1e300;
      const bool cached = c->plan_kc == kc_max && c->plan_tiles == tiles;  // the plan depends on nothing else
      if (cached) ns_best = c->plan_ns;
      for (int m = 1; m <= 32 && !cached; ++m) {
        const long long total = (long long)m * c->num_sms;
        if (total < n_ab && m < 32) continue;
        std::vector<int> ns((size_t)n_ab);
        long long used = 0;
        for (int i = 0; i < n_ab; ++i) {
          ns[i] = (int)std::min<long long>(kc_max, std::max<long long>(1, total * tiles[i] / total_tiles));
          used += ns[i];
        }
        // hand the rounding remainder to the blocks whose CTAs carry the most work
        for (long long left = n_ab <= 256 ? total - used : 0; left > 0; --left) {  // (quadratic in the block count)
          int arg = -1;
          double w = -1;
          for (int i = 0; i < n_ab; ++i) {
            const double wi = (double)tiles[i] * (double)((kc_max + ns[i] - 1) / ns[i]);
            if (ns[i] < kc_max && wi > w) w = wi, arg = i;
          }
          if (arg < 0) break;
          ns[arg] += 1;
          used += 1;
        }
        double wmax = 0;
        for (int i = 0; i < n_ab; ++i) wmax = std::max(wmax, (double)tiles[i] * (double)((kc_max + ns[i] - 1) / ns[i]));
        // makespan in tile-keyframes: waves x the slowest CTA, plus one resident-query load (~a tenth of a keyframe) per wave
        const double waves = (double)((used + c->num_sms - 1) / c->num_sms);
        const double cost = waves * (wmax + 0.1 * tc::kQTilesPerCta);
        if (cost < cost_best * 0.995) cost_best = cost, ns_best = ns;
        if (n_ab <= 4 && m >= 2) break;  // a few blocks: one or two waves are always enough
      }
      if (!cached) c->plan_kc = kc_max, c->plan_tiles = tiles, c->plan_ns = ns_best;
      std::vector<tc::CtaWork> work;
      for (int i = 0; i < n_ab; ++i)
        for (int s = 0; s < ns_best[i]; ++s) work.push_back(tc::CtaWork{i, s, ns_best[i]});
      LC_TRY(ensure(c, c->d_qimg, (size_t)n_ab * img_bytes));
      const size_t tables_cap = c->d_ablocks.cap;
      LC_TRY(ensure(c, c->d_ablocks, (size_t)n_ab * sizeof(tc::ABlock) + work.size() * sizeof(tc::CtaWork)));
      if (c->d_ablocks.cap != tables_cap) c->tables_dev = nullptr;  // a new allocation holds nothing yet
      for (int i = 0; i < n_ab; ++i) blocks[i].img = (const uint8_t*)c->d_qimg.p + (size_t)i * img_bytes;
      tc::ABlock* d_blocks = (tc::ABlock*)c->d_ablocks.p;
      tc::CtaWork* d_work = (tc::CtaWork*)(d_blocks + n_ab);
      {  // one upload for both tables, skipped when the device copy is already this (same shapes, same buffers)
        std::vector<uint8_t> img((size_t)n_ab * sizeof(tc::ABlock) + work.size() * sizeof(tc::CtaWork));
        memcpy(img.data(), blocks.data(), (size_t)n_ab * sizeof(tc::ABlock));
        memcpy(img.data() + (size_t)n_ab * sizeof(tc::ABlock), work.data(), work.size() * sizeof(tc::CtaWork));
        if (c->tables_dev != (const void*)d_blocks || c->tables_host != img) {
          LC_CUDA(c, cudaMemcpyAsync(d_blocks, img.data(), img.size(), cudaMemcpyHostToDevice, st));
          c->tables_dev = d_blocks;
          c->tables_host.swap(img);
        }
      }
      if (want_fp4) {
        fx::expand_queries_fp4_kernel<<<dim3(n_ab, fx::kQTilesPerCta), fx::kExpandThreads, 0, st>>>((const fx::ABlock*)d_blocks, d_frames,
                                                                                                  (uint8_t*)c->d_qimg.p, d_best != nullptr);
      } else {
        LC_TRY(fill_best(c, st, d_best, best_n));
        tc::expand_queries_kernel<<<n_ab, 256, 0, st>>>(d_blocks, d_frames, (uint8_t*)c->d_qimg.p);
      }
      LC_CUDA(c, cudaGetLastError());
      if (!c->smem_attr_set) {  // per device, so per context
        LC_CUDA(c, cudaFuncSetAttribute(tc::tensor_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemAlloc));
        LC_CUDA(c, cudaFuncSetAttribute(fx::mxf4_match_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fx::kSmemAlloc));
        LC_CUDA(c, cudaFuncSetAttribute(fx::mxf4_match_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fx::kSmemAlloc));
        c->smem_attr_set = true;
      }
      LC_TRY(ensure(c, c->d_scratch, (size_t)kc_max * n_frames * stride * elem));
      c->launches += 1;
      c->timing.kernel_launches += 1;
      bool any_best = false;
      long long total_nq = 0;
      for (int f = 0; f < n_frames; ++f) {
        any_best = any_best || h_frames[f].best_packed != nullptr;
        total_nq += h_frames[f].nq;
      }
      // winning keyframe of each query -> winning row (OpenCV's lowest-index order): its own launch after the last batch,
      // or fused with the finalize step when there is a single batch
      const int resolve_grid = (int)std::max<long long>(1, std::min<long long>((total_nq + 7) / 8, 4ll * c->num_sms));
      const bool fused_finish = want_fp4 && any_best && kc_max >= n_kf;
      for (int k0 = 0; k0 < n_kf; k0 += (int)kc_max) {
        const int kc = (int)std::min<long long>(kc_max, n_kf - k0);
        if (want_fp4) {
          // every compared (frame, keyframe, query) distance is stored exactly once: no scratch initialisation
          {
            MatchTimer mt(c, st);
            // specialisation of an issue-bound kernel: full-width train tiles only
            auto launch = [&](auto kern) {
              return launch_pdl(kern, dim3((unsigned)work.size()), dim3(fx::kThreads), fx::kSmemAlloc, st, d_kfs, k0, kc, d_frames,
                                (const fx::ABlock*)d_blocks, (const tc::CtaWork*)d_work, (uint16_t*)c->d_scratch.p, stride, kb);
            };
            LC_CUDA(c, full_tiles ? launch(fx::mxf4_match_kernel<true>) : launch(fx::mxf4_match_kernel<false>));
          }
          const int n_fin = (kc * n_frames + fx::kFinWarps - 1) / fx::kFinWarps;
          if (fused_finish)
            LC_CUDA(c, launch_pdl(fx::finish_kernel, dim3(n_fin + resolve_grid), dim3(256), 0, st, n_fin, d_kfs, n_kf, d_frames, n_frames,
                                  (const uint16_t*)c->d_scratch.p, stride, d_results, kb));
          else
            LC_CUDA(c, launch_pdl(fx::finalize_dist_kernel, dim3(n_fin), dim3(fx::kFinWarps * 32), 0, st, d_kfs, n_kf, k0, kc, d_frames, n_frames,
                                  (const uint16_t*)c->d_scratch.p, stride, d_results, kb));
        } else {
          LC_CUDA(c, cudaMemsetAsync(c->d_scratch.p, 0xff, (size_t)kc * n_frames * stride * 4, st));
          {
            MatchTimer mt(c, st);
            tc::tensor_match_kernel<<<(unsigned)work.size(), tc::kThreads, tc::kSmemAlloc, st>>>(
                d_kfs, k0, kc, d_frames, d_blocks, d_work, (uint32_t*)c->d_scratch.p, stride, kb);
          }
          LC_CUDA(c, cudaGetLastError());
          finalize_kernel<<<kc * n_frames, 256, 0, st>>>(d_kfs, n_kf, k0, kc, d_frames, (const uint32_t*)c->d_scratch.p, stride,
                                                         d_results, kb);
        }
        LC_CUDA(c, cudaGetLastError());
        c->launches += 2;
        c->timing.kernel_launches += 2;
      }
      if (want_fp4 && any_best && !fused_finish) {
        fx::resolve_best_kernel<<<resolve_grid, 256, 0, st>>>(d_kfs, d_frames, n_frames);
        LC_CUDA(c, cudaGetLastError());
        c->launches += 1;
        c->timing.kernel_launches += 1;
      }
      c->timing.path = want_fp4 ? 4 : 3;
      return LC_OK;
    }
  }

  LC_TRY(fill_best(c, st, d_best, best_n));
  // ---- whole-keyframe scan: enough items to fill every CTA slot many times over
  if (max_nq <= kFusedMaxNq && max_nq > 0) {
    const Shape sh = scan_shape(max_nq);
    const int occ = mode == MODE_POPC ? scan_occupancy<MODE_POPC>(sh.rq, sh.threads) : scan_occupancy<MODE_CSA>(sh.rq, sh.threads);
    const long long slots = (long long)c->num_sms * occ;
    if (n_items >= 16 * slots) {
      LC_TRY(ensure(c, c->d_counter, 256));
      LC_CUDA(c, cudaMemsetAsync(c->d_counter.p, 0, sizeof(int), st));
      const dim3 grid((unsigned)std::min<long long>(slots, n_items));
      MatchTimer mt(c, st);
      cudaError_t e = mode == MODE_POPC
                          ? launch_scan_m<MODE_POPC>(sh.rq, grid, sh.threads, st, d_kfs, n_kf, d_frames, n_frames, d_results, (int*)c->d_counter.p, kb)
                          : launch_scan_m<MODE_CSA>(sh.rq, grid, sh.threads, st, d_kfs, n_kf, d_frames, n_frames, d_results, (int*)c->d_counter.p, kb);
      LC_CUDA(c, e);
      c->launches += 1;
      c->timing.kernel_launches += 1;
      c->timing.path = 1;
      return LC_OK;
    }
  }

  // ---- split: (row segment, frame x keyframe, query chunk) grid + finalize, in batches bounded by the scratch budget
  c->timing.path = 2;
  if (max_nq == 0 || max_rows == 0) {  // nothing to compare: finalize alone writes the "skipped" records
    max_nq = std::max(max_nq, 1);
  }
  // widest register tile (fewest LDS per pair) that still yields enough CTAs to fill the machine
  Shape sh{};
  int seg = kTileRows;
  for (int rq : {8, 4, 2}) {
    sh.rq = rq;
    sh.threads = std::min(256, std::max(32, ((max_nq + 32 * rq - 1) / (32 * rq)) * 32));
    sh.chunks = (max_nq + sh.threads * rq - 1) / (sh.threads * rq);
    seg = kTileRows;
    while (seg > 32 && n_items * ((max_rows + seg - 1) / seg) * sh.chunks < 4ll * c->num_sms) seg >>= 1;
    if (n_items * ((max_rows + seg - 1) / seg) * sh.chunks >= 4ll * c->num_sms && seg >= 64) break;
  }
  const int n_seg = std::max(1, (max_rows + seg - 1) / seg);
  const int stride = (max_nq + 3) & ~3;
  // batches of keyframes [k0, k0 + kc) x all frames, bounded by the scratch budget and gridDim.y
  long long kc_max = std::max<long long>(1, (long long)(kScratchBudgetBytes / ((size_t)stride * 4)) / n_frames);
  kc_max = std::min<long long>(std::min<long long>(kc_max, std::max(1, 65535 / n_frames)), n_kf);
  LC_TRY(ensure(c, c->d_scratch, (size_t)kc_max * n_frames * stride * 4));
  for (int k0 = 0; k0 < n_kf; k0 += (int)kc_max) {
    const int kc = (int)std::min<long long>(kc_max, n_kf - k0);
    const int nb = kc * n_frames;
    LC_CUDA(c, cudaMemsetAsync(c->d_scratch.p, 0xff, (size_t)nb * stride * 4, st));
    if (max_rows > 0) {
      const dim3 grid(n_seg, nb, sh.chunks);
      MatchTimer mt(c, st);
      cudaError_t e = mode == MODE_POPC
                          ? launch_split_m<MODE_POPC>(sh.rq, grid, sh.threads, st, d_kfs, k0, kc, d_frames, (uint32_t*)c->d_scratch.p, stride, seg, kb)
                          : launch_split_m<MODE_CSA>(sh.rq, grid, sh.threads, st, d_kfs, k0, kc, d_frames, (uint32_t*)c->d_scratch.p, stride, seg, kb);
      LC_CUDA(c, e);
      c->launches += 1;
      c->timing.kernel_launches += 1;
    }
    finalize_kernel<<<nb, 256, 0, st>>>(d_kfs, n_kf, k0, kc, d_frames, (const uint32_t*)c->d_scratch.p, stride, d_results, kb);
    LC_CUDA(c, cudaGetLastError());
    c->launches += 1;
    c->timing.kernel_launches += 1;
  }
  return LC_OK;
}

// the frame table of a loop check whose rows lie concatenated at d_q
int build_frames(lc_ctx* c, int n_frames, const void* d_q, const int32_t* nq, const int32_t* cur_ids, uint64_t* d_best,
                 std::vector<QFrame>& hf, size_t* rows) {
  if ((uintptr_t)d_q % 16 != 0) return fail(c, LC_ERR_INVALID, "d_q must be 16-byte aligned");
  hf.resize((size_t)n_frames);
  size_t row = 0;
  for (int f = 0; f < n_frames; ++f) {
    if (nq[f] < 0 || nq[f] > LC_MAX_ROWS_PER_KF) return fail(c, LC_ERR_INVALID, "nq[%d] = %d out of range", f, nq[f]);
    hf[f] = QFrame{(const uint4*)d_q + 2 * row, nq[f], cur_ids[f], d_best ? d_best + row : nullptr};
    row += (size_t)nq[f];
  }
  *rows = row;
  return LC_OK;
}

// loop check of the frames in hf (device copy of the table at d_frames) against the local store
int loop_check_run(lc_ctx* c, cudaStream_t st, const std::vector<QFrame>& hf, const QFrame* d_frames, size_t rows,
                   lc_kf_result* d_results, uint64_t* d_best) {
  LC_TRY(sync_kf_table(c));
  return run_match(c, st, (const KfEntry*)c->d_kf.p, (int)c->h_kf.size(), c->max_n_local, d_frames, hf.data(), (int)hf.size(), d_results,
                   knobs_of(c->cfg), c->n_partial_kf == 0, d_best, rows);
}

// loop check of n_frames device-resident frames (rows concatenated at d_q) against the local store
int loop_check_device_impl(lc_ctx* c, cudaStream_t st, int n_frames, const void* d_q, const int32_t* nq,
                           const int32_t* cur_ids, lc_kf_result* d_results, uint64_t* d_best) {
  if (n_frames <= 0) return LC_OK;
  std::vector<QFrame> hf;
  size_t rows = 0;
  LC_TRY(build_frames(c, n_frames, d_q, nq, cur_ids, d_best, hf, &rows));
  LC_TRY(ensure(c, c->d_frames, hf.size() * sizeof(QFrame)));
  // pageable source: the runtime stages it before returning, so `hf` may die with this frame
  LC_CUDA(c, cudaMemcpyAsync(c->d_frames.p, hf.data(), hf.size() * sizeof(QFrame), cudaMemcpyHostToDevice, st));
  return loop_check_run(c, st, hf, (const QFrame*)c->d_frames.p, rows, d_results, d_best);
}

// matchFeatures on device-resident rows: one launch of pair_kernel (lc_kernels.cuh); nq, nt > 0
int pair_device_impl(lc_ctx* c, cudaStream_t st, const void* d_q, int nq, const KfEntry& train, const PairOut& po) {
  const int nt = train.n;
  // keys[nq] and {ticket counter, smallest distance} rest at 0xffffffff / {0, 0xffffffff} between calls (the kernel's last CTA puts them back)
  // A CUDA failure anywhere on the context (an aborted launch, a failed copy) may have left them dirty: both are put back
  // before the next launch, or every later pair match on this context would be silently wrong.
  const size_t need = (size_t)nq * 4;
  const bool dirty = c->pair_state_dirty;
  c->pair_state_dirty = false;
  if (c->d_pairkeys.cap < need || dirty) {
    LC_TRY(ensure(c, c->d_pairkeys, need));
    LC_CUDA(c, cudaMemsetAsync(c->d_pairkeys.p, 0xff, c->d_pairkeys.cap, st));
  }
  if (!c->d_paircnt.p || dirty) {
    LC_TRY(ensure(c, c->d_paircnt, 256));
    static const unsigned int idle[2] = {0u, 0xffffffffu};  // ticket counter, smallest distance
    LC_CUDA(c, cudaMemcpyAsync(c->d_paircnt.p, idle, sizeof idle, cudaMemcpyHostToDevice, st));
  }
  // The launch is a whole number of waves of near-equal CTAs, two per SM: the integer pipes bound the kernel, so an SM
  // that gets one CTA more than its neighbour sets the call's latency.  Two queries per thread measured fastest across
  // shapes (tools/pair_shapes.cu: 4 M pairs in 12.4 us per back-to-back launch, 8.3 us of it fixed latency).
  constexpr int kRq = 2;
  const int threads = std::min(256, std::max(32, ((nq + 32 * kRq - 1) / (32 * kRq)) * 32));
  const int chunks = (nq + threads * kRq - 1) / (threads * kRq);
  const long long segs_per_wave = std::max<long long>(1, 2ll * c->num_sms / chunks);
  const long long waves = std::max<long long>(1, (nt + kTileRows * segs_per_wave - 1) / (kTileRows * segs_per_wave));
  const int seg = (int)std::min<long long>(kTileRows, std::max<long long>(kPairMinSegRows, (nt + waves * segs_per_wave - 1) / (waves * segs_per_wave)));
  const long long n_seg = (nt + seg - 1) / seg;
  if (chunks > 65535) return fail(c, LC_ERR_INVALID, "pair match: %d query rows need %d chunks (max 65535)", nq, chunks);
  const dim3 grid((unsigned)n_seg, (unsigned)chunks);
  const Knobs kb = knobs_of(c->cfg);
  const uint4* q4 = (const uint4*)d_q;
  uint32_t* keys = (uint32_t*)c->d_pairkeys.p;
  unsigned int* state = (unsigned int*)c->d_paircnt.p;
  if (mode_of(c->cfg) == MODE_POPC)
    pair_kernel<kRq, MODE_POPC><<<grid, threads, 0, st>>>(q4, nq, train.rows, nt, seg, keys, state, po.train_idx, po.dist, po.good, po.summary, kb);
  else
    pair_kernel<kRq, MODE_CSA><<<grid, threads, 0, st>>>(q4, nq, train.rows, nt, seg, keys, state, po.train_idx, po.dist, po.good, po.summary, kb);
  LC_CUDA(c, cudaGetLastError());
  c->launches += 1;
  c->timing.kernel_launches += 1;
  c->timing.path = 5;
  return LC_OK;
}

int stage_query(lc_ctx* c, DevBuf& buf, const uint8_t* q, int64_t n, size_t stride) {
  LC_TRY(ensure(c, buf, (size_t)std::max<int64_t>(n, 1) * LC_DESC_BYTES));
  if (n > 0) LC_TRY(upload_rows(c, (uint8_t*)buf.p, q, n, stride));
  return LC_OK;
}

// both sides of a pair match in ONE host-to-device copy: [query rows | train rows] packed in the pinned upload buffer and
// landed in d_q (the call is latency-bound; a second copy costs as much as the kernel).  Larger pairs take the two-copy road.
int stage_pair(lc_ctx* c, const uint8_t* q, int64_t nq, size_t qs, const uint8_t* t, int64_t nt, size_t ts, const uint4** d_t) {
  const size_t q_bytes = ((size_t)nq * LC_DESC_BYTES + 255) & ~size_t(255), t_bytes = (size_t)nt * LC_DESC_BYTES;
  if (q_bytes + t_bytes > 2 * kPinHalfBytes) {
    LC_TRY(stage_query(c, c->d_q, q, nq, qs));
    LC_TRY(stage_query(c, c->d_t, t, nt, ts));
    *d_t = (const uint4*)c->d_t.p;
    return LC_OK;
  }
  LC_TRY(ensure(c, c->d_q, q_bytes + t_bytes));
  // earlier asynchronous uploads (lc_loop_check_begin) may still be reading the pinned halves
  LC_CUDA(c, cudaEventSynchronize(c->pin_ev[0]));
  LC_CUDA(c, cudaEventSynchronize(c->pin_ev[1]));
  auto pack = [](uint8_t* dst, const uint8_t* src, int64_t n, size_t stride) {
    if (stride == LC_DESC_BYTES) {
      memcpy(dst, src, (size_t)n * LC_DESC_BYTES);
    } else {
      for (int64_t r = 0; r < n; ++r) memcpy(dst + (size_t)r * LC_DESC_BYTES, src + (size_t)r * stride, LC_DESC_BYTES);
    }
  };
  pack(c->pin_up, q, nq, qs);
  pack(c->pin_up + q_bytes, t, nt, ts);
  LC_CUDA(c, cudaMemcpyAsync(c->d_q.p, c->pin_up, q_bytes + t_bytes, cudaMemcpyHostToDevice, c->stream));
  // the caller synchronises the stream before it returns, so the pinned buffer is idle again by then
  *d_t = (const uint4*)((const uint8_t*)c->d_q.p + q_bytes);
  return LC_OK;
}

// events = false: the call's path and launch count are still reported, its times read 0
void timing_begin(lc_ctx* c, bool events = true) {
  if (c->pending.active && c->pending.queued) {
    // a call between lc_loop_check_begin and _end: the pending check owns the events, so this call is not timed (its
    // path and launch count are still reported until _end puts the check's timing back)
    if (!c->pending.stashed) {
      c->pending.stashed = true;
      c->pending.stashed_on = c->timing_on;
      c->pending.stashed_timing = c->timing;
    }
    c->timing = lc_timing{};
    c->timing_on = false;
    return;
  }
  c->timing = lc_timing{};
  c->match_ev_used = 0;
  c->timing_on = events;
  if (events) cudaEventRecord(c->ev[0], c->stream);
}
// a pair match is a ~50 us call in which every event record costs ~3 us of stream time: timed only on request
bool pair_events(const lc_ctx* c) { return c->cfg.timing == LC_TIMING_ON; }
// a loop check is timed by default once it is large enough for six event records (~20 us) not to matter
bool loop_events(const lc_ctx* c, int64_t query_rows) {
  return c->cfg.timing == LC_TIMING_ON || (double)query_rows * (double)c->l_rows >= kTimedMinPairs;
}
// has_d2h = false: the kernels wrote their results straight into mapped pinned memory; ev[2] closes the call
// record_last = false: ev[3] was recorded when the work was queued (lc_loop_check_begin), not now
int timing_end(lc_ctx* c, bool has_d2h = true, bool record_last = true) {
  if (!c->timing_on) {
    LC_CUDA(c, cudaStreamSynchronize(c->stream));
    return LC_OK;
  }
  cudaEvent_t last = c->ev[has_d2h ? 3 : 2];
  if (has_d2h && record_last) LC_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
  LC_CUDA(c, cudaStreamSynchronize(c->stream));
  cudaEventElapsedTime(&c->timing.h2d_ms, c->ev[0], c->ev[1]);
  cudaEventElapsedTime(&c->timing.kernel_ms, c->ev[1], c->ev[2]);
  if (has_d2h) cudaEventElapsedTime(&c->timing.d2h_ms, c->ev[2], c->ev[3]);
  cudaEventElapsedTime(&c->timing.total_ms, c->ev[0], last);
  c->timing_on = false;
  if (c->timing.path == 5) {  // the pair kernel is the call's only kernel
    c->timing.match_ms = c->timing.kernel_ms;
    c->timing.match_launches = c->timing.kernel_launches;
  }
  for (size_t i = 0; i + 1 < c->match_ev_used; i += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->match_ev[i], c->match_ev[i + 1]) == cudaSuccess) {
      c->timing.match_ms += ms;
      c->timing.match_launches += 1;
    }
  }
  return LC_OK;
}

int pair_host_impl(lc_ctx* c, int32_t nq, const KfEntry& train, int32_t* train_idx, int32_t* dist, uint8_t* good,
                   int32_t* n_matches, int32_t* min_dist, int32_t* n_good) {
  // outputs: idx[nq] dist[nq] summary[2] good[nq].  Up to kPairDirectOutBytes the kernel's finalize step stores them
  // straight into the pinned result buffer (mapped into the device's address space under unified addressing: posted
  // writes over PCIe, visible to the host when the stream has drained) - a device-to-host copy would add a stream
  // operation that costs as much as the kernel.  Larger results go through device memory and one copy.
  const size_t off_dist = (size_t)nq * 4, off_sum = off_dist + (size_t)nq * 4, off_good = off_sum + 8;
  const size_t out_bytes = off_good + (size_t)nq;
  const bool direct = out_bytes <= kPairDirectOutBytes;
  LC_TRY(ensure_pin_down(c, out_bytes));
  if (!direct) LC_TRY(ensure(c, c->d_out, out_bytes));
  uint8_t* d = direct ? c->pin_down_dev : (uint8_t*)c->d_out.p;
  PairOut po{(int32_t*)d, (int32_t*)(d + off_dist), d + off_good, (int32_t*)(d + off_sum)};
  if (c->timing_on) LC_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
  LC_TRY(pair_device_impl(c, c->stream, c->d_q.p, nq, train, po));
  if (c->timing_on) LC_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
  if (!direct) LC_CUDA(c, cudaMemcpyAsync(c->pin_down, d, out_bytes, cudaMemcpyDeviceToHost, c->stream));
  LC_TRY(timing_end(c, !direct));
  const uint8_t* h = c->pin_down;
  if (train_idx) memcpy(train_idx, h, (size_t)nq * 4);
  if (dist) memcpy(dist, h + off_dist, (size_t)nq * 4);
  if (good) memcpy(good, h + off_good, (size_t)nq);
  int32_t sum[2];
  memcpy(sum, h + off_sum, 8);
  if (n_matches) *n_matches = nq;
  if (min_dist) *min_dist = sum[0];
  if (n_good) *n_good = sum[1];
  return LC_OK;
}

void empty_pair(int32_t* n_matches, int32_t* min_dist, int32_t* n_good) {
  if (n_matches) *n_matches = 0;
  if (min_dist) *min_dist = -1;
  if (n_good) *n_good = 0;
}

int check_rows(lc_ctx* c, const char* what, const uint8_t* p, int32_t n, size_t stride) {
  if (n < 0 || n > LC_MAX_ROWS_PER_KF) return fail(c, LC_ERR_INVALID, "%s: n = %d outside [0, %d]", what, n, LC_MAX_ROWS_PER_KF);
  if (n > 0 && !p) return fail(c, LC_ERR_INVALID, "%s: NULL descriptors with n = %d", what, n);
  if (n > 0 && stride < LC_DESC_BYTES) return fail(c, LC_ERR_INVALID, "%s: row_stride %zu < 32", what, stride);
  return LC_OK;
}

}  // namespace

// =================================================================== extern "C"
#define LC_API extern "C" __attribute__((visibility("default")))
// every entry point that touches the GPU runs inside an NVTX range named after it (ncu --nvtx --nvtx-include "lc_loop_check/",
// Nsight Systems timelines): the tracing hook of SURVEY.md section 5
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
#define LC_GUARD_BEGIN NvtxRange nvtx_range_(__func__); try {
#define LC_GUARD_END(c)                                          \
  }                                                              \
  catch (const std::bad_alloc&) { return fail((c), LC_ERR_NOMEM, "host allocation failed"); } \
  catch (const std::exception& e) { return fail((c), LC_ERR_INVALID, "%s", e.what()); }       \
  catch (...) { return fail((c), LC_ERR_INVALID, "unknown C++ exception"); }

// a timing-experiment build (-DLC_FX_DRAIN_EXPERIMENT=n, wrong results by design) answers with another version number, so it can
// never pass for the library: tests/test_abi_cpu.py compares this value with the header's
LC_API int32_t lc_abi_version(void) { return LC_ABI_VERSION + (LC_FX_DRAIN_EXPERIMENT ? 1000 : 0); }

LC_API void lc_default_config(lc_config* g) {
  if (!g) return;
  memset(g, 0, sizeof *g);
  g->struct_size = sizeof *g;
  g->device = 0;
  g->loop_threshold = 0.7;  // include/loop_closing.hpp:31
  g->min_loop_gap = 30;
  g->min_good = 50;
  g->filter_floor = 0;
  g->filter_strict = 0;
  g->min_init = -1;
  g->gap_strict = 0;
  g->thr_inclusive = 0;
  g->shard_rank = 0;
  g->shard_world = 1;
  g->kernel = LC_KERNEL_AUTO;
  g->timing = LC_TIMING_AUTO;
  g->slab_rows = 0;
}

LC_API const char* lc_last_error(const lc_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

LC_API void lc_destroy(lc_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->cfg.device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (int r = 0; r < kMaxPeers; ++r)
    if (c->xchg.ipc[r] && c->xchg.peer[r]) cudaIpcCloseMemHandle(c->xchg.peer[r]);
  if (c->xchg.base) cudaFree(c->xchg.base);
  for (void* p : c->slabs) cudaFree(p);
  for (void* p : c->l2s.slabs) cudaFree(p);
  for (DevBuf* b : {&c->l2s.d_kf, &c->d_l2q, &c->d_l2t, &c->d_l2norm, &c->d_l2res, &c->d_l2frames, &c->d_l2out, &c->d_l2img, &c->d_l2tables})
    if (b->p) cudaFree(b->p);
  for (DevBuf* b : {&c->d_kf, &c->d_q, &c->d_t, &c->d_results, &c->d_scratch, &c->d_frames, &c->d_pairkeys, &c->d_paircnt, &c->d_out, &c->d_counter, &c->d_qimg, &c->d_ablocks})
    if (b->p) cudaFree(b->p);
  if (c->pin_up) cudaFreeHost(c->pin_up);
  if (c->pin_down) cudaFreeHost(c->pin_down);
  if (c->pin_loop) cudaFreeHost(c->pin_loop);
  if (c->pin_frame) cudaFreeHost(c->pin_frame);
  if (c->pair_ev) cudaEventDestroy(c->pair_ev);
  for (auto& e : c->ev) if (e) cudaEventDestroy(e);
  for (auto& e : c->pin_ev) if (e) cudaEventDestroy(e);
  if (c->own_ev) cudaEventDestroy(c->own_ev);
  if (c->foreign_ev) cudaEventDestroy(c->foreign_ev);
  for (auto& e : c->match_ev) if (e) cudaEventDestroy(e);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

LC_API int32_t lc_create(const lc_config* cfg, lc_ctx** out) {
  if (!out) return fail(nullptr, LC_ERR_INVALID, "out is NULL");
  *out = nullptr;
  lc_ctx* c = nullptr;
  LC_GUARD_BEGIN
  lc_config g;
  lc_default_config(&g);
  if (cfg) {
    if (cfg->struct_size != sizeof(lc_config)) return fail(nullptr, LC_ERR_INVALID, "lc_config.struct_size %u != %zu", cfg->struct_size, sizeof(lc_config));
    g = *cfg;
  }
  LC_TRY(validate(nullptr, g));
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, LC_ERR_CUDA, "no usable CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
  if (g.device < 0 || g.device >= ndev) return fail(nullptr, LC_ERR_INVALID, "device %d not in [0, %d)", g.device, ndev);
  c = new lc_ctx();
  c->cfg = g;
  c->slab_rows = g.slab_rows > 0 ? g.slab_rows : kDefaultSlabRows;
  auto bail = [&](const char* what, cudaError_t err) {
    fail(nullptr, LC_ERR_CUDA, "%s: %s", what, cudaGetErrorString(err));
    lc_destroy(c);
    return (int32_t)LC_ERR_CUDA;
  };
  if ((e = cudaSetDevice(g.device)) != cudaSuccess) return bail("cudaSetDevice", e);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, g.device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
  if (prop.major != 10) {
    fail(nullptr, LC_ERR_CUDA, "device %d is sm_%d%d; liblc_b200 carries sm_100a code only", g.device, prop.major, prop.minor);
    lc_destroy(c);
    return LC_ERR_CUDA;
  }
  c->num_sms = prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  for (auto& ev : c->ev) if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail("cudaEventCreate", e);
  for (auto& ev : c->pin_ev) {
    if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    cudaEventRecord(ev, c->stream);
  }
  if ((e = cudaEventCreateWithFlags(&c->own_ev, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
  if ((e = cudaEventCreateWithFlags(&c->pair_ev, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
  if ((e = cudaEventCreateWithFlags(&c->foreign_ev, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
  if ((e = cudaMallocHost((void**)&c->pin_up, 2 * kPinHalfBytes)) != cudaSuccess) return bail("cudaMallocHost", e);
  *out = c;
  return LC_OK;
  LC_GUARD_END(nullptr)
}

LC_API int32_t lc_get_config(const lc_ctx* c, lc_config* out) {
  if (!c || !out) return LC_ERR_INVALID;
  *out = c->cfg;
  return LC_OK;
}

LC_API int32_t lc_set_params(lc_ctx* c, const lc_config* g) {
  if (!c) return LC_ERR_INVALID;
  if (!g || g->struct_size != sizeof(lc_config)) return fail(c, LC_ERR_INVALID, "bad lc_config");
  if (g->device != c->cfg.device || g->shard_rank != c->cfg.shard_rank || g->shard_world != c->cfg.shard_world)
    return fail(c, LC_ERR_INVALID, "device and sharding cannot change after lc_create");
  LC_TRY(validate(c, *g));
  const int64_t slab = c->cfg.slab_rows;
  c->cfg = *g;
  c->cfg.slab_rows = slab;
  return LC_OK;
}

// ------------------------------------------------------------------- store
LC_API int32_t lc_db_append(lc_ctx* c, int32_t frame_id, const uint8_t* desc, int32_t n, size_t stride, int64_t* kf_index) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  LC_TRY(enter_host(c));
  return store_append(c, frame_id, desc, n, stride, false, kf_index);
  LC_GUARD_END(c)
}

LC_API int32_t lc_db_append_many(lc_ctx* c, int32_t count, const int32_t* frame_ids, const int32_t* n_rows, const uint8_t* const* desc) {
  if (!c) return LC_ERR_INVALID;
  if (count < 0 || (count > 0 && (!frame_ids || !n_rows || !desc))) return fail(c, LC_ERR_INVALID, "bad arguments");
  LC_GUARD_BEGIN
  LC_TRY(enter_host(c));
  for (int32_t i = 0; i < count; ++i) LC_TRY(store_append(c, frame_ids[i], desc[i], n_rows[i], LC_DESC_BYTES, false, nullptr));
  return LC_OK;
  LC_GUARD_END(c)
}

// device-resident source rows: d_rows holds the rows of the OWNED keyframes of this call, back to back.  The copies
// run on the context's own stream (the store's order) behind an event recorded on `cuda_stream`, the stream that
// produced d_rows; that stream in turn waits for the copies, so stream-ordered reuse of d_rows is safe.
LC_API int32_t lc_db_append_many_device(lc_ctx* c, int32_t count, const int32_t* frame_ids, const int32_t* n_rows, const void* d_rows,
                                        void* cuda_stream) {
  if (!c) return LC_ERR_INVALID;
  if (count < 0 || (count > 0 && (!frame_ids || !n_rows))) return fail(c, LC_ERR_INVALID, "bad arguments");
  LC_GUARD_BEGIN
  LC_TRY(enter_host(c));
  cudaStream_t producer = (cudaStream_t)cuda_stream;
  if (producer && producer != c->stream) {
    LC_CUDA(c, cudaEventRecord(c->foreign_ev, producer));
    LC_CUDA(c, cudaStreamWaitEvent(c->stream, c->foreign_ev, 0));
  }
  const uint8_t* p = (const uint8_t*)d_rows;
  int rc = LC_OK;
  for (int32_t i = 0; i < count && rc == LC_OK; ++i) {
    const bool own = (c->g_kf % c->cfg.shard_world) == c->cfg.shard_rank;
    rc = store_append(c, frame_ids[i], own ? p : nullptr, n_rows[i], LC_DESC_BYTES, true, nullptr);
    if (own) p += (size_t)n_rows[i] * LC_DESC_BYTES;
  }
  if (producer && producer != c->stream) {  // also after a partial failure: some copies may be queued
    LC_CUDA(c, cudaEventRecord(c->own_ev, c->stream));
    LC_CUDA(c, cudaStreamWaitEvent(producer, c->own_ev, 0));
  }
  return rc;
  LC_GUARD_END(c)
}

LC_API int32_t lc_db_size(const lc_ctx* c, int64_t* n_kf, int64_t* n_rows, int64_t* n_local_kf, int64_t* n_local_rows) {
  if (!c) return LC_ERR_INVALID;
  if (n_kf) *n_kf = c->g_kf;
  if (n_rows) *n_rows = c->g_rows;
  if (n_local_kf) *n_local_kf = (int64_t)c->h_kf.size();
  if (n_local_rows) *n_local_rows = c->l_rows;
  return LC_OK;
}

LC_API int32_t lc_db_clear(lc_ctx* c) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  LC_TRY(enter_host(c));
  LC_CUDA(c, cudaStreamSynchronize(c->stream));
  for (void* p : c->slabs) cudaFree(p);
  c->slabs.clear();
  c->last_slab_cap = c->last_slab_used = 0;
  c->h_kf.clear();
  c->h_kf_global.clear();
  c->d_kf_synced = 0;
  c->g_kf = c->g_rows = c->l_rows = 0;
  c->max_n_local = 0;
  c->n_partial_kf = 0;
  return LC_OK;
  LC_GUARD_END(c)
}

static int find_local(lc_ctx* c, int64_t kf_index, size_t* slot) {
  if (kf_index < 0 || kf_index >= c->g_kf) return fail(c, LC_ERR_INVALID, "kf_index %lld not in [0, %lld)", (long long)kf_index, (long long)c->g_kf);
  if (kf_index % c->cfg.shard_world != c->cfg.shard_rank) return fail(c, LC_ERR_INVALID, "keyframe %lld is not stored on shard %d", (long long)kf_index, c->cfg.shard_rank);
  *slot = (size_t)(kf_index / c->cfg.shard_world);
  return LC_OK;
}

LC_API int32_t lc_db_read(lc_ctx* c, int64_t kf_index, uint8_t* desc_out, int32_t cap_rows, int32_t* n, int32_t* frame_id) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  LC_TRY(enter_host(c));
  size_t s;
  LC_TRY(find_local(c, kf_index, &s));
  const KfEntry& e = c->h_kf[s];
  if (n) *n = e.n;
  if (frame_id) *frame_id = e.frame_id;
  if (desc_out) {
    if (cap_rows < e.n) return fail(c, LC_ERR_CAPACITY, "cap_rows %d < %d", cap_rows, e.n);
    if (e.n > 0) {
      LC_CUDA(c, cudaMemcpyAsync(desc_out, e.rows, (size_t)e.n * LC_DESC_BYTES, cudaMemcpyDeviceToHost, c->stream));
      LC_CUDA(c, cudaStreamSynchronize(c->stream));
    }
  }
  return LC_OK;
  LC_GUARD_END(c)
}

// file format: "LCB200DB" u32 version, u32 shard_rank, u32 shard_world, i64 g_kf, i64 g_rows, i64 n_local,
// then per local keyframe {i64 global_index, i64 global_row_off, i32 frame_id, i32 n}, then the rows.
LC_API int32_t lc_db_save(lc_ctx* c, const char* path) {
  if (!c || !path) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  LC_TRY(enter_host(c));
  LC_CUDA(c, cudaStreamSynchronize(c->stream));
  FILE* f = fopen(path, "wb");
  if (!f) return fail(c, LC_ERR_IO, "cannot open %s for writing", path);
  bool ok = true;
  auto W = [&](const void* p, size_t n) { ok = ok && fwrite(p, 1, n, f) == n; };
  const uint32_t hdr[3] = {1u, (uint32_t)c->cfg.shard_rank, (uint32_t)c->cfg.shard_world};
  const int64_t cnt[3] = {c->g_kf, c->g_rows, (int64_t)c->h_kf.size()};
  W("LCB200DB", 8);
  W(hdr, sizeof hdr);
  W(cnt, sizeof cnt);
  for (size_t i = 0; i < c->h_kf.size(); ++i) {
    const int64_t a[2] = {c->h_kf_global[i], c->h_kf[i].global_row_off};
    const int32_t b[2] = {c->h_kf[i].frame_id, c->h_kf[i].n};
    W(a, sizeof a);
    W(b, sizeof b);
  }
  std::vector<uint8_t> buf;
  for (size_t i = 0; i < c->h_kf.size() && ok; ++i) {
    const KfEntry& e = c->h_kf[i];
    if (e.n == 0) continue;
    buf.resize((size_t)e.n * LC_DESC_BYTES);
    if (cudaMemcpy(buf.data(), e.rows, buf.size(), cudaMemcpyDeviceToHost) != cudaSuccess) {
      fclose(f);
      return fail(c, LC_ERR_CUDA, "device read failed while saving");
    }
    W(buf.data(), buf.size());
  }
  ok = (fclose(f) == 0) && ok;
  if (!ok) return fail(c, LC_ERR_IO, "short write to %s", path);
  return LC_OK;
  LC_GUARD_END(c)
}

LC_API int32_t lc_db_load(lc_ctx* c, const char* path) {
  if (!c || !path) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  LC_TRY(enter_host(c));
  if (c->g_kf != 0 || !c->h_kf.empty()) return fail(c, LC_ERR_INVALID, "lc_db_load needs an empty store (call lc_db_clear)");
  struct File {  // closed on every path out, exceptions included
    FILE* f;
    ~File() { if (f) fclose(f); }
  } file{fopen(path, "rb")};
  FILE* f = file.f;
  if (!f) return fail(c, LC_ERR_IO, "cannot open %s", path);
  bool ok = true;
  auto R = [&](void* p, size_t n) { ok = ok && fread(p, 1, n, f) == n; };
  char magic[8];
  uint32_t hdr[3];
  int64_t cnt[3];
  R(magic, 8);
  R(hdr, sizeof hdr);
  R(cnt, sizeof cnt);
  if (!ok || memcmp(magic, "LCB200DB", 8) != 0 || hdr[0] != 1u) return fail(c, LC_ERR_IO, "%s is not a liblc_b200 store file", path);
  const int64_t rank = c->cfg.shard_rank, world = c->cfg.shard_world;
  if ((int64_t)hdr[1] != rank || (int64_t)hdr[2] != world)
    return fail(c, LC_ERR_INVALID, "%s holds shard %u/%u, context is %d/%d", path, hdr[1], hdr[2], c->cfg.shard_rank, c->cfg.shard_world);
  // The header is not trusted: the counts must describe exactly this shard's slice of a store of g_kf keyframes
  // (keyframe k on rank k % world, local slot k / world - what find_local and the loop check's records rely on).
  const int64_t g_kf = cnt[0], g_rows = cnt[1], n_local = cnt[2];
  if (g_kf < 0 || g_rows < 0 || g_rows > LC_MAX_STORE_ROWS || n_local < 0 || n_local != (g_kf > rank ? (g_kf - rank + world - 1) / world : 0))
    return fail(c, LC_ERR_IO, "%s: header counts are inconsistent (%lld keyframes, %lld rows, %lld local on shard %lld/%lld)", path,
                (long long)g_kf, (long long)g_rows, (long long)n_local, (long long)rank, (long long)world);
  struct Rec { int64_t gi, off; int32_t fid, n; };
  {  // the table must fit the file: a corrupt count must not become a huge allocation
    const long pos = ftell(f);
    if (pos < 0 || fseek(f, 0, SEEK_END) != 0) return fail(c, LC_ERR_IO, "%s: cannot seek", path);
    const long end = ftell(f);
    if (end < 0 || fseek(f, pos, SEEK_SET) != 0) return fail(c, LC_ERR_IO, "%s: cannot seek", path);
    if ((uint64_t)n_local > (uint64_t)(end - pos) / 24u) return fail(c, LC_ERR_IO, "%s is truncated or corrupt (keyframe table)", path);
  }
  std::vector<Rec> recs((size_t)n_local);
  int64_t local_rows = 0, prev_end = 0;
  for (size_t i = 0; i < recs.size() && ok; ++i) {
    Rec& r = recs[i];
    R(&r.gi, 16);
    R(&r.fid, 8);
    if (!ok) break;
    // global index = slot * world + rank; 0 <= n <= LC_MAX_ROWS_PER_KF; the global row ranges ascend without overlap
    if (r.gi != (int64_t)i * world + rank || r.n < 0 || r.n > LC_MAX_ROWS_PER_KF || r.off < prev_end || r.off > g_rows - r.n)
      return fail(c, LC_ERR_IO, "%s: keyframe table entry %zu is corrupt (global index %lld, row offset %lld, %d rows)", path, i,
                  (long long)r.gi, (long long)r.off, r.n);
    prev_end = r.off + r.n;
    local_rows += r.n;
  }
  if (!ok) return fail(c, LC_ERR_IO, "%s is truncated or corrupt (keyframe table)", path);
  if (world == 1 && (local_rows != g_rows || (n_local > 0 && recs[0].off != 0)))
    return fail(c, LC_ERR_IO, "%s: row counts are inconsistent (%lld rows in the table, %lld in the header)", path, (long long)local_rows, (long long)g_rows);
  // from here on the store is being filled: every failure empties it again
  int rc = LC_OK;
  std::vector<uint8_t> buf;
  for (size_t i = 0; i < recs.size() && ok && rc == LC_OK; ++i) {
    const Rec& r = recs[i];
    buf.resize((size_t)r.n * LC_DESC_BYTES);
    if (r.n) R(buf.data(), buf.size());
    if (!ok) break;
    uint4* rows = nullptr;
    rc = slab_alloc(c, r.n, &rows);
    if (rc == LC_OK && r.n) rc = upload_rows(c, (uint8_t*)rows, buf.data(), r.n, LC_DESC_BYTES);
    if (rc != LC_OK) break;
    c->h_kf.push_back(KfEntry{rows, r.n, r.fid, r.off});
    c->h_kf_global.push_back(r.gi);
    c->l_rows += r.n;
    c->max_n_local = std::max(c->max_n_local, r.n);
    c->n_partial_kf += (r.n > 0 && !fx::full_tiles_only(r.n)) ? 1 : 0;
  }
  if (ok && rc == LC_OK && fgetc(f) != EOF) ok = false;  // trailing bytes: not a file lc_db_save wrote
  if (rc == LC_OK && ok && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(c, LC_ERR_CUDA, "upload failed while loading %s", path);
  if (!ok || rc != LC_OK) {
    const std::string why = c->err;
    lc_db_clear(c);
    if (!ok) return fail(c, LC_ERR_IO, "%s is truncated or corrupt (descriptor rows)", path);
    c->err = why;
    return rc;
  }
  c->g_kf = g_kf;
  c->g_rows = g_rows;
  return LC_OK;
  LC_GUARD_END(c)
}

// ------------------------------------------------------------------- matchFeatures
LC_API int32_t lc_match_pair(lc_ctx* c, const uint8_t* q, int32_t nq, size_t qs, const uint8_t* t, int32_t nt, size_t ts,
                             int32_t* train_idx, int32_t* dist, uint8_t* good, int32_t* n_matches, int32_t* min_dist, int32_t* n_good) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  LC_TRY(check_rows(c, "query", q, nq, qs));
  LC_TRY(check_rows(c, "train", t, nt, ts));
  if (nq == 0 || nt == 0) {  // cv::BFMatcher returns no matches when either side is empty
    empty_pair(n_matches, min_dist, n_good);
    return LC_OK;
  }
  LC_TRY(enter_host(c));
  timing_begin(c, pair_events(c));
  const uint4* d_t = nullptr;
  LC_TRY(stage_pair(c, q, nq, qs, t, nt, ts, &d_t));
  KfEntry train{d_t, nt, 0, 0};
  return pair_host_impl(c, nq, train, train_idx, dist, good, n_matches, min_dist, n_good);
  LC_GUARD_END(c)
}

LC_API int32_t lc_match_vs_keyframe(lc_ctx* c, const uint8_t* q, int32_t nq, size_t qs, int64_t kf_index, int32_t* train_idx,
                                    int32_t* dist, uint8_t* good, int32_t* n_matches, int32_t* min_dist, int32_t* n_good) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  LC_TRY(check_rows(c, "query", q, nq, qs));
  size_t s;
  LC_TRY(find_local(c, kf_index, &s));
  const KfEntry train = c->h_kf[s];
  if (nq == 0 || train.n == 0) {
    empty_pair(n_matches, min_dist, n_good);
    return LC_OK;
  }
  LC_TRY(enter_host(c));
  timing_begin(c, pair_events(c));
  LC_TRY(stage_query(c, c->d_q, q, nq, qs));
  return pair_host_impl(c, nq, train, train_idx, dist, good, n_matches, min_dist, n_good);
  LC_GUARD_END(c)
}

LC_API int32_t lc_match_pair_device(lc_ctx* c, const void* d_q, int32_t nq, const void* d_t, int32_t nt, int32_t* d_train_idx,
                                    int32_t* d_dist, uint8_t* d_good, int32_t* d_summary, void* cuda_stream) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  if (nq < 0 || nt < 0 || nq > LC_MAX_ROWS_PER_KF || nt > LC_MAX_ROWS_PER_KF) return fail(c, LC_ERR_INVALID, "row count out of range");
  if (((uintptr_t)d_q | (uintptr_t)d_t) % 16 != 0) return fail(c, LC_ERR_INVALID, "device rows must be 16-byte aligned");
  LC_CUDA(c, cudaSetDevice(c->cfg.device));
  StreamScope scope(c, cuda_stream);
  LC_TRY(scope.enter());
  cudaStream_t st = scope.st;
  c->timing_on = false;  // asynchronous entry point: nothing is timed
  if (nq == 0 || nt == 0) {
    if (d_summary) {
      const int32_t s[2] = {-1, 0};
      LC_CUDA(c, cudaMemcpyAsync(d_summary, s, sizeof s, cudaMemcpyHostToDevice, st));
    }
    return LC_OK;
  }
  KfEntry train{(const uint4*)d_t, nt, 0, 0};
  return pair_device_impl(c, st, d_q, nq, train, PairOut{d_train_idx, d_dist, d_good, d_summary});
  LC_GUARD_END(c)
}

LC_API int32_t lc_match_keyframes(lc_ctx* c, const uint8_t* q, int32_t nq, size_t qs, int32_t n_sel, const int64_t* kf_index,
                                  lc_match* matches, int64_t cap, int64_t* offsets, int32_t* min_dist) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  LC_TRY(check_rows(c, "query", q, nq, qs));
  if (n_sel < 0 || (n_sel > 0 && !kf_index) || !offsets || cap < 0 || (cap > 0 && !matches)) return fail(c, LC_ERR_INVALID, "bad arguments");
  std::vector<KfEntry> sel((size_t)n_sel);
  for (int32_t i = 0; i < n_sel; ++i) {
    size_t s;
    LC_TRY(find_local(c, kf_index[i], &s));
    sel[i] = c->h_kf[s];
  }
  offsets[0] = 0;
  if (n_sel == 0) return LC_OK;
  LC_TRY(enter_host(c));
  timing_begin(c);
  LC_TRY(stage_query(c, c->d_q, q, nq, qs));
  // per keyframe on the device: idx[nq] dist[nq] summary[2] good[nq], slices 16-byte aligned
  const size_t off_dist = (size_t)nq * 4, off_sum = off_dist + (size_t)nq * 4, off_good = off_sum + 8;
  const size_t slice = (off_good + (size_t)nq + 15) & ~size_t(15);
  LC_TRY(ensure(c, c->d_out, slice * (size_t)n_sel));
  LC_TRY(ensure_pin_down(c, slice * (size_t)n_sel));
  if (c->timing_on) LC_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
  for (int32_t i = 0; i < n_sel; ++i) {
    uint8_t* d = (uint8_t*)c->d_out.p + slice * (size_t)i;
    if (nq == 0 || sel[i].n == 0) {  // cv::BFMatcher returns no matches when either side is empty
      const int32_t none[2] = {-1, 0};
      LC_CUDA(c, cudaMemcpyAsync(d + off_sum, none, sizeof none, cudaMemcpyHostToDevice, c->stream));
      continue;
    }
    LC_TRY(pair_device_impl(c, c->stream, c->d_q.p, nq, sel[i], PairOut{(int32_t*)d, (int32_t*)(d + off_dist), d + off_good, (int32_t*)(d + off_sum)}));
  }
  if (c->timing_on) LC_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
  LC_CUDA(c, cudaMemcpyAsync(c->pin_down, c->d_out.p, slice * (size_t)n_sel, cudaMemcpyDeviceToHost, c->stream));
  LC_TRY(timing_end(c));
  int64_t total = 0;
  for (int32_t i = 0; i < n_sel; ++i) {
    const uint8_t* h = c->pin_down + slice * (size_t)i;
    int32_t sum[2];
    memcpy(sum, h + off_sum, 8);
    if (min_dist) min_dist[i] = sum[0];
    const bool empty = nq == 0 || sel[i].n == 0;
    const int32_t* idx = (const int32_t*)h;
    const int32_t* dist = (const int32_t*)(h + off_dist);
    const uint8_t* good = h + off_good;
    for (int32_t j = 0; !empty && j < nq; ++j) {
      if (!good[j]) continue;
      if (total < cap) matches[total] = lc_match{j, idx[j], dist[j]};
      ++total;
    }
    offsets[i + 1] = total;
  }
  if (total > cap) return fail(c, LC_ERR_CAPACITY, "matches_cap %lld < %lld", (long long)cap, (long long)total);
  return LC_OK;
  LC_GUARD_END(c)
}

// ------------------------------------------------------------------- detectLoops
LC_API int32_t lc_loop_check_batch_device(lc_ctx* c, int32_t n_frames, const void* d_q, const int32_t* nq, const int32_t* cur_ids,
                                          lc_kf_result* d_results, uint64_t* d_best, void* cuda_stream) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  if (n_frames < 0 || (n_frames > 0 && (!nq || !cur_ids))) return fail(c, LC_ERR_INVALID, "bad arguments");
  if (!d_results && !c->h_kf.empty() && n_frames > 0) return fail(c, LC_ERR_INVALID, "d_results is NULL");
  LC_CUDA(c, cudaSetDevice(c->cfg.device));
  StreamScope scope(c, cuda_stream);
  LC_TRY(scope.enter());
  cudaStream_t st = scope.st;
  c->timing_on = false;  // asynchronous entry point: nothing is timed
  return loop_check_device_impl(c, st, n_frames, d_q, nq, cur_ids, d_results, d_best);
  LC_GUARD_END(c)
}

LC_API int32_t lc_merge_shards_device(lc_ctx* c, const void* d_gathered, int64_t contrib_bytes, int32_t world, int32_t n_frames, int64_t slots,
                                      int64_t rows, lc_kf_result* d_records_out, uint64_t* d_best_out, void* cuda_stream) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  const int64_t n_rec = (int64_t)n_frames * slots;
  if (world < 1 || n_frames < 0 || slots < 0 || rows < 0 || contrib_bytes % 16 != 0 || contrib_bytes < n_rec * 16 + rows * 8)
    return fail(c, LC_ERR_INVALID, "lc_merge_shards_device: bad shape (contribution %lld B for %lld records + %lld bests)",
                (long long)contrib_bytes, (long long)n_rec, (long long)rows);
  if ((n_rec > 0 && (!d_gathered || !d_records_out)) || (rows > 0 && (!d_gathered || !d_best_out)))
    return fail(c, LC_ERR_INVALID, "lc_merge_shards_device: NULL buffer");
  if (((uintptr_t)d_gathered | (uintptr_t)d_records_out) % 16 != 0) return fail(c, LC_ERR_INVALID, "lc_merge_shards_device: buffers must be 16-byte aligned");
  if (n_rec * world + rows == 0) return LC_OK;
  LC_CUDA(c, cudaSetDevice(c->cfg.device));
  StreamScope scope(c, cuda_stream);
  LC_TRY(scope.enter());
  cudaStream_t st = scope.st;
  const long long items = std::max<long long>(n_rec * world, rows);
  const int grid = (int)std::min<long long>((items + 255) / 256, 8ll * c->num_sms);
  merge_shards_kernel<<<grid, 256, 0, st>>>((const uint8_t*)d_gathered, (size_t)contrib_bytes, world, n_rec, rows, (uint4*)d_records_out, d_best_out);
  LC_CUDA(c, cudaGetLastError());
  c->launches += 1;
  return LC_OK;
  LC_GUARD_END(c)
}

// ------------------------------------------------------------------- peer-memory exchange (no collective)
namespace {
void exchange_release(lc_ctx* c) {
  auto& x = c->xchg;
  for (int r = 0; r < kMaxPeers; ++r) {
    if (x.ipc[r] && x.peer[r]) cudaIpcCloseMemHandle(x.peer[r]);
    x.peer[r] = nullptr;
    x.ipc[r] = false;
  }
  if (x.base) cudaFree(x.base);
  x = lc_ctx::Exchange{};
}
}  // namespace

LC_API int32_t lc_exchange_create(lc_ctx* c, int64_t slot_bytes, void* ipc_handle_out, void** d_base_out) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  if (slot_bytes < 16 || slot_bytes % 16 != 0) return fail(c, LC_ERR_INVALID, "slot_bytes = %lld: a positive multiple of 16 is needed", (long long)slot_bytes);
  LC_TRY(enter_host(c));
  LC_CUDA(c, cudaStreamSynchronize(c->stream));
  exchange_release(c);
  auto& x = c->xchg;
  void* p = nullptr;
  LC_CUDA(c, cudaMalloc(&p, kXchgHeaderBytes + 2 * (size_t)slot_bytes));
  x.base = (uint8_t*)p;
  x.slot_bytes = (size_t)slot_bytes;
  LC_CUDA(c, cudaMemset(x.base, 0, kXchgHeaderBytes));  // no rank has signalled anything: sequence numbers start at 1
  if (ipc_handle_out) {
    static_assert(sizeof(cudaIpcMemHandle_t) == LC_IPC_HANDLE_BYTES, "LC_IPC_HANDLE_BYTES must be the size of a CUDA IPC handle");
    cudaIpcMemHandle_t h;
    LC_CUDA(c, cudaIpcGetMemHandle(&h, x.base));
    memcpy(ipc_handle_out, &h, sizeof h);
  }
  if (d_base_out) *d_base_out = x.base;
  return LC_OK;
  LC_GUARD_END(c)
}

LC_API int32_t lc_exchange_connect(lc_ctx* c, int32_t world, int32_t rank, const void* ipc_handles, void* const* d_bases) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  auto& x = c->xchg;
  if (!x.base) return fail(c, LC_ERR_INVALID, "lc_exchange_connect before lc_exchange_create");
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) return fail(c, LC_ERR_INVALID, "world %d / rank %d (at most %d ranks)", world, rank, kMaxPeers);
  if (world > 1 && !ipc_handles && !d_bases) return fail(c, LC_ERR_INVALID, "neither IPC handles nor mapped base pointers given");
  LC_TRY(enter_host(c));
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      x.peer[r] = x.base;
    } else if (d_bases) {  // peers of the same process (their buffers are already mapped here)
      if (!d_bases[r]) return fail(c, LC_ERR_INVALID, "d_bases[%d] is NULL", r);
      x.peer[r] = (uint8_t*)d_bases[r];
    } else {
      cudaIpcMemHandle_t h;
      memcpy(&h, (const uint8_t*)ipc_handles + (size_t)r * sizeof h, sizeof h);
      void* p = nullptr;
      LC_CUDA(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      x.peer[r] = (uint8_t*)p;
      x.ipc[r] = true;
    }
  }
  x.world = world;
  x.rank = rank;
  x.connected = true;
  return LC_OK;
  LC_GUARD_END(c)
}

LC_API int32_t lc_exchange_slot(lc_ctx* c, int64_t seq, void** d_slot, int64_t* slot_bytes) {
  if (!c) return LC_ERR_INVALID;
  if (!c->xchg.base || seq < 1) return fail(c, LC_ERR_INVALID, "no exchange buffer, or seq < 1");
  if (d_slot) *d_slot = c->xchg.base + kXchgHeaderBytes + (size_t)(seq & 1) * c->xchg.slot_bytes;
  if (slot_bytes) *slot_bytes = (int64_t)c->xchg.slot_bytes;
  return LC_OK;
}

LC_API int32_t lc_exchange_close(lc_ctx* c) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  LC_TRY(enter_host(c));
  LC_CUDA(c, cudaStreamSynchronize(c->stream));
  exchange_release(c);
  return LC_OK;
  LC_GUARD_END(c)
}

LC_API int32_t lc_merge_shards_peer_device(lc_ctx* c, int64_t seq, int32_t n_frames, int64_t slots, int64_t rows, lc_kf_result* d_records_out,
                                           uint64_t* d_best_out, int32_t phase, void* cuda_stream) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  const auto& x = c->xchg;
  if (!x.connected) return fail(c, LC_ERR_INVALID, "lc_merge_shards_peer_device before lc_exchange_connect");
  const int64_t n_rec = (int64_t)n_frames * slots;
  if (seq < 1 || n_frames < 0 || slots < 0 || rows < 0 || phase < 1 || phase > 3 || (size_t)(n_rec * 16 + rows * 8) > x.slot_bytes)
    return fail(c, LC_ERR_INVALID, "lc_merge_shards_peer_device: bad shape (%lld records + %lld bests in a slot of %zu B, seq %lld, phase %d)",
                (long long)n_rec, (long long)rows, x.slot_bytes, (long long)seq, phase);
  if ((phase & 2) && ((n_rec > 0 && !d_records_out) || (rows > 0 && !d_best_out))) return fail(c, LC_ERR_INVALID, "lc_merge_shards_peer_device: NULL output");
  if ((uintptr_t)d_records_out % 16 != 0) return fail(c, LC_ERR_INVALID, "lc_merge_shards_peer_device: d_records_out must be 16-byte aligned");
  LC_CUDA(c, cudaSetDevice(c->cfg.device));
  StreamScope scope(c, cuda_stream);
  LC_TRY(scope.enter());
  PeerBases pb{};
  for (int r = 0; r < x.world; ++r) pb.base[r] = x.peer[r];
  const long long items = std::max<long long>(n_rec * x.world, rows);
  const int grid = (int)std::max<long long>(1, std::min<long long>((items + 255) / 256, c->num_sms));  // one wave: block 0 must be resident
  merge_peers_kernel<<<grid, 256, 0, scope.st>>>(pb, x.world, x.rank, (unsigned long long)seq, x.slot_bytes, n_rec, rows, (uint4*)d_records_out,
                                                 d_best_out, phase, 5000000000ull);
  LC_CUDA(c, cudaGetLastError());
  c->launches += 1;
  return LC_OK;
  LC_GUARD_END(c)
}

LC_API int32_t lc_loop_check_device(lc_ctx* c, const void* d_q, int32_t nq, int32_t cur_id, lc_kf_result* d_results,
                                    uint64_t* d_best, void* cuda_stream) {
  return lc_loop_check_batch_device(c, 1, d_q, &nq, &cur_id, d_results, d_best, cuda_stream);
}

namespace {

// first half of a host-buffer loop check: upload, kernels, D2H into the pinned result buffer - all queued, nothing waited for
int loop_check_enqueue_impl(lc_ctx* c, int32_t n_frames, const uint8_t* const* q, const int32_t* nq, const int32_t* cur_ids, bool want_best) {
  if (n_frames < 0 || (n_frames > 0 && (!q || !nq || !cur_ids))) return fail(c, LC_ERR_INVALID, "bad arguments");
  const int64_t n_kf = (int64_t)c->h_kf.size();
  int64_t rows = 0;
  for (int f = 0; f < n_frames; ++f) {
    LC_TRY(check_rows(c, "query", q[f], nq[f], LC_DESC_BYTES));
    rows += nq[f];
  }
  c->pending = lc_ctx::Pending{};
  c->pending.active = true;
  c->pending.n_res = n_kf * n_frames;
  c->pending.rows = rows;
  if (n_frames == 0) return LC_OK;
  LC_TRY(enter_host(c));
  timing_begin(c, loop_events(c, rows));
  // device side: [query rows | frame table] in d_q, [per-keyframe records | per-query bests] in d_results - one copy in
  // and one copy out when the rows fit the pinned upload buffer (a small check is a chain of latencies, not bandwidth)
  const size_t row_bytes = ((size_t)rows * LC_DESC_BYTES + 255) & ~size_t(255), table_bytes = (size_t)n_frames * sizeof(QFrame);
  const size_t res_bytes = (size_t)c->pending.n_res * sizeof(lc_kf_result), best_bytes = want_best ? (size_t)rows * 8 : 0;
  LC_TRY(ensure(c, c->d_q, row_bytes + table_bytes));
  LC_TRY(ensure(c, c->d_results, std::max<size_t>(res_bytes + best_bytes, 16)));
  LC_TRY(ensure_pin_loop(c, res_bytes + best_bytes + 16));
  lc_kf_result* d_results = (lc_kf_result*)c->d_results.p;
  uint64_t* d_best = want_best ? (uint64_t*)((uint8_t*)c->d_results.p + res_bytes) : nullptr;  // res_bytes is a multiple of 16
  std::vector<QFrame> hf;
  size_t rows_built = 0;
  LC_TRY(build_frames(c, n_frames, c->d_q.p, nq, cur_ids, d_best, hf, &rows_built));
  const QFrame* d_frames = (const QFrame*)((const uint8_t*)c->d_q.p + row_bytes);
  if (row_bytes + table_bytes <= 2 * kPinHalfBytes) {
    LC_CUDA(c, cudaEventSynchronize(c->pin_ev[0]));  // earlier uploads have drained both pinned halves
    LC_CUDA(c, cudaEventSynchronize(c->pin_ev[1]));
    size_t off = 0;
    for (int f = 0; f < n_frames; ++f) {
      memcpy(c->pin_up + off, q[f], (size_t)nq[f] * LC_DESC_BYTES);
      off += (size_t)nq[f] * LC_DESC_BYTES;
    }
    memcpy(c->pin_up + row_bytes, hf.data(), table_bytes);
    LC_CUDA(c, cudaMemcpyAsync(c->d_q.p, c->pin_up, row_bytes + table_bytes, cudaMemcpyHostToDevice, c->stream));
    LC_CUDA(c, cudaEventRecord(c->pin_ev[0], c->stream));  // lc_loop_check_begin returns with the copy in flight
    LC_CUDA(c, cudaEventRecord(c->pin_ev[1], c->stream));
  } else {
    int64_t r0 = 0;
    for (int f = 0; f < n_frames; ++f) {
      if (nq[f] > 0) LC_TRY(upload_rows(c, (uint8_t*)c->d_q.p + (size_t)r0 * LC_DESC_BYTES, q[f], nq[f], LC_DESC_BYTES));
      r0 += nq[f];
    }
    LC_CUDA(c, cudaMemcpyAsync((void*)d_frames, hf.data(), table_bytes, cudaMemcpyHostToDevice, c->stream));  // pageable: staged before the call returns
  }
  if (c->timing_on) LC_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
  LC_TRY(loop_check_run(c, c->stream, hf, d_frames, rows_built, d_results, d_best));
  if (c->timing_on) LC_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
  if (res_bytes + best_bytes) LC_CUDA(c, cudaMemcpyAsync(c->pin_loop, c->d_results.p, res_bytes + best_bytes, cudaMemcpyDeviceToHost, c->stream));
  if (c->timing_on) LC_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
  c->pending.res_bytes = res_bytes;
  c->pending.best_bytes = best_bytes;
  c->pending.queued = true;
  return LC_OK;
}

int loop_check_enqueue(lc_ctx* c, int32_t n_frames, const uint8_t* const* q, const int32_t* nq, const int32_t* cur_ids, bool want_best) {
  if (c->pending.active) return fail(c, LC_ERR_INVALID, "a loop check is pending: call lc_loop_check_end first");
  const int rc = loop_check_enqueue_impl(c, n_frames, q, nq, cur_ids, want_best);
  if (rc != LC_OK) c->pending.active = false;  // nothing of a failed check stays pending
  return rc;
}

// second half: wait for the queued work and hand the results out
int loop_check_collect(lc_ctx* c, lc_kf_result* results, int64_t results_cap, uint64_t* best_packed) {
  if (!c->pending.active) return fail(c, LC_ERR_INVALID, "no loop check is pending");
  const lc_ctx::Pending p = c->pending;
  c->pending.active = false;
  if (p.stashed) {
    c->timing = p.stashed_timing;
    c->timing_on = p.stashed_on;
  }
  if (p.queued) {
    LC_TRY(enter_host(c));
    LC_TRY(timing_end(c, true, false));
  }
  if (p.n_res > 0 && (!results || results_cap < p.n_res))
    return fail(c, LC_ERR_CAPACITY, "results_cap %lld < %lld", (long long)results_cap, (long long)p.n_res);
  if (best_packed && p.rows > 0 && !p.best_bytes && p.queued) return fail(c, LC_ERR_INVALID, "the pending check was started without want_best");
  if (p.res_bytes) memcpy(results, c->pin_loop, p.res_bytes);
  if (p.best_bytes && best_packed) memcpy(best_packed, c->pin_loop + p.res_bytes, p.best_bytes);
  return LC_OK;
}

}  // namespace

LC_API int32_t lc_loop_check_batch(lc_ctx* c, int32_t n_frames, const uint8_t* const* q, const int32_t* nq, const int32_t* cur_ids,
                                   lc_kf_result* results, int64_t results_cap, uint64_t* best_packed) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  const int64_t n_res = (int64_t)c->h_kf.size() * std::max(n_frames, 0);
  if (n_res > 0 && (!results || results_cap < n_res)) return fail(c, LC_ERR_CAPACITY, "results_cap %lld < %lld", (long long)results_cap, (long long)n_res);
  LC_TRY(loop_check_enqueue(c, n_frames, q, nq, cur_ids, best_packed != nullptr));
  return loop_check_collect(c, results, results_cap, best_packed);
  LC_GUARD_END(c)
}

LC_API int32_t lc_loop_check_begin(lc_ctx* c, const uint8_t* q, int32_t nq, size_t qs, int32_t cur_id, int32_t want_best) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  LC_TRY(check_rows(c, "query", q, nq, qs));
  std::vector<uint8_t> packed;
  const uint8_t* p = q;
  if (qs != LC_DESC_BYTES && nq > 0) {  // strided cv::Mat: make the rows contiguous
    packed.resize((size_t)nq * LC_DESC_BYTES);
    for (int32_t i = 0; i < nq; ++i) memcpy(packed.data() + (size_t)i * LC_DESC_BYTES, q + (size_t)i * qs, LC_DESC_BYTES);
    p = packed.data();
  }
  return loop_check_enqueue(c, 1, &p, &nq, &cur_id, want_best != 0);  // upload_rows has copied the rows when this returns
  LC_GUARD_END(c)
}

LC_API int32_t lc_loop_check_end(lc_ctx* c, lc_kf_result* results, int64_t results_cap, int64_t* n_results, uint64_t* best_packed) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  if (c->frame.active) return fail(c, LC_ERR_INVALID, "a frame is pending: call lc_process_frame_end");
  if (n_results) *n_results = c->pending.active ? c->pending.n_res : 0;
  return loop_check_collect(c, results, results_cap, best_packed);
  LC_GUARD_END(c)
}

LC_API int32_t lc_loop_check(lc_ctx* c, const uint8_t* q, int32_t nq, size_t qs, int32_t cur_id, lc_kf_result* results,
                             int64_t results_cap, int64_t* n_results, uint64_t* best_packed) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  LC_TRY(check_rows(c, "query", q, nq, qs));
  if (n_results) *n_results = (int64_t)c->h_kf.size();
  if (qs == LC_DESC_BYTES || nq == 0) return lc_loop_check_batch(c, 1, &q, &nq, &cur_id, results, results_cap, best_packed);
  std::vector<uint8_t> packed((size_t)nq * LC_DESC_BYTES);  // strided cv::Mat: make the rows contiguous
  for (int32_t i = 0; i < nq; ++i) memcpy(packed.data() + (size_t)i * LC_DESC_BYTES, q + (size_t)i * qs, LC_DESC_BYTES);
  const uint8_t* p = packed.data();
  return lc_loop_check_batch(c, 1, &p, &nq, &cur_id, results, results_cap, best_packed);
  LC_GUARD_END(c)
}

LC_API int32_t lc_select_candidates(const lc_ctx* c, const lc_kf_result* results, int64_t n_results, int32_t nq, int32_t cur_id,
                                    lc_loop_candidate* out, int64_t out_cap, int64_t* n_out) {
  if (!c || n_results < 0 || (n_results > 0 && !results)) return LC_ERR_INVALID;
  int64_t n = 0;
  for (int64_t k = 0; k < n_results; ++k) {
    const lc_kf_result& r = results[k];
    if (r.n_good < 0) continue;
    const int32_t denom = std::min(nq, r.n_kf);
    if (denom <= 0) continue;
    const double sim = (double)r.n_good / (double)denom;  // README.md:126
    const bool over = c->cfg.thr_inclusive ? (sim >= c->cfg.loop_threshold) : (sim > c->cfg.loop_threshold);
    if (over && r.n_good >= c->cfg.min_good) {  // README.md:123-124
      if (out && n < out_cap) out[n] = lc_loop_candidate{cur_id, r.frame_id, r.n_good, sim};
      ++n;
    }
  }
  if (n_out) *n_out = n;
  return (out && n > out_cap) ? LC_ERR_CAPACITY : LC_OK;
}

LC_API int32_t lc_detect_loops(lc_ctx* c, const uint8_t* q, int32_t nq, size_t qs, int32_t cur_id, lc_loop_candidate* out,
                               int64_t out_cap, int64_t* n_out) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  if (c->cfg.shard_world != 1) return fail(c, LC_ERR_INVALID, "lc_detect_loops needs an unsharded context; gather lc_loop_check records and call lc_select_candidates");
  std::vector<lc_kf_result> res(c->h_kf.size());
  int64_t n = 0;
  LC_TRY(lc_loop_check(c, q, nq, qs, cur_id, res.data(), (int64_t)res.size(), &n, nullptr));
  const int rc = lc_select_candidates(c, res.data(), n, nq, cur_id, out, out_cap, n_out);
  if (rc == LC_ERR_CAPACITY) return fail(c, rc, "out_cap %lld too small", (long long)out_cap);
  return rc;
  LC_GUARD_END(c)
}

// ------------------------------------------------------------------- processFrame (include/loop_closing.hpp:34)
namespace {

// first half of lc_process_frame: append + matchFeatures(previous frame, this frame) + loop check, all queued on the
// context's stream behind ONE upload of the frame's rows (they land in the store; the pair match and the loop check read
// them there), nothing waited for
int process_frame_enqueue(lc_ctx* c, int32_t frame_id, const uint8_t* desc, int32_t n, size_t stride, bool pair_event, int64_t* kf_index) {
  if (c->cfg.shard_world != 1) return fail(c, LC_ERR_INVALID, "lc_process_frame needs an unsharded context (the previous frame must be resident here)");
  if (c->pending.active || c->frame.active) return fail(c, LC_ERR_INVALID, "a loop check or a frame is pending: call lc_loop_check_end / lc_process_frame_end first");
  LC_TRY(check_rows(c, "frame", desc, n, stride));
  LC_TRY(enter_host(c));
  const bool had_prev = !c->h_kf.empty();
  const KfEntry prev = had_prev ? c->h_kf.back() : KfEntry{nullptr, 0, 0, 0};
  const int64_t n_kf_before = (int64_t)c->h_kf.size();
  const int64_t rows_total = (int64_t)n;
  timing_begin(c, loop_events(c, rows_total));
  // ---- frames_.push_back: the rows go to their place in the store (the call's only bulk upload)
  int64_t k = -1;
  LC_TRY(store_append(c, frame_id, desc, n, stride, false, &k));
  if (kf_index) *kf_index = k;
  const KfEntry cur = c->h_kf.back();
  const int64_t n_kf = n_kf_before + 1;
  c->frame = lc_ctx::FramePending{};
  c->frame.kf_index = k;
  c->frame.frame_id = frame_id;
  c->frame.nq = n;
  c->frame.n_prev = prev.n;
  // ---- the two table entries of this frame: the keyframe table's new slot and the query-frame record, by kernel argument
  LC_TRY(ensure(c, c->d_frames, sizeof(QFrame)));
  const QFrame qf{cur.rows, n, frame_id, nullptr};
  if ((size_t)n_kf * sizeof(KfEntry) <= c->d_kf.cap && c->d_kf_synced == (size_t)n_kf_before) {
    poke_tables_kernel<<<1, 32, 0, c->stream>>>((KfEntry*)c->d_kf.p + n_kf_before, cur, (QFrame*)c->d_frames.p, qf);
    c->d_kf_synced = (size_t)n_kf;
  } else {  // the table has to grow (or was never uploaded): the ordinary road, then the frame record alone
    LC_TRY(sync_kf_table(c));
    poke_tables_kernel<<<1, 32, 0, c->stream>>>(nullptr, cur, (QFrame*)c->d_frames.p, qf);
  }
  LC_CUDA(c, cudaGetLastError());
  c->launches += 1;
  c->timing.kernel_launches += 1;
  if (c->timing_on) LC_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
  // ---- matchFeatures(prev, cur): query = the previous frame (resident), train = this frame (include/loop_closing.hpp:40,
  //      direction as src/main.cpp:1154)
  if (had_prev && prev.n > 0 && n > 0) {
    const int32_t nq = prev.n;
    const size_t off_dist = (size_t)nq * 4, off_sum = off_dist + (size_t)nq * 4, off_good = off_sum + 8, out_bytes = off_good + (size_t)nq;
    c->frame.pair_direct = out_bytes <= kPairDirectOutBytes;
    LC_TRY(ensure_pinned(c, c->pin_frame, c->pin_frame_dev, c->pin_frame_cap, out_bytes));
    if (!c->frame.pair_direct) LC_TRY(ensure(c, c->d_out, out_bytes));
    uint8_t* d = c->frame.pair_direct ? c->pin_frame_dev : (uint8_t*)c->d_out.p;
    const int path_before = c->timing.path;
    LC_TRY(pair_device_impl(c, c->stream, prev.rows, nq, cur, PairOut{(int32_t*)d, (int32_t*)(d + off_dist), d + off_good, (int32_t*)(d + off_sum)}));
    c->timing.path = path_before;  // the call reports the loop check's path
    if (!c->frame.pair_direct) LC_CUDA(c, cudaMemcpyAsync(c->pin_frame, d, out_bytes, cudaMemcpyDeviceToHost, c->stream));
    if (pair_event) LC_CUDA(c, cudaEventRecord(c->pair_ev, c->stream));
    c->frame.matched = true;
  }
  // ---- detectLoops(frame_id): this frame, read where it now lives in the store, against every stored keyframe (itself
  //      included when min_loop_gap <= 0, as frames_ holds it by then: include/loop_closing.hpp:69, README.md:94-102)
  c->pending = lc_ctx::Pending{};
  c->pending.active = true;
  c->pending.n_res = n_kf;
  c->pending.rows = n;
  const size_t res_bytes = (size_t)n_kf * sizeof(lc_kf_result);
  LC_TRY(ensure(c, c->d_results, std::max<size_t>(res_bytes, 16)));
  LC_TRY(ensure_pin_loop(c, res_bytes + 16));
  std::vector<QFrame> hf(1, qf);
  LC_TRY(run_match(c, c->stream, (const KfEntry*)c->d_kf.p, (int)n_kf, c->max_n_local, (const QFrame*)c->d_frames.p, hf.data(), 1,
                   (lc_kf_result*)c->d_results.p, knobs_of(c->cfg), c->n_partial_kf == 0, nullptr, 0));
  if (c->timing_on) LC_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
  LC_CUDA(c, cudaMemcpyAsync(c->pin_loop, c->d_results.p, res_bytes, cudaMemcpyDeviceToHost, c->stream));
  if (c->timing_on) LC_CUDA(c, cudaEventRecord(c->ev[3], c->stream));
  c->pending.res_bytes = res_bytes;
  c->pending.queued = true;
  c->frame.active = true;
  return LC_OK;
}

// matchFeatures(prev, cur) of the frame in flight; wait_all: the whole stream has been waited for already
int process_frame_matches(lc_ctx* c, bool wait_all, int32_t* train_idx, int32_t* dist, uint8_t* good, int32_t cap, int32_t* n_matches,
                          int32_t* min_dist, int32_t* n_good) {
  if (!c->frame.matched) {  // first frame, or an empty side: cv::BFMatcher returns no matches
    empty_pair(n_matches, min_dist, n_good);
    return LC_OK;
  }
  const int32_t nq = c->frame.n_prev;
  if ((train_idx || dist || good) && cap < nq) return fail(c, LC_ERR_CAPACITY, "cap %d < %d rows of the previous frame", cap, nq);
  if (!wait_all && !c->frame.pair_done) LC_CUDA(c, cudaEventSynchronize(c->pair_ev));
  c->frame.pair_done = true;
  const size_t off_dist = (size_t)nq * 4, off_sum = off_dist + (size_t)nq * 4, off_good = off_sum + 8;
  const uint8_t* h = c->pin_frame;
  if (train_idx) memcpy(train_idx, h, (size_t)nq * 4);
  if (dist) memcpy(dist, h + off_dist, (size_t)nq * 4);
  if (good) memcpy(good, h + off_good, (size_t)nq);
  int32_t sum[2];
  memcpy(sum, h + off_sum, 8);
  if (n_matches) *n_matches = nq;
  if (min_dist) *min_dist = sum[0];
  if (n_good) *n_good = sum[1];
  return LC_OK;
}

int process_frame_collect(lc_ctx* c, lc_loop_candidate* out, int64_t out_cap, int64_t* n_out, lc_kf_result* results, int64_t results_cap,
                          int64_t* n_results) {
  if (!c->frame.active) return fail(c, LC_ERR_INVALID, "no frame is pending");
  const lc_ctx::FramePending fr = c->frame;
  const int64_t n_res = c->pending.n_res;
  if (n_results) *n_results = n_res;
  c->frame.active = false;
  // the records stay in the pinned buffer: select from there, copy out only when asked to
  const bool want_records = results != nullptr;
  if (want_records && results_cap < n_res) {
    c->frame.active = true;  // nothing consumed: the caller may come back with room
    return fail(c, LC_ERR_CAPACITY, "results_cap %lld < %lld", (long long)results_cap, (long long)n_res);
  }
  {  // wait; with results == NULL the collect step must not insist on a destination
    const lc_ctx::Pending p = c->pending;
    c->pending.active = false;
    if (p.stashed) {
      c->timing = p.stashed_timing;
      c->timing_on = p.stashed_on;
    }
    if (p.queued) LC_TRY(timing_end(c, true, false));
    if (want_records && p.res_bytes) memcpy(results, c->pin_loop, p.res_bytes);
  }
  c->frame.pair_done = true;  // the stream has drained
  int64_t n = 0;
  const int rc = lc_select_candidates(c, (const lc_kf_result*)c->pin_loop, n_res, fr.nq, fr.frame_id, out, out_cap, &n);
  if (n_out) *n_out = n;
  if (rc == LC_ERR_CAPACITY) return fail(c, rc, "out_cap %lld < %lld loop candidates", (long long)out_cap, (long long)n);
  return rc;
}

}  // namespace

LC_API int32_t lc_process_frame_begin(lc_ctx* c, int32_t frame_id, const uint8_t* desc, int32_t n, size_t stride, int64_t* kf_index) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  if (kf_index) *kf_index = -1;
  const int rc = process_frame_enqueue(c, frame_id, desc, n, stride, true, kf_index);
  if (rc != LC_OK && !c->frame.active) {  // nothing of a failed frame stays pending (an appended keyframe stays appended)
    if (c->pending.active && !c->pending.queued) c->pending.active = false;
    c->frame = lc_ctx::FramePending{};
  }
  return rc;
  LC_GUARD_END(c)
}

LC_API int32_t lc_process_frame_matches(lc_ctx* c, int32_t* train_idx, int32_t* dist, uint8_t* good, int32_t cap, int32_t* n_matches,
                                        int32_t* min_dist, int32_t* n_good) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  if (!c->frame.active) return fail(c, LC_ERR_INVALID, "no frame is pending");
  LC_CUDA(c, cudaSetDevice(c->cfg.device));
  return process_frame_matches(c, false, train_idx, dist, good, cap, n_matches, min_dist, n_good);
  LC_GUARD_END(c)
}

LC_API int32_t lc_process_frame_end(lc_ctx* c, lc_loop_candidate* out, int64_t out_cap, int64_t* n_out, lc_kf_result* results,
                                    int64_t results_cap, int64_t* n_results) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  LC_CUDA(c, cudaSetDevice(c->cfg.device));
  return process_frame_collect(c, out, out_cap, n_out, results, results_cap, n_results);
  LC_GUARD_END(c)
}

LC_API int32_t lc_process_frame(lc_ctx* c, int32_t frame_id, const uint8_t* desc, int32_t n, size_t stride, int32_t* train_idx, int32_t* dist,
                                uint8_t* good, int32_t cap, int32_t* n_matches, int32_t* min_dist, int32_t* n_good, lc_loop_candidate* out,
                                int64_t out_cap, int64_t* n_out, int64_t* kf_index) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  if (kf_index) *kf_index = -1;
  if (!c->h_kf.empty() && (train_idx || dist || good) && cap < c->h_kf.back().n)  // before anything is appended
    return fail(c, LC_ERR_CAPACITY, "cap %d < %d rows of the previous frame", cap, c->h_kf.back().n);
  int rc = process_frame_enqueue(c, frame_id, desc, n, stride, false, kf_index);
  if (rc != LC_OK) {
    if (c->pending.active && !c->pending.queued) c->pending.active = false;
    if (!c->frame.active) c->frame = lc_ctx::FramePending{};
    return rc;
  }
  // one wait for everything; the candidates first (that drains the stream), then the matches from their pinned buffer
  rc = process_frame_collect(c, out, out_cap, n_out, nullptr, 0, nullptr);
  const int rc2 = process_frame_matches(c, true, train_idx, dist, good, cap, n_matches, min_dist, n_good);
  return rc != LC_OK ? rc : rc2;
  LC_GUARD_END(c)
}

// ------------------------------------------------------------------- SIFT / L2 (the shipped pipeline's matcher)
namespace {

constexpr size_t kL2SlabBytes = 256u << 20;

// OpenCV SIFT rows (CV_32F integer-valued, or CV_8U) -> contiguous u8 rows in `dst`
int l2_to_u8(lc_ctx* c, const void* desc, int32_t n, size_t stride, int32_t type, std::vector<uint8_t>& dst) {
  if (n < 0 || n > LC_MAX_ROWS_PER_KF) return fail(c, LC_ERR_INVALID, "n = %d outside [0, %d]", n, LC_MAX_ROWS_PER_KF);
  if (type != LC_DESC_F32 && type != LC_DESC_U8) return fail(c, LC_ERR_INVALID, "unknown desc_type %d", type);
  if (n > 0 && !desc) return fail(c, LC_ERR_INVALID, "NULL descriptors with n = %d", n);
  const size_t row = type == LC_DESC_F32 ? LC_L2_DIMS * sizeof(float) : LC_L2_DIMS;
  if (n > 0 && stride < row) return fail(c, LC_ERR_INVALID, "row_stride %zu < %zu", stride, row);
  dst.resize((size_t)n * LC_L2_DIMS);
  for (int32_t i = 0; i < n; ++i) {
    const uint8_t* src = (const uint8_t*)desc + (size_t)i * stride;
    uint8_t* out = dst.data() + (size_t)i * LC_L2_DIMS;
    if (type == LC_DESC_U8) {
      memcpy(out, src, LC_L2_DIMS);
    } else {
      const float* f = (const float*)src;
      for (int k = 0; k < LC_L2_DIMS; ++k) {
        const float v = f[k];
        if (!(v >= 0.f && v <= 255.f) || v != (float)(int)v)
          return fail(c, LC_ERR_INVALID, "row %d, element %d = %g: SIFT descriptors must be integers in [0, 255] (cv::SIFT's output)", i, k, (double)v);
        out[k] = (uint8_t)(int)v;
      }
    }
  }
  return LC_OK;
}

// u8 rows -> device buffer (16-byte aligned rows) + their norms
int l2_stage(lc_ctx* c, DevBuf& rows, size_t norm_off, const std::vector<uint8_t>& u8, int32_t n) {
  LC_TRY(ensure(c, rows, (size_t)std::max(n, 1) * LC_L2_DIMS));
  LC_TRY(ensure(c, c->d_l2norm, norm_off + (size_t)std::max(n, 1) * 4));
  if (n > 0) {
    LC_TRY(upload_rows(c, (uint8_t*)rows.p, u8.data(), (int64_t)n * 4, LC_DESC_BYTES));  // 128 B = four 32-byte units
    l2::norms_kernel<<<(n + 7) / 8, 256, 0, c->stream>>>((const uint4*)rows.p, n, (uint32_t*)((uint8_t*)c->d_l2norm.p + norm_off));
    LC_CUDA(c, cudaGetLastError());
    c->launches += 1;
  }
  return LC_OK;
}

// match n_frames query frames (device tables) against n_kf views
int l2_run(lc_ctx* c, const l2::L2Kf* d_kfs, int n_kf, const l2::L2Frame* d_frames, const l2::L2Frame* h_frames, int n_frames,
           l2::L2Result* d_results, l2::L2Knobs kb) {
  cudaStream_t st = c->stream;
  const int n_items = n_kf * n_frames;
  if (n_items == 0) return LC_OK;
  l2::init_results_kernel<<<(n_items + 255) / 256, 256, 0, st>>>(d_kfs, n_kf, d_frames, n_frames, d_results, kb);
  LC_CUDA(c, cudaGetLastError());
  c->launches += 1;
  std::vector<l2::ABlock> blocks;
  std::vector<int> tiles;
  long long total_tiles = 0;
  for (int f = 0; f < n_frames; ++f)
    for (int q0 = 0; q0 < h_frames[f].nq; q0 += l2::kQPerCta) {
      blocks.push_back(l2::ABlock{nullptr, f, q0, std::min(l2::kQPerCta, h_frames[f].nq - q0), 0});
      tiles.push_back((std::min(l2::kQPerCta, h_frames[f].nq - q0) + l2::kQTile - 1) / l2::kQTile);
      total_tiles += tiles.back();
    }
  const int n_ab = (int)blocks.size();
  if (n_ab == 0) return LC_OK;
  std::vector<tc::CtaWork> work;
  for (int i = 0; i < n_ab; ++i) {  // CTAs per query block: proportional to its tiles, about one CTA per SM in total
    const int ns = (int)std::min<long long>(n_kf, std::max<long long>(1, (long long)c->num_sms * tiles[i] / total_tiles));
    for (int s = 0; s < ns; ++s) work.push_back(tc::CtaWork{i, s, ns});
  }
  LC_TRY(ensure(c, c->d_l2img, (size_t)n_ab * l2::kABytes));
  LC_TRY(ensure(c, c->d_l2tables, (size_t)n_ab * sizeof(l2::ABlock) + work.size() * sizeof(tc::CtaWork)));
  for (int i = 0; i < n_ab; ++i) blocks[i].img = (const uint8_t*)c->d_l2img.p + (size_t)i * l2::kABytes;
  l2::ABlock* d_blocks = (l2::ABlock*)c->d_l2tables.p;
  tc::CtaWork* d_work = (tc::CtaWork*)(d_blocks + n_ab);
  LC_CUDA(c, cudaMemcpyAsync(d_blocks, blocks.data(), (size_t)n_ab * sizeof(l2::ABlock), cudaMemcpyHostToDevice, st));
  LC_CUDA(c, cudaMemcpyAsync(d_work, work.data(), work.size() * sizeof(tc::CtaWork), cudaMemcpyHostToDevice, st));
  l2::pack_queries_kernel<<<n_ab, 512, 0, st>>>(d_blocks, d_frames, (uint8_t*)c->d_l2img.p);
  LC_CUDA(c, cudaGetLastError());
  if (!c->l2_attr_set) {
    LC_CUDA(c, cudaFuncSetAttribute(l2::l2_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, l2::kSmemAlloc));
    c->l2_attr_set = true;
  }
  {
    MatchTimer mt(c, st);
    l2::l2_match_kernel<<<(unsigned)work.size(), l2::kThreads, l2::kSmemAlloc, st>>>(d_kfs, n_kf, d_frames, d_blocks, d_work, d_results, kb);
  }
  LC_CUDA(c, cudaGetLastError());
  c->launches += 2;
  c->timing.kernel_launches += 3;
  c->timing.path = 6;
  return LC_OK;
}

}  // namespace

LC_API int32_t lc_l2_db_append(lc_ctx* c, int32_t frame_id, const void* desc, int32_t n, size_t stride, int32_t type, int64_t* kf_index) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  LC_TRY(enter_host(c));
  auto& S = c->l2s;
  const int64_t k = S.g_kf;
  const bool own = (k % c->cfg.shard_world) == c->cfg.shard_rank;
  if (n < 0 || n > LC_MAX_ROWS_PER_KF) return fail(c, LC_ERR_INVALID, "n = %d outside [0, %d]", n, LC_MAX_ROWS_PER_KF);
  if (own) {
    std::vector<uint8_t> u8;
    LC_TRY(l2_to_u8(c, desc, n, stride, type, u8));
    // rows (128 B each) followed by the norms (4 B each), 256-byte aligned, inside one slab
    const size_t need = (((size_t)n * LC_L2_DIMS + 255) & ~size_t(255)) + (((size_t)n * 4 + 255) & ~size_t(255));
    if (S.slabs.empty() || S.slab_used + need > S.slab_cap) {
      const size_t cap = std::max(kL2SlabBytes, need);
      void* p = nullptr;
      LC_CUDA(c, cudaMalloc(&p, std::max<size_t>(cap, 256)));
      S.slabs.push_back(p);
      S.slab_cap = cap;
      S.slab_used = 0;
    }
    uint8_t* rows = (uint8_t*)S.slabs.back() + S.slab_used;
    uint32_t* norms = (uint32_t*)(rows + (((size_t)n * LC_L2_DIMS + 255) & ~size_t(255)));
    S.slab_used += need;
    if (n > 0) {
      LC_TRY(upload_rows(c, rows, u8.data(), (int64_t)n * 4, LC_DESC_BYTES));
      l2::norms_kernel<<<(n + 7) / 8, 256, 0, c->stream>>>((const uint4*)rows, n, norms);
      LC_CUDA(c, cudaGetLastError());
      c->launches += 1;
    }
    S.h_kf.push_back(l2::L2Kf{(const uint4*)rows, norms, n, frame_id});
  }
  if (kf_index) *kf_index = k;
  S.g_kf += 1;
  return LC_OK;
  LC_GUARD_END(c)
}

LC_API int32_t lc_l2_db_size(const lc_ctx* c, int64_t* n_kf, int64_t* n_local) {
  if (!c) return LC_ERR_INVALID;
  if (n_kf) *n_kf = c->l2s.g_kf;
  if (n_local) *n_local = (int64_t)c->l2s.h_kf.size();
  return LC_OK;
}

LC_API int32_t lc_l2_db_clear(lc_ctx* c) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  LC_TRY(enter_host(c));
  LC_CUDA(c, cudaStreamSynchronize(c->stream));
  for (void* p : c->l2s.slabs) cudaFree(p);
  c->l2s = lc_ctx::L2Store{};
  return LC_OK;
  LC_GUARD_END(c)
}

LC_API int32_t lc_l2_match_pair(lc_ctx* c, const void* q, int32_t nq, size_t qs, const void* t, int32_t nt, size_t ts, int32_t type,
                                double ratio, int32_t* train_idx, float* distance, float* second, uint8_t* good, int32_t* n_good) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  std::vector<uint8_t> qu, tu;
  LC_TRY(l2_to_u8(c, q, nq, qs, type, qu));
  LC_TRY(l2_to_u8(c, t, nt, ts, type, tu));
  if (n_good) *n_good = 0;
  if (nq == 0) return LC_OK;
  if (nt == 0) {  // knnMatch returns empty neighbour lists: nothing survives
    for (int32_t i = 0; i < nq; ++i) {
      if (train_idx) train_idx[i] = -1;
      if (distance) distance[i] = -1.f;
      if (second) second[i] = -1.f;
      if (good) good[i] = 0;
    }
    return LC_OK;
  }
  LC_TRY(enter_host(c));
  timing_begin(c);
  const size_t tn_off = ((size_t)nq * 4 + 255) & ~size_t(255);  // d_l2norm: query norms, then train norms
  LC_TRY(ensure(c, c->d_l2norm, tn_off + (size_t)nt * 4));     // grow once: the two stagings below must not reallocate it
  LC_TRY(l2_stage(c, c->d_l2q, 0, qu, nq));
  LC_TRY(l2_stage(c, c->d_l2t, tn_off, tu, nt));
  // device outputs: idx[nq] d1[nq] d2[nq] result(16 B) good[nq]
  const size_t o_d1 = (size_t)nq * 4, o_d2 = o_d1 * 2, o_res = o_d1 * 3, o_good = o_res + 16, bytes = o_good + (size_t)nq;
  LC_TRY(ensure(c, c->d_l2out, bytes));
  LC_TRY(ensure_pin_down(c, bytes));
  uint8_t* d = (uint8_t*)c->d_l2out.p;
  struct {
    l2::L2Kf kf;
    l2::L2Frame fr;
  } h;
  h.kf = l2::L2Kf{(const uint4*)c->d_l2t.p, (const uint32_t*)((const uint8_t*)c->d_l2norm.p + tn_off), nt, 0};
  h.fr = l2::L2Frame{(const uint4*)c->d_l2q.p, (const uint32_t*)c->d_l2norm.p, nq, 0, (int32_t*)d, (int32_t*)(d + o_d1), (int32_t*)(d + o_d2), d + o_good};
  LC_TRY(ensure(c, c->d_l2frames, sizeof h));
  LC_CUDA(c, cudaMemcpyAsync(c->d_l2frames.p, &h, sizeof h, cudaMemcpyHostToDevice, c->stream));
  if (c->timing_on) LC_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
  const l2::L2Kf* d_kf = (const l2::L2Kf*)c->d_l2frames.p;
  const l2::L2Frame* d_fr = (const l2::L2Frame*)((const uint8_t*)c->d_l2frames.p + offsetof(decltype(h), fr));
  LC_TRY(l2_run(c, d_kf, 1, d_fr, &h.fr, 1, (l2::L2Result*)(d + o_res), l2::L2Knobs{-(1 << 30), 0, ratio}));
  if (c->timing_on) LC_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
  LC_CUDA(c, cudaMemcpyAsync(c->pin_down, d, bytes, cudaMemcpyDeviceToHost, c->stream));
  LC_TRY(timing_end(c));
  const uint8_t* hp = c->pin_down;
  const int32_t *hi = (const int32_t*)hp, *h1 = (const int32_t*)(hp + o_d1), *h2 = (const int32_t*)(hp + o_d2);
  for (int32_t i = 0; i < nq; ++i) {  // distances as cv::DMatch carries them: sqrtf of the exact squared distance
    if (train_idx) train_idx[i] = hi[i];
    if (distance) distance[i] = h1[i] == l2::kNoD2 ? -1.f : sqrtf((float)h1[i]);
    if (second) second[i] = h2[i] == l2::kNoD2 ? -1.f : sqrtf((float)h2[i]);
    if (good) good[i] = hp[o_good + i];
  }
  if (n_good) *n_good = ((const l2::L2Result*)(hp + o_res))->n_good;
  return LC_OK;
  LC_GUARD_END(c)
}

LC_API int32_t lc_l2_loop_check(lc_ctx* c, const void* q, int32_t nq, size_t qs, int32_t type, int32_t cur_id, int32_t loop_gap,
                                double ratio, int32_t min_rows, lc_l2_result* results, int64_t cap, int64_t* n_results) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  std::vector<uint8_t> qu;
  LC_TRY(l2_to_u8(c, q, nq, qs, type, qu));
  auto& S = c->l2s;
  const int64_t n_kf = (int64_t)S.h_kf.size();
  if (n_results) *n_results = n_kf;
  if (n_kf > 0 && (!results || cap < n_kf)) return fail(c, LC_ERR_CAPACITY, "results_cap %lld < %lld", (long long)cap, (long long)n_kf);
  if (n_kf == 0) return LC_OK;
  LC_TRY(enter_host(c));
  timing_begin(c);
  if (S.synced != S.h_kf.size()) {  // the whole table again: views are appended rarely compared with rows
    LC_TRY(ensure(c, S.d_kf, S.h_kf.size() * sizeof(l2::L2Kf)));
    LC_CUDA(c, cudaMemcpyAsync(S.d_kf.p, S.h_kf.data(), S.h_kf.size() * sizeof(l2::L2Kf), cudaMemcpyHostToDevice, c->stream));
    LC_CUDA(c, cudaStreamSynchronize(c->stream));
    S.synced = S.h_kf.size();
  }
  LC_TRY(l2_stage(c, c->d_l2q, 0, qu, nq));
  l2::L2Frame fr{(const uint4*)c->d_l2q.p, (const uint32_t*)c->d_l2norm.p, nq, cur_id, nullptr, nullptr, nullptr, nullptr};
  LC_TRY(ensure(c, c->d_l2frames, sizeof fr));
  LC_CUDA(c, cudaMemcpyAsync(c->d_l2frames.p, &fr, sizeof fr, cudaMemcpyHostToDevice, c->stream));
  LC_TRY(ensure(c, c->d_l2res, (size_t)n_kf * sizeof(l2::L2Result)));
  LC_TRY(ensure_pin_down(c, (size_t)n_kf * sizeof(l2::L2Result)));
  if (c->timing_on) LC_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
  LC_TRY(l2_run(c, (const l2::L2Kf*)S.d_kf.p, (int)n_kf, (const l2::L2Frame*)c->d_l2frames.p, &fr, 1, (l2::L2Result*)c->d_l2res.p,
                l2::L2Knobs{loop_gap, min_rows, ratio}));
  if (c->timing_on) LC_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
  LC_CUDA(c, cudaMemcpyAsync(c->pin_down, c->d_l2res.p, (size_t)n_kf * sizeof(l2::L2Result), cudaMemcpyDeviceToHost, c->stream));
  LC_TRY(timing_end(c));
  static_assert(sizeof(lc_l2_result) == sizeof(l2::L2Result), "record layouts differ");
  memcpy(results, c->pin_down, (size_t)n_kf * sizeof(lc_l2_result));
  return LC_OK;
  LC_GUARD_END(c)
}

// The shipped loop search's matching step as a whole (src/main.cpp:1375-1389): every stored view `curr` of a range against
// every stored view at least loop_gap before it.  The query side is the resident store itself - nothing is uploaded.
LC_API int32_t lc_l2_loop_search(lc_ctx* c, int64_t first_view, int32_t n_query_views, int32_t loop_gap, double ratio, int32_t min_rows,
                                 lc_l2_result* results, int64_t cap, int64_t* n_views) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  auto& S = c->l2s;
  if (c->cfg.shard_world != 1) return fail(c, LC_ERR_INVALID, "lc_l2_loop_search needs an unsharded context (the query views must be resident here)");
  const int64_t n_kf = (int64_t)S.h_kf.size();
  if (n_views) *n_views = n_kf;
  if (first_view < 0 || n_query_views < 0 || first_view + n_query_views > n_kf)
    return fail(c, LC_ERR_INVALID, "query views [%lld, %lld) outside the %lld stored views", (long long)first_view, (long long)(first_view + n_query_views), (long long)n_kf);
  const int64_t need = (int64_t)n_query_views * n_kf;
  if (need > 0 && (!results || cap < need)) return fail(c, LC_ERR_CAPACITY, "results_cap %lld < %lld", (long long)cap, (long long)need);
  if (need == 0) return LC_OK;
  if (c->pending.active) return fail(c, LC_ERR_INVALID, "a loop check or a frame is pending (its timing events are in use)");
  LC_TRY(enter_host(c));
  timing_begin(c);
  if (S.synced != S.h_kf.size()) {
    LC_TRY(ensure(c, S.d_kf, S.h_kf.size() * sizeof(l2::L2Kf)));
    LC_CUDA(c, cudaMemcpyAsync(S.d_kf.p, S.h_kf.data(), S.h_kf.size() * sizeof(l2::L2Kf), cudaMemcpyHostToDevice, c->stream));
    LC_CUDA(c, cudaStreamSynchronize(c->stream));
    S.synced = S.h_kf.size();
  }
  if (c->timing_on) LC_CUDA(c, cudaEventRecord(c->ev[1], c->stream));
  constexpr int kChunk = 64;  // query views per launch: bounds the packed query images (4 x 128 KiB per 4000-row view)
  LC_TRY(ensure(c, c->d_l2frames, (size_t)kChunk * sizeof(l2::L2Frame)));
  LC_TRY(ensure(c, c->d_l2res, (size_t)kChunk * n_kf * sizeof(l2::L2Result)));
  LC_TRY(ensure_pin_down(c, (size_t)kChunk * n_kf * sizeof(l2::L2Result)));
  static_assert(sizeof(lc_l2_result) == sizeof(l2::L2Result), "record layouts differ");
  for (int32_t f0 = 0; f0 < n_query_views; f0 += kChunk) {
    const int nf = std::min<int>(kChunk, n_query_views - f0);
    std::vector<l2::L2Frame> hf((size_t)nf);
    for (int f = 0; f < nf; ++f) {
      const l2::L2Kf& v = S.h_kf[(size_t)(first_view + f0 + f)];
      hf[f] = l2::L2Frame{v.rows, v.norms, v.n, v.frame_id, nullptr, nullptr, nullptr, nullptr};
    }
    LC_CUDA(c, cudaMemcpyAsync(c->d_l2frames.p, hf.data(), hf.size() * sizeof(l2::L2Frame), cudaMemcpyHostToDevice, c->stream));  // pageable: staged before it returns
    LC_TRY(l2_run(c, (const l2::L2Kf*)S.d_kf.p, (int)n_kf, (const l2::L2Frame*)c->d_l2frames.p, hf.data(), nf, (l2::L2Result*)c->d_l2res.p,
                  l2::L2Knobs{loop_gap, min_rows, ratio}));
    const size_t bytes = (size_t)nf * n_kf * sizeof(l2::L2Result);
    LC_CUDA(c, cudaMemcpyAsync(c->pin_down, c->d_l2res.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    LC_CUDA(c, cudaStreamSynchronize(c->stream));
    memcpy(results + (size_t)f0 * n_kf, c->pin_down, bytes);
  }
  if (c->timing_on) LC_CUDA(c, cudaEventRecord(c->ev[2], c->stream));
  LC_TRY(timing_end(c, false));
  return LC_OK;
  LC_GUARD_END(c)
}

// ------------------------------------------------------------------- the FP4 pipe's measured peak
LC_API int32_t lc_probe_mxf4_peak(lc_ctx* c, int32_t milliseconds, double* mac_per_s, double* sm_mhz) {
  if (!c) return LC_ERR_INVALID;
  LC_GUARD_BEGIN
  if (milliseconds < 1 || milliseconds > 2000) return fail(c, LC_ERR_INVALID, "milliseconds = %d outside [1, 2000]", milliseconds);
  if (c->pending.active) return fail(c, LC_ERR_INVALID, "a loop check or a frame is pending (its timing events are in use)");
  LC_TRY(enter_host(c));
  LC_TRY(ensure(c, c->d_counter, (size_t)c->num_sms * sizeof(long long)));
  LC_CUDA(c, cudaFuncSetAttribute(fx::mxf4_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fx::kPeakSmem));
  cudaEvent_t e0 = c->ev[0], e1 = c->ev[1];
  double ms_run = 0.0, cyc = 0.0;
  long long iters = 2000;  // ~0.6 ms; the second run is sized from the first
  for (int pass = 0; pass < 2; ++pass) {
    LC_CUDA(c, cudaEventRecord(e0, c->stream));
    fx::mxf4_peak_kernel<<<c->num_sms, 128, fx::kPeakSmem, c->stream>>>(iters, (long long*)c->d_counter.p);
    LC_CUDA(c, cudaGetLastError());
    LC_CUDA(c, cudaEventRecord(e1, c->stream));
    LC_CUDA(c, cudaStreamSynchronize(c->stream));
    c->launches += 1;
    float ms = 0.f;
    LC_CUDA(c, cudaEventElapsedTime(&ms, e0, e1));
    ms_run = ms;
    if (pass == 0) iters = std::max<long long>(2000, (long long)((double)iters * milliseconds / std::max(ms, 0.05f)));
  }
  std::vector<long long> cycles((size_t)c->num_sms);
  LC_CUDA(c, cudaMemcpy(cycles.data(), c->d_counter.p, cycles.size() * sizeof(long long), cudaMemcpyDeviceToHost));
  for (long long v : cycles) cyc = std::max(cyc, (double)v);
  const double macs = (double)c->num_sms * (double)iters * 4.0 * fx::kQTile * fx::kPeakN * 64.0;
  if (mac_per_s) *mac_per_s = macs / (ms_run * 1e-3);
  if (sm_mhz) *sm_mhz = cyc / (ms_run * 1e-3) / 1e6;
  return LC_OK;
  LC_GUARD_END(c)
}

// ------------------------------------------------------------------- introspection
LC_API int32_t lc_get_last_timing(const lc_ctx* c, lc_timing* out) {
  if (!c || !out) return LC_ERR_INVALID;
  *out = c->timing;
  return LC_OK;
}
LC_API int64_t lc_launch_count(const lc_ctx* c) { return c ? c->launches : 0; }
LC_API int32_t lc_synchronize(lc_ctx* c) {
  if (!c) return LC_ERR_INVALID;
  LC_TRY(enter_host(c));
  LC_CUDA(c, cudaStreamSynchronize(c->stream));
  return LC_OK;
}
