// lc_kernels.cuh - sm_100a kernels for brute-force 256-bit Hamming matching.
//
// What they compute (reference include/loop_closing.hpp:40,48,73 and
// README.md:116-126): for every query row i and every keyframe k,
//     best_k[i] = min_j popc(q_i ^ t_kj)   with the LOWEST j on ties (cv::BFMatcher's strict "<" scan),
// then per keyframe min_dist_k = min_i best_k[i], n_good_k = #{i : best_k[i] <= max(2*min_dist_k, floor)},
// and per query the global best (dist << 32 | global_row) over all compared keyframes.
//
// Mapping to the machine
//   * One thread owns RQ query descriptors in registers (8 x 32-bit words each) for the whole scan;
//     train rows are staged in shared memory by 1-D TMA bulk copies (cp.async.bulk + mbarrier,
//     3-stage ring) and read back as warp-wide broadcast LDS.128, so one 32-byte row feeds
//     RQ x blockDim pairs.  HBM traffic is the DB read once: 0.016 B/pair.
//   * A pair costs 8 LOP3 + 8 POPC (LC_KERNEL_POPC) or 16 LOP3 + 4 POPC (LC_KERNEL_CSA: a carry-save
//     tree turns 7 words into ones/twos/fours planes).  POPC issues at 16/clk/SM, LOP3 at 64/clk/SM
//     (measured, profiles/r01_pipe_rates_b200.json), so the CSA form is bound at ~3.9 pairs/clk/SM
//     instead of 2.
//   * The running best is a single u32 key (dist << 20 | row) so min() alone gives OpenCV's
//     tie-break; two rows are folded per VIMNMX3.
//   * Whole-keyframe "scan" kernel: persistent CTAs pull (query frame, keyframe) items from an atomic
//     counter; the per-keyframe reduction (min, 2xmin filter, count) is fused in the epilogue.
//   * "Split" kernel for the shortest DBs: grid over (row segment, item, query chunk), atomicMin of the key
//     into scratch, then a tiny finalize kernel.
//   * "Pair" kernel for matchFeatures: the split grid with the finalize step (2 x min filter, per-query outputs)
//     done by the CTA that finishes last - one launch per match.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lc_abi.h"

namespace lc {

constexpr int kKeyShift = 20;                    // key = dist << 20 | row
constexpr uint32_t kKeyRowMask = (1u << kKeyShift) - 1;
constexpr int kTileRows = 256;                   // rows per smem stage (8 KiB)
constexpr int kStages = 3;
constexpr int kMaxThreads = 512;

enum : int { MODE_POPC = 0, MODE_CSA = 1 };

struct KfEntry {  // one stored keyframe of this shard (device table)
  const uint4* rows;       // n x 32 B, 32-byte aligned
  int32_t n;
  int32_t frame_id;
  int64_t global_row_off;  // rows of ALL keyframes (any shard) inserted before this one
};

struct QFrame {  // one query frame
  const uint4* q;          // nq x 32 B
  int32_t nq;
  int32_t cur_id;
  uint64_t* best_packed;   // nq entries or nullptr
};

struct Knobs {
  int32_t min_loop_gap, gap_strict, filter_floor, filter_strict, min_init;
  uint32_t key_unit;  // always 1 << kKeyShift; a runtime value on purpose (see ham_key)
};

// ---------------------------------------------------------------- PTX helpers (mbarrier + bulk copy)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "LC_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LC_DONE;\n\t"
      "bra LC_WAIT;\n\t"
      "LC_DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on `bar` (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------- distance -> key
__device__ __forceinline__ void full_add(uint32_t a, uint32_t b, uint32_t c, uint32_t& sum, uint32_t& carry) {
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(sum) : "r"(a), "r"(b), "r"(c));    // a ^ b ^ c
  asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(carry) : "r"(a), "r"(b), "r"(c));  // majority
}

// key = (hamming(q, t) << 20) + row.  The adds are IMADs (fma pipe) so the alu pipe only carries
// the LOP3s and the min; row is warp-uniform.
// `unit` is 1 << kKeyShift held in a register the compiler cannot see through (it comes from a kernel
// argument): with an immediate power of two ptxas strength-reduces the last IMAD to LEA, which lands
// on the alu pipe that already carries the 16 LOP3s and is the binding pipe of the carry-save form.
template <int MODE>
__device__ __forceinline__ uint32_t ham_key(const uint32_t (&q)[8], const uint4& a, const uint4& b, uint32_t row,
                                            uint32_t unit) {
  if (MODE == MODE_POPC) {
    uint32_t d = __popc(q[0] ^ a.x) + __popc(q[1] ^ a.y) + __popc(q[2] ^ a.z) + __popc(q[3] ^ a.w) +
                 __popc(q[4] ^ b.x) + __popc(q[5] ^ b.y) + __popc(q[6] ^ b.z) + __popc(q[7] ^ b.w);
    return d * unit + row;
  } else {
    // Spelled as explicit LOP3s: left to itself nvcc re-associates the tree into ~22 LOP3 per pair.
    const uint32_t x0 = q[0] ^ a.x, x1 = q[1] ^ a.y, x2 = q[2] ^ a.z, x3 = q[3] ^ a.w;
    const uint32_t x4 = q[4] ^ b.x, x5 = q[5] ^ b.y, x6 = q[6] ^ b.z, x7 = q[7] ^ b.w;
    // four full adders (sum = LOP3 0x96, carry = LOP3 0xE8): bit-plane count of x0..x6
    uint32_t s1, c1, s2, c2, ones, c3, twos, fours;
    full_add(x0, x1, x2, s1, c1);
    full_add(x3, x4, x5, s2, c2);
    full_add(s1, s2, x6, ones, c3);
    full_add(c1, c2, c3, twos, fours);
    uint32_t key = row;
    key = __popc(fours) * (4 * unit) + key;
    key = __popc(twos) * (2 * unit) + key;
    key = __popc(ones) * unit + key;
    key = __popc(x7) * unit + key;
    return key;
  }
}

// scan `cnt` staged rows (row index of tile[0] is `row0`), updating the RQ running keys
template <int RQ, int MODE>
__device__ __forceinline__ void scan_rows(const uint4* __restrict__ tile, int cnt, uint32_t row0,
                                          const uint32_t (&q)[RQ][8], uint32_t (&best)[RQ], uint32_t unit) {
  int r = 0;
#pragma unroll 1
  for (; r + 2 <= cnt; r += 2) {
    const uint4 a0 = tile[2 * r], b0 = tile[2 * r + 1], a1 = tile[2 * r + 2], b1 = tile[2 * r + 3];
#pragma unroll
    for (int s = 0; s < RQ; ++s) {
      const uint32_t k0 = ham_key<MODE>(q[s], a0, b0, row0 + r, unit);
      const uint32_t k1 = ham_key<MODE>(q[s], a1, b1, row0 + r + 1, unit);
      best[s] = min(best[s], min(k0, k1));  // VIMNMX3.U32
    }
  }
  if (r < cnt) {
    const uint4 a0 = tile[2 * r], b0 = tile[2 * r + 1];
#pragma unroll
    for (int s = 0; s < RQ; ++s) best[s] = min(best[s], ham_key<MODE>(q[s], a0, b0, row0 + r, unit));
  }
}

template <int RQ>
__device__ __forceinline__ void load_queries(const uint4* __restrict__ qrows, int nq, int first, int stride,
                                             uint32_t (&q)[RQ][8]) {
#pragma unroll
  for (int s = 0; s < RQ; ++s) {
    const int i = first + s * stride;
    uint4 a = make_uint4(0, 0, 0, 0), b = a;
    if (i < nq) {
      a = __ldg(qrows + 2 * (size_t)i);
      b = __ldg(qrows + 2 * (size_t)i + 1);
    }
    q[s][0] = a.x; q[s][1] = a.y; q[s][2] = a.z; q[s][3] = a.w;
    q[s][4] = b.x; q[s][5] = b.y; q[s][6] = b.z; q[s][7] = b.w;
  }
}

// block-wide min / sum of one int per thread (all threads get the result)
__device__ __forceinline__ int block_min(int v, int* red) {
  v = __reduce_min_sync(0xffffffffu, v);
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  int r = red[0];
  for (int i = 1; i < nw; ++i) r = min(r, red[i]);
  return r;
}
__device__ __forceinline__ int block_sum(int v, int* red) {
  v = __reduce_add_sync(0xffffffffu, v);
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  int r = 0;
  for (int i = 0; i < nw; ++i) r += red[i];
  return r;
}

__device__ __forceinline__ bool kf_eligible(int cur_id, int frame_id, const Knobs& kb) {
  const int gap = cur_id - frame_id;
  return kb.gap_strict ? (gap > kb.min_loop_gap) : (gap >= kb.min_loop_gap);
}
__device__ __forceinline__ int filter_threshold(int min_dist, const Knobs& kb) { return max(2 * min_dist, kb.filter_floor); }
__device__ __forceinline__ bool is_good(int d, int thr, const Knobs& kb) { return kb.filter_strict ? (d < thr) : (d <= thr); }

__device__ __forceinline__ void publish_best(uint64_t* bp, int i, uint32_t key, int64_t row_off) {
  const uint64_t packed = ((uint64_t)(key >> kKeyShift) << 32) | (uint64_t)(row_off + (int64_t)(key & kKeyRowMask));
  // bp only ever decreases, so a stale read can only cause a harmless extra atomic
  if (packed < *(volatile uint64_t*)(bp + i)) atomicMin((unsigned long long*)(bp + i), (unsigned long long)packed);
}

// ---------------------------------------------------------------- whole-keyframe scan kernel
// item = frame * n_kf + k.  blockDim.x * RQ >= max nq.  One CTA computes a whole (frame, keyframe)
// match and its reduction; results[item] is written exactly once.
template <int RQ, int MODE>
__global__ void __launch_bounds__(kMaxThreads, 1)
    kf_scan_kernel(const KfEntry* __restrict__ kfs, int n_kf, const QFrame* __restrict__ frames, int n_frames,
                   lc_kf_result* __restrict__ results, int* __restrict__ work_counter, Knobs kb) {
  __shared__ __align__(128) uint4 tiles[kStages][kTileRows * 2];
  __shared__ __align__(8) uint64_t full[kStages];
  __shared__ int red[kMaxThreads / 32];
  __shared__ int s_item;

  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  const int n_items = n_kf * n_frames;
  uint32_t q[RQ][8];
  int loaded_frame = -1;
  uint32_t seq = 0;  // tiles consumed by this CTA so far: stage = seq % kStages, parity = (seq / kStages) & 1

  for (;;) {
    if (tid == 0) s_item = atomicAdd(work_counter, 1);
    __syncthreads();
    const int item = s_item;
    __syncthreads();
    if (item >= n_items) break;
    const int f = item / n_kf, k = item - f * n_kf;
    const QFrame fr = frames[f];
    const KfEntry kf = kfs[k];
    if (!kf_eligible(fr.cur_id, kf.frame_id, kb) || kf.n <= 0 || fr.nq <= 0) {
      if (tid == 0) results[item] = lc_kf_result{kf.frame_id, kf.n, -1, -1};
      continue;
    }
    if (f != loaded_frame) {
      load_queries<RQ>(fr.q, fr.nq, tid, nt, q);
      loaded_frame = f;
    }
    uint32_t best[RQ];
#pragma unroll
    for (int s = 0; s < RQ; ++s) best[s] = 0xffffffffu;

    const int n_tiles = (kf.n + kTileRows - 1) / kTileRows;
    if (tid == 0) {
      for (int t = 0; t < kStages && t < n_tiles; ++t) {
        const int cnt = min(kTileRows, kf.n - t * kTileRows);
        const uint32_t st = (seq + t) % kStages;
        mbar_expect_tx(&full[st], cnt * 32);
        bulk_g2s(tiles[st], kf.rows + (size_t)t * kTileRows * 2, cnt * 32, &full[st]);
      }
    }
    for (int t = 0; t < n_tiles; ++t, ++seq) {
      const uint32_t st = seq % kStages;
      const int cnt = min(kTileRows, kf.n - t * kTileRows);
      mbar_wait(&full[st], (seq / kStages) & 1);
      scan_rows<RQ, MODE>(tiles[st], cnt, (uint32_t)(t * kTileRows), q, best, kb.key_unit);
      __syncthreads();  // every thread is done with this stage before it is refilled
      if (tid == 0 && t + kStages < n_tiles) {
        const int t2 = t + kStages;
        const int cnt2 = min(kTileRows, kf.n - t2 * kTileRows);
        mbar_expect_tx(&full[st], cnt2 * 32);
        bulk_g2s(tiles[st], kf.rows + (size_t)t2 * kTileRows * 2, cnt2 * 32, &full[st]);
      }
    }

    // ---- fused per-keyframe reduction: min distance, 2 x min filter, count (README.md:117,126)
    int dmin = 0x7fffffff;
#pragma unroll
    for (int s = 0; s < RQ; ++s)
      if (tid + s * nt < fr.nq) dmin = min(dmin, (int)(best[s] >> kKeyShift));
    dmin = block_min(dmin, red);
    if (kb.min_init >= 0) dmin = min(dmin, kb.min_init);
    const int thr = filter_threshold(dmin, kb);
    int cnt_good = 0;
#pragma unroll
    for (int s = 0; s < RQ; ++s)
      if (tid + s * nt < fr.nq) cnt_good += is_good((int)(best[s] >> kKeyShift), thr, kb) ? 1 : 0;
    cnt_good = block_sum(cnt_good, red);
    if (tid == 0) results[item] = lc_kf_result{kf.frame_id, kf.n, dmin, cnt_good};
    if (fr.best_packed) {
#pragma unroll
      for (int s = 0; s < RQ; ++s)
        if (tid + s * nt < fr.nq) publish_best(fr.best_packed, tid + s * nt, best[s], kf.global_row_off);
    }
  }
}

// ---------------------------------------------------------------- split kernel + finalize
// grid = (row segments, frames x keyframes [k0, k0 + kc), query chunks).  Each CTA scans `seg_rows` rows of
// one keyframe for blockDim.x * RQ queries and atomically folds its keys into
// scratch[(frame * kc + k - k0) * scratch_stride + i].
template <int RQ, int MODE>
__global__ void __launch_bounds__(kMaxThreads, 1)
    split_kernel(const KfEntry* __restrict__ kfs, int k0, int kc, const QFrame* __restrict__ frames,
                 uint32_t* __restrict__ scratch, int scratch_stride, int seg_rows, Knobs kb) {
  __shared__ __align__(128) uint4 tile[kTileRows * 2];
  __shared__ __align__(8) uint64_t full;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int f = blockIdx.y / kc, k = k0 + (blockIdx.y - f * kc);
  const QFrame fr = frames[f];
  const KfEntry kf = kfs[k];
  const int row0 = blockIdx.x * seg_rows;
  const int qfirst = blockIdx.z * nt * RQ;
  if (row0 >= kf.n || qfirst >= fr.nq || !kf_eligible(fr.cur_id, kf.frame_id, kb)) return;
  const int cnt = min(seg_rows, kf.n - row0);
  if (tid == 0) {
    mbar_init(&full, 1);
    mbar_fence_init();
    mbar_expect_tx(&full, cnt * 32);
    bulk_g2s(tile, kf.rows + (size_t)row0 * 2, cnt * 32, &full);
  }
  uint32_t q[RQ][8];
  load_queries<RQ>(fr.q, fr.nq, qfirst + tid, nt, q);
  uint32_t best[RQ];
#pragma unroll
  for (int s = 0; s < RQ; ++s) best[s] = 0xffffffffu;
  __syncthreads();  // barrier init visible before anyone waits on it
  mbar_wait(&full, 0);
  scan_rows<RQ, MODE>(tile, cnt, (uint32_t)row0, q, best, kb.key_unit);
  uint32_t* dst = scratch + (size_t)blockIdx.y * scratch_stride;
#pragma unroll
  for (int s = 0; s < RQ; ++s) {
    const int i = qfirst + tid + s * nt;
    if (i < fr.nq) atomicMin(dst + i, best[s]);
  }
}

// one CTA per (frame, keyframe in [k0, k0 + kc)): reduce the scratch keys of one match; results[frame * n_kf + k].
__global__ void __launch_bounds__(256)
    finalize_kernel(const KfEntry* __restrict__ kfs, int n_kf, int k0, int kc, const QFrame* __restrict__ frames,
                    const uint32_t* __restrict__ scratch, int scratch_stride, lc_kf_result* __restrict__ results, Knobs kb) {
  __shared__ int red[8];
  const int f = blockIdx.x / kc, k = k0 + (blockIdx.x - f * kc);
  const int item = f * n_kf + k;
  const QFrame fr = frames[f];
  const KfEntry kf = kfs[k];
  const int tid = threadIdx.x;
  if (!kf_eligible(fr.cur_id, kf.frame_id, kb) || kf.n <= 0 || fr.nq <= 0) {
    if (tid == 0) results[item] = lc_kf_result{kf.frame_id, kf.n, -1, -1};
    return;
  }
  const uint32_t* keys = scratch + (size_t)blockIdx.x * scratch_stride;
  int dmin = 0x7fffffff;
  for (int i = tid; i < fr.nq; i += blockDim.x) dmin = min(dmin, (int)(keys[i] >> kKeyShift));
  dmin = block_min(dmin, red);
  if (kb.min_init >= 0) dmin = min(dmin, kb.min_init);
  const int thr = filter_threshold(dmin, kb);
  int cnt_good = 0;
  for (int i = tid; i < fr.nq; i += blockDim.x) {
    const uint32_t key = keys[i];
    cnt_good += is_good((int)(key >> kKeyShift), thr, kb) ? 1 : 0;
    if (fr.best_packed) publish_best(fr.best_packed, i, key, kf.global_row_off);
  }
  cnt_good = block_sum(cnt_good, red);
  if (tid == 0) results[item] = lc_kf_result{kf.frame_id, kf.n, dmin, cnt_good};
}

// ---------------------------------------------------------------- pair kernel (matchFeatures): one launch
// The consecutive-frame match is latency-bound (4 M pairs), so the whole of it is ONE launch with every argument passed
// by value: grid = (row segments, query chunks) folds keys into keys[nq] as the split kernel does and the smallest
// distance it saw into state[1] (min over queries of min over rows = min over everything, so no second pass is needed);
// the CTA that finishes last (ticket counter state[0] behind __threadfence) runs the finalize step - 2 x min filter,
// per-query outputs, summary - in one sweep with its loads issued back to back (each is an L2 round trip), and puts
// keys[] and state[] back to their idle values (0xffffffff, {0, 0xffffffff}) for the next call: no memset, no table upload.
constexpr int kPairTailUnroll = 4;

template <int RQ, int MODE>
__global__ void __launch_bounds__(kMaxThreads, 1)
    pair_kernel(const uint4* __restrict__ qrows, int nq, const uint4* __restrict__ trows, int nt, int seg_rows,
                uint32_t* __restrict__ keys, unsigned int* __restrict__ state, int32_t* __restrict__ train_idx,
                int32_t* __restrict__ dist, uint8_t* __restrict__ good, int32_t* __restrict__ summary, Knobs kb) {
  __shared__ __align__(128) uint4 tile[kTileRows * 2];
  __shared__ __align__(8) uint64_t full;
  __shared__ int red[kMaxThreads / 32];
  __shared__ unsigned int ticket;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int row0 = blockIdx.x * seg_rows;
  const int qfirst = blockIdx.y * nthr * RQ;
  if (row0 < nt && qfirst < nq) {  // (the grid is sized so that this holds; the ticket below is taken either way)
    const int cnt = min(seg_rows, nt - row0);
    if (tid == 0) {
      mbar_init(&full, 1);
      mbar_fence_init();
      mbar_expect_tx(&full, cnt * 32);
      bulk_g2s(tile, trows + (size_t)row0 * 2, cnt * 32, &full);
    }
    uint32_t q[RQ][8];
    load_queries<RQ>(qrows, nq, qfirst + tid, nthr, q);
    uint32_t best[RQ];
#pragma unroll
    for (int s = 0; s < RQ; ++s) best[s] = 0xffffffffu;
    __syncthreads();  // barrier init visible before anyone waits on it
    mbar_wait(&full, 0);
    scan_rows<RQ, MODE>(tile, cnt, (uint32_t)row0, q, best, kb.key_unit);
    uint32_t lo = 0xffffffffu;
#pragma unroll
    for (int s = 0; s < RQ; ++s) {
      const int i = qfirst + tid + s * nthr;
      if (i < nq) {
        atomicMin(keys + i, best[s]);
        lo = min(lo, best[s]);
      }
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    if ((tid & 31) == 0) atomicMin(state + 1, lo >> kKeyShift);
  }
  __threadfence();  // this CTA's atomics are ordered before its ticket
  __syncthreads();
  if (tid == 0) ticket = atomicAdd(state, 1u);
  __syncthreads();
  if (ticket != gridDim.x * gridDim.y - 1) return;
  __threadfence();  // last CTA: every other CTA's atomics are visible (read past L1)
  int dmin = (int)__ldcg(state + 1);
  if (kb.min_init >= 0) dmin = min(dmin, kb.min_init);
  const int thr = filter_threshold(dmin, kb);
  int cnt_good = 0;
  for (int i0 = 0; i0 < nq; i0 += nthr * kPairTailUnroll) {
    uint32_t key[kPairTailUnroll];
#pragma unroll
    for (int u = 0; u < kPairTailUnroll; ++u) {
      const int i = i0 + u * nthr + tid;
      key[u] = i < nq ? __ldcg(keys + i) : 0u;
    }
#pragma unroll
    for (int u = 0; u < kPairTailUnroll; ++u) {
      const int i = i0 + u * nthr + tid;
      if (i >= nq) continue;
      keys[i] = 0xffffffffu;
      const int d = (int)(key[u] >> kKeyShift);
      const bool g = is_good(d, thr, kb);
      cnt_good += g ? 1 : 0;
      if (train_idx) train_idx[i] = (int32_t)(key[u] & kKeyRowMask);
      if (dist) dist[i] = d;
      if (good) good[i] = g ? 1 : 0;
    }
  }
  cnt_good = block_sum(cnt_good, red);
  if (tid == 0) {
    state[0] = 0u;
    state[1] = 0xffffffffu;
    if (summary) { summary[0] = dmin; summary[1] = cnt_good; }
  }
}

// ---------------------------------------------------------------- sharded merge (multi-GPU)
// gathered = the all-gathered contributions of `world` ranks, rank r at r * contrib_bytes:
// [records: n_frames x slots x 16 B | bests: rows x u64].  Out: records[f][slot * world + r] (keyframe k lives on rank
// k % world in slot k / world) and the unsigned min of the packed bests over the ranks - one launch after one collective.
__global__ void __launch_bounds__(256)
    merge_shards_kernel(const uint8_t* __restrict__ gathered, size_t contrib_bytes, int world, long long n_rec_per_rank, long long rows,
                        uint4* __restrict__ rec_out, uint64_t* __restrict__ best_out) {
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i < n_rec_per_rank * world; i += stride) {
    const int r = (int)(i % world);
    rec_out[i] = __ldg(reinterpret_cast<const uint4*>(gathered + (size_t)r * contrib_bytes) + i / world);
  }
  const size_t best_off = (size_t)n_rec_per_rank * 16;
  for (long long i = tid; i < rows; i += stride) {
    uint64_t m = ~0ull;
    for (int r = 0; r < world; ++r) m = min(m, __ldg(reinterpret_cast<const uint64_t*>(gathered + (size_t)r * contrib_bytes + best_off) + i));
    best_out[i] = m;
  }
}

// ---------------------------------------------------------------- the same merge over PEER MEMORY (no collective)
// Every rank owns an exchange buffer [ header: flags[r] = last sequence number rank r has signalled | slot 0 | slot 1 ] that
// its peers map (CUDA IPC between processes, plain peer access inside one).  A sharded check writes its [records | bests]
// contribution into slot (seq & 1) of its OWN buffer; then this one kernel
//   1. tells every peer "my contribution `seq` is complete" - a release store at system scope into the peer's flag array
//      (the contribution was written by earlier kernels of the same stream, so it is visible before the flag),
//   2. waits - acquire loads of its own flags, which the peers write - until every rank has signalled `seq`,
//   3. reads the peers' slots directly over NVLink (volatile loads: no stale line of an earlier sequence can answer) and
//      writes the merged records (global keyframe order) and the unsigned minimum of the packed bests locally.
// No staging buffer, no second kernel.  Slot reuse needs no extra barrier: a rank overwrites slot (seq & 1) at sequence
// seq + 2, after its kernel of sequence seq + 1 has seen every peer signal seq + 1, which a peer does only after its own
// kernel of sequence seq - the last reader of that slot - has finished (stream order).
// The grid is at most one wave, so block 0 (the signaller) is always resident.  A peer that never signals makes the wait
// give up after `timeout_ns` and trap: an error at the next synchronisation instead of a hung GPU.
constexpr int kMaxPeers = 16;
constexpr size_t kXchgHeaderBytes = 1024;
struct PeerBases {
  uint8_t* base[kMaxPeers];
};
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__global__ void __launch_bounds__(256)
    merge_peers_kernel(PeerBases peers, int world, int rank, unsigned long long seq, size_t slot_bytes, long long n_rec_per_rank, long long rows,
                       uint4* __restrict__ rec_out, uint64_t* __restrict__ best_out, int phase, unsigned long long timeout_ns) {
  if ((phase & 1) && blockIdx.x == 0 && (int)threadIdx.x < world) {
    __threadfence_system();
    unsigned long long* flag = reinterpret_cast<unsigned long long*>(peers.base[threadIdx.x]) + rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(seq) : "memory");
  }
  if (!(phase & 2)) return;
  if ((int)threadIdx.x < world) {
    const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(peers.base[rank]) + threadIdx.x;
    const unsigned long long t0 = global_timer_ns();
    for (;;) {
      unsigned long long v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
      if (v >= seq) break;
      if (global_timer_ns() - t0 > timeout_ns) __trap();
      __nanosleep(100);
    }
  }
  __syncthreads();
  const size_t slot_off = kXchgHeaderBytes + (size_t)(seq & 1ull) * slot_bytes;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i < n_rec_per_rank * world; i += stride) {
    const int r = (int)(i % world);
    rec_out[i] = __ldcv(reinterpret_cast<const uint4*>(peers.base[r] + slot_off) + i / world);
  }
  const size_t best_off = slot_off + (size_t)n_rec_per_rank * 16;
  for (long long i = tid; i < rows; i += stride) {
    uint64_t m = ~0ull;
    for (int r = 0; r < world; ++r)
      m = min(m, (uint64_t)__ldcv(reinterpret_cast<const unsigned long long*>(peers.base[r] + best_off) + i));
    best_out[i] = m;
  }
}

__global__ void fill_u64_kernel(uint64_t* p, size_t n, uint64_t v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// Table entries that change with every frame travel as kernel ARGUMENTS: a launch carries them to the device for a few
// microseconds, where a host-to-device copy of 24 bytes costs a copy-engine start-up and (from pageable memory) a
// staging step.  lc_process_frame puts the new keyframe's table entry and the query-frame record in place with it.
__global__ void poke_tables_kernel(KfEntry* kf_slot, KfEntry kf, QFrame* frame_slot, QFrame frame) {
  if (threadIdx.x == 0) {
    if (kf_slot) *kf_slot = kf;
    if (frame_slot) *frame_slot = frame;
  }
}

}  // namespace lc
