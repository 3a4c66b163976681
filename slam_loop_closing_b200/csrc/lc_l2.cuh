// lc_l2.cuh - the SHIPPED pipeline's matcher on the tensor cores: cv::BFMatcher(NORM_L2).knnMatch(k = 2) + Lowe's ratio
// test on SIFT descriptors (reference src/main.cpp:509-534; the loop search that calls it: :1375-1392).
//
// OpenCV SIFT descriptors are CV_32F rows of 128 values that are integers in [0, 255] (cv::SIFT quantises with
// saturate_cast<uchar>), so they are stored here as 128-byte u8 rows and
//     d2(q, t) = |q|^2 + |t|^2 - 2 q.t
// is an exact integer below 2^24: the dot products are a u8 x u8 -> s32 GEMM (tcgen05.mma kind::i8, unsigned operands),
// OpenCV's float accumulation is exact whatever its order, and DMatch::distance == sqrtf((float)d2) bit for bit.
//
// Per (query, keyframe) the kernel keeps the two smallest (d2, train row) in OpenCV's order (distance, then row) - the
// scan key is  -256 * dot + (|t|^2 * 128 + column)  = (d2 - |q|^2) * 128 + column, so integer min/max alone rank by
// distance and then by column - applies  best < ratio * second  exactly as main.cpp:528 does (float distances, double
// ratio) and adds the survivors to the keyframe's count; the pair API also gets the per-query neighbours.
//
// Shape: M = 128 queries x N = 128 train rows x K = 128 bytes (4 MMAs of K = 32) per tile, four 128-column s32
// accumulators (all of TMEM); up to 1024 resident queries per CTA; train rows go to shared memory as they are (a
// 128-byte row is one SWIZZLE_128B row), together with their 128 scan constants.  A streaming top-2 update costs
// 2.5 alu instructions per pair (against 0.5 for the Hamming maximum) and bound the first version of this kernel at
// about 640 cycles per tile against the MMA's 256; the epilogue is therefore a tournament: the minimum of every aligned
// group of 8 columns, the two smallest of the group minima (0.8 alu instructions per pair), and the 7 other rows of the
// single best group rescanned once per (query, view) at the end.  Even so a tile costs an epilogue thread about 140 dependent
// instructions behind a TMEM load; two warps per scheduler could not hide that (round 1: issue slots 58 % busy), so the
// epilogue is TWO sets of eight warps that drain alternate query tiles (four warps per scheduler, issue slots 73 %).  640
// threads would leave 96 registers per thread - too few for both 32-column loads of a half tile in flight - so the warp
// group of the loaders, the MMA issuer and the spare warp hands registers over with setmaxnreg (64 there, 104 in the
// epilogue; the loaders fetch a tile in two rounds of eight 16-byte loads per thread instead of one round of sixteen).
#pragma once
#include "lc_mxf4.cuh"

namespace lc {
namespace l2 {

using fx::mbar_arrive_a;
using fx::mbar_wait_a;
using fx::umma_commit_a;
using tc::elect_one;
using tc::fence_proxy_async;
using tc::tc_fence_after;
using tc::tc_fence_before;
using tc::tmem_ld32;
using tc::tmem_ld_wait;

constexpr int kDescBytes = 128;              // one SIFT descriptor
constexpr int kQTile = 128;                  // UMMA M
constexpr int kQTilesPerCta = 8;
constexpr int kQPerCta = kQTile * kQTilesPerCta;  // 1024
constexpr int kTN = 128;                     // train rows per B tile (UMMA N)
constexpr int kBStages = 3;
constexpr int kAccStages = 4;                // 4 x 128 columns = all of TMEM
constexpr int kATileBytes = kQTile * kDescBytes;   // 16 KiB
constexpr int kBTileBytes = kTN * kDescBytes;      // 16 KiB
constexpr int kABytes = kQTilesPerCta * kATileBytes;
constexpr int kOffB = kABytes;
constexpr int kOffKc = kOffB + kBStages * kBTileBytes;      // [kBStages][kTN] scan constants
constexpr int kOffXch = kOffKc + kBStages * kTN * 4;        // [tile][row] x {d1, i1, d2}: the column halves meet here
constexpr int kXchBytes = kQTilesPerCta * kQTile * 3 * 4;
constexpr int kOffBar = kOffXch + kXchBytes;
constexpr int kSmemAlloc = kOffBar + 256;    // 194,304 B
constexpr int kEpiSets = 2;                  // two sets of epilogue warps drain alternate query tiles (set = tile m & 1)
constexpr int kSetWarps = 8;                 // per set: warp w -> TMEM lane quadrant w % 4, column half w / 4
constexpr int kEpiWarps = kEpiSets * kSetWarps;
constexpr int kLoadWarps = 2;
constexpr int kMmaWarp = kEpiWarps + kLoadWarps;
constexpr int kThreads = 640;                // 20 warps (the last one idle), launched with 96 registers per thread
constexpr int kEpiRegs = 104;                // setmaxnreg: the epilogue warps take what the loader / MMA warp group gives up
constexpr int kOtherRegs = 64;
constexpr int kNone = 0x7f000000;            // scan constant of a pad column: never among the two smallest of real rows
constexpr int kNoD2 = 0x7fffffff;            // "no such neighbour"

struct L2Kf {                // one stored view of this shard
  const uint4* rows;         // n x 128 B
  const uint32_t* norms;     // n squared norms
  int32_t n;
  int32_t frame_id;
};
struct L2Frame {             // one query frame
  const uint4* q;            // nq x 128 B
  const uint32_t* norms;     // nq squared norms
  int32_t nq;
  int32_t cur_id;
  // optional per-query outputs (pair API): best row, best / second squared distance (kNoD2 if none), survives the ratio test
  int32_t* train_idx;
  int32_t* d1;
  int32_t* d2;
  uint8_t* good;
};
struct L2Result {            // one per (frame, view); mirrors lc_l2_result
  int32_t frame_id, n_kf, n_good, reserved;
};
struct L2Knobs {
  int32_t loop_gap, min_rows;
  double ratio;
};
using fx::ABlock;
using tc::CtaWork;

__device__ __forceinline__ bool view_eligible(const L2Frame& fr, const L2Kf& kf, const L2Knobs& kb) {
  // main.cpp:1380 "past <= curr - loopGap", :1384 "rows < 100 -> continue"
  return fr.cur_id - kf.frame_id >= kb.loop_gap && kf.n >= kb.min_rows && fr.nq >= kb.min_rows && kf.n > 0 && fr.nq > 0;
}
// main.cpp:521-532: two neighbours and best.distance < ratio * second.distance (float distances, double ratio)
__device__ __forceinline__ bool ratio_test(int d1, int d2, double ratio) {
  if (d1 == kNoD2 || d2 == kNoD2) return false;
  return (double)__fsqrt_rn((float)d1) < ratio * (double)__fsqrt_rn((float)d2);
}

// squared norms of u8 rows (one warp per row)
__global__ void norms_kernel(const uint4* __restrict__ rows, int n, uint32_t* __restrict__ norms) {
  const int lane = threadIdx.x & 31, r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  const uint32_t w = reinterpret_cast<const uint32_t*>(rows)[(size_t)r * 32 + lane];
  uint32_t s = __dp4a(w, w, 0u);
  s = __reduce_add_sync(0xffffffffu, s);
  if (lane == 0) norms[r] = s;
}

// A images: the query rows of every (frame, 1024-query block) in the swizzled shared-memory layout
__global__ void pack_queries_kernel(const ABlock* __restrict__ blocks, const L2Frame* __restrict__ frames, uint8_t* img_base) {
  const ABlock ab = blocks[blockIdx.x];
  const L2Frame fr = frames[ab.frame];
  uint8_t* img = img_base + (size_t)blockIdx.x * kABytes;
  for (int idx = threadIdx.x; idx < kQPerCta * 8; idx += blockDim.x) {
    const int ql = idx >> 3, ch = idx & 7, q = ab.q0 + ql;
    const uint4 v = q < fr.nq ? __ldg(fr.q + (size_t)q * 8 + ch) : make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(img + (size_t)(ql / kQTile) * kATileBytes + fx::tile_offset(ql & (kQTile - 1), ch)) = v;
  }
}

// records of the views that are not compared, zero counts for the others
__global__ void init_results_kernel(const L2Kf* __restrict__ kfs, int n_kf, const L2Frame* __restrict__ frames, int n_frames,
                                    L2Result* __restrict__ results, L2Knobs kb) {
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= n_kf * n_frames) return;
  const int f = item / n_kf, k = item - f * n_kf;
  const L2Kf kf = kfs[k];
  results[item] = L2Result{kf.frame_id, kf.n, view_eligible(frames[f], kf, kb) ? 0 : -1, 0};
}

// u8 x u8 -> s32, both K-major, M = 128, N = n (a/b format bits clear = unsigned)
__device__ __forceinline__ uint32_t instr_desc_u8(int n) { return (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kQTile >> 4) << 24); }

// two smallest of a stream of keys
struct Top2 {
  int m1, m2;
  __device__ __forceinline__ void reset() { m1 = m2 = 0x7fffffff; }
  __device__ __forceinline__ void push2(int a, int b) {  // 5 alu instructions per two keys
    const int lo = min(a, b), hi = max(a, b);
    const int t = max(m1, lo);
    m1 = min(m1, lo);
    m2 = min(m2, min(t, hi));
  }
};
__device__ __forceinline__ int group_min8(int k0, int k1, int k2, int k3, int k4, int k5, int k6, int k7) {
  return min(min(min(k0, min(k1, k2)), min(k3, k4)), min(min(k5, k6), k7));  // VIMNMX3 x 3 + VIMNMX
}
// running two nearest rows of one query inside one view, in (distance, row) order.  While the scan runs, d2 is the nearest
// row OUTSIDE the aligned group of 8 rows that holds the best one (see the tournament in the epilogue): every merge rule below
// holds for that meaning as well, because a losing side contributes its overall minimum and a winning side brings its own group.
struct Knn2 {
  int d1, i1, d2;
  __device__ __forceinline__ void reset() { d1 = d2 = kNoD2, i1 = -1; }
  // candidates (e1, j1), (e2, .) with e1 <= e2 and rows ABOVE every row seen so far: on equal distance the earlier row stays first
  __device__ __forceinline__ void merge_later(int e1, int j1, int e2) {
    if (e1 < d1) {
      d2 = e2 < d1 ? e2 : d1;
      d1 = e1;
      i1 = j1;
    } else if (e1 < d2) {
      d2 = e1;
    }
  }
  // candidates from an arbitrary other range of rows
  __device__ __forceinline__ void merge_any(int e1, int j1, int e2) {
    const bool first = e1 < d1 || (e1 == d1 && j1 < i1);
    if (first) {
      d2 = min(e2, d1);
      d1 = e1;
      i1 = j1;
    } else {
      d2 = min(d2, e1);
    }
  }
};

// smallest distance part |t|^2 - 2 q.t among the rows of the aligned group of 8 around row i1 of a view, i1 itself excluded
// (kNoD2 if there is none): 7 dot products of 128 bytes, once per (query, view)
__device__ __forceinline__ int rescan_group(const uint4* __restrict__ qrow, const L2Kf& kf, int i1) {
  uint4 q[8];
#pragma unroll
  for (int w = 0; w < 8; ++w) q[w] = __ldg(qrow + w);
  const int base = i1 & ~7;
  int best = kNoD2;
#pragma unroll 1
  for (int r = 0; r < 8; ++r) {
    const int j = base + r;
    if (j >= kf.n || j == i1) continue;
    const uint4* t = kf.rows + (size_t)j * 8;
    unsigned dot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const uint4 v = __ldg(t + w);
      dot = __dp4a(q[w].x, v.x, dot);
      dot = __dp4a(q[w].y, v.y, dot);
      dot = __dp4a(q[w].z, v.z, dot);
      dot = __dp4a(q[w].w, v.w, dot);
    }
    best = min(best, (int)__ldg(kf.norms + j) - 2 * (int)dot);
  }
  return best;
}

// One CTA per CtaWork entry: query block `ablock` against views k0 + slot, k0 + slot + n_slots, ...
__global__ void __launch_bounds__(kThreads, 1)
    l2_match_kernel(const L2Kf* __restrict__ kfs, int n_kf, const L2Frame* __restrict__ frames, const ABlock* __restrict__ blocks,
                    const CtaWork* __restrict__ work, L2Result* __restrict__ results, L2Knobs kb) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + kOffB;
  int* sKc = reinterpret_cast<int*>(smem + kOffKc);
  int* xch = reinterpret_cast<int*>(smem + kOffXch);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint32_t bar0 = smem_u32(bars);
  asm volatile("mov.u32 %0, %0;" : "+r"(bar0));  // keep the address in a register (see lc_mxf4.cuh)
  const uint32_t a_full = bar0;
  const uint32_t b_full = bar0 + 8;                       // [kBStages] one arrival per loader warp
  const uint32_t b_empty = b_full + 8 * kBStages;         // [kBStages] 1 commit + one arrival per epilogue warp (the constants)
  const uint32_t acc_full = b_empty + 8 * kBStages;       // [kAccStages] 1 commit
  const uint32_t acc_empty = acc_full + 8 * kAccStages;   // [kAccStages] one arrival per epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + 2 * kBStages + 2 * kAccStages);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const CtaWork cw = work[blockIdx.x];
  const int slot = cw.slot, n_slots = cw.n_slots;
  const ABlock ab = blocks[cw.ablock];
  const L2Frame fr = frames[ab.frame];
  const int nq_blk = min(kQPerCta, fr.nq - ab.q0);
  const int m_tiles = (nq_blk + kQTile - 1) / kQTile;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    mbar_init(bars, 1);
    for (int i = 0; i < kBStages; ++i) {
      mbar_init(bars + 1 + i, kLoadWarps);
      mbar_init(bars + 1 + kBStages + i, 1 + kEpiWarps);
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(bars + 1 + 2 * kBStages + i, 1);
      mbar_init(bars + 1 + 2 * kBStages + kAccStages + i, kSetWarps);  // the one set that drains this use of the stage
    }
    mbar_fence_init();
  }
  if (warp == kMmaWarp) tc::tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // warp groups 0-3 are the epilogue, warp group 4 holds the loaders, the MMA issuer and a spare warp
  if (warp >= kEpiWarps) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kOtherRegs));
  else asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpiRegs));

  if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      mbar_expect_tx(bars, (uint32_t)m_tiles * kATileBytes);
      for (int m = 0; m < m_tiles; ++m) bulk_g2s(sA + m * kATileBytes, ab.img + (size_t)m * kATileBytes, kATileBytes, bars);
    }
    __syncwarp();
    const bool leader = elect_one();
    mbar_wait_a(a_full, 0);
    const uint64_t adesc0 = tc::smem_desc(smem_u32(sA)), bdesc0 = tc::smem_desc(smem_u32(sB));
    const uint32_t idesc = instr_desc_u8(kTN);
    uint32_t par_bfull = 0, n_btiles = 0, st = 0, acc_par = 0;
    int bb = 0;
    for (int k = slot; k < n_kf; k += n_slots) {
      const L2Kf kf = kfs[k];
      if (!view_eligible(fr, kf, kb)) continue;
      for (int row0 = 0; row0 < kf.n; row0 += kTN) {
        const uint64_t bdesc = bdesc0 + (uint64_t)((bb * kBTileBytes) >> 4);
        mbar_wait_a(b_full + 8 * bb, (par_bfull >> bb) & 1);
        par_bfull ^= 1u << bb;
        tc_fence_after();
#pragma unroll 1
        for (int m = 0; m < m_tiles; ++m) {
          mbar_wait_a(acc_empty + 8 * st, acc_par ^ 1);
          tc_fence_after();
          if (leader) {
            const uint32_t d = tmem_base + st * kTN;
            const uint64_t adesc = adesc0 + (uint64_t)((m * kATileBytes) >> 4);
#pragma unroll
            for (int s = 0; s < 4; ++s)  // K = 128 bytes = 4 steps of 32
              tc::umma_i8(d, adesc + (uint64_t)(s * 2), bdesc + (uint64_t)(s * 2), idesc, s > 0 ? 1u : 0u);
            umma_commit_a(acc_full + 8 * st);
          }
          __syncwarp();
          if (++st == kAccStages) st = 0, acc_par ^= 1;
        }
        if (leader) umma_commit_a(b_empty + 8 * bb);
        __syncwarp();
        n_btiles += 1;
        bb = bb + 1 == kBStages ? 0 : bb + 1;
      }
    }
    // do not leave while a commit may still arrive on this CTA's barriers (a b_empty phase completes with the commit
    // and the eight epilogue arrivals of that use)
#pragma unroll
    for (int i = 0; i < kBStages; ++i) {
      const uint32_t n_i = (n_btiles + (kBStages - 1 - i)) / kBStages;  // phases completed on b_empty[i]
      if (n_i) mbar_wait_a(b_empty + 8 * i, (n_i - 1) & 1);
    }
  } else if (warp > kMmaWarp) {
    // spare warp (keeps the warp count a multiple of four, see lc_mxf4.cuh)
  } else if (warp >= kEpiWarps) {
    // ===================== B loaders (64 threads): rows as they are + the 128 scan constants =====================
    const int et = threadIdx.x - kEpiWarps * 32;
    uint32_t par_bempty = (1u << kBStages) - 1;
    int bb = 0;
    for (int k = slot; k < n_kf; k += n_slots) {
      const L2Kf kf = kfs[k];
      if (!view_eligible(fr, kf, kb)) continue;
      for (int row0 = 0; row0 < kf.n; row0 += kTN) {
        const int cnt = min(kTN, kf.n - row0);
        // 1024 16-byte chunks per tile, 64 threads: two rounds of eight loads in flight per thread (the registers of this
        // warp group have gone to the epilogue)
        uint4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int u = et + i * 64, r = u >> 3;
          v[i] = r < cnt ? __ldg(kf.rows + (size_t)(row0 + r) * 8 + (u & 7)) : make_uint4(0, 0, 0, 0);
        }
        int kc[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int c = et + i * 64;
          kc[i] = c < cnt ? (int)(__ldg(kf.norms + row0 + c) * 128u) + c : kNone + c;
        }
        mbar_wait_a(b_empty + 8 * bb, (par_bempty >> bb) & 1);
        par_bempty ^= 1u << bb;
        uint8_t* tile = sB + bb * kBTileBytes;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int u = et + i * 64;
          *reinterpret_cast<uint4*>(tile + fx::tile_offset(u >> 3, u & 7)) = v[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int u = et + (i + 8) * 64, r = u >> 3;
          v[i] = r < cnt ? __ldg(kf.rows + (size_t)(row0 + r) * 8 + (u & 7)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int u = et + (i + 8) * 64;
          *reinterpret_cast<uint4*>(tile + fx::tile_offset(u >> 3, u & 7)) = v[i];
        }
        sKc[bb * kTN + et] = kc[0];
        sKc[bb * kTN + et + 64] = kc[1];
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(b_full + 8 * bb);
        bb = bb + 1 == kBStages ? 0 : bb + 1;
      }
    }
  } else {
    // ===================== epilogue (512 threads = 2 sets x 8 warps; in a set warp w owns TMEM lanes 32 (w % 4).., column half w / 4) ==========
    // Set s drains the query tiles m with m % 2 == s; both sets walk the same tile sequence and count every tile, so each knows
    // the accumulator stage and barrier phase of its own tiles without looking at the other's.
    const int set = warp >> 3, w8 = warp & 7, quad = w8 & 3, half = w8 >> 2;
    const int row = quad * 32 + lane;
    const uint32_t tmem_rows = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 64);
    uint32_t t = 0, par_bfull = 0;
    int bb = 0;
    // |q|^2 of the queries this thread finishes: tiles m = set + 2 i with i % 2 == half
    int na[2];
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int q = ab.q0 + (set + 2 * (2 * jj + half)) * kQTile + row;
      na[jj] = q < fr.nq ? (int)__ldg(fr.norms + q) : 0;
    }
    for (int k = slot; k < n_kf; k += n_slots) {
      const L2Kf kf = kfs[k];
      if (!view_eligible(fr, kf, kb)) continue;
      Knn2 best[kQTilesPerCta / 2];  // best[i]: tile m = set + 2 i
#pragma unroll
      for (int i = 0; i < kQTilesPerCta / 2; ++i) best[i].reset();
      for (int row0 = 0; row0 < kf.n; row0 += kTN) {
        // the tile's scan constants (visible once the loaders' b_full arrival has been observed)
        mbar_wait_a(b_full + 8 * bb, (par_bfull >> bb) & 1);
        par_bfull ^= 1u << bb;
        const int4* kc4 = reinterpret_cast<const int4*>(sKc + bb * kTN + half * 64);
#pragma unroll
        for (int m = 0; m < kQTilesPerCta; ++m) {
          if (m < m_tiles) {
            if ((m & 1) == set) {
              const uint32_t st = t & (kAccStages - 1), acc_par = (t / kAccStages) & 1u;
              mbar_wait_a(acc_full + 8 * st, acc_par);
              tc_fence_after();
              int a[32], b[32];
              tmem_ld32(tmem_rows + st * kTN, a);
              tmem_ld32(tmem_rows + st * kTN + 32, b);
              tmem_ld_wait();
              fx::release_stage(acc_empty + 8 * st);
              // Tournament: the minimum of every aligned group of 8 columns (4 min3/min per 8 keys), then the two smallest
              // of the 8 group minima.  m1 is the half-tile's smallest key; m2 is the smallest key OUTSIDE m1's group - the
              // other 7 keys of that one group are looked at once per (query, view) at the very end (rescan_group), not here.
              Top2 t2;
              t2.reset();
#pragma unroll
              for (int j = 0; j < 32; j += 16) {  // key = -256 * dot + (|t|^2 * 128 + column)
                int g[4];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  const int4 k0 = kc4[(j >> 2) + 2 * u], k1 = kc4[(j >> 2) + 2 * u + 1];
                  const int4 k2 = kc4[8 + (j >> 2) + 2 * u], k3 = kc4[8 + (j >> 2) + 2 * u + 1];
                  const int c = j + 8 * u;
                  g[u] = group_min8(a[c] * -256 + k0.x, a[c + 1] * -256 + k0.y, a[c + 2] * -256 + k0.z, a[c + 3] * -256 + k0.w,
                                    a[c + 4] * -256 + k1.x, a[c + 5] * -256 + k1.y, a[c + 6] * -256 + k1.z, a[c + 7] * -256 + k1.w);
                  g[2 + u] = group_min8(b[c] * -256 + k2.x, b[c + 1] * -256 + k2.y, b[c + 2] * -256 + k2.z, b[c + 3] * -256 + k2.w,
                                        b[c + 4] * -256 + k3.x, b[c + 5] * -256 + k3.y, b[c + 6] * -256 + k3.z, b[c + 7] * -256 + k3.w);
                }
                t2.push2(g[0], g[1]);
                t2.push2(g[2], g[3]);
              }
              // tile-local (distance part, column) -> running neighbours; rows of later tiles are higher
              const int e1 = t2.m1 >> 7, j1 = row0 + (t2.m1 & 127), e2 = t2.m2 >> 7;
              best[m >> 1].merge_later(e1, j1, e2);
            }
            ++t;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_a(b_empty + 8 * bb);  // this warp is done with the stage's constants
        bb = bb + 1 == kBStages ? 0 : bb + 1;
      }
      // the two column halves of a row meet in shared memory: half h hands over the tiles it does not finish
      // (pad columns carry kNone >> 7 = 0xfe0000 + ... as distance part: far above any real one, see below)
#pragma unroll
      for (int i = 0; i < kQTilesPerCta / 2; ++i) {
        const int m = set + 2 * i;
        if ((i & 1) != half && m < m_tiles) {
          int* x = xch + (m * kQTile + row) * 3;
          x[0] = best[i].d1, x[1] = best[i].i1, x[2] = best[i].d2;
        }
      }
      if (set == 0) asm volatile("bar.sync 1, %0;" ::"n"(kSetWarps * 32) : "memory");  // one named barrier per set
      else asm volatile("bar.sync 2, %0;" ::"n"(kSetWarps * 32) : "memory");
      int cnt_good = 0;
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int m = set + 2 * (2 * jj + half), ql = m * kQTile + row;
        if (m < m_tiles && ql < nq_blk) {
          Knn2 r = half ? best[2 * jj + 1] : best[2 * jj];
          const int* x = xch + (m * kQTile + row) * 3;
          r.merge_any(x[0], x[1], x[2]);
          // distance parts -> squared distances; a pad column's part (>= kNone >> 7) means "no such row"
          const int limit = kNone >> 7;
          const int d1 = r.d1 >= limit ? kNoD2 : r.d1 + na[jj];
          int d2 = r.d2 >= limit ? kNoD2 : r.d2 + na[jj];
          // so far d2 is the nearest row OUTSIDE the best row's group of 8; the 7 rows inside it are the last candidates.
          // They can only lower d2, so a ratio test that already fails stays failed: the rescan is needed only for the
          // queries that pass so far (few, except against a view of the same place) and when the caller wants d2 itself.
          bool g = ratio_test(d1, d2, kb.ratio);
          if (d1 != kNoD2 && (g || d2 == kNoD2 || fr.d2 != nullptr)) {  // (no row outside the group at all: it decides alone)
            const int e = rescan_group(fr.q + (size_t)(ab.q0 + ql) * 8, kf, r.i1);
            if (e != kNoD2) d2 = min(d2, e + na[jj]);
            g = ratio_test(d1, d2, kb.ratio);
          }
          cnt_good += g ? 1 : 0;
          const int q = ab.q0 + ql;
          if (fr.train_idx) fr.train_idx[q] = d1 == kNoD2 ? -1 : r.i1;
          if (fr.d1) fr.d1[q] = d1;
          if (fr.d2) fr.d2[q] = d2;
          if (fr.good) fr.good[q] = g ? 1 : 0;
        }
      }
      cnt_good = __reduce_add_sync(0xffffffffu, cnt_good);
      if (lane == 0 && cnt_good) atomicAdd(&results[(size_t)ab.frame * n_kf + k].n_good, cnt_good);
      // this set's part of the exchange buffer is free again
      if (set == 0) asm volatile("bar.sync 1, %0;" ::"n"(kSetWarps * 32) : "memory");
      else asm volatile("bar.sync 2, %0;" ::"n"(kSetWarps * 32) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace l2
}  // namespace lc
