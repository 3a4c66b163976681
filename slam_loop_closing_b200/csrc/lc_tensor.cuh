// lc_tensor.cuh - Hamming matching on the 5th-generation tensor cores (tcgen05.mma kind::i8, TMEM).
//
// Identity: with query bits mapped to int8 +127 / -127 and train bits to +1 / -1,
//     D[q][t] = 127 * (256 - 2 * hamming(q, t)),
// so argmin_t hamming == argmax_t D and the whole match is a 128 x N x 256 int8 GEMM tile followed by a row-wise
// max.  OpenCV's tie-break (lowest train row) is folded into the GEMM as well: a ninth K=32 step multiplies a
// constant A (1 in k = 0) with a constant B (127 - column in k = 0), so D' = 127 * dot + (127 - column); dot is even,
// consecutive dots are 254 apart and the tie term is < 128, hence max(D') = (max dot, then lowest column) and the
// epilogue is ONE VIMNMX3 per two pairs - no per-pair multiply-add.  s32 accumulation is exact; the result is
// bit-identical to the popc kernels (same key, same tie-break).
//
// Why this shape:
//   * tcgen05.mma has no 1-bit kind and mma.sync b1 is emulated on sm_100a; kind::i8 at M=128 runs
//     8192 MAC/clk/SM = 32 descriptor pairs/clk/SM, 16x the POPC pipe.
//   * A (queries) is RESIDENT: one CTA owns up to 512 queries, expanded once by expand_queries_kernel
//     into the exact 128B-swizzled K-major smem image (4 tiles x 32 KiB) and pulled in with 1-D TMA bulk
//     copies.  Streaming A per tile would need ~32 B/clk/SM of L2 bandwidth chip-wide.
//   * B (keyframe rows) is expanded IN the kernel from the 32-byte packed rows (so the store stays
//     32 B/row in HBM): 4 expander warps turn 128 rows into a 32 KiB swizzled tile, 3 stages.
//   * D lives in TMEM (4 stages x 128 columns = all of it); 8 epilogue warps (lane quadrant x column half) read
//     it with tcgen05.ld.32x32b and fold 64 columns into one key per row: 1 IMAD + 1/2 VIMNMX3 per pair.  Four
//     stages matter: tcgen05.mma issue back-pressures the issuing thread (the queue is shallow), so any wait of the
//     MMA warp on a not-yet-drained accumulator is a bubble in the tensor pipe.
//   * two issuer warps (one elected lane each, alternating A tiles) issue the MMAs; mbarriers + tcgen05.commit
//     connect the three roles.
// Output: one u32 key (dist << 20 | row) per (frame, keyframe, query) into the scratch that
// finalize_kernel reduces - the same contract as split_kernel.
#pragma once
#include "lc_kernels.cuh"

namespace lc {
namespace tc {

constexpr int kQTile = 128;                 // UMMA M
constexpr int kQTilesPerCta = 4;            // resident A tiles
constexpr int kQPerCta = kQTile * kQTilesPerCta;
constexpr int kTN = 128;                    // train rows per B tile (UMMA N)
constexpr int kBStages = 3;                 // expanded B tiles in flight
constexpr int kAccStages = 4;               // TMEM accumulator stages: 4 x 128 columns = all 512
constexpr int kATileBytes = kQTile * 256;   // 32 KiB: 2 K-blocks x 128 rows x 128 B
constexpr int kBTileBytes = kTN * 256;      // 32 KiB
constexpr int kABytes = kQTilesPerCta * kATileBytes;
// Constant operands of the tie-break step (K-major, no swizzle: core matrix = 8 rows x 16 B).  Only k = 0 matters:
//   A_tie: every row is [1, 0, ..., 0] -> ONE row group (2 K-chunks x 128 B) re-read for all 16 groups (SBO = 0);
//   B_tie: row c holds 127 - c at k = 0 -> 16 groups x 128 B for K-chunk 0; K-chunk 1 aliases chunk 0 (LBO = 0), it is
//          multiplied by A's zeros.
constexpr int kTieABytes = 256;
constexpr int kTieBBytes = 16 * 128;
constexpr int kBarBytes = 256;
constexpr int kOffB = kABytes;
constexpr int kOffTieA = kOffB + kBStages * kBTileBytes;
constexpr int kOffTieB = kOffTieA + kTieABytes;
constexpr int kOffBar = kOffTieB + kTieBBytes;
constexpr int kSmemAlloc = kOffBar + kBarBytes;  // 231,936 B of the 232,448 B a CTA may have
constexpr int kEpiWarps = 8;                // warp w: TMEM lane quadrant w % 4, column half w / 4
constexpr int kExpWarps = 4;
constexpr int kMmaWarp = kEpiWarps + kExpWarps;
constexpr int kMmaWarps = 2;                // issuer w handles A tiles m = w, w + 2 (tcgen05.mma costs its issuing
                                            // thread ~48 clk, so one thread cannot feed 64-clk N=128 MMAs plus barriers)
constexpr int kThreads = (kMmaWarp + kMmaWarps) * 32;  // 448: warps 0-7 epilogue, 8-11 expanders, 12-13 MMA issuers
constexpr int kColsPerEpiWarp = kTN / 2;    // 64
constexpr int kAccStride = kTN;             // TMEM columns per accumulator stage
constexpr int kTmemCols = 512;
constexpr int kWordsPerExpThread = kTN * 8 / (kExpWarps * 32);  // 8

struct ABlock {          // one resident query block: 512 queries of one frame
  const uint8_t* img;    // expanded smem image (kABytes), written by expand_queries_kernel
  int32_t frame;
  int32_t q0;
  int32_t nq;            // queries of the block (the FP4 path sizes its blocks per launch; this kernel's are always full)
  int32_t reserved;
};
struct CtaWork {         // CTA -> (query block, its share of the keyframes): keyframes k0 + slot, + n_slots, ...
  int32_t ablock;
  int32_t slot;
  int32_t n_slots;
};

// ---------------------------------------------------------------- bit -> int8 expansion
// nibble (4 descriptor bits) -> 4 bytes: bit clear -> +MAG, bit set -> -MAG  (MAG = 1 for trains, 127 for queries)
template <int MAG>
__device__ __forceinline__ uint32_t expand_nibble(uint32_t n) {
  const uint32_t s = (n * 0x00204081u) & 0x01010101u;  // bit i -> byte i (0/1); the partial products never overlap
  constexpr uint32_t pos = (uint32_t)MAG & 0xffu, neg = (uint32_t)(-MAG) & 0xffu;
  return s * (neg - pos) + pos * 0x01010101u;          // 0 -> pos, 1 -> neg, no carries between bytes
}
// 16 bits -> one 16-byte swizzle chunk
template <int MAG>
__device__ __forceinline__ uint4 expand_half(uint32_t h) {
  return make_uint4(expand_nibble<MAG>(h & 15u), expand_nibble<MAG>((h >> 4) & 15u), expand_nibble<MAG>((h >> 8) & 15u),
                    expand_nibble<MAG>((h >> 12) & 15u));
}
// byte offset of (row r, 32-bit descriptor word w) inside a K-major SWIZZLE_128B tile of `rows` rows:
// [K-block = w / 4][row][128 B], 16-byte chunk index XOR (row & 7)  (Swizzle<3,4,3>)
__device__ __forceinline__ void tile_offsets(int rows, int r, int w, uint32_t& off0, uint32_t& off1) {
  const uint32_t base = (uint32_t)(w >> 2) * (uint32_t)rows * 128u + (uint32_t)r * 128u;
  const uint32_t j0 = (uint32_t)(w & 3) * 2u, x = (uint32_t)(r & 7);
  off0 = base + ((j0 ^ x) << 4);
  off1 = base + (((j0 + 1u) ^ x) << 4);
}

// one launch per loop check: every (frame, 512-query block) gets its A image in global memory
__global__ void expand_queries_kernel(const ABlock* __restrict__ blocks, const QFrame* __restrict__ frames, uint8_t* img_base) {
  const ABlock ab = blocks[blockIdx.x];
  const QFrame fr = frames[ab.frame];
  uint8_t* img = img_base + (size_t)blockIdx.x * kABytes;
  for (int idx = threadIdx.x; idx < kQPerCta * 8; idx += blockDim.x) {
    const int ql = idx >> 3, w = idx & 7, q = ab.q0 + ql;
    const uint32_t bits = q < fr.nq ? __ldg(reinterpret_cast<const uint32_t*>(fr.q) + (size_t)q * 8 + w) : 0u;
    uint32_t o0, o1;
    tile_offsets(kQTile, ql & (kQTile - 1), w, o0, o1);
    uint8_t* tile = img + (size_t)(ql / kQTile) * kATileBytes;
    *reinterpret_cast<uint4*>(tile + o0) = expand_half<127>(bits & 0xffffu);
    *reinterpret_cast<uint4*>(tile + o1) = expand_half<127>(bits >> 16);
  }
}

// ---------------------------------------------------------------- PTX: tcgen05 / mbarrier helpers
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug must trap, not hang the GPU
__device__ __forceinline__ void mbar_wait_guard(uint64_t* bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try(bar, parity); ++spins)
    if (spins > (1u << 24)) __trap();
}
__device__ __forceinline__ bool elect_one() {  // one lane of the (converged) warp
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> s32, issued by ONE thread
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// mbarrier arrives once every tcgen05 op issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (version 1): 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// K-major, no swizzle ("interleaved"): core matrix = 8 rows x 16 B contiguous; next 16-byte K chunk at +lbo, next
// 8-row group at +sbo
__device__ __forceinline__ uint64_t smem_desc_noswz(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46);
}
// s8 x s8 -> s32, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t instr_desc(int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kQTile >> 4) << 24);
}
// this warp's 32 TMEM lanes x 32 consecutive columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Fold this thread's row of columns [c0, c0 + 64) of one accumulator stage (cnt valid columns in the stage) into
// max_c D'[c].  Both TMEM loads are issued back to back and waited for once; four independent running maxima keep
// the VIMNMX3 dependency chains short.
__device__ __forceinline__ int fold_accumulator(uint32_t taddr, int c0, int cnt) {
  constexpr int NEG = (int)0x80000000;
  if (c0 >= cnt) return NEG;
  int v[kColsPerEpiWarp];
  int (&v0)[32] = *reinterpret_cast<int (*)[32]>(&v[0]);
  int (&v1)[32] = *reinterpret_cast<int (*)[32]>(&v[32]);
  tmem_ld32(taddr + c0, v0);
  if (c0 + 32 < cnt) tmem_ld32(taddr + c0 + 32, v1);
  tmem_ld_wait();
  int k0 = NEG, k1 = NEG, k2 = NEG, k3 = NEG;
  if (c0 + kColsPerEpiWarp <= cnt) {
#pragma unroll
    for (int j = 0; j < kColsPerEpiWarp; j += 8) {
      k0 = max(k0, max(v[j + 0], v[j + 1]));
      k1 = max(k1, max(v[j + 2], v[j + 3]));
      k2 = max(k2, max(v[j + 4], v[j + 5]));
      k3 = max(k3, max(v[j + 6], v[j + 7]));
    }
  } else {
    const int lim = cnt - c0;
#pragma unroll
    for (int j = 0; j < kColsPerEpiWarp; ++j)
      if (j < lim) k0 = max(k0, v[j]);
  }
  return max(max(k0, k1), max(k2, k3));
}

// ---------------------------------------------------------------- MMA issuer warp
// The whole warp runs the (uniform) control flow so descriptors live in uniform registers; one elected lane issues
// tcgen05.mma / tcgen05.commit.  (Issuing from inside an `if (lane == 0)` region makes ptxas wrap every MMA in an
// ELECT/R2UR.BROADCAST waterfall: ~100 clk per instruction.)  Accumulator stage == A tile index m, so every stage
// is used exactly once per B tile and one parity bit serves all of them.
template <int WI>
__device__ __forceinline__ void mma_issuer(const KfEntry* __restrict__ kfs, int k0, int kc, int slot, int n_slots, const QFrame& fr,
                                           const Knobs& kb, int m_tiles, uint32_t tmem_base, uint32_t sA_addr,
                                           uint64_t* a_full, uint64_t* b_full, uint64_t* b_empty, uint64_t* acc_full,
                                           uint64_t* acc_empty) {
  const bool leader = elect_one();
  mbar_wait_guard(a_full, 0);
  uint32_t par_bfull = 0, par_acc = 1, n_btiles = 0;
  int bb = 0;
  const uint64_t adesc0 = smem_desc(sA_addr), bdesc0 = smem_desc(sA_addr + kOffB);
  const uint64_t adesc_tie = smem_desc_noswz(sA_addr + kOffTieA, 128, 0), bdesc_tie = smem_desc_noswz(sA_addr + kOffTieB, 0, 128);
  for (int k = k0 + slot; k < k0 + kc; k += n_slots) {
    const KfEntry kf = kfs[k];
    if (!kf_eligible(fr.cur_id, kf.frame_id, kb) || kf.n <= 0) continue;
    for (int row0 = 0; row0 < kf.n; row0 += kTN) {
      const int cnt = min(kTN, kf.n - row0);
      const uint32_t idesc = instr_desc((cnt + 15) & ~15);
      const uint64_t bdesc = bdesc0 + (uint64_t)((bb * kBTileBytes) >> 4);
      mbar_wait_guard(&b_full[bb], (par_bfull >> bb) & 1);
      par_bfull ^= 1u << bb;
      tc_fence_after();
#pragma unroll
      for (int m = WI; m < kQTilesPerCta; m += kMmaWarps) {
        if (m < m_tiles) {
          mbar_wait_guard(&acc_empty[m], par_acc);
          tc_fence_after();
          if (leader) {
            const uint32_t d = tmem_base + (uint32_t)m * kAccStride;
            const uint64_t adesc = adesc0 + (uint64_t)((m * kATileBytes) >> 4);
#pragma unroll
            for (int s = 0; s < 8; ++s) {  // K = 256 = 2 K-blocks x 4 steps of 32 bytes
              const uint32_t ka = (uint32_t)(s >> 2) * (kQTile * 128) + (uint32_t)(s & 3) * 32;
              const uint32_t kbo = (uint32_t)(s >> 2) * (kTN * 128) + (uint32_t)(s & 3) * 32;
              umma_i8(d, adesc + (ka >> 4), bdesc + (kbo >> 4), idesc, s > 0 ? 1u : 0u);
            }
            umma_i8(d, adesc_tie, bdesc_tie, idesc, 1u);  // + (127 - column): the tie-break
            umma_commit(&acc_full[m]);
          }
        }
      }
      if (leader) umma_commit(&b_empty[bb]);  // arrives when this warp's MMAs reading the B stage have completed
      par_acc ^= 1;
      n_btiles += 1;
      bb = bb + 1 == kBStages ? 0 : bb + 1;
    }
  }
  // do not leave while a commit may still arrive on this CTA's barriers
#pragma unroll
  for (int i = 0; i < kBStages; ++i) {
    const uint32_t n_i = (n_btiles + (kBStages - 1 - i)) / kBStages;  // phases completed on b_empty[i]
    if (n_i) mbar_wait_guard(&b_empty[i], (n_i - 1) & 1);
  }
}

// ---------------------------------------------------------------- the kernel
// One CTA per CtaWork entry: query block `ablock` against keyframes k0 + slot, k0 + slot + n_slots, ...
// scratch[(frame * kc + (k - k0)) * stride + q] = dist << 20 | row  for every compared (frame, k, q).
__global__ void __launch_bounds__(kThreads, 1)
    tensor_match_kernel(const KfEntry* __restrict__ kfs, int k0, int kc, const QFrame* __restrict__ frames,
                        const ABlock* __restrict__ blocks, const CtaWork* __restrict__ work, uint32_t* __restrict__ scratch,
                        int scratch_stride, Knobs kb) {
  extern __shared__ __align__(1024) uint8_t smem[];  // SWIZZLE_128B operands need 1024-byte alignment
  uint8_t* sA = smem;
  uint8_t* sB = smem + kOffB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* a_full = bars;                          // 1 arrival + tx bytes
  uint64_t* b_full = bars + 1;                      // [kBStages] one arrival per expander thread
  uint64_t* b_empty = b_full + kBStages;            // [kBStages] one commit per MMA issuer warp
  uint64_t* acc_full = b_empty + kBStages;          // [kAccStages] 1 commit
  uint64_t* acc_empty = acc_full + kAccStages;      // [kAccStages] one arrival per epilogue thread
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kAccStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const CtaWork cw = work[blockIdx.x];
  const int slot = cw.slot, n_slots = cw.n_slots;
  const ABlock ab = blocks[cw.ablock];
  const QFrame fr = frames[ab.frame];
  const int nq_blk = min(kQPerCta, fr.nq - ab.q0);            // queries of this block (> 0 by construction)
  const int m_tiles = (nq_blk + kQTile - 1) / kQTile;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    mbar_init(a_full, 1);
    for (int i = 0; i < kBStages; ++i) {
      mbar_init(&b_full[i], kExpWarps * 32);
      mbar_init(&b_empty[i], kMmaWarps);
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], kEpiWarps * 32);
    }
    mbar_fence_init();
  }
  // constant operands of the tie-break MMA step: A = 1 at k = 0 (every row), B = 127 - column at k = 0
  for (int i = threadIdx.x; i < (kTieABytes + kTieBBytes) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + kOffTieA)[i] = 0u;
  __syncthreads();
  if (threadIdx.x < 8) smem[kOffTieA + threadIdx.x * 16] = 1;  // K-chunk 0, row r at +16 r, k = 0 is its first byte
  if (threadIdx.x < kTN) smem[kOffTieB + (threadIdx.x >> 3) * 128 + (threadIdx.x & 7) * 16] = (uint8_t)(kTN - 1 - threadIdx.x);
  fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= kMmaWarp) {
    // ===================== MMA issuers (2 warps) =====================
    if (warp == kMmaWarp && lane == 0) {
      mbar_expect_tx(a_full, (uint32_t)m_tiles * kATileBytes);
      for (int m = 0; m < m_tiles; ++m) bulk_g2s(sA + m * kATileBytes, ab.img + (size_t)m * kATileBytes, kATileBytes, a_full);
    }
    __syncwarp();
    if (warp == kMmaWarp)
      mma_issuer<0>(kfs, k0, kc, slot, n_slots, fr, kb, m_tiles, tmem_base, smem_u32(sA), a_full, b_full, b_empty, acc_full, acc_empty);
    else
      mma_issuer<1>(kfs, k0, kc, slot, n_slots, fr, kb, m_tiles, tmem_base, smem_u32(sA), a_full, b_full, b_empty, acc_full, acc_empty);
  } else if (warp >= kEpiWarps) {
    // ===================== B expanders (128 threads) =====================
    const int et = threadIdx.x - kEpiWarps * 32;
    uint32_t par_bempty = (1u << kBStages) - 1;
    int bb = 0;
    for (int k = k0 + slot; k < k0 + kc; k += n_slots) {
      const KfEntry kf = kfs[k];
      if (!kf_eligible(fr.cur_id, kf.frame_id, kb) || kf.n <= 0) continue;
      const uint32_t* words = reinterpret_cast<const uint32_t*>(kf.rows);
      for (int row0 = 0; row0 < kf.n; row0 += kTN) {
        const int nwords = min(kTN, kf.n - row0) * 8;
        uint32_t bits[kWordsPerExpThread];
#pragma unroll
        for (int i = 0; i < kWordsPerExpThread; ++i) {  // all loads in flight before the buffer is even free
          const int idx = et + i * (kExpWarps * 32);
          bits[i] = idx < nwords ? __ldg(words + (size_t)row0 * 8 + idx) : 0u;
        }
        mbar_wait_guard(&b_empty[bb], (par_bempty >> bb) & 1);
        par_bempty ^= 1u << bb;
        uint8_t* tile = sB + bb * kBTileBytes;
#pragma unroll
        for (int i = 0; i < kWordsPerExpThread; ++i) {
          const int idx = et + i * (kExpWarps * 32);
          if (idx < nwords) {
            uint32_t o0, o1;
            tile_offsets(kTN, idx >> 3, idx & 7, o0, o1);
            *reinterpret_cast<uint4*>(tile + o0) = expand_half<1>(bits[i] & 0xffffu);
            *reinterpret_cast<uint4*>(tile + o1) = expand_half<1>(bits[i] >> 16);
          }
        }
        fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
        mbar_arrive(&b_full[bb]);
        bb = bb + 1 == kBStages ? 0 : bb + 1;
      }
    }
  } else {
    // ===================== epilogue (256 threads: warp w owns TMEM lanes 32(w%4).. and columns 96(w/4)..) ====
    const int quad = warp & 3, half = warp >> 2;
    const int row = quad * 32 + lane;  // row inside an A tile
    const int c0 = half * kColsPerEpiWarp;
    uint32_t par_accfull = 0;  // accumulator stage == m: every stage completes once per B tile
    for (int k = k0 + slot; k < k0 + kc; k += n_slots) {
      const KfEntry kf = kfs[k];
      if (!kf_eligible(fr.cur_id, kf.frame_id, kb) || kf.n <= 0) continue;
      uint32_t best[kQTilesPerCta];  // dist << 20 | row, smaller is better
#pragma unroll
      for (int m = 0; m < kQTilesPerCta; ++m) best[m] = 0xffffffffu;
      for (int row0 = 0; row0 < kf.n; row0 += kTN) {
        const int cnt = min(kTN, kf.n - row0);
#pragma unroll
        for (int m = 0; m < kQTilesPerCta; ++m) {
          if (m < m_tiles) {
            mbar_wait_guard(&acc_full[m], par_accfull);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)m * kAccStride;
            const int dmax = fold_accumulator(taddr, c0, cnt);  // the TMEM loads have completed when this returns
            tc_fence_before();
            mbar_arrive(&acc_empty[m]);  // accumulator stage drained by this thread
            if (dmax != (int)0x80000000) {
              // D' = 127 * dot + tie, dot = 256 - 2 * ham, tie = 127 - column  =>  D' + 127 * 256 = 254 * (256 - ham) + tie
              const uint32_t u = (uint32_t)(dmax + 127 * 256);
              const uint32_t h = u / 254u, tie = u - h * 254u;
              best[m] = min(best[m], ((256u - h) << kKeyShift) | (uint32_t)(row0 + (kTN - 1) - (int)tie));
            }
          }
        }
        par_accfull ^= 1;
      }
      // the two column halves of a row meet in scratch
      uint32_t* dst = scratch + ((size_t)ab.frame * kc + (k - k0)) * scratch_stride + ab.q0;
#pragma unroll
      for (int m = 0; m < kQTilesPerCta; ++m) {
        const int q = m * kQTile + row;
        if (m < m_tiles && q < nq_blk && best[m] != 0xffffffffu) atomicMin(dst + q, best[m]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace tc
}  // namespace lc
