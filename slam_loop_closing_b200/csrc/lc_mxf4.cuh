// lc_mxf4.cuh - Hamming matching on the FP4 tensor pipe (tcgen05.mma kind::mxf4.block_scale, TMEM).
//
// Identity: query bits mapped to E2M1 -1.0 / +1.0 and train bits to 0.0 / +1.0 (all UE8M0 block scales = 1.0) give
//     D[q][t] = 2 popc(q & t) - popc(t) = popc(q) - hamming(q, t)     (f32 accumulation of <= 256 terms of 0, +/-1: exact),
// so min_t hamming = popc(q) - max_t D  (the +1 / -1 form of both sides, D = 256 - 2 hamming, is kept behind a macro).  kind::mxf4 runs 16384 MAC/clk/SM = 64 descriptor pairs/clk/SM, twice
// kind::i8 and 32x the POPC pipe (measured: tools/umma_fp4.cu, profiles/r01_umma_fp4_probe.txt).
//
// What this kernel produces is the per-(keyframe, query) MINIMUM DISTANCE only.  That is all the loop check needs per
// keyframe (min_dist, 2 x min filter, n_good: README.md:117-126); the train index is needed once per query, for the
// global best over the whole DB, and is recovered afterwards by resolve_best_kernel on the single winning keyframe
// of each query (lowest keyframe, then lowest row: OpenCV's order).  Dropping the index from the hot loop removes
// the tie-break MMA step of lc_tensor.cuh (1/5 of the MMA time here) and leaves a plain FMNMX3 fold as epilogue.
//
// Shape (all measured on B200):
//   * M = 128 queries x N = 224 train rows x K = 64 per MMA, 4 MMAs per tile = 448 cycles.  One mxf4 MMA costs N / 2
//     cycles for N >= 128 with both operands in shared memory; two 224-column f32 accumulators + the scale-factor
//     cells fill TMEM's 512 columns, and the long tile amortises the per-tile hand-offs best (three 128-column stages
//     were measured: not faster).
//   * A (queries) is RESIDENT: one CTA owns up to 1024 queries = 8 tiles x 16 KiB (a packed E2M1 row of 256 values is
//     128 B = one SWIZZLE_128B row), written once by expand_queries_fp4_kernel and pulled in with bulk copies.
//   * B (keyframe rows) is expanded IN the kernel from the 32-byte rows in HBM (the store stays 32 B/row): 4 expander
//     warps, 9 ALU ops per 32 descriptor bits (PRMT as a 2-bit -> byte lookup), 3 stages of 28 KiB.  A B tile is
//     reused by all resident A tiles.
//   * 8 epilogue warps (TMEM lane quadrant x column half) read the accumulator with tcgen05.ld (700 B/clk/SM measured
//     next to full-rate MMAs; 256 B/clk/SM are needed) and fold 112 columns into one running maximum per query with
//     FMNMX3: 448 warp instructions per tile on the half-rate alu pipe = half of a tile's cycles per SM sub-partition,
//     a quarter of the SM's issue slots.  With everything else the kernel is issue-bound as much as tensor-bound, so
//     per-tile overhead instructions are kept minimal; the drain (wake-up + load latency + fold, ~490 cycles per tile)
//     is what keeps the tensor pipe at 91 % (DESIGN.md 4c lists the variants that were measured and did not help).
//   * one MMA issuer warp; mbarriers + tcgen05.commit connect the three roles.
// Output: scratch[(frame * kc + (k - k0)) * stride + q] = min distance (u16), plain stores, every compared (frame, k, q)
// written exactly once - no memset, no atomics; per query, (min distance << 32 | first keyframe at that distance) is
// kept in registers over all keyframes of the CTA and folded into best_packed once at the end.
#pragma once
#include "lc_tensor.cuh"

namespace lc {
namespace fx {

using tc::elect_one;
using tc::fence_proxy_async;
using tc::tc_fence_after;
using tc::tc_fence_before;
using tc::tmem_ld32;
using tc::tmem_ld_wait;

constexpr int kQTile = 128;                  // UMMA M
constexpr int kQTilesPerCta = 8;             // resident A tiles
constexpr int kQPerCta = kQTile * kQTilesPerCta;  // 1024
constexpr int kTN = 224;                     // train rows per B tile (UMMA N of a full tile)
constexpr int kBStages = 3;
constexpr int kAccStages = 2;
constexpr int kAccStride = 256;              // TMEM columns between accumulator stages
constexpr int kSfCol = 224;                  // TMEM columns [224, 256): scale-factor cells, all UE8M0 1.0
constexpr int kATileBytes = kQTile * 128;    // 16 KiB
constexpr int kBTileBytes = kTN * 128;       // 28 KiB
constexpr int kABytes = kQTilesPerCta * kATileBytes;  // 128 KiB
constexpr int kAImageBytes = kABytes + kQPerCta * 2;  // a block's image in global memory: the A tiles, then popc(q) of its queries (u16)
constexpr int kOffB = kABytes;
constexpr int kOffXch = kOffB + kBStages * kBTileBytes;
constexpr int kXchBytes = 2 * kQTilesPerCta * kQTile * 4;  // [parity][tile][row] f32: the column halves meet here
constexpr int kOffBar = kOffXch + kXchBytes;
constexpr int kBarBytes = 256;
constexpr int kSmemAlloc = kOffBar + kBarBytes;  // 225,536 B of the 232,448 B a CTA may have
constexpr int kEpiWarps = 8;                 // warp w: TMEM lane quadrant w % 4, column half w / 4
constexpr int kExpWarps = 4;
constexpr int kMmaWarp = kEpiWarps + kExpWarps;
constexpr int kThreads = (kMmaWarp + 1) * 32;  // 416: warps 0-7 epilogue, 8-11 expanders, 12 MMA issuer
constexpr int kTmemCols = 512;
constexpr int kPairsPerExpThread = kTN * 4 / (kExpWarps * 32);  // 7 descriptor word pairs per thread and B tile

struct ABlock {          // one resident query block: up to 1024 queries of one frame
  const uint8_t* img;    // expanded smem image (kABytes), written by expand_queries_fp4_kernel
  int32_t frame;
  int32_t q0;
  int32_t nq;            // queries of this block: 128 x (1, 2, 4 or 8 tiles), less at the end of the frame.  Small stores get
                         // small blocks, so that (block, keyframe) units outnumber the SMs (run_match in lc_abi.cu)
  int32_t reserved;
};
using tc::CtaWork;

// ---------------------------------------------------------------- bit -> E2M1 expansion
// 32 descriptor bits -> 32 E2M1 nibbles (one 16-byte swizzle chunk).  PRMT is the lookup: a selector nibble holding two
// descriptor bits picks the byte (two E2M1 values) from a four-byte pool.  The resulting order of the 256 values along K
// is a fixed permutation of the bit order - irrelevant for a dot product as long as queries and train rows use the same
// one.  Two encodings (LC_FX_ZERO_ONE_TRAIN):
//   1 (default)  queries: bit clear -> -1.0 (0xA), set -> +1.0 (0x2);  train rows: clear -> 0.0 (0x0), set -> +1.0 (0x2):
//                D = 2 popc(q & t) - popc(t) = popc(q) - hamming(q, t), so min_t hamming = popc(q) - max_t D.  popc(q) is a
//                per-QUERY constant added once per (keyframe, query) result, and half of all products are exact zeros -
//                the multipliers and adders they would toggle are what a power-capped B200 spends its clock on.
//   0            both sides +1.0 / -1.0:  D = 256 - 2 hamming.
// LC_FX_DRAIN_EXPERIMENT: timing experiments of tools/gpu_r02c.sh (wrong results), 0 in every shipped build
#ifndef LC_FX_DRAIN_EXPERIMENT
#define LC_FX_DRAIN_EXPERIMENT 0
#endif
#ifndef LC_FX_ZERO_ONE_TRAIN
#define LC_FX_ZERO_ONE_TRAIN 1
#endif
template <bool QUERY>
__device__ __forceinline__ uint4 expand_word_fp4(uint32_t x) {
#if LC_FX_ZERO_ONE_TRAIN
  constexpr uint32_t pool = QUERY ? 0x222AA2AAu : 0x22200200u;  // bytes for the bit pairs 00, 01, 10, 11 (low nibble = low bit)
#else
  constexpr uint32_t pool = 0xAAA22A22u;
#endif
  const uint32_t lo = x & 0x33333333u, hi = (x >> 2) & 0x33333333u;
  return make_uint4(__byte_perm(pool, 0u, lo & 0xffffu), __byte_perm(pool, 0u, lo >> 16), __byte_perm(pool, 0u, hi & 0xffffu),
                    __byte_perm(pool, 0u, hi >> 16));
}
// byte offset of (row r, 32-bit descriptor word w) inside a K-major SWIZZLE_128B tile: row r at r * 128 B,
// 16-byte chunk index XOR (row & 7)  (Swizzle<3,4,3>)
__device__ __forceinline__ uint32_t tile_offset(int r, int w) { return (uint32_t)r * 128u + (((uint32_t)w ^ ((uint32_t)r & 7u)) << 4); }

// ---------------------------------------------------------------- programmatic dependent launch
// The kernels of a loop check (expand -> match -> finish) are launched with programmatic stream serialization: a kernel may
// be scheduled as soon as every CTA of its predecessor has called pdl_trigger() (all of them do so at their start), runs
// the part of its prologue that touches no global memory, and blocks in pdl_wait() until the predecessor grid has completed
// and its writes are visible.  That takes the launch latency of two kernels out of a small check; with an ordinary launch
// both calls are no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// one launch per loop check, grid (query blocks, 8): every 128-query tile of every (frame, 1024-query block) gets its A
// image in global memory.  fill_best: the per-query global best of those queries is also put to its idle value here
// (all ones), which saves the loop check a launch of its own for that.
constexpr int kExpandThreads = 512;
__global__ void __launch_bounds__(kExpandThreads)
    expand_queries_fp4_kernel(const ABlock* __restrict__ blocks, const QFrame* __restrict__ frames, uint8_t* img_base, int fill_best) {
  pdl_trigger();  // (its own inputs were uploaded by copies that completed before this kernel started)
  const ABlock ab = blocks[blockIdx.x];
  const QFrame fr = frames[ab.frame];
  const int tile = blockIdx.y;
  if (tile * kQTile >= ab.nq) return;  // a small block: its last tiles do not exist
  uint8_t* block_img = img_base + (size_t)blockIdx.x * kAImageBytes;
  uint8_t* img = block_img + (size_t)tile * kATileBytes;
  uint16_t* qpop = reinterpret_cast<uint16_t*>(block_img + kABytes) + tile * kQTile;
  for (int idx = threadIdx.x; idx < kQTile * 8; idx += kExpandThreads) {
    const int r = idx >> 3, w = idx & 7, q = ab.q0 + tile * kQTile + r;
    // rows past the end of the block: E2M1 0.0 everywhere - their products are exact zeros that cost the tensor pipe no
    // switching power (their accumulator rows are never read)
    const bool valid = tile * kQTile + r < ab.nq;
    const uint32_t bits = valid ? __ldg(reinterpret_cast<const uint32_t*>(fr.q) + (size_t)q * 8 + w) : 0u;
    *reinterpret_cast<uint4*>(img + tile_offset(r, w)) = valid ? expand_word_fp4<true>(bits) : make_uint4(0u, 0u, 0u, 0u);
    // popc(q) of the row: its eight words sit in eight neighbouring lanes
    int pc = __popc(bits);
    pc += __shfl_xor_sync(0xffffffffu, pc, 1);
    pc += __shfl_xor_sync(0xffffffffu, pc, 2);
    pc += __shfl_xor_sync(0xffffffffu, pc, 4);
    if (w == 0) qpop[r] = (uint16_t)pc;
  }
  if (fill_best && fr.best_packed && threadIdx.x < kQTile) {
    const int q = ab.q0 + tile * kQTile + threadIdx.x;
    if (tile * kQTile + (int)threadIdx.x < ab.nq) fr.best_packed[q] = ~0ull;
  }
}

// ---------------------------------------------------------------- PTX
// Barriers are addressed by their 32-bit shared-window address, computed once: converting a generic pointer costs
// five instructions per use, and this kernel is bound by instruction issue as much as by the tensor pipe.
__device__ __forceinline__ bool mbar_try_a(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug must trap, not hang the GPU
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_a(bar, parity); ++spins)
    if (spins > (1u << 24)) __trap();
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
// mbarrier arrives once every tcgen05 op issued so far by this thread has completed
__device__ __forceinline__ void umma_commit_a(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, E2M1 x E2M1 -> f32, block-32 UE8M0 scales read from TMEM; issued by ONE thread
__device__ __forceinline__ void umma_mxf4(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                          uint32_t tmem_sfa, uint32_t tmem_sfb) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
      : "memory");
}
// block-scaled instruction descriptor: A, B = E2M1 (format 1), K-major, scales UE8M0 (bit 23), scale ids 0, M = 128, N = n
__device__ __forceinline__ uint32_t instr_desc_mxf4(int n) {
  return (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (1u << 23) | ((uint32_t)(kQTile >> 4) << 24);
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory"); }

// running maxima of one thread: four independent FMNMX3 chains
struct Max4 {
  float m0, m1, m2, m3;
  __device__ __forceinline__ void reset() { m0 = m1 = m2 = m3 = -1024.f; }
  template <int N>
  __device__ __forceinline__ void fold(const int (&v)[N]) {
#pragma unroll
    for (int j = 0; j < N; j += 8) {
      m0 = fmaxf(m0, fmaxf(__int_as_float(v[j + 0]), __int_as_float(v[j + 1])));
      m1 = fmaxf(m1, fmaxf(__int_as_float(v[j + 2]), __int_as_float(v[j + 3])));
      m2 = fmaxf(m2, fmaxf(__int_as_float(v[j + 4]), __int_as_float(v[j + 5])));
      m3 = fmaxf(m3, fmaxf(__int_as_float(v[j + 6]), __int_as_float(v[j + 7])));
    }
  }
  __device__ __forceinline__ float get() const { return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)); }
};

// this warp has drained its part of an accumulator stage: ONE mbarrier arrival per warp (an arrival is a shared-memory
// atomic on one address; one per thread and tile would add up)
__device__ __forceinline__ void release_stage(uint32_t acc_empty) {
  tc_fence_before();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive_a(acc_empty);
}

#if LC_FX_DRAIN_EXPERIMENT == 4
// timing experiment only: the same cells through the 16x256b access shape (16 lanes x 8 columns per request)
template <int N>
__device__ __forceinline__ void tmem_ld_16x256(uint32_t taddr, int (&v)[4 * N]);
template <>
__device__ __forceinline__ void tmem_ld_16x256<8>(uint32_t taddr, int (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld_16x256<4>(uint32_t taddr, int (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld_16x256<2>(uint32_t taddr, int (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
#endif

// Max over this thread's row of columns [taddr, taddr + ncols) of one accumulator stage (ncols: multiple of 16, <= 112),
// then release the stage.  All loads of the half are issued at once (ptxas tracks each LDTM with its own scoreboard,
// so the FMNMX3s of the first 32 columns run while the rest is still in flight) and the stage goes back to the MMA
// warp as soon as the last load has landed - before the rest of the fold.
__device__ __forceinline__ float drain_accumulator(uint32_t taddr, int ncols, uint32_t acc_empty) {
  Max4 mx;
  mx.reset();
#if LC_FX_DRAIN_EXPERIMENT == 6    // WRONG RESULTS, timing only: no TMEM load at all, 56 FMNMX3 on eight opaque values
  if (ncols == kTN / 2) {
    release_stage(acc_empty);
    int s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("mov.b32 %0, %1;" : "=r"(s[i]) : "r"(taddr + i));
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = i + 1; j < 8; ++j) {  // 28 unordered pairs, twice (operands swapped): 56 distinct instructions
          const float x = fmaxf(__int_as_float(r ? s[j] : s[i]), __int_as_float(r ? s[i] : s[j]) + (r ? 1.f : 0.f) * 0.f);
          const int c = (i + j) & 3;
          if (c == 0) mx.m0 = fmaxf(mx.m0, x);
          if (c == 1) mx.m1 = fmaxf(mx.m1, x);
          if (c == 2) mx.m2 = fmaxf(mx.m2, x);
          if (c == 3) mx.m3 = fmaxf(mx.m3, x);
        }
  } else
#endif
#if LC_FX_DRAIN_EXPERIMENT == 4    // WRONG RESULTS, timing only: every cell loaded at once through the 16x256b shape, no fold
  if (ncols == kTN / 2) {
    int a0[32], a1[32], b0[16], b1[16], c0[8], c1[8];
    const uint32_t hi = taddr + (16u << 16);
    tmem_ld_16x256<8>(taddr, a0);
    tmem_ld_16x256<8>(hi, a1);
    tmem_ld_16x256<4>(taddr + 64, b0);
    tmem_ld_16x256<4>(hi + 64, b1);
    tmem_ld_16x256<2>(taddr + 96, c0);
    tmem_ld_16x256<2>(hi + 96, c1);
    tmem_ld_wait();
    release_stage(acc_empty);
    mx.m0 = fmaxf(__int_as_float(a0[0]), fmaxf(__int_as_float(a1[0]), __int_as_float(b0[0])));
    mx.m1 = fmaxf(__int_as_float(b1[0]), fmaxf(__int_as_float(c0[0]), __int_as_float(c1[0])));
  } else
#endif
  if (ncols == kTN / 2) {  // the full tile: 112 = 32 + 32 + 32 + 16 columns
    int a[32], b[32], c[32], d[16];
    tmem_ld32(taddr, a);
    tmem_ld32(taddr + 32, b);
    tmem_ld32(taddr + 64, c);
    tmem_ld16(taddr + 96, d);
    tmem_ld_wait();
    release_stage(acc_empty);
#if LC_FX_DRAIN_EXPERIMENT == 1    // WRONG RESULTS, timing only: all TMEM loads, (almost) no fold
    mx.m0 = fmaxf(__int_as_float(a[0]), __int_as_float(b[0]));
    mx.m1 = fmaxf(__int_as_float(c[0]), __int_as_float(d[0]));
#elif LC_FX_DRAIN_EXPERIMENT == 2  // WRONG RESULTS, timing only: 64 of the 112 columns loaded and folded
    mx.fold(a);
    mx.fold(b);
#elif LC_FX_DRAIN_EXPERIMENT == 3  // WRONG RESULTS, timing only: all loads, 64 of the 112 columns folded
    mx.fold(a);
    mx.fold(b);
    mx.m1 = fmaxf(mx.m1, fmaxf(__int_as_float(c[0]), __int_as_float(d[0])));
#else
    mx.fold(a);
    mx.fold(b);
    mx.fold(c);
    mx.fold(d);
#endif
  } else {  // the last tile of a keyframe: 16 columns at a time
#pragma unroll 1
    for (int col = 0; col < ncols; col += 16) {
      int v[16];
      tmem_ld16(taddr + col, v);
      tmem_ld_wait();
      mx.fold(v);
    }
    release_stage(acc_empty);
  }
  return mx.get();
}

// columns of a tile with `cnt` valid rows: rounded up to 32 (pad rows are copies of the last valid row, so they never
// change a maximum) - every column half is then a whole number of 16-column loads
__host__ __device__ __forceinline__ int tile_cols(int cnt) { return (cnt + 31) & ~31; }
// does a keyframe of n rows consist of full-width tiles only?
inline bool full_tiles_only(int n) { return n > 0 && tile_cols(n - (n - 1) / kTN * kTN) == kTN; }

// ---------------------------------------------------------------- the kernel
// One CTA per CtaWork entry: query block `ablock` against keyframes k0 + slot, k0 + slot + n_slots, ...
// FULL: every B tile of every keyframe of the launch spans all 224 columns (true for the usual 2000-row keyframes: 8 x 224 +
// 208 -> 224 with pad rows); the host knows, and the per-tile column arithmetic and its branches drop out of an
// instruction stream that is the kernel's bottleneck.
template <bool FULL>
__global__ void __launch_bounds__(kThreads, 1)
    mxf4_match_kernel(const KfEntry* __restrict__ kfs, int k0, int kc, const QFrame* __restrict__ frames,
                      const ABlock* __restrict__ blocks, const CtaWork* __restrict__ work, uint16_t* __restrict__ scratch,
                      int scratch_stride, Knobs kb) {
  extern __shared__ __align__(1024) uint8_t smem[];  // SWIZZLE_128B operands need 1024-byte alignment
  uint8_t* sA = smem;
  uint8_t* sB = smem + kOffB;
  float* xch = reinterpret_cast<float*>(smem + kOffXch);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  // barrier i of a group lives at group base + 8 * i (shared-window addresses)
  // (laundered through an opaque asm: otherwise ptxas rematerialises the address - S2R SR_CgaCtaId + 4 more - on the
  // critical path of every stage release when registers get tight)
  uint32_t bar0 = smem_u32(bars);
  asm volatile("mov.u32 %0, %0;" : "+r"(bar0));
  const uint32_t a_full = bar0;                               // 1 arrival + tx bytes
  const uint32_t b_full = bar0 + 8;                           // [kBStages] one arrival per expander warp
  const uint32_t b_empty = b_full + 8 * kBStages;             // [kBStages] 1 commit
  const uint32_t acc_full = b_empty + 8 * kBStages;           // [kAccStages] 1 commit
  const uint32_t acc_empty = acc_full + 8 * kAccStages;       // [kAccStages] one arrival per epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + 2 * kBStages + 2 * kAccStages);

  pdl_trigger();
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform by construction
  const int lane = threadIdx.x & 31;
  const CtaWork cw = work[blockIdx.x];
  const int slot = cw.slot, n_slots = cw.n_slots;
  const ABlock ab = blocks[cw.ablock];
  const QFrame fr = frames[ab.frame];
  const int nq_blk = ab.nq;                                   // queries of this block (> 0 by construction)
  const int m_tiles = (nq_blk + kQTile - 1) / kQTile;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    mbar_init(bars, 1);
    for (int i = 0; i < kBStages; ++i) {
      mbar_init(bars + 1 + i, kExpWarps);
      mbar_init(bars + 1 + kBStages + i, 1);
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(bars + 1 + 2 * kBStages + i, 1);
      mbar_init(bars + 1 + 2 * kBStages + kAccStages + i, kEpiWarps);
    }
    mbar_fence_init();
  }
  if (warp == kMmaWarp) tc::tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp < 4) {  // scale factors: every cell of columns [224, 256) = four UE8M0 1.0 bytes, all 128 lanes
    for (int c = 0; c < 32; ++c) tmem_st1(tmem_base + ((uint32_t)(warp * 32) << 16) + kSfCol + c, 0x7f7f7f7fu);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // everything above touched shared and tensor memory only; from here on the query images (expand kernel), the distance
  // scratch (the previous batch's finalize kernel still reads it) and the per-query bests are in play
  pdl_wait();

  if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    // The whole warp runs the (uniform) control flow so descriptors live in uniform registers; one elected lane
    // issues tcgen05.mma / tcgen05.commit.
    if (lane == 0) {
      mbar_expect_tx(bars, (uint32_t)m_tiles * kATileBytes);
      for (int m = 0; m < m_tiles; ++m) bulk_g2s(sA + m * kATileBytes, ab.img + (size_t)m * kATileBytes, kATileBytes, bars);
    }
    __syncwarp();
    const bool leader = elect_one();
    mbar_wait_a(a_full, 0);
    const uint64_t adesc0 = tc::smem_desc(smem_u32(sA)), bdesc0 = tc::smem_desc(smem_u32(sB));
    const uint32_t sfa = tmem_base + kSfCol, sfb = tmem_base + kSfCol + 8;
    uint32_t par_bfull = 0, n_btiles = 0, t = 0;  // t: accumulator tiles issued so far (stage t & 1)
    int bb = 0;
    for (int k = k0 + slot; k < k0 + kc; k += n_slots) {
      const KfEntry kf = kfs[k];
      if (!kf_eligible(fr.cur_id, kf.frame_id, kb) || kf.n <= 0) continue;
      for (int row0 = 0; row0 < kf.n; row0 += kTN) {
        const uint32_t idesc = instr_desc_mxf4(FULL ? kTN : tile_cols(min(kTN, kf.n - row0)));
        const uint64_t bdesc = bdesc0 + (uint64_t)((bb * kBTileBytes) >> 4);
        mbar_wait_a(b_full + 8 * bb, (par_bfull >> bb) & 1);
        par_bfull ^= 1u << bb;
        tc_fence_after();
#pragma unroll 1
        for (int m = 0; m < m_tiles; ++m, ++t) {
          const uint32_t st = t & 1u, par = (t >> 1) & 1u;
          mbar_wait_a(acc_empty + 8 * st, par ^ 1);
          tc_fence_after();
          if (leader) {
            const uint32_t d = tmem_base + st * kAccStride;
            const uint64_t adesc = adesc0 + (uint64_t)((m * kATileBytes) >> 4);
#pragma unroll
            for (int s = 0; s < 4; ++s)  // K = 256 = 4 steps of 64 E2M1 values = 32 bytes along the 128-byte row
              umma_mxf4(d, adesc + (uint64_t)(s * 2), bdesc + (uint64_t)(s * 2), idesc, s > 0 ? 1u : 0u, sfa, sfb);
            umma_commit_a(acc_full + 8 * st);
          }
          __syncwarp();
        }
        if (leader) umma_commit_a(b_empty + 8 * bb);  // arrives when the MMAs reading this B stage have completed
        __syncwarp();
        n_btiles += 1;
        bb = bb + 1 == kBStages ? 0 : bb + 1;
      }
    }
    // do not leave while a commit may still arrive on this CTA's barriers
#pragma unroll
    for (int i = 0; i < kBStages; ++i) {
      const uint32_t n_i = (n_btiles + (kBStages - 1 - i)) / kBStages;  // phases completed on b_empty[i]
      if (n_i) mbar_wait_a(b_empty + 8 * i, (n_i - 1) & 1);
    }
  } else if (warp >= kEpiWarps) {
    // ===================== B expanders (128 threads) =====================
    // Thread et owns the descriptor word pair (2 wp, 2 wp + 1) of rows r0 + 32 i, i = 0..6: row & 7 is the same for all
    // seven, so the swizzled shared-memory offsets and the global offsets differ by constants.
    const int et = threadIdx.x - kEpiWarps * 32;
    const int r0 = et >> 2, wp = et & 3;
    const uint32_t off_lo = tile_offset(r0, 2 * wp), off_hi = tile_offset(r0, 2 * wp + 1);
    uint32_t par_bempty = (1u << kBStages) - 1;
    int bb = 0;
    for (int k = k0 + slot; k < k0 + kc; k += n_slots) {
      const KfEntry kf = kfs[k];
      if (!kf_eligible(fr.cur_id, kf.frame_id, kb) || kf.n <= 0) continue;
      const uint2* pairs = reinterpret_cast<const uint2*>(kf.rows);  // 4 word pairs per 32-byte row
      for (int row0 = 0; row0 < kf.n; row0 += kTN) {
        const int cnt = min(kTN, kf.n - row0);
        uint2 bits[kPairsPerExpThread];
        if (cnt == kTN) {  // all loads in flight before the buffer is even free
#pragma unroll
          for (int i = 0; i < kPairsPerExpThread; ++i) bits[i] = __ldg(pairs + (size_t)(row0 + r0) * 4 + wp + i * 128);
        } else {           // last tile: pad rows repeat the last valid row
#pragma unroll
          for (int i = 0; i < kPairsPerExpThread; ++i) bits[i] = __ldg(pairs + (size_t)(row0 + min(r0 + 32 * i, cnt - 1)) * 4 + wp);
        }
        const int rows_used = FULL ? kTN : tile_cols(cnt);
        mbar_wait_a(b_empty + 8 * bb, (par_bempty >> bb) & 1);
        par_bempty ^= 1u << bb;
        uint8_t* tile = sB + bb * kBTileBytes;
#pragma unroll
        for (int i = 0; i < kPairsPerExpThread; ++i) {
          if (r0 + 32 * i < rows_used) {
            *reinterpret_cast<uint4*>(tile + off_lo + i * 4096) = expand_word_fp4<false>(bits[i].x);
            *reinterpret_cast<uint4*>(tile + off_hi + i * 4096) = expand_word_fp4<false>(bits[i].y);
          }
        }
        fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive_a(b_full + 8 * bb);
        bb = bb + 1 == kBStages ? 0 : bb + 1;
      }
    }
  } else {
    // ===================== epilogue (256 threads: warp w owns TMEM lanes 32 (w % 4).., column half w / 4) ====
    const int quad = warp & 3, half = warp >> 2;
    const int row = quad * 32 + lane;  // row inside an A tile
    const uint32_t tmem_rows = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint32_t t = 0, kpar = 0;
    // this thread finishes the queries (tile m, row) with m % 2 == half; their running global best (distance, then
    // lowest keyframe: k only grows, so "<" keeps the first) stays in registers until the CTA is done
    uint32_t gbest_d[kQTilesPerCta / 2], gbest_k[kQTilesPerCta / 2];
#pragma unroll
    for (int j = 0; j < kQTilesPerCta / 2; ++j) gbest_d[j] = 0xffffffffu, gbest_k[j] = 0u;
#if LC_FX_ZERO_ONE_TRAIN
    int qpop[kQTilesPerCta / 2];  // popc of this thread's queries (expand_queries_fp4_kernel left them behind the A tiles): hamming = popc(q) - D
    {
      const uint16_t* pops = reinterpret_cast<const uint16_t*>(ab.img + kABytes);
#pragma unroll
      for (int j = 0; j < kQTilesPerCta / 2; ++j) qpop[j] = (int)__ldg(pops + (2 * j + half) * kQTile + row);
    }
#endif
    for (int k = k0 + slot; k < k0 + kc; k += n_slots) {
      const KfEntry kf = kfs[k];
      if (!kf_eligible(fr.cur_id, kf.frame_id, kb) || kf.n <= 0) continue;
      float best[kQTilesPerCta];
#pragma unroll
      for (int m = 0; m < kQTilesPerCta; ++m) best[m] = -1024.f;
      for (int row0 = 0; row0 < kf.n; row0 += kTN) {
        const int hcols = FULL ? kTN / 2 : tile_cols(min(kTN, kf.n - row0)) >> 1;  // columns per half: a multiple of 16
#pragma unroll
        for (int m = 0; m < kQTilesPerCta; ++m) {
          if (m < m_tiles) {
            const uint32_t st = t & 1u, par = (t >> 1) & 1u;
            mbar_wait_a(acc_full + 8 * st, par);
            tc_fence_after();
            best[m] = fmaxf(best[m], drain_accumulator(tmem_rows + st * kAccStride + (uint32_t)(half * hcols), hcols, acc_empty + 8 * st));
            ++t;
          }
        }
      }
      // the two column halves of a row meet in shared memory: half h hands over the tiles it does not finish
      float* x = xch + kpar * (kQTilesPerCta * kQTile);
#pragma unroll
      for (int m = 0; m < kQTilesPerCta; ++m)
        if ((m & 1) != half && m < m_tiles) x[m * kQTile + row] = best[m];
      epi_bar_sync();
      uint16_t* dst = scratch + ((size_t)ab.frame * kc + (k - k0)) * scratch_stride + ab.q0;
#pragma unroll
      for (int j = 0; j < kQTilesPerCta / 2; ++j) {
        const int m = 2 * j + half, q = m * kQTile + row;
        if (m < m_tiles && q < nq_blk) {
          const float mine = half ? best[2 * j + 1] : best[2 * j];  // static register indices
          const float dot = fmaxf(mine, x[m * kQTile + row]);
#if LC_FX_ZERO_ONE_TRAIN
          const uint32_t d = (uint32_t)(qpop[j] - (int)dot);       // dot = popc(q) - min hamming
#else
          const uint32_t d = (uint32_t)(256 - (int)dot) >> 1;      // dot = 256 - 2 * min hamming
#endif
          dst[q] = (uint16_t)d;
          if (d < gbest_d[j]) gbest_d[j] = d, gbest_k[j] = (uint32_t)k;
        }
      }
      kpar ^= 1;  // the next keyframe writes the other exchange buffer: one barrier per keyframe is enough
    }
    if (fr.best_packed) {
#pragma unroll
      for (int j = 0; j < kQTilesPerCta / 2; ++j) {
        const int q = (2 * j + half) * kQTile + row;
        if (gbest_d[j] != 0xffffffffu)
          atomicMin((unsigned long long*)(fr.best_packed + ab.q0 + q), ((unsigned long long)gbest_d[j] << 32) | gbest_k[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tc::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------- the pipe's measured peak (roofline denominator)
// One CTA per SM issues `iters` x 4 kind::mxf4 MMAs (M = 128, N = 256, K = 64, operands in shared memory with the match
// kernel's value distribution: A = -1 / +1, B = 0 / +1) back to back into one accumulator and waits for the last one:
// nothing but the tensor pipe.  MACs / elapsed time is what bench.py divides the match kernel's rate by, and
// cycles / elapsed time is the SM clock the pipe really ran at.
constexpr int kPeakN = 256;
constexpr int kPeakSmem = kATileBytes + kPeakN * 128 + 1024;
__global__ void __launch_bounds__(128, 1) mxf4_peak_kernel(long long iters, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint32_t rng = 0x9e3779b9u * (blockIdx.x * 128 + threadIdx.x + 1);
  for (int i = threadIdx.x; i < (kATileBytes + kPeakN * 128) / 16; i += blockDim.x) {
    rng = rng * 1664525u + 1013904223u;
    const uint32_t bits = rng >> 8;
    reinterpret_cast<uint4*>(smem)[i] = i < kATileBytes / 16 ? expand_word_fp4<true>(bits) : expand_word_fp4<false>(bits);
  }
  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tc::tmem_alloc(&tmem_slot, kTmemCols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  for (int c = 0; c < 64; ++c) tmem_st1(tmem_base + ((uint32_t)(warp * 32) << 16) + kPeakN + c, 0x7f7f7f7fu);  // scales = 1.0
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    const uint32_t idesc = instr_desc_mxf4(kPeakN);
    const uint64_t adesc = tc::smem_desc(smem_u32(smem)), bdesc = tc::smem_desc(smem_u32(smem) + kATileBytes);
    const bool leader = elect_one();
    const long long t0 = clock64();
    for (long long it = 0; it < iters; ++it) {
      if (leader) {
#pragma unroll
        for (int s = 0; s < 4; ++s)
          umma_mxf4(tmem_base, adesc + (uint64_t)(s * 2), bdesc + (uint64_t)(s * 2), idesc, (it | s) ? 1u : 0u, tmem_base + kPeakN, tmem_base + kPeakN + 32);
      }
      __syncwarp();
    }
    if (leader) umma_commit_a(smem_u32(&bar));
    __syncwarp();
    mbar_wait_a(smem_u32(&bar), 0);  // bounded: a tensor pipe that never answers traps instead of hanging the GPU
    if (threadIdx.x == 0) cycles_out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tc::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------- finalize (distance-only scratch)
// one WARP per (frame, keyframe in [k0, k0 + kc)): min distance, 2 x min filter, count -> results[frame * n_kf + k].
// The 16-bit distances of one match (<= 4 KiB) are read once with 128-bit loads and stay in registers for both passes.
constexpr int kFinWarps = 8;
constexpr int kFinMaxVec = 16;  // 16 x 32 lanes x 8 distances = 4096 queries per pass
__device__ __forceinline__ void finalize_dist_body(int block, const KfEntry* __restrict__ kfs, int n_kf, int k0, int kc,
                                                   const QFrame* __restrict__ frames, int n_frames, const uint16_t* __restrict__ scratch,
                                                   int scratch_stride, lc_kf_result* __restrict__ results, const Knobs& kb) {
  const int lane = threadIdx.x & 31;
  const int idx = block * kFinWarps + (threadIdx.x >> 5);
  if (idx >= kc * n_frames) return;
  const int f = idx / kc, k = k0 + (idx - f * kc);
  const int item = f * n_kf + k;
  const QFrame fr = frames[f];
  const KfEntry kf = kfs[k];
  if (!kf_eligible(fr.cur_id, kf.frame_id, kb) || kf.n <= 0 || fr.nq <= 0) {
    if (lane == 0) results[item] = lc_kf_result{kf.frame_id, kf.n, -1, -1};
    return;
  }
  const uint4* dists = reinterpret_cast<const uint4*>(scratch + (size_t)idx * scratch_stride);
  int dmin = 0x7fffffff, cnt_good = 0;
  for (int base = 0; base < fr.nq; base += kFinMaxVec * 256) {  // one pass unless nq > 4096
    uint4 v[kFinMaxVec];
    uint32_t mn = 0xffffffffu;
#pragma unroll
    for (int i = 0; i < kFinMaxVec; ++i) {
      const int e = base + (i * 32 + lane) * 8;  // first query of this 128-bit vector
      v[i] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
      if (e < fr.nq) {
        v[i] = __ldg(dists + (e >> 3));
        if (e + 8 > fr.nq) {  // the vector straddles nq: entries past it were never written
          uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (e + j >= fr.nq) w[j >> 1] |= 0xffffu << (16 * (j & 1));
          v[i] = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      mn = __vminu2(mn, __vminu2(__vminu2(v[i].x, v[i].y), __vminu2(v[i].z, v[i].w)));
    }
    if (base == 0 && fr.nq <= kFinMaxVec * 256) {  // the common case: everything is in registers
      dmin = (int)min(mn & 0xffffu, mn >> 16);
      dmin = __reduce_min_sync(0xffffffffu, dmin);
      if (kb.min_init >= 0) dmin = min(dmin, kb.min_init);
      const int thr = filter_threshold(dmin, kb) - (kb.filter_strict ? 1 : 0);  // good iff d <= thr
      const uint32_t thr2 = (uint32_t)max(thr, -1) * 0x10001u;                  // thr < 0: nothing passes
      uint32_t acc = 0;
      if (thr >= 0) {
#pragma unroll
        for (int i = 0; i < kFinMaxVec; ++i)  // __vsetleu2: 1 per 16-bit lane with a <= b
          acc += __vsetleu2(v[i].x, thr2) + __vsetleu2(v[i].y, thr2) + __vsetleu2(v[i].z, thr2) + __vsetleu2(v[i].w, thr2);
      }
      cnt_good = __reduce_add_sync(0xffffffffu, (int)((acc & 0xffffu) + (acc >> 16)));
    } else {  // nq > 4096: min over all passes first, then a second sweep for the count
      dmin = min(dmin, (int)min(mn & 0xffffu, mn >> 16));
    }
  }
  if (fr.nq > kFinMaxVec * 256) {
    dmin = __reduce_min_sync(0xffffffffu, dmin);
    if (kb.min_init >= 0) dmin = min(dmin, kb.min_init);
    const int thr = filter_threshold(dmin, kb);
    const uint16_t* d16 = scratch + (size_t)idx * scratch_stride;
    for (int i = lane; i < fr.nq; i += 32) cnt_good += is_good((int)d16[i], thr, kb) ? 1 : 0;
    cnt_good = __reduce_add_sync(0xffffffffu, cnt_good);
  }
  if (lane == 0) results[item] = lc_kf_result{kf.frame_id, kf.n, dmin, cnt_good};
}
__global__ void __launch_bounds__(kFinWarps * 32)
    finalize_dist_kernel(const KfEntry* __restrict__ kfs, int n_kf, int k0, int kc, const QFrame* __restrict__ frames, int n_frames,
                         const uint16_t* __restrict__ scratch, int scratch_stride, lc_kf_result* __restrict__ results, Knobs kb) {
  pdl_trigger();
  pdl_wait();
  finalize_dist_body(blockIdx.x, kfs, n_kf, k0, kc, frames, n_frames, scratch, scratch_stride, results, kb);
}

// best_packed[i] = dist << 32 | local keyframe index  ->  dist << 32 | global row of the first row of that keyframe
// at that distance (lowest keyframe, then lowest row: OpenCV's multi-image order).  One warp per query.
constexpr int kResolveGroups = 8;
__device__ __forceinline__ void resolve_best_body(int block, int n_blocks, const KfEntry* __restrict__ kfs, const QFrame* __restrict__ frames,
                                                  int n_frames) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int f = 0; f < n_frames; ++f) {
    const QFrame fr = frames[f];
    if (!fr.best_packed) continue;
    for (int i = block * wpb + (threadIdx.x >> 5); i < fr.nq; i += n_blocks * wpb) {
      const uint64_t packed = fr.best_packed[i];
      if (packed == ~0ull) continue;  // nothing was compared
      const int d = (int)(packed >> 32);
      const KfEntry kf = kfs[(uint32_t)packed];
      const uint4 qa = __ldg(fr.q + 2 * (size_t)i), qb = __ldg(fr.q + 2 * (size_t)i + 1);
      int found = -1;
      // kResolveGroups independent 32-row groups in flight per step: a step is one L2 round trip, and the steps of a
      // query are serial (the scan stops at the first hit), so their number is the kernel's latency
      for (int j0 = 0; j0 < kf.n && found < 0; j0 += 32 * kResolveGroups) {
        uint32_t hit[kResolveGroups];
        uint4 a[kResolveGroups], b[kResolveGroups];
#pragma unroll
        for (int u = 0; u < kResolveGroups; ++u) {  // all loads first (rows past the end: the last row again)
          const int j = min(j0 + 32 * u + lane, kf.n - 1);
          a[u] = __ldg(kf.rows + 2 * (size_t)j);
          b[u] = __ldg(kf.rows + 2 * (size_t)j + 1);
        }
#pragma unroll
        for (int u = 0; u < kResolveGroups; ++u) {
          const int dj = __popc(qa.x ^ a[u].x) + __popc(qa.y ^ a[u].y) + __popc(qa.z ^ a[u].z) + __popc(qa.w ^ a[u].w) +
                         __popc(qb.x ^ b[u].x) + __popc(qb.y ^ b[u].y) + __popc(qb.z ^ b[u].z) + __popc(qb.w ^ b[u].w);
          hit[u] = __ballot_sync(0xffffffffu, dj == d && j0 + 32 * u + lane < kf.n);
        }
#pragma unroll
        for (int u = kResolveGroups - 1; u >= 0; --u)
          if (hit[u]) found = j0 + 32 * u + __ffs(hit[u]) - 1;  // the lowest group wins
      }
#if !LC_FX_DRAIN_EXPERIMENT
      if (found < 0) __trap();  // the distance came from this keyframe: a miss is a kernel bug
#endif
      if (lane == 0) fr.best_packed[i] = ((uint64_t)d << 32) | (uint64_t)(kf.global_row_off + found);
    }
  }
}
__global__ void __launch_bounds__(256)
    resolve_best_kernel(const KfEntry* __restrict__ kfs, const QFrame* __restrict__ frames, int n_frames) {
  pdl_wait();
  resolve_best_body(blockIdx.x, gridDim.x, kfs, frames, n_frames);
}
// Both steps after the matching kernel in one launch when all keyframes went through in one batch (they are independent:
// one reads the distance scratch, the other the per-query bests): blocks [0, n_fin) finalize, the rest resolve.
static_assert(kFinWarps * 32 == 256, "finish_kernel runs both bodies with 256 threads");
__global__ void __launch_bounds__(256)
    finish_kernel(int n_fin, const KfEntry* __restrict__ kfs, int n_kf, const QFrame* __restrict__ frames, int n_frames,
                  const uint16_t* __restrict__ scratch, int scratch_stride, lc_kf_result* __restrict__ results, Knobs kb) {
  pdl_wait();
  if ((int)blockIdx.x < n_fin)
    finalize_dist_body(blockIdx.x, kfs, n_kf, 0, n_kf, frames, n_frames, scratch, scratch_stride, results, kb);
  else
    resolve_best_body(blockIdx.x - n_fin, gridDim.x - n_fin, kfs, frames, n_frames);
}

}  // namespace fx
}  // namespace lc
