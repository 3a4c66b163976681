// tcgen05.mma kind::mxf4 (E2M1 x E2M1, UE8M0 block-32 scales) probe for the Hamming formulation:
//   1. correctness: one 128 x N x 256 tile of +1/-1 expanded bits + the tie-break step, checked on the host;
//   2. issue rate (cycles per MMA) and sustained whole-chip MAC rate (wall clock, power cap included) next to kind::i8;
//   3. tcgen05.ld read bandwidth with 8 warps.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_fp4 tools/umma_fp4.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {  // K-major SWIZZLE_128B, 8-row groups 1024 B apart
  return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_i8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5,%5,%5,%5}, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_mxf4(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc, uint32_t sfa, uint32_t sfb) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(sfa), "r"(sfb) : "memory");
}
__device__ __forceinline__ uint32_t idesc_i8(int n) { return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (8u << 24); }
// block-scaled: a/b format E2M1 (1), scale format UE8M0 (bit 23), sf ids: byte offset of the first scale inside a 32-bit TMEM cell
__device__ __forceinline__ uint32_t idesc_mxf4(int n, uint32_t a_sf_id, uint32_t b_sf_id) {
  return (b_sf_id << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (1u << 23) | (8u << 24) | (a_sf_id << 29);
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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

// 8 descriptor bits -> 8 E2M1 nibbles: bit clear -> +1.0 (0x2), bit set -> -1.0 (0xA); bit i -> nibble i
__host__ __device__ __forceinline__ uint32_t expand8_fp4(uint32_t x) {
  x = (x | (x << 12)) & 0x000F000Fu;
  x = (x | (x << 6)) & 0x03030303u;
  x = (x | (x << 3)) & 0x11111111u;
  return x * 8u + 0x22222222u;
}

constexpr int kSfCol = 448;        // TMEM columns [448, 512) hold the (uniform) scale factors
constexpr int kSmemOps = 200 * 1024;

struct Setup {
  uint8_t* smem;
  uint64_t* bar;
  uint32_t tb;
};

// ---------------------------------------------------------------- 1. correctness
// A: 128 query rows (32 B packed bits each), B: N train rows.  Tile image: row r at r * 128 B, 16-byte chunk w (= the
// 32 nibbles of descriptor word w) at chunk index w ^ (r & 7).  D' = S * dot + (N - 1 - c) via a fifth (tie) step.
__global__ void __launch_bounds__(128, 1) correctness(const uint32_t* qbits, const uint32_t* tbits, int N, float* out, uint32_t sfa_cell, int tie) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  uint8_t* sA = smem;                 // 128 x 128 B
  uint8_t* sB = smem + 16384;         // 256 x 128 B
  uint8_t* sTA = smem + 16384 + 32768;  // tie A: 128 rows x 128 B (only the first 32 B of each row are read)
  uint8_t* sTB = sTA + 16384;           // tie B: 256 rows x 128 B
  const int tid = threadIdx.x;
  for (int i = tid; i < 128 * 8; i += 128) {
    const int r = i >> 3, w = i & 7;
    const uint32_t b = qbits[i];
    *(uint4*)(sA + r * 128 + ((w ^ (r & 7)) << 4)) = make_uint4(expand8_fp4(b & 255), expand8_fp4((b >> 8) & 255), expand8_fp4((b >> 16) & 255), expand8_fp4(b >> 24));
  }
  for (int i = tid; i < N * 8; i += 128) {
    const int r = i >> 3, w = i & 7;
    const uint32_t b = tbits[i];
    *(uint4*)(sB + r * 128 + ((w ^ (r & 7)) << 4)) = make_uint4(expand8_fp4(b & 255), expand8_fp4((b >> 8) & 255), expand8_fp4((b >> 16) & 255), expand8_fp4(b >> 24));
  }
  // tie operands: A_tie = 1.0 in all 64 k; B_tie[c] = e2m1 decomposition of (N - 1 - c): sixes then the remainder
  for (int i = tid; i < 128 * 2; i += 128) {
    const int r = i >> 1, ch = i & 1;
    *(uint4*)(sTA + r * 128 + ((ch ^ (r & 7)) << 4)) = make_uint4(0x22222222u, 0x22222222u, 0x22222222u, 0x22222222u);
  }
  for (int c = tid; c < N; c += 128) {
    uint8_t nib[64];
    int t = N - 1 - c, k = 0;
    for (int j = 0; j < 64; ++j) nib[j] = 0;
    while (t >= 6) { nib[k++] = 0x7; t -= 6; }        // 6.0 = 0 11 1
    if (t == 5) { nib[k++] = 0x6; nib[k++] = 0x2; }     // 4 + 1
    else if (t == 4) nib[k++] = 0x6;                    // 4.0 = 0 11 0
    else if (t == 3) nib[k++] = 0x5;                    // 3.0 = 0 10 1
    else if (t == 2) nib[k++] = 0x4;                    // 2.0 = 0 10 0
    else if (t == 1) nib[k++] = 0x2;                    // 1.0 = 0 01 0
    for (int ch = 0; ch < 2; ++ch) {
      uint32_t w4[4];
      for (int j = 0; j < 4; ++j) {
        uint32_t w = 0;
        for (int e = 0; e < 8; ++e) w |= (uint32_t)nib[ch * 32 + j * 8 + e] << (4 * e);
        w4[j] = w;
      }
      *(uint4*)(sTB + c * 128 + ((ch ^ (c & 7)) << 4)) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
    }
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = tslot;
  const int warp = tid >> 5;
  // uniform scale factors: SFA cells = sfa_cell in columns [448, 480), SFB cells = 1.0 (127) in [480, 512)
  for (int c = 0; c < 32; ++c) tmem_st1(tb + ((uint32_t)(warp * 32) << 16) + kSfCol + c, sfa_cell);
  for (int c = 0; c < 32; ++c) tmem_st1(tb + ((uint32_t)(warp * 32) << 16) + kSfCol + 32 + c, 0x7f7f7f7fu);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (tid == 0) {
    const uint64_t da = smem_desc(smem_u32(sA)), db = smem_desc(smem_u32(sB));
    for (int s = 0; s < 4; ++s)
      umma_mxf4(tb, da + (uint64_t)((s * 32) >> 4), db + (uint64_t)((s * 32) >> 4), idesc_mxf4(N, 0, 0), s > 0, tb + kSfCol, tb + kSfCol + 32);
    if (tie) umma_mxf4(tb, smem_desc(smem_u32(sTA)), smem_desc(smem_u32(sTB)), idesc_mxf4(N, 2, 0), 1u, tb + kSfCol, tb + kSfCol + 32);
    commit(&bar);
  }
  __syncthreads();
  wait_bar(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;");
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + c0, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u));
}

// ---------------------------------------------------------------- 2. MMA rate
// KIND 0 = i8 (K = 32 elements = 32 B per MMA), 1 = mxf4 (K = 64 elements = 32 B per MMA).  Each iteration issues
// `per_tile` MMAs into accumulator (it % n_acc); operands are random +/-1 images already in shared memory.
template <int KIND>
__global__ void __launch_bounds__(128, 1) rate(int N, long long iters, int n_acc, int per_tile, long long* out, int commit_each = 0) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[2];
  __shared__ uint32_t tslot;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[1])));
  }
  uint32_t rng = 0x9e3779b9u * (blockIdx.x * 128 + threadIdx.x + 1);
  for (int i = threadIdx.x; i < kSmemOps / 4; i += blockDim.x) {
    rng = rng * 1664525u + 1013904223u;
    const uint32_t bits = rng >> 8;
    uint32_t w;
    if (KIND == 0) { const uint32_t s = ((bits & 15u) * 0x00204081u) & 0x01010101u; w = s * 0xFEu + 0x01010101u; }
    else w = expand8_fp4(bits & 255u);
    ((uint32_t*)smem)[i] = w;
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = tslot;
  const int warp = threadIdx.x >> 5;
  if (KIND == 1) {
    for (int c = 0; c < 64; ++c) tmem_st1(tb + ((uint32_t)(warp * 32) << 16) + kSfCol + c, 0x7f7f7f7fu);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (threadIdx.x < 32) {
    const uint32_t idesc = KIND == 0 ? idesc_i8(N) : idesc_mxf4(N, 0, 0);
    const uint64_t da = smem_desc(smem_u32(smem)), db = smem_desc(smem_u32(smem) + 64 * 1024);
    uint32_t el;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
    const int acc_stride = N <= 128 ? 128 : 256;  // accumulators stay below column 448 when n_acc * stride <= 448 (or 1 x 256)
    long long t0 = clock64();
    for (long long it = 0; it < iters; ++it) {
      const uint32_t d = tb + (uint32_t)(it % n_acc) * acc_stride;
      if (el) {
        for (int s = 0; s < per_tile; ++s) {
          // i8: 8 steps of 32 B over two 128-byte K blocks; mxf4: 4 steps of 32 B inside one 128-byte row
          const uint32_t ka = KIND == 0 ? (uint32_t)((s >> 2) & 1) * 16384 + (uint32_t)(s & 3) * 32 : (uint32_t)(s & 3) * 32;
          const uint32_t kb = KIND == 0 ? (uint32_t)((s >> 2) & 1) * 32768 + (uint32_t)(s & 3) * 32 : (uint32_t)(s & 3) * 32;
          if (KIND == 0) umma_i8(d, da + (ka >> 4), db + (kb >> 4), idesc, s > 0);
          else umma_mxf4(d, da + (ka >> 4), db + (kb >> 4), idesc, s > 0, tb + kSfCol, tb + kSfCol + 32);
        }
        if (commit_each) commit(&bar2[it & 1]);
      }
      __syncwarp();
    }
    if (el) commit(&bar);
    __syncwarp();
    wait_bar(&bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u));
}

// ---------------------------------------------------------------- 3. TMEM read bandwidth
__global__ void __launch_bounds__(256, 1) tmem_read(int iters, int ncols, long long* out, uint32_t* sink) {
  __shared__ uint32_t tslot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = tslot;
  const int warp = threadIdx.x >> 5, quad = warp & 3, half = warp >> 2;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int c = 0; c < ncols; c += 64) {
      uint32_t v0[32], v1[32];
      const uint32_t ta = tb + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 256 + ((c + it * 64) & 255));
      tmem_ld32(ta, v0);
      tmem_ld32(ta + 32, v1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; j += 2) acc = max(acc, max(v0[j] ^ v1[j], v0[j + 1] ^ v1[j + 1]));
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345u) sink[threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u));
}

// MMA issue (warp 0) into alternating accumulators while warps 4-11 stream tcgen05.ld over both accumulators
// (no synchronisation between them: bandwidth interference only)
__global__ void __launch_bounds__(384, 1) both(int N, long long iters, int readers, long long* out, uint32_t* sink) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  __shared__ volatile int done;
  uint32_t rng = 0x9e3779b9u * (blockIdx.x * 384 + threadIdx.x + 1);
  for (int i = threadIdx.x; i < kSmemOps / 4; i += blockDim.x) {
    rng = rng * 1664525u + 1013904223u;
    ((uint32_t*)smem)[i] = expand8_fp4((rng >> 8) & 255u);
  }
  if (threadIdx.x == 0) {
    done = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = tslot;
  const int warp = threadIdx.x >> 5;
  if (warp < 4) {
    for (int c = 0; c < 32; ++c) tmem_st1(tb + ((uint32_t)(warp * 32) << 16) + 480 + c, 0x7f7f7f7fu);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (warp == 0) {
    const uint32_t idesc = idesc_mxf4(N, 0, 0);
    const uint64_t da = smem_desc(smem_u32(smem)), db = smem_desc(smem_u32(smem) + 64 * 1024);
    uint32_t el;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
    long long t0 = clock64();
    for (long long it = 0; it < iters; ++it) {
      const uint32_t d = tb + (uint32_t)(it & 1) * 240;
      if (el)
        for (int s = 0; s < 4; ++s) umma_mxf4(d, da + (uint64_t)((s * 32) >> 4), db + (uint64_t)((s * 32) >> 4), idesc, s > 0, tb + 480, tb + 496);
      __syncwarp();
    }
    if (el) commit(&bar);
    __syncwarp();
    wait_bar(&bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x] = t1 - t0; done = 1; }
  } else if (warp >= 4 && readers) {
    const int w = warp - 4, quad = w & 3, half = w >> 2;
    float acc = 0.f;
    long long nloops = 0;
    while (!done) {
      ++nloops;
      for (int c = 0; c < 192; c += 64) {
        uint32_t v0[32], v1[32];
        const uint32_t ta = tb + ((uint32_t)(quad * 32) << 16) + (uint32_t)(half * 240 + c);
        tmem_ld32(ta, v0);
        tmem_ld32(ta + 32, v1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; j += 2) acc = fmaxf(acc, fmaxf(fmaxf(__uint_as_float(v0[j]), __uint_as_float(v1[j])), fmaxf(__uint_as_float(v0[j + 1]), __uint_as_float(v1[j + 1]))));
      }
    }
    if (acc == 12345.f) sink[threadIdx.x] = 1;
    if (threadIdx.x == 128) out[148 + blockIdx.x] = nloops;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u));
}

// two issuer warps taking turns tile by tile (4 MMAs + commit each); a named barrier hands the turn over right after the
// fourth MMA, so one warp's commit overlaps the other warp's MMA issue
__global__ void __launch_bounds__(128, 1) rate2(int N, int iters, long long* out, int handoff_before_commit) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tslot;
  uint32_t rng = 0x9e3779b9u * (blockIdx.x * 128 + threadIdx.x + 1);
  for (int i = threadIdx.x; i < kSmemOps / 4; i += blockDim.x) {
    rng = rng * 1664525u + 1013904223u;
    ((uint32_t*)smem)[i] = expand8_fp4((rng >> 8) & 255u);
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = tslot;
  const int warp = threadIdx.x >> 5;
  for (int c = 0; c < 32; ++c) tmem_st1(tb + ((uint32_t)(warp * 32) << 16) + 480 + c, 0x7f7f7f7fu);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (warp < 2) {
    const uint32_t idesc = idesc_mxf4(N, 0, 0);
    const uint64_t da = smem_desc(smem_u32(smem)), db = smem_desc(smem_u32(smem) + 64 * 1024);
    uint32_t el;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
    const uint32_t d = tb + (uint32_t)warp * 240;
    long long t0 = clock64();
    if (warp == 1) asm volatile("bar.sync 3, 64;" ::: "memory");  // wait for warp 0's first turn
    for (int it = 0; it < iters; ++it) {
      if (el)
        for (int s = 0; s < 4; ++s) umma_mxf4(d, da + (uint64_t)((s * 32) >> 4), db + (uint64_t)((s * 32) >> 4), idesc, s > 0, tb + 480, tb + 496);
      __syncwarp();
      if (handoff_before_commit) {
        asm volatile("bar.arrive %0, 64;" ::"r"(2 + (warp ^ 1)) : "memory");   // the other warp may go
        if (el) commit(&bar[warp]);
      } else {
        if (el) commit(&bar[warp]);
        asm volatile("bar.arrive %0, 64;" ::"r"(2 + (warp ^ 1)) : "memory");
      }
      __syncwarp();
      if (it + 1 < iters || warp == 0) asm volatile("bar.sync %0, 64;" ::"r"(2 + warp) : "memory");  // my next turn
    }
    // wait for my last commit
    wait_bar(&bar[warp], (iters - 1) & 1);
    long long t1 = clock64();
    if (threadIdx.x == 32) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u));
}

// pure TMEM read bandwidth: `nwarps` warps, each issuing `depth` x32 loads back to back before one wait (no math),
// optionally while warp 0 keeps the tensor pipe busy
template <int DEPTH>
__global__ void __launch_bounds__(544, 1) tmem_bw(int iters, int nwarps, int with_mma, long long* out, uint32_t* sink, int mode = 0) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  __shared__ int done;  // reader warps that have finished
  for (int i = threadIdx.x; i < kSmemOps / 4; i += blockDim.x) ((uint32_t*)smem)[i] = expand8_fp4((i * 2654435761u >> 8) & 255u);
  if (threadIdx.x == 0) {
    done = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = tslot;
  const int warp = threadIdx.x >> 5;
  if (warp < 4) {
    for (int c = 0; c < 32; ++c) tmem_st1(tb + ((uint32_t)(warp * 32) << 16) + 480 + c, 0x7f7f7f7fu);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (warp == 0) {
    if (with_mma) {
      const uint32_t idesc = idesc_mxf4(224, 0, 0);
      const uint64_t da = smem_desc(smem_u32(smem)), db = smem_desc(smem_u32(smem) + 64 * 1024);
      uint32_t el;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
      long long it = 0;
      const long long m0 = clock64();
      while (*(volatile int*)&done < nwarps && it < (1ll << 22)) {  // bounded: a protocol slip must not hang the GPU
        const uint32_t d = tb + (uint32_t)(it++ & 1) * 240;
        if (el)
          for (int s = 0; s < 4; ++s) umma_mxf4(d, da + (uint64_t)((s * 32) >> 4), db + (uint64_t)((s * 32) >> 4), idesc, s > 0, tb + 480, tb + 496);
        __syncwarp();
      }
      if (el) commit(&bar);
      __syncwarp();
      wait_bar(&bar, 0);
      if (threadIdx.x == 0) { out[148 * 16 + blockIdx.x * 2] = clock64() - m0; out[148 * 16 + blockIdx.x * 2 + 1] = it * 4; }
    }
  } else if (warp <= nwarps) {
    const int w = warp - 1, quad = warp & 3;
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      uint32_t v[DEPTH][32];
#pragma unroll
      for (int k = 0; k < DEPTH; ++k) tmem_ld32(tb + ((uint32_t)(quad * 32) << 16) + (uint32_t)(((w >> 2) * 32 * DEPTH + k * 32 + it * 64) & 255), v[k]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (mode & 1) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      if (mode & 2) {
        float m0 = -1e30f, m1 = -1e30f, m2 = -1e30f, m3 = -1e30f;
#pragma unroll
        for (int k = 0; k < DEPTH; ++k)
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            m0 = fmaxf(m0, fmaxf(__uint_as_float(v[k][j]), __uint_as_float(v[k][j + 1])));
            m1 = fmaxf(m1, fmaxf(__uint_as_float(v[k][j + 2]), __uint_as_float(v[k][j + 3])));
            m2 = fmaxf(m2, fmaxf(__uint_as_float(v[k][j + 4]), __uint_as_float(v[k][j + 5])));
            m3 = fmaxf(m3, fmaxf(__uint_as_float(v[k][j + 6]), __uint_as_float(v[k][j + 7])));
          }
        acc ^= __float_as_uint(fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
      } else {
#pragma unroll
        for (int k = 0; k < DEPTH; ++k) acc ^= v[k][0] ^ v[k][31];
      }
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) { out[blockIdx.x * 16 + w] = t1 - t0; atomicAdd(&done, 1); }
    if (acc == 0x12345u) sink[threadIdx.x] = acc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u));
}

// MMA rate (warp 0, N = 224, alternating accumulators) while `npoll` other warps spin on mbarrier.try_wait of a phase that
// never completes: does barrier polling steal shared-memory bandwidth from the tensor core's operand fetch?
__global__ void __launch_bounds__(544, 1) poll_vs_mma(int iters, int npoll, int sleep_ns, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, never;
  __shared__ uint32_t tslot;
  __shared__ int done;
  for (int i = threadIdx.x; i < kSmemOps / 4; i += blockDim.x) ((uint32_t*)smem)[i] = expand8_fp4((i * 2654435761u >> 8) & 255u);
  if (threadIdx.x == 0) {
    done = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&never)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = tslot;
  const int warp = threadIdx.x >> 5;
  if (warp < 4) {
    for (int c = 0; c < 32; ++c) tmem_st1(tb + ((uint32_t)(warp * 32) << 16) + 480 + c, 0x7f7f7f7fu);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (warp == 0) {
    const uint32_t idesc = idesc_mxf4(224, 0, 0);
    const uint64_t da = smem_desc(smem_u32(smem)), db = smem_desc(smem_u32(smem) + 64 * 1024);
    uint32_t el;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tb + (uint32_t)(it & 1) * 240;
      if (el)
        for (int s = 0; s < 4; ++s) umma_mxf4(d, da + (uint64_t)((s * 32) >> 4), db + (uint64_t)((s * 32) >> 4), idesc, s > 0, tb + 480, tb + 496);
      __syncwarp();
    }
    if (el) commit(&bar);
    __syncwarp();
    wait_bar(&bar, 0);
    if (threadIdx.x == 0) { out[blockIdx.x] = clock64() - t0; *(volatile int*)&done = 1; }
  } else if (warp <= npoll) {
    long long guard = 0;
    while (!*(volatile int*)&done && guard++ < (1ll << 24)) {
      uint32_t ok;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&never)), "r"(0u) : "memory");
      if (sleep_ns) __nanosleep(sleep_ns);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u));
}

static int popc32(uint32_t x) { return __builtin_popcount(x); }

int main(int argc, char** argv) {
  setvbuf(stdout, nullptr, _IOLBF, 0);
  const double sustain_s = argc > 1 ? atof(argv[1]) : 1.5;
  long long* d; CK(cudaMalloc(&d, 20 * 148 * 8));
  uint32_t* sink; CK(cudaMalloc(&sink, 4096));
  CK(cudaFuncSetAttribute(correctness, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
  CK(cudaFuncSetAttribute(rate<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
  CK(cudaFuncSetAttribute(rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
  CK(cudaFuncSetAttribute(both, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
  CK(cudaFuncSetAttribute(rate2, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));

  // ---- 1. correctness
  for (int N : {128, 256}) {
    for (int tie : {0, 1}) {
      std::vector<uint32_t> q(128 * 8), t(N * 8);
      srand(7 + N);
      for (auto& x : q) x = (uint32_t)rand() ^ ((uint32_t)rand() << 16);
      for (auto& x : t) x = (uint32_t)rand() ^ ((uint32_t)rand() << 16);
      for (int w = 0; w < 8; ++w) { t[5 * 8 + w] = q[3 * 8 + w]; t[9 * 8 + w] = q[3 * 8 + w]; t[7 * 8 + w] = ~q[4 * 8 + w]; }
      uint32_t *dq, *dt; float* dout;
      CK(cudaMalloc(&dq, q.size() * 4)); CK(cudaMalloc(&dt, t.size() * 4)); CK(cudaMalloc(&dout, 128 * N * 4));
      CK(cudaMemcpy(dq, q.data(), q.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(dt, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
      const int log2s = N == 128 ? 7 : 8;  // main scale 2^7 / 2^8 (bytes 0,1 of the SFA cell), tie scale 1 (bytes 2,3)
      const uint32_t cell = (uint32_t)(127 + log2s) | ((uint32_t)(127 + log2s) << 8) | (127u << 16) | (127u << 24);
      correctness<<<1, 128, 201 * 1024>>>(dq, dt, N, dout, cell, tie);
      CK(cudaDeviceSynchronize());
      std::vector<float> out(128 * N);
      CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
      long long bad = 0; int shown = 0;
      for (int r = 0; r < 128; ++r)
        for (int c = 0; c < N; ++c) {
          int ham = 0;
          for (int w = 0; w < 8; ++w) ham += popc32(q[r * 8 + w] ^ t[c * 8 + w]);
          const float want = (float)((1 << log2s) * (256 - 2 * ham) + (tie ? (N - 1 - c) : 0));
          if (out[(size_t)r * N + c] != want) {
            if (shown++ < 6) printf("  mismatch r=%d c=%d got %.1f want %.1f (ham %d)\n", r, c, out[(size_t)r * N + c], want, ham);
            ++bad;
          }
        }
      printf("correctness mxf4 N=%d tie=%d: %lld mismatches of %d\n", N, tie, bad, 128 * N);
      cudaFree(dq); cudaFree(dt); cudaFree(dout);
    }
  }

  // ---- 2. rates
  long long h[148];
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  printf("kind   N n_acc per_tile  cycles/MMA  MAC/clk/SM   | sustained: s  TMAC/s  eff_MHz(from cycles)\n");
  for (int kind : {0, 1})
    for (int N : {128, 256})
      for (int n_acc : {1, 2}) {
        if (N == 256 && n_acc == 2 && kind == 1) continue;  // would overlap the scale-factor columns
        const int per_tile = kind == 0 ? 8 : 4;
        const int kelem = kind == 0 ? 32 : 64;
        long long iters = 4096;
        for (int rep = 0; rep < 2; ++rep) {
          if (kind == 0) rate<0><<<148, 128, 201 * 1024>>>(N, iters, n_acc, per_tile, d); else rate<1><<<148, 128, 201 * 1024>>>(N, iters, n_acc, per_tile, d);
          CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost));
        double mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
        const double per = mx / (iters * per_tile);
        // sustained run
        const long long iters2 = (long long)(sustain_s * 1.9e9 / (per * per_tile));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        if (kind == 0) rate<0><<<148, 128, 201 * 1024>>>(N, iters2, n_acc, per_tile, d); else rate<1><<<148, 128, 201 * 1024>>>(N, iters2, n_acc, per_tile, d);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        CK(cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost));
        double mx2 = 0; for (int i = 0; i < 148; ++i) mx2 = h[i] > mx2 ? h[i] : mx2;
        const double macs = 148.0 * iters2 * per_tile * 128.0 * N * kelem;
        printf("%s %4d %3d %5d %12.1f %12.0f   | %8.2f %8.1f %8.0f\n", kind ? "mxf4" : "i8  ", N, n_acc, per_tile, per, 128.0 * N * kelem / per,
               ms / 1e3, macs / (ms * 1e-3) / 1e12, mx2 / (ms * 1e-3) / 1e6);
      }

  // ---- 2b. mxf4 N sweep (cycles per MMA, one accumulator)
  for (int N : {128, 160, 192, 208, 224, 240, 248, 256}) {
    const long long iters = 4096;
    for (int rep = 0; rep < 2; ++rep) { rate<1><<<148, 128, 201 * 1024>>>(N, iters, 1, 4, d); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost));
    double mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const double per = mx / (iters * 4);
    printf("mxf4 N=%3d: %6.1f cycles/MMA (floor %5.1f)  %6.0f MAC/clk/SM\n", N, per, N / 2.0, 128.0 * N * 64 / per);
  }
  // ---- 2b'. one commit per tile (4 MMAs) into alternating accumulators
  for (int ce : {0, 1}) {
    const long long iters = 4096;
    for (int rep = 0; rep < 2; ++rep) { rate<1><<<148, 128, 201 * 1024>>>(224, iters, 2, 4, d, ce); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost));
    double mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("mxf4 N=224 2 accumulators, commit per tile=%d: %6.1f cycles/MMA\n", ce, mx / (iters * 4));
  }
  for (int hb : {0, 1}) {
    const int iters = 2048;  // per warp
    for (int rep = 0; rep < 2; ++rep) { rate2<<<148, 128, 201 * 1024>>>(224, iters, d, hb); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost));
    double mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("mxf4 N=224, two issuer warps taking turns, commit per tile, hand-off %s commit: %6.1f cycles/MMA\n", hb ? "before" : "after", mx / (2.0 * iters * 4));
  }
  // ---- 2b''. N sweep with a tight issue loop (the `both` kernel without readers): is there a per-MMA floor?
  for (int N : {32, 64, 96, 112, 128, 160, 192, 224}) {
    const long long iters = 4096;
    for (int rep = 0; rep < 2; ++rep) { both<<<148, 384, 201 * 1024>>>(N, iters, 0, d, sink); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost));
    double mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("mxf4 N=%3d tight issue loop: %6.1f cycles/MMA (N/2 = %5.1f)\n", N, mx / (iters * 4), N / 2.0);
  }
  // ---- 2c. MMA (warp 0) and TMEM reads (warps 4-11) at the same time
  for (int N : {224, 240}) {
    for (int readers : {0, 1}) {
      const long long iters = 4096;
      for (int rep = 0; rep < 2; ++rep) { both<<<148, 384, 201 * 1024>>>(N, iters, readers, d, sink); CK(cudaDeviceSynchronize()); }
      CK(cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost));
      double mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      long long h2[148]; CK(cudaMemcpy(h2, d + 148, 148 * 8, cudaMemcpyDeviceToHost));
      printf("mxf4 N=%3d 2 accumulators, readers=%d: %6.1f cycles/MMA", N, readers, mx / (iters * 4));
      if (readers) printf("   readers: %.1f B/clk/SM while the MMAs run", (double)h2[0] * 3 * 8 * 2 * 32 * 32 * 4 / (double)h[0]);
      printf("\n");
    }
  }

  // ---- 2d. mbarrier polling next to the MMAs
  CK(cudaFuncSetAttribute(poll_vs_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
  for (int sl : {0, 100})
    for (int np : {0, 1, 4, 8, 12, 16}) {
      const int iters = 4096;
      for (int rep = 0; rep < 2; ++rep) { poll_vs_mma<<<148, 544, 201 * 1024>>>(iters, np, sl, d); CK(cudaDeviceSynchronize()); }
      CK(cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost));
      double mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("mxf4 N=224 with %2d warps polling mbarrier.try_wait (nanosleep %3d): %6.1f cycles/MMA\n", np, sl, mx / (iters * 4.0));
    }
  // ---- 3a. pure TMEM read bandwidth (no math): warps x loads in flight, with and without MMAs running
  {
    CK(cudaFuncSetAttribute(tmem_bw<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
    CK(cudaFuncSetAttribute(tmem_bw<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
    CK(cudaFuncSetAttribute(tmem_bw<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
    static long long hh[148 * 16];
    for (int with_mma : {0, 1})
      for (int nw : {4, 8, 16})
        for (int depth : {1, 2, 4}) {
          const int iters = 4000;
          for (int rep = 0; rep < 2; ++rep) {
            if (depth == 1) tmem_bw<1><<<148, 544, 201 * 1024>>>(iters, nw, with_mma, d, sink);
            else if (depth == 2) tmem_bw<2><<<148, 544, 201 * 1024>>>(iters, nw, with_mma, d, sink);
            else tmem_bw<4><<<148, 544, 201 * 1024>>>(iters, nw, with_mma, d, sink);
            CK(cudaDeviceSynchronize());
          }
          CK(cudaMemcpy(hh, d, sizeof hh, cudaMemcpyDeviceToHost));
          double mx = 0; for (int w = 0; w < nw; ++w) mx = hh[w] > mx ? hh[w] : mx;
          long long mm[2]; CK(cudaMemcpy(mm, d + 148 * 16, sizeof mm, cudaMemcpyDeviceToHost));
          printf("tmem read: %2d warps x %d x32 loads in flight, mma=%d: %7.1f B/clk/SM", nw, depth, with_mma, (double)iters * depth * 4096.0 * nw / mx);
          if (with_mma) printf("   MMA: %6.1f cycles each meanwhile", (double)mm[0] / (double)mm[1]);
          printf("\n");
        }
  }
  // ---- 3b. readers that also fence / fold, next to the MMAs (8 warps, 4 loads in flight)
  for (int mode : {0, 1, 2, 3}) {
    const int iters = 4000;
    static long long hh2[148 * 16];
    for (int rep = 0; rep < 2; ++rep) { tmem_bw<4><<<148, 544, 201 * 1024>>>(iters, 8, 1, d, sink, mode); CK(cudaDeviceSynchronize()); }
    CK(cudaMemcpy(hh2, d, sizeof hh2, cudaMemcpyDeviceToHost));
    double mx = 0; for (int w = 0; w < 8; ++w) mx = hh2[w] > mx ? hh2[w] : mx;
    long long mm[2]; CK(cudaMemcpy(mm, d + 148 * 16, sizeof mm, cudaMemcpyDeviceToHost));
    printf("8 reader warps, fences=%d fold=%d: %7.1f B/clk/SM read,  MMA %6.1f cycles each meanwhile\n", mode & 1, (mode >> 1) & 1, (double)iters * 4 * 4096.0 * 8 / mx, (double)mm[0] / (double)mm[1]);
  }
  // ---- 3. TMEM read bandwidth (8 warps, 32x32b.x32 pairs)
  for (int rep = 0; rep < 2; ++rep) { tmem_read<<<148, 256>>>(2000, 256, d, sink); CK(cudaDeviceSynchronize()); }
  CK(cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost));
  { double mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const double bytes = 2000.0 * (256 / 64) * 8 /*warps*/ * 2 * 32 * 32 * 4;
    printf("tcgen05.ld 32x32b.x32, 8 warps + 1 max per 2 values: %.1f B/clk/SM (%.0f cycles)\n", bytes / mx, mx); }
  return 0;
}
