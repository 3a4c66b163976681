// tcgen05.mma.sp kind::mxf4 probe (sparse A, pair-granular 4:8, K = 128 operand positions per instruction) - groundwork for
// DESIGN.md section 9.1 (six descriptor bits per chunk of eight positions: three sparse instead of four dense MMAs per tile).
//   1. layout discovery: A (compressed) = 1.0 in chosen slots, B = one-hot (row n is 1.0 at position n), metadata written
//      with tcgen05.st in a few patterns -> D[m][n] != 0 says which operand position the hardware pairs with which slot;
//   2. rate: cycles per sparse MMA (K = 128) next to the dense one (K = 64), N = 224 / 256, one accumulator.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_sp_fp4 tools/umma_sp_fp4.cu
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
__device__ __forceinline__ void umma_mxf4(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc, uint32_t sfa, uint32_t sfb) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(sfa), "r"(sfb) : "memory");
}
__device__ __forceinline__ void umma_mxf4_sp(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc, uint32_t sfa, uint32_t sfb, uint32_t meta) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.sp.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, [%7], %3, [%5], [%6], p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(sfa), "r"(sfb), "r"(meta) : "memory");
}
// block-scaled descriptor: E2M1 x E2M1, UE8M0 scales, M = 128; bit 2 = sparse, bits [0, 2) = metadata id
__device__ __forceinline__ uint32_t idesc_mxf4(int n, bool sparse, uint32_t id2) {
  return id2 | (sparse ? 4u : 0u) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (1u << 23) | (8u << 24);
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {  // bounded: a pipe that never answers traps
  uint32_t ok = 0;
  for (uint32_t spins = 0; !ok; ++spins) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (spins > (1u << 22)) __trap();
  }
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

constexpr int kSfCol = 448;    // TMEM columns [448, 512): scale cells, all UE8M0 1.0
constexpr int kMetaCol = 384;  // TMEM columns [384, 400): sparse metadata

// the six valid (idx0 < idx1) selections of two pairs out of four, as a nibble idx1 << 2 | idx0
__host__ __device__ inline uint32_t combo(int i) { const uint32_t c[6] = {0x4, 0x8, 0xC, 0x9, 0xD, 0xE}; return c[i % 6]; }

// ---------------------------------------------------------------- 1. layout discovery
// a_mode 0: every compressed slot of A is 1.0; 1: row m has 1.0 in slot (m % 64) only.
// meta_mode 0: nibble 0x4 everywhere; 1: 0xE everywhere; 2: chunk c gets combo(c); 3: lane (row) m gets combo(m) in every chunk;
//           4: column j of the metadata region holds combo(j) everywhere (which columns does an instruction read?)
// (metadata id2 != 0 in the descriptor is an illegal instruction for kind::mxf4: found with this probe)
__global__ void __launch_bounds__(128, 1) discover(int a_mode, int meta_mode, uint32_t id2, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  uint8_t* sA = smem;          // 128 rows x 128 B (the instruction reads the first 32 B = 64 compressed values of every row)
  uint8_t* sB = smem + 16384;  // 128 rows x 128 B (the instruction reads the first 64 B = 128 positions of every row)
  const int tid = threadIdx.x;
  for (int i = tid; i < (16384 + 16384) / 4; i += 128) ((uint32_t*)smem)[i] = 0u;
  __syncthreads();
  {  // A: row = tid
    const int r = tid;
    for (int slot = 0; slot < 64; ++slot) {
      const bool on = a_mode == 0 || slot == (r % 64);
      if (!on) continue;
      const int chunk = slot / 32, byte = (slot % 32) / 2, hi = slot & 1;
      sA[r * 128 + ((chunk ^ (r & 7)) << 4) + byte] |= (uint8_t)(0x2u << (4 * hi));
    }
    // B: row n = tid is 1.0 at position n
    const int n = tid, k = n;
    const int chunk = k / 32, byte = (k % 32) / 2, hi = k & 1;
    sB[n * 128 + ((chunk ^ (n & 7)) << 4) + byte] |= (uint8_t)(0x2u << (4 * hi));
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
  const uint32_t lanes = tb + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < 64; ++c) tmem_st1(lanes + kSfCol + c, 0x7f7f7f7fu);
  for (int c = 0; c < 128; ++c) tmem_st1(lanes + c, 0u);  // the accumulator starts clean (the MMA overwrites it anyway)
  for (int j = 0; j < 16; ++j) {
    uint32_t w = 0;
    for (int e = 0; e < 8; ++e) {
      const int chunk = (j * 8 + e) % 16;
      uint32_t nib = 0x4;
      if (meta_mode == 1) nib = 0xE;
      if (meta_mode == 2) nib = combo(chunk);
      if (meta_mode == 3) nib = combo(tid);
      if (meta_mode == 4) nib = combo(j);
      w |= nib << (4 * e);
    }
    tmem_st1(lanes + kMetaCol + j, w);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (tid == 0) {
    umma_mxf4_sp(tb, smem_desc(smem_u32(sA)), smem_desc(smem_u32(sB)), idesc_mxf4(128, true, id2), 0u, tb + kSfCol, tb + kSfCol + 32, tb + kMetaCol);
    commit(&bar);
  }
  __syncthreads();
  wait_bar(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;");
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(lanes + c0, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[(size_t)tid * 128 + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u));
}

// ---------------------------------------------------------------- 2. rate
template <bool SPARSE>
__global__ void __launch_bounds__(128, 1) rate(int N, long long iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  uint32_t rng = 0x9e3779b9u * (blockIdx.x * 128 + threadIdx.x + 1);
  for (int i = threadIdx.x; i < (16384 + 65536) / 4; i += blockDim.x) {  // A 16 KiB, B 256 rows x 256 B
    rng = rng * 1664525u + 1013904223u;
    uint32_t w = 0;
    for (int e = 0; e < 8; ++e) w |= (((rng >> (8 + e)) & 1u) ? 0xAu : 0x2u) << (4 * e);  // -1.0 / +1.0
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
  const uint32_t lanes = tb + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < 64; ++c) tmem_st1(lanes + kSfCol + c, 0x7f7f7f7fu);
  for (int j = 0; j < 16; ++j) tmem_st1(lanes + kMetaCol + j, 0x44444444u);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (threadIdx.x < 32) {
    const uint32_t idesc = idesc_mxf4(N, SPARSE, 0);
    const uint64_t da = smem_desc(smem_u32(smem)), db = smem_desc(smem_u32(smem) + 16384);
    uint32_t el;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
    const long long t0 = clock64();
    for (long long it = 0; it < iters; ++it) {
      if (el) {
        for (int s = 0; s < 4; ++s) {  // A advances 32 B per step; B 32 B (dense) or 64 B (sparse, two steps per 128-byte row)
          const uint32_t ka = (uint32_t)s * 32, kb = SPARSE ? (uint32_t)(s & 1) * 64 : (uint32_t)s * 32;
          if (SPARSE) umma_mxf4_sp(tb, da + (ka >> 4), db + (kb >> 4), idesc, (it | s) ? 1u : 0u, tb + kSfCol, tb + kSfCol + 32, tb + kMetaCol);
          else umma_mxf4(tb, da + (ka >> 4), db + (kb >> 4), idesc, (it | s) ? 1u : 0u, tb + kSfCol, tb + kSfCol + 32);
        }
      }
      __syncwarp();
    }
    if (el) commit(&bar);
    __syncwarp();
    wait_bar(&bar, 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u));
}

static void show(const char* what, const std::vector<float>& o, std::initializer_list<int> rows) {
  printf("%s\n", what);
  for (int r : rows) {
    printf("  row %3d: ", r);
    int shown = 0;
    for (int n = 0; n < 128; ++n) {
      const float v = o[(size_t)r * 128 + n];
      if (v != 0.f && shown++ < 40) printf("%d%s ", n, v == 1.f ? "" : (v == 2.f ? "(x2)" : "(?)"));
    }
    printf("| %d non-zero\n", shown);
  }
}

int main(int argc, char** argv) {
  const int what = argc > 1 ? atoi(argv[1]) : 3;
  CK(cudaFuncSetAttribute(discover, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  float* d_out;
  CK(cudaMalloc(&d_out, 128 * 128 * 4));
  std::vector<float> o(128 * 128);
  if (what & 1) {
    struct Case { int a_mode, meta_mode; uint32_t id2; const char* name; };
    const Case cases[] = {
        {0, 0, 0, "A all ones, metadata nibble 0x4 (pairs 0,1) everywhere"},
        {0, 1, 0, "A all ones, metadata nibble 0xE (pairs 2,3) everywhere"},
        {0, 2, 0, "A all ones, chunk c uses combo(c) = 4 8 C 9 D E 4 8 ..."},
        {0, 3, 0, "A all ones, lane m uses combo(m)"},
        {0, 4, 0, "A all ones, metadata column j holds combo(j): which columns are read (id2 = 0)"},
        {1, 0, 0, "A: row m has slot m % 64 only, nibble 0x4 everywhere (slot -> position)"},
        {1, 1, 0, "A: row m has slot m % 64 only, nibble 0xE everywhere"},
    };
    for (const Case& c : cases) {
      CK(cudaMemset(d_out, 0, 128 * 128 * 4));
      discover<<<1, 128, 64 * 1024>>>(c.a_mode, c.meta_mode, c.id2, d_out);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(o.data(), d_out, 128 * 128 * 4, cudaMemcpyDeviceToHost));
      show(c.name, o, {0, 1, 2, 5, 33, 64, 127});
    }
  }
  if (what & 2) {
    long long* d;
    CK(cudaMalloc(&d, 148 * 8));
    long long h[148];
    CK(cudaFuncSetAttribute(rate<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(rate<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    for (int N : {128, 224, 256}) {
      for (int sp = 0; sp < 2; ++sp) {
        const long long iters = 4096;
        for (int rep = 0; rep < 2; ++rep) {
          if (sp) rate<true><<<148, 128, 100 * 1024>>>(N, iters, d); else rate<false><<<148, 128, 100 * 1024>>>(N, iters, d);
          CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost));
        double mx = 0;
        for (int i = 0; i < 148; ++i) mx = h[i] > mx ? (double)h[i] : mx;
        const double per = mx / (iters * 4);
        printf("%s mxf4 N=%3d: %6.1f cycles per MMA (K = %3d operand positions): %6.0f positions x rows x columns / clk / SM\n", sp ? "sparse" : "dense ", N, per,
               sp ? 128 : 64, 128.0 * N * (sp ? 128 : 64) / per);
      }
    }
  }
  return 0;
}
