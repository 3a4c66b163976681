// tcgen05.mma cta_group::2 (a CTA pair shares one MMA: M = 256, each SM holds its 128 A rows and HALF of the B tile) rate probe,
// dense kind::mxf4 (K = 64) and sparse (.sp, K = 128): does halving the train-tile bytes an SM has to read lift the sparse
// instruction off its shared-memory bound (tools/umma_sp_fp4.cu: 144 cycles at N = 224 on one SM)?  DESIGN.md section 9.1.
// Operand contents are random +-1.0 (timing only).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_cg2_rate tools/umma_cg2_rate.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {  // K-major SWIZZLE_128B, 8-row groups 1024 B apart
  return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t idesc_mxf4(int m, int n, bool sparse) {
  return (sparse ? 4u : 0u) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (1u << 23) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {  // bounded: a pipe that never answers traps
  uint32_t ok = 0;
  for (uint32_t spins = 0; !ok; ++spins) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (spins > (1u << 22)) __trap();
  }
}

constexpr int kSfCol = 448, kMetaCol = 384;

template <bool SPARSE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate2(int N, long long iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  uint32_t rng = 0x9e3779b9u * (blockIdx.x * 128 + threadIdx.x + 1);
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) {  // A: 128 rows x 128 B; this CTA's half of B: up to 128 rows x 256 B
    rng = rng * 1664525u + 1013904223u;
    uint32_t w = 0;
    for (int e = 0; e < 8; ++e) w |= (((rng >> (8 + e)) & 1u) ? 0xAu : 0x2u) << (4 * e);
    ((uint32_t*)smem)[i] = w;
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {  // the same warp of both CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = tslot;
  const int warp = threadIdx.x >> 5;
  const uint32_t lanes = tb + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < 64; ++c) tmem_st1(lanes + kSfCol + c, 0x7f7f7f7fu);      // every scale = 1.0, in both CTAs
  for (int j = 0; j < 16; ++j) tmem_st1(lanes + kMetaCol + j, 0x44444444u);    // metadata: pairs 0, 1 of every chunk
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  cluster_sync();  // both CTAs' operands, scales and metadata are in place
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (rank == 0 && threadIdx.x < 32) {
    const uint32_t idesc = idesc_mxf4(256, N, SPARSE);
    const uint64_t da = smem_desc(smem_u32(smem)), db = smem_desc(smem_u32(smem) + 16384);
    uint32_t el;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
    const long long t0 = clock64();
    for (long long it = 0; it < iters; ++it) {
      if (el) {
        for (int s = 0; s < 4; ++s) {
          const uint32_t ka = (uint32_t)s * 32, kb = SPARSE ? (uint32_t)(s & 1) * 64 : (uint32_t)s * 32;
          const uint32_t acc = (it | s) ? 1u : 0u;
          if (SPARSE)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.sp.cta_group::2.kind::mxf4.block_scale.block32 [%0], %1, %2, [%7], %3, [%5], [%6], p;\n\t}" ::"r"(tb),
                         "l"(da + (ka >> 4)), "l"(db + (kb >> 4)), "r"(idesc), "r"(acc), "r"(tb + kSfCol), "r"(tb + kSfCol + 32), "r"(tb + kMetaCol)
                         : "memory");
          else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(tb),
                         "l"(da + (ka >> 4)), "l"(db + (kb >> 4)), "r"(idesc), "r"(acc), "r"(tb + kSfCol), "r"(tb + kSfCol + 32)
                         : "memory");
        }
      }
      __syncwarp();
    }
    if (el) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
    wait_bar(&bar, 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x >> 1] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  cluster_sync();  // the partner stays until the pair's MMAs are done
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u));
}

int main() {
  long long* d;
  CK(cudaMalloc(&d, 74 * 8));
  long long h[74];
  CK(cudaFuncSetAttribute(rate2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  CK(cudaFuncSetAttribute(rate2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  for (int N : {128, 224, 256}) {
    for (int sp = 0; sp < 2; ++sp) {
      const long long iters = 4096;
      for (int rep = 0; rep < 2; ++rep) {
        if (sp) rate2<true><<<148, 128, 64 * 1024>>>(N, iters, d); else rate2<false><<<148, 128, 64 * 1024>>>(N, iters, d);
        CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(h, d, 74 * 8, cudaMemcpyDeviceToHost));
      double mx = 0;
      for (int i = 0; i < 74; ++i) mx = h[i] > mx ? (double)h[i] : mx;
      printf("cta_group::2 %s mxf4 M=256 N=%3d: %6.1f cycles per MMA (K = %3d positions; per SM: 128 rows of A + %3d rows of B)\n", sp ? "sparse" : "dense ", N,
             mx / (iters * 4), sp ? 128 : 64, N / 2);
    }
  }
  return 0;
}
