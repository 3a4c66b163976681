// tcgen05.mma issue-rate probe: cycles per MMA instruction (M=128, K=32 bytes, SS operands) for several N,
// same-accumulator chains vs alternating accumulators.  Operands are whatever is in shared memory.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_rate tools/umma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
template <int KIND>  // 0 = i8, 1 = f8f6f4 (e4m3)
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5,%5,%5,%5}, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, {%5,%5,%5,%5}, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}

template <int KIND>
__global__ void __launch_bounds__(128, 1) probe(int N, int iters, int n_acc, int commit_every, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x01010101u * (i & 1);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
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
  if (threadIdx.x < 32) {  // whole warp, uniform control flow; one elected lane issues
    // i8: c=S32(2) a=S8(1) b=S8(1); f8f6f4: c=F32(1) a=E4M3(0) b=E4M3(0)
    const uint32_t idesc = (KIND == 0 ? ((2u << 4) | (1u << 7) | (1u << 10)) : (1u << 4)) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint64_t da = smem_desc(smem_u32(smem)), db = smem_desc(smem_u32(smem) + 64 * 1024);
    uint32_t el;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tb + (uint32_t)(it & (n_acc - 1)) * 256;
      if (el) {
#pragma unroll
        for (int s = 0; s < 8; ++s)
          umma<KIND>(d, da + (uint64_t)(((s >> 2) * 16384 + (s & 3) * 32) >> 4), db + (uint64_t)(((s >> 2) * 32768 + (s & 3) * 32) >> 4), idesc, s > 0);
      }
      if (commit_every && (it % commit_every) == commit_every - 1) {
        if (el) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
      }
      __syncwarp();
    }
    if (el) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u));
}

int main() {
  long long* d; CK(cudaMalloc(&d, 148 * 8));
  CK(cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
  CK(cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
  long long h[148];
  printf("kind  N  n_acc commit_every_tiles  grid  cycles/MMA   MAC/clk/SM\n");
  for (int kind = 0; kind < 1; ++kind)
    for (int grid : {148})
      for (int N : {128, 256})
        for (int n_acc : {1, 2})
          for (int ks : {0, 1, 2, 4}) {
            const int iters = 512;
            for (int rep = 0; rep < 2; ++rep) {
              if (kind == 0) probe<0><<<grid, 128, 201 * 1024>>>(N, iters, n_acc, ks, d);
              else probe<1><<<grid, 128, 201 * 1024>>>(N, iters, n_acc, ks, d);
              CK(cudaDeviceSynchronize());
            }
            CK(cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost));
            double mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            double per = mx / (iters * 8);
            printf("%s %4d %3d %5d %6d %10.1f %12.0f\n", kind ? "f8 " : "i8 ", N, n_acc, ks, grid, per, 128.0 * N * 32 / per);
          }
  return 0;
}
