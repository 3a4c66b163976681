// Probe for a packed 16-bit epilogue of the kind::mxf4 Hamming GEMM:
//   1. pre-fill the f32 accumulator with C = 1.5 * 2^23 (bits 0x4B400000), accumulate the four K = 64 MMAs ON TOP of it:
//      is the sum still exact, i.e. are the bits 0x4B400000 + D with D = 256 - 2 * hamming (so that the low 16 bits of every
//      accumulator cell hold D as a signed 16-bit integer)?
//   2. does tcgen05.ld .pack::16b return the LOW halves of two adjacent columns in one register?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/umma_bias_pack tools/umma_bias_pack.cu
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
__device__ __forceinline__ uint32_t idesc_mxf4(int n) { return (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (1u << 23) | (8u << 24); }
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  for (int spins = 0; !ok && spins < (1 << 22); ++spins)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
#define R16(v, o) "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7]), \
                  "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]), "=r"(v[o + 12]), "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15])
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : R16(v, 0) : "r"(taddr) : "memory");
}
// .pack::16b: the same 16 columns... or 32?  Found out by looking at what comes back.
__device__ __forceinline__ void tmem_ld16_pack(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : R16(v, 0) : "r"(taddr) : "memory");
}
__host__ __device__ __forceinline__ uint32_t expand8_fp4(uint32_t x) {
  x = (x | (x << 12)) & 0x000F000Fu;
  x = (x | (x << 6)) & 0x03030303u;
  x = (x | (x << 3)) & 0x11111111u;
  return x * 8u + 0x22222222u;
}
constexpr int kSfCol = 448;
constexpr uint32_t kBias = 0x4B400000u;  // 1.5 * 2^23: ulp 1, and room for negative sums below it

__global__ void __launch_bounds__(128, 1) probe(const uint32_t* qbits, const uint32_t* tbits, int N, uint32_t bias, uint32_t* raw_out, uint32_t* pack_out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 16384;
  const int tid = threadIdx.x, warp = tid >> 5;
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
  const uint32_t lanes = tb + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < 64; ++c) tmem_st1(lanes + kSfCol + c, 0x7f7f7f7fu);  // all block scales 1.0
  for (int c = 0; c < N; ++c) tmem_st1(lanes + c, bias);                   // the accumulator starts at the bias
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (tid == 0) {
    const uint64_t da = smem_desc(smem_u32(sA)), db = smem_desc(smem_u32(sB));
    for (int s = 0; s < 4; ++s)
      umma_mxf4(tb, da + (uint64_t)((s * 32) >> 4), db + (uint64_t)((s * 32) >> 4), idesc_mxf4(N), 1u, tb + kSfCol, tb + kSfCol + 32);
    commit(&bar);
  }
  __syncthreads();
  wait_bar(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;");
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(lanes + c0, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) raw_out[(size_t)tid * N + c0 + j] = v[j];
    tmem_ld16_pack(lanes + c0, v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) pack_out[(size_t)tid * N + c0 + j] = v[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u));
}

int main() {
  const int N = 224;
  std::vector<uint32_t> q(128 * 8), t(N * 8);
  unsigned s = 7;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s; };
  for (auto& x : q) x = rnd() ^ (rnd() >> 7);
  for (auto& x : t) x = rnd() ^ (rnd() >> 7);
  for (int w = 0; w < 8; ++w) { t[5 * 8 + w] = q[3 * 8 + w]; t[9 * 8 + w] = ~q[4 * 8 + w]; }  // distance 0 and 256
  for (int w = 0; w < 8; ++w) t[17 * 8 + w] = q[11 * 8 + w] ^ (w < 5 ? 0xffffffffu : 0u);     // distance 160: a negative sum
  uint32_t *dq, *dt, *draw, *dpack;
  CK(cudaMalloc(&dq, q.size() * 4)); CK(cudaMalloc(&dt, t.size() * 4));
  CK(cudaMalloc(&draw, 128 * N * 4)); CK(cudaMalloc(&dpack, 128 * N * 4));
  CK(cudaMemcpy(dq, q.data(), q.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dt, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  for (uint32_t bias : {0u, kBias, 0x4B000000u}) {
    CK(cudaMemset(draw, 0, 128 * N * 4)); CK(cudaMemset(dpack, 0, 128 * N * 4));
    probe<<<1, 128, 64 * 1024>>>(dq, dt, N, bias, draw, dpack);
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> raw(128 * N), pack(128 * N);
    CK(cudaMemcpy(raw.data(), draw, raw.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(pack.data(), dpack, pack.size() * 4, cudaMemcpyDeviceToHost));
    long bad_raw = 0, bad_lo = 0, neg = 0;
    for (int r = 0; r < 128; ++r)
      for (int c = 0; c < N; ++c) {
        int ham = 0;
        for (int w = 0; w < 8; ++w) ham += __builtin_popcount(q[r * 8 + w] ^ t[c * 8 + w]);
        const int D = 256 - 2 * ham;
        neg += D < 0;
        uint32_t want;
        if (bias == 0) { float f = (float)D; memcpy(&want, &f, 4); } else want = bias + (uint32_t)D;
        bad_raw += raw[r * N + c] != want;
        bad_lo += (raw[r * N + c] & 0xffffu) != ((uint32_t)D & 0xffffu);
      }
    printf("bias 0x%08x: %ld of %d accumulator cells differ from the exact value (%ld negative sums); low-16 != D in %ld cells\n", bias, bad_raw,
           128 * N, neg, bad_lo);
    // what does .pack::16b return?  print row 0: raw columns 0..7 next to packed registers 0..7 of the first two 16-column loads
    printf("  row 0 raw  c0..7 : "); for (int c = 0; c < 8; ++c) printf("%08x ", raw[c]); printf("\n");
    printf("  row 0 pack r0..7 : "); for (int c = 0; c < 8; ++c) printf("%08x ", pack[c]); printf("\n");
    printf("  row 0 raw  c16..23: "); for (int c = 16; c < 24; ++c) printf("%08x ", raw[c]); printf("\n");
    printf("  row 0 pack r16..23: "); for (int c = 16; c < 24; ++c) printf("%08x ", pack[c]); printf("\n");
  }
  return 0;
}
