// Pipe-rate microbenchmarks for sm_100a: defines the integer-popc roofline denominator
// (SURVEY.md §8d asks for POPC/clk/SM to be measured, not assumed).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/pipe_rates tools/pipe_rates.cu
// Run:   tools/pipe_rates > gpurun_out/pipe_rates.json
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;
constexpr int ILP = 8;

// ---- POPC only: ILP independent chains x = popc(x) ^ k (1 POPC + 1 LOP3 per step; LOP3 is 4x cheaper)
__global__ void k_popc(uint32_t* out, uint32_t seed, long long* cyc) {
  uint32_t x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = seed + threadIdx.x * 7 + i;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = __popc(x[i]) ^ seed;
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- LOP3 only
__global__ void k_lop3(uint32_t* out, uint32_t seed, long long* cyc) {
  uint32_t x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = seed + threadIdx.x * 7 + i;
  uint32_t a = seed * 3 + threadIdx.x, b = seed * 5 + 1;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      // maj(x,a,b) xor-ish 3-input function, not reducible
      asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(x[i]) : "r"(a), "r"(b));
      asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(b), "r"(a));
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- IMAD only
__global__ void k_imad(uint32_t* out, uint32_t seed, long long* cyc) {
  uint32_t x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = seed + threadIdx.x * 7 + i;
  uint32_t a = seed * 3 + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(seed));
      asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(seed), "r"(a));
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- VIMNMX / min only
__global__ void k_min(uint32_t* out, uint32_t seed, long long* cyc) {
  uint32_t x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = seed + threadIdx.x * 7 + i;
  uint32_t a = seed * 3 + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      asm volatile("min.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
      asm volatile("max.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(seed));
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- register-resident Hamming inner loops: RQ queries per thread vs a rotating train descriptor
template <int RQ, int MODE>
__global__ void k_hamming(uint32_t* out, uint32_t seed, long long* cyc) {
  uint32_t q[RQ][8];
  uint32_t best[RQ];
#pragma unroll
  for (int r = 0; r < RQ; ++r) {
    best[r] = 0xffffffffu;
#pragma unroll
    for (int w = 0; w < 8; ++w) q[r][w] = (seed + threadIdx.x * 977u + r * 131u + w) * 2654435761u;
  }
  uint32_t t[8];
#pragma unroll
  for (int w = 0; w < 8; ++w) t[w] = seed * (w + 3);
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int w = 0; w < 8; ++w) t[w] = t[w] * 1664525u + 1013904223u;  // 8 IMAD per train, amortised over RQ
#pragma unroll
    for (int r = 0; r < RQ; ++r) {
      uint32_t key;
      if (MODE == 0) {  // naive: 8 xor + 8 popc + adds
        uint32_t d = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) d += __popc(q[r][w] ^ t[w]);
        key = (d << 20) | (uint32_t)it;
      } else {  // carry-save: 7 words -> ones/twos/fours (4 full adders), + word 7 : 4 popc
        uint32_t x0 = q[r][0] ^ t[0], x1 = q[r][1] ^ t[1], x2 = q[r][2] ^ t[2], x3 = q[r][3] ^ t[3];
        uint32_t x4 = q[r][4] ^ t[4], x5 = q[r][5] ^ t[5], x6 = q[r][6] ^ t[6], x7 = q[r][7] ^ t[7];
        uint32_t s1 = x0 ^ x1 ^ x2, c1 = (x0 & x1) | (x2 & (x0 | x1));
        uint32_t s2 = x3 ^ x4 ^ x5, c2 = (x3 & x4) | (x5 & (x3 | x4));
        uint32_t ones = s1 ^ s2 ^ x6, c3 = (s1 & s2) | (x6 & (s1 | s2));
        uint32_t twos = c1 ^ c2 ^ c3, fours = (c1 & c2) | (c3 & (c1 | c2));
        uint32_t d = __popc(ones) + __popc(x7) + 2 * __popc(twos) + 4 * __popc(fours);
        key = (d << 20) | (uint32_t)it;
      }
      best[r] = min(best[r], key);
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int r = 0; r < RQ; ++r) s ^= best[r];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- legacy int8 tensor MMA (mma.sync m16n8k32 u8.u8 s32): rate of the only int8 path mma.sync offers
constexpr int MMA_ITERS = 2048;
template <int NACC>
__global__ void k_imma(uint32_t* out, uint32_t seed, long long* cyc) {
  int c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0; }
  uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  long long t0 = clock64();
  for (int it = 0; it < MMA_ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      asm volatile(
          "mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
          : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
          : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
  long long t1 = clock64();
  int s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- 1-bit mma (and.popc m16n8k256): ptxas emulates it on sm_100a - measure anyway
template <int NACC>
__global__ void k_bmma(uint32_t* out, uint32_t seed, long long* cyc) {
  int c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0; }
  uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  long long t0 = clock64();
  for (int it = 0; it < MMA_ITERS / 4; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      asm volatile(
          "mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
          : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
          : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
  long long t1 = clock64();
  int s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

struct Res { double ms; double cyc_max; };

template <typename F>
static int run(const char* name, F launch, int blocks, int threads, double ops_per_thread, const char* unit, int nsm, bool last = false) {
  uint32_t* out; long long* cyc;
  CK(cudaMalloc(&out, sizeof(uint32_t) * blocks * threads));
  CK(cudaMalloc(&cyc, sizeof(long long) * blocks));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; ++w) launch(out, cyc);
  CK(cudaDeviceSynchronize());
  float best_ms = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0));
    launch(out, cyc);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best_ms) best_ms = ms;
  }
  CK(cudaGetLastError());
  long long* h = (long long*)malloc(sizeof(long long) * blocks);
  CK(cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
  double cmax = 0, csum = 0;
  for (int i = 0; i < blocks; ++i) { if (h[i] > cmax) cmax = (double)h[i]; csum += (double)h[i]; }
  free(h);
  double total_ops = ops_per_thread * (double)blocks * threads;
  // per-SM per-clock rate using the mean in-kernel cycle count of one block and blocks-per-SM co-residency
  double blocks_per_sm = (double)blocks / nsm;
  double per_clk_sm = ops_per_thread * threads * blocks_per_sm / (csum / blocks);
  printf("  {\"name\": \"%s\", \"blocks\": %d, \"threads\": %d, \"ms\": %.5f, \"cycles_mean\": %.0f, \"cycles_max\": %.0f, "
         "\"%s_per_clk_per_sm\": %.3f, \"G%s_per_s\": %.2f, \"eff_mhz\": %.0f}%s\n",
         name, blocks, threads, best_ms, csum / blocks, cmax, unit, per_clk_sm, unit, total_ops / (best_ms * 1e6),
         (csum / blocks) / (best_ms * 1e3), last ? "" : ",");
  cudaFree(out); cudaFree(cyc);
  return 0;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"results\": [\n", p.name, nsm, p.clockRate);
  const int T = 256;
  for (int bps : {1, 2, 4}) {
    int B = nsm * bps;
    char nm[64];
    snprintf(nm, sizeof nm, "popc_bps%d", bps);
    run(nm, [&](uint32_t* o, long long* c) { k_popc<<<B, T>>>(o, 12345u, c); }, B, T, (double)ITERS * ILP, "popc", nsm);
    snprintf(nm, sizeof nm, "lop3_bps%d", bps);
    run(nm, [&](uint32_t* o, long long* c) { k_lop3<<<B, T>>>(o, 12345u, c); }, B, T, (double)ITERS * ILP * 2, "lop3", nsm);
    snprintf(nm, sizeof nm, "imad_bps%d", bps);
    run(nm, [&](uint32_t* o, long long* c) { k_imad<<<B, T>>>(o, 12345u, c); }, B, T, (double)ITERS * ILP * 2, "imad", nsm);
    snprintf(nm, sizeof nm, "minmax_bps%d", bps);
    run(nm, [&](uint32_t* o, long long* c) { k_min<<<B, T>>>(o, 12345u, c); }, B, T, (double)ITERS * ILP * 2, "minmax", nsm);
  }
  for (int bps : {1, 2}) {
    int B = nsm * bps;
    char nm[64];
    snprintf(nm, sizeof nm, "hamming_naive_rq4_bps%d", bps);
    run(nm, [&](uint32_t* o, long long* c) { k_hamming<4, 0><<<B, T>>>(o, 12345u, c); }, B, T, (double)ITERS * 4, "pair", nsm);
    snprintf(nm, sizeof nm, "hamming_naive_rq8_bps%d", bps);
    run(nm, [&](uint32_t* o, long long* c) { k_hamming<8, 0><<<B, T>>>(o, 12345u, c); }, B, T, (double)ITERS * 8, "pair", nsm);
    snprintf(nm, sizeof nm, "hamming_csa_rq4_bps%d", bps);
    run(nm, [&](uint32_t* o, long long* c) { k_hamming<4, 1><<<B, T>>>(o, 12345u, c); }, B, T, (double)ITERS * 4, "pair", nsm);
    snprintf(nm, sizeof nm, "hamming_csa_rq8_bps%d", bps);
    run(nm, [&](uint32_t* o, long long* c) { k_hamming<8, 1><<<B, T>>>(o, 12345u, c); }, B, T, (double)ITERS * 8, "pair", nsm);
  }
  for (int bps : {1, 2, 4}) {
    int B = nsm * bps;
    char nm[64];
    // ops = int8 MACs: one m16n8k32 = 4096 MACs per warp => 128 per thread
    snprintf(nm, sizeof nm, "imma_u8_m16n8k32_acc8_bps%d", bps);
    run(nm, [&](uint32_t* o, long long* c) { k_imma<8><<<B, T>>>(o, 12345u, c); }, B, T, (double)MMA_ITERS * 8 * 128, "mac", nsm);
    // one m16n8k256 b1 = 128 descriptor pairs per warp => 4 per thread
    snprintf(nm, sizeof nm, "bmma_b1_m16n8k256_acc4_bps%d", bps);
    run(nm, [&](uint32_t* o, long long* c) { k_bmma<4><<<B, T>>>(o, 12345u, c); }, B, T, (double)(MMA_ITERS / 4) * 4 * 4, "pair", nsm, bps == 4);
  }
  printf("]}\n");
  return 0;
}
