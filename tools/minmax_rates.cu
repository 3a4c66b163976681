// Which pipes do the max instructions use on sm_100a, and at what rate?  FMNMX3 / FMNMX / HMNMX2 / F2FP.PACK / VIMNMX3
// alone and in pairs (a pair on different pipes adds up, on the same pipe it does not).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/minmax_rates tools/minmax_rates.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, ILP = 8;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed, long long* cyc) {
  float a[ILP], b[ILP];
  __half2 h[ILP], g[ILP];
  int x[ILP], y[ILP];
  for (int i = 0; i < ILP; ++i) {
    a[i] = seed * (threadIdx.x + i); b[i] = seed + i;
    h[i] = __floats2half2_rn(a[i], b[i]); g[i] = __floats2half2_rn(b[i], a[i]);
    x[i] = (int)a[i]; y[i] = (int)b[i] + i;
  }
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MODE == 0 || MODE == 5 || MODE == 6) a[i] = fmaxf(fmaxf(a[i], b[i]), seed);              // FMNMX3
      if (MODE == 1) a[i] = fmaxf(a[i], b[i] + 0.f * it);                                          // FMNMX (+FADD folded?) 
      if (MODE == 2 || MODE == 5) h[i] = __hmax2(h[i], g[i]);                                      // HMNMX2
      if (MODE == 3 || MODE == 6) g[i] = __floats2half2_rn(b[i], (float)it);                       // F2FP.PACK_AB
      if (MODE == 4) x[i] = max(max(x[i], y[i]), it);                                              // VIMNMX3
      if (MODE == 7) { h[i] = __hmax2(h[i], g[i]); g[i] = __floats2half2_rn(b[i], (float)it); }    // HMNMX2 + F2FP
    }
    if (MODE == 0 || MODE == 5 || MODE == 6) for (int i = 0; i < ILP; ++i) b[i] = a[(i + 1) % ILP];
    if (MODE == 2 || MODE == 5 || MODE == 7) for (int i = 0; i < ILP; ++i) g[i] = h[(i + 1) % ILP];
  }
  const long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < ILP; ++i) s += a[i] + b[i] + __low2float(h[i]) + __high2float(g[i]) + x[i] + y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double ops_per_iter_per_thread) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 148 * 8);
  for (int rep = 0; rep < 2; ++rep) { k<MODE><<<148, 256>>>(out, 1.0001f, cyc); cudaDeviceSynchronize(); }
  long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  double mx = 0; for (auto v : h) mx = v > mx ? v : mx;
  printf("%-34s %8.1f thread-ops/clk/SM  (%.0f cycles)\n", name, ops_per_iter_per_thread * ITERS * 256 / mx, mx);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("FMNMX3 (fmaxf of 3)", ILP);
  run<1>("FMNMX (fmaxf of 2)", ILP);
  run<2>("HMNMX2 (__hmax2)", ILP);
  run<3>("F2FP.PACK_AB (f32,f32 -> f16x2)", ILP);
  run<4>("VIMNMX3 (int max of 3)", ILP);
  run<5>("FMNMX3 + HMNMX2 together", 2 * ILP);
  run<6>("FMNMX3 + F2FP together", 2 * ILP);
  run<7>("HMNMX2 + F2FP together", 2 * ILP);
  return 0;
}
