// Shape sweep of the single-launch pair kernel (lc_kernels.cuh: pair_kernel) on one 2000 x 2000 match: for each
// (RQ, threads, rows per CTA) the average time of back-to-back launches (launch overhead pipelined away), with the
// outputs in device memory and in mapped pinned host memory.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/pair_shapes tools/pair_shapes.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../slam_loop_closing_b200/csrc/lc_kernels.cuh"
using namespace lc;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int RQ>
void launch(dim3 grid, int threads, cudaStream_t st, const uint4* q, int nq, const uint4* t, int nt, int seg, uint32_t* keys, unsigned* state,
            int32_t* idx, int32_t* dist, uint8_t* good, int32_t* sum, Knobs kb) {
  pair_kernel<RQ, MODE_CSA><<<grid, threads, 0, st>>>(q, nq, t, nt, seg, keys, state, idx, dist, good, sum, kb);
}

int main(int argc, char** argv) {
  const int nq = argc > 1 ? atoi(argv[1]) : 2000, nt = argc > 2 ? atoi(argv[2]) : 2000;
  std::vector<uint8_t> hq((size_t)nq * 32), ht((size_t)nt * 32);
  unsigned s = 1;
  for (auto& b : hq) b = (uint8_t)((s = s * 1664525u + 1013904223u) >> 24);
  for (auto& b : ht) b = (uint8_t)((s = s * 1664525u + 1013904223u) >> 24);
  uint4 *q, *t;
  uint32_t* keys;
  unsigned* state;
  uint8_t *out_dev, *out_host, *out_host_dev;
  const size_t out_bytes = (size_t)nq * 9 + 8;
  CK(cudaMalloc(&q, hq.size())); CK(cudaMalloc(&t, ht.size())); CK(cudaMalloc(&keys, (size_t)nq * 4)); CK(cudaMalloc(&state, 256));
  CK(cudaMalloc(&out_dev, out_bytes));
  CK(cudaHostAlloc(&out_host, out_bytes, cudaHostAllocMapped)); CK(cudaHostGetDevicePointer(&out_host_dev, out_host, 0));
  CK(cudaMemcpy(q, hq.data(), hq.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(t, ht.data(), ht.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(keys, 0xff, (size_t)nq * 4));
  const unsigned idle[2] = {0u, 0xffffffffu};
  CK(cudaMemcpy(state, idle, 8, cudaMemcpyHostToDevice));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  Knobs kb{30, 0, 0, 0, -1, 1u << kKeyShift};
  int ref_sum[2] = {-2, -2};
  for (int host_out = 0; host_out < 2; ++host_out)
    for (int rq : {8, 4, 2})
      for (int threads : {128, 256, 512}) {
        const int chunks = (nq + threads * rq - 1) / (threads * rq);
        for (int ctas : {74, 148, 296, 592}) {
          const int segs = std::max(1, ctas / chunks);
          const int seg = std::min(256, std::max(2, (nt + segs - 1) / segs));
          const dim3 grid((nt + seg - 1) / seg, chunks);
          uint8_t* o = host_out ? out_host_dev : out_dev;
          int32_t *idx = (int32_t*)o, *dist = (int32_t*)(o + (size_t)nq * 4), *sum = (int32_t*)(o + (size_t)nq * 8);
          uint8_t* good = o + (size_t)nq * 8 + 8;
          const int iters = 50;
          for (int i = -5; i < iters; ++i) {
            if (i == 0) CK(cudaEventRecord(e0, st));
            switch (rq) {
              case 8: launch<8>(grid, threads, st, q, nq, t, nt, seg, keys, state, idx, dist, good, sum, kb); break;
              case 4: launch<4>(grid, threads, st, q, nq, t, nt, seg, keys, state, idx, dist, good, sum, kb); break;
              default: launch<2>(grid, threads, st, q, nq, t, nt, seg, keys, state, idx, dist, good, sum, kb); break;
            }
          }
          CK(cudaEventRecord(e1, st));
          CK(cudaStreamSynchronize(st));
          float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
          int hs[2];
          CK(cudaMemcpy(hs, sum, 8, cudaMemcpyDeviceToHost));
          if (ref_sum[0] == -2) { ref_sum[0] = hs[0]; ref_sum[1] = hs[1]; }
          printf("%s rq %d threads %3d chunks %d seg %3d grid %3u x %d = %4u CTAs: %6.2f us/launch  summary {%d,%d}%s\n", host_out ? "host-out" : "dev-out ",
                 rq, threads, chunks, seg, grid.x, chunks, grid.x * grid.y, ms * 1e3 / iters, hs[0], hs[1],
                 (hs[0] == ref_sum[0] && hs[1] == ref_sum[1]) ? "" : "  MISMATCH");
        }
      }
  return 0;
}
