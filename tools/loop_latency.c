/* Latency of lc_loop_check through the C ABI from a plain C caller against small keyframe stores (the reference's own
 * regime: <= ~70 eligible keyframes of 2000 descriptors).
 * gcc -O2 -I include tools/loop_latency.c slam_loop_closing_b200/csrc/liblc_b200.so -Wl,-rpath,$PWD/slam_loop_closing_b200/csrc -o build/loop_latency
 * usage: loop_latency [iters] */
#define _POSIX_C_SOURCE 199309L
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include "lc_abi.h"

static double now_us(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}
static int cmp(const void* a, const void* b) { return (*(const double*)a > *(const double*)b) - (*(const double*)a < *(const double*)b); }
static unsigned g_seed = 12345u;
static void fill(uint8_t* p, size_t n) {
  for (size_t i = 0; i < n; ++i) p[i] = (uint8_t)((g_seed = g_seed * 1664525u + 1013904223u) >> 24);
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 200, n = 2000;
  const int Ks[] = {1, 10, 30, 67, 100, 400, 1000};
  uint8_t* q = malloc((size_t)n * 32);
  uint8_t* kf = malloc((size_t)n * 32);
  fill(q, (size_t)n * 32);
  double* us = malloc(sizeof(double) * (size_t)iters);
  for (size_t ki = 0; ki < sizeof Ks / sizeof Ks[0]; ++ki) {
    const int K = Ks[ki];
    lc_config cfg;
    lc_default_config(&cfg);
    cfg.loop_threshold = 0.15;
    cfg.min_loop_gap = 0;
    lc_ctx* c = NULL;
    if (lc_create(&cfg, &c) != 0) { fprintf(stderr, "lc_create: %s\n", lc_last_error(NULL)); return 1; }
    for (int k = 0; k < K; ++k) {
      fill(kf, (size_t)n * 32);
      if (lc_db_append(c, k, kf, n, 32, NULL) != 0) { fprintf(stderr, "%s\n", lc_last_error(c)); return 1; }
    }
    lc_kf_result* res = malloc(sizeof(lc_kf_result) * (size_t)K);
    uint64_t* best = malloc(sizeof(uint64_t) * (size_t)n);
    int64_t nres = 0;
    lc_timing tm;
    for (int i = -10; i < iters; ++i) {
      const double t0 = now_us();
      if (lc_loop_check(c, q, n, 32, K + 30, res, K, &nres, best) != 0) { fprintf(stderr, "%s\n", lc_last_error(c)); return 1; }
      const double t1 = now_us();
      if (i >= 0) us[i] = t1 - t0;
    }
    lc_get_last_timing(c, &tm);
    qsort(us, (size_t)iters, sizeof(double), cmp);
    printf("C caller, loop check 2000 x (%4d keyframes x 2000): median %7.1f us, p10 %7.1f, p90 %7.1f; path %d, %d launches; %.0f Gpairs/s\n", K,
           us[iters / 2], us[iters / 10], us[iters * 9 / 10], tm.path, tm.kernel_launches, (double)K * n * n / us[iters / 2] * 1e-3);
    free(res);
    free(best);
    lc_destroy(c);
  }
  return 0;
}
