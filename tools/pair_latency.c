/* Latency of lc_match_pair through the C ABI from a plain C caller (no Python in the loop): BASELINE config 2.
 * gcc -O2 -I include tools/pair_latency.c slam_loop_closing_b200/csrc/liblc_b200.so -Wl,-rpath,$PWD/slam_loop_closing_b200/csrc -o build/pair_latency
 * usage: pair_latency [nq] [nt] [iters] [timing: 0 = default (pair calls not event-timed), 1 = timed] */
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

int main(int argc, char** argv) {
  const int nq = argc > 1 ? atoi(argv[1]) : 2000, nt = argc > 2 ? atoi(argv[2]) : 2000, iters = argc > 3 ? atoi(argv[3]) : 500, timing = argc > 4 ? atoi(argv[4]) : 0;
  uint8_t* q = malloc((size_t)nq * 32);
  uint8_t* t = malloc((size_t)nt * 32);
  unsigned s = 12345u;
  for (size_t i = 0; i < (size_t)nq * 32; ++i) q[i] = (uint8_t)((s = s * 1664525u + 1013904223u) >> 24);
  for (size_t i = 0; i < (size_t)nt * 32; ++i) t[i] = (uint8_t)((s = s * 1664525u + 1013904223u) >> 24);
  int32_t* idx = malloc((size_t)nq * 4);
  int32_t* dist = malloc((size_t)nq * 4);
  uint8_t* good = malloc((size_t)nq);
  double* us = malloc(sizeof(double) * (size_t)iters);
  lc_ctx* c = NULL;
  lc_config cfg;
  lc_default_config(&cfg);
  cfg.timing = timing;
  if (lc_create(&cfg, &c) != 0) { fprintf(stderr, "lc_create: %s\n", lc_last_error(NULL)); return 1; }
  int32_t nm, md, ng;
  double h2d = 0, ker = 0, d2h = 0, mk = 0;
  lc_timing tm;
  for (int i = -20; i < iters; ++i) {
    const double t0 = now_us();
    if (lc_match_pair(c, q, nq, 32, t, nt, 32, idx, dist, good, &nm, &md, &ng) != 0) { fprintf(stderr, "%s\n", lc_last_error(c)); return 1; }
    const double t1 = now_us();
    if (i < 0) continue;
    us[i] = t1 - t0;
    lc_get_last_timing(c, &tm);
    h2d += tm.h2d_ms; ker += tm.kernel_ms; d2h += tm.d2h_ms; mk += tm.match_ms;
  }
  qsort(us, (size_t)iters, sizeof(double), cmp);
  printf("C caller, pair %dx%d path %d: median %.1f us, p10 %.1f, p90 %.1f (device means: h2d %.1f us, kernels %.1f us [match kernel %.1f], d2h %.1f us); "
         "min_dist %d n_good %d; %.1f Gpairs/s\n", nq, nt, tm.path, us[iters / 2], us[iters / 10], us[iters * 9 / 10], h2d / iters * 1e3,
         ker / iters * 1e3, mk / iters * 1e3, d2h / iters * 1e3, md, ng, (double)nq * nt / us[iters / 2] * 1e-3);
  lc_destroy(c);
  return 0;
}
