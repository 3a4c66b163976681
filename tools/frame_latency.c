/* Per-frame latency of processFrame's matcher side through the C ABI from a plain C caller, in the reference's own
 * regime (README.md:150-165: ~97 frames of 2000 ORB descriptors, min_loop_gap 30, threshold 0.15):
 *   separate: lc_match_pair(prev, cur) + lc_db_append(cur) + lc_detect_loops(cur)   (three uploads, three waits)
 *   fused:    lc_process_frame(cur)                                                 (one upload, one wait)
 * gcc -O2 -I include tools/frame_latency.c slam_loop_closing_b200/csrc/liblc_b200.so -Wl,-rpath,$PWD/slam_loop_closing_b200/csrc -o build/frame_latency
 * usage: frame_latency [frames] [rows] [timing: 1 = also print the device-side split of the fused call (event-timed run)] */
#define _POSIX_C_SOURCE 199309L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "lc_abi.h"

static double now_us(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}
static int cmp(const void* a, const void* b) { return (*(const double*)a > *(const double*)b) - (*(const double*)a < *(const double*)b); }
static double median(double* v, int n) {
  qsort(v, (size_t)n, sizeof(double), cmp);
  return v[n / 2];
}

int main(int argc, char** argv) {
  const int F = argc > 1 ? atoi(argv[1]) : 97, n = argc > 2 ? atoi(argv[2]) : 2000, timed = argc > 3 ? atoi(argv[3]) : 0;
  uint8_t* frames = malloc((size_t)F * n * 32);
  unsigned s = 4242u;
  for (size_t i = 0; i < (size_t)F * n * 32; ++i) frames[i] = (uint8_t)((s = s * 1664525u + 1013904223u) >> 24);
  /* the walk returns: the last third of the frames are copies of early ones with every 8th row replaced */
  for (int f = 2 * F / 3; f < F; ++f) {
    memcpy(frames + (size_t)f * n * 32, frames + (size_t)(f - 2 * F / 3) * n * 32, (size_t)n * 32);
    for (int r = 0; r < n; r += 8) frames[((size_t)f * n + r) * 32] ^= 0xff;
  }
  int32_t* idx = malloc((size_t)n * 4);
  int32_t* dist = malloc((size_t)n * 4);
  uint8_t* good = malloc((size_t)n);
  lc_loop_candidate* out = malloc(sizeof(lc_loop_candidate) * (size_t)F);
  double* us = malloc(sizeof(double) * (size_t)F);
  long loops[2] = {0, 0};
  long good_sum[2] = {0, 0};
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) { /* rep 0 warms up (buffer growth, first launches) */
      lc_config cfg;
      lc_default_config(&cfg);
      cfg.loop_threshold = 0.15;
      lc_ctx* c = NULL;
      if (lc_create(&cfg, &c) != 0) { fprintf(stderr, "lc_create: %s\n", lc_last_error(NULL)); return 1; }
      loops[mode] = good_sum[mode] = 0;
      for (int f = 0; f < F; ++f) {
        const uint8_t* cur = frames + (size_t)f * n * 32;
        int32_t nm = 0, md = 0, ng = 0;
        int64_t n_out = 0;
        const double t0 = now_us();
        int rc = 0;
        if (mode == 0) {
          if (f) rc |= lc_match_pair(c, cur - (size_t)n * 32, n, 32, cur, n, 32, idx, dist, good, &nm, &md, &ng);
          rc |= lc_db_append(c, f, cur, n, 32, NULL);
          rc |= lc_detect_loops(c, cur, n, 32, f, out, F, &n_out);
        } else {
          rc = lc_process_frame(c, f, cur, n, 32, idx, dist, good, n, &nm, &md, &ng, out, F, &n_out, NULL);
        }
        us[f] = now_us() - t0;
        if (rc != 0) { fprintf(stderr, "frame %d: %s\n", f, lc_last_error(c)); return 1; }
        loops[mode] += (long)n_out;
        good_sum[mode] += ng;
      }
      if (rep == 1) {
        double tot = 0;
        for (int f = 0; f < F; ++f) tot += us[f];
        const int tail = F / 3;
        printf("%-8s %d frames x %d rows: total %.2f ms, median per frame over the last %d frames (%d-%d eligible keyframes) %.1f us; "
               "%ld loop candidates, %ld good consecutive matches\n", mode ? "fused" : "separate", F, n, tot * 1e-3, tail, F - tail - 30, F - 31,
               median(us + F - tail, tail), loops[mode], good_sum[mode]);
      }
      lc_destroy(c);
    }
  }
  if (loops[0] != loops[1] || good_sum[0] != good_sum[1]) { fprintf(stderr, "MISMATCH between the two modes\n"); return 2; }
  if (timed) { /* the fused call again with CUDA events on (they cost ~20 us per call, so only the split is meaningful) */
    lc_config cfg;
    lc_default_config(&cfg);
    cfg.loop_threshold = 0.15;
    cfg.timing = LC_TIMING_ON;
    lc_ctx* c = NULL;
    if (lc_create(&cfg, &c) != 0) return 1;
    double h2d = 0, ker = 0, d2h = 0, tot = 0, mk = 0;
    int cnt = 0, launches = 0;
    for (int f = 0; f < F; ++f) {
      int32_t nm, md, ng;
      int64_t n_out;
      if (lc_process_frame(c, f, frames + (size_t)f * n * 32, n, 32, idx, dist, good, n, &nm, &md, &ng, out, F, &n_out, NULL) != 0) return 1;
      if (f >= F - F / 3) {
        lc_timing tm;
        lc_get_last_timing(c, &tm);
        h2d += tm.h2d_ms; ker += tm.kernel_ms; d2h += tm.d2h_ms; tot += tm.total_ms; mk += tm.match_ms; launches = tm.kernel_launches;
        ++cnt;
      }
    }
    printf("fused, device side (means over the last %d frames): upload + table poke %.1f us, kernels %.1f us (loop-check matching kernel %.1f us), "
           "download %.1f us, first to last %.1f us; %d launches\n", cnt, h2d / cnt * 1e3, ker / cnt * 1e3, mk / cnt * 1e3, d2h / cnt * 1e3, tot / cnt * 1e3, launches);
    lc_destroy(c);
  }
  return 0;
}
