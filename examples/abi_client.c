/* abi_client.c - a plain C99 client of liblc_b200.so: what the reference's LoopClosingSystem would do through the C ABI
 * (include/lc_abi.h), with no C++, OpenCV, Python or torch in sight.
 *   gcc -std=c99 -Iinclude examples/abi_client.c -Lslam_loop_closing_b200/csrc -llc_b200 -Wl,-rpath,... -o abi_client
 * Without a usable sm_100 device lc_create fails with a message (there is no CPU fallback) and the program says so. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lc_abi.h"

static void fill(uint8_t* p, size_t n, unsigned seed) {
  for (size_t i = 0; i < n; ++i) {
    seed = seed * 1664525u + 1013904223u;
    p[i] = (uint8_t)(seed >> 24);
  }
}

int main(void) {
  lc_config cfg;
  lc_default_config(&cfg);
  cfg.loop_threshold = 0.15; /* README.md:108 */
  cfg.min_loop_gap = 30;
  printf("liblc_b200 ABI %d, sizeof(lc_config) = %u\n", (int)lc_abi_version(), (unsigned)cfg.struct_size);
  lc_ctx* ctx = NULL;
  if (lc_create(&cfg, &ctx) != LC_OK) {
    printf("lc_create: %s\n", lc_last_error(NULL));
    return 0; /* expected on a machine without a B200 */
  }
  enum { N = 500, K = 64 };
  uint8_t* desc = (uint8_t*)malloc((size_t)N * LC_DESC_BYTES);
  for (int f = 0; f < K; ++f) { /* processFrame: frames_.push_back(frame) -> lc_db_append */
    fill(desc, (size_t)N * LC_DESC_BYTES, 1000u + (unsigned)f);
    if (lc_db_append(ctx, f, desc, N, LC_DESC_BYTES, NULL) != LC_OK) return 1;
  }
  fill(desc, (size_t)N * LC_DESC_BYTES, 1000u + 7u); /* the current frame revisits keyframe 7 */
  lc_loop_candidate cand[K];
  int64_t n = 0;
  if (lc_detect_loops(ctx, desc, N, LC_DESC_BYTES, 100, cand, K, &n) != LC_OK) {
    printf("lc_detect_loops: %s\n", lc_last_error(ctx));
    return 1;
  }
  printf("%lld loop candidates for frame 100\n", (long long)n);
  int found = 0;
  for (int64_t i = 0; i < n; ++i)
    if (cand[i].matched_frame_id == 7 && cand[i].num_matches == N) found = 1;
  printf("keyframe 7 %s\n", found ? "recognised with all matches" : "NOT recognised");
  free(desc);
  lc_destroy(ctx);
  return found ? 0 : 2;
}
