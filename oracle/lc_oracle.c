/*
 * lc_oracle.c - plain-C CPU restatement of the hot path (see lc_oracle.h: TEST INFRASTRUCTURE ONLY).
 *
 * Follows, step by step, SURVEY.md Appendix A, which restates:
 *   - cv::BFMatcher(NORM_HAMMING)::match semantics behind `matcher_`
 *     (/root/reference/include/loop_closing.hpp:73; README.md:116) - one best train row per
 *     query row, strict "<" update while scanning train rows in ascending order, so the lowest
 *     train index wins ties; no matches when either side is empty;
 *   - LoopClosingSystem::matchFeatures (include/loop_closing.hpp:40; README.md:117) - keep
 *     matches with distance <= 2 * min distance;
 *   - LoopClosingSystem::detectLoops (include/loop_closing.hpp:48; README.md:121-126) - compare
 *     against every stored frame at least min_loop_gap frames back, similarity =
 *     matches / min(n1, n2), candidate iff similarity > threshold and matches >= 50.
 * Build: make -C oracle   (gcc -O3 -mpopcnt -pthread)
 */
#include "lc_oracle.h"

#include <limits.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---- minimal pthread parallel-for with dynamic (atomic counter) scheduling ---- */
typedef void (*lco_body)(int index, void* arg);
typedef struct {
  lco_body body;
  void* arg;
  int n;
  int next;
} lco_pf;
static void* lco_pf_worker(void* v) {
  lco_pf* pf = (lco_pf*)v;
  for (;;) {
    int i = __atomic_fetch_add(&pf->next, 1, __ATOMIC_RELAXED);
    if (i >= pf->n) break;
    pf->body(i, pf->arg);
  }
  return 0;
}
static void lco_parallel_for(int n, int nthreads, lco_body body, void* arg) {
  lco_pf pf = {body, arg, n, 0};
  if (nthreads > n) nthreads = n;
  if (nthreads <= 1) {
    lco_pf_worker(&pf);
    return;
  }
  pthread_t tid[256];
  if (nthreads > 256) nthreads = 256;
  int started = 0;
  for (int t = 1; t < nthreads; ++t)
    if (pthread_create(&tid[started], 0, lco_pf_worker, &pf) == 0) ++started;
  lco_pf_worker(&pf);
  for (int t = 0; t < started; ++t) pthread_join(tid[t], 0);
}

void lco_default_params(lco_params* p) {
  p->loop_threshold = 0.7; /* header default; README/main passes 0.15 */
  p->min_loop_gap = 30;
  p->min_good = 50;
  p->filter_floor = 0;
  p->filter_strict = 0;
  p->min_init = -1;
  p->gap_strict = 0;
  p->thr_inclusive = 0;
}

static inline int hamming256(const uint8_t* a, const uint8_t* b) {
  uint64_t x[4], y[4];
  memcpy(x, a, 32);
  memcpy(y, b, 32);
  return __builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) +
         __builtin_popcountll(x[2] ^ y[2]) + __builtin_popcountll(x[3] ^ y[3]);
}

/* best train row of one query row: ascending scan, strict improvement only */
static inline void best_of_row(const uint8_t* qi, const uint8_t* t, int nt, size_t ts, int32_t* idx,
                               int32_t* dist) {
  int best = INT_MAX, bj = -1;
  for (int j = 0; j < nt; ++j) {
    int d = hamming256(qi, t + (size_t)j * ts);
    if (d < best) {
      best = d;
      bj = j;
    }
  }
  *idx = bj;
  *dist = best;
}

typedef struct {
  const uint8_t *q, *t;
  int nq, nt;
  size_t qs, ts;
  int32_t *idx, *dist;
} match_job;
#define MATCH_CHUNK 64
static void match_body(int c, void* v) {
  match_job* m = (match_job*)v;
  int hi = (c + 1) * MATCH_CHUNK < m->nq ? (c + 1) * MATCH_CHUNK : m->nq;
  for (int i = c * MATCH_CHUNK; i < hi; ++i)
    best_of_row(m->q + (size_t)i * m->qs, m->t, m->nt, m->ts, &m->idx[i], &m->dist[i]);
}

int lco_match(const uint8_t* q, int nq, size_t qs, const uint8_t* t, int nt, size_t ts,
              int32_t* train_idx, int32_t* dist, int nthreads) {
  if (nq <= 0 || nt <= 0) return 0;
  match_job m = {q, t, nq, nt, qs, ts, train_idx, dist};
  lco_parallel_for((nq + MATCH_CHUNK - 1) / MATCH_CHUNK, nthreads, match_body, &m);
  return nq;
}

void lco_filter(const int32_t* dist, int n, const lco_params* p, uint8_t* good, int32_t* min_dist,
                int32_t* n_good) {
  if (n <= 0) {
    *min_dist = -1;
    *n_good = 0;
    return;
  }
  int mn = p->min_init >= 0 ? p->min_init : INT_MAX;
  for (int i = 0; i < n; ++i)
    if (dist[i] < mn) mn = dist[i];
  int thr = 2 * mn;
  if (thr < p->filter_floor) thr = p->filter_floor;
  int cnt = 0;
  for (int i = 0; i < n; ++i) {
    int g = p->filter_strict ? (dist[i] < thr) : (dist[i] <= thr);
    if (good) good[i] = (uint8_t)g;
    cnt += g;
  }
  *min_dist = mn;
  *n_good = cnt;
}

typedef struct {
  const uint8_t *q, *db;
  int nq;
  size_t qs;
  const int64_t* row_off;
  const int32_t *n_kf, *frame_id;
  int cur;
  const lco_params* p;
  lco_kf_result* res;
  uint64_t* best_packed;
} loop_job;

static void loop_body(int k, void* v) {
  loop_job* L = (loop_job*)v;
  const lco_params* p = L->p;
  lco_kf_result* r = &L->res[k];
  r->frame_id = L->frame_id[k];
  r->n_kf = L->n_kf[k];
  r->min_dist = -1;
  r->n_good = -1;
  int gap = L->cur - L->frame_id[k];
  int eligible = p->gap_strict ? (gap > p->min_loop_gap) : (gap >= p->min_loop_gap);
  if (!eligible || L->nq <= 0 || L->n_kf[k] <= 0) return;
  int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)L->nq);
  int32_t* dist = idx + L->nq;
  const uint8_t* t = L->db + (size_t)L->row_off[k] * 32;
  for (int i = 0; i < L->nq; ++i)
    best_of_row(L->q + (size_t)i * L->qs, t, L->n_kf[k], 32, &idx[i], &dist[i]);
  lco_filter(dist, L->nq, p, 0, &r->min_dist, &r->n_good);
  if (L->best_packed) {
    for (int i = 0; i < L->nq; ++i) {
      uint64_t pk = ((uint64_t)(uint32_t)dist[i] << 32) | (uint64_t)(L->row_off[k] + idx[i]);
      /* order-independent integer min (the multi-GPU merge is the same operation) */
      uint64_t old = __atomic_load_n(&L->best_packed[i], __ATOMIC_RELAXED);
      while (pk < old &&
             !__atomic_compare_exchange_n(&L->best_packed[i], &old, pk, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
      }
    }
  }
  free(idx);
}

void lco_loop_check(const uint8_t* q, int nq, size_t qs, const uint8_t* db, const int64_t* row_off,
                    const int32_t* n_kf, const int32_t* frame_id, int K, int cur,
                    const lco_params* p, lco_kf_result* res, uint64_t* best_packed, int nthreads) {
  if (best_packed)
    for (int i = 0; i < nq; ++i) best_packed[i] = UINT64_MAX;
  loop_job L = {q, db, nq, qs, row_off, n_kf, frame_id, cur, p, res, best_packed};
  lco_parallel_for(K, nthreads, loop_body, &L);
}
