/*
 * l2_oracle.c - plain-C CPU restatement of the SHIPPED pipeline's matcher (TEST INFRASTRUCTURE ONLY, like lc_oracle.c).
 *
 * Follows /root/reference/src/main.cpp:
 *   - matchFeatures (:509-534): cv::BFMatcher(cv::NORM_L2, false).knnMatch(desc1, desc2, knn, 2); a query is kept iff it
 *     has two neighbours and best.distance < ratio * second.distance (float distances, double ratio);
 *   - the loop search (:1375-1392): for past <= curr - loopGap, skip pairs where either side has < 100 rows, ratio 0.7,
 *     candidate for geometric verification iff matches.size() >= 300.
 * Descriptors are OpenCV SIFT descriptors (main.cpp:497-504): CV_32F rows of 128 values that are integers in [0, 255]
 * (cv::SIFT quantises with saturate_cast<uchar>), so squared distances are exact integers below 2^24, OpenCV's float
 * accumulation is exact whatever its order, and DMatch::distance == sqrtf((float)d2).  knnMatch orders neighbours by
 * (distance, train index).
 *
 * PARITY PIN: python cv2 4.13 BFMatcher(NORM_L2).knnMatch on real cv2.SIFT descriptors (oracle/cv2_l2_ref.py,
 * tests/golden/l2_golden.npz made by tests/golden/make_l2_golden.py) - the reference has no tests of its own.
 * Build: make -C oracle
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct {
  const uint8_t *q, *t;
  int nq, nt;
  int32_t *idx1, *d1, *idx2, *d2;
  int next;
} l2o_job;

static void* l2o_worker(void* v) {
  l2o_job* j = (l2o_job*)v;
  for (;;) {
    const int i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
    if (i >= j->nq) break;
    const uint8_t* a = j->q + (size_t)i * 128;
    int32_t b1 = INT32_MAX, b2 = INT32_MAX, i1 = -1, i2 = -1;
    for (int r = 0; r < j->nt; ++r) {
      const uint8_t* b = j->t + (size_t)r * 128;
      int32_t s = 0;
      for (int k = 0; k < 128; ++k) {
        const int32_t d = (int32_t)a[k] - (int32_t)b[k];
        s += d * d;
      }
      /* ascending scan with strict "<": among equal distances the lower train index ranks first */
      if (s < b1) { b2 = b1; i2 = i1; b1 = s; i1 = r; }
      else if (s < b2) { b2 = s; i2 = r; }
    }
    j->idx1[i] = i1; j->d1[i] = i1 < 0 ? -1 : b1;
    j->idx2[i] = i2; j->d2[i] = i2 < 0 ? -1 : b2;
  }
  return 0;
}

/* knnMatch(k = 2) on u8 rows of 128: best and second (squared distances, -1 / index -1 when there is no such row) */
void l2o_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx1, int32_t* d1, int32_t* idx2, int32_t* d2,
              int nthreads) {
  l2o_job j = {q, t, nq, nt, idx1, d1, idx2, d2, 0};
  pthread_t tid[256];
  int started = 0;
  if (nthreads > 256) nthreads = 256;
  for (int k = 1; k < nthreads; ++k)
    if (pthread_create(&tid[started], 0, l2o_worker, &j) == 0) ++started;
  l2o_worker(&j);
  for (int k = 0; k < started; ++k) pthread_join(tid[k], 0);
}

/* Lowe ratio test of main.cpp:521-532 on squared integer distances; returns the number of good matches */
int l2o_ratio(const int32_t* d1, const int32_t* d2, int nq, double ratio, uint8_t* good) {
  int n = 0;
  for (int i = 0; i < nq; ++i) {
    int g = 0;
    if (d1[i] >= 0 && d2[i] >= 0) {
      const float best = sqrtf((float)d1[i]), second = sqrtf((float)d2[i]);
      g = (double)best < ratio * (double)second;
    }
    if (good) good[i] = (uint8_t)g;
    n += g;
  }
  return n;
}
