// C++ parity test of the host side in the reference's own language: loop_closing::b200::LoopClosingSystem
// (include/loop_closing_system_b200.hpp) drives a descriptor sequence through the C ABI on the GPU - one submission per frame
// (fused) and as three separate calls - and every result is compared with the C oracle (oracle/lc_oracle.h, the checker):
// the good-match count of each matchFeatures(previous, current) and the LoopCandidate list (ids, match counts, similarity
// doubles bit for bit) of each detectLoops.  Built and run by tests/test_cpp_system.py; OpenCV's types come from tests/stubs.
//   exit 0 + "PASS ..."                on a B200
//   exit 0 + "... no CPU fallback ..." where lc_create finds no usable device
//   exit 1                             on any mismatch
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "loop_closing_system_b200.hpp"
#include "lc_oracle.h"

using loop_closing::b200::LoopCandidate;
using loop_closing::b200::LoopClosingSystem;

static uint32_t g_rng = 20261018u;
static uint32_t rnd() { return g_rng = g_rng * 1664525u + 1013904223u; }

// a walk that returns to its first places: frame i >= revisit is a noisy copy of frame i - revisit (bit flips, a third of the
// rows replaced, rows rotated); consecutive frames share about half of their rows
static std::vector<cv::Mat> make_sequence(int n_frames, int revisit) {
  std::vector<cv::Mat> frames;
  for (int i = 0; i < n_frames; ++i) {
    const int n = i == 7 ? 0 : 500 + (int)(rnd() % 700);  // one empty frame: no matches, no loop, not an error
    cv::Mat d(n, 32, CV_8UC1);
    const cv::Mat* src = nullptr;
    if (i >= revisit && frames[i - revisit].rows) src = &frames[i - revisit];
    else if (i && frames[i - 1].rows && i % 3) src = &frames[i - 1];
    for (int r = 0; r < n; ++r) {
      unsigned char* row = d.ptr<unsigned char>(r);
      const bool fresh = !src || rnd() % 100 < (i >= revisit ? 30u : 50u);
      if (fresh) {
        for (int b = 0; b < 32; ++b) row[b] = (unsigned char)(rnd() >> 24);
      } else {
        std::memcpy(row, src->ptr<unsigned char>((r * 7 + 3) % src->rows), 32);
        for (int f = 0; f < 10; ++f)
          if (rnd() % 2) row[rnd() % 32] ^= (unsigned char)(1u << (rnd() % 8));
      }
    }
    frames.push_back(d);
  }
  return frames;
}

static int fail(const char* what, int frame) {
  std::fprintf(stderr, "MISMATCH: %s at frame %d\n", what, frame);
  return 1;
}

int main() {
  const int F = 64, revisit = 38;
  const double thr = 0.15;
  const int gap = 30;
  const std::vector<cv::Mat> frames = make_sequence(F, revisit);
  try {
    LoopClosingSystem fused(thr, gap, true), plain(thr, gap, false);
    lco_params p;
    lco_default_params(&p);
    p.loop_threshold = thr;
    p.min_loop_gap = gap;
    std::vector<unsigned char> db;
    std::vector<int64_t> row_off;
    std::vector<int32_t> n_kf, ids;
    size_t loops_seen = 0;
    long total_good = 0;
    for (int i = 0; i < F; ++i) {
      const int id = 100 + 3 * i;  // frame ids are not store indices (frame_skip = 3, README.md:110)
      const cv::Mat& d = frames[i];
      fused.processDescriptors(d, id);
      plain.processDescriptors(d, id);
      // ---- the oracle's answer for this frame
      if (i) {
        const cv::Mat& q = frames[i - 1];
        std::vector<int32_t> idx(q.rows), dist(q.rows);
        std::vector<uint8_t> good(q.rows);
        int32_t min_dist = -1, n_good = 0;
        if (lco_match(q.data, q.rows, 32, d.data, d.rows, 32, idx.data(), dist.data(), 4) > 0) lco_filter(dist.data(), q.rows, &p, good.data(), &min_dist, &n_good);
        if (fused.consecutiveMatchCounts().back() != n_good || plain.consecutiveMatchCounts().back() != n_good) return fail("matchFeatures(prev, cur) good-match count", i);
        total_good += n_good;
      }
      row_off.push_back((int64_t)(db.size() / 32));
      n_kf.push_back(d.rows);
      ids.push_back(id);
      if (d.rows) db.insert(db.end(), d.data, d.data + (size_t)d.rows * 32);
      std::vector<lco_kf_result> rec(i + 1);
      lco_loop_check(d.data, d.rows, 32, db.data(), row_off.data(), n_kf.data(), ids.data(), i + 1, id, &p, rec.data(), nullptr, 4);
      std::vector<LoopCandidate> want;
      for (const auto& r : rec) {
        const int denom = d.rows < r.n_kf ? d.rows : r.n_kf;
        if (r.n_good < 0 || denom <= 0) continue;
        const double sim = (double)r.n_good / (double)denom;  // README.md:126
        if (sim > p.loop_threshold && r.n_good >= p.min_good) want.push_back(LoopCandidate{id, r.frame_id, r.n_good, sim});
      }
      for (const LoopClosingSystem* s : {&fused, &plain}) {
        const auto& got = s->getLoopClosures();
        if (got.size() != loops_seen + want.size()) return fail("number of loop closures", i);
        for (size_t j = 0; j < want.size(); ++j) {
          const LoopCandidate& g = got[loops_seen + j];
          if (g.current_frame_id != want[j].current_frame_id || g.matched_frame_id != want[j].matched_frame_id || g.num_matches != want[j].num_matches ||
              std::memcmp(&g.similarity_score, &want[j].similarity_score, sizeof(double)) != 0)
            return fail("LoopCandidate", i);
        }
      }
      loops_seen += want.size();
    }
    if (loops_seen == 0) return fail("the sequence closed no loop: the test would prove nothing", F);
    if (fused.getFrames().size() != (size_t)F || fused.detectLoops(-12345).size() != 0) return fail("getFrames / detectLoops(unknown id)", F);
    const char* out_dir = std::getenv("LC_TEST_OUT");
    if (out_dir) fused.saveResults(out_dir);
    std::printf("PASS: %d frames through loop_closing::b200::LoopClosingSystem (fused and unfused) == C oracle: %zu loop closures, %ld good consecutive matches\n",
                F, loops_seen, total_good);
    return 0;
  } catch (const std::exception& e) {
    std::printf("LoopClosingSystem could not start: %s\n", e.what());  // no B200 here: "no CPU fallback"
    return std::strstr(e.what(), "no CPU fallback") || std::strstr(e.what(), "sm_100a code only") ? 0 : 1;
  }
}
