// loop_closing_b200.hpp - C++ adapter: the reference's LoopClosingSystem hot path on liblc_b200.
//
// What a maintainer of F-Fer/SLAM-Loop-Closing would add: the class interface of
// include/loop_closing.hpp:29-80 stays byte-for-byte; only the bodies of
//     matchFeatures  (include/loop_closing.hpp:40)
//     detectLoops    (include/loop_closing.hpp:48)
// and one line after frames_.push_back (include/loop_closing.hpp:69, in processFrame) change.  The
// cv::BFMatcher member `matcher_` (include/loop_closing.hpp:73) is replaced by an lc_ctx*.
//
// This header needs OpenCV's C++ headers (cv::Mat, cv::DMatch).  They are not installed in the build
// image of this repo, so it is not compiled by the test-suite; the executable proof of the same
// drop-in is slam_loop_closing_b200/loop_closing.py, which binds the same C ABI.
#pragma once

#include <opencv2/core.hpp>
#include <opencv2/features2d.hpp>

#include <stdexcept>
#include <string>
#include <vector>

#include "lc_abi.h"

namespace loop_closing {
namespace b200 {

struct LoopCandidate {  // == loop_closing::LoopCandidate (include/loop_closing.hpp:22-27) == lc_loop_candidate
  int current_frame_id;
  int matched_frame_id;
  int num_matches;
  double similarity_score;
};
static_assert(sizeof(LoopCandidate) == sizeof(lc_loop_candidate), "layout must match the C ABI");

class HammingMatcher {
 public:
  HammingMatcher(double loop_threshold = 0.7, int min_loop_gap = 30, int device = 0) {
    lc_config cfg;
    lc_default_config(&cfg);
    cfg.device = device;
    cfg.loop_threshold = loop_threshold;
    cfg.min_loop_gap = min_loop_gap;
    if (lc_create(&cfg, &ctx_) != LC_OK) throw std::runtime_error(lc_last_error(nullptr));
  }
  ~HammingMatcher() { lc_destroy(ctx_); }
  HammingMatcher(const HammingMatcher&) = delete;
  HammingMatcher& operator=(const HammingMatcher&) = delete;

  // body of LoopClosingSystem::matchFeatures(frame1, frame2): BFMatcher(NORM_HAMMING).match + 2 x min filter
  std::vector<cv::DMatch> matchFeatures(const cv::Mat& desc1, const cv::Mat& desc2) {
    check(desc1);
    check(desc2);
    const int nq = desc1.rows;
    std::vector<int32_t> idx(nq), dist(nq);
    std::vector<uint8_t> good(nq);
    int32_t n = 0, min_dist = 0, n_good = 0;
    ok(lc_match_pair(ctx_, desc1.data, nq, desc1.step[0], desc2.data, desc2.rows, desc2.step[0], idx.data(), dist.data(),
                     good.data(), &n, &min_dist, &n_good));
    std::vector<cv::DMatch> out;
    out.reserve(n_good);
    for (int i = 0; i < n; ++i)
      if (good[i]) out.emplace_back(i, idx[i], 0, static_cast<float>(dist[i]));
    return out;
  }

  // after frames_.push_back(frame) in processFrame
  void addFrame(int frame_id, const cv::Mat& descriptors) {
    check(descriptors);
    ok(lc_db_append(ctx_, frame_id, descriptors.data, descriptors.rows, descriptors.rows ? descriptors.step[0] : 32, nullptr));
  }

  // body of LoopClosingSystem::detectLoops(current_frame_id)
  std::vector<LoopCandidate> detectLoops(int current_frame_id, const cv::Mat& current_descriptors) {
    check(current_descriptors);
    int64_t n_kf = 0;
    ok(lc_db_size(ctx_, &n_kf, nullptr, nullptr, nullptr));
    std::vector<LoopCandidate> out(static_cast<size_t>(n_kf));
    int64_t n = 0;
    ok(lc_detect_loops(ctx_, current_descriptors.data, current_descriptors.rows,
                       current_descriptors.rows ? current_descriptors.step[0] : 32, current_frame_id,
                       reinterpret_cast<lc_loop_candidate*>(out.data()), n_kf, &n));
    out.resize(static_cast<size_t>(n));
    return out;
  }

  lc_ctx* handle() { return ctx_; }

 private:
  static void check(const cv::Mat& d) {
    if (!d.empty() && (d.type() != CV_8UC1 || d.cols != LC_DESC_BYTES)) throw std::invalid_argument("descriptors must be CV_8UC1 n x 32");
  }
  void ok(int32_t rc) {
    if (rc != LC_OK) throw std::runtime_error(lc_last_error(ctx_));
  }
  lc_ctx* ctx_ = nullptr;
};

}  // namespace b200
}  // namespace loop_closing
