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
// image of this repo: the test-suite compiles it against a minimal stand-in for those two types
// (tests/stubs/opencv2/), which checks every call against include/lc_abi.h; the executable proof of the same
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
    last_rows_ = descriptors.rows;
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

  // detectLoops split in two, so that processFrame can detect the next image's features while the GPU scans the store:
  // beginDetectLoops copies the descriptors and returns at once, finishDetectLoops waits and applies README.md:121-126
  void beginDetectLoops(int current_frame_id, const cv::Mat& current_descriptors) {
    check(current_descriptors);
    pending_nq_ = current_descriptors.rows;
    pending_id_ = current_frame_id;
    ok(lc_loop_check_begin(ctx_, current_descriptors.data, current_descriptors.rows,
                           current_descriptors.rows ? current_descriptors.step[0] : 32, current_frame_id, 0));
  }
  std::vector<LoopCandidate> finishDetectLoops() {
    int64_t n_kf = 0, n = 0, n_out = 0;
    ok(lc_db_size(ctx_, &n_kf, nullptr, nullptr, nullptr));
    std::vector<lc_kf_result> rec(static_cast<size_t>(n_kf));
    ok(lc_loop_check_end(ctx_, rec.data(), n_kf, &n, nullptr));
    std::vector<LoopCandidate> out(static_cast<size_t>(n));
    ok(lc_select_candidates(ctx_, rec.data(), n, pending_nq_, pending_id_, reinterpret_cast<lc_loop_candidate*>(out.data()), n, &n_out));
    out.resize(static_cast<size_t>(n_out));
    return out;
  }

  // The matcher side of a whole processFrame (include/loop_closing.hpp:34) in one GPU submission: matchFeatures(previous
  // frame, this frame) + frames_.push_back + detectLoops(frame_id) behind a single upload of the descriptors.
  //   beginFrame(id, desc);                                   // everything queued, returns at once
  //   auto matches = frameMatches();                          // waits for the pair match only
  //   estimatePose(prev, cur, matches, R, t); triangulatePoints(...);   // host, while the GPU scans the store
  //   auto loops = finishFrame();                             // waits for the loop check
  void beginFrame(int frame_id, const cv::Mat& descriptors) {
    check(descriptors);
    frame_prev_rows_ = last_rows_;
    ok(lc_process_frame_begin(ctx_, frame_id, descriptors.data, descriptors.rows, descriptors.rows ? descriptors.step[0] : 32, nullptr));
    last_rows_ = descriptors.rows;
  }
  std::vector<cv::DMatch> frameMatches() {
    const int nq = frame_prev_rows_;
    std::vector<int32_t> idx(nq), dist(nq);
    std::vector<uint8_t> good(nq);
    int32_t n = 0, min_dist = 0, n_good = 0;
    ok(lc_process_frame_matches(ctx_, idx.data(), dist.data(), good.data(), nq, &n, &min_dist, &n_good));
    std::vector<cv::DMatch> out;
    out.reserve(n_good);
    for (int i = 0; i < n; ++i)
      if (good[i]) out.emplace_back(i, idx[i], 0, static_cast<float>(dist[i]));
    return out;
  }
  std::vector<LoopCandidate> finishFrame() {
    int64_t n_kf = 0, n = 0;
    ok(lc_db_size(ctx_, &n_kf, nullptr, nullptr, nullptr));
    std::vector<LoopCandidate> out(static_cast<size_t>(n_kf));
    ok(lc_process_frame_end(ctx_, reinterpret_cast<lc_loop_candidate*>(out.data()), n_kf, &n, nullptr, 0, nullptr));
    out.resize(static_cast<size_t>(n));
    return out;
  }

  // matchFeatures(current, frames_[k]) for the loop candidates in ONE call: the DMatch lists estimatePose /
  // triangulatePoints (include/loop_closing.hpp:43-53) verify a loop with.  kf_index[i] = position of the matched frame
  // in frames_ (the store shares its insertion order).
  std::vector<std::vector<cv::DMatch>> matchCandidates(const cv::Mat& current_descriptors, const std::vector<int64_t>& kf_index) {
    check(current_descriptors);
    const int nq = current_descriptors.rows, n_sel = static_cast<int>(kf_index.size());
    std::vector<lc_match> m(static_cast<size_t>(nq) * kf_index.size());
    std::vector<int64_t> offs(kf_index.size() + 1);
    ok(lc_match_keyframes(ctx_, current_descriptors.data, nq, nq ? current_descriptors.step[0] : 32, n_sel, kf_index.data(), m.data(),
                          static_cast<int64_t>(m.size()), offs.data(), nullptr));
    std::vector<std::vector<cv::DMatch>> out(kf_index.size());
    for (size_t i = 0; i < kf_index.size(); ++i)
      for (int64_t j = offs[i]; j < offs[i + 1]; ++j)
        out[i].emplace_back(m[j].query_idx, m[j].train_idx, 0, static_cast<float>(m[j].distance));
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
  int pending_nq_ = 0, pending_id_ = 0;
  int last_rows_ = 0, frame_prev_rows_ = 0;  // rows of the frame appended last / of the previous frame of the frame in flight
};

// The SHIPPED pipeline (src/main.cpp): matchFeatures (:509-534) and the matching step of its loop search (:1378-1392)
class SiftMatcher {
 public:
  explicit SiftMatcher(int device = 0) {
    lc_config cfg;
    lc_default_config(&cfg);
    cfg.device = device;
    if (lc_create(&cfg, &ctx_) != LC_OK) throw std::runtime_error(lc_last_error(nullptr));
  }
  ~SiftMatcher() { lc_destroy(ctx_); }
  SiftMatcher(const SiftMatcher&) = delete;
  SiftMatcher& operator=(const SiftMatcher&) = delete;

  // void matchFeatures(desc1, desc2, goodMatches, ratio): knnMatch(k = 2) + Lowe's ratio test
  void matchFeatures(const cv::Mat& desc1, const cv::Mat& desc2, std::vector<cv::DMatch>& goodMatches, double ratio = 0.75) {
    goodMatches.clear();
    const int nq = desc1.rows;
    std::vector<int32_t> idx(nq);
    std::vector<float> dist(nq);
    std::vector<uint8_t> good(nq);
    int32_t n_good = 0;
    ok(lc_l2_match_pair(ctx_, desc1.data, nq, nq ? desc1.step[0] : 512, desc2.data, desc2.rows, desc2.rows ? desc2.step[0] : 512,
                        type_of(desc1), ratio, idx.data(), dist.data(), nullptr, good.data(), &n_good));
    goodMatches.reserve(n_good);
    for (int i = 0; i < nq; ++i)
      if (good[i]) goodMatches.emplace_back(i, idx[i], 0, dist[i]);
  }
  // allDescriptors.push_back(desc) of the shipped main(): the view joins the device-resident store
  void addView(int view, const cv::Mat& desc) {
    ok(lc_l2_db_append(ctx_, view, desc.data, desc.rows, desc.rows ? desc.step[0] : 512, type_of(desc), nullptr));
  }
  // matches.size() of (curr, past) for every stored view past <= curr - loopGap, -1 where main.cpp:1384 skips the pair
  std::vector<lc_l2_result> loopMatchCounts(const cv::Mat& curr_desc, int curr, int loopGap, double ratio = 0.7, int min_rows = 100) {
    int64_t n_local = 0, n = 0;
    ok(lc_l2_db_size(ctx_, nullptr, &n_local));
    std::vector<lc_l2_result> out(static_cast<size_t>(n_local));
    ok(lc_l2_loop_check(ctx_, curr_desc.data, curr_desc.rows, curr_desc.rows ? curr_desc.step[0] : 512, type_of(curr_desc), curr, loopGap,
                        ratio, min_rows, out.data(), n_local, &n));
    out.resize(static_cast<size_t>(n));
    return out;
  }
  // the whole double loop of main.cpp:1375-1389 in one call: counts[curr - loopGap][past] for curr = loopGap .. numViews - 1
  std::vector<std::vector<lc_l2_result>> loopSearch(int loopGap, double ratio = 0.7, int min_rows = 100) {
    int64_t n_views = 0;
    ok(lc_l2_db_size(ctx_, &n_views, nullptr));
    const int n_query = n_views > loopGap ? static_cast<int>(n_views - loopGap) : 0;
    std::vector<lc_l2_result> flat(static_cast<size_t>(n_query) * static_cast<size_t>(n_views));
    ok(lc_l2_loop_search(ctx_, loopGap, n_query, loopGap, ratio, min_rows, flat.data(), static_cast<int64_t>(flat.size()), nullptr));
    std::vector<std::vector<lc_l2_result>> out(static_cast<size_t>(n_query));
    for (int f = 0; f < n_query; ++f) out[f].assign(flat.begin() + static_cast<size_t>(f) * n_views, flat.begin() + static_cast<size_t>(f + 1) * n_views);
    return out;
  }

 private:
  static int32_t type_of(const cv::Mat& d) {
    if (d.empty() || (d.type() == CV_32FC1 && d.cols == LC_L2_DIMS)) return LC_DESC_F32;
    if (d.type() == CV_8UC1 && d.cols == LC_L2_DIMS) return LC_DESC_U8;
    throw std::invalid_argument("SIFT descriptors must be CV_32F or CV_8U, n x 128");
  }
  void ok(int32_t rc) {
    if (rc != LC_OK) throw std::runtime_error(lc_last_error(ctx_));
  }
  lc_ctx* ctx_ = nullptr;
};

}  // namespace b200
}  // namespace loop_closing
