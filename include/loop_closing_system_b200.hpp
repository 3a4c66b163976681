// loop_closing_system_b200.hpp - the reference's LoopClosingSystem on liblc_b200, in the reference's own language.
//
// Same public interface as loop_closing::LoopClosingSystem (reference include/loop_closing.hpp:29-66: constructor
// defaults, processFrame, detectFeatures, matchFeatures, estimatePose, detectLoops, triangulatePoints, visualizeMatches,
// getFrames, getLoopClosures, saveResults; Frame :12-19, LoopCandidate :22-27), following the workflow README.md:92-146
// describes - the reference's own src/loop_closing.cpp is absent from its snapshot.  What differs is the member behind
// matchFeatures / detectLoops: the cv::Ptr<cv::BFMatcher> matcher_ (:73) is a b200::HammingMatcher, i.e. an lc_ctx of
// include/lc_abi.h, and a frame's three matcher calls go to the GPU as ONE submission (lc_process_frame_begin / _matches /
// _end) while estimatePose / triangulatePoints run on the host.
//
// ORB detection, essential-matrix RANSAC, recoverPose, triangulation and drawing stay in OpenCV (BASELINE north_star): they
// are compiled when OpenCV's calib3d header is present (LC_WITH_OPENCV).  This repo's build image has no OpenCV C++
// headers; there the class compiles against tests/stubs/opencv2 and is driven through processDescriptors() - processFrame
// for a caller that already holds the ORB descriptors - which is what tests/cpp/system_check.cpp runs on the B200 against
// the C oracle.
#pragma once

#include "loop_closing_b200.hpp"

#if !defined(LC_WITH_OPENCV)
#if defined(__has_include)
#if __has_include(<opencv2/calib3d.hpp>)
#define LC_WITH_OPENCV 1
#endif
#endif
#endif
#ifndef LC_WITH_OPENCV
#define LC_WITH_OPENCV 0
#endif
#if LC_WITH_OPENCV
#include <opencv2/calib3d.hpp>
#include <opencv2/imgcodecs.hpp>
#include <opencv2/imgproc.hpp>
#endif

#include <cstdio>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace loop_closing {
namespace b200 {

struct Frame {  // include/loop_closing.hpp:12-19
  int id = 0;
  cv::Mat image;
  std::vector<cv::KeyPoint> keypoints;
  cv::Mat descriptors;
  cv::Mat pose;  // 4x4 transformation matrix
  std::vector<cv::Point3f> points3D;
};

class LoopClosingSystem {
 public:
  // fused: one GPU submission per frame (lc_process_frame_*); false: the three separate calls of the unfused drop-in
  explicit LoopClosingSystem(double loop_threshold = 0.7, int min_loop_gap = 30, bool fused = true, int device = 0)
      : matcher_(loop_threshold, min_loop_gap, device), loop_threshold_(loop_threshold), min_loop_gap_(min_loop_gap), fused_(fused) {}

  // Process a single frame (include/loop_closing.hpp:34): ORB on the host, matching on the GPU
  void processFrame(const cv::Mat& image, int frame_id) {
#if LC_WITH_OPENCV
    Frame frame;
    frame.id = frame_id;
    frame.image = image;
    frame.pose = cv::Mat::eye(4, 4, CV_64F);
    detectFeatures(frame);
    process(frame, true);
#else
    (void)image, (void)frame_id;
    throw std::runtime_error("processFrame needs OpenCV (ORB); without it use processDescriptors");
#endif
  }
  // processFrame for a caller that already holds the frame's ORB descriptors (CV_8UC1, n x 32): the hot path only
  void processDescriptors(const cv::Mat& descriptors, int frame_id) {
    Frame frame;
    frame.id = frame_id;
    frame.descriptors = descriptors.clone();
    process(frame, false);
  }

  // Match features between two frames (:40, README.md:116-117): BFMatcher(NORM_HAMMING).match + 2 x min-distance filter
  std::vector<cv::DMatch> matchFeatures(const Frame& frame1, const Frame& frame2) {
    return matcher_.matchFeatures(frame1.descriptors, frame2.descriptors);
  }

  // Check for loop closure (:48, README.md:121-126)
  std::vector<LoopCandidate> detectLoops(int current_frame_id) {
    for (size_t i = frames_.size(); i-- > 0;)
      if (frames_[i].id == current_frame_id) return matcher_.detectLoops(current_frame_id, frames_[i].descriptors);
    return {};
  }

#if LC_WITH_OPENCV
  // Detect features in an image (:37, README.md:114-115): ORB, 2000 features, 256-bit descriptors
  void detectFeatures(Frame& frame) {
    if (!feature_detector_) feature_detector_ = cv::ORB::create(2000);
    cv::Mat gray = frame.image;
    if (frame.image.channels() == 3) cv::cvtColor(frame.image, gray, cv::COLOR_BGR2GRAY);
    feature_detector_->detectAndCompute(gray, cv::noArray(), frame.keypoints, frame.descriptors);
  }
  // Estimate relative pose between two frames (:43, README.md:128-132): essential matrix by RANSAC, at least 8 matches
  bool estimatePose(const Frame& frame1, const Frame& frame2, const std::vector<cv::DMatch>& matches, cv::Mat& R, cv::Mat& t) {
    if (matches.size() < 8) return false;
    std::vector<cv::Point2f> p1, p2;
    for (const auto& m : matches) {
      p1.push_back(frame1.keypoints[m.queryIdx].pt);
      p2.push_back(frame2.keypoints[m.trainIdx].pt);
    }
    cv::Mat mask, E = cv::findEssentialMat(p1, p2, K_, cv::RANSAC, 0.999, 1.0, mask);
    if (E.rows != 3 || E.cols != 3) return false;
    return cv::recoverPose(E, p1, p2, K_, R, t, mask) > 0;
  }
  // Triangulate points from two views (:51, README.md:134-138): points behind the camera or beyond 100 units are dropped
  void triangulatePoints(const Frame& frame1, const Frame& frame2, const std::vector<cv::DMatch>& matches, std::vector<cv::Point3f>& points3D) {
    points3D.clear();
    if (matches.empty()) return;
    cv::Mat P1 = K_ * frame1.pose.inv()(cv::Rect(0, 0, 4, 3)), P2 = K_ * frame2.pose.inv()(cv::Rect(0, 0, 4, 3));
    std::vector<cv::Point2f> p1, p2;
    for (const auto& m : matches) {
      p1.push_back(frame1.keypoints[m.queryIdx].pt);
      p2.push_back(frame2.keypoints[m.trainIdx].pt);
    }
    cv::Mat X;
    cv::triangulatePoints(P1, P2, p1, p2, X);
    X.convertTo(X, CV_64F);
    const cv::Mat world_to_cam2 = frame2.pose.inv();
    for (int i = 0; i < X.cols; ++i) {
      const double w = X.at<double>(3, i);
      if (std::abs(w) < 1e-12) continue;
      const cv::Mat p = (cv::Mat_<double>(4, 1) << X.at<double>(0, i) / w, X.at<double>(1, i) / w, X.at<double>(2, i) / w, 1.0);
      const cv::Mat c = world_to_cam2 * p;
      if (c.at<double>(2) > 0 && cv::norm(p.rowRange(0, 3)) < 100.0)
        points3D.emplace_back((float)p.at<double>(0), (float)p.at<double>(1), (float)p.at<double>(2));
    }
  }
  // Visualize matches between two frames (:56)
  cv::Mat visualizeMatches(const Frame& frame1, const Frame& frame2, const std::vector<cv::DMatch>& matches) {
    cv::Mat out;
    cv::drawMatches(frame1.image, frame1.keypoints, frame2.image, frame2.keypoints, matches, out);
    return out;
  }
#endif

  const std::vector<Frame>& getFrames() const { return frames_; }                          // :60
  const std::vector<LoopCandidate>& getLoopClosures() const { return loop_closures_; }     // :63
  // good-match count of every matchFeatures(previous, current): what the pose stage was handed (for tests and logs)
  const std::vector<int>& consecutiveMatchCounts() const { return consecutive_matches_; }

  // Save results to file (:66, format of README.md:150-165): <output_dir>/loop_closures.txt
  void saveResults(const std::string& output_dir) {
    std::ofstream f(output_dir + "/loop_closures.txt");
    if (!f) throw std::runtime_error("cannot write " + output_dir + "/loop_closures.txt");
    f << "Total frames processed: " << frames_.size() << "\n" << "Loop closures detected: " << loop_closures_.size() << "\n\n"
      << "Loop Closures Detected:\n======================\n\n";
    for (const auto& lc : loop_closures_) {
      char sim[32];
      std::snprintf(sim, sizeof sim, "%.6g", lc.similarity_score);
      f << "Frame " << lc.current_frame_id << " <-> Frame " << lc.matched_frame_id << "\n  Matches: " << lc.num_matches << "\n  Similarity: " << sim << "\n\n";
    }
  }

 private:
  // README.md:94-102: pair match with the previous frame, pose + triangulation, store the frame, loop check
  void process(Frame& frame, bool geometry) {
    if (fused_) matcher_.beginFrame(frame.id, frame.descriptors);  // pair match + append + loop check queued on the GPU
    if (!frames_.empty()) {
      const Frame& prev = frames_.back();
      const std::vector<cv::DMatch> matches = fused_ ? matcher_.frameMatches() : matchFeatures(prev, frame);
      consecutive_matches_.push_back((int)matches.size());
      (void)geometry;
#if LC_WITH_OPENCV
      cv::Mat R, t;
      if (geometry && estimatePose(prev, frame, matches, R, t)) {  // on the host, while the GPU scans the store
        cv::Mat T = cv::Mat::eye(4, 4, CV_64F);
        R.copyTo(T(cv::Rect(0, 0, 3, 3)));
        t.copyTo(T(cv::Rect(3, 0, 1, 3)));
        frame.pose = prev.pose * T.inv();
        triangulatePoints(prev, frame, matches, frame.points3D);
      }
#endif
    }
    frames_.push_back(frame);  // include/loop_closing.hpp:69
    std::vector<LoopCandidate> loops;
    if (fused_) {
      loops = matcher_.finishFrame();
    } else {
      matcher_.addFrame(frame.id, frame.descriptors);
      loops = detectLoops(frame.id);
    }
    loop_closures_.insert(loop_closures_.end(), loops.begin(), loops.end());
  }

  std::vector<Frame> frames_;
  std::vector<LoopCandidate> loop_closures_;
  std::vector<int> consecutive_matches_;
#if LC_WITH_OPENCV
  cv::Ptr<cv::ORB> feature_detector_;
  cv::Mat K_ = (cv::Mat_<double>(3, 3) << 800, 0, 640, 0, 800, 360, 0, 0, 1);  // README.md:136
#endif
  HammingMatcher matcher_;  // replaces cv::Ptr<cv::BFMatcher> matcher_ (include/loop_closing.hpp:73)
  double loop_threshold_;
  int min_loop_gap_;
  bool fused_;
};

}  // namespace b200
}  // namespace loop_closing
