// Minimal stand-in for the two OpenCV types include/loop_closing_b200.hpp touches, so that the adapter can be compiled
// (syntax + every call against include/lc_abi.h) in an image without OpenCV's C++ headers.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstddef>
#include <cstdint>
#define CV_8UC1 0
#define CV_32FC1 5
namespace cv {
struct Mat {
  int rows = 0, cols = 0;
  unsigned char* data = nullptr;
  struct Step {
    size_t v[2] = {0, 0};
    size_t operator[](int i) const { return v[i]; }
  } step;
  int type_ = CV_8UC1;
  int type() const { return type_; }
  bool empty() const { return rows == 0 || data == nullptr; }
};
struct DMatch {
  DMatch() = default;
  DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
  int queryIdx = -1, trainIdx = -1, imgIdx = -1;
  float distance = 0.f;
};
}  // namespace cv
