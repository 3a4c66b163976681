// Minimal stand-in for the OpenCV types the C++ host side (include/loop_closing_b200.hpp, include/loop_closing_system_b200.hpp)
// touches, so that it can be compiled and RUN in an image without OpenCV's C++ headers: cv::Mat (an owning n x cols byte /
// float matrix with the members the host side uses), cv::DMatch, cv::KeyPoint, cv::Point3f.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
#define CV_8UC1 0
#define CV_32FC1 5
#define CV_64FC1 6
namespace cv {
struct Mat {
  int rows = 0, cols = 0;
  unsigned char* data = nullptr;
  struct Step {
    size_t v[2] = {0, 0};
    size_t operator[](int i) const { return v[i]; }
  } step;
  int type_ = CV_8UC1;
  std::shared_ptr<std::vector<unsigned char>> buf_;
  Mat() = default;
  Mat(int r, int c, int type) : rows(r), cols(c), type_(type) {
    const size_t es = type == CV_32FC1 ? 4 : type == CV_64FC1 ? 8 : 1;
    buf_ = std::make_shared<std::vector<unsigned char>>((size_t)r * c * es);
    data = buf_->data();
    step.v[0] = (size_t)c * es;
    step.v[1] = es;
  }
  Mat clone() const {
    Mat m(rows, cols, type_);
    for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step.v[0], data + (size_t)r * step.v[0], m.step.v[0]);
    return m;
  }
  template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step.v[0]); }
  template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step.v[0]); }
  int type() const { return type_; }
  bool empty() const { return rows == 0 || data == nullptr; }
};
struct DMatch {
  DMatch() = default;
  DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
  int queryIdx = -1, trainIdx = -1, imgIdx = -1;
  float distance = 0.f;
};
struct Point2f {
  float x = 0.f, y = 0.f;
};
struct Point3f {
  float x = 0.f, y = 0.f, z = 0.f;
};
struct KeyPoint {
  Point2f pt;
};
}  // namespace cv
