// Minimal stand-in for <opencv2/core.hpp> (cv::Mat as a row-major 16-bit image), compile-check only -- see Eigen/Core next to it.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>
namespace cv {
struct Mat {
  int rows = 0, cols = 0; size_t step = 0; std::vector<uint16_t> data;
  Mat() {}
  Mat(int r, int c) : rows(r), cols(c), step((size_t)c * sizeof(uint16_t)), data((size_t)r * c) {}
  template <class T> T *ptr(int r) { return reinterpret_cast<T *>(data.data()) + (size_t)r * cols; }
  template <class T> const T *ptr(int r) const { return reinterpret_cast<const T *>(data.data()) + (size_t)r * cols; }
};
}  // namespace cv
