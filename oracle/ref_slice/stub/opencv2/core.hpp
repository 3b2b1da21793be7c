// Minimal stand-in for <opencv2/core.hpp>: the reference's Ellipsoid.cpp pulls src/Polygon/Polygon.hpp, which only needs the 2-D point
// types.  Used by oracle/ref_slice/build_ref.sh when OpenCV's C++ headers are absent (test infrastructure, not product code).
#pragma once
#include <cassert>
#include <cmath>
namespace cv {
template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  template <typename U> Point_(const Point_<U> &o) : x(static_cast<T>(o.x)), y(static_cast<T>(o.y)) {}
  Point_ operator+(const Point_ &o) const { return Point_(x + o.x, y + o.y); }
  Point_ operator-(const Point_ &o) const { return Point_(x - o.x, y - o.y); }
  Point_ operator*(T s) const { return Point_(x * s, y * s); }
  Point_ &operator+=(const Point_ &o) { x += o.x; y += o.y; return *this; }
  bool operator==(const Point_ &o) const { return x == o.x && y == o.y; }
};
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
}  // namespace cv
