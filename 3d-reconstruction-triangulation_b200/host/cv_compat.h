// cv_compat.h -- the two OpenCV value types that appear in the reference's Triangulator interface.
// With OpenCV available, build with -DTRI_HAVE_OPENCV and the real cv::Point2d / cv::Point3d are used
// (a maintainer of the reference integrates that way); otherwise these layout-
// compatible stand-ins keep the host side free of any OpenCV dependency.
#pragma once
#ifdef TRI_HAVE_OPENCV
#include <opencv2/core.hpp>
#else
#include <cmath>
namespace cv {
struct Point2d {
  double x = 0, y = 0;
  Point2d() = default;
  Point2d(double x_, double y_) : x(x_), y(y_) {}
};
struct Point3d {
  double x = 0, y = 0, z = 0;
  Point3d() = default;
  Point3d(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
  Point3d operator-(const Point3d& o) const { return {x - o.x, y - o.y, z - o.z}; }
  Point3d operator*(double s) const { return {x * s, y * s, z * s}; }
  bool operator==(const Point3d& o) const { return x == o.x && y == o.y && z == o.z; }
  bool operator!=(const Point3d& o) const { return !(*this == o); }
};
inline double norm(const Point3d& p) { return std::sqrt(p.x * p.x + p.y * p.y + p.z * p.z); }
}  // namespace cv
#endif
