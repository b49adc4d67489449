// pm_types.hpp -- the slice of the reference's data model that the hot path touches.
//
// With -DPM_USE_REFERENCE_HEADERS the reference's own datatypes.h is used (it drags in
// <opencv2/opencv.hpp> and <Eigen/Dense>, Mapper/libMapper/datatypes.h:7-8).  Without it -- neither
// library exists in this image -- a source-compatible minimal mirror is declared here: same
// namespace, same member names (featCoord.x/.y, featDesc.desc/.type), so that code written against
// the reference (FeatureMatcher.cpp:11-25, utils.cpp:165-177) compiles unchanged.
#pragma once

#include <array>
#include <memory>
#include <vector>

#ifdef PM_USE_REFERENCE_HEADERS
#include "datatypes.h"
#include "Camera.h"
namespace pmshim { using Matrix3d = Eigen::Matrix3d; }
#else
#ifndef CV_32F
#define CV_32F 5
#endif
#ifndef CV_8U
#define CV_8U 0
#endif
namespace reconstructor::Core {
// int pixel coordinates: the detector truncates sub-pixel positions (FeatureDetector.cpp:28-29)
template <typename coordType = int>
struct FeatCoord {
  FeatCoord() = default;
  FeatCoord(coordType x_, coordType y_) : x(x_), y(y_) {}
  virtual ~FeatCoord() = default;
  coordType x{}, y{};
};
// float descriptor, type hard-wired to CV_32F (datatypes.h:70-71)
struct FeatDesc {
  FeatDesc() = default;
  template <typename It>
  FeatDesc(It first, It last) : desc(first, last) {}
  std::vector<float> desc;
  int type = CV_32F;
};
template <typename coordType = int>
struct Feature {
  Feature() = default;
  Feature(FeatCoord<coordType> c, FeatDesc d) : featCoord(c), featDesc(std::move(d)) {}
  virtual ~Feature() = default;
  FeatCoord<coordType> featCoord;
  FeatDesc featDesc;
  int landmarkId = -1;
};
template <typename coordType = int>
using FeaturePtr = std::shared_ptr<Feature<coordType>>;
}  // namespace reconstructor::Core

// Camera.h:12-127: only the members estimateEssential reads (getMatrixCV / getDistortCV: fX, fY, cX, cY, k1, k2)
class PinholeCamera {
 public:
  PinholeCamera() = default;
  PinholeCamera(int height, int width, double fX_, double fY_) : fX(fX_), fY(fY_), cX(width / 2), cY(height / 2) {}
  double fX = 1, fY = 1, cX = 0, cY = 0, k1 = 0, k2 = 0;
};

namespace pmshim {
// Stand-in for Eigen::Matrix3d (column-major like Eigen, operator()(row, col)).
struct Matrix3d {
  std::array<double, 9> m{};
  double& operator()(int r, int c) { return m[c * 3 + r]; }
  double operator()(int r, int c) const { return m[c * 3 + r]; }
  static Matrix3d Zero() { return Matrix3d{}; }
  bool isZero() const { for (double v : m) if (v != 0.0) return false; return true; }
};
}  // namespace pmshim
#endif
