// CudaGeometricFilter.hpp -- drop-in for GeometricFilter::estimateFundamental
// (Mapper/libMapper/GeometricFilter.h:33-35, GeometricFilter.cpp:39-61).
//
// Contract kept: inputs are equal-length, already matched, ordered by ascending query index;
// `inlierMatchIds` is pushed back into (must arrive empty); on failure the all-zero matrix is
// returned and the vector stays EMPTY, upon which the caller drops the pair
// (SequentialReconstructor.cpp:253-256).
// estimateEssential (GeometricFilter.h:23-27, GeometricFilter.cpp:10-37): same signature; returns E of unit norm.  The
// reference hands no mask to cv::findEssentialMat (:25-33), so its inlierMatchIds comes back EMPTY whatever the data;
// that is kept by default (a caller counting "essentialInliers", .cpp:360-367, sees 0 as before) and
// setFillEssentialMask(true) returns the RANSAC mask instead.
#pragma once

#include <vector>

#include "CudaFeatureMatcher.hpp"

namespace reconstructor::Core {

class CudaGeometricFilter {
 public:
  explicit CudaGeometricFilter(std::shared_ptr<PairMatchDevice> dev = nullptr)
      : dev_(dev ? std::move(dev) : std::make_shared<PairMatchDevice>()) {}

  pmshim::Matrix3d estimateFundamental(const std::vector<FeaturePtr<>>& features1,
                                       const std::vector<FeaturePtr<>>& features2,
                                       std::vector<bool>& inlierMatchIds) {
    const int m = static_cast<int>(features1.size());
    std::vector<float> p1(2 * static_cast<size_t>(m)), p2(2 * static_cast<size_t>(m));
    for (int i = 0; i < m; ++i) {      // featuresToCvPoints: int -> float (utils.cpp:165-177)
      p1[2 * i] = static_cast<float>(features1[i]->featCoord.x);
      p1[2 * i + 1] = static_cast<float>(features1[i]->featCoord.y);
      p2[2 * i] = static_cast<float>(features2[i]->featCoord.x);
      p2[2 * i + 1] = static_cast<float>(features2[i]->featCoord.y);
    }
    std::vector<uint8_t> mask(static_cast<size_t>(m > 0 ? m : 1));
    double F[9];
    int32_t status = PM_PAIR_DROPPED, iters = 0;
    const int rc = pm_filter_pair_F(dev_->handle(), p1.data(), p2.data(), m, F, mask.data(), &status, &iters);
    pmshim::Matrix3d out = pmshim::Matrix3d::Zero();
    if (rc != PM_OK || status != PM_PAIR_FILTERED) return out;       // GeometricFilter.cpp:50-53
    for (int i = 0; i < m; ++i) inlierMatchIds.push_back(mask[i] != 0);   // writeInliersToVector
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) out(r, c) = F[3 * r + c];               // cvMatToEigen3d
    return out;
  }

  pmshim::Matrix3d estimateEssential(const std::vector<FeaturePtr<>>& features1,
                                     const std::vector<FeaturePtr<>>& features2, const PinholeCamera& intrinsics1,
                                     const PinholeCamera& intrinsics2, std::vector<bool>& inlierMatchIds) {
    const int m = static_cast<int>(features1.size());
    std::vector<float> p1(2 * static_cast<size_t>(m)), p2(2 * static_cast<size_t>(m));
    for (int i = 0; i < m; ++i) {      // featuresToCvPoints (utils.cpp:165-177)
      p1[2 * i] = static_cast<float>(features1[i]->featCoord.x);
      p1[2 * i + 1] = static_cast<float>(features1[i]->featCoord.y);
      p2[2 * i] = static_cast<float>(features2[i]->featCoord.x);
      p2[2 * i + 1] = static_cast<float>(features2[i]->featCoord.y);
    }
    const pm_camera c1{intrinsics1.fX, intrinsics1.fY, intrinsics1.cX, intrinsics1.cY, intrinsics1.k1, intrinsics1.k2};
    const pm_camera c2{intrinsics2.fX, intrinsics2.fY, intrinsics2.cX, intrinsics2.cY, intrinsics2.k1, intrinsics2.k2};
    std::vector<uint8_t> mask(static_cast<size_t>(m > 0 ? m : 1));
    double E[9];
    int32_t status = PM_PAIR_DROPPED, iters = 0;
    const int rc = pm_filter_pair_E(dev_->handle(), p1.data(), p2.data(), m, &c1, &c2, E, mask.data(), &status, &iters);
    pmshim::Matrix3d out = pmshim::Matrix3d::Zero();
    if (rc != PM_OK || status != PM_PAIR_FILTERED) return out;
    if (fill_essential_mask_)
      for (int i = 0; i < m; ++i) inlierMatchIds.push_back(mask[i] != 0);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) out(r, c) = E[3 * r + c];               // cvMatToEigen3d
    return out;
  }
  void setFillEssentialMask(bool on) { fill_essential_mask_ = on; }

 private:
  std::shared_ptr<PairMatchDevice> dev_;
  bool fill_essential_mask_ = false;
};

}  // namespace reconstructor::Core
