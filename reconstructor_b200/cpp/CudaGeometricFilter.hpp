// CudaGeometricFilter.hpp -- drop-in for GeometricFilter::estimateFundamental
// (Mapper/libMapper/GeometricFilter.h:33-35, GeometricFilter.cpp:39-61).
//
// Contract kept: inputs are equal-length, already matched, ordered by ascending query index;
// `inlierMatchIds` is pushed back into (must arrive empty); on failure the all-zero matrix is
// returned and the vector stays EMPTY, upon which the caller drops the pair
// (SequentialReconstructor.cpp:253-256).  estimateEssential is a "next" row and is
// not provided here.
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

 private:
  std::shared_ptr<PairMatchDevice> dev_;
};

}  // namespace reconstructor::Core
