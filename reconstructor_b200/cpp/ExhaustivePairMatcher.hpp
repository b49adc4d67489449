// ExhaustivePairMatcher.hpp -- batched replacement of SequentialReconstructor::matchFeatures(bool)
// (Mapper/libMapper/SequentialReconstructor.cpp:199-279): ONE pm_match_all_pairs call instead of the
// OpenMP double loop, then a `featureMatches`-shaped view (SequentialReconstructor.h:226) with both
// (i,j) and the mirrored (j,i) entries.  Fixes the reference's unsynchronised writes to the shared
// map and its schedule-dependent choice of which direction gets matched: query = lower image id.
#pragma once

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "CudaFeatureMatcher.hpp"

namespace reconstructor::Core {

struct pm_pair_hash {      // not the reference's h1 ^ h2 (SequentialReconstructor.h:51-62): (i,j)/(j,i) collide there
  size_t operator()(const std::pair<int, int>& p) const {
    return std::hash<long long>{}((static_cast<long long>(p.first) << 32) ^ static_cast<unsigned>(p.second));
  }
};
template <class Hash = pm_pair_hash>
using FeatureMatchesT = std::unordered_map<std::pair<int, int>, std::unordered_map<int, int>, Hash>;

// ImageMatcher plugin (Mapper/libMapper/ImageMatcher.h:14-24): pair pre-selection by global-descriptor retrieval on the
// device instead of FakeImgMatcher's "every image with every other one" (ImageMatcher.cpp:6-24; README.md:40 lists image
// retrieval as a todo).  Same call shape as ImageMatcher::match minus the unused path map; imgMatches receives BOTH
// directions of every selected pair, like FakeImgMatcher.  topK <= 0: all pairs (FakeImgMatcher's result).
class CudaRetrievalImgMatcher {
 public:
  explicit CudaRetrievalImgMatcher(std::shared_ptr<PairMatchDevice> dev = nullptr, int topK = 0)
      : dev_(dev ? std::move(dev) : std::make_shared<PairMatchDevice>()), top_k_(topK) {}
  int match(const std::unordered_map<int, std::vector<FeaturePtr<>>>& features,
            std::unordered_map<int, std::vector<int>>& imgMatches) {
    for (const auto& kv : features) {
      const int rc = dev_->upload(kv.first, kv.second);
      if (rc != PM_OK) return rc;
    }
    int32_t* pairs = nullptr;
    int64_t n = 0;
    std::vector<int32_t> ids;                     // only the images handed over: the handle may hold other uploads
    for (const auto& kv : features) ids.push_back(kv.first);
    const int rc = pm_select_pairs_among(dev_->handle(), ids.data(), static_cast<int>(ids.size()), top_k_, &pairs, &n, nullptr);
    if (rc != PM_OK) return rc;
    for (const auto& kv : features) imgMatches[kv.first];                    // an entry for every image, as the reference
    for (int64_t p = 0; p < n; ++p) {
      imgMatches[pairs[2 * p]].push_back(pairs[2 * p + 1]);
      imgMatches[pairs[2 * p + 1]].push_back(pairs[2 * p]);
    }
    pm_free_pairs(pairs);
    return PM_OK;
  }
  // The reference's call shape, ImageMatcher::match(imgIds2Paths, features, imgMatches) (ImageMatcher.h:18-21,
  // SequentialReconstructor.cpp:114): the path map is not read (FakeImgMatcher does not read it either); void like the
  // reference, the status is kept in lastStatus().
  template <class PathMap>
  void match(const PathMap& /*imgIds2Paths*/, const std::unordered_map<int, std::vector<FeaturePtr<>>>& features,
             std::unordered_map<int, std::vector<int>>& imgMatches) {
    last_status_ = match(features, imgMatches);
  }
  int lastStatus() const { return last_status_; }
  const std::shared_ptr<PairMatchDevice>& device() const { return dev_; }

 private:
  std::shared_ptr<PairMatchDevice> dev_;
  int top_k_;
  int last_status_ = PM_OK;
};

class ExhaustivePairMatcher {
 public:
  explicit ExhaustivePairMatcher(std::shared_ptr<PairMatchDevice> dev = nullptr)
      : dev_(dev ? std::move(dev) : std::make_shared<PairMatchDevice>()) {}

  // features: imgId -> features;  imgMatches: imgId -> matched image ids (FakeImgMatcher: all others).
  // Returns a pm status; on success featureMatches holds what the reference's loop would hold.
  template <class Hash>
  int matchFeatures(const std::unordered_map<int, std::vector<FeaturePtr<>>>& features,
                    const std::unordered_map<int, std::vector<int>>& imgMatches,
                    FeatureMatchesT<Hash>& featureMatches) {
    for (const auto& kv : features) {
      const int rc = dev_->upload(kv.first, kv.second);
      if (rc != PM_OK) return rc;
    }
    std::vector<int32_t> pairs;
    for (const auto& kv : imgMatches)
      for (int j : kv.second) {
        const int i = kv.first;
        if (i == j) continue;
        const auto back = imgMatches.find(j);
        const bool mirrored = back != imgMatches.end() &&
                              std::find(back->second.begin(), back->second.end(), i) != back->second.end();
        if (i < j || !mirrored) { pairs.push_back(i); pairs.push_back(j); }   // canonical direction once
      }
    if (pairs.empty()) { last_device_ms_ = 0; return PM_OK; }   // nothing to match (an empty list is NOT "all pairs")
    pm_csr_result* res = nullptr;
    const int rc = pm_match_all_pairs(dev_->handle(), pairs.data(), static_cast<int64_t>(pairs.size() / 2), &res);
    if (rc != PM_OK) return rc;
    for (int64_t p = 0; p < res->n_pairs; ++p) {
      if (res->status[p] == PM_PAIR_DROPPED) continue;                         // .cpp:253-256
      const int i = res->pair_ij[2 * p], j = res->pair_ij[2 * p + 1];
      for (int64_t k = res->offsets[p]; k < res->offsets[p + 1]; ++k) {
        if (!res->inlier[k]) continue;
        featureMatches[{i, j}][res->q[k]] = res->t[k];                         // .cpp:259-267 / :272-275
        featureMatches[{j, i}][res->t[k]] = res->q[k];                         // mirror, .cpp:219-227
      }
    }
    last_device_ms_ = res->device_ms;
    int src = PM_OK;
    if (!result_cache_.empty()) src = pm_save_result(res, result_cache_.c_str());   // "save intermediate steps", README.md:39
    pm_free_result(res);
    return src;
  }
  // Optional on-disk cache of the match result (written after every matchFeatures call) and of the ingested images.
  void setResultCache(std::string path) { result_cache_ = std::move(path); }
  int saveImages(const std::string& path) { return pm_save_images(dev_->handle(), path.c_str()); }
  int loadImages(const std::string& path) { return pm_load_images(dev_->handle(), path.c_str()); }
  double lastDeviceMs() const { return last_device_ms_; }
  const std::shared_ptr<PairMatchDevice>& device() const { return dev_; }

 private:
  std::shared_ptr<PairMatchDevice> dev_;
  double last_device_ms_ = 0;
  std::string result_cache_;
};

}  // namespace reconstructor::Core
