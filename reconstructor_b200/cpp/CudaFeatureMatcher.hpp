// CudaFeatureMatcher.hpp -- drop-in FeatureMatcher plugin over the C ABI.
//
// Keeps the reference's abstract base and call signature (Mapper/libMapper/FeatureMatcher.h:11-27):
//     void matchFeatures(features1, features2, std::map<int,int>& matches, imgShape1, imgShape2)
// `matches` is appended into (caller passes an empty map, SequentialReconstructor.cpp:216),
// key = index into features1, value = index into features2, every value at most once.
// Replaces FlannMatcher (FeatureMatcher.cpp:27-65): exact brute-force 2-NN on the GPU instead of
// FLANN, then the same ratio test and first-come uniqueness.
//
// Per-image device cache: the reference hands over COPIES of the same vector<shared_ptr<Feature>>
// for every pair an image takes part in (.cpp:213-214) and re-packs them every time
// (featDescToCV, FeatureMatcher.cpp:11-25).  Here an image is packed + uploaded once.  The cache key is
// the identity of EVERY Feature object of the vector (an order-sensitive hash of the shared_ptr targets,
// the first and last pointer and the size): the copies the reference makes share those objects, while a
// vector that was freed and rebuilt -- even at the same address and of the same size -- holds other objects.
// The cache also keeps the shared_ptrs alive, so a cached address cannot be recycled under it.  Features
// edited in place are not detected: call invalidate() (it also frees the device rows).
#pragma once

#include <atomic>
#include <cstdint>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/pairmatch_b200.h"
#include "pm_types.hpp"

namespace reconstructor::Core {

#ifndef PM_USE_REFERENCE_HEADERS
class FeatureMatcher {
 public:
  explicit FeatureMatcher(const bool /*featNormalization*/ = false) {}
  virtual void matchFeatures(const std::vector<FeaturePtr<>>& features1,
                             const std::vector<FeaturePtr<>>& features2, std::map<int, int>& matches,
                             const std::pair<int, int> imgShape1, const std::pair<int, int> imgShape2) = 0;
  virtual ~FeatureMatcher() {}
};
#endif

// Owns one pm_handle; shared by the matcher, the filter and the batched loop.
class PairMatchDevice {
 public:
  explicit PairMatchDevice(const pm_params* p = nullptr, const std::vector<int>& devices = {}) {
    pm_params prm;
    if (p) prm = *p; else pm_default_params(&prm);
    const int rc = pm_create(&prm, devices.empty() ? nullptr : devices.data(), static_cast<int>(devices.size()), &h_);
    // like the reference's constructors, failure here is the one place that throws
    // (std::runtime_error, SequentialReconstructor.cpp:28,40,49)
    if (rc != PM_OK) throw std::runtime_error(std::string("pairmatch_b200: ") + pm_last_error(nullptr));
  }
  ~PairMatchDevice() { pm_destroy(h_); }
  PairMatchDevice(const PairMatchDevice&) = delete;
  PairMatchDevice& operator=(const PairMatchDevice&) = delete;
  pm_handle handle() const { return h_; }

  // Flattens vector<FeaturePtr<>> (AoS, one heap vector per feature) into row-major descriptors +
  // int32 xy and uploads them under `img_id`.
  int upload(int img_id, const std::vector<FeaturePtr<>>& f) {
    const int n = static_cast<int>(f.size());
    const int dim = n ? static_cast<int>(f[0]->featDesc.desc.size()) : 0;
    std::vector<float> desc(static_cast<size_t>(n) * dim);
    std::vector<int32_t> xy(static_cast<size_t>(n) * 2);
    for (int i = 0; i < n; ++i) {
      if (static_cast<int>(f[i]->featDesc.desc.size()) != dim) return PM_ERR_INVALID;
      std::copy(f[i]->featDesc.desc.begin(), f[i]->featDesc.desc.end(), desc.begin() + static_cast<size_t>(i) * dim);
      xy[2 * i] = static_cast<int32_t>(f[i]->featCoord.x);
      xy[2 * i + 1] = static_cast<int32_t>(f[i]->featCoord.y);
    }
    return pm_set_image(h_, img_id, desc.data(), n, dim > 0 ? dim : 4, PM_DESC_F32, xy.data());
  }

 private:
  pm_handle h_ = nullptr;
};

class CudaExhaustiveMatcher : public FeatureMatcher {
 public:
  explicit CudaExhaustiveMatcher(std::shared_ptr<PairMatchDevice> dev = nullptr)
      : dev_(dev ? std::move(dev) : std::make_shared<PairMatchDevice>()) {}

  void matchFeatures(const std::vector<FeaturePtr<>>& features1, const std::vector<FeaturePtr<>>& features2,
                     std::map<int, int>& matches, const std::pair<int, int> /*imgShape1*/,
                     const std::pair<int, int> /*imgShape2*/) override {
    if (features1.empty() || features2.empty()) return;   // the reference only asserts (:39)
    int a, b;
    {
      std::lock_guard<std::mutex> lk(mu_);   // called from up to 4 OpenMP threads (.cpp:202)
      a = resident(features1);
      b = resident(features2);
    }
    if (a < 0 || b < 0) { fail(PM_ERR_INVALID); return; }
    std::vector<int32_t> q(features1.size()), t(features1.size());
    pm_pair_result r{};
    r.capacity = static_cast<int32_t>(features1.size());
    r.q = q.data(); r.t = t.data(); r.inlier = nullptr;
    const int rc = pm_match_pair(dev_->handle(), a, b, &r);
    if (rc != PM_OK) { fail(rc); return; }    // no exceptions inside the OpenMP region
    for (int i = 0; i < r.n_matches; ++i) matches[q[i]] = t[i];
  }

  // Drops the cache AND the device rows behind it.
  void invalidate() {
    std::lock_guard<std::mutex> lk(mu_);
    for (auto& kv : cache_) pm_remove_image(dev_->handle(), kv.second.id);
    cache_.clear();
  }
  // First error any thread has seen since the last clearStatus() (PM_OK if none); the virtual call itself cannot
  // report one (void, FeatureMatcher.h:18-22) and must not throw inside the OpenMP region.
  int lastStatus() const { return first_error_.load(std::memory_order_acquire); }
  void clearStatus() { first_error_.store(PM_OK, std::memory_order_release); }
  const std::shared_ptr<PairMatchDevice>& device() const { return dev_; }

 private:
  void fail(int rc) {
    int expected = PM_OK;
    first_error_.compare_exchange_strong(expected, rc, std::memory_order_acq_rel);
  }
  struct Key {
    const void *first, *last;
    size_t n, h;
    bool operator==(const Key& o) const { return first == o.first && last == o.last && n == o.n && h == o.h; }
  };
  struct KeyHash { size_t operator()(const Key& k) const { return k.h ^ (k.n * 1000003u); } };
  struct Entry { int id; std::vector<FeaturePtr<>> keep; };
  static Key key_of(const std::vector<FeaturePtr<>>& f) {
    size_t h = 1469598103934665603ull;
    for (const auto& p : f) { h ^= std::hash<const void*>{}(p.get()); h *= 1099511628211ull; }
    return Key{f.front().get(), f.back().get(), f.size(), h};
  }
  int resident(const std::vector<FeaturePtr<>>& f) {
    const Key k = key_of(f);
    auto it = cache_.find(k);
    if (it != cache_.end()) return it->second.id;
    const int id = next_id_++;
    if (dev_->upload(id, f) != PM_OK) return -1;
    cache_.emplace(k, Entry{id, f});
    return id;
  }
  std::shared_ptr<PairMatchDevice> dev_;
  std::unordered_map<Key, Entry, KeyHash> cache_;
  std::mutex mu_;
  int next_id_ = 1 << 20;                   // away from the ids the batched loop uses
  std::atomic<int> first_error_{PM_OK};
};

}  // namespace reconstructor::Core
