// shim_selftest.cpp -- exercises the C++ drop-in classes the way SequentialReconstructor uses its
// plugins (per-pair virtual calls from several threads) and checks them against the batched loop.
// Built by `make` (host compiler only); run on a GPU box by tests/test_gpu_shim.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>

#include "CudaGeometricFilter.hpp"
#include "ExhaustivePairMatcher.hpp"

using namespace reconstructor::Core;

static std::vector<FeaturePtr<>> make_image(int id, int n, const std::vector<std::vector<float>>& world,
                                            std::mt19937& g) {
  std::vector<FeaturePtr<>> f;
  std::vector<int> ids(world.size());
  for (size_t i = 0; i < ids.size(); ++i) ids[i] = static_cast<int>(i);
  std::shuffle(ids.begin(), ids.end(), g);
  std::normal_distribution<float> noise(0.f, 4.f);
  const double ang = 0.15 * id;
  for (int k = 0; k < n; ++k) {
    const int l = ids[k];
    std::vector<float> d(world[l]);
    for (auto& v : d) v = std::min(255.f, std::max(0.f, std::round(v + noise(g))));
    // a plane-induced motion is degenerate for F; use a crude 3-D scene: x shifts with depth
    const double X = (l % 97) / 97.0 - 0.5, Y = ((l / 97) % 89) / 89.0 - 0.5, Z = 4.0 + (l % 13) / 6.0;
    const double xc = X * std::cos(ang) + Z * std::sin(ang), zc = -X * std::sin(ang) + Z * std::cos(ang) + 0.3 * id;
    const int u = static_cast<int>(1200.0 * xc / zc + 1024), v = static_cast<int>(1200.0 * Y / zc + 768);
    f.push_back(std::make_shared<Feature<>>(FeatCoord<>(u, v), FeatDesc(d.begin(), d.end())));
  }
  return f;
}

int main() {
  std::mt19937 g(7);
  std::uniform_int_distribution<int> u8(0, 160);
  std::vector<std::vector<float>> world(1500, std::vector<float>(128));
  for (auto& w : world) for (auto& v : w) v = static_cast<float>(u8(g));
  const int n_img = 4;
  std::unordered_map<int, std::vector<FeaturePtr<>>> features;
  std::unordered_map<int, std::vector<int>> imgMatches;
  for (int i = 0; i < n_img; ++i) features[i] = make_image(i, 600 - 20 * i, world, g);
  for (int i = 0; i < n_img; ++i) for (int j = 0; j < n_img; ++j) if (i != j) imgMatches[i].push_back(j);

  auto dev = std::make_shared<PairMatchDevice>();
  CudaExhaustiveMatcher matcher(dev);
  CudaGeometricFilter filter(dev);
  ExhaustivePairMatcher loop(dev);

  // (1) the reference's loop body, per pair, from 4 threads on the shared plugin objects
  FeatureMatchesT<> viaPlugins;
  std::mutex mu;
  std::vector<std::thread> th;
  for (int tid = 0; tid < 4; ++tid)
    th.emplace_back([&, tid] {
      int k = 0;
      for (int i = 0; i < n_img; ++i)
        for (int j = i + 1; j < n_img; ++j, ++k) {
          if (k % 4 != tid) continue;
          auto f1 = features[i], f2 = features[j];                      // copies, like .cpp:213-214
          std::map<int, int> cur;
          matcher.matchFeatures(f1, f2, cur, {0, 0}, {0, 0});
          std::unordered_map<int, int> kept;
          if (cur.size() >= 7) {
            std::vector<FeaturePtr<>> m1, m2;
            for (auto& [a, b] : cur) { m1.push_back(f1[a]); m2.push_back(f2[b]); }
            std::vector<bool> inl;
            filter.estimateFundamental(m1, m2, inl);
            if (inl.empty()) continue;
            int c = 0;
            for (auto& [a, b] : cur) { if (inl[c]) kept[a] = b; ++c; }
          } else {
            for (auto& [a, b] : cur) kept[a] = b;
          }
          std::lock_guard<std::mutex> lk(mu);
          for (auto& [a, b] : kept) { viaPlugins[{i, j}][a] = b; viaPlugins[{j, i}][b] = a; }
        }
    });
  for (auto& t : th) t.join();

  // (2) the batched loop
  FeatureMatchesT<> viaBatch;
  const int rc = loop.matchFeatures(features, imgMatches, viaBatch);
  if (rc != PM_OK) { std::printf("batched loop failed: %s\n", pm_last_error(dev->handle())); return 2; }

  size_t total = 0;
  if (viaPlugins.size() != viaBatch.size()) { std::printf("pair count differs %zu vs %zu\n", viaPlugins.size(), viaBatch.size()); return 1; }
  for (auto& [key, m] : viaBatch) {
    auto it = viaPlugins.find(key);
    if (it == viaPlugins.end() || it->second != m) { std::printf("pair (%d,%d) differs\n", key.first, key.second); return 1; }
    total += m.size();
  }
  if (viaBatch.size() != static_cast<size_t>(n_img * (n_img - 1)) || total < 400) { std::printf("too few matches: %zu pairs, %zu matches\n", viaBatch.size(), total); return 1; }
  if (matcher.lastStatus() != PM_OK) { std::printf("plugin reported status %d\n", matcher.lastStatus()); return 1; }

  // (3) an empty image-match list is an empty job, not "all pairs of the handle"
  {
    FeatureMatchesT<> none;
    std::unordered_map<int, std::vector<int>> noMatches;
    if (loop.matchFeatures(features, noMatches, none) != PM_OK || !none.empty()) { std::printf("empty list ran pairs\n"); return 1; }
  }
  // (4) the plugin's image cache keys on the Feature objects: a rebuilt vector of the same size (same address or not)
  //     is a new image; invalidate() frees the device rows
  {
    std::map<int, int> before, after, again;
    matcher.matchFeatures(features[0], features[1], before, {0, 0}, {0, 0});
    std::vector<FeaturePtr<>> rebuilt;                       // image 1 with the descriptors of image 2 (same size as ...)
    for (size_t k = 0; k < features[1].size(); ++k)
      rebuilt.push_back(std::make_shared<Feature<>>(features[1][k]->featCoord, features[2][k % features[2].size()]->featDesc));
    matcher.matchFeatures(features[0], rebuilt, after, {0, 0}, {0, 0});
    if (before == after) { std::printf("stale upload served for a rebuilt feature vector\n"); return 1; }
    matcher.invalidate();
    matcher.matchFeatures(features[0], features[1], again, {0, 0}, {0, 0});
    if (before != again || matcher.lastStatus() != PM_OK) { std::printf("results changed after invalidate()\n"); return 1; }
  }
  // (5) estimateEssential with the reference's signature: E of unit norm, x2' E x1 ~ 0 for the inliers of F; the
  //     reference's inlierMatchIds stays empty (no mask is passed to cv::findEssentialMat) unless asked otherwise
  {
    auto f1 = features[0], f2 = features[1];
    std::map<int, int> cur;
    matcher.matchFeatures(f1, f2, cur, {0, 0}, {0, 0});
    std::vector<FeaturePtr<>> m1, m2;
    for (auto& [a, b] : cur) { m1.push_back(f1[a]); m2.push_back(f2[b]); }
    PinholeCamera cam(1536, 2048, 1200.0, 1200.0);
    std::vector<bool> inl;
    auto E = filter.estimateEssential(m1, m2, cam, cam, inl);
    if (!inl.empty() || E.isZero()) { std::printf("estimateEssential: reference behaviour not kept\n"); return 1; }
    filter.setFillEssentialMask(true);
    auto E2 = filter.estimateEssential(m1, m2, cam, cam, inl);
    size_t n_in = 0;
    for (bool b : inl) n_in += b;
    double nrm = 0;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { nrm += E2(r, c) * E2(r, c); if (E2(r, c) != E(r, c)) { std::printf("E not deterministic\n"); return 1; } }
    if (inl.size() != m1.size() || n_in < m1.size() / 2 || std::fabs(nrm - 1.0) > 1e-9) { std::printf("estimateEssential: %zu of %zu inliers, |E|^2 = %g\n", n_in, m1.size(), nrm); return 1; }
  }
  // (6) the ImageMatcher plugin: top-k retrieval lists are symmetric, without self-matches; topK = 0 is FakeImgMatcher
  {
    CudaRetrievalImgMatcher fake(dev, 0), top1(dev, 1);
    std::unordered_map<int, std::vector<int>> all, near;
    std::unordered_map<int, std::string> paths;            // the reference's first argument (never read)
    fake.match(paths, features, all);                       // the reference's call shape (ImageMatcher.h:18-21)
    if (fake.lastStatus() != PM_OK || top1.match(features, near) != PM_OK) { std::printf("retrieval failed\n"); return 1; }
    for (int i = 0; i < n_img; ++i) {
      if (all[i].size() != static_cast<size_t>(n_img - 1) || near[i].empty() || near[i].size() > all[i].size()) { std::printf("retrieval lists wrong for image %d\n", i); return 1; }
      for (int j : near[i]) {
        const auto& back = near[j];
        if (j == i || std::find(back.begin(), back.end(), i) == back.end()) { std::printf("retrieval list not symmetric\n"); return 1; }
      }
    }
  }
  std::printf("SHIM_OK pairs=%zu matches=%zu device_ms=%.3f\n", viaBatch.size(), total, loop.lastDeviceMs());
  return 0;
}
