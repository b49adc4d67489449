"""Development aid: where does the end-to-end step go?  host time of the ingest call(s), of pm_match_all_pairs, batch timeline."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from reconstructor_b200 import api, shard, synth

n_img, kp = 100, 8192
w = synth.World("sift", kp, seed=0xB200 + 2)
pinned = []
for i in range(n_img):
    d, xy, _ = w.image(i, n_img, 0.0)
    pinned.append((torch.from_numpy(np.ascontiguousarray(d, np.float32)).pin_memory(),
                   torch.from_numpy(np.ascontiguousarray(xy, np.int32)).pin_memory()))
pairs = shard.all_pairs(n_img)
pm = api.PairMatcher(reserve_keypoints=n_img * kp)
ids = list(range(n_img)); dp = [t.data_ptr() for t, _ in pinned]; ns = [t.shape[0] for t, _ in pinned]; xp = [x.data_ptr() for _, x in pinned]
for rep in range(5):
    t0 = time.perf_counter()
    pm.set_images_ptr_async(ids, dp, ns, 128, api.DESC_F32, xp)
    t1 = time.perf_counter()
    r = pm.match_all_pairs(pairs, copy=False)
    t2 = time.perf_counter()
    s = int(r["n_inliers"].sum()); pm.free_result(r)
    t3 = time.perf_counter()
    print("rep %d: ingest call %.2f ms | match_all_pairs %.2f ms (device_ms %.2f) | read+free %.2f ms | total %.2f" %
          (rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), r.get("device_ms", -1) if isinstance(r, dict) else -1, 1e3 * (t3 - t2), 1e3 * (t3 - t0)))
# resident images: the pure matching call
for rep in range(3):
    t1 = time.perf_counter(); r = pm.match_all_pairs(pairs, copy=False); t2 = time.perf_counter(); pm.free_result(r)
    print("resident: match_all_pairs %.2f ms" % (1e3 * (t2 - t1)))
# upload alone
for rep in range(3):
    t0 = time.perf_counter(); pm.set_images_ptr_async(ids, dp, ns, 128, api.DESC_F32, xp); t1 = time.perf_counter(); pm.sync_images(); t2 = time.perf_counter()
    print("upload alone: call %.2f ms, until resident %.2f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t0)))
