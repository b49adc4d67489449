"""Development aid: 512-bit binary rows, kind::mxf4 form (default) against the kind::f8f6f4 form (debug bit 16)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from reconstructor_b200 import api

rng = np.random.default_rng(3)
n_img, kp = 40, 8192
base = rng.integers(0, 256, (4 * kp, 64), dtype=np.uint8)
imgs = []
for i in range(n_img):
    ids = rng.permutation(4 * kp)[:kp]
    d = base[ids].copy()
    d ^= ((rng.random((kp, 64)) < 0.03) * rng.integers(1, 256, (kp, 64))).astype(np.uint8)
    xy = rng.integers(0, 1500, (kp, 2)).astype(np.int32)
    imgs.append((d, xy))
ref = None
for flags, name in ((0, "mxf4"), (1 << 16, "e4m3")):
    with api.PairMatcher(debug_flags=flags, do_filter=0, reserve_keypoints=n_img * kp) as pm:
        for i, (d, xy) in enumerate(imgs):
            pm.set_image(i, d, xy)
        r = pm.match_all_pairs(copy=False); pm.free_result(r)
        pm.reset_stats()
        ms = 0.0
        for _ in range(3):
            r = pm.match_all_pairs(copy=False); ms += r["device_ms"]
            out = {k: np.array(r[k]) for k in ("offsets", "q", "t")}
            pm.free_result(r)
        st = pm.stats()
        n_pairs = n_img * (n_img - 1) // 2
        print(f"{name}: {n_pairs / (ms / 3 * 1e-3):.0f} pairs/s, {ms / 3:.2f} ms/step, knn {st['knn_ms'] / max(st['knn_launches'], 1):.3f} ms/launch, matches {out['offsets'][-1]}")
        if ref is None: ref = out
        else: print("identical:", all(np.array_equal(ref[k], out[k]) for k in ref))
