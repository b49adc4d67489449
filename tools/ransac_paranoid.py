"""Development aid: runs many RANSAC problems through the PM_RANSAC_PARANOID build (ab/libpm_paranoid.so), which
re-checks every point the conservative classifier decided against the literal OpenCV formula and prints mismatches."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
from reconstructor_b200 import api, synth

g = np.load("tests/golden/fmat_scenes.npz")
rng = np.random.default_rng(0)
with api.PairMatcher() as pm:
    for k in range(int(g["n_scenes"])):
        pm.estimate_fundamental(g[f"s{k}_p1"], g[f"s{k}_p2"])
    # scaled / shifted coordinates: sub-pixel, large offsets, huge scales
    for scale, shift in ((1.0, 0.0), (0.37, 0.0), (1.0, 50000.0), (977.0, 0.0), (1e-3, 0.0), (1.0, 3e6)):
        for k in (20, 30, 40, 50):
            p1 = (g[f"s{k}_p1"] * scale + shift).astype(np.float32)
            p2 = (g[f"s{k}_p2"] * scale + shift).astype(np.float32)
            pm.estimate_fundamental(p1, p2)
    for mode in (0, 1):
        for frac in (0.0, 0.3, 0.5):
            w = synth.World("orb", 4096, seed=7)
            imgs = [w.image(i, 50, frac)[:2] for i in range(4)]
            with api.PairMatcher(residual_mode=mode) as p2m:
                for i, (d, xy) in enumerate(imgs):
                    p2m.set_image(i, d, xy)
                r = p2m.match_all_pairs()
                print("mode", mode, "frac", frac, "iters", r["ransac_iters"].tolist(), "inl", r["n_inliers"].tolist())
print("paranoid run finished")
