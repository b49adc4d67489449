"""Development aid: per-phase cycle counts of the RANSAC kernel (needs the -DPROFILE build in ab/libpm_prof.so)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from reconstructor_b200 import api, synth

frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
w = synth.World("sift", 8192, seed=0xB200 + 2)
imgs = [w.image(i, 100, frac)[:2] for i in range(3)]
with api.PairMatcher() as pm:
    for i, (d, xy) in enumerate(imgs):
        pm.set_image(i, d, xy)
    r = pm.match_all_pairs()
    print("iters", r["ransac_iters"], "inl", r["n_inliers"], "put", np.diff(r["offsets"]))
