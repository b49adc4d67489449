"""Development aid: random two-view scenes (sizes 8..3000, 0..80 % outliers, integer / sub-pixel / shifted / scaled coordinates,
both samplers, both residual modes): staged filter == per-pair kernel (debug bit 21) == CPU filter (oracle), masks / F / iterations."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import orc
from reconstructor_b200 import api


def rot(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]); Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def scene(rng, n, of, subpixel, scale, shift):
    X = rng.uniform(-1, 1, (n, 3)) * np.array([1.0, 0.8, 0.6]) + np.array([0, 0, 5.0])
    f, cx, cy = 1200.0, 1024.0, 768.0
    R = rot(0.1 * rng.standard_normal(), rng.uniform(0.05, 0.35) * rng.choice([-1, 1]), 0.05 * rng.standard_normal())
    t = np.array([rng.uniform(0.3, 1.0), 0.1 * rng.standard_normal(), 0.1 * rng.standard_normal()])
    proj = lambda P: np.stack([f * P[:, 0] / P[:, 2] + cx, f * P[:, 1] / P[:, 2] + cy], 1)
    p1 = proj(X) + 0.5 * rng.standard_normal((n, 2))
    p2 = proj(X @ R.T + t) + 0.5 * rng.standard_normal((n, 2))
    bad = rng.random(n) < of
    p2[bad] = rng.uniform(0, 1, (int(bad.sum()), 2)) * np.array([2048, 1536])
    if not subpixel:
        p1, p2 = np.trunc(p1), np.trunc(p2)
    return (p1 * scale + shift).astype(np.float32), (p2 * scale + shift).astype(np.float32)


lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = n_long = 0
for seed in range(lo, hi):
    rng = np.random.default_rng(777 + seed)
    sampler = int(rng.integers(0, 2)); resid = int(rng.integers(0, 2))
    cap = int(rng.choice([1000, 1000, 1000, 300, 2000, 25, 281]))
    n = int(rng.choice([8, 9, 15, 40, 100, 333, 1000, 2050, 3000]))
    of = float(rng.choice([0.0, 0.2, 0.4, 0.5, 0.6, 0.8]))
    scale, shift = [(1.0, 0.0), (1.0, 0.0), (0.37, 0.0), (1.0, 50000.0), (31.0, 0.0)][int(rng.integers(0, 5))]
    p1, p2 = scene(rng, n, of, bool(rng.integers(0, 2)), scale, shift)
    osamp = orc.SAMPLER_PHILOX if sampler else orc.SAMPLER_OPENCV_MWC
    prm = orc.default_params(residual_mode=resid, sampler=osamp, seed=seed, max_iters=cap)
    ns, Fo, mo, tr = orc.find_fundamental(p1, p2, prm)
    res = []
    for flags in (0, 1 << 21):
        with api.PairMatcher(sampler=sampler, residual_mode=resid, seed=seed, ransac_max_iters=cap, debug_flags=flags) as pm:
            res.append(pm.estimate_fundamental(p1, p2))
    (F, mask, st, it), (F1, mask1, st1, it1) = res
    ok = st == st1 and it == it1 and np.array_equal(mask, mask1) and np.array_equal(F, F1)
    ok = ok and ((st == api.PAIR_FILTERED) == (ns > 0)) and it == tr.iters_run and (ns == 0 or np.array_equal(mask, mo))
    n_long += it > 24
    if not ok:
        bad += 1
        print("FAIL seed", seed, "n", n, "of", of, "sampler", sampler, "resid", resid, "cap", cap, "iters", it, it1, tr.iters_run,
              "inl", int(mask.sum()), int(mask1.sum()), int(mo.sum()))
print("ransac fuzz done: seeds %d..%d, failures %d, scenes past the hand-over %d" % (lo, hi, bad, n_long))
