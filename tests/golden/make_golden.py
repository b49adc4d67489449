#!/usr/bin/env python
"""Generates the golden vectors under tests/golden/ from the reference's arithmetic carrier.

Run HERE (the build container) only:   python tests/golden/make_golden.py

The reference has no tests, fixtures or golden vectors of its own, and it
cannot be compiled in this image.  Its hot path is three OpenCV calls
(FeatureMatcher.cpp:29,49; GeometricFilter.cpp:47), so the goldens are outputs of those
calls in Python cv2 (version recorded in MANIFEST.json) on seeded inputs, plus -- for the
real-data fixture -- SIFT features of three of the reference's own sample images
(/root/reference/data/0000-0002.jpg, detected the way SequentialReconstructor::detectFeatures
does: resize to max side 512 with the %8 rule (utils.cpp:61-99), SIFT, int-truncated coords
(FeatureDetector.cpp:24-29)).  /root/reference is read ONLY by this script; tests read the
.npz files.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402

from oracle import cv2_ref  # noqa: E402
from reconstructor_b200 import synth  # noqa: E402

cv2.setNumThreads(1)


def two_view_scene(rng, n, outlier_frac, subpixel=False, noise=0.5):
    """Random two-view geometry -> matched pixel coordinates (n x 2 each)."""
    X = rng.uniform(-1, 1, (n, 3)) * np.array([1.0, 0.8, 0.6]) + np.array([0, 0, 5.0])
    f, cx, cy = 1200.0, 1024.0, 768.0
    ang = rng.uniform(0.05, 0.35) * rng.choice([-1, 1])
    R = cv2.Rodrigues(np.array([0.1 * rng.standard_normal(), ang, 0.05 * rng.standard_normal()]))[0]
    t = np.array([rng.uniform(0.3, 1.0), 0.1 * rng.standard_normal(), 0.1 * rng.standard_normal()])
    def proj(P):
        return np.stack([f * P[:, 0] / P[:, 2] + cx, f * P[:, 1] / P[:, 2] + cy], 1)
    p1 = proj(X) + noise * rng.standard_normal((n, 2))
    p2 = proj(X @ R.T + t) + noise * rng.standard_normal((n, 2))
    bad = rng.random(n) < outlier_frac
    p2[bad] = rng.uniform(0, 1, (int(bad.sum()), 2)) * np.array([2048, 1536])
    if not subpixel:
        p1, p2 = np.trunc(p1), np.trunc(p2)
    return p1.astype(np.float32), p2.astype(np.float32)


def gen_fmat():
    rng = np.random.default_rng(20261018)
    scenes = {}
    k = 0
    cfgs = []
    for n in (7, 8, 10, 14, 15, 16, 20, 33, 64, 100, 257, 500, 1000, 2048):
        for of in (0.0, 0.2, 0.5, 0.65):
            cfgs.append((n, of, False))
    for n in (15, 40, 300, 1200):
        for of in (0.1, 0.4):
            cfgs.append((n, of, True))
    for n, of, sub in cfgs:
        p1, p2 = two_view_scene(rng, n, of, subpixel=sub)
        F, mask = cv2_ref.estimate_fundamental(p1, p2)
        scenes[f"s{k}_p1"] = p1
        scenes[f"s{k}_p2"] = p2
        if F is None:
            scenes[f"s{k}_ok"] = np.array(0)
        else:
            Ffull, _ = cv2.findFundamentalMat(p1.reshape(-1, 1, 2), p2.reshape(-1, 1, 2))
            scenes[f"s{k}_ok"] = np.array(1)
            scenes[f"s{k}_F"] = np.asarray(Ffull, np.float64)     # 3x3 or 9x3 (n == 7)
            scenes[f"s{k}_mask"] = mask
        scenes[f"s{k}_meta"] = np.array([n, of, float(sub)])
        k += 1
    # degenerate: identical points, collinear points
    p = np.tile(np.array([[100.0, 200.0]], np.float32), (20, 1))
    for name, (a, b) in {"same": (p, p),
                         "line": (np.stack([np.arange(20.0), 2 * np.arange(20.0)], 1).astype(np.float32),
                                  np.stack([np.arange(20.0), 3 * np.arange(20.0)], 1).astype(np.float32))}.items():
        F, mask = cv2_ref.estimate_fundamental(a, b)
        scenes[f"deg_{name}_p1"] = a
        scenes[f"deg_{name}_p2"] = b
        scenes[f"deg_{name}_ok"] = np.array(0 if F is None else 1)
    scenes["n_scenes"] = np.array(k)
    np.savez_compressed(os.path.join(HERE, "fmat_scenes.npz"), **scenes)
    return k


def gen_cubic():
    rng = np.random.default_rng(7)
    C, N, R = [], [], []
    while len(C) < 400:
        c = rng.standard_normal(4) * 10 ** rng.uniform(-1, 1, 4)
        if len(C) % 9 == 0:
            c[0] = 0.0
        if len(C) % 63 == 0:
            c[1] = 0.0
        n, r = cv2.solveCubic(c.reshape(4, 1))
        r = r.reshape(-1)
        if not np.all(np.isfinite(r)):
            continue
        C.append(c); N.append(n); R.append(np.pad(r, (0, 3 - len(r))))
    np.savez_compressed(os.path.join(HERE, "cubic.npz"), coeffs=np.array(C), n=np.array(N, np.int32),
                        roots=np.array(R))


def gen_knn_and_pairs():
    out = {}
    for kind, n in (("sift", 384), ("orb", 500), ("superpoint", 320)):
        w = synth.World(kind, n, seed=0xB200 + 1)
        imgs = [w.image(i, n_images_on_ring=3) for i in range(3)]
        for i, (d, xy, _) in enumerate(imgs):
            if kind == "sift":
                out[f"{kind}_desc{i}"] = d.astype(np.uint8)
            else:
                out[f"{kind}_desc{i}"] = d
            out[f"{kind}_xy{i}"] = xy
        # planted ties / duplicates: train rows 5, 77, 200 all equal query row 10,
        # and train rows 300, 31 are the same near-copy of query row 20 (equal non-zero distance)
        d1 = imgs[1][0].copy()
        d1[5] = imgs[0][0][10]; d1[77] = d1[5]; d1[200] = d1[5]
        near = imgs[0][0][20].copy()
        near[0] = near[0] ^ 1 if kind == "orb" else (near[0] + (1 if kind == "sift" else 0.01))
        d1[300] = near; d1[31] = near
        out[f"{kind}_desc1_ties"] = d1.astype(np.uint8) if kind == "sift" else d1
        for (a, b, tag) in ((0, 1, "01"), (0, 2, "02"), (1, 2, "12")):
            da, db = imgs[a][0], imgs[b][0]
            idx, dist = cv2_ref.knn2_bf(da, db)
            out[f"{kind}_{tag}_idx"] = idx
            out[f"{kind}_{tag}_dist"] = dist
            r = cv2_ref.match_pair(da, imgs[a][1], db, imgs[b][1], matcher="bf")
            out[f"{kind}_{tag}_status"] = np.array(1 if r["status"] == "ok" else 0)
            out[f"{kind}_{tag}_q"] = r["q"]; out[f"{kind}_{tag}_t"] = r["t"]
            out[f"{kind}_{tag}_nput"] = np.array(r["n_putative"])
        idx, dist = cv2_ref.knn2_bf(imgs[0][0], d1)
        out[f"{kind}_ties_idx"] = idx
        out[f"{kind}_ties_dist"] = dist
    np.savez_compressed(os.path.join(HERE, "knn_pairs.npz"), **out)


def detect_like_reference(path):
    img = cv2.imread(path)
    img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
    rows, cols = img.shape[:2]
    if rows > cols:
        if rows > 512:
            cs = 512 * (cols / rows); cs = cs - np.fmod(cs, 8)
            img = cv2.resize(img, (int(cs), 512))
    elif cols > 512:
        rs = 512 * (rows / cols); rs = rs - np.fmod(rs, 8)
        img = cv2.resize(img, (512, int(rs)))
    gray = cv2.cvtColor(img, cv2.COLOR_RGB2GRAY)
    kps, desc = cv2.SIFT_create().detectAndCompute(gray, None)
    xy = np.array([[int(k.pt[0]), int(k.pt[1])] for k in kps], np.int32)
    return desc.astype(np.float32), xy


def gen_fountain():
    data = "/root/reference/data"
    out = {}
    feats = []
    for i in range(3):
        d, xy = detect_like_reference(os.path.join(data, f"{i:04d}.jpg"))
        assert np.all(d == np.rint(d)) and d.min() >= 0 and d.max() <= 255
        feats.append((d, xy))
        out[f"desc{i}"] = d.astype(np.uint8)
        out[f"xy{i}"] = xy
    for (a, b, tag) in ((0, 1, "01"), (0, 2, "02"), (1, 2, "12")):
        r = cv2_ref.match_pair(feats[a][0], feats[a][1], feats[b][0], feats[b][1], matcher="bf")
        out[f"{tag}_status"] = np.array(1 if r["status"] == "ok" else 0)
        out[f"{tag}_q"] = r["q"]; out[f"{tag}_t"] = r["t"]; out[f"{tag}_nput"] = np.array(r["n_putative"])
        idx, dist = cv2_ref.knn2_bf(feats[a][0], feats[b][0])
        out[f"{tag}_idx"] = idx; out[f"{tag}_dist"] = dist
        rf = cv2_ref.match_pair(feats[a][0], feats[a][1], feats[b][0], feats[b][1], matcher="flann")
        out[f"{tag}_flann_q"] = rf["q"]; out[f"{tag}_flann_t"] = rf["t"]
    np.savez_compressed(os.path.join(HERE, "fountain.npz"), **out)
    return [f[0].shape[0] for f in feats]


def main():
    nsc = gen_fmat()
    gen_cubic()
    gen_knn_and_pairs()
    nk = gen_fountain() if os.path.isdir("/root/reference/data") else None
    man = dict(cv2=cv2.__version__, numpy=np.__version__, fmat_scenes=nsc, fountain_keypoints=nk,
               note="generated by tests/golden/make_golden.py; cv2 carries the reference's arithmetic")
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
        json.dump(man, f, indent=1)
    print(man)


def gen_eight_point():
    """cv2.findFundamentalMat(FM_8POINT) over the inliers of cv2's own RANSAC mask, for the scenes of
    fmat_scenes.npz (pins the optional normalised 8-point refit, pm_params.refit_8point)."""
    z = np.load(os.path.join(HERE, "fmat_scenes.npz"))
    out = {}
    for k in range(int(z["n_scenes"])):
        if not int(z[f"s{k}_ok"]):
            continue
        p1, p2, mask = z[f"s{k}_p1"], z[f"s{k}_p2"], z[f"s{k}_mask"].astype(bool)
        if p1.shape[0] < 15 or mask.sum() < 8:
            continue
        F, _ = cv2.findFundamentalMat(p1[mask], p2[mask], cv2.FM_8POINT)
        if F is not None and F.shape == (3, 3):
            out[f"s{k}_F8"] = F
    out["scenes"] = np.array(sorted(int(k[1:-3]) for k in out if k.endswith("_F8")), np.int32)
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(HERE, "eight_point.npz"), **out)




def gen_essential():
    """cv2.findEssentialMat(p1, p2, K1, d1, K2, d2) -- the call of GeometricFilter::estimateEssential
    (GeometricFilter.cpp:26-31) with its defaults -- on seeded two-view scenes: equal and unequal cameras, with and
    without the radial distortion of PinholeCamera (Camera.h:113-123), N from 5 to 2000, 0-50 % outliers."""
    rng = np.random.default_rng(20261019)
    out, k = {}, 0
    for n in (5, 6, 8, 12, 20, 40, 100, 300, 1000, 2000):
        for of in (0.0, 0.25, 0.5):
            for cam_mode in (0, 1):
                if n <= 8 and of > 0:
                    continue
                p1, p2 = two_view_scene(rng, n, of, subpixel=False)
                if cam_mode == 0:          # the reference's cameras at chooseInitialPair: equal, no distortion
                    c1 = c2 = np.array([1200.0, 1200.0, 1024.0, 768.0, 0.0, 0.0])
                else:
                    c1 = np.array([1180.0, 1210.0, 1000.0, 770.0, 0.03, -0.01])
                    c2 = np.array([1230.0, 1195.0, 1040.0, 760.0, -0.02, 0.015])
                K1 = np.array([[c1[0], 0, c1[2]], [0, c1[1], c1[3]], [0, 0, 1.0]])
                K2 = np.array([[c2[0], 0, c2[2]], [0, c2[1], c2[3]], [0, 0, 1.0]])
                d1 = np.array([c1[4], c1[5], 0.0, 0.0]).reshape(4, 1)
                d2 = np.array([c2[4], c2[5], 0.0, 0.0]).reshape(4, 1)
                E, mask = cv2.findEssentialMat(p1.reshape(-1, 1, 2), p2.reshape(-1, 1, 2), K1, d1, K2, d2)
                ok = E is not None and E.shape[0] >= 3
                out[f"e{k}_p1"] = p1; out[f"e{k}_p2"] = p2; out[f"e{k}_c1"] = c1; out[f"e{k}_c2"] = c2
                out[f"e{k}_ok"] = np.array(int(ok))
                out[f"e{k}_E"] = np.asarray(E[:3], np.float64) if ok else np.zeros((3, 3))
                out[f"e{k}_mask"] = mask.reshape(-1).astype(np.uint8) if ok and mask is not None else np.zeros(n, np.uint8)
                k += 1
    out["n_scenes"] = np.array(k)
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(HERE, "essential.npz"), **out)
    return k


if __name__ == "__main__":
    # --eight-point / --essential: only (re)generate that file (the others stay as committed)
    if "--eight-point" in sys.argv:
        gen_eight_point()
    elif "--essential" in sys.argv:
        print("essential scenes:", gen_essential())
    else:
        main()
        gen_eight_point()
        gen_essential()
