"""Pins the C oracle (oracle/pm_oracle.c) against the golden vectors generated from cv2, the
library that carries the reference's arithmetic (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import orc


@pytest.fixture(scope="module")
def scenes(golden_dir):
    return np.load(os.path.join(golden_dir, "fmat_scenes.npz"))


@pytest.fixture(scope="module")
def knn(golden_dir):
    return np.load(os.path.join(golden_dir, "knn_pairs.npz"))


@pytest.fixture(scope="module")
def fountain(golden_dir):
    return np.load(os.path.join(golden_dir, "fountain.npz"))


def test_cubic_matches_cv2(golden_dir):
    g = np.load(os.path.join(golden_dir, "cubic.npz"))
    for c, n, r in zip(g["coeffs"], g["n"], g["roots"]):
        no, ro = orc.solve_cubic(c)
        assert no == n
        # same root ORDER; values equal up to the conditioning of the trigonometric form
        np.testing.assert_allclose(ro[:n], r[:n], rtol=1e-7, atol=1e-9)


def test_update_num_iters():
    assert orc.update_num_iters(0.99, 0.0, 7, 1000) == 0
    assert orc.update_num_iters(0.99, 1.0, 7, 1000) == 1000
    assert orc.update_num_iters(0.99, 0.5, 7, 1000) == 587
    assert orc.update_num_iters(0.99, 0.5, 7, 300) == 300
    assert orc.update_num_iters(0.99, 0.05, 7, 1000) == 4


@pytest.mark.parametrize("kind", ["sift", "orb", "superpoint"])
def test_knn_matches_bfmatcher(knn, kind):
    for tag, a, b in (("01", "desc0", "desc1"), ("02", "desc0", "desc2"), ("12", "desc1", "desc2"),
                      ("ties", "desc0", "desc1_ties")):
        da, db = knn[f"{kind}_{a}"], knn[f"{kind}_{b}"]
        gi, gd = knn[f"{kind}_{tag}_idx"], knn[f"{kind}_{tag}_dist"]
        if kind == "orb":
            oi, od = orc.knn2_hamming(da, db)
            assert (oi == gi).all()
            assert (od.astype(np.float32) == gd).all()
        else:
            oi, o2 = orc.knn2_l2(da.astype(np.float32), db.astype(np.float32))
            od = np.sqrt(o2).astype(np.float32)
            if kind == "sift":       # integer-valued: bit exact
                assert (oi == gi).all()
                assert (od == gd).all()
            else:                    # fp64 arbiter; BFMatcher within 1e-5 rel, idx may differ only at near-ties
                np.testing.assert_allclose(od, gd, rtol=1e-5)
                diff = np.nonzero((oi != gi).any(axis=1))[0]
                for r in diff:            # the column BFMatcher chose instead must tie the arbiter's in fp64
                    for k in range(2):
                        if oi[r, k] != gi[r, k]:
                            alt = float(((da[r].astype(np.float64) - db[gi[r, k]].astype(np.float64)) ** 2).sum())
                            assert abs(alt - o2[r, k]) <= 1e-6 * o2[r, k], (kind, tag, r, k, alt, o2[r, k])
                assert len(diff) <= 1


def test_knn_tie_break_lowest_index(knn):
    # train rows 5/77/200 duplicate query 10; rows 31/300 are equal near-copies of query 20
    for kind in ("sift", "orb"):
        da, db = knn[f"{kind}_desc0"], knn[f"{kind}_desc1_ties"]
        if kind == "orb":
            oi, od = orc.knn2_hamming(da, db)
        else:
            oi, od = orc.knn2_l2(da.astype(np.float32), db.astype(np.float32))
        assert tuple(oi[10]) == (5, 77) and od[10, 0] == 0 and od[10, 1] == 0
        assert tuple(oi[20]) == (31, 300) and od[20, 0] == od[20, 1] and od[20, 0] > 0
        assert not np.isin(oi[:, 0], [77, 200, 300]).any()


def test_fmat_masks_identical_to_cv2(scenes):
    n_sc = int(scenes["n_scenes"])
    checked = exceptions = 0
    for k in range(n_sc):
        p1, p2 = scenes[f"s{k}_p1"], scenes[f"s{k}_p2"]
        n = p1.shape[0]
        ok = int(scenes[f"s{k}_ok"])
        ns, F, mask, tr = orc.find_fundamental(p1, p2)
        if n >= 15:
            assert (ns > 0) == bool(ok)
            if ok:
                checked += 1
                gm = scenes[f"s{k}_mask"]
                if not (gm == mask).all():
                    exceptions += 1
                    continue
                gF = scenes[f"s{k}_F"][:3]
                np.testing.assert_allclose(F[0], gF, rtol=0, atol=1e-9 * max(1.0, np.abs(gF).max()))
        elif n == 7:
            gF = scenes[f"s{k}_F"].reshape(-1, 3, 3)
            assert ns == gF.shape[0] and mask.all()
            # same model SET (order depends on the null-space basis)
            for Fo in F:
                assert min(np.abs(Fo - g).max() for g in gF) < 1e-7 * max(1.0, np.abs(Fo).max())
        else:
            # 8..14: OpenCV runs LMedS there (noise-determined); we run RANSAC
            assert ns in (0, 1)
            if ns:
                assert mask.sum() >= 7
    assert checked >= 40
    assert exceptions == 0


def test_fmat_degenerate_fails(scenes):
    for name in ("same", "line"):
        p1, p2 = scenes[f"deg_{name}_p1"], scenes[f"deg_{name}_p2"]
        assert int(scenes[f"deg_{name}_ok"]) == 0
        ns, F, mask, tr = orc.find_fundamental(p1, p2)
        assert ns == 0


def test_sampler_is_deterministic_and_distinct(scenes):
    p1, p2 = scenes["s20_p1"], scenes["s20_p2"]
    a = orc.sample_subsets(p1, p2, 50)
    b = orc.sample_subsets(p1, p2, 50)
    assert (a == b).all() and a.shape == (50, 7)
    for row in a:
        assert len(set(row.tolist())) == 7
        assert row.min() >= 0 and row.max() < p1.shape[0]


@pytest.mark.parametrize("kind", ["sift", "orb", "superpoint"])
def test_pair_body_matches_cv2(knn, kind):
    for tag, a, b in (("01", 0, 1), ("02", 0, 2), ("12", 1, 2)):
        da, db = knn[f"{kind}_desc{a}"], knn[f"{kind}_desc{b}"]
        if kind != "orb":
            da, db = da.astype(np.float32), db.astype(np.float32)
        r = orc.match_pair(da, knn[f"{kind}_xy{a}"], db, knn[f"{kind}_xy{b}"])
        assert r["n_putative"] == int(knn[f"{kind}_{tag}_nput"])
        assert (r["status"] == "ok") == bool(knn[f"{kind}_{tag}_status"])
        assert (r["q"] == knn[f"{kind}_{tag}_q"]).all() and (r["t"] == knn[f"{kind}_{tag}_t"]).all()


def test_fountain_real_images(fountain):
    """Three of the reference's own sample images (data/0000-0002.jpg) through the pair body."""
    for tag, a, b in (("01", 0, 1), ("02", 0, 2), ("12", 1, 2)):
        da, db = fountain[f"desc{a}"].astype(np.float32), fountain[f"desc{b}"].astype(np.float32)
        oi, o2 = orc.knn2_l2(da, db)
        assert (oi == fountain[f"{tag}_idx"]).all()
        assert (np.sqrt(o2).astype(np.float32) == fountain[f"{tag}_dist"]).all()
        r = orc.match_pair(da, fountain[f"xy{a}"], db, fountain[f"xy{b}"])
        assert r["n_putative"] == int(fountain[f"{tag}_nput"])
        assert (r["q"] == fountain[f"{tag}_q"]).all() and (r["t"] == fountain[f"{tag}_t"]).all()


def test_small_and_empty_inputs():
    q = np.zeros((3, 32), np.uint8); t = np.zeros((1, 32), np.uint8)
    idx, dist = orc.knn2_hamming(q, t)
    assert (idx[:, 0] == 0).all() and (idx[:, 1] == -1).all()
    oq, ot = orc.ratio_unique(idx, dist.astype(np.float32), 1)
    assert len(oq) == 0                       # < 2 neighbours => no match
    ns, F, mask, tr = orc.find_fundamental(np.zeros((6, 2), np.float32), np.zeros((6, 2), np.float32))
    assert ns == 0


# ---------------------------------------------------------------------------------------------
# Philox sampler and the 8-point refit (SURVEY App. B3 / north star "8-point")
# ---------------------------------------------------------------------------------------------
def test_philox4x32_known_answers():
    """Published Philox4x32-10 vectors (Random123 kat_vectors): pins the counter-based generator."""
    assert orc.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert orc.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert orc.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_sampler_properties(scenes):
    p1, p2 = scenes["s30_p1"], scenes["s30_p2"]
    n = p1.shape[0]
    seen = set()
    for it in range(200):
        ok, idx = orc.philox_subset(p1, p2, 0x1234, it)
        assert ok and len(set(idx.tolist())) == 7 and idx.min() >= 0 and idx.max() < n
        ok2, idx2 = orc.philox_subset(p1, p2, 0x1234, it)
        assert np.array_equal(idx, idx2)                    # pure function of (seed, iteration)
        seen.add(tuple(idx.tolist()))
    assert len(seen) == 200
    assert not np.array_equal(orc.philox_subset(p1, p2, 0x1235, 0)[1], orc.philox_subset(p1, p2, 0x1234, 0)[1])
    assert orc.pair_seed(7, 1, 2) != orc.pair_seed(7, 2, 1) and orc.pair_seed(7, 1, 2) != orc.pair_seed(8, 1, 2)
    # the filter with this sampler finds the same geometry as with OpenCV's stream (statistical agreement)
    prm = orc.default_params(sampler=orc.SAMPLER_PHILOX, seed=99)
    ns, F, mask, tr = orc.find_fundamental(p1, p2, prm)
    ns0, F0, mask0, tr0 = orc.find_fundamental(p1, p2)
    inter = int((mask & mask0).sum()); union = int((mask | mask0).sum())
    assert ns == 1 and ns0 == 1 and inter / union > 0.9


def test_eight_point_matches_cv2_golden(scenes, golden_dir):
    g = np.load(os.path.join(golden_dir, "eight_point.npz"))
    assert len(g["scenes"]) >= 40
    for k in g["scenes"]:
        p1, p2, mask = scenes[f"s{k}_p1"], scenes[f"s{k}_p2"], scenes[f"s{k}_mask"]
        ok, F = orc.eight_point(p1, p2, mask)
        assert ok
        Fg = g[f"s{k}_F8"]
        assert np.abs(F - Fg).max() <= 1e-8 * np.abs(Fg).max(), (int(k), np.abs(F - Fg).max())
    # degenerate: fewer than 8 points, identical points
    assert not orc.eight_point(scenes["s30_p1"][:7], scenes["s30_p2"][:7])[0]
    same = np.tile(np.array([[5.0, 9.0]], np.float32), (20, 1))
    assert not orc.eight_point(same, same)[0]


# ---------------------------------------------------------------------------------------------
# GeometricFilter::estimateEssential = cv::findEssentialMat
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def essential(golden_dir):
    return np.load(os.path.join(golden_dir, "essential.npz"))


def test_essential_oracle_matches_cv2_golden(essential):
    """Masks identical to cv2.findEssentialMat on every golden scene (equal / unequal / distorted cameras, N = 5..2000,
    up to 50 % outliers); E equal up to sign within 1e-4 (cv2's float undistortion is vectorised) whenever N > 5 -- for
    N = 5 cv2 returns the first model of ITS solver order, which is not reproducible (any returned model must fit)."""
    g = essential
    n = int(g["n_scenes"])
    assert n >= 40
    for k in range(n):
        p1, p2, c1, c2 = g[f"e{k}_p1"], g[f"e{k}_p2"], g[f"e{k}_c1"], g[f"e{k}_c2"]
        ok, E, m, tr = orc.find_essential(p1, p2, orc.Camera(*c1), orc.Camera(*c2))
        assert ok == bool(int(g[f"e{k}_ok"]))
        assert np.array_equal(m, g[f"e{k}_mask"]), k
        Eg = g[f"e{k}_E"]
        if p1.shape[0] > 5:
            assert min(np.abs(E - Eg).max(), np.abs(E + Eg).max()) < 1e-4, k
        assert abs(np.linalg.norm(E) - 1.0) < 1e-12
        # an essential matrix: two equal singular values and a zero one
        sv = np.linalg.svd(E, compute_uv=False)
        assert abs(sv[0] - sv[1]) < 1e-6 and sv[2] < 1e-6, (k, sv)


def test_five_point_solver_exact_geometry():
    """Noise-free normalised correspondences of a known motion: one of the models is [t]x R."""
    rng = np.random.default_rng(5)
    for trial in range(20):
        X = rng.uniform(-1, 1, (5, 3)) + np.array([0, 0, 5.0])
        a = 0.3 * rng.standard_normal(3)
        th = np.linalg.norm(a); kx = a / th
        Kx = np.array([[0, -kx[2], kx[1]], [kx[2], 0, -kx[0]], [-kx[1], kx[0], 0]])
        R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
        t = rng.standard_normal(3)
        Y = X @ R.T + t
        m1, m2 = X[:, :2] / X[:, 2:], Y[:, :2] / Y[:, 2:]
        Es = orc.five_point(m1, m2)
        tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
        Et = tx @ R; Et /= np.linalg.norm(Et)
        assert 1 <= len(Es) <= 10
        assert min(min(np.abs(E - Et).max(), np.abs(E + Et).max()) for E in Es) < 1e-8, trial
        for E in Es:      # every model satisfies the five epipolar constraints
            r = np.einsum("ni,ij,nj->n", np.c_[m2, np.ones(5)], E, np.c_[m1, np.ones(5)])
            assert np.abs(r).max() < 1e-9


def test_essential_small_and_degenerate(essential):
    cam = orc.Camera(1200, 1200, 1024, 768, 0, 0)
    ok, E, m, tr = orc.find_essential(np.zeros((4, 2), np.float32), np.ones((4, 2), np.float32), cam, cam)
    assert not ok and not E.any()
    same = np.tile(np.array([[700.0, 500.0]], np.float32), (30, 1))
    ok, E, m, tr = orc.find_essential(same, same, cam, cam)      # all-identical points: whatever comes back is finite
    assert np.isfinite(E).all()


def test_retrieval_restatement_degenerates_to_fake_img_matcher():
    """top_k >= n - 1 (or <= 0) gives FakeImgMatcher's list (every image with every other one, ImageMatcher.cpp:6-24) in
    the canonical order of the pair loop; a finite top_k gives a subset without self-pairs or duplicates."""
    from oracle import retrieval_ref
    from reconstructor_b200 import shard
    rng = np.random.default_rng(3)
    imgs = [rng.integers(0, 256, (40, 32), dtype=np.uint8) for _ in range(7)]
    allp = retrieval_ref.select_pairs(imgs, 0)
    assert np.array_equal(allp, shard.all_pairs(7)) and np.array_equal(retrieval_ref.select_pairs(imgs, 6), allp)
    p2 = retrieval_ref.select_pairs(imgs, 2)
    assert len(p2) < len(allp) and (p2[:, 0] < p2[:, 1]).all() and len({tuple(p) for p in p2.tolist()}) == len(p2)
    S = retrieval_ref.similarity(imgs)
    assert np.allclose(np.diag(S), 1.0) and np.allclose(S, S.T)
