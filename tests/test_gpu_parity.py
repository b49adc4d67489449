"""GPU parity tests: the CUDA path through the C ABI vs the oracle and the committed goldens.

Bar: bit-exact for integer / byte / index work (Hamming, integer-valued SIFT, match lists,
inlier masks); L2 distances of real-valued descriptors within 1e-5 relative (tolerance of the
north star), their indices equal to the fp64 brute force except at fp64 near-ties (< 1e-6 rel).
"""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import orc  # noqa: E402  (the checker)
from reconstructor_b200 import api, synth  # noqa: E402


@pytest.fixture(scope="module")
def knn(golden_dir):
    return np.load(os.path.join(golden_dir, "knn_pairs.npz"))


@pytest.fixture(scope="module")
def scenes(golden_dir):
    return np.load(os.path.join(golden_dir, "fmat_scenes.npz"))


@pytest.fixture(scope="module")
def fountain(golden_dir):
    return np.load(os.path.join(golden_dir, "fountain.npz"))


def _load3(pm, g, kind, ties=False):
    for i in range(3):
        d = g[f"{kind}_desc{i}"]
        if ties and i == 1:
            d = g[f"{kind}_desc1_ties"]
        if kind == "sift":
            d = d.astype(np.float32)
        pm.set_image(i, d, g[f"{kind}_xy{i}"])


# ---------------------------------------------------------------------------------------------
# kNN
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", [0, 2])
def test_hamming_knn_bit_exact(knn, variant):
    with api.PairMatcher(debug_flags=variant) as pm:
        _load3(pm, knn, "orb")
        for tag, a, b in (("01", 0, 1), ("02", 0, 2), ("12", 1, 2)):
            idx, dist = pm.knn_pair(a, b)
            assert np.array_equal(idx, knn[f"orb_{tag}_idx"])
            assert np.array_equal(dist, knn[f"orb_{tag}_dist"])


@pytest.mark.parametrize("force_simt", [0, 1, 4, 8, 12, 20, 28])
def test_l2_sift_knn_bit_exact(knn, force_simt):
    """flags 0/4/8/12 are the single-CTA tcgen05 path (four epilogue variants), 20 the CTA-pair kernel, 1 the fp32 SIMT kernel; all must
    equal cv::BFMatcher."""
    with api.PairMatcher(debug_flags=force_simt) as pm:
        _load3(pm, knn, "sift")
        for tag, a, b in (("01", 0, 1), ("02", 0, 2), ("12", 1, 2)):
            idx, dist = pm.knn_pair(a, b)
            bad = np.nonzero((idx != knn[f"sift_{tag}_idx"]).any(axis=1))[0]
            assert len(bad) == 0, (tag, bad[:10], idx[bad[:5]], knn[f"sift_{tag}_idx"][bad[:5]])
            assert np.array_equal(dist, knn[f"sift_{tag}_dist"])


def test_l2_sift_u8_dtype_matches_f32(knn):
    with api.PairMatcher() as pm:
        for i in range(3):
            pm.set_image(i, knn[f"sift_desc{i}"], knn[f"sift_xy{i}"], dtype=api.DESC_U8)
        idx, dist = pm.knn_pair(0, 1)
        assert np.array_equal(idx, knn["sift_01_idx"]) and np.array_equal(dist, knn["sift_01_dist"])


@pytest.mark.parametrize("kind,flags", [("orb", 0), ("sift", 0), ("sift", 1), ("sift", 4), ("sift", 12), ("sift", 20), ("sift", 28)])
def test_knn_ties_lowest_index(knn, kind, flags):
    with api.PairMatcher(debug_flags=flags) as pm:
        _load3(pm, knn, kind, ties=True)
        idx, dist = pm.knn_pair(0, 1)
        assert np.array_equal(idx, knn[f"{kind}_ties_idx"])
        assert np.array_equal(dist, knn[f"{kind}_ties_dist"])
        assert tuple(idx[10]) == (5, 77) and tuple(idx[20]) == (31, 300)


def test_l2_superpoint_within_tolerance(knn):
    with api.PairMatcher() as pm:
        _load3(pm, knn, "superpoint")
        for tag, a, b in (("01", 0, 1), ("12", 1, 2)):
            idx, dist = pm.knn_pair(a, b)
            oi, o2 = orc.knn2_l2(knn[f"superpoint_desc{a}"], knn[f"superpoint_desc{b}"])
            np.testing.assert_allclose(dist, np.sqrt(o2), rtol=1e-5)          # north-star tolerance
            np.testing.assert_allclose(dist, knn[f"superpoint_{tag}_dist"], rtol=1e-5)
            for r in np.nonzero((idx != oi).any(axis=1))[0]:                    # only fp64 near-ties
                assert abs(o2[r, 1] - o2[r, 0]) <= 1e-6 * o2[r, 1]


def test_tc_accumulators_exact(knn):
    """The raw TMEM accumulators of the first 256x128 block equal nb - 2 a.b as integers."""
    a = knn["sift_desc0"].astype(np.int64); b = knn["sift_desc1"].astype(np.int64)
    with api.PairMatcher() as pm:
        _load3(pm, knn, "sift")
        idx, dist, acc = pm.debug_tc_dump(0, 1)
    want = (b[:128] ** 2).sum(1)[None, :] - 2 * (a[:256] @ b[:128].T)
    got = acc.astype(np.int64)
    bad = np.argwhere(got != want)
    assert len(bad) == 0, (len(bad), bad[:8], got[tuple(bad[0])], want[tuple(bad[0])])


def test_l2_sift_extreme_values_exact():
    """All-255 / all-0 rows: 2*a.b reaches 16,646,400 < 2^24 -- the fp32 accumulator limit."""
    rng = np.random.default_rng(3)
    q = rng.integers(0, 256, (300, 128)).astype(np.float32)
    t = rng.integers(0, 256, (260, 128)).astype(np.float32)
    q[:40] = 255; t[:30] = 255; q[40:60] = 0; t[30:50] = 0
    t[100] = q[7]
    oi, o2 = orc.knn2_l2(q, t)
    with api.PairMatcher() as pm:
        pm.set_image(0, q); pm.set_image(1, t)
        idx, dist = pm.knn_pair(0, 1)
    assert np.array_equal(idx, oi)
    assert np.array_equal(dist, np.sqrt(o2).astype(np.float32))


@pytest.mark.parametrize("kind", ["orb", "sift", "sift-epi1", "sift-epi2", "sift-epi3", "sift-pair", "sift-pair192", "superpoint"])
def test_knn_ragged_and_tiny(kind):
    """Ragged sizes (not multiples of any tile), 1-row and 2-row train sets, empty images."""
    flags = {"sift-epi1": 4, "sift-epi2": 8, "sift-epi3": 12, "sift-pair": 20, "sift-pair192": 28}.get(kind, 0)
    kind = kind.split("-")[0]
    w = synth.World(kind, 700, seed=11)
    full = [w.image(i, 4)[0] for i in range(2)]
    cases = [(1, 1), (5, 1), (3, 2), (129, 127), (257, 300), (700, 513), (511, 700), (300, 65), (260, 190)]
    with api.PairMatcher(debug_flags=flags) as pm:
        for nq, nt in cases:
            q, t = full[0][:nq], full[1][:nt]
            pm.set_image(0, q); pm.set_image(1, t)
            idx, dist = pm.knn_pair(0, 1)
            if kind == "orb":
                oi, od = orc.knn2_hamming(q, t)
                od = od.astype(np.float32)
                od[oi < 0] = np.inf
                assert np.array_equal(idx, oi) and np.array_equal(dist, od), (nq, nt)
            else:
                oi, o2 = orc.knn2_l2(q, t)
                od = np.sqrt(o2).astype(np.float32)
                if kind == "sift":
                    assert np.array_equal(idx, oi) and np.array_equal(dist, od), (nq, nt)
                else:
                    assert np.array_equal(idx, oi), (nq, nt)
                    np.testing.assert_allclose(dist, od, rtol=1e-5)
        # empty query image: nothing to do, empty train image: no neighbours
        pm.set_image(2, full[0][:0]); pm.set_image(3, full[1][:10])
        r = pm.match_pair(2, 3)
        assert len(r["q"]) == 0
        pm.set_image(4, full[0][:10]); pm.set_image(5, full[1][:0])
        idx, dist = pm.knn_pair(4, 5)
        assert (idx == -1).all() and np.isinf(dist).all()
        assert len(pm.match_pair(4, 5)["q"]) == 0


# ---------------------------------------------------------------------------------------------
# ratio + uniqueness
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["orb", "sift", "superpoint"])
@pytest.mark.parametrize("mode", [api.UNIQUE_FIRST_WINS, api.MUTUAL_NN, api.UNIQUE_NONE])
def test_ratio_unique_modes(knn, kind, mode):
    d0, d1 = knn[f"{kind}_desc0"], knn[f"{kind}_desc1"]
    if kind != "orb":
        d0, d1 = d0.astype(np.float32), d1.astype(np.float32)
    if kind == "orb":
        oi, od = orc.knn2_hamming(d0, d1); od = od.astype(np.float32)
    else:
        oi, o2 = orc.knn2_l2(d0, d1); od = np.sqrt(o2).astype(np.float32)
    bq = orc.best_query(d0, d1) if mode == api.MUTUAL_NN else None
    wq, wt = orc.ratio_unique(oi, od, d1.shape[0], 0.7, mode, bq)
    with api.PairMatcher(unique_mode=mode) as pm:
        pm.set_image(0, d0); pm.set_image(1, d1)
        r = pm.match_pair(0, 1)
    assert np.array_equal(r["q"], wq) and np.array_equal(r["t"], wt)
    assert len(wq) > 20


# ---------------------------------------------------------------------------------------------
# epipolar filter
# ---------------------------------------------------------------------------------------------
def test_fmat_masks_identical_to_cv2_and_oracle(scenes):
    n_sc = int(scenes["n_scenes"])
    checked = 0
    with api.PairMatcher() as pm:
        for k in range(n_sc):
            p1, p2 = scenes[f"s{k}_p1"], scenes[f"s{k}_p2"]
            n = p1.shape[0]
            F, mask, st, it = pm.estimate_fundamental(p1, p2)
            ns, Fo, mo, tr = orc.find_fundamental(p1, p2)
            # GPU <-> repo CPU filter: identical in every regime
            assert (st == api.PAIR_FILTERED) == (ns > 0), k
            if ns > 0:
                assert np.array_equal(mask, mo), (k, n, int(mask.sum()), int(mo.sum()))
                if n > 7:
                    assert it == tr.iters_run, (k, it, tr.iters_run)
                    np.testing.assert_allclose(F, Fo[0], rtol=0, atol=1e-9 * max(1.0, np.abs(Fo[0]).max()))
            # GPU <-> cv2.findFundamentalMat for N >= 15 (RANSAC regime of OpenCV)
            if n >= 15 and int(scenes[f"s{k}_ok"]):
                assert np.array_equal(mask, scenes[f"s{k}_mask"]), k
                checked += 1
    assert checked >= 40


def test_fmat_degenerate_and_small(scenes):
    with api.PairMatcher() as pm:
        for name in ("same", "line"):
            F, mask, st, it = pm.estimate_fundamental(scenes[f"deg_{name}_p1"], scenes[f"deg_{name}_p2"])
            assert st == api.PAIR_DROPPED and not mask.any() and not F.any()
        F, mask, st, it = pm.estimate_fundamental(np.zeros((6, 2), np.float32), np.ones((6, 2), np.float32))
        assert st == api.PAIR_DROPPED
        F, mask, st, it = pm.estimate_fundamental(np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32))
        assert st == api.PAIR_DROPPED


def test_fmat_sampson_mode_matches_oracle(scenes):
    prm = orc.default_params(residual_mode=orc.RESID_SAMPSON)
    with api.PairMatcher(residual_mode=api.RESID_SAMPSON) as pm:
        for k in (20, 24, 30, 38, 44, 50):
            p1, p2 = scenes[f"s{k}_p1"], scenes[f"s{k}_p2"]
            F, mask, st, it = pm.estimate_fundamental(p1, p2)
            ns, Fo, mo, tr = orc.find_fundamental(p1, p2, prm)
            assert (st == api.PAIR_FILTERED) == (ns > 0)
            assert np.array_equal(mask, mo), k


# ---------------------------------------------------------------------------------------------
# whole pair body / batched loop
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["orb", "sift", "superpoint"])
def test_pair_body_matches_cv2_golden(knn, kind):
    with api.PairMatcher() as pm:
        _load3(pm, knn, kind)
        for tag, a, b in (("01", 0, 1), ("02", 0, 2), ("12", 1, 2)):
            r = pm.match_filter_pair(a, b)
            assert len(r["q"]) == int(knn[f"{kind}_{tag}_nput"])
            assert (r["status"] != api.PAIR_DROPPED) == bool(knn[f"{kind}_{tag}_status"])
            keep = r["inlier"].astype(bool)
            assert np.array_equal(r["q"][keep], knn[f"{kind}_{tag}_q"])
            assert np.array_equal(r["t"][keep], knn[f"{kind}_{tag}_t"])


def test_fountain_real_images(fountain):
    """SIFT features of the reference's own sample images data/0000-0002.jpg."""
    with api.PairMatcher() as pm:
        for i in range(3):
            pm.set_image(i, fountain[f"desc{i}"].astype(np.float32), fountain[f"xy{i}"])
        res = pm.match_all_pairs()
        fm = api.feature_matches_view(res)
        for p, (tag, a, b) in enumerate((("01", 0, 1), ("02", 0, 2), ("12", 1, 2))):
            assert tuple(res["pair_ij"][p]) == (a, b)
            idx, dist = pm.knn_pair(a, b)
            assert np.array_equal(idx, fountain[f"{tag}_idx"]) and np.array_equal(dist, fountain[f"{tag}_dist"])
            want = dict(zip(fountain[f"{tag}_q"].tolist(), fountain[f"{tag}_t"].tolist()))
            assert fm[(a, b)] == want
            assert fm[(b, a)] == {t: q for q, t in want.items()}


@pytest.mark.parametrize("kind,n_img,n_kp", [("orb", 6, 700), ("sift", 6, 650), ("superpoint", 4, 400)])
def test_match_all_pairs_equals_oracle(kind, n_img, n_kp):
    w = synth.World(kind, n_kp, seed=21)
    imgs = []
    for i in range(n_img):
        d, xy, _ = w.image(i, n_img, outlier_frac=0.3 if i % 2 else 0.0)
        cut = n_kp - 37 * i                       # ragged keypoint counts
        imgs.append((d[:cut], xy[:cut]))
    with api.PairMatcher(batch_pairs=4) as pm:    # several batches in flight
        for i, (d, xy) in enumerate(imgs):
            pm.set_image(i, d, xy)
        res = pm.match_all_pairs()
        assert res["n_pairs"] == n_img * (n_img - 1) // 2
        for p, (i, j) in enumerate(res["pair_ij"]):
            ref = orc.match_pair(imgs[i][0], imgs[i][1], imgs[j][0], imgs[j][1])
            a, b = res["offsets"][p], res["offsets"][p + 1]
            assert b - a == ref["n_putative"], (p, b - a, ref["n_putative"])
            assert (res["status"][p] == api.PAIR_DROPPED) == (ref["status"] == "dropped")
            keep = res["inlier"][a:b].astype(bool)
            assert np.array_equal(res["q"][a:b][keep], ref["q"]), p
            assert np.array_equal(res["t"][a:b][keep], ref["t"]), p
        # explicit pair list incl. a reversed pair and a repeated pair
        sub = np.array([[2, 1], [0, 3], [0, 3]], np.int32)
        r2 = pm.match_all_pairs(sub)
        ref = orc.match_pair(imgs[2][0], imgs[2][1], imgs[1][0], imgs[1][1])
        a, b = r2["offsets"][0], r2["offsets"][1]
        assert np.array_equal(r2["q"][a:b][r2["inlier"][a:b].astype(bool)], ref["q"])
        a1, b1, b2 = r2["offsets"][1], r2["offsets"][2], r2["offsets"][3]
        assert np.array_equal(r2["q"][a1:b1], r2["q"][b1:b2])


def test_min_matches_gate_and_unfiltered_branch():
    """< 7 putative matches: everything is kept unfiltered (SequentialReconstructor.cpp:270-276)."""
    w = synth.World("orb", 400, seed=5)
    d0, xy0, _ = w.image(0, 2); d1, xy1, _ = w.image(1, 2)
    with api.PairMatcher() as pm:
        pm.set_image(0, d0[:12], xy0[:12]); pm.set_image(1, d1, xy1)
        r = pm.match_filter_pair(0, 1)
        ref = orc.match_pair(d0[:12], xy0[:12], d1, xy1)
        assert len(r["q"]) == ref["n_putative"] < 7
        assert r["status"] == api.PAIR_UNFILTERED and r["inlier"].all()
        assert np.array_equal(r["q"], ref["q"])
    with api.PairMatcher(do_filter=0) as pm:
        pm.set_image(0, d0, xy0); pm.set_image(1, d1, xy1)
        r = pm.match_filter_pair(0, 1)
        assert r["status"] == api.PAIR_UNFILTERED and r["inlier"].all() and len(r["q"]) > 20


def test_match_descriptors_compat_call(knn):
    d0, d1 = knn["orb_desc0"], knn["orb_desc1"]
    with api.PairMatcher() as pm:
        r = pm.match_descriptors(d0, d1)
        oi, od = orc.knn2_hamming(d0, d1)
        wq, wt = orc.ratio_unique(oi, od.astype(np.float32), d1.shape[0])
        assert np.array_equal(r["q"], wq) and np.array_equal(r["t"], wt)
        r2 = pm.match_descriptors(d1[:100], d0)           # smaller re-use of the temp slots
        oi, od = orc.knn2_hamming(d1[:100], d0)
        wq, wt = orc.ratio_unique(oi, od.astype(np.float32), d0.shape[0])
        assert np.array_equal(r2["q"], wq) and np.array_equal(r2["t"], wt)


def test_errors_are_reported_not_thrown(knn):
    with api.PairMatcher() as pm:
        pm.set_image(0, knn["orb_desc0"])
        with pytest.raises(api.PairMatchError) as e:
            pm.set_image(1, knn["sift_desc1"].astype(np.float32))      # shape/dtype mismatch
        assert e.value.code == api.ERR_INVALID
        with pytest.raises(api.PairMatchError) as e:
            pm._n[9] = 4
            pm.knn_pair(0, 9)                                           # image not set
        assert e.value.code == api.ERR_STATE


# ---------------------------------------------------------------------------------------------
# full size (BASELINE.json configs): oracle on one pair + size-independent properties
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["sift", "orb"])
def test_full_size_8192_pair(kind):
    w = synth.World(kind, 8192, seed=0xB200)
    imgs = [w.image(i, 100)[:2] for i in range(3)]
    flags = [0, 1, 4, 8, 12, 20, 28] if kind == "sift" else [0, 2]
    outs = []
    for f in flags:
        with api.PairMatcher(debug_flags=f) as pm:
            for i, (d, xy) in enumerate(imgs):
                pm.set_image(i, d, xy)
            outs.append(pm.knn_pair(0, 1))
            if f == 0:
                res = pm.match_all_pairs()
    # tensor path == SIMT path (sift) / popc == carry-save popc (orb), bit for bit
    for o in outs[1:]:
        assert np.array_equal(outs[0][0], o[0]) and np.array_equal(outs[0][1], o[1])
    # against the oracle on the same pair
    if kind == "orb":
        oi, od = orc.knn2_hamming(imgs[0][0], imgs[1][0]); od = od.astype(np.float32)
    else:
        oi, o2 = orc.knn2_l2(imgs[0][0], imgs[1][0]); od = np.sqrt(o2).astype(np.float32)
    assert np.array_equal(outs[0][0], oi) and np.array_equal(outs[0][1], od)
    ref = orc.match_pair(imgs[0][0], imgs[0][1], imgs[1][0], imgs[1][1])
    a, b = res["offsets"][0], res["offsets"][1]
    keep = res["inlier"][a:b].astype(bool)
    assert b - a == ref["n_putative"]
    assert np.array_equal(res["q"][a:b][keep], ref["q"]) and np.array_equal(res["t"][a:b][keep], ref["t"])
    # properties: ascending unique queries, unique trains, planted overlap recovered
    for p in range(3):
        a, b = res["offsets"][p], res["offsets"][p + 1]
        q, t = res["q"][a:b], res["t"][a:b]
        assert (np.diff(q) > 0).all() and len(np.unique(t)) == len(t)
        assert b - a > 1500 and res["n_inliers"][p] > 0.85 * (b - a)


# ---------------------------------------------------------------------------------------------
# values-only tensor epilogue + fix-up (the default batched path) vs the general kernels
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("flags", [0, 2048, 128, 256])
@pytest.mark.parametrize("mode", [api.UNIQUE_FIRST_WINS, api.MUTUAL_NN, api.UNIQUE_NONE])
def test_batched_fast_path_equals_oracle_all_modes(mode, flags):
    w = synth.World("sift", 900, seed=33)
    imgs = []
    for i in range(5):
        d, xy, _ = w.image(i, 5, outlier_frac=0.2)
        cut = 900 - 61 * i
        d = d[:cut].copy(); xy = xy[:cut]
        if i == 1:                       # duplicates: equal distances inside one 32-column chunk and across chunks
            d[40] = d[7]; d[300] = d[7]; d[301] = d[7]
        imgs.append((d, xy))
    with api.PairMatcher(unique_mode=mode, batch_pairs=3, debug_flags=flags) as pm:
        for i, (d, xy) in enumerate(imgs):
            pm.set_image(i, d, xy)
        res = pm.match_all_pairs()
    for p, (i, j) in enumerate(res["pair_ij"]):
        ref = orc.match_pair(imgs[i][0], imgs[i][1], imgs[j][0], imgs[j][1], unique_mode=mode)
        a, b = res["offsets"][p], res["offsets"][p + 1]
        assert b - a == ref["n_putative"], (mode, p, b - a, ref["n_putative"])
        keep = res["inlier"][a:b].astype(bool)
        assert np.array_equal(res["q"][a:b][keep], ref["q"]) and np.array_equal(res["t"][a:b][keep], ref["t"]), (mode, p)


def test_batched_fast_path_equals_general_kernel_full_size():
    """8192-keypoint images: the values-only kernel + fix-up must reproduce the general tensor kernel
    and the SIMT kernel bit for bit (CSR arrays identical)."""
    w = synth.World("sift", 8192, seed=0xB200 + 2)
    imgs = [w.image(i, 100, outlier_frac=0.3 if i == 2 else 0.0)[:2] for i in range(6)]
    outs = []
    # values-only kernels (byte form on kind::i8 = default, fp16 form: CTA pair 256-col, single-CTA, pair 192-col),
    # general tensor kernels (single-CTA, CTA pair), fp32 SIMT
    # + byte form with one query row set per cluster (16384) and its 64-register build
    for flags in (0, 2048, 16384, 16384 + 12288, 12288, 128, 256, 64, 64 + 20, 1):
        with api.PairMatcher(debug_flags=flags) as pm:
            for i, (d, xy) in enumerate(imgs):
                pm.set_image(i, d, xy)
            outs.append(pm.match_all_pairs())
    for o in outs[1:]:
        for k in ("offsets", "q", "t", "inlier", "status", "n_inliers", "ransac_iters"):
            assert np.array_equal(outs[0][k], o[k]), k
        assert np.array_equal(outs[0]["F"], o["F"])
    assert outs[0]["offsets"][-1] > 15 * 1500


def test_multi_device_handle_matches_single_device():
    """In-process multi-GPU: pairs dealt to one host thread per device, descriptors replicated."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    w = synth.World("orb", 1200, seed=9)
    imgs = [w.image(i, 8)[:2] for i in range(8)]
    outs = []
    for devs in ([0], [0, 1]):
        with api.PairMatcher(devices=devs, batch_pairs=5) as pm:
            for i, (d, xy) in enumerate(imgs):
                pm.set_image(i, d, xy)
            outs.append(pm.match_all_pairs())
    for k in ("pair_ij", "offsets", "q", "t", "inlier", "status", "n_inliers"):
        assert np.array_equal(outs[0][k], outs[1][k]), k


# ---------------------------------------------------------------------------------------------
# real-valued rows on the tensor cores (fp16 scores + certified exact fp32 re-rank) vs the SIMT kernel
# ---------------------------------------------------------------------------------------------
def _unit_rows(rng, n, dim):
    x = rng.standard_normal((n, dim)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True).astype(np.float32)
    x *= np.float32(0.999)                      # |x|^2 <= 1 after fp32 rounding
    return x


@pytest.mark.parametrize("dim", [128, 256])
def test_float_tensor_knn_equals_simt_bit_exact(dim):
    """pm_knn_pair through MODE 3 + L2F_NEED_FULL equals the fp32 SIMT kernel bit for bit, incl. planted
    duplicates (ties -> lowest index), ragged sizes and single-chunk train sets."""
    rng = np.random.default_rng(dim)
    base = _unit_rows(rng, 3000, dim)
    q = base[:1500].copy()
    t = (base[700:2900] + 0.02 * rng.standard_normal((2200, dim)).astype(np.float32)).astype(np.float32)
    t /= (np.linalg.norm(t, axis=1, keepdims=True) * 1.001).astype(np.float32)
    t[5] = t[900]; t[1700] = t[900]; t[901] = t[900]          # exact ties inside and across chunks
    t[40] = q[3]; t[41] = q[3]                                  # zero distance, tie within a chunk
    cases = [(1500, 2200), (257, 300), (129, 17), (5, 1), (3, 2), (700, 16), (64, 2049)]
    got = {}
    for flags in (0, 1):
        with api.PairMatcher(debug_flags=flags) as pm:
            for nq, nt in cases:
                pm.set_image(0, q[:nq]); pm.set_image(1, t[:nt])
                got[(flags, nq, nt)] = pm.knn_pair(0, 1)
            st = pm.stats()
        if flags == 0:
            assert st["rerank_rows"] > 0, "tensor path did not run"
            assert st["rerank_worst_err"] < 1.0, st
        else:
            assert st["rerank_rows"] == 0
    for nq, nt in cases:
        a, b = got[(0, nq, nt)], got[(1, nq, nt)]
        bad = np.nonzero((a[0] != b[0]).any(axis=1))[0]
        assert len(bad) == 0, (nq, nt, bad[:8], a[0][bad[:4]], b[0][bad[:4]])
        assert np.array_equal(a[1], b[1]), (nq, nt)
    oi, o2 = orc.knn2_l2(q, t)
    idx, dist = got[(0, 1500, 2200)]
    np.testing.assert_allclose(dist, np.sqrt(o2), rtol=1e-5)
    for r in np.nonzero((idx != oi).any(axis=1))[0]:
        assert abs(o2[r, 1] - o2[r, 0]) <= 1e-6 * o2[r, 1] or o2[r, 0] == 0


@pytest.mark.parametrize("mode", [api.UNIQUE_FIRST_WINS, api.MUTUAL_NN, api.UNIQUE_NONE])
def test_float_tensor_batched_equals_simt_all_modes(mode):
    w = synth.World("superpoint", 1100, seed=77)
    imgs = []
    for i in range(5):
        d, xy, _ = w.image(i, 5, outlier_frac=0.2)
        cut = 1100 - 83 * i
        d = d[:cut].copy(); xy = xy[:cut]
        if i == 1:
            d[40] = d[7]; d[300] = d[7]; d[301] = d[7]
        if i == 3:
            d[:] = d[0]                         # every row identical: all distances tie (certified to fail the ratio test)
        imgs.append((d, xy))
    outs = []
    for flags in (0, 1):
        with api.PairMatcher(unique_mode=mode, batch_pairs=3, debug_flags=flags) as pm:
            for i, (d, xy) in enumerate(imgs):
                pm.set_image(i, d, xy)
            outs.append(pm.match_all_pairs())
            st = pm.stats()
            assert (st["rerank_rows"] > 0) == (flags == 0)
            assert st["rerank_worst_err"] < 1.0
    for k in ("offsets", "q", "t", "inlier", "status", "n_inliers", "ransac_iters", "F"):
        assert np.array_equal(outs[0][k], outs[1][k]), (mode, k)
    assert outs[0]["offsets"][-1] > 500


def test_float_tensor_overflow_scan_and_fallbacks():
    """Identical rows force the exhaustive scan in FULL mode; non-unit and non-finite images use the SIMT kernel."""
    rng = np.random.default_rng(5)
    q = _unit_rows(rng, 300, 256)
    t = np.repeat(_unit_rows(rng, 1, 256), 400, axis=0)         # 25 chunks with the same score
    with api.PairMatcher() as pm:
        pm.set_image(0, q); pm.set_image(1, t)
        idx, dist = pm.knn_pair(0, 1)
        st = pm.stats()
    assert st["rerank_overflow"] == 300
    assert (idx[:, 0] == 0).all() and (idx[:, 1] == 1).all()
    with api.PairMatcher(debug_flags=1) as pm:
        pm.set_image(0, q); pm.set_image(1, t)
        i2, d2 = pm.knn_pair(0, 1)
    assert np.array_equal(idx, i2) and np.array_equal(dist, d2)
    big = (3.0 * _unit_rows(rng, 200, 256)).astype(np.float32)
    with api.PairMatcher() as pm:
        pm.set_image(0, q); pm.set_image(1, big)
        idx, dist = pm.knn_pair(0, 1)
        assert pm.stats()["rerank_rows"] == 0                   # |x|^2 = 9 > bound: SIMT kernel
    oi, o2 = orc.knn2_l2(q, big)
    assert np.array_equal(idx, oi)
    np.testing.assert_allclose(dist, np.sqrt(o2), rtol=1e-5)


def test_float_tensor_full_size_superpoint():
    """8192 x 8192 x 256-d (BASELINE config #4 shape): tensor path == SIMT path bit for bit, CSR identical."""
    w = synth.World("superpoint", 8192, seed=0xB200 + 4)
    imgs = [w.image(i, 100, outlier_frac=0.3 if i == 2 else 0.0)[:2] for i in range(4)]
    outs, knn_rows = [], []
    for flags in (0, 1):
        with api.PairMatcher(debug_flags=flags) as pm:
            for i, (d, xy) in enumerate(imgs):
                pm.set_image(i, d, xy)
            knn_rows.append(pm.knn_pair(0, 1))
            outs.append(pm.match_all_pairs())
            st = pm.stats()
            if flags == 0:
                assert st["rerank_rows"] > 0 and st["rerank_worst_err"] < 1.0, st
                print("rerank stats", {k: st[k] for k in st if k.startswith("rerank")})
    assert np.array_equal(knn_rows[0][0], knn_rows[1][0]) and np.array_equal(knn_rows[0][1], knn_rows[1][1])
    for k in ("offsets", "q", "t", "inlier", "status", "n_inliers", "ransac_iters", "F"):
        assert np.array_equal(outs[0][k], outs[1][k]), k
    oi, o2 = orc.knn2_l2(imgs[0][0], imgs[1][0])
    np.testing.assert_allclose(knn_rows[0][1], np.sqrt(o2), rtol=1e-5)
    for r in np.nonzero((knn_rows[0][0] != oi).any(axis=1))[0]:
        assert abs(o2[r, 1] - o2[r, 0]) <= 1e-6 * o2[r, 1]
    assert outs[0]["offsets"][-1] > 6 * 1500
    # the batched CSR of full-size pairs against the oracle's pair body (fp64 arbiter for the kNN stage), one clean
    # pair and one with 30 % outliers
    for p, (i, j) in enumerate(outs[0]["pair_ij"]):
        if (i, j) not in ((0, 1), (1, 2)):
            continue
        ref = orc.match_pair(imgs[i][0], imgs[i][1], imgs[j][0], imgs[j][1])
        a, b = outs[0]["offsets"][p], outs[0]["offsets"][p + 1]
        keep = outs[0]["inlier"][a:b].astype(bool)
        assert ref["status"] == "ok" and outs[0]["status"][p] == api.PAIR_FILTERED
        assert b - a == ref["n_putative"], (i, j, b - a, ref["n_putative"])
        assert np.array_equal(outs[0]["q"][a:b][keep], ref["q"]) and np.array_equal(outs[0]["t"][a:b][keep], ref["t"])


# ---------------------------------------------------------------------------------------------
# binary descriptors on the tensor cores (E4M3 {0,1} operands, exact) vs the XOR/popc kernel and the oracle
# ---------------------------------------------------------------------------------------------
FORCE_POPC = 1 << 10
FORCE_E4M3_512 = 1 << 16          # binary rows: the kind::f8f6f4 kernel (E4M3 forms) instead of the kind::mxf4 one


@pytest.mark.parametrize("mode", [api.UNIQUE_FIRST_WINS, api.MUTUAL_NN, api.UNIQUE_NONE])
def test_hamming_tensor_batched_equals_popc_and_oracle(mode):
    w = synth.World("orb", 1000, seed=55)
    imgs = []
    for i in range(5):
        d, xy, _ = w.image(i, 5, outlier_frac=0.2)
        cut = 1000 - 71 * i
        d = d[:cut].copy(); xy = xy[:cut]
        if i == 1:                       # exact duplicates inside one 16-column chunk and across chunks
            d[40] = d[7]; d[300] = d[7]; d[301] = d[7]
        if i == 2:
            d[5] = 0; d[6] = 255         # all-zero / all-one rows (|b| = 0 and 256)
        imgs.append((d, xy))
    outs = []
    for flags in (0, FORCE_POPC):
        with api.PairMatcher(unique_mode=mode, batch_pairs=3, debug_flags=flags) as pm:
            for i, (d, xy) in enumerate(imgs):
                pm.set_image(i, d, xy)
            outs.append(pm.match_all_pairs())
    for k in ("offsets", "q", "t", "inlier", "status", "n_inliers", "ransac_iters", "F"):
        assert np.array_equal(outs[0][k], outs[1][k]), (mode, k)
    res = outs[0]
    for p, (i, j) in enumerate(res["pair_ij"]):
        ref = orc.match_pair(imgs[i][0], imgs[i][1], imgs[j][0], imgs[j][1], unique_mode=mode)
        a, b = res["offsets"][p], res["offsets"][p + 1]
        assert b - a == ref["n_putative"], (mode, p, b - a, ref["n_putative"])
        keep = res["inlier"][a:b].astype(bool)
        assert np.array_equal(res["q"][a:b][keep], ref["q"]) and np.array_equal(res["t"][a:b][keep], ref["t"]), (mode, p)


def test_hamming_tensor_512_bit_and_ragged():
    rng = np.random.default_rng(12)
    base = rng.integers(0, 256, (1500, 64), dtype=np.uint8)
    imgs = []
    for i, n in enumerate((700, 513, 129, 17, 1)):
        ids = rng.permutation(1500)[:n]
        d = base[ids].copy()
        flip = rng.random((n, 64)) < 0.04
        d ^= (flip * rng.integers(1, 256, (n, 64))).astype(np.uint8)
        imgs.append(d)
    outs = []
    for flags in (0, FORCE_POPC):
        with api.PairMatcher(debug_flags=flags, do_filter=0, batch_pairs=4) as pm:
            for i, d in enumerate(imgs):
                pm.set_image(i, d)
            outs.append(pm.match_all_pairs())
    for k in ("offsets", "q", "t", "status"):
        assert np.array_equal(outs[0][k], outs[1][k]), k
    res = outs[0]
    for p, (i, j) in enumerate(res["pair_ij"]):
        oi, od = orc.knn2_hamming(imgs[i], imgs[j])
        wq, wt = orc.ratio_unique(oi, od.astype(np.float32), imgs[j].shape[0])
        a, b = res["offsets"][p], res["offsets"][p + 1]
        assert np.array_equal(res["q"][a:b], wq) and np.array_equal(res["t"][a:b], wt), (i, j)
    assert res["offsets"][-1] > 300


def test_hamming_512_bit_fp4_form_every_popcount_and_sizes():
    """512-bit rows as E2M1 values on kind::mxf4 (two K atoms, 9 K-steps; the default since round 2) vs the E4M3 form
    (17 K-steps) vs the XOR/popc kernel: image sizes around the row-set / tile / chunk boundaries, and train rows of
    EVERY popcount 0..512 (the norm block encodes |b| = 36 n + 6 u + v in E2M1 digits)."""
    rng = np.random.default_rng(21)
    pops = np.concatenate([np.arange(513), rng.integers(0, 513, 400)])
    every = np.zeros((len(pops), 512), np.uint8)
    for r, c in enumerate(pops):
        every[r, rng.permutation(512)[:c]] = 1
    every = np.packbits(every, axis=1, bitorder="little")
    base = rng.integers(0, 256, (1400, 64), dtype=np.uint8)
    imgs = [every]
    for n in (769, 513, 512, 193, 192, 97, 33):
        ids = rng.permutation(1400)[:n]
        d = base[ids].copy()
        d ^= ((rng.random((n, 64)) < 0.04) * rng.integers(1, 256, (n, 64))).astype(np.uint8)
        imgs.append(d)
    imgs.append(every[rng.permutation(len(every))[:600]] ^ np.uint8(1))     # near-duplicates of the popcount rows
    outs = []
    for flags in (0, FORCE_E4M3_512, FORCE_POPC):
        with api.PairMatcher(debug_flags=flags, do_filter=0, batch_pairs=5, unique_mode=api.UNIQUE_NONE) as pm:
            for i, d in enumerate(imgs):
                pm.set_image(i, d)
            outs.append(pm.match_all_pairs())
    for o in outs[1:]:
        for k in ("offsets", "q", "t", "status"):
            assert np.array_equal(outs[0][k], o[k]), k
    res = outs[0]
    for p in (0, 7, len(res["pair_ij"]) - 1):
        i, j = res["pair_ij"][p]
        oi, od = orc.knn2_hamming(imgs[i], imgs[j])
        wq, wt = orc.ratio_unique(oi, od.astype(np.float32), imgs[j].shape[0], mode=orc.UNIQUE_NONE)
        a, b = res["offsets"][p], res["offsets"][p + 1]
        assert np.array_equal(res["q"][a:b], wq) and np.array_equal(res["t"][a:b], wt), (i, j)
    assert res["offsets"][-1] > 1000


def test_hamming_tensor_full_size_equals_popc():
    w = synth.World("orb", 8192, seed=0xB200 + 3)
    imgs = [w.image(i, 100, outlier_frac=0.3 if i == 2 else 0.0)[:2] for i in range(5)]
    outs = []
    for flags in (0, FORCE_POPC, 1 << 26):
        with api.PairMatcher(debug_flags=flags) as pm:
            for i, (d, xy) in enumerate(imgs):
                pm.set_image(i, d, xy)
            outs.append(pm.match_all_pairs())
    for o in outs[1:]:
        for k in ("offsets", "q", "t", "inlier", "status", "n_inliers", "ransac_iters", "F"):
            assert np.array_equal(outs[0][k], o[k]), k
    assert outs[0]["offsets"][-1] > 10 * 1500


FORCE_E4M3 = 1 << 16
FORCE_I8 = 1 << 18
FORCE_FP4_UNPACKED = 1 << 26      # kind::mxf4 with one train row per accumulator column (default: packed pairs)


@pytest.mark.parametrize("mode", [api.UNIQUE_FIRST_WINS, api.MUTUAL_NN])
def test_hamming_i8_two_set_kernel_sizes_around_row_sets(mode):
    """256-bit rows as E2M1 values on kind::mxf4 (default: two train rows per accumulator column, scale factors 1 and
    2^10; also the one-row form) vs bytes on kind::i8 vs the E4M3 form vs the XOR/popc kernel, with image sizes on both
    sides of the 128 / 256 / 512-row boundaries of a work item (one or two resident query row sets), of the 192 / 256 /
    384-row train tiles (packed form: 192 + 192, second half absent or ragged) and of the 32-column fix-up chunks,
    duplicates inside and across chunks, all-zero and all-one rows (Hamming distance 256 in either packed field), and one
    image holding every popcount 0..256 (all digits of the norm block)."""
    rng = np.random.default_rng(77 + mode)
    base = rng.integers(0, 256, (1400, 32), dtype=np.uint8)
    imgs = []
    for i, n in enumerate((1025, 769, 577, 513, 512, 511, 385, 384, 257, 256, 225, 193, 192, 97, 33, 31, 1)):
        ids = rng.permutation(1400)[:n]
        d = base[ids].copy()
        flip = rng.random((n, 32)) < 0.05
        d ^= (flip * rng.integers(1, 256, (n, 32))).astype(np.uint8)
        if n > 300:
            d[40] = d[7]; d[41] = d[7]; d[300] = d[7]
            d[5] = 0; d[6] = 255
        imgs.append(d)
    ramp = np.zeros((257, 256), np.uint8)
    for k in range(257):
        ramp[k, rng.permutation(256)[:k]] = 1                  # row k has exactly k set bits
    imgs.insert(3, np.packbits(ramp, axis=1))
    outs = []
    for flags in (0, FORCE_FP4_UNPACKED, FORCE_I8, FORCE_E4M3, FORCE_POPC):
        with api.PairMatcher(unique_mode=mode, debug_flags=flags, do_filter=0, batch_pairs=7) as pm:
            for i, d in enumerate(imgs):
                pm.set_image(i, d)
            outs.append(pm.match_all_pairs())
    for o in outs[1:]:
        for k in ("offsets", "q", "t", "status"):
            assert np.array_equal(outs[0][k], o[k]), (mode, k)
    res = outs[0]
    for p, (i, j) in enumerate(res["pair_ij"]):
        if mode != api.UNIQUE_FIRST_WINS or p % 6:
            continue
        oi, od = orc.knn2_hamming(imgs[i], imgs[j])
        wq, wt = orc.ratio_unique(oi, od.astype(np.float32), imgs[j].shape[0])
        a, b = res["offsets"][p], res["offsets"][p + 1]
        assert np.array_equal(res["q"][a:b], wq) and np.array_equal(res["t"][a:b], wt), (i, j)
    assert res["offsets"][-1] > 1000


# ---------------------------------------------------------------------------------------------
# integer-valued 128-d rows as bytes on kind::i8 (default batched SIFT path) vs the fp16 form and the oracle
# ---------------------------------------------------------------------------------------------
FORCE_F16 = 1 << 11


def _csr_equal(a, b, keys=("offsets", "q", "t", "inlier", "status", "n_inliers", "ransac_iters", "F")):
    for k in keys:
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("mode", [api.UNIQUE_FIRST_WINS, api.MUTUAL_NN, api.UNIQUE_NONE])
def test_i8_form_parity_ties_force_rescans(mode):
    """Tiny value range: many train rows tie in D' = 127*sum(a) - a.b + floor(|b|^2/2) with different parities of
    |b|^2, and exact duplicates tie completely -- the fix-up must rescan those rows and still match the oracle."""
    rng = np.random.default_rng(101 + mode)
    imgs = []
    for i, n in enumerate((700, 643, 300, 17, 1)):
        d = rng.integers(0, 3, (n, 128)).astype(np.float32)
        if n > 400:
            d[1::7] = d[0]                                     # duplicates of one row all over the image
            d[5] = 0
        imgs.append(d)
    imgs[1][:200] = imgs[0][:200]                              # exact matches (d = 0) between images 0 and 1
    imgs[1][200:260] = np.minimum(imgs[0][200:260] + (rng.random((60, 128)) < 0.01), 255)   # d^2 = 0..3
    imgs[1][300] = imgs[0][10]; imgs[1][500] = imgs[0][10]     # d = 0 in three chunks: D' ties, rescan for sure
    outs = []
    for flags in (0, FORCE_F16):
        with api.PairMatcher(unique_mode=mode, batch_pairs=4, do_filter=0, debug_flags=flags) as pm:
            for i, d in enumerate(imgs):
                pm.set_image(i, d)
            outs.append(pm.match_all_pairs())
            st = pm.stats()
        if flags == 0:
            assert st["rerank_overflow"] > 0, st               # the rescan path ran
    _csr_equal(outs[0], outs[1], ("offsets", "q", "t", "status"))
    res = outs[0]
    for p, (i, j) in enumerate(res["pair_ij"]):
        oi, o2 = orc.knn2_l2(imgs[i], imgs[j])
        od = np.sqrt(o2).astype(np.float32)
        bq = orc.best_query(imgs[i], imgs[j]) if mode == api.MUTUAL_NN else None
        wq, wt = orc.ratio_unique(oi, od, imgs[j].shape[0], 0.7, mode, bq)
        a, b = res["offsets"][p], res["offsets"][p + 1]
        assert np.array_equal(res["q"][a:b], wq) and np.array_equal(res["t"][a:b], wt), (mode, i, j)
    assert res["offsets"][-1] > 100


def test_i8_form_norm_limit_and_fallback():
    """|b|^2 up to 4,032,059 rides in the 31-digit norm block; beyond it the image stays on the fp16 form."""
    rng = np.random.default_rng(9)
    def make(n, heavy):
        d = rng.integers(0, 40, (n, 128)).astype(np.float32)
        d[:heavy, :62] = 255                                   # 62 * 255^2 = 4,031,550 + small rest
        d[:heavy, 62:] = rng.integers(0, 2, (heavy, 66))
        return d
    a, b = make(500, 40), make(450, 30)
    b[100:140] = a[60:100]
    c = a.copy(); c[3, :] = 255                                # one all-255 row: |b|^2 = 8.3 M -> no byte form
    for imgs in ((a, b), (c, b)):
        outs = []
        for flags in (0, FORCE_F16, 1):
            with api.PairMatcher(do_filter=0, debug_flags=flags) as pm:
                for i, d in enumerate(imgs):
                    pm.set_image(i, d)
                outs.append(pm.match_all_pairs())
        _csr_equal(outs[0], outs[1], ("offsets", "q", "t", "status"))
        _csr_equal(outs[0], outs[2], ("offsets", "q", "t", "status"))
        assert outs[0]["offsets"][-1] >= 40


def test_i8_form_full_size_equals_f16_form_with_outliers_and_mutual():
    w = synth.World("sift", 8192, seed=0xB200 + 7)
    imgs = [w.image(i, 100, outlier_frac=0.5 if i == 1 else 0.0)[:2] for i in range(4)]
    for mode in (api.UNIQUE_FIRST_WINS, api.MUTUAL_NN):
        outs = []
        for flags in (0, FORCE_F16):
            with api.PairMatcher(unique_mode=mode, debug_flags=flags) as pm:
                for i, (d, xy) in enumerate(imgs):
                    pm.set_image(i, d, xy, dtype=api.DESC_U8 if flags == 0 else None)
                outs.append(pm.match_all_pairs())
        _csr_equal(outs[0], outs[1])
        assert outs[0]["offsets"][-1] > 6 * 800


@pytest.mark.parametrize("kind", ["sift", "orb", "superpoint", "sp128"])
def test_async_ingest_gives_identical_results(kind):
    """pm_set_image_async (no host synchronisation per image; facts resolved at first use) == pm_set_image."""
    import torch
    w = synth.World("superpoint" if kind == "sp128" else kind, 1200, seed=77)
    imgs = [w.image(i, 6)[:2] for i in range(6)]
    if kind == "sp128":                                        # real-valued 128-d rows: fp16 forms packed at first use
        imgs = [(np.ascontiguousarray(d[:, :128] / np.linalg.norm(d[:, :128], axis=1, keepdims=True)).astype(np.float32), xy)
                for d, xy in imgs]
    pinned = [(torch.from_numpy(np.ascontiguousarray(d)).pin_memory(),
               torch.from_numpy(np.ascontiguousarray(xy, dtype=np.int32)).pin_memory()) for d, xy in imgs]
    dim = imgs[0][0].shape[1] * (8 if kind == "orb" else 1)
    dt = api.DESC_U8_BITS if kind == "orb" else api.DESC_F32
    outs = []
    for asyn in (False, True):
        with api.PairMatcher(batch_pairs=4) as pm:
            for rep in range(2):                               # re-ingest over resident images too
                if asyn and rep == 1:                          # the whole set in one call
                    pm.set_images_ptr_async(list(range(len(pinned))), [td.data_ptr() for td, _ in pinned],
                                            [td.shape[0] for td, _ in pinned], dim, dt, [tx.data_ptr() for _, tx in pinned])
                else:
                    for i, (td, tx) in enumerate(pinned):
                        pm.set_image_ptr(i, td.data_ptr(), td.shape[0], dim, dt, tx.data_ptr(), asynchronous=asyn)
                if rep == 0 and asyn:
                    pm.sync_images()
            outs.append(pm.match_all_pairs())
            i1, d1 = pm.knn_pair(0, 1)
            outs[-1]["knn"] = (i1, d1)
    _csr_equal(outs[0], outs[1])
    assert np.array_equal(outs[0]["knn"][0], outs[1]["knn"][0]) and np.array_equal(outs[0]["knn"][1], outs[1]["knn"][1])
    assert outs[0]["offsets"][-1] > (15 * 200 if kind != "sp128" else 15 * 50)


@pytest.mark.parametrize("kind", ["sift", "orb"])
def test_batch_async_ingest_shuffled_pairs_and_reingest(kind):
    """pm_set_images_async for a set of 14 images of different sizes: a shuffled explicit pair list that leaves two images
    untouched (their uploads complete behind the call), then the untouched images through other entry points, then the
    same ids again with other sizes (rows re-used and re-allocated) and the implicit all-pairs list: everything equals the
    synchronous path.  (Also the parity test of tools/r02/deferred_ingest_wip.patch, a lazily queued variant that was
    measured to be worth < 1 % end to end -- the first batches wait for the PCIe upload, not for the host -- and not applied.)"""
    import torch
    w = synth.World(kind, 900, seed=91)
    rng = np.random.default_rng(5)
    sizes = [900, 650, 333, 777, 64, 900, 128, 500, 1, 420, 810, 256, 7, 600]
    def make(sz):
        out = []
        for i, n in enumerate(sz):
            d, xy = w.image(i, len(sz))[:2]
            out.append((torch.from_numpy(np.ascontiguousarray(d[:n])).pin_memory(),
                        torch.from_numpy(np.ascontiguousarray(xy[:n], dtype=np.int32)).pin_memory()))
        return out
    sets = [make(sizes), make(sizes[::-1])]
    dim = sets[0][0][0].shape[1] * (8 if kind == "orb" else 1)
    dt = api.DESC_U8_BITS if kind == "orb" else api.DESC_F32
    pairs = np.array([(i, j) for j in range(12) for i in range(j)], dtype=np.int32)      # images 12, 13 stay untouched
    pairs = pairs[rng.permutation(len(pairs))]
    outs = []
    for deferred in (False, True):
        res = []
        with api.PairMatcher(batch_pairs=5) as pm:
            for rnd, pinned in enumerate(sets):
                ids = list(range(len(pinned)))
                if deferred:
                    pm.set_images_ptr_async(ids, [td.data_ptr() for td, _ in pinned], [td.shape[0] for td, _ in pinned], dim, dt,
                                            [tx.data_ptr() for _, tx in pinned])
                else:
                    for i, (td, tx) in enumerate(pinned):
                        pm.set_image_ptr(i, td.data_ptr(), td.shape[0], dim, dt, tx.data_ptr(), asynchronous=False)
                if rnd == 0:
                    res.append(pm.match_all_pairs(pairs))
                    assert pm.lib.pm_num_keypoints(pm.h, 13) == pinned[13][0].shape[0]
                    res.append({"knn": pm.knn_pair(13, 11)})
                else:
                    res.append(pm.match_all_pairs())               # implicit list: includes the deferred ids
        outs.append(res)
    _csr_equal(outs[0][0], outs[1][0])
    _csr_equal(outs[0][2], outs[1][2])
    assert np.array_equal(outs[0][1]["knn"][0], outs[1][1]["knn"][0]) and np.array_equal(outs[0][1]["knn"][1], outs[1][1]["knn"][1])
    assert outs[0][0]["offsets"][-1] > 2000 and outs[0][2]["n_pairs"] == 91


# ---------------------------------------------------------------------------------------------
# real-valued rows quantised to s8 on kind::i8 (default batched SuperPoint path without cross-check) vs fp16 forms
# ---------------------------------------------------------------------------------------------
FORCE_FP16_FORMS = 1 << 15
S8_ONE_ROW_SET = 1 << 19          # the s8 candidate kernel with one query row set per cluster (default: two)
S8_SIX_KEYS = 1 << 20             # two row sets, six chunk keys + chunk re-rank (default: argmin epilogue + one-column re-rank)
S8_WITH_NORM = 1 << 23      # unit-norm rows: keep the norm K-step of the s8 search (default: dropped)


@pytest.mark.parametrize("dim", [128, 256])
@pytest.mark.parametrize("mode", [api.UNIQUE_FIRST_WINS, api.UNIQUE_NONE, api.MUTUAL_NN])
def test_s8_quantised_candidates_equal_fp16_forms_and_simt(dim, mode):
    """The candidate stage only proposes; the certified fp32 re-rank decides.  Outputs must be identical whichever
    operand precision proposed (s8 / fp16) and identical to the fp32 SIMT search."""
    rng = np.random.default_rng(200 + dim + mode)
    base = _unit_rows(rng, 3000, dim)
    imgs = []
    for i, n in enumerate((1100, 1000, 700, 300, 40)):
        ids = rng.permutation(3000)[:n]
        d = base[ids] + 0.35 * rng.standard_normal((n, dim)).astype(np.float32) / np.sqrt(dim)
        d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        if i == 1:
            d[11] = d[5]; d[700] = d[5]                     # duplicates: ties in the candidate scores
        imgs.append(d)
    outs = []
    for flags in (0, FORCE_FP16_FORMS, 1, S8_ONE_ROW_SET, S8_SIX_KEYS, S8_WITH_NORM):
        with api.PairMatcher(unique_mode=mode, batch_pairs=4, do_filter=0, debug_flags=flags) as pm:
            for i, d in enumerate(imgs):
                pm.set_image(i, d)
            outs.append(pm.match_all_pairs())
            st = pm.stats()
        if flags in (0, S8_ONE_ROW_SET, S8_SIX_KEYS, S8_WITH_NORM) and mode != api.MUTUAL_NN:
            assert st["rerank_rows"] > 0 and st["rerank_worst_err"] < 1.0, st
    for o in outs[1:]:
        _csr_equal(outs[0], o, ("offsets", "q", "t", "status"))
    assert outs[0]["offsets"][-1] > 300


def test_s8_quantised_fallback_on_large_entries_and_full_size():
    """Rows with an entry above 0.5 cannot be quantised with scale 254: the image stays on the fp16 forms.  Full-size
    SuperPoint-like images: s8 candidates == fp16 candidates, CSR identical including the RANSAC outputs."""
    rng = np.random.default_rng(5)
    a = _unit_rows(rng, 600, 256); b = _unit_rows(rng, 500, 256)
    b[100:200] = a[50:150]
    a2 = a.copy(); a2[7] = 0; a2[7, 3] = 0.8; a2[7, 4] = 0.6      # unit norm, max |x| = 0.8
    for imgs in ((a, b), (a2, b)):
        outs = []
        for flags in (0, FORCE_FP16_FORMS):
            with api.PairMatcher(do_filter=0, debug_flags=flags) as pm:
                for i, d in enumerate(imgs):
                    pm.set_image(i, d)
                outs.append(pm.match_all_pairs())
        _csr_equal(outs[0], outs[1], ("offsets", "q", "t", "status"))
        assert outs[0]["offsets"][-1] >= 99
    w = synth.World("superpoint", 8192, seed=0xB200 + 9)
    imgs = [w.image(i, 100, outlier_frac=0.3 if i == 1 else 0.0)[:2] for i in range(4)]
    outs = []
    for flags in (0, FORCE_FP16_FORMS, S8_ONE_ROW_SET, S8_SIX_KEYS, S8_WITH_NORM):
        with api.PairMatcher(debug_flags=flags) as pm:
            for i, (d, xy) in enumerate(imgs):
                pm.set_image(i, d, xy)
            outs.append(pm.match_all_pairs())
            st = pm.stats()
        assert st["rerank_worst_err"] < 1.0, st
    for o in outs[1:]:
        _csr_equal(outs[0], o)
    assert outs[0]["offsets"][-1] > 6 * 1000


def test_s8_unit_norm_form_only_for_unit_norm_train_images():
    """The s8 search drops its norm K-step when every TRAIN row of the batch has |b|^2 within 2^-10 of 1 (SuperPoint's
    L2-normalised rows); images that are not normalised (here: rows scaled to norms 0.8 .. 1.0) keep it.  Whatever
    the mix, the output equals the fp16 forms and the fp32 SIMT search."""
    rng = np.random.default_rng(77)
    base = _unit_rows(rng, 2500, 256)
    def img(n, scaled):
        ids = rng.permutation(2500)[:n]
        d = base[ids] + 0.3 * rng.standard_normal((n, 256)).astype(np.float32) / 16
        d = d / np.linalg.norm(d, axis=1, keepdims=True)
        if scaled:
            d = d * rng.uniform(0.8, 1.0, (n, 1))
        return d.astype(np.float32)
    imgs = [img(900, False), img(800, True), img(700, False), img(600, True), img(520, False)]
    outs = []
    for flags in (0, S8_WITH_NORM, FORCE_FP16_FORMS, 1):
        for bp in ((0, 1) if flags == 0 else (0,)):              # batch of one pair: the form is chosen per batch
            with api.PairMatcher(do_filter=0, debug_flags=flags, batch_pairs=bp) as pm:
                for i, d in enumerate(imgs):
                    pm.set_image(i, d)
                outs.append(pm.match_all_pairs())
                assert pm.stats()["rerank_worst_err"] < 1.0
    for o in outs[1:]:
        _csr_equal(outs[0], o, ("offsets", "q", "t", "status"))
    assert outs[0]["offsets"][-1] > 1000


@pytest.mark.parametrize("kind", ["sift", "superpoint", "orb"])
def test_batched_paths_edge_sizes_equal_reference_kernels(kind):
    """Image sizes around every tiling boundary of the batched tensor paths (empty image, 1 row, 32-column chunks,
    256-row tiles, the 512-row work item of the two-set kernel): default path == SIMT / popc path, array for array."""
    rng = np.random.default_rng(31)
    sizes = (1, 2, 31, 32, 33, 255, 256, 257, 511, 512, 513, 769, 0)
    if kind == "sift":
        base = rng.integers(0, 60, (1200, 128)).astype(np.float32)
        noise = lambda n: rng.integers(-3, 4, (n, 128))
        make = lambda ids: np.clip(base[ids] + noise(len(ids)), 0, 255).astype(np.float32)
        ref_flags = 1
    elif kind == "superpoint":
        base = _unit_rows(rng, 1200, 256)
        def make(ids):
            d = base[ids] + 0.3 * rng.standard_normal((len(ids), 256)).astype(np.float32) / 16
            return (0.999 * d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        ref_flags = 1
    else:
        base = rng.integers(0, 256, (1200, 32), dtype=np.uint8)
        def make(ids):
            d = base[ids].copy()
            d ^= ((rng.random(d.shape) < 0.05) * rng.integers(1, 256, d.shape)).astype(np.uint8)
            return d
        ref_flags = FORCE_POPC
    imgs = [make(rng.permutation(1200)[:n]) if n else np.zeros((0, base.shape[1]), base.dtype) for n in sizes]
    outs = []
    for flags in (0, ref_flags):
        with api.PairMatcher(do_filter=0, batch_pairs=16, debug_flags=flags) as pm:
            for i, d in enumerate(imgs):
                pm.set_image(i, d)
            outs.append(pm.match_all_pairs())
    _csr_equal(outs[0], outs[1], ("pair_ij", "offsets", "q", "t", "status"))
    assert outs[0]["offsets"][-1] > 2000


@pytest.mark.parametrize("seed", list(range(6)))
@pytest.mark.parametrize("kind", ["sift", "superpoint"])
def test_fuzz_batched_paths_vs_reference_kernels(kind, seed):
    """Random image sizes, value ranges, duplicated and shared rows, all uniqueness modes: the default batched path
    (byte / s8 forms on kind::i8 + fix-up / re-rank) must equal the fp32 SIMT path array for array."""
    rng = np.random.default_rng(1000 * seed + (7 if kind == "sift" else 13))
    n_img = int(rng.integers(3, 7))
    pool = 1500
    if kind == "sift":
        hi = int(rng.choice([2, 4, 30, 120, 256]))
        base = rng.integers(0, hi, (pool, 128)).astype(np.float32)
    else:
        base = _unit_rows(rng, pool, int(rng.choice([128, 256])))
    imgs = []
    for i in range(n_img):
        n = int(rng.choice([1, 5, 31, 33, 64, 200, 257, 400, 700, 1030]))
        ids = rng.integers(0, pool, n)                           # with repetition: duplicate rows inside an image
        d = base[ids].copy()
        if kind == "sift":
            amp = int(rng.choice([0, 1, 3]))
            if amp:
                d = np.clip(d + rng.integers(-amp, amp + 1, d.shape), 0, 255).astype(np.float32)
        else:
            d = d + float(rng.choice([0.0, 0.05, 0.3])) * rng.standard_normal(d.shape).astype(np.float32) / np.sqrt(d.shape[1])
            d = (0.999 * d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        imgs.append(d)
    mode = int(rng.choice([api.UNIQUE_FIRST_WINS, api.MUTUAL_NN, api.UNIQUE_NONE]))
    ratio = float(rng.choice([0.7, 0.95, 0.5]))
    outs = []
    for flags in (0, 1):
        with api.PairMatcher(unique_mode=mode, ratio=ratio, do_filter=0, batch_pairs=int(rng.choice([2, 5, 64])) if flags == 0 else 64,
                             debug_flags=flags) as pm:
            for i, d in enumerate(imgs):
                pm.set_image(i, d)
            outs.append(pm.match_all_pairs())
    _csr_equal(outs[0], outs[1], ("pair_ij", "offsets", "q", "t", "status"))

# ---------------------------------------------------------------------------------------------
# on-disk cache: resume from files, identical results
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["orb", "sift", "sift-u8", "superpoint"])
def test_cache_resume_gives_identical_results(kind, tmp_path):
    from reconstructor_b200 import cache
    base = kind.split("-")[0]
    dt = api.DESC_U8 if kind == "sift-u8" else None
    imgs = synth.make_set(base, 4, 500, seed=17)
    if kind == "sift-u8":
        imgs = [(d.astype(np.uint8), xy) for d, xy in imgs]
    img_file, res_file = str(tmp_path / "images.pmb"), str(tmp_path / "result.pmb")
    with api.PairMatcher() as pm:
        for i, (d, xy) in enumerate(imgs):
            pm.set_image(10 * i, d, xy if i != 2 else None, dtype=dt)       # sparse ids, one image without xy
        first = pm.match_all_pairs(save_to=res_file)
        pm.save_images(img_file)
    # the file written by the library is readable by the numpy mirror and holds exactly what was ingested
    recs = cache.read_images(img_file)
    assert [r["id"] for r in recs] == [0, 10, 20, 30] and recs[2]["xy"] is None
    for r, (d, xy) in zip(recs, imgs):
        assert np.array_equal(r["desc"], d) and (r["xy"] is None or np.array_equal(r["xy"], xy))
    # resume: fresh handle, images from the cache file
    with api.PairMatcher() as pm:
        pm.load_images(img_file)
        again = pm.match_all_pairs()
    saved = api.load_result(res_file)
    for k in ("pair_ij", "offsets", "q", "t", "inlier", "F", "status", "n_inliers", "ransac_iters"):
        assert np.array_equal(first[k], again[k]), (kind, k)
        assert np.array_equal(first[k], saved[k]), (kind, k)
    assert first["offsets"][-1] > 100
    # a file produced WITHOUT the library (sharded extraction on another machine) ingests the same way
    alt = str(tmp_path / "alt.pmb")
    cache.write_images(alt, {10 * i: (d, xy if i != 2 else None) for i, (d, xy) in enumerate(imgs)}, dtype=dt)
    with api.PairMatcher() as pm:
        pm.load_images(alt)
        third = pm.match_all_pairs()
    assert np.array_equal(first["q"], third["q"]) and np.array_equal(first["t"], third["t"])


# ---------------------------------------------------------------------------------------------
# round 2: Philox sampler + seed, 8-point refit, boundary behaviour (SURVEY 8b / 8c leg 5)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("resid", [api.RESID_SYMMETRIC_EPIPOLAR, api.RESID_SAMPSON])
def test_philox_sampler_masks_identical_to_cpu_filter(scenes, resid):
    """'with any other sampler: GPU <-> CPU-filter identity'.  Same Philox key -> same hypothesis stream ->
    identical masks, iteration counts and F."""
    n_sc = int(scenes["n_scenes"])
    checked = 0
    with api.PairMatcher(sampler=api.SAMPLER_PHILOX, residual_mode=resid, seed=5) as pm:
        for k in range(n_sc):
            p1, p2 = scenes[f"s{k}_p1"], scenes[f"s{k}_p2"]
            for key in (0, 0xB200, 2**63 + 12345 + k):
                F, mask, st, it = pm.estimate_fundamental(p1, p2, pair_key=key)
                prm = orc.default_params(residual_mode=resid, sampler=orc.SAMPLER_PHILOX, seed=key)
                ns, Fo, mo, tr = orc.find_fundamental(p1, p2, prm)
                assert (st == api.PAIR_FILTERED) == (ns > 0), (k, key)
                if ns > 0:
                    assert np.array_equal(mask, mo), (k, key, int(mask.sum()), int(mo.sum()))
                    if p1.shape[0] > 7:
                        assert it == tr.iters_run, (k, key, it, tr.iters_run)
                        np.testing.assert_allclose(F, Fo[0], rtol=0, atol=1e-9 * max(1.0, np.abs(Fo[0]).max()))
                    checked += 1
        # pm_filter_pair_F without an explicit key uses pm_params.seed
        p1, p2 = scenes["s30_p1"], scenes["s30_p2"]
        a = pm.estimate_fundamental(p1, p2)
        b = pm.estimate_fundamental(p1, p2, pair_key=5)
        assert np.array_equal(a[1], b[1]) and a[3] == b[3]
    assert checked >= 150


def test_philox_batched_loop_uses_pair_keys_and_is_partition_invariant():
    """Per-pair key = f(seed, i, j): the batched result equals the CPU pair body run with orc.pair_seed, whatever the
    batch size or the order / subset of the pair list (a pair's result does not depend on who ran it)."""
    w = synth.World("sift", 700, seed=91)
    imgs = [w.image(i, 6, outlier_frac=0.4)[:2] for i in range(5)]
    assert api.pair_seed(77, 1, 3) == orc.pair_seed(77, 1, 3)
    outs = []
    for bp, order in ((0, None), (3, None), (2, "rev")):
        with api.PairMatcher(sampler=api.SAMPLER_PHILOX, seed=77, batch_pairs=bp) as pm:
            for i, (d, xy) in enumerate(imgs):
                pm.set_image(i, d, xy)
            pairs = np.array([(i, j) for j in range(5) for i in range(j)], np.int32)
            if order == "rev":
                pairs = pairs[::-1].copy()
            outs.append(pm.match_all_pairs(pairs))
    res = outs[0]
    for p, (i, j) in enumerate(res["pair_ij"]):
        prm = orc.default_params(sampler=orc.SAMPLER_PHILOX, seed=orc.pair_seed(77, int(i), int(j)))
        ref = orc.match_pair(imgs[i][0], imgs[i][1], imgs[j][0], imgs[j][1], params=prm)
        a, b = res["offsets"][p], res["offsets"][p + 1]
        keep = res["inlier"][a:b].astype(bool)
        assert ref["status"] == "ok" and res["status"][p] == api.PAIR_FILTERED
        assert np.array_equal(res["q"][a:b][keep], ref["q"]) and np.array_equal(res["t"][a:b][keep], ref["t"]), (i, j)
    _csr_equal(outs[0], outs[1])
    rev = outs[2]
    for p, (i, j) in enumerate(res["pair_ij"]):
        r = len(res["pair_ij"]) - 1 - p
        assert tuple(rev["pair_ij"][r]) == (i, j)
        assert np.array_equal(rev["inlier"][rev["offsets"][r]:rev["offsets"][r + 1]],
                              res["inlier"][res["offsets"][p]:res["offsets"][p + 1]])
        assert rev["ransac_iters"][r] == res["ransac_iters"][p]


@pytest.mark.parametrize("sampler", [api.SAMPLER_OPENCV_MWC, api.SAMPLER_PHILOX])
@pytest.mark.parametrize("resid", [api.RESID_SYMMETRIC_EPIPOLAR, api.RESID_SAMPSON])
def test_staged_filter_equals_one_kernel_filter_and_cpu_filter(scenes, sampler, resid):
    """Pairs that need more than the first 24 iterations continue as device-wide stages (sample / solve / score / select
    kernels over mega-rounds of 256 iterations, ransac.cu); debug_flags bit 21 keeps every iteration in the per-pair
    kernel.  Same subsets, models, counts and selection order: masks, F and iteration counts are identical, and equal to
    the CPU filter, for every iteration cap around the mega-round boundaries."""
    ks = [k for k in range(int(scenes["n_scenes"])) if scenes[f"s{k}_meta"][1] >= 0.4 and scenes[f"s{k}_p1"].shape[0] >= 33]
    assert len(ks) >= 10
    osamp = orc.SAMPLER_PHILOX if sampler == api.SAMPLER_PHILOX else orc.SAMPLER_OPENCV_MWC
    handed = 0
    for cap in (24, 25, 56, 100, 279, 280, 281, 1000, 2000):
        with api.PairMatcher(sampler=sampler, residual_mode=resid, seed=9, ransac_max_iters=cap) as pm, \
             api.PairMatcher(sampler=sampler, residual_mode=resid, seed=9, ransac_max_iters=cap, debug_flags=1 << 21) as pm1:
            for k in ks if cap in (280, 1000) else ks[::3]:
                p1, p2 = scenes[f"s{k}_p1"], scenes[f"s{k}_p2"]
                F, mask, st, it = pm.estimate_fundamental(p1, p2)
                F1, mask1, st1, it1 = pm1.estimate_fundamental(p1, p2)
                assert st == st1 and it == it1 and np.array_equal(mask, mask1) and np.array_equal(F, F1), (cap, k, it, it1)
                prm = orc.default_params(residual_mode=resid, sampler=osamp, seed=9, max_iters=cap)
                ns, Fo, mo, tr = orc.find_fundamental(p1, p2, prm)
                assert (st == api.PAIR_FILTERED) == (ns > 0) and it == tr.iters_run, (cap, k, it, tr.iters_run)
                if ns > 0:
                    assert np.array_equal(mask, mo), (cap, k)
                handed += it > 24
    assert handed >= 20


def test_staged_filter_batched_loop_with_outliers_equals_one_kernel_filter():
    """The batched loop with half of the keypoints displaced (every pair runs to the iteration cap): CSR arrays, F and
    iteration counts of the staged filter equal those of the per-pair kernel."""
    w = synth.World("sift", 1500, seed=17)
    imgs = [w.image(i, 8, outlier_frac=0.5)[:2] for i in range(6)]
    outs = []
    for flags in (1 << 22, 1 << 21, 0):       # always staged / never staged / by the scheduling hint
        with api.PairMatcher(debug_flags=flags, batch_pairs=4) as pm:
            for i, (d, xy) in enumerate(imgs):
                pm.set_image(i, d, xy)
            outs.append(pm.match_all_pairs())
    _csr_equal(outs[0], outs[1])
    _csr_equal(outs[0], outs[2])
    assert np.array_equal(outs[0]["ransac_iters"], outs[1]["ransac_iters"]) and np.array_equal(outs[0]["F"], outs[1]["F"])
    assert (outs[0]["ransac_iters"] > 24).sum() >= 10


def _random_two_view(rng, n, of, subpixel, scale, shift):
    def rot(rx, ry, rz):
        cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
        Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]); Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        return np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ Ry @ Rx
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


def test_fuzz_epipolar_filter_staged_per_pair_and_cpu_filter():
    """Random two-view scenes (8..3000 matches, 0..80 % outliers, integer / sub-pixel / shifted / scaled coordinates, both
    samplers, both residual modes, several iteration caps): the staged filter, the per-pair kernel and the CPU filter agree
    on status, mask, iteration count (and the two GPU forms on F bit for bit).  tools/r02/fuzz_ransac.py ran 300 more."""
    n_long = 0
    for seed in range(40):
        rng = np.random.default_rng(777 + seed)
        sampler = int(rng.integers(0, 2)); resid = int(rng.integers(0, 2))
        cap = int(rng.choice([1000, 1000, 1000, 300, 2000, 25, 281]))
        n = int(rng.choice([8, 9, 15, 40, 100, 333, 1000, 2050, 3000]))
        of = float(rng.choice([0.0, 0.2, 0.4, 0.5, 0.6, 0.8]))
        scale, shift = [(1.0, 0.0), (1.0, 0.0), (0.37, 0.0), (1.0, 50000.0), (31.0, 0.0)][int(rng.integers(0, 5))]
        p1, p2 = _random_two_view(rng, n, of, bool(rng.integers(0, 2)), scale, shift)
        osamp = orc.SAMPLER_PHILOX if sampler else orc.SAMPLER_OPENCV_MWC
        prm = orc.default_params(residual_mode=resid, sampler=osamp, seed=seed, max_iters=cap)
        ns, Fo, mo, tr = orc.find_fundamental(p1, p2, prm)
        res = []
        for flags in (0, 1 << 21):
            with api.PairMatcher(sampler=sampler, residual_mode=resid, seed=seed, ransac_max_iters=cap, debug_flags=flags) as pm:
                res.append(pm.estimate_fundamental(p1, p2))
        (F, mask, st, it), (F1, mask1, st1, it1) = res
        assert st == st1 and it == it1 and np.array_equal(mask, mask1) and np.array_equal(F, F1), (seed, it, it1)
        assert (st == api.PAIR_FILTERED) == (ns > 0) and it == tr.iters_run, (seed, it, tr.iters_run)
        if ns > 0:
            assert np.array_equal(mask, mo), seed
        n_long += it > 24
    assert n_long >= 15


def test_eight_point_refit_matches_cpu_filter_and_cv2(scenes, golden_dir):
    g8 = np.load(os.path.join(golden_dir, "eight_point.npz"))
    prm = orc.default_params(refit_8point=1)
    with api.PairMatcher(refit_8point=1) as pm, api.PairMatcher() as pm0:
        for k in g8["scenes"]:
            p1, p2 = scenes[f"s{k}_p1"], scenes[f"s{k}_p2"]
            F, mask, st, it = pm.estimate_fundamental(p1, p2)
            F0, mask0, st0, it0 = pm0.estimate_fundamental(p1, p2)
            ns, Fo, mo, tr = orc.find_fundamental(p1, p2, prm)
            assert st == api.PAIR_FILTERED and ns == 1
            assert np.array_equal(mask, mask0) and it == it0          # the refit changes F only
            assert np.array_equal(mask, mo)
            assert np.abs(F - Fo[0]).max() <= 1e-9 * np.abs(Fo[0]).max(), int(k)
            # cv2's own FM_8POINT over the same inliers (the GPU mask equals cv2's for these scenes)
            assert np.array_equal(mask, scenes[f"s{k}_mask"])
            Fg = g8[f"s{k}_F8"]
            assert np.abs(F - Fg).max() <= 1e-7 * np.abs(Fg).max(), int(k)
    # batched loop: refit on every filtered pair, masks untouched
    w = synth.World("orb", 600, seed=3)
    imgs = [w.image(i, 4, outlier_frac=0.3)[:2] for i in range(4)]
    outs = []
    for refit in (0, 1):
        with api.PairMatcher(refit_8point=refit) as pm:
            for i, (d, xy) in enumerate(imgs):
                pm.set_image(i, d, xy)
            outs.append(pm.match_all_pairs())
    _csr_equal(outs[0], outs[1], ("offsets", "q", "t", "inlier", "status", "n_inliers", "ransac_iters"))
    assert not np.array_equal(outs[0]["F"], outs[1]["F"])
    for p in range(outs[1]["n_pairs"]):
        a, b = outs[1]["offsets"][p], outs[1]["offsets"][p + 1]
        i, j = outs[1]["pair_ij"][p]
        q, t, m = outs[1]["q"][a:b], outs[1]["t"][a:b], outs[1]["inlier"][a:b]
        ok, Fo = orc.eight_point(imgs[i][1][q].astype(np.float32), imgs[j][1][t].astype(np.float32), m)
        assert ok and np.abs(outs[1]["F"][p] - Fo).max() <= 1e-9 * np.abs(Fo).max()


def test_failed_batched_call_leaves_no_stale_batches():
    """an unknown image id deep in the pair list must fail the call without leaving slots busy; the next,
    shorter call must return exactly its own pairs."""
    w = synth.World("sift", 500, seed=12)
    imgs = [w.image(i, 14)[:2] for i in range(14)]
    with api.PairMatcher(batch_pairs=8) as pm:
        for i, (d, xy) in enumerate(imgs):
            pm.set_image(i, d, xy)
        good = np.array([(i, j) for j in range(14) for i in range(j)], np.int32)
        ref = pm.match_all_pairs(good)
        bad = good.copy()
        bad[70] = (3, 999)                                   # unknown id at pair index >= 64
        with pytest.raises(api.PairMatchError) as e:
            pm.match_all_pairs(bad)
        assert e.value.code == api.ERR_STATE and "999" in str(e.value)
        short = pm.match_all_pairs(good[:5])
        assert short["n_pairs"] == 5
        for k in ("q", "t", "inlier"):
            assert np.array_equal(short[k], ref[k][:ref["offsets"][5]]), k
        assert np.array_equal(short["offsets"], ref["offsets"][:6])
        again = pm.match_all_pairs(good)
        _csr_equal(again, ref)


def test_all_pairs_sentinel_empty_list_and_remove_image():
    w = synth.World("orb", 400, seed=8)
    imgs = [w.image(i, 4)[:2] for i in range(4)]
    with api.PairMatcher() as pm:
        for i, (d, xy) in enumerate(imgs):
            pm.set_image(i, d, xy)
        allp = pm.match_all_pairs()                          # pairs = NULL, n_pairs = PM_ALL_PAIRS
        assert allp["n_pairs"] == 6
        empty = pm.match_all_pairs(np.zeros((0, 2), np.int32))
        assert empty["n_pairs"] == 0 and len(empty["q"]) == 0
        res = C.POINTER(api.CsrResult)()
        assert pm.lib.pm_match_all_pairs(pm.h, None, 3, C.byref(res)) == api.ERR_INVALID
        assert pm.lib.pm_match_all_pairs(pm.h, None, 0, C.byref(res)) == api.OK and res.contents.n_pairs == 0
        pm.lib.pm_free_result(res)
        # removing an image frees its rows; a later image re-uses them (the arena does not grow)
        n_before = pm.stats()["n_images"]
        pm.remove_image(1)
        assert pm.stats()["n_images"] == n_before - 1
        with pytest.raises(api.PairMatchError):
            pm.match_pair(0, 1)
        with pytest.raises(api.PairMatchError):
            pm.remove_image(1)
        pm.set_image(7, imgs[1][0], imgs[1][1])
        again = pm.match_all_pairs(np.array([(0, 7), (7, 2), (7, 3)], np.int32))
        ref = {tuple(int(x) for x in ij): r for r, ij in enumerate(allp["pair_ij"])}
        for p, key in enumerate(((0, 1), (1, 2), (1, 3))):
            r = ref[key]
            assert np.array_equal(again["q"][again["offsets"][p]:again["offsets"][p + 1]],
                                  allp["q"][allp["offsets"][r]:allp["offsets"][r + 1]])
        # per-pair calls with growing temporary images re-use released rows too
        for n in (50, 120, 300, 90):
            r = pm.match_descriptors(imgs[0][0][:n], imgs[2][0][:n])
            assert len(r["q"]) > 0


def test_tensor_peaks_are_measured_and_plausible():
    with api.PairMatcher() as pm:
        f16 = pm.measure_tensor_peak(api.PEAK_KIND_F16)
        i8 = pm.measure_tensor_peak(api.PEAK_KIND_I8)
        fp4 = pm.measure_tensor_peak(api.PEAK_KIND_MXF4)
    print("tensor peaks TFLOP/s: f16 %.0f  i8 %.0f  mxf4 %.0f" % (f16 / 1e12, i8 / 1e12, fp4 / 1e12))
    assert 0.8e15 < f16 < 2.6e15 and 1.6e15 < i8 < 5.2e15 and 3.0e15 < fp4 < 10.5e15
    assert 1.6 < i8 / f16 < 2.4 and 1.5 < fp4 / i8 < 2.4


# ---------------------------------------------------------------------------------------------
# round 2: GeometricFilter::estimateEssential
# ---------------------------------------------------------------------------------------------
def test_essential_masks_identical_to_cv2_and_cpu_filter(golden_dir):
    g = np.load(os.path.join(golden_dir, "essential.npz"))
    n = int(g["n_scenes"])
    checked = 0
    with api.PairMatcher() as pm:
        for k in range(n):
            p1, p2, c1, c2 = g[f"e{k}_p1"], g[f"e{k}_p2"], g[f"e{k}_c1"], g[f"e{k}_c2"]
            E, mask, st, it = pm.estimate_essential(p1, p2, api.Camera(*c1), api.Camera(*c2))
            ok, Eo, mo, tr = orc.find_essential(p1, p2, orc.Camera(*c1), orc.Camera(*c2))
            # GPU <-> repo CPU filter: identical masks, iteration counts, E
            assert (st == api.PAIR_FILTERED) == ok, k
            assert np.array_equal(mask, mo), (k, int(mask.sum()), int(mo.sum()))
            if p1.shape[0] > 5:
                assert it == tr.iters_run, (k, it, tr.iters_run)
            assert np.abs(E - Eo).max() < 1e-9, (k, np.abs(E - Eo).max())
            # GPU <-> cv2.findEssentialMat: identical masks; E up to sign (N = 5: cv2 keeps the first model of ITS order)
            assert np.array_equal(mask, g[f"e{k}_mask"]), k
            Eg = g[f"e{k}_E"]
            if p1.shape[0] > 5:
                assert min(np.abs(E - Eg).max(), np.abs(E + Eg).max()) < 1e-4, k
            checked += 1
        # too few points / degenerate input
        cam = api.Camera(1200, 1200, 1024, 768, 0, 0)
        E, mask, st, it = pm.estimate_essential(np.zeros((4, 2), np.float32), np.ones((4, 2), np.float32), cam, cam)
        assert st == api.PAIR_DROPPED and not E.any()
        E, mask, st, it = pm.estimate_essential(np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32), cam, cam)
        assert st == api.PAIR_DROPPED
        with pytest.raises(api.PairMatchError):
            pm.estimate_essential(g["e10_p1"], g["e10_p2"], api.Camera(0, 1200, 1, 1, 0, 0), cam)
    assert checked >= 40
    # Philox sampler: GPU <-> CPU filter identity
    with api.PairMatcher(sampler=api.SAMPLER_PHILOX, seed=4242) as pm:
        for k in (10, 20, 30, 40, 45):
            p1, p2, c1, c2 = g[f"e{k}_p1"], g[f"e{k}_p2"], g[f"e{k}_c1"], g[f"e{k}_c2"]
            E, mask, st, it = pm.estimate_essential(p1, p2, api.Camera(*c1), api.Camera(*c2))
            ok, Eo, mo, tr = orc.find_essential(p1, p2, orc.Camera(*c1), orc.Camera(*c2), sampler=orc.SAMPLER_PHILOX, seed=4242)
            assert (st == api.PAIR_FILTERED) == ok and np.array_equal(mask, mo) and it == tr.iters_run, k
            assert np.abs(E - Eo).max() < 1e-9


# ---------------------------------------------------------------------------------------------
# round 2: pair pre-selection by retrieval (the ImageMatcher plugin)
# ---------------------------------------------------------------------------------------------
def _windowed_images(kind, n_img, n_kp, rng):
    """Image i observes landmarks of a window around i on a line of landmarks: neighbours overlap, far images do not."""
    L = 4 * n_kp
    if kind == "orb":
        base = rng.integers(0, 256, (L, 32), dtype=np.uint8)
    elif kind == "sift":
        base = rng.integers(0, 100, (L, 128)).astype(np.float32)
    else:
        base = _unit_rows(rng, L, 256)
    imgs = []
    for i in range(n_img):
        lo = int(i * (L - 2 * n_kp) / max(n_img - 1, 1))
        ids = lo + rng.permutation(2 * n_kp)[:n_kp]
        d = base[ids].copy()
        if kind == "orb":
            d ^= ((rng.random(d.shape) < 0.03) * rng.integers(1, 256, d.shape)).astype(np.uint8)
        elif kind == "sift":
            d = np.clip(d + rng.integers(-2, 3, d.shape), 0, 255).astype(np.float32)
        else:
            d = d + 0.2 * rng.standard_normal(d.shape).astype(np.float32) / 16
            d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        imgs.append(d)
    return imgs


@pytest.mark.parametrize("kind", ["sift", "orb", "superpoint"])
def test_pair_preselection_matches_cpu_restatement(kind):
    from oracle import retrieval_ref
    rng = np.random.default_rng(77)
    imgs = _windowed_images(kind, 12, 300, rng)
    with api.PairMatcher() as pm:
        for i, d in enumerate(imgs):
            pm.set_image(i, d)
        for k in (1, 3, 5):
            pairs, S = pm.select_pairs(k, want_scores=True)
            ref = retrieval_ref.select_pairs(imgs, k)
            assert np.array_equal(pairs, ref), (kind, k)
            np.testing.assert_allclose(S, retrieval_ref.similarity(imgs), rtol=0, atol=1e-12)
            assert len(pairs) >= 11 * k // 2
        # neighbours on the line are what retrieval finds: (almost) every (i, i+1) is selected at k = 3, all at k = 4
        # (the global descriptor of 300 random rows is noisy: i+-1 and i+-2 score within the noise of each other)
        p3 = {tuple(p) for p in pm.select_pairs(3).tolist()}
        assert sum((i, i + 1) in p3 for i in range(11)) >= 10
        p4 = {tuple(p) for p in pm.select_pairs(4).tolist()}
        assert all((i, i + 1) in p4 for i in range(11))
        # a subset of the handle's images (pm_select_pairs_among): the restatement over those images alone
        sub = [9, 1, 4, 6, 2, 10, 7]
        ps, Ss = pm.select_pairs(2, want_scores=True, ids=sub)
        srt = sorted(sub)
        ref_s = retrieval_ref.select_pairs([imgs[i] for i in srt], 2)
        assert np.array_equal(ps, np.asarray(srt, np.int32)[ref_s])
        np.testing.assert_allclose(Ss, retrieval_ref.similarity([imgs[i] for i in srt]), rtol=0, atol=1e-12)
        with pytest.raises(api.PairMatchError):
            pm.select_pairs(2, ids=[0, 1, 99])
        # top_k >= n - 1 (or <= 0): FakeImgMatcher's all-pairs list, in the order of the implicit list
        allp = pm.select_pairs(0)
        assert np.array_equal(allp, retrieval_ref.select_pairs(imgs, 0)) and len(allp) == 66
        assert np.array_equal(pm.select_pairs(11), allp)
        # the selected list drives the batched loop
        res = pm.match_all_pairs(pm.select_pairs(2))
        full = pm.match_all_pairs()
        idx = {tuple(p): r for r, p in enumerate(full["pair_ij"].tolist())}
        for p, ij in enumerate(res["pair_ij"].tolist()):
            r = idx[tuple(ij)]
            assert np.array_equal(res["q"][res["offsets"][p]:res["offsets"][p + 1]], full["q"][full["offsets"][r]:full["offsets"][r + 1]])
