"""The reference's per-pair body restated over Python cv2.  TEST INFRASTRUCTURE ONLY.

The reference (C++17) cannot be compiled in this image (no OpenCV/Eigen/PCL/Ceres C++
packages), but all arithmetic on its hot path is inside three OpenCV calls, and Python cv2
(4.13.0 here) exposes exactly those:

  FlannMatcher::matchFeatures            Mapper/libMapper/FeatureMatcher.cpp:32-65
      cv::DescriptorMatcher::create(FLANNBASED) :29, knnMatch(.., 2) :49,
      ratio 0.7 :55 (FeatureMatcher.h:45), first-come uniqueness :58-62
  GeometricFilter::estimateFundamental   Mapper/libMapper/GeometricFilter.cpp:39-61
      cv::findFundamentalMat(p1, p2, mask) :47, empty F => no mask :50-53
  pair body                              Mapper/libMapper/SequentialReconstructor.cpp:213-276

Used (a) to pin oracle/pm_oracle.c, (b) by tests/golden/make_golden.py to generate the
committed golden vectors, (c) as the timed CPU arm of bench.py (--impl reference and
cpu_baseline), one worker process per core like the reference's OpenMP-over-pairs loop.
"""
from __future__ import annotations

import numpy as np

try:  # cv2 is in the image; gate so that importing the module never fails
    import cv2
    cv2.setNumThreads(1)
except Exception:  # pragma: no cover
    cv2 = None

RATIO = np.float32(0.7)          # FeatureMatcher.h:45
MIN_MATCHES = 7                  # SequentialReconstructor.cpp:237


def have_cv2() -> bool:
    return cv2 is not None


def knn2_flann(desc1: np.ndarray, desc2: np.ndarray):
    """FeatureMatcher.cpp:27-30,48-49 -- what the reference really runs (approximate)."""
    m = cv2.DescriptorMatcher_create(cv2.DescriptorMatcher_FLANNBASED)
    return _unpack(m.knnMatch(np.ascontiguousarray(desc1, np.float32),
                              np.ascontiguousarray(desc2, np.float32), 2), desc1.shape[0])


def knn2_bf(desc1: np.ndarray, desc2: np.ndarray):
    """cv::BFMatcher with the matching norm -- the exact comparator of the north star."""
    if desc1.dtype == np.uint8:
        m = cv2.BFMatcher(cv2.NORM_HAMMING)
        return _unpack(m.knnMatch(np.ascontiguousarray(desc1), np.ascontiguousarray(desc2), 2),
                       desc1.shape[0])
    m = cv2.BFMatcher(cv2.NORM_L2)
    return _unpack(m.knnMatch(np.ascontiguousarray(desc1, np.float32),
                              np.ascontiguousarray(desc2, np.float32), 2), desc1.shape[0])


def _unpack(knn, nq):
    """DMatch rows -> (idx, dist) arrays.  The DMatch objects are what cv2 hands back (the C++ reference reads the same
    fields in place); flat generator expressions keep the Python share of the timed CPU arm at ~1 % of a pair."""
    if nq and all(len(row) == 2 for row in knn):
        idx = np.fromiter((m.trainIdx for row in knn for m in row), np.int32, 2 * nq).reshape(nq, 2)
        dist = np.fromiter((m.distance for row in knn for m in row), np.float32, 2 * nq).reshape(nq, 2)
        return idx, dist
    idx = np.full((nq, 2), -1, np.int32)
    dist = np.full((nq, 2), np.inf, np.float32)
    for i, row in enumerate(knn):
        for k, dm in enumerate(row[:2]):
            idx[i, k] = dm.trainIdx
            dist[i, k] = dm.distance
    return idx, dist


def ratio_unique(idx: np.ndarray, dist: np.ndarray, ratio=RATIO):
    """FeatureMatcher.cpp:51-64 (float compare on non-squared distances, first query wins), vectorised:
    the first occurrence of every train index among the rows that pass, in ascending query order."""
    ratio = np.float32(ratio)
    d = np.asarray(dist, np.float32)
    ok = (idx[:, 0] >= 0) & (idx[:, 1] >= 0) & (d[:, 0] < (ratio * d[:, 1]).astype(np.float32))
    q = np.nonzero(ok)[0]
    t = idx[q, 0]
    _, first = np.unique(t, return_index=True)
    first.sort()
    return q[first].astype(np.int32), t[first].astype(np.int32)


def ratio_unique_loop(idx: np.ndarray, dist: np.ndarray, ratio=RATIO):
    """The same as the literal loop of FeatureMatcher.cpp:51-64 (kept as the cross-check of ratio_unique)."""
    q_out, t_out, seen = [], [], set()
    ratio = np.float32(ratio)
    for i in range(idx.shape[0]):
        if idx[i, 0] < 0 or idx[i, 1] < 0:
            continue
        if np.float32(dist[i, 0]) < np.float32(ratio * np.float32(dist[i, 1])):
            t = int(idx[i, 0])
            if t not in seen:
                seen.add(t)
                q_out.append(i)
                t_out.append(t)
    return np.asarray(q_out, np.int32), np.asarray(t_out, np.int32)


def estimate_fundamental(xy1: np.ndarray, xy2: np.ndarray):
    """GeometricFilter.cpp:39-61.  Returns (F 3x3 or None, mask uint8 [n] or None)."""
    p1 = np.ascontiguousarray(xy1, np.float32).reshape(-1, 1, 2)   # utils.cpp:165-177
    p2 = np.ascontiguousarray(xy2, np.float32).reshape(-1, 1, 2)
    F, mask = cv2.findFundamentalMat(p1, p2)
    if F is None or F.size == 0:
        return None, None
    return np.asarray(F[:3, :3], np.float64), mask.reshape(-1).astype(np.uint8)  # utils.cpp:192-203


def match_pair(desc1, xy1, desc2, xy2, matcher="flann", do_filter=True):
    """SequentialReconstructor.cpp:213-276 for one (i, j) pair.

    Returns dict(status, q, t, n_putative)."""
    if matcher == "flann":
        idx, dist = knn2_flann(desc1, desc2)
    else:
        idx, dist = knn2_bf(desc1, desc2)
    q, t = ratio_unique(idx, dist)
    nput = len(q)
    if do_filter and nput >= MIN_MATCHES:
        F, mask = estimate_fundamental(np.asarray(xy1)[q], np.asarray(xy2)[t])
        if F is None:
            return dict(status="dropped", q=q[:0], t=t[:0], n_putative=nput)
        keep = mask.astype(bool)
        return dict(status="ok", q=q[keep], t=t[keep], n_putative=nput, F=F)
    return dict(status="ok", q=q, t=t, n_putative=nput)


def time_pair_body(desc1, xy1, desc2, xy2, matcher="flann"):
    """One pair through the reference body; returns the number of surviving matches.
    Used by the bench's CPU arm (only the cv2 calls + the ratio/unique loop are inside)."""
    r = match_pair(desc1, xy1, desc2, xy2, matcher=matcher)
    return len(r["q"])
