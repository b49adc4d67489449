"""ctypes binding of the C oracle (oracle/libpm_oracle.so).  TEST INFRASTRUCTURE ONLY.

Every function cites the reference lines it follows in oracle/pm_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpm_oracle.so")

UNIQUE_FIRST_WINS, MUTUAL_NN, UNIQUE_NONE = 0, 1, 2
RESID_SYMMETRIC_EPIPOLAR, RESID_SAMPSON = 0, 1


SAMPLER_OPENCV_MWC, SAMPLER_PHILOX = 0, 1


class RansacParams(C.Structure):
    _fields_ = [("threshold", C.c_double), ("confidence", C.c_double),
                ("max_iters", C.c_int), ("residual_mode", C.c_int),
                ("sampler", C.c_int), ("refit_8point", C.c_int), ("seed", C.c_uint64)]


class RansacTrace(C.Structure):
    _fields_ = [("iters_run", C.c_int), ("niters_final", C.c_int), ("best_iter", C.c_int),
                ("best_model", C.c_int), ("best_count", C.c_int), ("models_tested", C.c_int)]


def build(force: bool = False) -> str:
    """Compile the C oracle (gcc).  Building the checker is not using it."""
    src = os.path.join(_HERE, "pm_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)
             or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src),
                                                  os.path.getmtime(os.path.join(_HERE, "pm_essential.c")),
                                                  os.path.getmtime(os.path.join(_HERE, "pm_oracle.h"))))
    if force or stale:
        env = dict(os.environ)
        env.pop("CC", None)
        r = subprocess.run(["make", "-C", _HERE, "-B"], env=env, capture_output=True, text=True)
        if r.returncode != 0:
            # toolchains without libgomp: rebuild single-threaded
            r = subprocess.run(["make", "-C", _HERE, "-B",
                                "CFLAGS=-O2 -fPIC -ffp-contract=off -fno-fast-math -std=gnu11 "
                                "-Wno-unknown-pragmas -mpopcnt"],
                               env=env, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_find_fundamental.restype = C.c_int
        _lib.orc_match_pair.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def default_params(threshold=3.0, confidence=0.99, max_iters=1000,
                   residual_mode=RESID_SYMMETRIC_EPIPOLAR, sampler=SAMPLER_OPENCV_MWC, refit_8point=0,
                   seed=0) -> RansacParams:
    return RansacParams(threshold, confidence, max_iters, residual_mode, sampler, refit_8point, seed)


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    lib().orc_philox4x32_10(c, C.c_uint32(key[0]), C.c_uint32(key[1]))
    return list(c)


def pair_seed(seed: int, i: int, j: int) -> int:
    f = lib().orc_pair_seed
    f.restype = C.c_uint64
    f.argtypes = [C.c_uint64, C.c_int32, C.c_int32]
    return int(f(seed, i, j))


def philox_subset(xy1, xy2, seed: int, it: int):
    xy1 = np.ascontiguousarray(xy1, np.float32); xy2 = np.ascontiguousarray(xy2, np.float32)
    idx = np.zeros(7, np.int32)
    f = lib().orc_philox_subset
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_void_p]
    ok = f(xy1.ctypes.data, xy2.ctypes.data, xy1.shape[0], seed, it, idx.ctypes.data)
    return bool(ok), idx


def eight_point(xy1, xy2, mask=None):
    xy1 = np.ascontiguousarray(xy1, np.float32); xy2 = np.ascontiguousarray(xy2, np.float32)
    F = np.zeros(9, np.float64)
    mp = None
    if mask is not None:
        mask = np.ascontiguousarray(mask, np.uint8)
        mp = _p(mask, C.c_uint8)
    ok = lib().orc_eight_point(_p(xy1, C.c_float), _p(xy2, C.c_float), xy1.shape[0], mp, _p(F, C.c_double))
    return bool(ok), F.reshape(3, 3)


def knn2_hamming(q: np.ndarray, t: np.ndarray):
    """q,t: uint8 [n, nbytes].  Returns (idx int32 [nq,2], dist int32 [nq,2])."""
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    nq, nb = q.shape; nt = t.shape[0]
    idx = np.empty((nq, 2), np.int32); dist = np.empty((nq, 2), np.int32)
    rc = lib().orc_knn2_hamming(_p(q, C.c_uint8), nq, _p(t, C.c_uint8), nt, nb,
                                _p(idx, C.c_int32), _p(dist, C.c_int32))
    assert rc == 0
    return idx, dist


def knn2_l2(q: np.ndarray, t: np.ndarray):
    """q,t: float32 [n, dim].  Returns (idx int32 [nq,2], d2 float64 [nq,2]) (squared)."""
    q = np.ascontiguousarray(q, np.float32); t = np.ascontiguousarray(t, np.float32)
    nq, dim = q.shape; nt = t.shape[0]
    idx = np.empty((nq, 2), np.int32); d2 = np.empty((nq, 2), np.float64)
    rc = lib().orc_knn2_l2(_p(q, C.c_float), nq, _p(t, C.c_float), nt, dim,
                           _p(idx, C.c_int32), _p(d2, C.c_double))
    assert rc == 0
    return idx, d2


def best_query(q: np.ndarray, t: np.ndarray) -> np.ndarray:
    nt = t.shape[0]
    out = np.empty(nt, np.int32)
    if q.dtype == np.uint8:
        q = np.ascontiguousarray(q); t = np.ascontiguousarray(t)
        lib().orc_best_query_hamming(_p(q, C.c_uint8), q.shape[0], _p(t, C.c_uint8), nt,
                                     q.shape[1], _p(out, C.c_int32))
    else:
        q = np.ascontiguousarray(q, np.float32); t = np.ascontiguousarray(t, np.float32)
        lib().orc_best_query_l2(_p(q, C.c_float), q.shape[0], _p(t, C.c_float), nt,
                                q.shape[1], _p(out, C.c_int32))
    return out


def ratio_unique(idx, dist, nt, ratio=0.7, mode=UNIQUE_FIRST_WINS, best_q=None):
    idx = np.ascontiguousarray(idx, np.int32); dist = np.ascontiguousarray(dist, np.float32)
    nq = idx.shape[0]
    oq = np.empty(max(nq, 1), np.int32); ot = np.empty(max(nq, 1), np.int32)
    bq = None
    if best_q is not None:
        best_q = np.ascontiguousarray(best_q, np.int32)
        bq = _p(best_q, C.c_int32)
    n = lib().orc_ratio_unique(_p(idx, C.c_int32), _p(dist, C.c_float), nq, nt,
                               C.c_float(ratio), mode, bq, _p(oq, C.c_int32), _p(ot, C.c_int32))
    return oq[:n].copy(), ot[:n].copy()


def find_fundamental(xy1, xy2, params: RansacParams | None = None):
    """Returns (n_solutions, F [n_solutions or 1, 3, 3], mask uint8 [n], trace)."""
    xy1 = np.ascontiguousarray(xy1, np.float32); xy2 = np.ascontiguousarray(xy2, np.float32)
    n = xy1.shape[0]
    prm = params or default_params()
    F = np.zeros(27, np.float64); mask = np.zeros(max(n, 1), np.uint8); tr = RansacTrace()
    ns = lib().orc_find_fundamental(_p(xy1, C.c_float), _p(xy2, C.c_float), n, C.byref(prm),
                                    _p(F, C.c_double), _p(mask, C.c_uint8), C.byref(tr))
    return ns, F.reshape(3, 3, 3)[:max(ns, 1)].copy(), mask[:n].copy(), tr


def solve_cubic(coeffs):
    c = np.ascontiguousarray(coeffs, np.float64); r = np.zeros(3, np.float64)
    n = lib().orc_solve_cubic(_p(c, C.c_double), _p(r, C.c_double))
    return n, r


def seven_point(xy1, xy2):
    xy1 = np.ascontiguousarray(xy1, np.float32); xy2 = np.ascontiguousarray(xy2, np.float32)
    F = np.zeros(27, np.float64)
    n = lib().orc_seven_point(_p(xy1, C.c_float), _p(xy2, C.c_float), _p(F, C.c_double))
    return n, F.reshape(3, 3, 3)[:max(n, 0)].copy()


def residuals(F, xy1, xy2, mode=RESID_SYMMETRIC_EPIPOLAR):
    F = np.ascontiguousarray(F, np.float64).reshape(9)
    xy1 = np.ascontiguousarray(xy1, np.float32); xy2 = np.ascontiguousarray(xy2, np.float32)
    n = xy1.shape[0]; err = np.empty(n, np.float32)
    lib().orc_residuals(_p(F, C.c_double), _p(xy1, C.c_float), _p(xy2, C.c_float), n, mode,
                        _p(err, C.c_float))
    return err


def update_num_iters(p, ep, model_points, max_iters):
    f = lib().orc_update_num_iters
    f.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int]
    return f(p, ep, model_points, max_iters)


def sample_subsets(xy1, xy2, iters):
    xy1 = np.ascontiguousarray(xy1, np.float32); xy2 = np.ascontiguousarray(xy2, np.float32)
    out = np.zeros((iters, 7), np.int32)
    k = lib().orc_sample_subsets(_p(xy1, C.c_float), _p(xy2, C.c_float), xy1.shape[0], iters,
                                 _p(out, C.c_int32))
    return out[:k]


def match_pair(desc1, xy1, desc2, xy2, ratio=0.7, unique_mode=UNIQUE_FIRST_WINS,
               min_matches=7, params: RansacParams | None = None, do_filter=True):
    """Whole per-pair body.  Returns dict(status, q, t, n_putative, F).

    status: 'ok' (q,t = surviving matches) or 'dropped' (F estimation failed)."""
    kind = 1 if desc1.dtype == np.uint8 else 0
    if kind == 1:
        desc1 = np.ascontiguousarray(desc1, np.uint8); desc2 = np.ascontiguousarray(desc2, np.uint8)
        d1p, d2p = _p(desc1, C.c_uint8), _p(desc2, C.c_uint8)
    else:
        desc1 = np.ascontiguousarray(desc1, np.float32); desc2 = np.ascontiguousarray(desc2, np.float32)
        d1p, d2p = _p(desc1, C.c_float), _p(desc2, C.c_float)
    xy1 = np.ascontiguousarray(xy1, np.int32); xy2 = np.ascontiguousarray(xy2, np.int32)
    n1, dim = desc1.shape; n2 = desc2.shape[0]
    oq = np.empty(max(n1, 1), np.int32); ot = np.empty(max(n1, 1), np.int32)
    nput = C.c_int(0); F = np.zeros(9, np.float64)
    prm = params or default_params()
    n = lib().orc_match_pair(kind, C.cast(d1p, C.c_void_p), _p(xy1, C.c_int32), n1,
                             C.cast(d2p, C.c_void_p), _p(xy2, C.c_int32), n2, dim,
                             C.c_float(ratio), unique_mode, min_matches,
                             C.byref(prm) if do_filter else None,
                             _p(oq, C.c_int32), _p(ot, C.c_int32), C.byref(nput),
                             _p(F, C.c_double))
    if n < 0:
        return dict(status="dropped", q=oq[:0].copy(), t=ot[:0].copy(),
                    n_putative=nput.value, F=F.reshape(3, 3))
    return dict(status="ok", q=oq[:n].copy(), t=ot[:n].copy(), n_putative=nput.value,
                F=F.reshape(3, 3))


class Camera(C.Structure):
    """PinholeCamera (Camera.h:127): fx, fy, cx, cy, k1, k2."""
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("k1", C.c_double), ("k2", C.c_double)]


def five_point(m1, m2):
    m1 = np.ascontiguousarray(m1, np.float64); m2 = np.ascontiguousarray(m2, np.float64)
    Es = np.zeros(90, np.float64)
    n = lib().orc_five_point(_p(m1, C.c_double), _p(m2, C.c_double), _p(Es, C.c_double))
    return Es.reshape(10, 3, 3)[:n].copy()


def normalize_for_essential(xy, own: Camera, c1: Camera, c2: Camera):
    xy = np.ascontiguousarray(xy, np.float32)
    out = np.zeros((xy.shape[0], 2), np.float64)
    lib().orc_normalize_for_essential(_p(xy, C.c_float), xy.shape[0], C.byref(own), C.byref(c1), C.byref(c2),
                                      _p(out, C.c_double))
    return out


def find_essential(xy1, xy2, cam1: Camera, cam2: Camera, prob=0.999, threshold=1.0, max_iters=1000,
                   sampler=SAMPLER_OPENCV_MWC, seed=0):
    """cv::findEssentialMat(p1, p2, K1, d1, K2, d2) restated.  Returns (ok, E 3x3, mask uint8 [n], trace)."""
    xy1 = np.ascontiguousarray(xy1, np.float32); xy2 = np.ascontiguousarray(xy2, np.float32)
    n = xy1.shape[0]
    E = np.zeros(9, np.float64); mask = np.zeros(max(n, 1), np.uint8); tr = RansacTrace()
    f = lib().orc_find_essential
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Camera), C.POINTER(Camera), C.c_double, C.c_double,
                  C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.POINTER(RansacTrace)]
    ok = f(xy1.ctypes.data, xy2.ctypes.data, n, C.byref(cam1), C.byref(cam2), prob, threshold, max_iters, sampler,
           seed, E.ctypes.data, mask.ctypes.data, C.byref(tr))
    return bool(ok), E.reshape(3, 3), mask[:n].copy(), tr
