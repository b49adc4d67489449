"""ctypes binding of libpairmatch_b200.so (include/pairmatch_b200.h).

Host-side mirror (Python flavour) of the reference's plugin interface for the hot path:

  FeatureMatcher::matchFeatures              Mapper/libMapper/FeatureMatcher.h:18-22
  GeometricFilter::estimateFundamental       Mapper/libMapper/GeometricFilter.h:33-35
  SequentialReconstructor::matchFeatures     Mapper/libMapper/SequentialReconstructor.cpp:199-279

The C++ flavour (what a maintainer of the reference links) is reconstructor_b200/cpp/.
This module never imports the oracle and has no CPU path: if the CUDA library or a device
is missing it raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PM_B200_LIB: A/B runs of two builds on the same GPU box (development aid); default = the in-tree build
LIB_PATH = os.environ.get("PM_B200_LIB") or os.path.join(_HERE, "libpairmatch_b200.so")

DESC_F32, DESC_U8_BITS, DESC_U8 = 0, 1, 2
UNIQUE_FIRST_WINS, MUTUAL_NN, UNIQUE_NONE = 0, 1, 2
RESID_SYMMETRIC_EPIPOLAR, RESID_SAMPSON = 0, 1
SAMPLER_OPENCV_MWC, SAMPLER_PHILOX = 0, 1
PEAK_KIND_F16, PEAK_KIND_I8, PEAK_KIND_MXF4 = 0, 1, 2
ALL_PAIRS = -1
PAIR_UNFILTERED, PAIR_FILTERED, PAIR_DROPPED = 0, 1, 2
OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_OOM, ERR_STATE, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6

EXPORTS = [
    "pm_default_params", "pm_create", "pm_destroy", "pm_last_error", "pm_version", "pm_set_image",
    "pm_set_image_device", "pm_set_image_async", "pm_set_images_async", "pm_set_image_device_async", "pm_sync_images", "pm_num_keypoints", "pm_knn_pair", "pm_match_pair",
    "pm_match_descriptors", "pm_filter_pair_F", "pm_match_filter_pair", "pm_match_all_pairs",
    "pm_free_result", "pm_get_stats", "pm_reset_stats", "pm_measure_popc_peak",
    "pm_save_images", "pm_load_images", "pm_save_result", "pm_load_result",
    "pm_filter_pair_F_seeded", "pm_pair_seed", "pm_remove_image", "pm_measure_tensor_peak", "pm_debug_tc_dump",
    "pm_comm_get_unique_id", "pm_comm_init", "pm_ingest_allgather", "pm_filter_pair_E",
    "pm_select_pairs", "pm_select_pairs_among", "pm_free_pairs",
]


class Params(C.Structure):
    _fields_ = [("ratio", C.c_float), ("unique_mode", C.c_int32), ("min_matches", C.c_int32),
                ("do_filter", C.c_int32), ("ransac_threshold", C.c_double),
                ("ransac_confidence", C.c_double), ("ransac_max_iters", C.c_int32),
                ("residual_mode", C.c_int32), ("sampler", C.c_int32), ("batch_pairs", C.c_int32),
                ("reserve_keypoints", C.c_int64), ("debug_flags", C.c_int32), ("refit_8point", C.c_int32),
                ("seed", C.c_uint64), ("essential_confidence", C.c_double), ("essential_threshold", C.c_double)]


class Camera(C.Structure):
    """pm_camera = PinholeCamera of the reference (Camera.h:127)."""
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("k1", C.c_double), ("k2", C.c_double)]


class PairResult(C.Structure):
    _fields_ = [("capacity", C.c_int32), ("n_matches", C.c_int32), ("n_inliers", C.c_int32),
                ("status", C.c_int32), ("ransac_iters", C.c_int32), ("q", C.POINTER(C.c_int32)),
                ("t", C.POINTER(C.c_int32)), ("inlier", C.POINTER(C.c_uint8)), ("F", C.c_double * 9)]


class CsrResult(C.Structure):
    _fields_ = [("n_pairs", C.c_int64), ("pair_ij", C.POINTER(C.c_int32)),
                ("offsets", C.POINTER(C.c_int64)), ("q", C.POINTER(C.c_int32)),
                ("t", C.POINTER(C.c_int32)), ("inlier", C.POINTER(C.c_uint8)),
                ("F", C.POINTER(C.c_double)), ("status", C.POINTER(C.c_int32)),
                ("n_inliers", C.POINTER(C.c_int32)), ("ransac_iters", C.POINTER(C.c_int32)),
                ("device_ms", C.c_double), ("owner_", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("pairs_matched", C.c_int64), ("putative_matches", C.c_int64),
                ("inlier_matches", C.c_int64), ("kernel_launches", C.c_int64),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("knn_ms", C.c_double),
                ("knn_launches", C.c_int64), ("knn_work", C.c_double), ("device_id", C.c_int32),
                ("n_images", C.c_int32), ("rerank_rows", C.c_int64), ("rerank_chunks", C.c_int64),
                ("rerank_overflow", C.c_int64), ("rerank_worst_err", C.c_double)]


class PairMatchError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pairmatch_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load_library() -> C.CDLL:
    """Loads the in-tree CUDA library; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        lib.pm_last_error.restype = C.c_char_p
        lib.pm_last_error.argtypes = [C.c_void_p]
        lib.pm_version.restype = C.c_char_p
        lib.pm_create.argtypes = [C.POINTER(Params), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]
        lib.pm_destroy.argtypes = [C.c_void_p]
        lib.pm_set_image.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.pm_set_image_device.argtypes = lib.pm_set_image.argtypes
        lib.pm_set_image_async.argtypes = lib.pm_set_image.argtypes
        lib.pm_set_image_device_async.argtypes = lib.pm_set_image.argtypes
        lib.pm_sync_images.argtypes = [C.c_void_p]
        lib.pm_set_images_async.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                            C.c_void_p]
        lib.pm_num_keypoints.argtypes = [C.c_void_p, C.c_int]
        lib.pm_knn_pair.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.pm_debug_tc_dump.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pm_match_pair.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(PairResult)]
        lib.pm_match_filter_pair.argtypes = lib.pm_match_pair.argtypes
        lib.pm_match_descriptors.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                             C.c_int, C.POINTER(PairResult)]
        lib.pm_filter_pair_F.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                         C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        lib.pm_filter_pair_F_seeded.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_void_p,
                                                C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        lib.pm_filter_pair_E.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Camera), C.POINTER(Camera),
                                         C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        lib.pm_pair_seed.argtypes = [C.c_uint64, C.c_int32, C.c_int32]
        lib.pm_pair_seed.restype = C.c_uint64
        lib.pm_remove_image.argtypes = [C.c_void_p, C.c_int]
        lib.pm_select_pairs.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int64), C.c_void_p]
        lib.pm_select_pairs_among.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_int32)),
                                              C.POINTER(C.c_int64), C.c_void_p]
        lib.pm_free_pairs.argtypes = [C.POINTER(C.c_int32)]
        lib.pm_comm_get_unique_id.argtypes = [C.c_void_p]
        lib.pm_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.pm_ingest_allgather.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_int]
        lib.pm_measure_tensor_peak.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        lib.pm_match_all_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_int64,
                                           C.POINTER(C.POINTER(CsrResult))]
        lib.pm_free_result.argtypes = [C.POINTER(CsrResult)]
        lib.pm_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        lib.pm_reset_stats.argtypes = [C.c_void_p]
        lib.pm_measure_popc_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        lib.pm_save_images.argtypes = [C.c_void_p, C.c_char_p]
        lib.pm_load_images.argtypes = [C.c_void_p, C.c_char_p]
        lib.pm_save_result.argtypes = [C.POINTER(CsrResult), C.c_char_p]
        lib.pm_load_result.argtypes = [C.c_char_p, C.POINTER(C.POINTER(CsrResult))]
        _lib = lib
    return _lib


def pair_seed(seed: int, i: int, j: int) -> int:
    """The Philox key the batched loop derives for pair (i, j) under pm_params.seed."""
    return int(load_library().pm_pair_seed(seed, i, j))


def comm_unique_id() -> bytes:
    """pm_comm_get_unique_id (rank 0); ship the bytes to the other ranks by any means."""
    buf = (C.c_uint8 * 128)()
    rc = load_library().pm_comm_get_unique_id(buf)
    if rc != OK:
        raise PairMatchError(rc, load_library().pm_last_error(None).decode())
    return bytes(buf)


def default_params() -> Params:
    p = Params()
    load_library().pm_default_params(C.byref(p))
    return p


def _desc_args(desc: np.ndarray, dtype: int | None):
    """Maps a numpy descriptor matrix to (contiguous array, dim, dtype enum)."""
    if dtype is None:
        dtype = DESC_U8_BITS if desc.dtype == np.uint8 else DESC_F32
    if dtype == DESC_U8_BITS:
        a = np.ascontiguousarray(desc, np.uint8)
        return a, a.shape[1] * 8, dtype
    if dtype == DESC_U8:
        a = np.ascontiguousarray(desc, np.uint8)
        return a, a.shape[1], dtype
    a = np.ascontiguousarray(desc, np.float32)
    return a, a.shape[1], dtype


class PairMatcher:
    """Handle on the device pipeline (one per process; thread-safe like the reference's plugins)."""

    def __init__(self, devices=None, **kw):
        self.lib = load_library()
        self.params = default_params()
        for k, v in kw.items():
            if not hasattr(self.params, k):
                raise TypeError(f"unknown parameter {k}")
            setattr(self.params, k, v)
        self.h = C.c_void_p()
        if devices is None:
            rc = self.lib.pm_create(C.byref(self.params), None, 0, C.byref(self.h))
        else:
            arr = (C.c_int * len(devices))(*devices)
            rc = self.lib.pm_create(C.byref(self.params), arr, len(devices), C.byref(self.h))
        if rc != OK:
            raise PairMatchError(rc, self.lib.pm_last_error(None).decode())
        self._n = {}

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.pm_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != OK:
            raise PairMatchError(rc, self.lib.pm_last_error(self.h).decode())

    # -- ingest ---------------------------------------------------------------------------------
    def set_image(self, img_id: int, desc: np.ndarray, xy: np.ndarray | None = None, dtype=None):
        a, dim, dt = _desc_args(desc, dtype)
        xyp = None
        if xy is not None:
            xy = np.ascontiguousarray(xy, np.int32)
            assert xy.shape == (a.shape[0], 2)
            xyp = xy.ctypes.data
        self._check(self.lib.pm_set_image(self.h, img_id, a.ctypes.data, a.shape[0], dim, dt, xyp))
        self._n[img_id] = a.shape[0]

    def set_image_ptr(self, img_id: int, desc_ptr: int, n: int, dim: int, dtype: int, xy_ptr: int | None,
                      on_device=False, asynchronous=False):
        """Raw-pointer variant (pinned host buffers or device buffers).  asynchronous: pm_set_image_async -- the
        buffers must stay alive and unchanged until sync_images() or a matching call that uses the image."""
        if on_device:
            f = self.lib.pm_set_image_device_async if asynchronous else self.lib.pm_set_image_device
        else:
            f = self.lib.pm_set_image_async if asynchronous else self.lib.pm_set_image
        self._check(f(self.h, img_id, desc_ptr, n, dim, dtype, xy_ptr))
        self._n[img_id] = n

    def set_images_ptr_async(self, ids, desc_ptrs, ns, dim: int, dtype: int, xy_ptrs=None):
        """pm_set_images_async: one call for a whole image set held in (pinned) host buffers."""
        n = len(ids)
        a_ids = (C.c_int * n)(*ids)
        a_desc = (C.c_void_p * n)(*desc_ptrs)
        a_ns = (C.c_int * n)(*ns)
        a_xy = (C.c_void_p * n)(*xy_ptrs) if xy_ptrs is not None else None
        self._check(self.lib.pm_set_images_async(self.h, n, a_ids, a_desc, a_ns, dim, dtype, a_xy))
        for i, k in zip(ids, ns):
            self._n[i] = k

    def sync_images(self):
        self._check(self.lib.pm_sync_images(self.h))

    # -- collective ingest (extraction sharded over ranks, one process per GPU) --------------------
    def comm_init(self, unique_id: bytes, rank: int, n_ranks: int):
        """pm_comm_init: unique_id = the 128 bytes of comm_unique_id() from rank 0."""
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(self.lib.pm_comm_init(self.h, buf, rank, n_ranks))
        self._comm = (rank, n_ranks)

    def ingest_allgather(self, n_images_total: int, n_keypoints: int, dim: int, dtype: int, wire_dtype: int,
                         own_desc_ptr: int, own_xy_ptr: int | None, own_on_device=False):
        """pm_ingest_allgather: image k lives on rank k % n_ranks; own_* = this rank's images in ascending id order."""
        self._check(self.lib.pm_ingest_allgather(self.h, n_images_total, n_keypoints, dim, dtype, wire_dtype,
                                                 own_desc_ptr, own_xy_ptr, 1 if own_on_device else 0))
        for i in range(n_images_total):
            self._n[i] = n_keypoints

    # -- per-pair (compat path of the virtual calls) ---------------------------------------------
    def knn_pair(self, i: int, j: int):
        n = self._n[i]
        idx = np.empty((n, 2), np.int32); dist = np.empty((n, 2), np.float32)
        self._check(self.lib.pm_knn_pair(self.h, i, j, idx.ctypes.data, dist.ctypes.data))
        return idx, dist

    def debug_tc_dump(self, i: int, j: int):
        n = self._n[i]
        idx = np.empty((n, 2), np.int32); dist = np.empty((n, 2), np.float32)
        acc = np.zeros((256, 128), np.float32)
        self._check(self.lib.pm_debug_tc_dump(self.h, i, j, idx.ctypes.data, dist.ctypes.data, acc.ctypes.data))
        return idx, dist, acc

    def _pair(self, fn, *args, cap):
        q = np.empty(max(cap, 1), np.int32); t = np.empty(max(cap, 1), np.int32)
        inl = np.zeros(max(cap, 1), np.uint8)
        r = PairResult()
        r.capacity = cap
        r.q = q.ctypes.data_as(C.POINTER(C.c_int32)); r.t = t.ctypes.data_as(C.POINTER(C.c_int32))
        r.inlier = inl.ctypes.data_as(C.POINTER(C.c_uint8))
        self._check(fn(self.h, *args, C.byref(r)))
        m = r.n_matches
        return dict(q=q[:m].copy(), t=t[:m].copy(), inlier=inl[:m].copy(), status=r.status,
                    n_inliers=r.n_inliers, iters=r.ransac_iters, F=np.array(r.F[:]).reshape(3, 3))

    def match_pair(self, i: int, j: int):
        """FeatureMatcher::matchFeatures for resident images -> dict(q, t)."""
        return self._pair(self.lib.pm_match_pair, i, j, cap=self._n[i])

    def match_filter_pair(self, i: int, j: int):
        """Pair body of SequentialReconstructor::matchFeatures (match, >=7 gate, filter)."""
        return self._pair(self.lib.pm_match_filter_pair, i, j, cap=self._n[i])

    def match_descriptors(self, desc1: np.ndarray, desc2: np.ndarray, dtype=None):
        a, dim, dt = _desc_args(desc1, dtype)
        b, dim2, _ = _desc_args(desc2, dtype)
        assert dim == dim2
        return self._pair(self.lib.pm_match_descriptors, a.ctypes.data, a.shape[0], b.ctypes.data,
                          b.shape[0], dim, dt, cap=a.shape[0])

    def estimate_fundamental(self, xy1: np.ndarray, xy2: np.ndarray, pair_key: int | None = None):
        """GeometricFilter::estimateFundamental -> (F 3x3, mask uint8 [M], status, iters).
        pair_key: explicit Philox key of this call (pm_filter_pair_F_seeded; only read with SAMPLER_PHILOX)."""
        p1 = np.ascontiguousarray(xy1, np.float32); p2 = np.ascontiguousarray(xy2, np.float32)
        m = p1.shape[0]
        F = np.zeros(9, np.float64); mask = np.zeros(max(m, 1), np.uint8)
        st = C.c_int32(0); it = C.c_int32(0)
        if pair_key is None:
            self._check(self.lib.pm_filter_pair_F(self.h, p1.ctypes.data, p2.ctypes.data, m, F.ctypes.data,
                                                  mask.ctypes.data, C.byref(st), C.byref(it)))
        else:
            self._check(self.lib.pm_filter_pair_F_seeded(self.h, p1.ctypes.data, p2.ctypes.data, m, pair_key,
                                                         F.ctypes.data, mask.ctypes.data, C.byref(st), C.byref(it)))
        return F.reshape(3, 3), mask[:m], st.value, it.value

    def estimate_essential(self, xy1: np.ndarray, xy2: np.ndarray, cam1: Camera, cam2: Camera):
        """GeometricFilter::estimateEssential -> (E 3x3 of unit norm, mask uint8 [M], status, iters)."""
        p1 = np.ascontiguousarray(xy1, np.float32); p2 = np.ascontiguousarray(xy2, np.float32)
        m = p1.shape[0]
        E = np.zeros(9, np.float64); mask = np.zeros(max(m, 1), np.uint8)
        st = C.c_int32(0); it = C.c_int32(0)
        self._check(self.lib.pm_filter_pair_E(self.h, p1.ctypes.data, p2.ctypes.data, m, C.byref(cam1), C.byref(cam2),
                                              E.ctypes.data, mask.ctypes.data, C.byref(st), C.byref(it)))
        return E.reshape(3, 3), mask[:m], st.value, it.value

    def remove_image(self, img_id: int):
        self._check(self.lib.pm_remove_image(self.h, img_id))
        self._n.pop(img_id, None)

    # -- batched loop ----------------------------------------------------------------------------
    def match_all_pairs(self, pairs: np.ndarray | None = None, copy=True, save_to: str | None = None):
        """Whole pair loop.  Returns dict of numpy arrays (CSR): pair_ij, offsets, q, t, inlier, F,
        status, n_inliers, ransac_iters, device_ms.  save_to: also write the result cache file."""
        res = C.POINTER(CsrResult)()
        if pairs is None:
            self._check(self.lib.pm_match_all_pairs(self.h, None, ALL_PAIRS, C.byref(res)))
        else:
            pairs = np.ascontiguousarray(pairs, np.int32)
            self._check(self.lib.pm_match_all_pairs(self.h, pairs.ctypes.data, pairs.shape[0], C.byref(res)))
        if save_to is not None:
            self._check(self.lib.pm_save_result(res, os.fsencode(save_to)))
        return _csr_to_dict(self.lib, res, copy)

    def select_pairs(self, top_k: int, want_scores=False, ids=None):
        """pm_select_pairs[_among]: pair pre-selection by global-descriptor retrieval (the ImageMatcher plugin,
        ImageMatcher.h:18-21) over every image of the handle, or over `ids`.
        Returns pairs int32 [P, 2] (and the n x n similarity matrix in ascending image-id order)."""
        ptr = C.POINTER(C.c_int32)(); n = C.c_int64(0)
        scores = None
        ids_a = None if ids is None else np.ascontiguousarray(ids, np.int32)
        if want_scores:
            k = len(self._n) if ids_a is None else len(ids_a)
            scores = np.zeros((k, k), np.float64)
        sp = scores.ctypes.data if want_scores else None
        if ids_a is None:
            self._check(self.lib.pm_select_pairs(self.h, top_k, C.byref(ptr), C.byref(n), sp))
        else:
            self._check(self.lib.pm_select_pairs_among(self.h, ids_a.ctypes.data, len(ids_a), top_k, C.byref(ptr), C.byref(n), sp))
        pairs = np.ctypeslib.as_array(ptr, shape=(max(n.value, 1) * 2,))[:2 * n.value].copy().reshape(-1, 2)
        self.lib.pm_free_pairs(ptr)
        return (pairs, scores) if want_scores else pairs

    # -- on-disk cache ---------------------------------------------------------------------------
    def save_images(self, path: str):
        self._check(self.lib.pm_save_images(self.h, os.fsencode(path)))

    def load_images(self, path: str):
        """Ingests every image of a cache file (pm_save_images or cache.write_images)."""
        self._check(self.lib.pm_load_images(self.h, os.fsencode(path)))
        from . import cache
        for rec in cache.read_images(path, headers_only=True):
            self._n[rec["id"]] = rec["n"]

    def free_result(self, out):
        if "_handle" in out:
            self.lib.pm_free_result(out.pop("_handle"))

    def stats(self) -> dict:
        s = Stats()
        self._check(self.lib.pm_get_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def reset_stats(self):
        self._check(self.lib.pm_reset_stats(self.h))

    def measure_popc_peak(self) -> float:
        v = C.c_double(0)
        self._check(self.lib.pm_measure_popc_peak(self.h, C.byref(v)))
        return v.value

    def measure_tensor_peak(self, kind: int) -> float:
        """FLOP/s of the pure tcgen05.mma issue loop of one MMA kind (PEAK_KIND_*)."""
        v = C.c_double(0)
        self._check(self.lib.pm_measure_tensor_peak(self.h, kind, C.byref(v)))
        return v.value


def _csr_to_dict(lib, res, copy=True) -> dict:
    r = res.contents
    n = r.n_pairs

    def arr(ptr, count, dt):
        if count == 0:
            return np.empty(0, dt)
        a = np.ctypeslib.as_array(ptr, shape=(count,))
        return a.copy() if copy else a
    out = dict(n_pairs=n, device_ms=r.device_ms)
    out["offsets"] = arr(r.offsets, n + 1, np.int64)
    total = int(out["offsets"][n]) if n >= 0 else 0
    out["pair_ij"] = arr(r.pair_ij, 2 * n, np.int32).reshape(-1, 2)
    out["q"] = arr(r.q, total, np.int32); out["t"] = arr(r.t, total, np.int32)
    out["inlier"] = arr(r.inlier, total, np.uint8)
    out["F"] = arr(r.F, 9 * n, np.float64).reshape(-1, 3, 3)
    out["status"] = arr(r.status, n, np.int32)
    out["n_inliers"] = arr(r.n_inliers, n, np.int32)
    out["ransac_iters"] = arr(r.ransac_iters, n, np.int32)
    if copy:
        lib.pm_free_result(res)
    else:
        out["_handle"] = res
    return out


def load_result(path: str) -> dict:
    """Reads a result cache file through the C ABI (no device needed)."""
    lib = load_library()
    res = C.POINTER(CsrResult)()
    rc = lib.pm_load_result(os.fsencode(path), C.byref(res))
    if rc != 0:
        raise PairMatchError(rc, lib.pm_last_error(None).decode())
    return _csr_to_dict(lib, res, True)


def feature_matches_view(res: dict, mirror=True) -> dict:
    """Materialises the reference's `featureMatches` container
    (unordered_map<pair<int,int>, unordered_map<int,int>>, SequentialReconstructor.h:226) from the
    CSR result: entries exist for (i,j) and, mirrored, for (j,i) (.cpp:219-227); dropped pairs
    have no entry (.cpp:253-256); pairs without any surviving match have no entry either."""
    fm = {}
    off = res["offsets"]
    for p in range(res["n_pairs"]):
        if res["status"][p] == PAIR_DROPPED:
            continue
        a, b = int(off[p]), int(off[p + 1])
        keep = res["inlier"][a:b].astype(bool)
        q = res["q"][a:b][keep]; t = res["t"][a:b][keep]
        if len(q) == 0:
            continue
        i, j = map(int, res["pair_ij"][p])
        fm[(i, j)] = dict(zip(q.tolist(), t.tolist()))
        if mirror:
            fm[(j, i)] = dict(zip(t.tolist(), q.tolist()))
    return fm
