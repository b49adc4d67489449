"""Synthetic image sets for the pair-matching path.

Counter-based RNG (Philox) keyed by (seed, image_id): any rank regenerates any image
identically, nothing is shipped.  A world of L = 4*N landmarks (base descriptor + 3-D point)
is observed by cameras on a ring; every image draws N landmarks without replacement, so a
pair shares ~N/4 true correspondences.  Keypoints are projections + N(0, 0.5 px) noise,
truncated to int like the reference's detector does (FeatureDetector.cpp:28-29).

Descriptor value domains follow the reference's producers:
  'sift'       128 floats, integer-valued in [0,255]   (FeatureDetector.cpp:20-24)
  'superpoint' 256 floats, unit L2 norm                (FeatureSuperPoint.cpp:195-205)
  'orb'        32 bytes = 256 bits                     (FeatureDetector.cpp:9,19, commented out)
"""
from __future__ import annotations

import numpy as np

W, H = 2048, 1536
_WORLD_ID = 0xFFFFFFFF


def _rng(seed: int, image_id: int) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[int(seed) & 0xFFFFFFFFFFFFFFFF, int(image_id)]))


class World:
    """Landmarks shared by all images of one synthetic set."""

    def __init__(self, kind: str, n_kp: int, seed: int = 0xB200, n_landmarks: int | None = None):
        assert kind in ("sift", "superpoint", "orb")
        self.kind, self.n_kp, self.seed = kind, int(n_kp), int(seed)
        self.L = int(n_landmarks) if n_landmarks else 4 * self.n_kp
        g = _rng(seed, _WORLD_ID)
        # 3-D points inside a ball of radius 0.9
        p = g.standard_normal((self.L, 3))
        p /= np.linalg.norm(p, axis=1, keepdims=True)
        p *= 0.9 * g.random((self.L, 1)) ** (1.0 / 3.0)
        self.xyz = p
        if kind == "sift":
            b = np.abs(g.standard_normal((self.L, 128)))
            b /= np.linalg.norm(b, axis=1, keepdims=True)
            b = np.minimum(b, 0.2)
            b /= np.linalg.norm(b, axis=1, keepdims=True)
            self.base = np.clip(np.floor(512.0 * b), 0, 255).astype(np.float32)
        elif kind == "superpoint":
            b = g.standard_normal((self.L, 256))
            b /= np.linalg.norm(b, axis=1, keepdims=True)
            self.base = b.astype(np.float32)
        else:
            self.base = g.integers(0, 2, size=(self.L, 256), dtype=np.uint8)

    def image(self, image_id: int, n_images_on_ring: int = 100, outlier_frac: float = 0.0):
        """Returns (desc, xy_int32 [N,2], landmark_ids [N])."""
        g = _rng(self.seed, image_id)
        N = self.n_kp
        ids = g.permutation(self.L)[:N]
        # camera on a ring of radius 4 looking at the origin
        ang = 2.0 * np.pi * (image_id % max(n_images_on_ring, 1)) / max(n_images_on_ring, 1)
        ang += 0.05 * g.standard_normal()
        c = np.array([4.0 * np.cos(ang), 0.3 * np.sin(3 * ang), 4.0 * np.sin(ang)])
        z = -c / np.linalg.norm(c)
        x = np.cross(np.array([0.0, 1.0, 0.0]), z); x /= np.linalg.norm(x)
        y = np.cross(z, x)
        R = np.stack([x, y, z])                       # world -> camera
        pc = (self.xyz[ids] - c) @ R.T
        f = 1.2 * max(W, H)                           # Camera.h:45-54
        u = f * pc[:, 0] / pc[:, 2] + W / 2 + 0.5 * g.standard_normal(N)
        v = f * pc[:, 1] / pc[:, 2] + H / 2 + 0.5 * g.standard_normal(N)
        if outlier_frac > 0:
            bad = g.random(N) < outlier_frac
            u = np.where(bad, g.random(N) * W, u)
            v = np.where(bad, g.random(N) * H, v)
        xy = np.stack([np.trunc(u), np.trunc(v)], axis=1).astype(np.int32)
        b = self.base[ids]
        if self.kind == "sift":
            d = np.clip(np.rint(b + 6.0 * g.standard_normal(b.shape, dtype=np.float32)), 0, 255)
            desc = d.astype(np.float32)
        elif self.kind == "superpoint":
            d = b + np.float32(0.35 / 16.0) * g.standard_normal(b.shape, dtype=np.float32)
            d /= np.linalg.norm(d, axis=1, keepdims=True)
            desc = d.astype(np.float32)
        else:
            flip = (g.random(b.shape, dtype=np.float32) < 0.05).astype(np.uint8)
            desc = np.packbits(b ^ flip, axis=1, bitorder="little")
        return np.ascontiguousarray(desc), xy, ids


def make_set(kind: str, n_images: int, n_kp: int, seed: int = 0xB200, outlier_frac: float = 0.0,
             image_ids=None):
    """Convenience: list of (desc, xy) for image ids 0..n_images-1 (or the given ids)."""
    w = World(kind, n_kp, seed)
    ids = range(n_images) if image_ids is None else image_ids
    out = []
    for i in ids:
        d, xy, _ = w.image(i, n_images_on_ring=n_images, outlier_frac=outlier_frac)
        out.append((d, xy))
    return out


def all_pairs(n_images: int) -> np.ndarray:
    """Canonical pair list: query = lower image id; int32 [P,2]."""
    i, j = np.triu_indices(n_images, k=1)
    return np.stack([i, j], axis=1).astype(np.int32)
