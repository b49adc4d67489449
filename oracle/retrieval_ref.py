"""CPU restatement (numpy, float64) of the pair pre-selection of reconstructor_b200/csrc/retrieval.cu.
TEST INFRASTRUCTURE ONLY.

The reference has only FakeImgMatcher (every image with every other one, Mapper/libMapper/ImageMatcher.cpp:6-24) and a
todo "image matcher (apply some image retrieval ...)" (README.md:40); the rule restated here is this repo's own:
global descriptor = unit-length sum of the unit-length local descriptors (binary rows: of their +-1 bit vectors),
cosine similarity, per image the top_k most similar other images (ties to the lower index), canonical pairs (i < j) =
union of both directions, ordered like the all-pairs list (every image against all earlier ones).
With top_k >= n - 1 it is FakeImgMatcher's list."""
from __future__ import annotations

import numpy as np


def global_descriptor(desc: np.ndarray) -> np.ndarray:
    if desc.dtype == np.uint8:                       # binary rows: bit b of a row = bit (b % 8) of byte b // 8
        bits = np.unpackbits(desc, axis=1, bitorder="little").astype(np.float64)
        g = 2.0 * bits.sum(axis=0) - desc.shape[0]
    else:
        d = desc.astype(np.float64)
        n = np.sqrt((d * d).sum(axis=1, keepdims=True))
        g = (d * np.where(n > 0, 1.0 / np.where(n > 0, n, 1.0), 0.0)).sum(axis=0)
    nn = np.sqrt((g * g).sum())
    return g / nn if nn > 0 else g * 0.0


def similarity(descs) -> np.ndarray:
    G = np.stack([global_descriptor(d) for d in descs])
    return G @ G.T


def select_pairs(descs, top_k: int) -> np.ndarray:
    n = len(descs)
    if top_k <= 0 or top_k >= n - 1:
        j, i = np.tril_indices(n, k=-1)
        return np.stack([i, j], axis=1).astype(np.int32)
    S = similarity(descs)
    pairs = set()
    for a in range(n):
        s = S[a].copy(); s[a] = -np.inf
        order = sorted(range(n), key=lambda b: (-s[b], b))[:top_k]
        for b in order:
            if b != a:
                pairs.add((max(a, b), min(a, b)))
    out = sorted(pairs)
    return np.array([(i, j) for (j, i) in out], np.int32).reshape(-1, 2)
