"""Multi-GPU partition of the pair loop: image pairs are independent units, so the
pair list is dealt out in equally sized shares -- one per rank / device -- with every descriptor
set replicated on every GPU and NO data-path collective.  A pair's result does not depend on
which GPU ran it (the RANSAC stream is a function of the pair's own matches only).

Shares are BLOCK-CYCLIC over the image-ordered pair list (blocks of SHARD_BLOCK pairs, rank r takes
blocks r, r + R, ...; the ragged end is split evenly): every rank's share starts with pairs of the
first images, so with asynchronous / collective ingest all ranks start matching while later images
are still arriving.  Contiguous shares made the last rank wait for the whole image set (its first
pair already needs the last third of the images): 7.6 ms of a 43 ms end-to-end step at 2 GPUs."""
from __future__ import annotations

import numpy as np


def images_for_world(world: int, base_images: int = 100) -> int:
    """Weak scaling: smallest image count whose all-pairs list holds >= world * C(base,2) pairs,
    i.e. the per-GPU share stays ~ C(base,2) pairs (4,950 for the 100-image config)."""
    target = world * base_images * (base_images - 1) // 2
    n = base_images
    while n * (n - 1) // 2 < target:
        n += 1
    return n


def all_pairs(n_images: int) -> np.ndarray:
    """Canonical pair list, query = lower image id (SequentialReconstructor.cpp:203-227 run
    sequentially visits (i,j), i<j first and mirrors (j,i)).  Order: every image against all earlier
    ones -- (0,1), (0,2), (1,2), (0,3), ... -- so that the first batches need only the first few images
    (they overlap the asynchronous upload of the rest) and consecutive pairs share their train image."""
    j, i = np.tril_indices(n_images, k=-1)
    return np.stack([i, j], axis=1).astype(np.int32)


SHARD_BLOCK = 64


def shard_bounds(n_pairs: int, world: int) -> np.ndarray:
    """Sizes as prefix sums: shares differ by at most one pair."""
    return (np.arange(world + 1, dtype=np.int64) * n_pairs) // world


def shard_index(n_pairs: int, rank: int, world: int, block: int = SHARD_BLOCK) -> np.ndarray:
    """Indices (ascending) of the pairs of `rank`: whole rounds of `world` blocks are dealt block by block, the
    remaining < world * block pairs are split contiguously and evenly."""
    if world <= 1:
        return np.arange(n_pairs, dtype=np.int64)
    rounds = n_pairs // (block * world)
    head = (np.arange(rounds, dtype=np.int64)[:, None] * (block * world) + rank * block +
            np.arange(block, dtype=np.int64)[None, :]).reshape(-1)
    done = rounds * block * world
    b = shard_bounds(n_pairs - done, world)
    return np.concatenate([head, done + np.arange(b[rank], b[rank + 1], dtype=np.int64)])


def shard_pairs(pairs: np.ndarray, rank: int, world: int) -> np.ndarray:
    return np.ascontiguousarray(pairs[shard_index(len(pairs), rank, world)])
