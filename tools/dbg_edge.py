import sys
import numpy as np
sys.path.insert(0, ".")
from reconstructor_b200 import api
rng = np.random.default_rng(31)
sizes = (1, 2, 31, 32, 33, 255, 256, 257, 511, 512, 513, 769, 0)
base = rng.integers(0, 60, (1200, 128)).astype(np.float32)
noise = lambda n: rng.integers(-3, 4, (n, 128))
make = lambda ids: np.clip(base[ids] + noise(len(ids)), 0, 255).astype(np.float32)
imgs = [make(rng.permutation(1200)[:n]) if n else np.zeros((0, 128), np.float32) for n in sizes]
outs = []
for flags in (0, 1, 2048, 16384):
    with api.PairMatcher(do_filter=0, batch_pairs=16, debug_flags=flags) as pm:
        for i, d in enumerate(imgs):
            pm.set_image(i, d)
        outs.append(pm.match_all_pairs())
ref = outs[1]
for name, o in zip(("i8x2", "simt", "f16", "i8x1"), outs):
    bad = []
    for p, (i, j) in enumerate(o["pair_ij"]):
        a, b = o["offsets"][p], o["offsets"][p + 1]
        ra, rb = ref["offsets"][p], ref["offsets"][p + 1]
        if (b - a) != (rb - ra) or not np.array_equal(o["q"][a:b], ref["q"][ra:rb]) or not np.array_equal(o["t"][a:b], ref["t"][ra:rb]):
            bad.append((int(i), int(j), sizes[i], sizes[j], int(b - a), int(rb - ra)))
    print(name, "bad pairs:", bad[:12], len(bad))
