"""Development aid: random sets of 256-bit binary descriptors -- image sizes 1..2100 around the 96 / 192 / 384-row boundaries of
the packed-pair kind::mxf4 kernel, near-duplicates, all-zero / all-one rows, images replaced by smaller and larger ones and
removed (arena rows reused: stale pair norm blocks) -- matched with the default kernel, the one-row kernel (debug bit 26) and the
XOR/popc kernel (bit 10): the CSR results must be identical."""
import sys
import numpy as np
sys.path.insert(0, ".")
from reconstructor_b200 import api

SIZES = [1, 2, 31, 33, 95, 96, 97, 191, 192, 193, 200, 287, 288, 289, 383, 384, 385, 400, 575, 576, 577, 700, 767, 768, 769,
         1000, 1151, 1152, 1153, 1500, 2100]


def image(rng, base, n):
    d = base[rng.permutation(len(base))[:n]].copy()
    d ^= ((rng.random((n, 32)) < 0.05) * rng.integers(1, 256, (n, 32))).astype(np.uint8)
    if n > 8:
        d[rng.integers(0, n)] = 0
        d[rng.integers(0, n)] = 255
        d[rng.integers(0, n)] = d[0]
    return d


lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0
for seed in range(lo, hi):
    rng = np.random.default_rng(4242 + seed)
    base = rng.integers(0, 256, (2100, 32), dtype=np.uint8)
    mode = int(rng.choice([api.UNIQUE_FIRST_WINS, api.MUTUAL_NN, api.UNIQUE_NONE]))
    n_img = int(rng.integers(3, 7))
    script = [("set", i, int(rng.choice(SIZES))) for i in range(n_img)]
    for _ in range(int(rng.integers(0, 4))):                   # replace / remove / add again
        i = int(rng.integers(0, n_img))
        script.append((str(rng.choice(["set", "set", "remove"])), i, int(rng.choice(SIZES))))
    data = {}
    for k, (op, i, n) in enumerate(script):
        data[k] = image(np.random.default_rng(seed * 100 + k), base, n) if op == "set" else None
    outs = []
    for flags in (0, 1 << 26, 1 << 10):
        with api.PairMatcher(unique_mode=mode, debug_flags=flags, do_filter=0, batch_pairs=int(rng.choice([0, 3, 7]))) as pm:
            live = set()
            for k, (op, i, n) in enumerate(script):
                if op == "set":
                    pm.set_image(i, data[k]); live.add(i)
                elif i in live:
                    pm.remove_image(i); live.discard(i)
            outs.append(pm.match_all_pairs() if len(live) >= 2 else None)
    if outs[0] is None:
        continue
    for o in outs[1:]:
        for key in ("pair_ij", "offsets", "q", "t", "status"):
            if not np.array_equal(outs[0][key], o[key]):
                bad += 1
                print("FAIL seed", seed, "mode", mode, "key", key, "script", script)
                break
print("orb packed fuzz done: seeds %d..%d, failures %d" % (lo, hi, bad))
