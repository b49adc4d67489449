"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, the
header is plain C, there is no CPU fallback, and the host-side sharding logic (incl. a 2-rank gloo
run) covers the pair list exactly once."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from reconstructor_b200 import api, shard, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pairmatch_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pm_[a-zA-Z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = api.load_library()
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(api.EXPORTS)
    assert b"sm_100a" in lib.pm_version()


def test_header_is_plain_c(tmp_path):
    c = tmp_path / "t.c"
    c.write_text('#include "pairmatch_b200.h"\nint main(void){pm_params p; pm_default_params(&p); '
                 'return p.min_matches == 7 ? 0 : 1;}\n')
    exe = tmp_path / "t"
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.dirname(HEADER), str(c),
                           "-o", str(exe), api.LIB_PATH, "-Wl,-rpath," + os.path.dirname(api.LIB_PATH)])
    assert subprocess.call([str(exe)]) == 0


def test_default_params_follow_the_reference():
    p = api.default_params()
    assert abs(p.ratio - 0.7) < 1e-7                 # FeatureMatcher.h:45
    assert p.min_matches == 7 and p.do_filter == 1   # SequentialReconstructor.cpp:237
    assert (p.ransac_threshold, p.ransac_confidence, p.ransac_max_iters) == (3.0, 0.99, 1000)
    assert p.unique_mode == api.UNIQUE_FIRST_WINS and p.residual_mode == api.RESID_SYMMETRIC_EPIPOLAR


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.PairMatchError) as e:
        api.PairMatcher()
    assert e.value.code == api.ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "reconstructor_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "pm_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_feature_matches_view_mirrors_and_drops():
    res = dict(n_pairs=3, pair_ij=np.array([[0, 1], [0, 2], [1, 2]], np.int32),
               offsets=np.array([0, 3, 5, 5], np.int64), q=np.array([1, 4, 9, 2, 3], np.int32),
               t=np.array([7, 0, 5, 8, 6], np.int32), inlier=np.array([1, 0, 1, 1, 1], np.uint8),
               status=np.array([api.PAIR_FILTERED, api.PAIR_DROPPED, api.PAIR_UNFILTERED], np.int32))
    fm = api.feature_matches_view(res)
    assert fm[(0, 1)] == {1: 7, 9: 5} and fm[(1, 0)] == {7: 1, 5: 9}
    assert (0, 2) not in fm and (1, 2) not in fm


def test_synth_is_deterministic_and_in_domain():
    for kind in ("sift", "orb", "superpoint"):
        a = synth.World(kind, 256, seed=3).image(5, 10)
        b = synth.World(kind, 256, seed=3).image(5, 10)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    d = synth.World("sift", 256, seed=3).image(0, 10)[0]
    assert d.dtype == np.float32 and d.shape == (256, 128) and (d == np.rint(d)).all() and d.min() >= 0 and d.max() <= 255
    d = synth.World("superpoint", 256, seed=3).image(0, 10)[0]
    np.testing.assert_allclose(np.linalg.norm(d, axis=1), 1.0, rtol=1e-5)
    d = synth.World("orb", 256, seed=3).image(0, 10)[0]
    assert d.dtype == np.uint8 and d.shape == (256, 32)


def test_weak_scaling_image_counts():
    assert shard.images_for_world(1) == 100
    for w in (2, 4, 8):
        n = shard.images_for_world(w)
        per = n * (n - 1) // 2 / w
        assert 4950 <= per < 4950 * 1.03


def test_shards_cover_pairs_once():
    pairs = shard.all_pairs(37)
    assert len(pairs) == 37 * 36 // 2 and (pairs[:, 0] < pairs[:, 1]).all()
    for world in (1, 2, 3, 8):
        parts = [shard.shard_pairs(pairs, r, world) for r in range(world)]
        idx = [shard.shard_index(len(pairs), r, world) for r in range(world)]
        assert np.array_equal(np.sort(np.concatenate(idx)), np.arange(len(pairs)))      # every pair exactly once
        assert all((np.diff(i) > 0).all() for i in idx if len(i) > 1)                   # image order kept inside a share
        assert all(np.array_equal(pairs[i], p) for i, p in zip(idx, parts))
        sizes = [len(p) for p in parts]
        assert max(sizes) - min(sizes) <= 1
    # block-cyclic: every rank's share starts among the first pairs (early images), for the overlap with the ingest
    big = shard.all_pairs(142)
    for r in range(8):
        assert shard.shard_index(len(big), r, 8)[0] == r * shard.SHARD_BLOCK
        assert shard.shard_pairs(big, r, 8)[0, 1] <= 33


_GLOO_WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
from reconstructor_b200 import shard
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = shard.images_for_world(world, 20)
pairs = shard.all_pairs(n)
mine = shard.shard_pairs(pairs, rank, world)
# every rank processes its share (here: a checksum of the pair ids stands in for the device work)
local = torch.tensor([len(mine), int((mine[:, 0].astype(np.int64) * 100003 + mine[:, 1]).sum())], dtype=torch.int64)
allv = [torch.zeros_like(local) for _ in range(world)]
dist.all_gather(allv, local)
tot = sum(int(v[0]) for v in allv); chk = sum(int(v[1]) for v in allv)
want = int((pairs[:, 0].astype(np.int64) * 100003 + pairs[:, 1]).sum())
t = torch.tensor([float(rank + 1)], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)       # the bench takes the max over ranks of the device time
assert tot == len(pairs) and chk == want and t.item() == world, (tot, len(pairs), chk, want)
dist.barrier()
if rank == 0: print("GLOO_OK", world, n, tot)
"""


def test_two_rank_gloo_partition(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "GLOO_OK 2" in r.stdout


# ---------------------------------------------------------------------------------------------
# on-disk cache: the C ABI reader/writer against the numpy mirror (no device needed for results)
# ---------------------------------------------------------------------------------------------
def _fake_result(rng, n_pairs=7):
    counts = rng.integers(0, 40, n_pairs)
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    m = int(off[-1])
    return dict(n_pairs=n_pairs, device_ms=12.5, pair_ij=rng.integers(0, 9, (n_pairs, 2)).astype(np.int32), offsets=off,
                q=rng.integers(0, 8192, m).astype(np.int32), t=rng.integers(0, 8192, m).astype(np.int32),
                inlier=rng.integers(0, 2, m).astype(np.uint8), F=rng.standard_normal((n_pairs, 3, 3)),
                status=rng.integers(0, 3, n_pairs).astype(np.int32), n_inliers=counts.astype(np.int32),
                ransac_iters=rng.integers(0, 1000, n_pairs).astype(np.int32))


def test_result_cache_roundtrip_c_abi_vs_numpy_mirror(tmp_path):
    import ctypes as C

    from reconstructor_b200 import api, cache
    rng = np.random.default_rng(3)
    for n_pairs in (7, 1, 0):
        want = _fake_result(rng, n_pairs)
        a, b = str(tmp_path / f"a{n_pairs}.pmb"), str(tmp_path / f"b{n_pairs}.pmb")
        cache.write_result(a, want)                         # numpy writer -> C reader
        got = api.load_result(a)
        for k in ("pair_ij", "offsets", "q", "t", "inlier", "F", "status", "n_inliers", "ransac_iters"):
            assert np.array_equal(got[k], want[k]), (n_pairs, k)
        assert got["n_pairs"] == n_pairs and got["device_ms"] == 12.5
        lib = api.load_library()                            # C reader -> C writer: byte-identical file
        res = C.POINTER(api.CsrResult)()
        assert lib.pm_load_result(a.encode(), C.byref(res)) == 0
        assert lib.pm_save_result(res, b.encode()) == 0
        lib.pm_free_result(res)
        assert open(a, "rb").read() == open(b, "rb").read()
        back = cache.read_result(b)                         # C writer -> numpy reader
        assert np.array_equal(back["q"], want["q"]) and np.array_equal(back["F"], want["F"])


def test_cache_rejects_corrupt_files(tmp_path):
    from reconstructor_b200 import api, cache
    rng = np.random.default_rng(4)
    p = str(tmp_path / "r.pmb")
    cache.write_result(p, _fake_result(rng))
    raw = bytearray(open(p, "rb").read())
    for mutate in ("flip", "truncate", "magic", "kind"):
        bad = bytearray(raw)
        if mutate == "flip":
            bad[len(bad) // 2] ^= 0x10
        elif mutate == "truncate":
            bad = bad[:-8]
        elif mutate == "magic":
            bad[0] = ord("X")
        else:
            bad[12] = 1                                     # claims to be an image file
        q = str(tmp_path / f"bad_{mutate}.pmb")
        open(q, "wb").write(bytes(bad))
        with pytest.raises(api.PairMatchError) as e:
            api.load_result(q)
        assert e.value.code == api.ERR_INVALID
        with pytest.raises(ValueError):
            cache.read_result(q)
    with pytest.raises(api.PairMatchError):
        api.load_result(str(tmp_path / "missing.pmb"))


def test_image_cache_numpy_roundtrip(tmp_path):
    from reconstructor_b200 import cache, synth
    imgs = {}
    for kind, i in (("orb", 3), ("orb", 1)):
        d, xy = synth.make_set(kind, 1, 50 + i, seed=i)[0]
        imgs[i] = (d, xy if i == 3 else None)
    p = str(tmp_path / "img.pmb")
    cache.write_images(p, imgs)
    back = cache.read_images(p)
    assert [r["id"] for r in back] == [1, 3] and back[0]["xy"] is None
    assert np.array_equal(back[1]["desc"], imgs[3][0]) and np.array_equal(back[1]["xy"], imgs[3][1])
    assert back[1]["dim"] == 256 and back[1]["dtype"] == cache.DESC_U8_BITS


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the driver's CPU arm) prints one JSON line with the contract's keys; it runs the
    cv2 restatement of the reference body on the host cores and needs no GPU."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--images", "6", "--kp", "512",
                        "--steps", "1", "--warmup", "0", "--cpu-seconds", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]
