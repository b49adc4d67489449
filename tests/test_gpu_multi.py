"""Multi-GPU tests (need >= 2 devices; skipped on a single-GPU lease): the collective ingest under torchrun and the
in-process multi-device handle.  `gpurun --gpus 2 -- 'python -m pytest tests/test_gpu_multi.py -m gpu'`."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def test_collective_ingest_matches_replicated_ingest():
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(ROOT, "tests", "mgpu_ingest_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for kind in ("sift", "orb", "superpoint"):
        assert f"MGPU_INGEST_OK {kind} world={world}" in r.stdout


def test_in_process_multi_device_handle():
    """pm_create(..., n_dev > 1): one host thread per device, results identical to the single-device handle."""
    if _n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    import numpy as np
    from reconstructor_b200 import api, synth
    for kind in ("sift", "superpoint"):
        imgs = synth.make_set(kind, 8, 900, seed=4)
        outs = []
        for devs in ([0], list(range(min(_n_gpus(), 4)))):
            with api.PairMatcher(devices=devs, batch_pairs=5) as pm:
                for i, (d, xy) in enumerate(imgs):
                    pm.set_image(i, d, xy)
                outs.append(pm.match_all_pairs())
        for k in ("pair_ij", "offsets", "q", "t", "inlier", "status", "n_inliers", "ransac_iters", "F"):
            assert np.array_equal(outs[0][k], outs[1][k]), (kind, k)
