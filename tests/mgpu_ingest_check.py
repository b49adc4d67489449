"""Multi-GPU check of the collective ingest (run by tests/test_gpu_multi.py under torchrun, one rank per GPU):
pm_ingest_allgather (extraction sharded by image id mod N, NCCL all-gather in the wire dtype) must leave every rank
with exactly the image set a replicated pm_set_image ingest gives -- same CSR for the rank's share of the pair list --
and the shares of all ranks together must equal the single-GPU result of rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reconstructor_b200 import api, shard, synth  # noqa: E402

KEYS = ("offsets", "q", "t", "inlier", "status", "n_inliers", "ransac_iters", "F")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for kind, n_img, kp in (("sift", 7, 700), ("orb", 6, 512), ("superpoint", 5, 600)):
        imgs = synth.make_set(kind, n_img, kp, seed=31)
        pairs = shard.all_pairs(n_img)
        mine = shard.shard_pairs(pairs, rank, world)
        with api.PairMatcher(devices=[local]) as ref:
            for i, (d, xy) in enumerate(imgs):
                ref.set_image(i, d, xy)
            want = ref.match_all_pairs(mine)
            full = ref.match_all_pairs(pairs) if rank == 0 else None
        own = list(range(rank, n_img, world))
        d_own = torch.from_numpy(np.stack([imgs[i][0] for i in own])).pin_memory()
        x_own = torch.from_numpy(np.stack([imgs[i][1] for i in own])).pin_memory()
        dt = api.DESC_U8_BITS if kind == "orb" else api.DESC_F32
        dim = 256 if kind == "orb" else imgs[0][0].shape[1]
        for wire in ([api.DESC_U8, api.DESC_F32] if kind == "sift" else [dt]):
            with api.PairMatcher(devices=[local]) as pm:
                uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
                if rank == 0:
                    uid.copy_(torch.frombuffer(bytearray(api.comm_unique_id()), dtype=torch.uint8))
                dist.broadcast(uid, src=0)
                pm.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
                for rep in range(2):                              # twice: the second ingest re-sets every image
                    pm.ingest_allgather(n_img, kp, dim, dt, wire, d_own.data_ptr(), x_own.data_ptr())
                    got = pm.match_all_pairs(mine)                # (no pm_sync_images: the loop resolves what it needs)
                    pm.sync_images()
                    for k in KEYS:
                        assert np.array_equal(got[k], want[k]), (kind, wire, rep, k, rank)
                # the shares of all ranks, concatenated in rank order, are the single-GPU result
                parts = [None] * world
                dist.all_gather_object(parts, {k: got[k] for k in ("q", "t", "inlier", "n_inliers", "status")})
                if rank == 0:
                    for k in ("q", "t", "inlier", "n_inliers", "status"):
                        assert np.array_equal(np.concatenate([p[k] for p in parts]), full[k]), (kind, k)
        if kind == "sift":                                        # the u8 wire promise is checked on the device
            bad = d_own.clone(); bad[0, 3, 5] = 0.5
            bad = bad.pin_memory()
            with api.PairMatcher(devices=[local]) as pm:
                uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
                if rank == 0:
                    uid.copy_(torch.frombuffer(bytearray(api.comm_unique_id()), dtype=torch.uint8))
                dist.broadcast(uid, src=0)
                pm.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
                pm.ingest_allgather(n_img, kp, dim, dt, api.DESC_U8, bad.data_ptr(), x_own.data_ptr())
                try:
                    pm.sync_images()
                    raise SystemExit("non-integral rows were accepted on the u8 wire")
                except api.PairMatchError as e:
                    assert e.code == api.ERR_INVALID
        if rank == 0:
            print(f"MGPU_INGEST_OK {kind} world={world}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
