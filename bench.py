#!/usr/bin/env python
"""bench.py -- image pairs/sec of the exhaustive pair-matching path (kNN k=2 + ratio + uniqueness +
F-matrix RANSAC) on synthetic descriptor sets of BASELINE.json's shape.

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one process per GPU)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path (cv2 FLANN +
                                                           findFundamentalMat) on the host cores

A step = one pass of the whole pair loop over the workload:
  N=1   configs[1]: 100 images x 8192 SIFT keypoints (128-d), all 4,950 pairs.
  N>1   weak scaling: the image count grows so that every GPU keeps ~4,950 pairs; the pair list is
        split into contiguous shares, descriptors are replicated, no data-path collective.
`value`  = pairs/s with descriptors resident in HBM, results delivered to host memory (CSR);
`e2e`    = the same through the C ABI with HOST buffers: per step every image is re-ingested from
           pinned host memory (H2D + pack) and the CSR result is read back (D2H).
Timing: CUDA events inside the library on its own streams (pm_csr_result.device_ms), max over ranks;
the descriptor working set (100 x 4.7 MB fp16 operand forms + 400 MB fp32) exceeds the 126 MB L2.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from reconstructor_b200 import shard, synth  # noqa: E402

METRIC = "image pairs/sec (exhaustive kNN match + epipolar RANSAC)"
UNIT = "pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kind", default="sift", choices=["sift", "orb", "superpoint"])
    ap.add_argument("--images", type=int, default=100)
    ap.add_argument("--kp", type=int, default=8192)
    ap.add_argument("--outlier-frac", type=float, default=0.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sync-ingest", action="store_true", help="e2e leg: pm_set_image instead of pm_set_image_async")
    ap.add_argument("--no-stages", action="store_true", help="skip the matcher-only / RANSAC attribution pass")
    ap.add_argument("--replicated-ingest", action="store_true", help="N > 1: every rank uploads every image from its own "
                    "host copy instead of the sharded upload + all-gather")
    ap.add_argument("--sharded-ingest", action="store_true",
                    help="N>1: every rank owns images k = rank mod N, descriptors are all-gathered over NCCL and "
                         "ingested from device memory (the exchange step of sharded extraction, SURVEY 8e)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--debug-flags", type=int, default=0, help="pm_params.debug_flags (kernel variants)")
    ap.add_argument("--batch-pairs", type=int, default=0)
    ap.add_argument("--max-pairs", type=int, default=0, help="time a seeded random sample of this many pairs of the all-pairs "
                    "list instead of all of them (config #5: thousands of images, SURVEY 8d)")
    ap.add_argument("--dev-ratio", type=float, default=None, help="development: Lowe ratio of the main arm (a tiny value "
                    "empties the tail kernels and isolates the kNN kernel)")
    ap.add_argument("--dev-no-filter", action="store_true", help="development: main arm without the epipolar filter")
    return ap.parse_args()


def workload_name(a, n_img, n_pairs):
    dim = {"sift": "128-d float", "orb": "256-bit binary", "superpoint": "256-d float"}[a.kind]
    return (f"{n_img} images {a.kind.upper()} {a.kp} kp ({dim}), exhaustive {n_pairs} pairs, "
            f"kNN k=2 + ratio 0.7 + first-wins uniqueness + F-RANSAC")


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own per-pair body (FLANN knnMatch + ratio/unique + findFundamentalMat),
# one worker process per host core like its OpenMP-over-pairs loop.  Executes oracle/ -- allowed
# here only (cpu_baseline and --impl reference).
# ------------------------------------------------------------------------------------------------
_CPU_IMGS = None


def _cpu_pair(ij):
    from oracle import cv2_ref
    i, j = ij
    (d1, x1), (d2, x2) = _CPU_IMGS[i], _CPU_IMGS[j]
    if d1.dtype == np.uint8:       # what the reference does with ORB bytes: convertTo(CV_32F) + FLANN L2
        d1 = d1.astype(np.float32)  # (FeatureDetector.cpp:24; 32-d float vectors)
        d2 = d2.astype(np.float32)
    return cv2_ref.time_pair_body(d1, x1, d2, x2, matcher="flann")


def _c_oracle_pair(ij):
    from oracle import orc
    i, j = ij
    (d1, x1), (d2, x2) = _CPU_IMGS[i], _CPU_IMGS[j]
    return len(orc.match_pair(d1, x1, d2, x2)["q"])


def _flann_sample(ij):
    """FLANN top-1 indices and the reference body's final match set for one pair (recall report)."""
    from oracle import cv2_ref
    i, j = ij
    (d1, x1), (d2, x2) = _CPU_IMGS[i], _CPU_IMGS[j]
    f1, f2 = d1.astype(np.float32), d2.astype(np.float32)
    idx, _ = cv2_ref.knn2_flann(f1, f2)
    r = cv2_ref.match_pair(f1, x1, f2, x2, matcher="flann")
    return idx[:, 0].copy(), np.stack([r["q"], r["t"]], axis=1)


def flann_samples(imgs, sample_pairs):
    """Runs before any CUDA context exists (fork).  Returns {pair: (top1, final (q,t))} or None."""
    import multiprocessing as mp
    global _CPU_IMGS
    from oracle import cv2_ref
    if not cv2_ref.have_cv2():
        return None
    _CPU_IMGS = imgs
    with mp.get_context("fork").Pool(min(len(sample_pairs), 4)) as pool:
        out = pool.map(_flann_sample, sample_pairs, chunksize=1)
    return dict(zip(sample_pairs, out))


def cpu_arm(imgs, pairs, seconds, steps=1, warmup=0, pairs_per_step=None, workers=None):
    """Returns dict(value pairs/s, cores, kind, sample, ms_per_step)."""
    import multiprocessing as mp
    global _CPU_IMGS
    from oracle import cv2_ref
    _CPU_IMGS = imgs
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    use_cv2 = cv2_ref.have_cv2()
    fn = _cpu_pair if use_cv2 else _c_oracle_pair
    if workers is None or not use_cv2:
        workers = cores if use_cv2 else 1          # the C oracle is already OpenMP-parallel inside
    rng = np.random.default_rng(1)
    n_step = pairs_per_step or max(2 * workers, 16)
    ctx = mp.get_context("fork")
    times = []
    done_pairs = 0
    with ctx.Pool(workers) as pool:
        pool.map(fn, [tuple(pairs[k]) for k in rng.choice(len(pairs), min(workers, len(pairs)), replace=False)])
        k = 0
        t_start = time.perf_counter()
        while True:
            sel = [tuple(pairs[x]) for x in rng.choice(len(pairs), min(n_step, len(pairs)), replace=False)]
            t0 = time.perf_counter()
            pool.map(fn, sel, chunksize=1)
            dt = time.perf_counter() - t0
            if k >= warmup:
                times.append(dt); done_pairs += len(sel)
            k += 1
            if steps > 1 or warmup > 0:
                if k >= steps + warmup:
                    break
            elif time.perf_counter() - t_start >= seconds:
                break
    tot = sum(times)
    return dict(value=done_pairs / tot, unit=UNIT, cores=workers,
                kind="port",
                sample=(f"{done_pairs} random pairs of the workload in {len(times)} steps of {n_step}, "
                        + ("cv2 %s FLANN knnMatch(k=2)+ratio+unique+findFundamentalMat, one process per core"
                           % cv2_ref.cv2.__version__ if use_cv2 else
                           "oracle/pm_oracle.c exact brute force + RANSAC (cv2 not importable), OpenMP")),
                ms_per_step=1e3 * tot / max(len(times), 1), pairs_per_step=n_step)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.p = gpu_index, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.p:
            self.p.terminate()
            try:
                self.p.wait(timeout=2)
            except Exception:
                pass
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def main():
    a = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(a.gpus, 1)

    n_img = a.images if n_gpus == 1 else shard.images_for_world(n_gpus, a.images)
    pairs = shard.all_pairs(n_img)
    n_all = len(pairs)
    if a.max_pairs and a.max_pairs < n_all:
        # pairs are iid here, so a random sample of the list is a steady-state sample of the whole job; sorted so that
        # the list keeps its image-by-image order
        pairs = pairs[np.sort(np.random.default_rng(0xB200).choice(n_all, a.max_pairs, replace=False))]
    cfg = dict(workload=workload_name(a, n_img, n_all), images=n_img, keypoints=a.kp, kind=a.kind,
               pairs=int(len(pairs)), partition=f"pairs sharded over {n_gpus} GPU(s), descriptors replicated",
               l2="inputs larger than L2 (no flush needed)", outlier_frac=a.outlier_frac,
               debug_flags=a.debug_flags)
    if len(pairs) < n_all:
        cfg["pairs_timed"] = int(len(pairs))
        cfg["sample"] = f"seeded random sample of {len(pairs)} of the {n_all} pairs per step"

    if a.impl == "reference":
        if rank != 0:
            return
        w = synth.World(a.kind, a.kp, seed=0xB200 + 2)
        need = sorted(set(np.random.default_rng(1).choice(n_img, min(n_img, 24), replace=False).tolist()))
        imgs = {i: w.image(i, n_img, a.outlier_frac)[:2] for i in need}
        sub = np.array([(i, j) for x, i in enumerate(need) for j in need[x + 1:]], np.int32)
        r = cpu_arm(imgs, sub, a.cpu_seconds, steps=a.steps, warmup=a.warmup)
        line = dict(metric=METRIC, value=r["value"], unit=UNIT, n_gpus=n_gpus, steps=a.steps, warmup=a.warmup,
                    ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f32", data="synthetic", impl="reference", config=cfg,
                    cpu_baseline=dict(value=r["value"], unit=UNIT, cores=r["cores"], kind=r["kind"], sample=r["sample"]),
                    e2e=dict(value=r["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return

    # ---- CPU baseline first (fork before any CUDA context exists), rank 0 at N=1 only -----------
    cpu = None
    w = synth.World(a.kind, a.kp, seed=0xB200 + 2)
    # N > 1: every rank uploads only its own share of the images and the rest arrives by one NCCL all-gather over
    # NVLink (the exchange step of sharded extraction) -- replicating the upload would push the whole
    # descriptor set through every rank's PCIe link each step (measured at N = 8: 24 ms of a 73 ms e2e step)
    sharded = world > 1 and not a.replicated_ingest
    n_local = -(-n_img // world)                       # images per rank when extraction is sharded
    own_ids = list(range(rank, n_local * world, world)) if sharded else list(range(n_img))
    imgs = [w.image(i, n_img, a.outlier_frac)[:2] for i in own_ids]
    if n_gpus == 1 and rank == 0 and not a.no_cpu_baseline:
        sub_ids = list(range(min(n_img, 24)))
        sub = np.array([(i, j) for x, i in enumerate(sub_ids) for j in sub_ids[x + 1:]], np.int32)
        r = cpu_arm({i: imgs[i] for i in sub_ids}, sub, a.cpu_seconds)
        cpu = dict(value=r["value"], unit=UNIT, cores=r["cores"], kind=r["kind"], sample=r["sample"])
        if r["cores"] > 4:      # the reference's own thread cap (MAX_NUM_THREADS 4, SequentialReconstructor.h:17)
            r4 = cpu_arm({i: imgs[i] for i in sub_ids}, sub, min(a.cpu_seconds, 6.0), workers=4)
            cpu["at_4_workers"] = dict(value=r4["value"], unit=UNIT, cores=4, sample=r4["sample"])
        flann = flann_samples({i: imgs[i] for i in sub_ids}, [tuple(int(x) for x in sub[k]) for k in (0, 1, len(sub) // 2, len(sub) - 1)][:len(sub)])
    else:
        flann = None

    import torch
    import torch.distributed as dist
    from reconstructor_b200 import api

    torch.cuda.set_device(local)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    mine = shard.shard_pairs(pairs, rank, world)
    # pinned host copies of every image (what the reference-facing call is handed: float rows + int xy)
    pinned = []
    for d, xy in imgs:
        td = torch.from_numpy(np.ascontiguousarray(d)).pin_memory()
        tx = torch.from_numpy(np.ascontiguousarray(xy)).pin_memory()
        pinned.append((td, tx))
    dim = imgs[0][0].shape[1] * (8 if a.kind == "orb" else 1)
    dt = api.DESC_U8_BITS if a.kind == "orb" else api.DESC_F32

    dev_kw = {}
    if a.dev_ratio is not None: dev_kw["ratio"] = a.dev_ratio
    if a.dev_no_filter: dev_kw["do_filter"] = 0
    pm = api.PairMatcher(devices=[local], reserve_keypoints=n_img * a.kp, debug_flags=a.debug_flags,
                         batch_pairs=a.batch_pairs, **dev_kw)

    if sharded:
        # sharded extraction: this rank "extracted" images rank, rank + N, ...; descriptors and keypoints are
        # all-gathered over NCCL (equal counts per image) and ingested from device memory
        d_own = torch.stack([td for td, _ in pinned]).pin_memory()
        x_own = torch.stack([tx for _, tx in pinned]).pin_memory()
        cfg["partition"] += "; extraction sharded by image id mod N, one NCCL all-gather of descriptors + keypoints per step"

    def ingest(asynchronous=False):
        if not sharded:
            # asynchronous: pm_set_image_async from pinned buffers -- the uploads are queued and the first batches
            # of pm_match_all_pairs run while the later images are still on their way (the pair list is ordered)
            if asynchronous:
                pm.set_images_ptr_async(list(range(len(pinned))), [td.data_ptr() for td, _ in pinned],
                                        [td.shape[0] for td, _ in pinned], dim, dt, [tx.data_ptr() for _, tx in pinned])
                return
            for i, (td, tx) in enumerate(pinned):
                pm.set_image_ptr(i, td.data_ptr(), td.shape[0], dim, dt, tx.data_ptr())
            return
        d_dev = d_own.cuda(non_blocking=True); x_dev = x_own.cuda(non_blocking=True)
        all_d = torch.empty((world,) + tuple(d_dev.shape), dtype=d_dev.dtype, device="cuda")
        all_x = torch.empty((world,) + tuple(x_dev.shape), dtype=x_dev.dtype, device="cuda")
        dist.all_gather_into_tensor(all_d, d_dev)
        dist.all_gather_into_tensor(all_x, x_dev)
        torch.cuda.synchronize()
        for slot in range(n_local):
            for r in range(world):
                img = slot * world + r
                if img < n_img:
                    pm.set_image_ptr(img, all_d[r, slot].data_ptr(), a.kp, dim, dt, all_x[r, slot].data_ptr(), on_device=True,
                                     asynchronous=asynchronous)
        if asynchronous:
            pm.sync_images()            # all_d / all_x go out of scope on return

    ingest()
    for _ in range(a.warmup):
        r = pm.match_all_pairs(mine, copy=False); pm.free_result(r)

    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    pm.reset_stats()
    if sampler:
        sampler.start()
    dev_ms = 0.0
    t0 = time.perf_counter()
    matches = inliers = 0
    for _ in range(a.steps):
        r = pm.match_all_pairs(mine, copy=False)
        dev_ms += r["device_ms"]
        matches = int(r["offsets"][-1]); inliers = int(r["n_inliers"].sum())
        pm.free_result(r)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    clocks = sampler.stop() if sampler else None
    st = pm.stats()
    dev_ms = allmax(dev_ms)
    wall_ms = allmax(wall_ms)
    launches = allsum(st["kernel_launches"])
    total_pairs = len(pairs)
    ms_per_step = dev_ms / a.steps
    value = total_pairs / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (kNN), rank 0 --------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    knn_s = st["knn_ms"] * 1e-3
    orb_form = "popc" if (a.debug_flags & 1024) else "e4m3" if (a.debug_flags & 65536) else "i8" if (a.debug_flags & 262144) else "fp4"
    if a.kind == "orb" and orb_form != "popc":
        # default binary path: Hamming = |a| + |b| - 2 a.b on the tensor cores (E2M1 {0,+-1} operands on kind::mxf4, exact);
        # one popc32 of the fixed numerator = 32 bit compares = 64 FLOP of the contraction
        pk = peaks.get("bf16_tflops_sustained")
        popc_peak = pm.measure_popc_peak()
        mult = 4.0 if orb_form == "fp4" else 2.0
        roof = dict(bound="tensor", achieved=64.0 * st["knn_work"] / knn_s / 1e12, peak=mult * (pk if pk else 1400.0),
                    unit="TFLOP/s",
                    peak_source=(("%g x MEASURED_PEAKS.json bf16_tflops_sustained (%s; no measured figure of that kind in "
                                  "MEASURED_PEAKS.json)" % (mult, "fp4 dense (kind::mxf4) = 4 x bf16 on B200" if orb_form == "fp4" else
                                                            "fp8 / int8 dense = 2 x bf16 on B200")) if pk
                                 else "%g x fallback 1.4 PFLOP/s (B200_PROFILING.md)" % mult),
                    operands={"fp4": "E2M1 {0, +-1} values on tcgen05 kind::mxf4 (64 values of K per instruction, all-ones scale "
                                     "factors), f32 accumulate, exact integers", "i8": "u8 x s8 on tcgen05 kind::i8, s32 accumulate",
                              "e4m3": "E4M3 {0,1} values on tcgen05 kind::f8f6f4, f32 accumulate, exact integers"}[orb_form],
                    frac_of_bf16_peak=64.0 * st["knn_work"] / knn_s / 1e12 / (pk if pk else 1400.0),
                    popc32_equiv=dict(achieved=st["knn_work"] / knn_s / 1e12, popc_pipe_peak=popc_peak / 1e12, unit="Tpopc32/s",
                                      note="the XOR/popc kernel of the north star (--debug-flags 1024) is bounded by popc_pipe_peak"))
    elif a.kind == "orb":
        popc_peak = pm.measure_popc_peak()
        roof = dict(bound="popc", achieved=st["knn_work"] / knn_s / 1e12, peak=popc_peak / 1e12, unit="Tpopc32/s",
                    peak_source="measured live: pm_measure_popc_peak (dependent-chain POPC micro-benchmark)")
    else:
        pk = peaks.get("bf16_tflops_sustained")
        roof = dict(bound="tensor", achieved=st["knn_work"] / knn_s / 1e12, peak=pk if pk else 1400.0, unit="TFLOP/s",
                    peak_source="MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if pk
                    else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)")
    roof["frac"] = roof["achieved"] / roof["peak"]
    if a.kind == "sift" and not (a.debug_flags & 2048):
        roof["operands"] = ("int8 (tcgen05 kind::i8, u8 x s8 -> s32): the MMA rate of this kind is 2 x the bf16 rate the peak "
                            "is quoted in, so frac can reach 2; ncu: tensor pipe (imma) active 90.4 % of elapsed "
                            "(profiles/r01_i8x2_kernel_ncu_full.md)")
        roof["frac_of_int8_peak"] = roof["frac"] / 2.0
    if a.kind == "superpoint" and not (a.debug_flags & 32768):
        roof["operands"] = ("int8 (tcgen05 kind::i8, rows quantised to s8 for the candidate stage; exact fp32 re-rank follows): "
                            "2 x the bf16 MMA rate the peak is quoted in")
        roof["frac_of_int8_peak"] = roof["frac"] / 2.0
    roof["kernel"] = {"sift": "l2_top2_tc2_kernel" if (a.debug_flags & (2048 | 16384)) else "l2_i8x2_kernel", "orb": {"popc": "hamming_top2_kernel", "e4m3": "l2_top2_tc2_kernel<T2Cfg<256,2,2>,2,false,1>", "i8": "l2_i8x2_kernel<2,false,2,0>", "fp4": "l2_i8x2_kernel<2,false,1,1>"}[orb_form], "superpoint": "l2_top2_tc2_kernel<T2Cfg<256,2,4>,3>" if (a.debug_flags & 32768) else "l2_top2_tc2_kernel<T2Cfg<256,2,2>,3,true,3>"}[a.kind]
    roof["launches"] = st["knn_launches"]
    roof["avg_launch_ms"] = st["knn_ms"] / max(st["knn_launches"], 1)
    roof["share_of_step"] = st["knn_ms"] / max(dev_ms, 1e-9) if world == 1 else None
    if a.kind == "superpoint":
        roof["rerank"] = {k[7:]: st[k] for k in st if k.startswith("rerank_")}
    # DRAM bytes per launch of that kernel from the committed ncu --set full capture, scaled to this
    # run's pairs per launch (null when no capture exists for the kernel)
    roof["traffic"] = None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(roof["kernel"])
        if t:
            pairs_per_launch = len(mine) * a.steps / max(st["knn_launches"], 1)
            roof["traffic"] = t["bytes_per_launch"] * pairs_per_launch / t["pairs_per_launch"]
            roof["traffic_source"] = t["source"]
    except Exception:
        pass

    # ---- attribution: the same workload with the epipolar filter switched off (matcher only); the filter's cost
    #      is the difference ("matcher-only and RANSAC-only pairs/s") -----------------------------
    stages = None
    if not a.no_stages and world == 1:
        pm2 = api.PairMatcher(devices=[local], reserve_keypoints=n_img * a.kp, debug_flags=a.debug_flags,
                              batch_pairs=a.batch_pairs, do_filter=0)
        for i, (td, tx) in enumerate(pinned):
            pm2.set_image_ptr(i, td.data_ptr(), td.shape[0], dim, dt, tx.data_ptr())
        r = pm2.match_all_pairs(mine, copy=False); pm2.free_result(r)
        torch.cuda.synchronize()
        m_ms = 0.0
        for _ in range(a.steps):
            r = pm2.match_all_pairs(mine, copy=False)
            m_ms += r["device_ms"]
            pm2.free_result(r)
        pm2.close()
        m_ms /= a.steps
        f_ms = max(ms_per_step - m_ms, 0.0)
        stages = dict(matcher_only=dict(value=total_pairs / (m_ms * 1e-3), unit=UNIT, ms_per_step=m_ms,
                                        what="kNN + ratio + uniqueness + CSR, do_filter = 0"),
                      epipolar_filter=dict(ms_per_step=f_ms, value=(total_pairs / (f_ms * 1e-3)) if f_ms > 0 else None,
                                           unit=UNIT, derived="full step minus matcher-only step",
                                           share_of_step=f_ms / ms_per_step))

    # ---- recall of the reference's approximate FLANN search against this exact search ------
    if cpu is not None and flann:
        agree = tot = inter = union = 0
        for (i, j), (top1, final) in flann.items():
            gi, _ = pm.knn_pair(i, j)
            agree += int((gi[:, 0] == top1).sum()); tot += len(top1)
            g = pm.match_filter_pair(i, j)
            keep = g["inlier"].astype(bool)
            gs = set(zip(g["q"][keep].tolist(), g["t"][keep].tolist()))
            fs = set(map(tuple, final.tolist()))
            inter += len(gs & fs); union += len(gs | fs)
        cpu["recall"] = dict(pairs=len(flann), flann_top1_agreement=agree / max(tot, 1),
                             final_match_set_iou=inter / max(union, 1),
                             note="reference FLANN (approximate) vs this library's exact search, same pairs")

    # ---- e2e: host buffers in, host CSR out, every step ---------------------------------------------
    e2e = None
    if not a.no_e2e:
        use_async = not a.sync_ingest
        ingest(use_async); r = pm.match_all_pairs(mine, copy=False); pm.free_result(r)      # warm
        barrier()
        pm.reset_stats()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            ingest(use_async)
            r = pm.match_all_pairs(mine, copy=False)
            _ = int(r["n_inliers"].sum())                     # the step's result is read on the host
            pm.free_result(r)
        barrier()
        e_ms = allmax(1e3 * (time.perf_counter() - t0)) / a.steps
        st2 = pm.stats()
        e2e = dict(value=total_pairs / (e_ms * 1e-3), unit=UNIT, ms_per_step=e_ms,
                   h2d_bytes_per_step=int(allsum(st2["h2d_bytes"]) / a.steps) +
                   (int(allsum(d_own.numel() * d_own.element_size() + x_own.numel() * x_own.element_size())) if sharded else 0),
                   d2h_bytes_per_step=int(allsum(st2["d2h_bytes"]) / a.steps),
                   timing="host wall clock around %s x images + match_all_pairs, max over ranks" %
                          ("set_images_async" if use_async and not sharded else "NCCL all-gather + set_image_device_async" if use_async else "set_image"))
        if sharded:
            e2e["allgather_bytes_per_step"] = int(world * (d_own.numel() * d_own.element_size() + x_own.numel() * x_own.element_size()))

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=n_gpus, steps=a.steps, warmup=a.warmup,
                    ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype={"sift": "f16 operands / f32 accumulate (exact integers)" if (a.debug_flags & 2048) else
                           "u8 x s8 operands / s32 accumulate (kind::i8, exact integers)", "orb": {"popc": "u32 popc", "e4m3": "e4m3 {0,1} operands / f32 accumulate (exact integers)", "i8": "u8 x s8 operands / s32 accumulate (exact integers)", "fp4": "e2m1 {0,+-1} operands (kind::mxf4) / f32 accumulate (exact integers)"}[orb_form],
                           "superpoint": "f16 operands / f32 accumulate candidates + exact f32 re-rank" if (a.debug_flags & 32768) else
                           "s8 operands / s32 accumulate candidates (kind::i8) + exact f32 re-rank"}[a.kind],
                    data="synthetic", config=cfg, wall_ms_per_step=wall_ms / a.steps,
                    putative_matches_per_step=matches, inliers_per_step=inliers,
                    roofline=roof, stages=stages, cpu_baseline=cpu, e2e=e2e, gpu_launches=int(launches), clocks=clocks)
        print(json.dumps(line))
    pm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
