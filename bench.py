#!/usr/bin/env python
"""bench.py -- image pairs/sec of the exhaustive pair-matching path (kNN k=2 + ratio + uniqueness +
F-matrix RANSAC) on synthetic descriptor sets of BASELINE.json's shapes.

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one process per GPU)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path (cv2 FLANN +
                                                           findFundamentalMat) on the host cores

A step = one pass of the whole pair loop over a workload.

Headline line (`value`, `e2e`, `roofline`, `stages`, `cpu_baseline`):
  N=1   configs[1]: 100 images x 8192 SIFT keypoints (128-d), all 4,950 pairs.
  N>1   weak scaling of that config: the image count grows so that every GPU keeps ~4,950 pairs.
`configs` (same JSON line) carries the other BASELINE.json configurations, each with value / e2e / roofline / clocks:
  orb500          configs[2]: 500 images x 8192 ORB keypoints (256 bit), 124,750 pairs        -- STRONG scaling over N
  superpoint1000  configs[3]: 1,000 images x 8192 SuperPoint keypoints (256-d), 499,500 pairs -- STRONG scaling over N
`stages.ransac_heavy` is the headline workload with 50 % outlier keypoints (RANSAC runs hundreds of iterations).

The pair list is dealt in equal block-cyclic shares (one per rank; shard.py), descriptors are replicated on every GPU, the timed loop
has no data-path collective.
`value`  = pairs/s with descriptors resident in HBM, results delivered to host memory (CSR);
`e2e`    = the same through the C ABI with HOST buffers: per step every image is re-ingested from pinned host memory
           (N=1: H2D + pack; N>1: extraction sharded by image id mod N -- every rank uploads only its own images and
           pm_ingest_allgather replicates them over NVLink) and the CSR result is read back (D2H).
Timing: CUDA events inside the library on its own streams (pm_csr_result.device_ms), max over ranks; the descriptor
working sets exceed the 126 MB L2 (no flush needed).  Roofline peaks are MEASURED in the same run with the same MMA kind
(pm_measure_tensor_peak); MEASURED_PEAKS.json's bf16 figure is reported next to them.
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
DIM_TXT = {"sift": "128-d float", "orb": "256-bit binary", "superpoint": "256-d float"}


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
    ap.add_argument("--no-stages", action="store_true", help="skip the matcher-only / RANSAC attribution passes")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` object (ORB-500, SuperPoint-1000)")
    ap.add_argument("--orb-images", type=int, default=500, help="configs.orb500 (BASELINE configs[2])")
    ap.add_argument("--sp-images", type=int, default=1000, help="configs.superpoint1000 (BASELINE configs[3])")
    ap.add_argument("--cfg-steps", type=int, default=2, help="timed steps of the `configs` / ransac_heavy workloads")
    ap.add_argument("--replicated-ingest", action="store_true", help="N > 1: every rank uploads every image from its own "
                    "host copy instead of the sharded upload + all-gather")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--debug-flags", type=int, default=0, help="pm_params.debug_flags (kernel variants)")
    ap.add_argument("--batch-pairs", type=int, default=0)
    ap.add_argument("--max-pairs", type=int, default=0, help="time a seeded random sample of this many pairs of the all-pairs "
                    "list instead of all of them (config #5: thousands of images, SURVEY 8d)")
    ap.add_argument("--dev-ratio", type=float, default=None, help="development: Lowe ratio of the main arm (a tiny value "
                    "empties the tail kernels and isolates the kNN kernel)")
    ap.add_argument("--dev-no-filter", action="store_true", help="development: main arm without the epipolar filter")
    return ap.parse_args()


def workload_name(kind, n_img, kp, n_pairs):
    return (f"{n_img} images {kind.upper()} {kp} kp ({DIM_TXT[kind]}), exhaustive {n_pairs} pairs, "
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
                        + ("cv2 %s FLANN knnMatch(k=2)+ratio+unique+findFundamentalMat, one process per core (the DMatch "
                           "rows are read from Python: ~1 %% of a pair)" % cv2_ref.cv2.__version__ if use_cv2 else
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
        return self

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


# ------------------------------------------------------------------------------------------------
class Dist:
    """torch.distributed plumbing: barrier, max / sum over ranks, broadcast of a few bytes."""

    def __init__(self, world, rank, local):
        import torch
        self.torch, self.world, self.rank, self.local = torch, world, rank, local
        torch.cuda.set_device(local)
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            self.dist = dist

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def _red(self, x, op):
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def allmax(self, x):
        return self._red(x, self.dist.ReduceOp.MAX) if self.dist else x

    def allsum(self, x):
        return self._red(x, self.dist.ReduceOp.SUM) if self.dist else x

    def bcast_bytes(self, b: bytes | None, n: int) -> bytes:
        if not self.dist:
            return b
        t = self.torch.zeros(n, dtype=self.torch.uint8, device="cuda")
        if self.rank == 0:
            t.copy_(self.torch.frombuffer(bytearray(b), dtype=self.torch.uint8))
        self.dist.broadcast(t, src=0)
        return bytes(t.cpu().numpy().tobytes())

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


class Workload:
    """One image set on this rank: pinned host copies of the images this rank 'extracted', a PairMatcher, the pair share."""

    def __init__(self, a, D: Dist, kind, n_img, outlier_frac, pairs, images_np=None, **pm_kw):
        import torch
        from reconstructor_b200 import api
        self.a, self.D, self.kind, self.n_img, self.api = a, D, kind, n_img, api
        self.pairs_all = pairs
        self.mine = shard.shard_pairs(pairs, D.rank, D.world)
        self.sharded = D.world > 1 and not a.replicated_ingest
        own = list(range(D.rank, n_img, D.world)) if self.sharded else list(range(n_img))
        self.own = own
        self.dt = api.DESC_U8_BITS if kind == "orb" else api.DESC_F32
        # integer-valued SIFT rows travel as bytes between the ranks (4x fewer bytes); the rank's own upload is fp32,
        # what FeatDesc holds (datatypes.h:70-71)
        self.wire = api.DESC_U8 if kind == "sift" else self.dt
        cols = 32 if kind == "orb" else (128 if kind == "sift" else 256)
        self.dim = 256 if kind == "orb" else cols
        self.d_own = torch.empty((max(len(own), 1), a.kp, cols), dtype=torch.uint8 if kind == "orb" else torch.float32,
                                 pin_memory=True)
        self.x_own = torch.empty((max(len(own), 1), a.kp, 2), dtype=torch.int32, pin_memory=True)
        dn, xn = self.d_own.numpy(), self.x_own.numpy()
        w = synth.World(kind, a.kp, seed=0xB200 + 2)
        for k, i in enumerate(own):
            d, xy = images_np[i] if images_np is not None else w.image(i, n_img, outlier_frac)[:2]
            dn[k] = d; xn[k] = xy
        self.pm = api.PairMatcher(devices=[D.local], reserve_keypoints=n_img * a.kp, debug_flags=a.debug_flags,
                                  batch_pairs=a.batch_pairs, **pm_kw)
        if self.sharded:
            uid = D.bcast_bytes(api.comm_unique_id() if D.rank == 0 else None, 128)
            self.pm.comm_init(uid, D.rank, D.world)

    def ingest(self, asynchronous=False, pm=None):
        pm = pm or self.pm
        a = self.a
        if self.sharded:
            pm.ingest_allgather(self.n_img, a.kp, self.dim, self.dt, self.wire, self.d_own.data_ptr(), self.x_own.data_ptr())
            if not asynchronous:
                pm.sync_images()
            return
        ids = self.own
        if asynchronous:
            # pm_set_images_async from pinned buffers: the uploads are queued and the first batches of pm_match_all_pairs
            # run while the later images are still on their way (the pair list is ordered image by image)
            pm.set_images_ptr_async(ids, [self.d_own[k].data_ptr() for k in range(len(ids))], [a.kp] * len(ids), self.dim,
                                    self.dt, [self.x_own[k].data_ptr() for k in range(len(ids))])
            return
        for k, i in enumerate(ids):
            pm.set_image_ptr(i, self.d_own[k].data_ptr(), a.kp, self.dim, self.dt, self.x_own[k].data_ptr())

    def run_resident(self, steps, warmup, sample_clocks=True):
        """Device-timed steps with the images resident.  Returns dict(ms_per_step, wall_ms_per_step, stats, clocks, ...)."""
        pm, D = self.pm, self.D
        for _ in range(warmup):
            r = pm.match_all_pairs(self.mine, copy=False); pm.free_result(r)
        sampler = ClockSampler(D.local).start() if (D.rank == 0 and sample_clocks) else None
        D.barrier()
        pm.reset_stats()
        dev_ms = 0.0
        matches = inliers = 0
        t0 = time.perf_counter()
        for _ in range(steps):
            r = pm.match_all_pairs(self.mine, copy=False)
            dev_ms += r["device_ms"]
            matches = int(r["offsets"][-1]); inliers = int(r["n_inliers"].sum())
            pm.free_result(r)
        D.barrier()
        wall_ms = 1e3 * (time.perf_counter() - t0)
        clocks = sampler.stop() if sampler else None
        st = pm.stats()
        return dict(ms_per_step=D.allmax(dev_ms) / steps, wall_ms_per_step=D.allmax(wall_ms) / steps, stats=st, clocks=clocks,
                    launches=D.allsum(st["kernel_launches"]), matches=matches, inliers=inliers)

    def run_e2e(self, steps, warm=True):
        pm, D, a = self.pm, self.D, self.a
        use_async = not a.sync_ingest
        if warm:
            self.ingest(use_async); r = pm.match_all_pairs(self.mine, copy=False); pm.free_result(r)
        D.barrier()
        pm.reset_stats()
        t0 = time.perf_counter()
        trace = os.environ.get("PM_BENCH_TRACE")
        for _ in range(steps):
            ta = time.perf_counter()
            self.ingest(use_async)
            tb = time.perf_counter()
            r = pm.match_all_pairs(self.mine, copy=False)
            tc = time.perf_counter()
            _ = int(r["n_inliers"].sum())                     # the step's result is read on the host
            if trace:                                         # development: host-side split of the end-to-end step
                print("[bench trace] rank %d: ingest call %.2f ms, match call %.2f ms (device %.2f ms)" %
                      (D.rank, 1e3 * (tb - ta), 1e3 * (tc - tb), r["device_ms"]), file=sys.stderr)
            pm.free_result(r)
        D.barrier()
        e_ms = D.allmax(1e3 * (time.perf_counter() - t0)) / steps
        st2 = pm.stats()
        out = dict(value=len(self.pairs_all) / (e_ms * 1e-3), unit=UNIT, ms_per_step=e_ms,
                   h2d_bytes_per_step=int(D.allsum(st2["h2d_bytes"]) / steps),
                   d2h_bytes_per_step=int(D.allsum(st2["d2h_bytes"]) / steps),
                   timing="host wall clock around %s + pm_match_all_pairs, max over ranks" %
                          ("pm_ingest_allgather (own images H2D, NCCL all-gather in the wire dtype, device ingest)" if self.sharded
                           else "pm_set_images_async" if use_async else "pm_set_image x images"))
        if self.sharded:
            row = {self.api.DESC_F32: 4 * self.dim, self.api.DESC_U8: self.dim, self.api.DESC_U8_BITS: self.dim // 8}[self.wire]
            out["allgather_bytes_per_step"] = int(self.n_img * a.kp * (row + 8) * (D.world - 1))
            out["wire_dtype"] = {self.api.DESC_F32: "f32", self.api.DESC_U8: "u8", self.api.DESC_U8_BITS: "bits"}[self.wire]
        return out

    def close(self):
        self.pm.close()
        self.d_own = self.x_own = None


# ------------------------------------------------------------------------------------------------
def kernel_name(kind, flags):
    if kind == "sift":
        return "l2_top2_tc2_kernel" if (flags & (2048 | 16384)) else "l2_i8x2_kernel"
    if kind == "orb":
        form = orb_form(flags)
        return {"popc": "hamming_top2_kernel", "e4m3": "l2_top2_tc2_kernel<T2Cfg<256,2,2>,2,false,1>",
                "i8": "l2_i8x2_kernel<2,false,2,0>", "fp4": "l2_i8x2_kernel<2,3,1,1,0,0,1>", "fp4-1row": "l2_i8x2_kernel<2,false,1,1>"}[form]
    if flags & 32768:
        return "l2_top2_tc2_kernel<T2Cfg<256,2,4>,3>"
    if flags & 524288:
        return "l2_top2_tc2_kernel<T2Cfg<256,2,2>,3,true,3>"
    if flags & 1048576:
        return "l2_i8x2_kernel<3,2,2,0,3>"
    # unit-norm rows (SuperPoint, the synthetic set): the form without the norm K-step; debug bit 23 keeps it
    return "l2_i8x2_kernel<4,2,2,0,3>" if (flags & 8388608) else "l2_i8x2_kernel<4,2,2,0,3,1>"


def orb_form(flags):
    # fp4 = the packed-pair form (two train rows per accumulator column, the default); debug bit 26 keeps one row per column
    return ("popc" if (flags & 1024) else "e4m3" if (flags & 65536) else "i8" if (flags & 262144) else
            "fp4-1row" if (flags & 67108864) else "fp4")


_PEAK_CACHE = {}


def measured_peak(pm, api, kind_id):
    """FLOP/s of the pure tcgen05.mma issue loop of one MMA kind, measured once per process."""
    if kind_id not in _PEAK_CACHE:
        _PEAK_CACHE[kind_id] = pm.measure_tensor_peak(kind_id)
    return _PEAK_CACHE[kind_id]


def roofline(wl: Workload, res, peaks, world):
    """Roofline of the dominant (kNN) kernel of a resident run: algorithmic work / CUDA-event time of its launches."""
    a, api, pm, st = wl.a, wl.api, wl.pm, res["stats"]
    kind, flags = wl.kind, a.debug_flags
    knn_s = max(st["knn_ms"] * 1e-3, 1e-12)
    bf16 = peaks.get("bf16_tflops_sustained")
    if kind == "orb" and orb_form(flags) == "popc":
        pk = pm.measure_popc_peak()
        roof = dict(bound="popc", achieved=st["knn_work"] / knn_s / 1e12, peak=pk / 1e12, unit="Tpopc32/s",
                    peak_source="measured live: pm_measure_popc_peak (dependent-chain POPC micro-benchmark)")
    else:
        if kind == "orb":
            # Hamming = |a| + |b| - 2 a.b on the tensor cores; one popc32 of the fixed numerator = 32 bit
            # compares = 64 FLOP of the contraction
            form = orb_form(flags)
            ach = 64.0 * st["knn_work"] / knn_s / 1e12
            kid, kname = ((api.PEAK_KIND_MXF4, "kind::mxf4") if form.startswith("fp4") else
                          (api.PEAK_KIND_I8, "kind::i8 (= kind::f8f6f4 rate)"))
            ops = {"fp4": "E2M1 {0, +-1} values on tcgen05 kind::mxf4 (64 values of K per instruction); two train rows per f32 "
                          "accumulator column through the UE8M0 block scales 2^10 | 1 (9 K-steps per 384 train rows), exact integers",
                   "fp4-1row": "E2M1 {0, +-1} values on tcgen05 kind::mxf4 (64 values of K per instruction, all-ones scale factors), "
                               "f32 accumulate, exact integers", "i8": "u8 x s8 on tcgen05 kind::i8, s32 accumulate",
                   "e4m3": "E4M3 {0,1} values on tcgen05 kind::f8f6f4, f32 accumulate, exact integers"}[form]
        else:
            ach = st["knn_work"] / knn_s / 1e12
            fp16 = (kind == "sift" and (flags & 2048)) or (kind == "superpoint" and (flags & 32768))
            kid, kname = (api.PEAK_KIND_F16, "kind::f16") if fp16 else (api.PEAK_KIND_I8, "kind::i8")
            ops = ("f16 operands, f32 accumulate" if fp16 else
                   "u8 x s8 -> s32 (exact integers)" if kind == "sift" else
                   "rows quantised to s8 for the candidate stage (s8 x s8 -> s32); the exact fp32 re-rank follows")
        pk = measured_peak(pm, api, kid) / 1e12
        roof = dict(bound="tensor", achieved=ach, peak=pk, unit="TFLOP/s", operands=ops,
                    peak_source=f"measured live in this run: pm_measure_tensor_peak({kname}) -- the tcgen05.mma issue loop of the "
                                "kernel (cta_group::2, M=256, operands resident in shared memory), nothing else",
                    bf16_sustained_peak=bf16, frac_of_bf16_sustained=(ach / bf16) if bf16 else None,
                    bf16_peak_source="MEASURED_PEAKS.json bf16_tflops_sustained (cuBLAS, other MMA kind: a frac above 1 is expected "
                                     "for kind::i8 (2x) and kind::mxf4 (4x))")
        if kind == "orb":
            roof["popc32_equiv"] = dict(achieved=st["knn_work"] / knn_s / 1e12, unit="Tpopc32/s",
                                        note="the XOR/popc kernel of the north star (--debug-flags 1024) is bounded by the popc pipe")
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["kernel"] = kernel_name(kind, flags)
    roof["launches"] = st["knn_launches"]
    roof["avg_launch_ms"] = st["knn_ms"] / max(st["knn_launches"], 1)
    roof["share_of_step"] = st["knn_ms"] / max(res["ms_per_step"] * max(res.get("steps", 1), 1), 1e-9) if world == 1 else None
    if kind == "superpoint":
        roof["rerank"] = {k[7:]: st[k] for k in st if k.startswith("rerank_")}
    roof["traffic"] = None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(roof["kernel"])
        if t:
            pairs_per_launch = len(wl.mine) * max(res.get("steps", 1), 1) / max(st["knn_launches"], 1)
            roof["traffic"] = t["bytes_per_launch"] * pairs_per_launch / t["pairs_per_launch"]
            roof["traffic_source"] = t["source"]
    except Exception:
        pass
    return roof


def run_config(a, D, peaks, kind, n_img, outlier_frac, steps, warmup, e2e_steps, tag):
    """One entry of `configs`: the whole job on all N GPUs (strong scaling)."""
    pairs = shard.all_pairs(n_img)
    t0 = time.perf_counter()
    wl = Workload(a, D, kind, n_img, outlier_frac, pairs)
    t_gen = time.perf_counter() - t0
    wl.ingest()
    res = wl.run_resident(steps, warmup)
    res["steps"] = steps
    roof = roofline(wl, res, peaks, D.world)
    e2e = wl.run_e2e(e2e_steps, warm=False) if e2e_steps > 0 else None
    wl.close()
    return dict(tag=tag, workload=workload_name(kind, n_img, a.kp, len(pairs)), kind=kind, images=n_img, keypoints=a.kp,
                pairs=int(len(pairs)), n_gpus=D.world, scaling="strong", outlier_frac=outlier_frac,
                value=len(pairs) / (res["ms_per_step"] * 1e-3), unit=UNIT, ms_per_step=res["ms_per_step"],
                wall_ms_per_step=res["wall_ms_per_step"], steps=steps, warmup=warmup, putative_matches_per_step=res["matches"]
                if D.world == 1 else None, e2e=e2e, roofline=roof, clocks=res["clocks"], gpu_launches=int(res["launches"]),
                setup_s=round(t_gen, 1))


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
    cfg = dict(workload=workload_name(a.kind, n_img, a.kp, n_all), images=n_img, keypoints=a.kp, kind=a.kind,
               pairs=int(len(pairs)), partition=f"pairs sharded over {n_gpus} GPU(s) (equal block-cyclic shares), descriptors replicated",
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

    # ---- headline images as numpy first: the CPU baseline forks before any CUDA context exists (rank 0, N=1) -------
    sharded = world > 1 and not a.replicated_ingest
    w = synth.World(a.kind, a.kp, seed=0xB200 + 2)
    own_ids = list(range(rank, n_img, world)) if sharded else list(range(n_img))
    images_np = {i: w.image(i, n_img, a.outlier_frac)[:2] for i in own_ids}
    cpu, flann = None, None
    if n_gpus == 1 and rank == 0 and not a.no_cpu_baseline:
        sub_ids = list(range(min(n_img, 24)))
        sub = np.array([(i, j) for x, i in enumerate(sub_ids) for j in sub_ids[x + 1:]], np.int32)
        r = cpu_arm({i: images_np[i] for i in sub_ids}, sub, a.cpu_seconds)
        cpu = dict(value=r["value"], unit=UNIT, cores=r["cores"], kind=r["kind"], sample=r["sample"])
        if r["cores"] > 4:      # the reference's own thread cap (MAX_NUM_THREADS 4, SequentialReconstructor.h:17)
            r4 = cpu_arm({i: images_np[i] for i in sub_ids}, sub, min(a.cpu_seconds, 6.0), workers=4)
            cpu["at_4_workers"] = dict(value=r4["value"], unit=UNIT, cores=4, sample=r4["sample"])
        flann = flann_samples({i: images_np[i] for i in sub_ids},
                              [tuple(int(x) for x in sub[k]) for k in (0, 1, len(sub) // 2, len(sub) - 1)][:len(sub)])

    D = Dist(world, rank, local)
    from reconstructor_b200 import api
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    dev_kw = {}
    if a.dev_ratio is not None: dev_kw["ratio"] = a.dev_ratio
    if a.dev_no_filter: dev_kw["do_filter"] = 0
    wl = Workload(a, D, a.kind, n_img, a.outlier_frac, pairs, images_np=images_np, **dev_kw)
    images_np = None
    if sharded:
        cfg["partition"] += ("; extraction sharded by image id mod N: per step every rank uploads its own images and "
                             "pm_ingest_allgather (NCCL, wire dtype %s) replicates them" % ("u8" if a.kind == "sift" else "as ingested"))
    wl.ingest()
    res = wl.run_resident(a.steps, a.warmup)
    res["steps"] = a.steps
    ms_per_step = res["ms_per_step"]
    total_pairs = len(pairs)
    value = total_pairs / (ms_per_step * 1e-3)
    roof = roofline(wl, res, peaks, world)

    # ---- attribution: the same workload with the epipolar filter switched off (matcher only); the filter's cost is the
    #      difference ("matcher-only and RANSAC-only pairs/s") -------------------------------------------
    stages = None
    if not a.no_stages and world == 1:
        pm2 = api.PairMatcher(devices=[local], reserve_keypoints=n_img * a.kp, debug_flags=a.debug_flags,
                              batch_pairs=a.batch_pairs, do_filter=0)
        wl.ingest(pm=pm2)
        r = pm2.match_all_pairs(wl.mine, copy=False); pm2.free_result(r)
        D.torch.cuda.synchronize()
        m_ms = 0.0
        for _ in range(a.steps):
            r = pm2.match_all_pairs(wl.mine, copy=False)
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
            gi, _ = wl.pm.knn_pair(i, j)
            agree += int((gi[:, 0] == top1).sum()); tot += len(top1)
            g = wl.pm.match_filter_pair(i, j)
            keep = g["inlier"].astype(bool)
            gs = set(zip(g["q"][keep].tolist(), g["t"][keep].tolist()))
            fs = set(map(tuple, final.tolist()))
            inter += len(gs & fs); union += len(gs | fs)
        cpu["recall"] = dict(pairs=len(flann), flann_top1_agreement=agree / max(tot, 1),
                             final_match_set_iou=inter / max(union, 1),
                             note="reference FLANN (approximate) vs this library's exact search, same pairs")

    # ---- e2e: host buffers in, host CSR out, every step ---------------------------------------------
    e2e = None if a.no_e2e else wl.run_e2e(a.steps)
    wl.close()

    # ---- the headline workload with 50 % outlier keypoints: the filter runs hundreds of iterations ---------
    if not a.no_stages and a.outlier_frac == 0.0:
        rh = run_config(a, D, peaks, a.kind, n_img, 0.5, a.cfg_steps, 1, 0, "ransac_heavy")
        heavy = dict(value=rh["value"], unit=UNIT, ms_per_step=rh["ms_per_step"], outlier_frac=0.5, steps=rh["steps"],
                     what="the headline workload with half of the keypoints moved to random pixels",
                     kNN_share_of_step=rh["roofline"]["share_of_step"], clocks=rh["clocks"])
        stages = dict(stages or {}, ransac_heavy=heavy)

    # ---- the other BASELINE.json configurations (strong scaling: the same job on all N GPUs) --------------------------
    configs = None
    if not a.no_configs:
        configs = {}
        configs["orb500"] = run_config(a, D, peaks, "orb", a.orb_images, 0.0, a.cfg_steps, 1, 1, "BASELINE.json configs[2]")
        configs["superpoint1000"] = run_config(a, D, peaks, "superpoint", a.sp_images, 0.0, max(1, a.cfg_steps - 1), 1, 1,
                                               "BASELINE.json configs[3]")

    if rank == 0:
        flags = a.debug_flags
        dtype = {"sift": "f16 operands / f32 accumulate (exact integers)" if (flags & 2048) else
                 "u8 x s8 operands / s32 accumulate (kind::i8, exact integers)",
                 "orb": {"popc": "u32 popc", "e4m3": "e4m3 {0,1} operands / f32 accumulate (exact integers)",
                         "i8": "u8 x s8 operands / s32 accumulate (exact integers)",
                         "fp4": "e2m1 {0,+-1} operands (kind::mxf4) / f32 accumulate (exact integers)",
                         "fp4-1row": "e2m1 {0,+-1} operands (kind::mxf4) / f32 accumulate (exact integers)"}[orb_form(flags)],
                 "superpoint": "f16 operands / f32 accumulate candidates + exact f32 re-rank" if (flags & 32768) else
                 "s8 operands / s32 accumulate candidates (kind::i8) + exact f32 re-rank"}[a.kind]
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=n_gpus, steps=a.steps, warmup=a.warmup,
                    ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype=dtype,
                    data="synthetic", config=cfg, wall_ms_per_step=res["wall_ms_per_step"],
                    putative_matches_per_step=res["matches"], inliers_per_step=res["inliers"],
                    roofline=roof, stages=stages, cpu_baseline=cpu, e2e=e2e, gpu_launches=int(res["launches"]),
                    clocks=res["clocks"], configs=configs)
        print(json.dumps(line))
    D.close()


if __name__ == "__main__":
    main()
