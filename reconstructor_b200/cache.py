"""Host-side mirror of the on-disk cache format of libpairmatch_b200 (csrc/api.cu, "On-disk cache").

The reference keeps everything in memory and lists "save intermediate steps" as a todo (README.md:39); SURVEY 8f
rank 2 asks for a binary format for the ingested descriptors and for the CSR match result
(`featureMatches`, SequentialReconstructor.h:226, in flat form).  One little-endian container:

    header (40 bytes)   magic "PMB200\\0\\1" | u32 version = 1 | u32 kind (1 images, 2 result) | u64 count |
                        u64 payload bytes | u64 checksum
    payload             sections, each zero-padded to a multiple of 8 bytes
    checksum            s1 ^ rotl64(s2, 32) with s1 = sum of the payload's u64 words, s2 = sum of (i+1) * word_i
                        (both mod 2^64)

kind 1 (images), per image:  i32[6] {id, n, dim, dtype, has_xy, 0} | descriptors (n*dim f32 / n*dim u8 /
                             n*dim/8 bytes of bits) | n*2 i32 xy if has_xy
kind 2 (result):             i64[4] {n_pairs, n_matches, 0, 0} | f64 device_ms | pair_ij i32[2P] | offsets i64[P+1] |
                             q i32[M] | t i32[M] | inlier u8[M] | F f64[9P] | status i32[P] | n_inliers i32[P] |
                             ransac_iters i32[P]

This module reads and writes the same bytes with numpy only, so files can be produced and inspected on machines
without a GPU (sharded extraction, tests); it shares no code with the library.
"""
from __future__ import annotations

import struct

import numpy as np

MAGIC = b"PMB200\x00\x01"
KIND_IMAGES, KIND_RESULT = 1, 2
DESC_F32, DESC_U8_BITS, DESC_U8 = 0, 1, 2


def checksum(payload: bytes) -> int:
    w = np.frombuffer(payload, dtype="<u8")
    with np.errstate(over="ignore"):
        s1 = int(w.sum(dtype=np.uint64))
        s2 = int((w * np.arange(1, len(w) + 1, dtype=np.uint64)).sum(dtype=np.uint64))
    return s1 ^ (((s2 << 32) | (s2 >> 32)) & 0xFFFFFFFFFFFFFFFF)


def _pad(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


def _write(path: str, kind: int, count: int, sections: list[bytes]):
    payload = b"".join(_pad(s) for s in sections)
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<IIQQQ", 1, kind, count, len(payload), checksum(payload)))
        f.write(payload)


def _read(path: str, kind: int) -> tuple[int, bytes]:
    with open(path, "rb") as f:
        head = f.read(40)
        payload = f.read()
    if len(head) != 40 or head[:8] != MAGIC:
        raise ValueError(f"{path}: not a pairmatch_b200 cache file")
    version, k, count, nbytes, ck = struct.unpack("<IIQQQ", head[8:])
    if version != 1 or k != kind or nbytes != len(payload) or nbytes % 8 or checksum(payload) != ck:
        raise ValueError(f"{path}: version / kind / size / checksum mismatch")
    return count, payload


class _Cursor:
    def __init__(self, buf: bytes):
        self.buf, self.pos = buf, 0

    def take(self, dtype, count) -> np.ndarray:
        n = np.dtype(dtype).itemsize * count
        if self.pos + n > len(self.buf):
            raise ValueError("truncated section")
        a = np.frombuffer(self.buf, dtype=dtype, count=count, offset=self.pos)
        self.pos += n + (-n % 8)
        return a


def write_images(path: str, images: dict[int, tuple[np.ndarray, np.ndarray | None]], dtype: int | None = None):
    """images: {id: (descriptors [n, dim] float32 or uint8, xy int32 [n, 2] or None)}."""
    sections = []
    for img_id in sorted(images):
        desc, xy = images[img_id]
        dt = dtype if dtype is not None else (DESC_U8_BITS if desc.dtype == np.uint8 else DESC_F32)
        if dt == DESC_F32:
            d = np.ascontiguousarray(desc, "<f4"); dim = d.shape[1]
        else:
            d = np.ascontiguousarray(desc, np.uint8); dim = d.shape[1] * (8 if dt == DESC_U8_BITS else 1)
        sections.append(np.array([img_id, d.shape[0], dim, dt, 0 if xy is None else 1, 0], "<i4").tobytes())
        sections.append(d.tobytes())
        if xy is not None:
            sections.append(np.ascontiguousarray(xy, "<i4").reshape(-1, 2).tobytes())
    _write(path, KIND_IMAGES, len(images), sections)


def read_images(path: str, headers_only: bool = False) -> list[dict]:
    count, payload = _read(path, KIND_IMAGES)
    c, out = _Cursor(payload), []
    for _ in range(count):
        img_id, n, dim, dt, has_xy, _z = (int(v) for v in c.take("<i4", 6))
        if dt == DESC_F32:
            desc = c.take("<f4", n * dim).reshape(n, dim)
        elif dt == DESC_U8:
            desc = c.take(np.uint8, n * dim).reshape(n, dim)
        else:
            desc = c.take(np.uint8, n * (dim // 8)).reshape(n, dim // 8)
        xy = c.take("<i4", 2 * n).reshape(n, 2) if has_xy else None
        rec = dict(id=img_id, n=n, dim=dim, dtype=dt)
        if not headers_only:
            rec.update(desc=desc, xy=xy)
        out.append(rec)
    if c.pos != len(payload):
        raise ValueError(f"{path}: trailing bytes")
    return out


def write_result(path: str, res: dict):
    """res: the dict PairMatcher.match_all_pairs returns."""
    p = int(res["n_pairs"]); m = int(res["offsets"][p]) if p > 0 else 0
    sections = [np.array([p, m, 0, 0], "<i8").tobytes(), np.array([res.get("device_ms", 0.0)], "<f8").tobytes(),
                np.ascontiguousarray(res["pair_ij"], "<i4").tobytes(),
                np.ascontiguousarray(res["offsets"] if p > 0 else [0], "<i8").tobytes(),
                np.ascontiguousarray(res["q"], "<i4").tobytes(), np.ascontiguousarray(res["t"], "<i4").tobytes(),
                np.ascontiguousarray(res["inlier"], np.uint8).tobytes(), np.ascontiguousarray(res["F"], "<f8").tobytes(),
                np.ascontiguousarray(res["status"], "<i4").tobytes(), np.ascontiguousarray(res["n_inliers"], "<i4").tobytes(),
                np.ascontiguousarray(res["ransac_iters"], "<i4").tobytes()]
    _write(path, KIND_RESULT, p, sections)


def read_result(path: str) -> dict:
    count, payload = _read(path, KIND_RESULT)
    c = _Cursor(payload)
    p, m, _a, _b = (int(v) for v in c.take("<i8", 4))
    if p != count:
        raise ValueError(f"{path}: pair count mismatch")
    out = dict(n_pairs=p, device_ms=float(c.take("<f8", 1)[0]))
    out["pair_ij"] = c.take("<i4", 2 * p).reshape(-1, 2)
    out["offsets"] = c.take("<i8", p + 1)
    out["q"] = c.take("<i4", m); out["t"] = c.take("<i4", m); out["inlier"] = c.take(np.uint8, m)
    out["F"] = c.take("<f8", 9 * p).reshape(-1, 3, 3)
    out["status"] = c.take("<i4", p); out["n_inliers"] = c.take("<i4", p); out["ransac_iters"] = c.take("<i4", p)
    if c.pos != len(payload):
        raise ValueError(f"{path}: trailing bytes")
    return out
