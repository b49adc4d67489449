"""Host-side restatement of the arithmetic behind the packed-pair Hamming kernel (csrc/l2_tc2.cu, l2_i8x2_kernel PK;
csrc/pack.cu, pack_bits_kernel): the E2M1 norm block with its bias slots, the fp32 accumulator that holds two distances, and
the shift + 16x2 minima that read them.  No GPU: numpy float32 / uint32 only.  The kernel itself is compared bit for bit with
the XOR/popc kernel and the oracle in tests/test_gpu_parity.py."""
import numpy as np

E2M1 = np.array([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0], np.float32)          # value of code c (sign bit clear)
CODE = {0: 0, 1: 2, 2: 4, 3: 5, 4: 6, 6: 7}


def norm_digits(cnt, ns6=7):
    """pack.cu fp4_norm_digit: |b| = 36 n + 6 u + v in slots of weight 6 (ns6 + 2 of them) and 1 (2 of them)."""
    n36 = min(cnt // 36, ns6); r = cnt - 36 * n36; u, v = divmod(r, 6)
    d = [6 if s < n36 else 0 for s in range(ns6)]
    d += [min(u, 4), u - min(u, 4), min(v, 4), v - min(v, 4)]
    return d


def query_block():
    """64 slots: weights 6 x9, 1 x2, bias 6 x14, 2 -- and the same again in the second K block (slots 32..)."""
    blk = [6] * 9 + [1] * 2 + [6] * 14 + [2] + [0] * 6
    return np.array(blk + blk, np.float32)


def train_block(cnt):
    blk = norm_digits(cnt) + [6] * 14 + [4] + [0] * 6
    assert all(x in CODE for x in blk)                                           # every digit is an E2M1 value
    return np.array(blk + [0] * 32, np.float32)


def test_norm_block_adds_popcount_plus_bias_for_every_popcount():
    q = query_block()
    for cnt in range(257):
        t = train_block(cnt)
        assert len(t) == 64 and float(q @ t) == cnt + 512, cnt
        # pair norm block (t4x): first K block = this row, second = the row 192 further on; block scales 2^10 | 1
        for cnt2 in (0, 1, 37, 255, 256):
            t2 = train_block(cnt2)
            pair = np.concatenate([t[:32], t2[:32]])
            assert 1024.0 * float(q[:32] @ pair[:32]) + float(q[32:] @ pair[32:]) == 1024 * (cnt + 512) + cnt2 + 512


def _fields(v):
    bits = v.astype(np.float32).view(np.uint32)
    w = (bits.astype(np.uint64) * 4 & 0xFFFFFFFF).astype(np.uint32)              # IMAD.SHL by 2
    return bits, w >> 16, w & 0xFFFF


def test_two_distances_per_accumulator_every_pair_of_values():
    h_hi, h_lo = np.meshgrid(np.arange(257), np.arange(257), indexing="ij")
    v = (2.0 ** 19 + 1024.0 * h_hi + 512.0 + h_lo).astype(np.float32)           # exact: integers below 2^20
    assert np.array_equal(v.astype(np.int64), 2 ** 19 + 1024 * h_hi + 512 + h_lo)
    bits, hi16, lo16 = _fields(v)
    assert np.all(bits >> 23 == 127 + 19)                                        # one binade
    assert np.array_equal(hi16 & 0x1FF, h_hi) and np.all(hi16 >> 9 == 0x2400 >> 9)
    assert np.array_equal((lo16 >> 6) - 512, h_lo) and np.all(lo16 & 63 == 0)
    # a second row past the image adds between 256 and 768 (bias +- the main K-steps of whatever row lies there), or
    # nothing at all (zero rows, or a tile whose second sub-tile is not multiplied): the high field is untouched
    for stray in (0, 256, 300, 768):
        b2, hi2, _ = _fields((2.0 ** 19 + 1024.0 * h_hi[:, 0] + stray).astype(np.float32))
        assert np.array_equal(hi2 & 0x1FF, h_hi[:, 0]) and np.all(b2 >> 23 == 146)


def test_one_tree_of_16x2_minima_reduces_both_fields():
    rng = np.random.default_rng(3)
    for _ in range(200):
        h_hi = rng.integers(0, 257, 32); h_lo = rng.integers(0, 257, 32)
        lim_hi = int(rng.integers(1, 33)); lim_lo = int(rng.integers(0, lim_hi + 1))      # ragged chunk
        _, hi16, lo16 = _fields((2.0 ** 19 + 1024.0 * h_hi + 512.0 + h_lo).astype(np.float32))
        hi16 = np.where(np.arange(32) >= lim_hi, 0xFFFF, hi16)                  # w = 0xFFFFFFFF
        lo16 = np.where(np.arange(32) >= lim_lo, 0xFFFF, lo16)                  # w |= 0x0000FFFF
        assert int(hi16.min()) & 0x1FF == h_hi[:lim_hi].min()
        if lim_lo > 0:
            assert (int(lo16.min()) >> 6) - 512 == h_lo[:lim_lo].min()
