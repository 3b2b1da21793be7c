"""Bit-level model of the block-exact ordered sum used by the LM kernels (csrc/graph_dev.cuh::warp_ordered_sum): the same decisions in
Python integers, checked against one-by-one fp64 addition on adversarial inputs.  The CUDA routine itself is checked on the device by
tests/test_gpu_parity.py::test_ordered_sum_in_steps_of_32_is_bit_identical_to_the_plain_loop; this file pins the ARGUMENT (why a block
of non-negative terms can be added with integer arithmetic while the running sum stays inside one binade) without a GPU."""
import struct

import numpy as np

KM = (1 << 52) - 1


def bits(x):
    return struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def dbl(b):
    return struct.unpack("<d", struct.pack("<Q", b))[0]


def inc(es, c):
    """ordered_sum_inc: increment of the integer mantissa of a sum in binade `es` by the term c, or ok = False."""
    if c == 0.0:
        return 0, True
    cb = bits(c); ec = (cb >> 52) & 0x7ff; shift = es - ec
    if (cb >> 63) != 0 or ec == 0 or ec == 0x7ff or shift < 0:
        return 0, False
    if shift >= 54:
        return 0, True
    mc = (cb & KM) | (1 << 52)
    if shift == 0:
        return mc, True
    q = mc >> shift; f = mc & ((1 << shift) - 1); half = 1 << (shift - 1)
    if f == half:
        return 0, False
    return q + (1 if f > half else 0), True


def commit(s, total):
    sb = bits(s); ms = (sb & KM) | (1 << 52)
    if total >= (1 << 53) or ms + total >= (1 << 53):
        return s, False
    return np.float64(dbl((sb & ~KM) | ((ms + total) & KM))), True


def try_block(s, terms):
    sb = bits(s); es = (sb >> 52) & 0x7ff
    if (sb >> 63) != 0 or not (0 < es < 0x7ff):
        return s, False
    total = 0
    for c in terms:
        d, ok = inc(es, c)
        if not ok:
            return s, False
        total += d
    return commit(s, total)


def block_sum(v, acc=0.0):
    s = np.float64(acc); n = len(v); steps = [0, 0, 0]
    for base in range(0, n, 256):
        blk = [np.float64(x) for x in v[base:base + 256]]
        s2, ok = try_block(s, blk)
        if ok:
            s = s2; steps[0] += 1
            continue
        for j in range(0, len(blk), 32):
            ch = blk[j:j + 32]
            s2, ok = try_block(s, ch)
            if ok:
                s = s2; steps[1] += 1
            else:
                steps[2] += 1
                for c in ch:
                    s = np.float64(s + c)
    return float(s), steps


def plain(v, acc=0.0):
    s = np.float64(acc)
    for x in v:
        s = np.float64(s + np.float64(x))
    return float(s)


def test_block_sum_equals_the_one_by_one_sum_bit_for_bit():
    rng = np.random.default_rng(1)
    fast = slow = 0
    for trial in range(90):
        n = int(rng.integers(0, 1500)); kind = trial % 6
        if kind == 0:
            v = rng.exponential(10.0, n) ** 2
        elif kind == 1:
            v = 10.0 ** rng.uniform(-30, 30, n)
        elif kind == 2:
            v = rng.exponential(1.0, n)
            if n:
                v[rng.integers(0, n, max(1, n // 10))] = 0.0
        elif kind == 3:
            v = np.abs(rng.normal(size=n)) * 2.0 ** rng.integers(-60, 60, n)
        elif kind == 4:
            v = rng.exponential(100.0, n)
            if n > 5:
                v[3] = -1.0; v[n // 2] = np.inf if trial % 12 == 4 else 5e-324
        else:   # engineered ties and near-ties at the half-ulp of a running sum of 1
            v = np.r_[1.0, [2.0 ** -53] * 40, [2.0 ** -52] * 40, [1.5 * 2.0 ** -53] * 40, rng.exponential(1e-15, n)]
        acc = float(rng.choice([0.0, 1e-300, 3.7, 1e300])) if trial % 7 == 0 else 0.0
        with np.errstate(all="ignore"):
            (a, st), b = block_sum(list(v), acc), plain(list(v), acc)
        assert (a != a and b != b) or bits(a) == bits(b), (kind, a, b)
        fast += st[0] + st[1]; slow += st[2]
    assert fast > slow   # the shortcut still carries most steps of this adversarial mix (wide dynamic ranges, ties, odd values)


def test_chi_square_like_vector_takes_the_shortcut_almost_everywhere():
    rng = np.random.default_rng(2)
    v = list(rng.exponential(20.0, 25050))
    (a, st), b = block_sum(v), plain(v)
    assert bits(a) == bits(b) and st[2] <= 40 and st[0] >= 80   # a few dozen one-by-one steps (binade crossings) out of 98 blocks / 783 steps
