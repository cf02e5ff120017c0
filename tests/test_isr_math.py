"""Closed forms used by the device ISR (csrc/isr_device.cuh) vs the oracle's literal restatement of the
reference's integer helpers (which is pinned against the compiled reference)."""
import ctypes as C
import math

import numpy as np


def _closed_isqrt(L: int) -> int:
    x = int(math.sqrt(L) + 0.5) or 1
    while x * (x - 1) > L:
        x -= 1
    while (x + 1) * x <= L:
        x += 1
    return x


def test_sqrt_newton_closed_form(oracle_lib):
    """sqrt_newton(L) == max{x : x(x-1) <= L} for 0 < L < 2^31 (exhaustive run: 2^31-1 values, 0 mismatches,
    40 CPU-s; here: every boundary x(x-1)-1, x(x-1), x(x-1)+1 plus a stride sample)."""
    L = oracle_lib.Oracle.lib()
    L.orc_isqrt.argtypes = [C.c_long]
    L.orc_isqrt.restype = C.c_uint
    vals = set()
    for x in range(1, 46342):
        for d in (-1, 0, 1):
            v = x * (x - 1) + d
            if 0 < v < 2**31:
                vals.add(v)
    rng = np.random.default_rng(0)
    vals.update(int(v) for v in rng.integers(1, 2**31, size=200000))
    vals.update([1, 2, 3, 4, 5, 6, 7, 255, 256, 65535, 65536, 2**24 - 1, 2**24, 2**31 - 1])
    for v in vals:
        assert L.orc_isqrt(v) == _closed_isqrt(v), v
