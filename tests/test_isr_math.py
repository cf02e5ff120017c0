"""Closed forms used by the device ISR (csrc/isr_device.cuh) vs the oracle's literal restatement of the
reference's integer helpers (which is pinned against the compiled reference)."""
import ctypes as C
import math

import numpy as np
import pytest


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


@pytest.mark.gpu
def test_device_isr_helpers_vs_reference_helpers(oracle_lib):
    """The branch-free device forms of fix_atan2 / sqrt_newton / the DLL division (csrc/isr_device.cuh),
    evaluated on the GPU through the C ABI, vs the oracle's literal restatement of osgpsisr.c:148-231
    (itself pinned against the compiled reference): bit exact on boundary and random arguments."""
    from gnss_sdr_ru_b200 import lib

    L = lib.lib()
    O = oracle_lib.Oracle.lib()
    O.orc_isqrt.argtypes = [C.c_long]
    O.orc_isqrt.restype = C.c_uint
    O.orc_atan2.argtypes = [C.c_long, C.c_long]
    O.orc_atan2.restype = C.c_long
    rng = np.random.default_rng(11)
    n = 400_000
    # atan2 operands: what the ISR passes (abs < 2^23 for the FLL pair, shorts for the PLL pair), edges, wide values
    y = rng.integers(-(1 << 23), 1 << 23, size=n, dtype=np.int64)
    x = rng.integers(-(1 << 23), 1 << 23, size=n, dtype=np.int64)
    y[: n // 4] = rng.integers(-32768, 32768, size=n // 4)
    x[: n // 4] = rng.integers(0, 32769, size=n // 4)
    edge = np.array([0, 1, -1, 2, -2, 3, 255, 256, 32767, -32768, 32768, (1 << 23) - 1, -(1 << 23), (1 << 30) - 1, -(1 << 30) + 1,
                     (1 << 30), (1 << 31) - 1, -(1 << 31) + 1], dtype=np.int64)
    ey, ex = np.meshgrid(edge, edge)
    k = ey.size
    y[n // 4 : n // 4 + k] = ey.ravel()
    x[n // 4 : n // 4 + k] = ex.ravel()
    # equal / nearly equal magnitudes (the case split of fix_atan2)
    m = rng.integers(1, 1 << 22, size=1000)
    for j, (sy, sx, d) in enumerate([(1, 1, 0), (1, -1, 0), (-1, 1, 0), (-1, -1, 0), (1, 1, 1), (1, -1, 1), (-1, 1, 1), (-1, -1, 1)]):
        lo = n // 2 + j * 1000
        y[lo : lo + 1000] = sy * (m + d)
        x[lo : lo + 1000] = sx * m
    # sqrt arguments: every boundary x(x-1)+{-1,0,1} once, random values, zero / negative / >= 2^31
    Ls = rng.integers(1, 1 << 31, size=n, dtype=np.int64)
    xs = np.arange(1, 46342, dtype=np.int64)
    b = np.concatenate([xs * (xs - 1) - 1, xs * (xs - 1), xs * (xs - 1) + 1])
    b = b[(b > 0) & (b < (1 << 31))]
    Ls[: b.size] = b
    Ls[b.size : b.size + 8] = [0, -5, 1, 2, (1 << 31) - 1, 1 << 31, (1 << 32) + 12345, (1 << 33) + 77]
    # DLL division
    num = rng.integers(-(1 << 30) + 1, 1 << 30, size=n, dtype=np.int64)
    den = rng.integers(1, 1 << 20, size=n, dtype=np.int64)
    se = rng.integers(1, 46342, size=n // 2)
    sl = rng.integers(1, 46342, size=n // 2)
    num[: n // 2] = 8192 * (se - sl)
    den[: n // 2] = se + sl
    y32, x32, num32, den32 = (np.ascontiguousarray(a.astype(np.int32)) for a in (y, x, num, den))
    at = np.zeros(n, np.int32)
    sq = np.zeros(n, np.uint32)
    dv = np.zeros(n, np.int32)
    h = L.gnssb200_open(0, None)
    assert h
    try:
        lib.check(L.gnssb200_isr_math_eval(h, n, y32.ctypes.data, x32.ctypes.data, at.ctypes.data, Ls.ctypes.data, sq.ctypes.data,
                                           num32.ctypes.data, den32.ctypes.data, dv.ctypes.data), "isr_math_eval")
    finally:
        L.gnssb200_close(h)
    want_dv = np.where(num >= 0, num // den, -((-num) // den))
    assert np.array_equal(dv.astype(np.int64), want_dv)
    step = 7  # the oracle helpers are called one by one through ctypes: every 7th random value, all edges
    idx = np.unique(np.concatenate([np.arange(0, n, step), np.arange(n // 4, n // 4 + k), np.arange(n // 2, n // 2 + 8000)]))
    for i in idx:
        assert int(at[i]) == O.orc_atan2(int(y32[i]), int(x32[i])), (int(y32[i]), int(x32[i]))
    idx = np.unique(np.concatenate([np.arange(0, n, step), np.arange(0, b.size + 8)]))
    for i in idx:
        assert int(sq[i]) == O.orc_isqrt(int(Ls[i])), int(Ls[i])
