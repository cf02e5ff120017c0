"""Development check (CPU, NumPy) of the half-chip-segment partition used by track_seg.cuh: the per-thread segment
ownership, the float64 start-sample formula and the dump hand-over reproduce the per-sample closed forms
 on random block parameters.  Not product code; the GPU kernel is checked against the oracle
in tests/test_track_gpu.py."""
import numpy as np


def brute(v, kph0, kinc, hc0, w1, stale_idx, tbl, n):
    i = np.arange(n, dtype=np.uint64)
    wb = ((np.uint64(kph0) + i * np.uint64(kinc)) >> np.uint64(32)).astype(np.int64)
    inA = wb < w1
    rel = wb - w1
    hh = np.where(inA, hc0 + wb, np.where(rel == 0, stale_idx, rel))
    bits = tbl[hh]
    return (bits[inA] * v[inA]).sum(), (bits[~inA] * v[~inA]).sum()


def seg_start(m, kph0, kinc):
    x = np.float64(m) * 4294967296.0 + (-np.float64(kph0) - 0.5)
    return int(np.float64(x) * (np.float64(1.0) / np.float64(kinc))) + 1


def seg_model(v, kph0, kinc, hc0, w1, stale_idx, tbl, n, NT, H, SLOTS=8):
    wtot = (kph0 + n * kinc) >> 32
    NF = wtot - 1 if wtot > 0 else 0
    A = B = 0
    alias = tbl.copy()
    alias[0] = tbl[stale_idx]
    for tid in range(NT):
        if tid < NT - 1:
            first = tid * H
            m0 = 1 + first
            nv = min(NF - first, H) if NF > first else 0
            if m0 < w1 < m0 + nv:
                nv = w1 - m0
        else:
            m0, nv = 1, 0
            if w1 >= 2:
                ts = (w1 - 2) // H
                first = ts * H
                end = 1 + first + (min(NF - first, H) if NF > first else 0)
                if ts < NT - 1 and w1 < end:
                    m0, nv = w1, end - w1
        if nv == 0:
            m0 = 1
        clsB = m0 >= w1
        if clsB:
            rel = m0 - w1
            src, base = (alias, 0) if rel == 0 else (tbl, rel)
        else:
            src, base = tbl, hc0 + m0
        s = seg_start(m0, kph0, kinc)
        exact = -((-(m0 * 2**32 - kph0)) // kinc)
        assert s == exact, (s, exact)
        ks = (kph0 + s * kinc) % 2**32
        assert ks < kinc
        acc = 0
        for j in range(H):
            u = ks + (SLOTS - 1) * kinc
            eight = u < 2**32
            cnt = SLOTS if eight else SLOTS - 1
            ks = (u + (kinc if eight else 0)) % 2**32
            if j < nv:
                assert s + cnt <= n
                acc += src[base + j] * v[s : s + cnt].sum()
            s += cnt
        if clsB:
            B += acc
        else:
            A += acc
    OWNED = (NT - 1) * H
    head_end = n if wtot == 0 else min(seg_start(1, kph0, kinc), n)
    mt = 1 + min(NF, OWNED)
    tail_start = n if mt > wtot else min(seg_start(mt, kph0, kinc), n)
    for i in list(range(head_end)) + list(range(tail_start, n)):
        wb = (kph0 + i * kinc) >> 32
        if wb < w1:
            A += tbl[hc0 + wb] * v[i]
        else:
            rel = wb - w1
            B += tbl[stale_idx if rel == 0 else rel] * v[i]
    return A, B


def main():
    rng = np.random.default_rng(1)
    tbl = rng.integers(-1, 2, 4096)
    for it in range(3000):
        n = int(rng.choice([8192, 8192, 8192, 4000, 4096, 6000, 512, 64]))
        kinc = int(rng.integers(2**29, 2**32 // 7 + 1)) if it % 3 else 549218880 + int(rng.integers(-3000, 3000))
        kph0 = int(rng.integers(0, 2**32))
        wtot = (kph0 + n * kinc) >> 32
        hc0 = int(rng.integers(0, 2046))
        mode = it % 4
        if mode == 0:
            w1 = 2046 - hc0 if hc0 < 2045 else 1
        else:
            w1 = int(rng.integers(1, max(2, wtot + 3)))
        stale_idx = hc0 + w1
        v = rng.integers(-9, 10, n)
        for NT, H in ((96, 11), (192, 6), (384, 3)):
            a = brute(v, kph0, kinc, hc0, w1, stale_idx, tbl, n)
            b = seg_model(v, kph0, kinc, hc0, w1, stale_idx, tbl, n, NT, H)
            assert a == b, (it, n, kinc, kph0, hc0, w1, NT, H, a, b)
        # 15 or 16 samples per half chip (GLONASS channels): H = 6 / 3 / 2 segments per thread
        kinc16 = int(rng.integers(2**28, 2**32 // 15 + 1)) if it % 3 else 274340960 + int(rng.integers(-2000, 2000))
        wtot16 = (kph0 + n * kinc16) >> 32
        w1b = 1022 - hc0 % 1022 if mode == 0 else int(rng.integers(1, max(2, wtot16 + 3)))
        hc0b = hc0 % 1022
        for NT, H in ((96, 6), (192, 3), (384, 2)):
            a = brute(v, kph0, kinc16, hc0b, w1b, hc0b + w1b, tbl, n)
            b = seg_model(v, kph0, kinc16, hc0b, w1b, hc0b + w1b, tbl, n, NT, H, SLOTS=16)
            assert a == b, ("16", it, n, kinc16, kph0, hc0b, w1b, NT, H, a, b)
    print("segment partition model: ok")


if __name__ == "__main__":
    main()
