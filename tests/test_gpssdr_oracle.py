"""GPS-SDR fixed-point acquisition, CPU side: the restatement's primitives against the
reference's own code compiled in place, the regenerated code table against the reference's header."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import gpssdr_oracle_api as G

needs_ref = pytest.mark.skipif(not os.path.isdir(G.RT), reason="reference tree not present")


def test_code_table_hash_is_stable():
    from gnss_sdr_ru_b200 import gpssdr_codes

    t = gpssdr_codes.fft_codes()
    assert t.shape == (51, 2048, 2) and t.dtype == np.int16
    assert int(np.hypot(t[..., 0].astype(float), t[..., 1].astype(float)).max() + 0.5) == 512  # scaled to 9 signed bits
    assert tuple(t[0, 0]) == (11, 0) and tuple(t[0, 1]) == (-90, -89)  # first entries of prn_codes.h
    assert gpssdr_codes.table_sha256() == "be9e4e21b1888d1861f858876b84d322b35caf822cdd708b6e67d9c1fa32fc24"


@needs_ref
def test_code_table_equals_reference_header():
    from gnss_sdr_ru_b200 import gpssdr_codes

    txt = open(os.path.join(G.RT, "accessories", "prn_codes.h")).read()
    body = txt[txt.index("{") + 1 : txt.index("}")]
    ref = np.array([int(x) for x in re.split(r"[,\s]+", body.strip()) if x], dtype=np.int16)
    assert ref.size == 208896
    assert np.array_equal(gpssdr_codes.fft_codes().reshape(-1), ref)


@needs_ref
def test_primitives_match_compiled_reference():
    L, R = G.lib(), G.ref()
    rng = np.random.default_rng(5)
    R1 = np.zeros(16, dtype=np.int32)
    R2 = np.array([0, 0, 0, 0, 0, 0, 0, 1, 0, 1, 0, 1, 1, 1, 1, 1], dtype=np.int32)
    Rall = np.ones(16, dtype=np.int32)
    for amp in (40, 900, 20000):  # small, typical, wrapping (int16 overflow must wrap identically)
        for Rp in (R1, R2, Rall):
            for inv in (0, 1):
                for n in (2048, 32):
                    x = rng.integers(-amp, amp + 1, size=(n, 2)).astype(np.int16)
                    a, b = x.copy(), x.copy()
                    L.gso_fft(a.ctypes.data, n, Rp.ctypes.data, inv, 1)
                    R.gsr_fft(b.ctypes.data, n, Rp.ctypes.data, inv, 1)
                    assert np.array_equal(a, b), (amp, inv, n)
    for shift in (9, 10, 14):
        A = rng.integers(-23000, 23001, size=(5000, 2)).astype(np.int16)  # products stay inside int32 (beyond that the C code overflows)
        B = rng.integers(-23000, 23001, size=(5000, 2)).astype(np.int16)
        c1, c2 = np.zeros_like(A), np.zeros_like(A)
        L.gso_cmulsc(A.ctypes.data, B.ctypes.data, c1.ctypes.data, 5000, shift)
        Ar, Br = A.copy(), B.copy()
        R.gsr_cmulsc(Ar.ctypes.data, Br.ctypes.data, c2.ctypes.data, 5000, shift)
        assert np.array_equal(c1, c2)
        a2 = A.copy()
        R.gsr_cmuls(a2.ctypes.data, Br.ctypes.data, 5000, shift)  # in-place form
        assert np.array_equal(c1, a2)
    for f in (-112.5, -12.5, 87.5):
        w1, w2 = np.zeros((10, 4), np.int16), np.zeros((10, 4), np.int16)
        L.gso_wipeoff_gen(w1.ctypes.data, f, 1000.0, 10)
        R.gsr_wipeoff_gen(w2.ctypes.data, f, 1000.0, 10)
        assert np.array_equal(w1, w2)
        d = rng.integers(-3000, 3000, size=(10, 2)).astype(np.int16)
        i1, q1, i2, q2 = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        L.gso_cacc(d.ctypes.data, w1.ctypes.data, 10, C.byref(i1), C.byref(q1))
        dr = d.copy()
        R.gsr_cacc(dr.ctypes.data, w2.ctypes.data, 10, C.byref(i2), C.byref(q2))
        assert (i1.value, q1.value) == (i2.value, q2.value)
    for f in (-38400.0, -38650.0, -38900.0, -39150.0):
        s1, s2 = np.zeros((20480, 2), np.int16), np.zeros((20480, 2), np.int16)
        L.gso_sine_gen(s1.ctypes.data, f, 2048000.0, 20480)
        R.gsr_sine_gen(s2.ctypes.data, f, 2048000.0, 20480)
        assert np.array_equal(s1, s2)
    m = rng.integers(-180, 181, size=(4096, 2)).astype(np.int16)
    m1, m2 = m.copy(), m.copy()
    L.gso_cmag(m1.ctypes.data, 4096)
    R.gsr_cmag(m2.ctypes.data, 4096)
    assert np.array_equal(m1, m2)
    p = m1.view(np.int32).reshape(-1)
    i1, g1, i2, g2 = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    L.gso_max(p.ctypes.data, C.byref(i1), C.byref(g1), p.size)
    pr = p.copy()
    R.gsr_max(pr.ctypes.data, C.byref(i2), C.byref(g2), p.size)
    assert (i1.value, g1.value) == (i2.value, g2.value) == (int(np.argmax(p)), int(p.max()))


def test_primitives_match_committed_reference_outputs():
    """same pin without the reference tree: outputs of the reference's compiled primitives committed as
    tests/golden/gpssdr_ref_golden.npz (made by tests/golden/make_gpssdr_golden.py)"""
    L = G.lib()
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gpssdr_ref_golden.npz"))
    pats = {"R1": np.zeros(16, np.int32), "R2": np.array([0, 0, 0, 0, 0, 0, 0, 1, 0, 1, 0, 1, 1, 1, 1, 1], np.int32), "Rall": np.ones(16, np.int32)}
    n = 0
    for name, Rp in pats.items():
        for inv in (0, 1):
            for amp in (60, 20000):
                x = g[f"fft_{name}_{inv}_{amp}_in"].copy()
                L.gso_fft(x.ctypes.data, 2048, Rp.ctypes.data, inv, 1)
                assert np.array_equal(x, g[f"fft_{name}_{inv}_{amp}_out"])
                n += 1
    for shift in (9, 10, 14):
        A, B = g[f"cmulsc_{shift}_a"].copy(), g[f"cmulsc_{shift}_b"].copy()
        c = np.zeros_like(A)
        L.gso_cmulsc(A.ctypes.data, B.ctypes.data, c.ctypes.data, A.shape[0], shift)
        assert np.array_equal(c, g[f"cmulsc_{shift}_c"])
    for k, f in enumerate((-38400.0, -38650.0, -38900.0, -39150.0)):
        s = np.zeros((20480, 2), np.int16)
        L.gso_sine_gen(s.ctypes.data, f, 2048000.0, 20480)
        assert np.array_equal(s, g[f"sine_{k}"])  # float-phase sinf/cosf of this libm (same image as the reference build)
    dft = np.zeros((10, 10, 4), np.int16)
    for r in range(10):
        L.gso_wipeoff_gen(dft[r].ctypes.data, float(np.float32(r) * 25.0 - 112.5), 1000.0, 10)
    assert np.array_equal(dft, g["dft_rows"])
    d, want = g["cacc_in"], g["cacc_out"]
    for i in range(d.shape[0]):
        for r in range(10):
            a, b = C.c_int32(), C.c_int32()
            di = np.ascontiguousarray(d[i])
            L.gso_cacc(di.ctypes.data, dft[r].ctypes.data, 10, C.byref(a), C.byref(b))
            assert (a.value, b.value) == tuple(want[i, r])
    assert n == 12


def test_restated_pipeline_finds_planted_satellites():
    """doPrepIF + doAcqStrong / doAcqWeak of the restatement on a synthetic 2.048 Msps record: the planted
    satellites come out at their code phase and Doppler, absent ones stay far below."""
    from gnss_sdr_ru_b200 import gpssdr_codes

    rng = np.random.default_rng(8)
    chips = gpssdr_codes.prn_gen()
    codes = gpssdr_codes.fft_codes()

    def record(ms, sats):
        n = ms * 2048
        t = np.arange(n) / 2048000.0
        x = 8.0 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
        for sv, amp, dopp, off in sats:
            ci = (np.floor((np.arange(n) + off) * 1023.0 / 2048.0)).astype(np.int64) % 1023
            x += amp * chips[ci, sv] * np.exp(2j * np.pi * (38400.0 + dopp) * t)
        out = np.empty((n, 2), dtype=np.int16)
        out[:, 0], out[:, 1] = np.round(x.real), np.round(x.imag)
        return out

    o = G.GpsSdrAcquisition(fif=38400.0)
    try:
        rec1 = record(1, [(4, 4.0, 1500.0, 700)])
        o.doPrepIF(0, rec1)
        s = o.doAcqStrong(codes[4], -4000, 4000)
        a = o.doAcqStrong(codes[9], -4000, 4000)
        assert s["code_phase"] == 700 and abs(s["doppler"] - 1500) <= 250 and s["magnitude"] > 4 * a["magnitude"]
        rec = record(310, [(4, 0.5, 1500.0, 700)])
        o.doPrepIF(2, rec)
        w = o.doAcqWeak(codes[4], 1000, 2000)
        b = o.doAcqWeak(codes[9], 1000, 2000)
        # doAcqWeak reports the index of the power matrix column: (2048 - offset) % 2048 like the strong search's raw index
        assert w["code_phase"] in (2048 - 700, 2048 - 701, 2048 - 699) and abs(w["doppler"] - 112.5 - 1500) <= 50
        assert w["magnitude"] > 5 * b["magnitude"]
    finally:
        o.close()


def _medium_inputs():
    import hashlib

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gpssdr_ref_golden.npz"))
    prior = np.random.default_rng(77).integers(-20, 21, size=(310 * 2048, 2)).astype(np.int16)
    assert hashlib.sha256(prior.tobytes()).digest() == g["medium_prior_sha256"].tobytes(), "numpy changed its integer stream"
    return g, prior


def test_medium_matches_committed_reference_outputs():
    """doPrepIF(1) + doAcqMedium of the restatement against the same search composed from the REFERENCE's compiled
    primitives (tests/gpssdr_refpipe.py, outputs committed by make_gpssdr_golden.py): on a new object, where rows
    40-69 are zero, and after a 310-ms preparation, whose rows 40-69 the medium search reads (acquisition.cpp:340)."""
    from gnss_sdr_ru_b200 import gpssdr_codes

    g, prior = _medium_inputs()
    codes = gpssdr_codes.fft_codes()
    for tag in ("fresh", "prior"):
        o = G.GpsSdrAcquisition(fif=38400.0)
        try:
            if tag == "prior":
                o.doPrepIF(2, prior)
            o.doPrepIF(1, g["medium_rec"])
            for (sv, dmin, dmax), want in zip(g["medium_cases"], g[f"medium_{tag}"]):
                r = o.doAcqMedium(codes[sv], int(dmin), int(dmax))
                assert (r["code_phase"], r["doppler"], r["magnitude"]) == tuple(int(v) for v in want), (tag, sv, r, want)
        finally:
            o.close()
    # the planted satellites: sv 4 at +1040 Hz / offset 700, sv 20 at -2480 Hz / offset 1500 (index = 2048 - offset);
    # the absent sv 9 is decided by the stale rows once they hold something
    assert g["medium_fresh"][0][0] == 2048 - 700 and g["medium_fresh"][1][0] == 2048 - 1500 + 1
    assert tuple(g["medium_prior"][2]) != tuple(g["medium_fresh"][2]) and g["medium_prior"][2][1] % 1000 >= 500


@needs_ref
def test_medium_matches_reference_composition_live():
    """the same comparison with the reference's primitives called now (one kHz bin, both histories)"""
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import gpssdr_refpipe as RP
    from gnss_sdr_ru_b200 import gpssdr_codes

    _, prior = _medium_inputs()
    codes = gpssdr_codes.fft_codes()
    rng = np.random.default_rng(31)
    rec = rng.integers(-30, 31, size=(10 * 2048, 2)).astype(np.int16)
    ra = RP.RefAcquisition(RP.Prims(G.ref(), "gsr_"), n_rows=70)
    o = G.GpsSdrAcquisition(fif=38400.0)
    try:
        for with_prior in (False, True):
            if with_prior:
                ra.prep(2, prior, max_rows=70)
                o.doPrepIF(2, prior)
            ra.prep(1, rec)
            o.doPrepIF(1, rec)
            want = RP.RefAcquisition.pick(ra.medium_cells(codes[13], -1000, -1000))
            assert o.doAcqMedium(codes[13], -1000, -1000) == want, (with_prior, want)
    finally:
        o.close()


WEAK_REC = dict(seed=11, ms=310, sats=[(4, 1, 1480, 700)], noise=12)
STRONG_REC = dict(seed=12, ms=1, sats=[(7, 4, -2300, 300), (22, 3, 3900, 1777)], noise=12)


def composed_records():
    """the integer-only records of make_gpssdr_golden.py, checked against the hashes it stored"""
    import hashlib
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import gpssdr_refpipe as RP

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gpssdr_ref_golden.npz"))
    wrec, srec = RP.int_record(**WEAK_REC), RP.int_record(**STRONG_REC)
    assert hashlib.sha256(wrec.tobytes()).digest() == g["weak_rec_sha256"].tobytes()
    assert hashlib.sha256(srec.tobytes()).digest() == g["strong_rec_sha256"].tobytes()
    return g, wrec, srec


def test_strong_and_weak_match_committed_reference_compositions():
    """doPrepIF + doAcqStrong / doAcqWeak of the restatement against the same searches composed from the REFERENCE's
    compiled primitives (every (kHz bin, offset, alignment) cell through gsr_cmulsc / gsr_fft / gsr_cacc / gsr_cmag /
    gsr_max; outputs committed by make_gpssdr_golden.py)"""
    from gnss_sdr_ru_b200 import gpssdr_codes

    g, wrec, srec = composed_records()
    codes = gpssdr_codes.fft_codes()
    o = G.GpsSdrAcquisition(fif=38400.0)
    try:
        o.doPrepIF(0, srec)
        for (sv, dmin, dmax), want in zip(g["strong_cases"], g["strong_ref"]):
            r = o.doAcqStrong(codes[sv], int(dmin), int(dmax))
            assert (r["code_phase"], r["doppler"], r["magnitude"]) == tuple(int(v) for v in want), (sv, r, want)
        o.doPrepIF(2, wrec)
        for (sv, dmin, dmax), want in zip(g["weak_cases"], g["weak_ref"]):
            r = o.doAcqWeak(codes[sv], int(dmin), int(dmax))
            assert (r["code_phase"], r["doppler"], r["magnitude"]) == tuple(int(v) for v in want), (sv, r, want)
    finally:
        o.close()
    assert abs(int(g["strong_ref"][0][0]) - 300) <= 1 and abs(int(g["strong_ref"][1][0]) - 1777) <= 1
    assert int(g["weak_ref"][0][2]) > 10 * int(g["weak_ref"][1][2])


@needs_ref
def test_strong_matches_reference_composition_live():
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import gpssdr_refpipe as RP
    from gnss_sdr_ru_b200 import gpssdr_codes

    codes = gpssdr_codes.fft_codes()
    rec = RP.int_record(99, 1, [(15, 5, 640, 1234)], noise=700)  # large noise: the unscaled forward FFT wraps in int16
    ra = RP.RefAcquisition(RP.Prims(G.ref(), "gsr_"), n_rows=4)
    ra.prep(0, rec)
    o = G.GpsSdrAcquisition(fif=38400.0)
    try:
        o.doPrepIF(0, rec)
        for sv in (15, 2):
            assert o.doAcqStrong(codes[sv], -3000, 3000) == RP.RefAcquisition.pick_strong(ra.strong_cells(codes[sv], -3000, 3000))
    finally:
        o.close()
