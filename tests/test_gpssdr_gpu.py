"""GPS-SDR fixed-point acquisition on the GPU against the restatement (oracle/gpssdr_oracle.c, pinned against
the reference's compiled primitives): magnitude, code phase and Doppler bit exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FS, FIF = 2048000.0, 38400.0


def _record(rng, ms, sats, sigma=8.0):  # AGC_BITS 6 (RT/includes/config.h:101): samples within about +-32
    """complex int16 at 2.048 Msps: sum of C/A signals (sv 0-based, amplitude, doppler Hz, code offset samples) + noise"""
    from gnss_sdr_ru_b200 import gpssdr_codes

    chips = gpssdr_codes.prn_gen()
    n = ms * 2048
    t = np.arange(n) / FS
    x = sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    for sv, amp, dopp, off in sats:
        chip_idx = (np.floor((np.arange(n) + off) * 1023.0 / 2048.0)).astype(np.int64) % 1023
        data = np.where((np.arange(n) // (20 * 2048) + sv) % 3 == 0, -1.0, 1.0)  # some 20-ms data bits
        x += amp * chips[chip_idx, sv] * data * np.exp(2j * np.pi * (FIF + dopp) * t + 1j * 0.3 * sv)
    out = np.empty((n, 2), dtype=np.int16)
    out[:, 0] = np.round(x.real)
    out[:, 1] = np.round(x.imag)
    return out


def test_weak_and_strong_match_oracle():
    from gnss_sdr_ru_b200 import gpssdr_codes
    from gnss_sdr_ru_b200.gpssdr_acq import Acquisition
    from oracle import gpssdr_oracle_api as G

    rng = np.random.default_rng(12)
    codes = gpssdr_codes.fft_codes()
    acq = Acquisition(fif=FIF)
    try:
        # ---- weak: 310 ms, three satellites present (one weak), two absent ----
        rec = _record(rng, 310, [(3, 1.2, 1730.0, 517), (17, 0.6, -2210.0, 1201), (30, 0.3, 480.0, 77)])
        svs = [3, 17, 30, 8, 24]
        got = acq.doAcqWeak(rec, svs, -3000, 3000)
        o = G.GpsSdrAcquisition(fif=FIF)
        o.doPrepIF(2, rec)
        for sv, g in zip(svs, got):
            w = o.doAcqWeak(codes[sv], -3000, 3000)
            assert (g["code_phase"], g["doppler"], g["magnitude"]) == (w["code_phase"], w["doppler"], w["magnitude"]), (sv, g, w)
        # the two stronger ones are where they were put: the reported Doppler is lcv*1000 + lcv2*250 + r*25 while DFT row r
        # sits at r*25 - 112.5 Hz (acquisition.cpp:108,551), hence the 112.5 Hz offset
        assert abs(got[0]["doppler"] - 112.5 - 1730) <= 100 and abs(got[1]["doppler"] - 112.5 + 2210) <= 100
        assert got[0]["magnitude"] > 5 * got[3]["magnitude"] and got[1]["magnitude"] > 3 * got[4]["magnitude"]
        # ---- strong: 1 ms ----
        rec1 = _record(rng, 1, [(5, 4.0, 2400.0, 300), (11, 3.0, -3600.0, 1999)])
        svs1 = [5, 11, 2]
        got1 = acq.doAcqStrong(rec1, svs1, -5000, 5000)
        o.doPrepIF(0, rec1)
        for sv, g in zip(svs1, got1):
            w = o.doAcqStrong(codes[sv], -5000, 5000)
            assert (g["code_phase"], g["doppler"], g["magnitude"]) == (w["code_phase"], w["doppler"], w["magnitude"]), (sv, g, w)
        assert got1[0]["code_phase"] == 300 and got1[1]["code_phase"] == 1999
        # ---- large amplitudes: int16 wrap-around inside the unscaled forward FFT must wrap identically ----
        rec2 = _record(rng, 1, [(7, 900.0, 1000.0, 100)], sigma=700.0)
        got2 = acq.doAcqStrong(rec2, [7, 9], -2000, 2000)
        o.doPrepIF(0, rec2)
        for sv, g in zip([7, 9], got2):
            w = o.doAcqStrong(codes[sv], -2000, 2000)
            assert (g["code_phase"], g["doppler"], g["magnitude"]) == (w["code_phase"], w["doppler"], w["magnitude"]), (sv, g, w)
        o.close()
    finally:
        acq.close()


def test_medium_matches_oracle_and_committed_reference_outputs():
    """doAcqMedium on the device: a new object (rows 40-69 zero) and after a 310-ms preparation whose rows the search reads
    (acquisition.cpp:340), against the restatement and against the outputs of the search composed from the reference's own
    compiled primitives (tests/golden/gpssdr_ref_golden.npz)."""
    import os

    from gnss_sdr_ru_b200 import gpssdr_codes
    from gnss_sdr_ru_b200.gpssdr_acq import Acquisition
    from oracle import gpssdr_oracle_api as G

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gpssdr_ref_golden.npz"))
    prior = np.random.default_rng(77).integers(-20, 21, size=(310 * 2048, 2)).astype(np.int16)
    rec = g["medium_rec"]
    codes = gpssdr_codes.fft_codes()
    acq = Acquisition(fif=FIF)
    try:
        for tag, pr in (("fresh", None), ("prior", (2, prior))):
            for (sv, dmin, dmax), want in zip(g["medium_cases"], g[f"medium_{tag}"]):
                got = acq.doAcqMedium(rec, [int(sv)], int(dmin), int(dmax), prior=pr)[0]
                assert (got["code_phase"], got["doppler"], got["magnitude"]) == tuple(int(v) for v in want), (tag, sv, got, want)
                assert got["type"] == 1 and got["success"] == 1
        # a wider search over several satellites in one call, random record, both histories, against the restatement
        rng = np.random.default_rng(3)
        rec2 = _record(rng, 10, [(6, 1.5, -1290.0, 900), (27, 1.0, 3330.0, 64)])
        svs = [6, 27, 1, 15]
        for pr in (None, (2, prior), (1, rec)):
            got = acq.doAcqMedium(rec2, svs, -4000, 4000, prior=pr)
            o = G.GpsSdrAcquisition(fif=FIF)
            if pr is not None:
                o.doPrepIF(pr[0], pr[1])
            o.doPrepIF(1, rec2)
            for sv, gg in zip(svs, got):
                w = o.doAcqMedium(codes[sv], -4000, 4000)
                assert (gg["code_phase"], gg["doppler"], gg["magnitude"]) == (w["code_phase"], w["doppler"], w["magnitude"]), (sv, gg, w)
            o.close()
        with pytest.raises(Exception):
            acq.doAcqMedium(rec2[:100], svs, -4000, 4000)
    finally:
        acq.close()


def test_strong_and_weak_match_committed_reference_compositions():
    """the device against the searches composed from the reference's own compiled primitives (make_gpssdr_golden.py)"""
    from gnss_sdr_ru_b200.gpssdr_acq import Acquisition
    from test_gpssdr_oracle import composed_records

    g, wrec, srec = composed_records()
    acq = Acquisition(fif=FIF)
    try:
        for (sv, dmin, dmax), want in zip(g["strong_cases"], g["strong_ref"]):
            r = acq.doAcqStrong(srec, [int(sv)], int(dmin), int(dmax))[0]
            assert (r["code_phase"], r["doppler"], r["magnitude"]) == tuple(int(v) for v in want), (sv, r, want)
        for (sv, dmin, dmax), want in zip(g["weak_cases"], g["weak_ref"]):
            r = acq.doAcqWeak(wrec, [int(sv)], int(dmin), int(dmax))[0]
            assert (r["code_phase"], r["doppler"], r["magnitude"]) == tuple(int(v) for v in want), (sv, r, want)
    finally:
        acq.close()
