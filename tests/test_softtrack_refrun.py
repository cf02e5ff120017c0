"""The Scilab receiver's loop closure against the reference's OWN saved run: SCI/GLONASS/L1/trackingResults.dat holds
1500 ms of trackResults (six correlator sums, both discriminators, both filtered outputs, carrier and code frequency,
sample position per code period), the settings, acqResults and the channel table of one real GLONASS L1 recording
(postProcessing.sce:143).  The recording itself is not in the repository, so the correlator sums cannot be recomputed,
but everything downstream of them can: oracle/softtrack_oracle.py, fed with the recorded sums, must reproduce every
other recorded series (bit for bit, the phase discriminator to one ulp of atan)."""
import os

import numpy as np
import pytest

from oracle import scilab_save
from oracle import softtrack_oracle as so

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/trunk/GNSS_SOFTWARE_RECEIVERS/POSTPROCESSING_SCILAB_RECEIVERS/GLONASS/L1/trackingResults.dat"


RUNS = {"l1": ("scilab_track_golden.npz", SRC), "l2": ("scilab_track_golden_l2.npz", SRC.replace("/L1/", "/L2/"))}


def _golden(run="l1"):
    return np.load(os.path.join(HERE, "golden", RUNS[run][0]))


def _settings(g, **kw):
    return so.TrackSettings(samplingFreq=float(g["settings_samplingFreq"]), IF=float(g["settings_IF"]), L1_IF_step=float(g["settings_L1_IF_step"]),
                            codeFreqBasis=float(g["settings_codeFreqBasis"]), codeLength=int(g["settings_codeLength"]),
                            skipNumberOfSamples=int(g["settings_skipNumberOfBytes"]),  # that run's file position was counted in samples
                            msToProcess=int(g["settings_msToProcess"]), numberOfChannels=int(g["settings_numberOfChannels"]),
                            dllDampingRatio=float(g["settings_dllDampingRatio"]), dllNoiseBandwidth=float(g["settings_dllNoiseBandwidth"]),
                            dllCorrelatorSpacing=float(g["settings_dllCorrelatorSpacing"]), pllNoiseBandwidth=float(g["settings_pllNoiseBandwidth"]),
                            fllNoiseBandwidth=float(g["settings_fllNoiseBandwidth"]), **kw)


@pytest.mark.skipif(not os.path.exists(SRC), reason="reference tree not present")
@pytest.mark.parametrize("run", ["l1", "l2"])
def test_committed_fixture_is_the_reference_file(run):
    import sys

    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_scilab_track_golden as M

    SRC = RUNS[run][1]
    live, g = M.extract(SRC), _golden(run)
    assert sorted(live) == sorted(g.files)
    for k in g.files:
        assert np.array_equal(live[k], g[k], equal_nan=True) if g[k].dtype.kind == "f" else np.array_equal(live[k], g[k]), k
    v = scilab_save.load(SRC)
    assert sorted(v) == ["acqResults", "channel", "settings", "trackRdsults"]  # sic: the stored name
    assert str(v["settings"]["fileName"].ravel()[0]).endswith("FFF005.DAT")


def test_prerun_channel_table_equals_saved_one():
    """preRun.sci:66-81 on the saved acqResults gives the saved channel table (one detected signal, FCH -4)"""
    g = _golden()
    acq = dict(carrFreq=g["acq_carrFreq"], codePhase=g["acq_codePhase"], peakMetric=g["acq_peakMetric"], freqChannel=g["acq_freqChannel"])
    ch = so.pre_run(acq, _settings(g))
    assert len(ch) == 1 and list(g["track_status"]) == ["T", "-"]
    assert (ch[0]["FCH"], ch[0]["acquiredFreq"], ch[0]["codePhase"]) == (g["channel_FCH"][0], g["channel_acquiredFreq"][0], g["channel_codePhase"][0])
    assert ch[0]["SVN"] == g["channel_SVN"][0]
    assert g["channel_acquiredFreq"][1] == 0 and g["channel_codePhase"][1] == 0


@pytest.mark.parametrize("run", ["l1", "l2"])
def test_loop_closure_reproduces_saved_run_bit_for_bit(run):
    """l1: SCI/GLONASS/L1/trackingResults.dat (IF 1 MHz, channel -4 at -1248.5 kHz); l2: SCI/GLONASS/L2/trackingResults.dat
    (IF 2 MHz, 437.5 kHz channel step, signal at +248.5 kHz): two independent runs of the same loop"""
    g = _golden(run)
    rec = {f: g["track_" + f] for f in ("I_E", "I_P", "I_L", "Q_E", "Q_P", "Q_L")}
    channel = dict(FCH=int(g["channel_FCH"][0]), acquiredFreq=float(g["channel_acquiredFreq"][0]), codePhase=int(g["channel_codePhase"][0]))
    # the saved run predates the code-aiding term and the remainder correction of absoluteSample (tracking.sci:366, :379)
    r = so.replay(rec, channel, _settings(g, codeAiding=False, absSampleRemCorr=False))
    n = int(g["settings_msToProcess"])
    assert n == 1500 and all(len(r[f]) == n for f in r)
    # identical expressions in IEEE double: the DLL side and both NCO frequencies come out bit for bit ...
    for f in ("dllDiscr", "dllDiscrFilt", "codeFreq"):
        assert np.array_equal(r[f], g["track_" + f]), f
    d = np.abs(r["carrFreq"] - g["track_carrFreq"])  # basis + pllDiscrFilt: exact on l1, one ulp on l2
    assert d.max() <= np.spacing(np.abs(g["track_carrFreq"]).max()) and (run != "l1" or d.max() == 0)
    # ... the phase discriminator to the last bit of atan() of this libm against Scilab's, its filter output (a running
    # sum of 1500 of those times k1 = 69) to 1e-13
    assert np.abs(r["pllDiscr"] - g["track_pllDiscr"]).max() <= 2 ** -53
    assert np.abs(r["pllDiscrFilt"] - g["track_pllDiscrFilt"]).max() <= 1e-13
    # block sizes and the code-phase remainder: the file position after every code period, 1500 exact integers
    assert np.array_equal(r["absoluteSample"], g["track_absoluteSample"])
    assert {15999, 16000} <= set(r["blksize"]) <= {15999, 16000, 16001}
    # the loops did something: the carrier moved by tens of Hz and settled, the prompt arm holds the power
    assert 5 < np.ptp(g["track_carrFreq"]) < 200
    assert np.mean(g["track_I_P"][500:] ** 2) > 10 * np.mean(g["track_Q_P"][500:] ** 2)


def test_todays_variants_differ_only_where_the_source_says():
    """with today's lines (code aiding :367, remainder correction :384) the carrier side is untouched, the code frequency
    gains the aiding term and the sample position its fractional correction"""
    g = _golden()
    rec = {f: g["track_" + f] for f in ("I_E", "I_P", "I_L", "Q_E", "Q_P", "Q_L")}
    channel = dict(FCH=int(g["channel_FCH"][0]), acquiredFreq=float(g["channel_acquiredFreq"][0]), codePhase=int(g["channel_codePhase"][0]))
    s = _settings(g)
    r = so.replay(rec, channel, s)
    assert np.array_equal(r["carrFreq"], g["track_carrFreq"]) and np.array_equal(r["dllDiscrFilt"], g["track_dllDiscrFilt"])
    aid = (r["carrFreq"] - (s.IF + s.L1_IF_step * channel["FCH"])) / ((s.GLONASS_zero_channel + channel["FCH"] * s.L1_IF_step) / s.codeFreqBasis)
    assert np.allclose(r["codeFreq"] - g["track_codeFreq"], aid, rtol=0, atol=1e-9) and np.all(np.abs(aid) < 1.0)


def test_time_mark_is_found_in_the_recorded_prompt_series():
    """findTimeMarks (restated, SCI/GLONASS/L1/findTimeMarks.sci:25-66) on the prompt values the reference recorded from a
    real satellite: the 30-chip GLONASS time mark is there exactly once in these 1.5 s, on the 10-ms bit grid of the data,
    whichever sign the PLL locked with."""
    from oracle import navbits_oracle as nb

    g = _golden()
    ip = g["track_I_P"]
    first, active = nb.findTimeMarks(["T"], [ip])
    first_neg, _ = nb.findTimeMarks(["T"], [-ip])
    assert list(active) == [1] and first[0] == first_neg[0] == 433
    edges = np.nonzero(np.diff(np.sign(ip[300:])))[0] + 300 + 2  # 1-based index of the first millisecond after a sign change
    assert np.all(edges % 10 == first[0] % 10)
