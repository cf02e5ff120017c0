"""GPU parity of the floating-point (Scilab) tracking against its NumPy float64 restatement
(oracle/softtrack_oracle.py; parity unpinned by any reference output, see its header).

Tolerance: the correlator sums are double-precision sums of 16000 products evaluated in a different
order and with a different libm sincos (1-2 ulp); through the closed loop the difference stays at the
1e-10 level.  Asserted: block sizes identical (absoluteSample equal to 1e-6 sample), correlator outputs
within 1e-7 relative of the prompt magnitude, NCO frequencies within 1e-6 Hz."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _compare(res, ora, ms):
    for f in ("I_E", "I_P", "I_L", "Q_E", "Q_P", "Q_L"):
        scale = np.abs(ora["I_P"]).mean() + np.abs(ora["Q_P"]).mean()
        assert len(res[f]) == len(ora[f]) == ms
        assert np.max(np.abs(res[f] - ora[f])) <= 1e-7 * scale, (f, np.max(np.abs(res[f] - ora[f])), scale)
    assert np.max(np.abs(res["absoluteSample"] - ora["absoluteSample"])) <= 1e-6
    assert np.max(np.abs(res["carrFreq"] - ora["carrFreq"])) <= 1e-6
    assert np.max(np.abs(res["codeFreq"] - ora["codeFreq"])) <= 1e-6
    for f in ("dllDiscr", "dllDiscrFilt", "pllDiscr", "pllDiscrFilt"):
        assert np.max(np.abs(res[f] - ora[f])) <= 1e-8, f


def test_glonass_acq_prerun_tracking_chain(oracle_lib):
    """acquisition -> preRun -> tracking on a GLONASS record, every stage on the GPU, every stage compared"""
    from oracle import pcps_oracle as po, softtrack_oracle as so
    from gnss_sdr_ru_b200.acquisition import AcquisitionEngine, Settings
    from gnss_sdr_ru_b200.softtrack import SoftTrackingEngine, TrackSettings, preRun
    from gnss_sdr_ru_b200.synth import Sat, make_record

    sats = [Sat(system="glonass", prn=2, cn0_dbhz=50, doppler_hz=1234.0, code_phase_chips=100.7, data_seed=3, data_rate_hz=100.0),
            Sat(system="glonass", prn=-3, cn0_dbhz=47, doppler_hz=-2100.0, code_phase_chips=300.2, data_seed=4, data_rate_hz=100.0)]
    ms = 400
    rec = make_record(sats, 16000 * (ms + 15), seed=5)
    fch = [-3, 0, 2]
    acq_eng = AcquisitionEngine()
    acq = acq_eng.acquisition(rec[: 2 * 16000 * 11], Settings.glonass(acqSatelliteList=fch))
    ts = TrackSettings(msToProcess=ms)
    channel = preRun(acq, ts)
    oacq = po.acquisition(po.to_complex(rec[: 2 * 16000 * 11]), po.AcqSettings.glonass(svList=fch))
    oacqd = {k: [r[k] for r in oacq] for k in ("peakMetric", "carrFreq", "codePhase", "freqChannel")}
    ochannel = so.pre_run(oacqd, so.TrackSettings(msToProcess=ms))
    assert [(c["FCH"], c["codePhase"], c["acquiredFreq"]) for c in channel] == [(c["FCH"], c["codePhase"], c["acquiredFreq"]) for c in ochannel]
    assert [c["FCH"] for c in channel] == [2, -3]
    res = SoftTrackingEngine(handle=acq_eng.h).tracking(rec, channel, ts)
    for r, c in zip(res, ochannel):
        ora = so.tracking(rec, c, so.TrackSettings(msToProcess=ms))
        _compare(r, ora, ms)
        # the loop is locked: prompt power in the in-phase arm, frequency near the truth
        assert np.abs(r["I_P"][-100:]).mean() > 5 * np.abs(r["Q_P"][-100:]).mean()
    assert abs(res[0]["carrFreq"][-1] - (1e6 + 2 * 562500 + 1234.0)) < 20.0


def test_gps_soft_tracking(oracle_lib):
    from oracle import softtrack_oracle as so
    from gnss_sdr_ru_b200.softtrack import SoftTrackingEngine, TrackSettings
    from gnss_sdr_ru_b200.synth import Sat, make_record

    sats = [Sat(prn=7, cn0_dbhz=49, doppler_hz=-1500.0, code_phase_chips=200.0, data_seed=9)]
    ms = 300
    rec = make_record(sats, 16000 * (ms + 3), seed=6)
    # code phase 200 chips at sample 0 -> the code starts (1023-200)/1023*16000 samples later (1-based +1)
    cp = int(round((1023.0 - 200.0) * 16000.0 / 1023.0)) + 1
    channel = [dict(FCH=7, codePhase=cp, acquiredFreq=2.42e6 - 1500.0 + 40.0)]
    ts = TrackSettings.gps(msToProcess=ms)
    res = SoftTrackingEngine().tracking(rec, channel, ts)
    ora = so.tracking(rec, channel[0], so.TrackSettings.gps(msToProcess=ms))
    _compare(res[0], ora, ms)


def test_soft_tracking_record_end(oracle_lib):
    """the record ends before msToProcess code periods: both sides stop at the same block"""
    from oracle import softtrack_oracle as so
    from gnss_sdr_ru_b200.softtrack import SoftTrackingEngine, TrackSettings
    from gnss_sdr_ru_b200.synth import Sat, make_record

    rec = make_record([Sat(system="glonass", prn=0, cn0_dbhz=50, doppler_hz=300.0, code_phase_chips=0.0)], 16000 * 20 + 777, seed=8)
    channel = [dict(FCH=0, codePhase=1, acquiredFreq=1e6 + 300.0)]
    res = SoftTrackingEngine().tracking(rec, channel, TrackSettings(msToProcess=50))
    ora = so.tracking(rec, channel[0], so.TrackSettings(msToProcess=50))
    assert len(res[0]["I_P"]) == len(ora["I_P"]) == 20
    _compare(res[0], ora, 20)
