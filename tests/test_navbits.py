"""Navigation-message search: the oracle against known answers on CPU, the CUDA path
(through the C ABI) against the oracle on GPU."""
import numpy as np
import pytest

from oracle import navbits_oracle as O


def _gps_ip(rng, n_ms, first_ms0, amp=900.0, sigma=250.0, invert=False, integer=False):
    """prompt values of one channel: parity-correct subframes whose first TLM bit starts at 0-based ms first_ms0"""
    lead_sub = -(-first_ms0 // 6000)  # whole subframes before the one that starts at first_ms0
    nsub = lead_sub + (n_ms - first_ms0) // 6000 + 2
    bits = O.gps_nav_bits(nsub, rng)
    p0 = 6000 * lead_sub - first_ms0
    sig = np.repeat(2 * bits - 1, 20)[p0 : p0 + n_ms].astype(np.float64)
    if invert:
        sig = -sig
    x = amp * sig + sigma * rng.standard_normal(n_ms)
    return np.round(x).astype(np.int32) if integer else x


def test_parity_routine_accepts_encoded_words_and_rejects_errors():
    rng = np.random.default_rng(1)
    bits = O.gps_nav_bits(3, rng)
    pm = 2 * bits - 1
    for inv in (1, -1):
        for w in range(1, 30):
            nd = inv * pm[30 * w - 2 : 30 * w + 30]
            assert O.navPartyChk(nd) == -nd[1]
            bad = nd.copy()
            bad[5 + (w % 20)] *= -1
            assert O.navPartyChk(bad) == 0
    z = (2 * bits - 1)[28:60].copy()
    z[7] = 0  # a zero "bit" (sign(0) = 0) can never pass
    assert O.navPartyChk(z) == 0


def test_oracle_finds_planted_subframe_and_time_mark():
    rng = np.random.default_rng(2)
    n_ms = 5000 + 6000 + 6000 + 1300
    first0 = 5000 + 1234  # 0-based ms of the first preamble bit after the search offset
    ip = np.stack([_gps_ip(rng, n_ms, first0), _gps_ip(rng, n_ms, first0 + 77, invert=True), 50.0 * rng.standard_normal(n_ms)])
    first, act = O.findPreambles("TT-", ip)
    assert list(first) == [first0 + 1, first0 + 77 + 1, 0] and act == [1, 2]
    # GLONASS: time mark = the 30-bit pattern at 10 ms per bit, in transmission order (tm_bits is stored reversed)
    tm = np.array([-1, 1, 1, -1, 1, -1, -1, 1, -1, -1, -1, -1, 1, -1, 1, -1, 1, 1, 1, -1, 1, 1, -1, -1, -1, 1, 1, 1, 1, 1])
    n2 = 4000
    x = np.repeat(2 * rng.integers(0, 2, size=n2 // 10) - 1, 10).astype(np.float64)
    x[1700:2000] = np.repeat(tm[::-1], 10)
    y = -x
    fs, act2 = O.findTimeMarks("TT", np.stack([x * 500 + 20 * rng.standard_normal(n2), y * 400]))
    assert list(fs) == [1701, 1701] and act2 == [1, 2]


@pytest.mark.gpu
def test_gpu_matches_oracle_gps_and_glonass():
    from gnss_sdr_ru_b200.navbits import NavBitsEngine

    rng = np.random.default_rng(3)
    eng = NavBitsEngine()
    try:
        n_ms = 5000 + 2 * 6000 + 2500
        rows, status = [], ""
        for c in range(10):
            first0 = 5000 + int(rng.integers(0, 5990))
            if c == 3:
                rows.append(60.0 * rng.standard_normal(n_ms))          # noise only: nothing found
            elif c == 5:
                r = _gps_ip(rng, n_ms, first0, sigma=700.0)               # noisy: false preamble hits to reject by parity
                rows.append(r)
            elif c == 7:
                r = _gps_ip(rng, n_ms, first0)
                r[first0 + 3 * 20 : first0 + 3 * 20 + 5] = 0.0            # exact zeros inside the preamble
                rows.append(r)
            else:
                rows.append(_gps_ip(rng, n_ms, first0, invert=bool(c & 1)))
            status += "-" if c == 8 else "T"
        ip = np.stack(rows)
        want, want_act = O.findPreambles(status, ip)
        got, got_act = eng.findPreambles(status, ip)
        assert np.array_equal(got, want) and got_act == want_act
        assert (want > 0).sum() >= 6 and want[3] == 0 and want[8] == 0
        # integer dump values (what gnssb200_track_run leaves in its dump records)
        ipi = np.stack([_gps_ip(rng, n_ms, 5000 + 321 + 40 * c, integer=True) for c in range(4)])
        want, want_act = O.findPreambles("TTTT", ipi)
        got, got_act = eng.findPreambles("TTTT", ipi)
        assert np.array_equal(got, want) and got_act == want_act and (want > 0).all()
        # too short for the search offset / for a second preamble: nothing, no crash
        got, got_act = eng.findPreambles("TT", ip[:2, :4000])
        assert list(got) == [0, 0] and got_act == []
        w2, a2 = O.findPreambles("TT", ip[:2, :9000])
        g2, ga2 = eng.findPreambles("TT", ip[:2, :9000])
        assert np.array_equal(g2, w2) and ga2 == a2
        # GLONASS time marks
        tm = np.array([-1, 1, 1, -1, 1, -1, -1, 1, -1, -1, -1, -1, 1, -1, 1, -1, 1, 1, 1, -1, 1, 1, -1, -1, -1, 1, 1, 1, 1, 1])
        n2 = 7000
        gl = []
        for c in range(6):
            x = np.repeat(2 * rng.integers(0, 2, size=n2 // 10) - 1, 10).astype(np.float64)
            if c != 4:
                p = int(rng.integers(0, n2 - 2300))
                for rep in range(p, n2 - 300, 2000):
                    x[rep : rep + 300] = np.repeat(tm[::-1], 10) * (1 if c & 1 else -1)
            gl.append(x * 300.0 + 120.0 * rng.standard_normal(n2))
        gl = np.stack(gl)
        want, want_act = O.findTimeMarks("TTT-TT", gl)
        got, got_act = eng.findTimeMarks("TTT-TT", gl)
        assert np.array_equal(got, want) and got_act == want_act
        assert (want > 0).sum() == 4
    finally:
        eng.close()


@pytest.mark.gpu
def test_time_marks_from_device_softtrack_layout():
    """strided device input: the I_P field of a [n_ch][ms][13] double buffer (gnssb200_softtrack output layout)"""
    import torch

    from gnss_sdr_ru_b200.navbits import NAV_F64, NavBitsEngine

    rng = np.random.default_rng(4)
    tm = np.array([-1, 1, 1, -1, 1, -1, -1, 1, -1, -1, -1, -1, 1, -1, 1, -1, 1, 1, 1, -1, 1, 1, -1, -1, -1, 1, 1, 1, 1, 1])
    n_ch, n_ms = 3, 3000
    out = rng.standard_normal((n_ch, n_ms, 13)) * 1000.0
    for c in range(n_ch):
        x = np.repeat(2 * rng.integers(0, 2, size=n_ms // 10) - 1, 10).astype(np.float64)
        x[400 + 100 * c : 700 + 100 * c] = np.repeat(tm[::-1], 10)
        out[c, :, 1] = 800.0 * x + 50.0 * rng.standard_normal(n_ms)
    d = torch.from_numpy(out).cuda()
    eng = NavBitsEngine()
    try:
        got, act = eng.findTimeMarks_device(d.data_ptr() + 8, NAV_F64, n_ms * 13 * 8, 13 * 8, n_ch, n_ms)
    finally:
        eng.close()
    want, wact = O.findTimeMarks("TTT", out[:, :, 1])
    assert np.array_equal(got, want) and act == wact and list(want) == [401, 501, 601]


@pytest.mark.gpu
def test_chain_integer_tracking_to_preambles_on_device():
    """The whole integer-receiver chain without leaving the device: a 19 s GPS record carrying parity-correct
    subframes (generated on the GPU), closed-loop tracking (search -> confirm -> pull-in -> tracking), then
    findPreambles reading the prompt values straight from the device dump records (int32 acc[2], 48-byte
    stride).  The result must equal the oracle's on the downloaded records, and the subframe period must be
    visible in them."""
    import ctypes as C

    import torch

    from gnss_sdr_ru_b200 import abi
    from gnss_sdr_ru_b200.lib import check, lib
    from gnss_sdr_ru_b200.navbits import NAV_I32, NavBitsEngine
    from gnss_sdr_ru_b200.receiver import TrackingEngine
    from gnss_sdr_ru_b200.scenarios import TrackScenario, synth_sat_array
    from gnss_sdr_ru_b200.synth import Sat

    rng = np.random.default_rng(21)
    NSB, nblk = 8192, 37110  # 19.0 s
    sats, prns = [], [0] * 12
    for i, (prn, dop) in enumerate(((5, 1210.0), (12, -2890.0), (23, 2240.0), (30, -760.0))):
        bits = O.gps_nav_bits(4, rng).astype(np.uint8)  # 4 subframes, repeated cyclically
        sats.append(Sat(prn=prn, doppler_hz=dop, cn0_dbhz=50.0, code_phase_chips=100.0 + 211.0 * i, data_bits=bits))
        prns[2 * i] = prn
    eng = TrackingEngine(n_streams=1)
    eng.simple_cold_allocate(0, prns)
    for i, s in enumerate(sats):  # start each channel in the Doppler bin of its satellite
        eng.warm_start(0, 2 * i, int(round(s.doppler_hz / 1000.0)))
    eng.upload()
    L = lib()
    n = NSB * nblk
    d_if = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
    arr, nsat = synth_sat_array([TrackScenario(sats=sats, prns=[], n_freq=[])])
    check(L.gnssb200_synth(eng.h, d_if.data_ptr(), n // 2, abi.FMT_PACKED2, 1, n, C.addressof(arr), nsat, 4242, None), "synth")
    cap = 19200
    d_dumps = torch.zeros((12, cap, 48), dtype=torch.uint8, device="cuda")
    d_cnt = torch.zeros(12, dtype=torch.int32, device="cuda")
    eng.run_device(d_if.data_ptr(), n // 2, nblk, NSB, abi.FMT_PACKED2, d_dumps_ptr=d_dumps.data_ptr(), dump_cap=cap, d_count_ptr=d_cnt.data_ptr())
    torch.cuda.synchronize()
    eng.download()
    cnt = d_cnt.cpu().numpy()
    states = [int(eng.rx[0].chan[ch].state) for ch in range(12)]
    assert sum(states[2 * i] == 4 for i in range(4)) >= 3, states  # CHANNEL_TRACKING (a channel may still be pulling in: the reference's bit sync is slow)
    n_ms = int(cnt[[0, 2, 4, 6]].min())
    assert n_ms > 18500
    nav = NavBitsEngine(handle=eng.h)
    status = "".join("T" if (ch % 2 == 0 and ch < 8) else "-" for ch in range(12))
    first, act = nav.findPreambles_device(d_dumps.data_ptr() + 16, NAV_I32, cap * 48, 48, 12, n_ms, status)
    recs = d_dumps.cpu().numpy().view(abi.DUMP_DTYPE).reshape(12, cap)
    ip = recs["acc"][:, :n_ms, 2].astype(np.int32)
    want, wact = O.findPreambles(status, ip)
    assert np.array_equal(first, want) and act == wact
    assert len(act) >= 3 and set(act) <= {1, 3, 5, 7}, (first, act)
    for ch in [c - 1 for c in act]:  # a second preamble one subframe later, where the reference looks for it
        sgn = np.sign(ip[ch])
        pat = np.repeat(np.array([1, -1, -1, -1, 1, -1, 1, 1]), 20)
        k0 = int(first[ch]) - 1
        assert abs(int(np.dot(sgn[k0 + 6000 : k0 + 6160], pat))) > 153


@pytest.mark.gpu
def test_chain_glonass_acquisition_tracking_time_marks_on_device():
    """The GLONASS Scilab chain on the GPU: record synthesised on the device (three frequency channels, 100 sym/s
    strings that end in the 30-symbol time mark), acquisition -> preRun -> floating-point tracking with the
    results left in HBM -> findTimeMarks reading the I_P field of that buffer.  Indices equal the oracle's on the
    downloaded buffer, and consecutive marks are one string (2000 ms) apart."""
    import ctypes as C

    import torch

    from gnss_sdr_ru_b200 import abi
    from gnss_sdr_ru_b200.acquisition import AcquisitionEngine, Settings
    from gnss_sdr_ru_b200.lib import check, lib
    from gnss_sdr_ru_b200.navbits import NAV_F64, NavBitsEngine
    from gnss_sdr_ru_b200.scenarios import TrackScenario, synth_sat_array
    from gnss_sdr_ru_b200.softtrack import SoftTrackingEngine, TrackSettings, preRun
    from gnss_sdr_ru_b200.synth import Sat

    rng = np.random.default_rng(33)
    tm = np.array([-1, 1, 1, -1, 1, -1, -1, 1, -1, -1, -1, -1, 1, -1, 1, -1, 1, 1, 1, -1, 1, 1, -1, -1, -1, 1, 1, 1, 1, 1])
    ms = 4300
    sats = []
    for i, (k, dop) in enumerate(((-4, 1810.0), (1, -2440.0), (5, 630.0))):
        string = np.concatenate([2 * rng.integers(0, 2, size=170) - 1, tm[::-1]])  # 1.7 s of symbols + 0.3 s time mark
        bits = ((1 - np.tile(string, 3)) // 2).astype(np.uint8)
        sats.append(Sat(system="glonass", prn=k, cn0_dbhz=49.0, doppler_hz=dop, code_phase_chips=60.0 + 133.0 * i, data_bits=bits, data_rate_hz=100.0))
    L = lib()
    acq_eng = AcquisitionEngine()
    n = 16000 * (ms + 20)
    d_rec = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
    arr, nsat = synth_sat_array([TrackScenario(sats=sats, prns=[], n_freq=[])])
    check(L.gnssb200_synth(acq_eng.h, d_rec.data_ptr(), 2 * n, abi.FMT_INT8_IQ, 1, n, C.addressof(arr), nsat, 909, None), "synth")
    head = d_rec[: 2 * 16000 * 11].cpu().numpy().view(np.int8)
    acq = acq_eng.acquisition(head, Settings.glonass(acqSatelliteList=[-4, 0, 1, 5]))
    ts = TrackSettings(msToProcess=ms)
    channel = preRun(acq, ts)
    assert sorted(c["FCH"] for c in channel) == [-4, 1, 5]
    n_ch = len(channel)
    d_out = torch.zeros((n_ch, ms, 13), dtype=torch.float64, device="cuda")
    d_done = torch.zeros(n_ch, dtype=torch.int32, device="cuda")
    SoftTrackingEngine(handle=acq_eng.h).tracking_device(d_rec.data_ptr(), n, channel, ts, d_out.data_ptr(), d_done.data_ptr())
    assert d_done.cpu().tolist() == [ms] * n_ch
    nav = NavBitsEngine(handle=acq_eng.h)
    first, act = nav.findTimeMarks_device(d_out.data_ptr() + 8, NAV_F64, ms * 13 * 8, 13 * 8, n_ch, ms)
    ip = d_out.cpu().numpy()[:, :, 1]
    want, wact = O.findTimeMarks("T" * n_ch, ip)
    assert np.array_equal(first, want) and act == wact == [1, 2, 3]
    pat = -np.repeat(tm[::-1], 10)
    for ch in range(n_ch):
        k0 = int(first[ch]) - 1
        assert 1500 <= k0 <= 1900  # the first string's mark starts 1.7 s into the record (minus the code phase the tracking starts at)
        assert abs(int(np.dot(np.sign(ip[ch, k0 + 2000 : k0 + 2300]), pat))) > 290  # and again one string later


@pytest.mark.gpu
def test_chain_gps_scilab_acquisition_tracking_preambles_on_device():
    """The GPS Scilab chain on the GPU: 12.6 s record with parity-correct subframes synthesised on the device,
    acquisition -> preRun -> floating-point tracking (results stay in HBM) -> findPreambles on the I_P field.
    Indices equal the oracle's on the downloaded buffer; the subframe start is where the signal put it."""
    import ctypes as C

    import torch

    from gnss_sdr_ru_b200 import abi
    from gnss_sdr_ru_b200.acquisition import AcquisitionEngine, Settings
    from gnss_sdr_ru_b200.lib import check, lib
    from gnss_sdr_ru_b200.navbits import NAV_F64, NavBitsEngine
    from gnss_sdr_ru_b200.scenarios import TrackScenario, synth_sat_array
    from gnss_sdr_ru_b200.softtrack import SoftTrackingEngine, TrackSettings, preRun
    from gnss_sdr_ru_b200.synth import Sat

    rng = np.random.default_rng(44)
    ms = 12600
    sats = []
    for i, (prn, dop) in enumerate(((6, 2150.0), (21, -3320.0))):
        bits = O.gps_nav_bits(3, rng).astype(np.uint8)
        sats.append(Sat(prn=prn, cn0_dbhz=48.0, doppler_hz=dop, code_phase_chips=250.0 + 301.0 * i, data_bits=bits))
    L = lib()
    acq_eng = AcquisitionEngine()
    n = 16000 * (ms + 20)
    d_rec = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
    arr, nsat = synth_sat_array([TrackScenario(sats=sats, prns=[], n_freq=[])])
    check(L.gnssb200_synth(acq_eng.h, d_rec.data_ptr(), 2 * n, abi.FMT_INT8_IQ, 1, n, C.addressof(arr), nsat, 1212, None), "synth")
    st = Settings.gps(acqSearchBand=14.0, acqCohIntegration=4, acqSatelliteList=[6, 13, 21])
    head = d_rec[: 2 * 16000 * 9].cpu().numpy().view(np.int8)
    acq = acq_eng.acquisition(head, st)
    ts = TrackSettings.gps(msToProcess=ms)
    channel = preRun(acq, ts)
    assert sorted(c["FCH"] for c in channel) == [6, 21]
    n_ch = len(channel)
    d_out = torch.zeros((n_ch, ms, 13), dtype=torch.float64, device="cuda")
    d_done = torch.zeros(n_ch, dtype=torch.int32, device="cuda")
    SoftTrackingEngine(handle=acq_eng.h).tracking_device(d_rec.data_ptr(), n, channel, ts, d_out.data_ptr(), d_done.data_ptr())
    assert d_done.cpu().tolist() == [ms] * n_ch
    nav = NavBitsEngine(handle=acq_eng.h)
    first, act = nav.findPreambles_device(d_out.data_ptr() + 8, NAV_F64, ms * 13 * 8, 13 * 8, n_ch, ms)
    ip = d_out.cpu().numpy()[:, :, 1]
    want, wact = O.findPreambles("T" * n_ch, ip)
    assert np.array_equal(first, want) and act == wact == [1, 2]
    # subframes start every 6000 ms of signal; the first one after the 5000 ms search offset is the second (6000 ms),
    # seen from where the tracking started (the acquired code phase, < 1 ms into the record)
    for ch in range(n_ch):
        assert 5995 <= int(first[ch]) <= 6002, first
