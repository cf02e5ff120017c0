"""The NumPy restatement of acquisition.sci: regression against its committed fixture and checks of the
rules it restates (bin grid, sampled code, exclusion ranges).  Parity of this path is unpinned by any
reference output (oracle/pcps_oracle.py header)."""
import os

import numpy as np

from gnss_sdr_ru_b200.synth import unpack2

HERE = os.path.dirname(os.path.abspath(__file__))


def test_grid_and_code_rules(oracle_lib):
    from oracle import pcps_oracle as po

    g = po.AcqSettings.glonass()
    assert po.samples_per_code(g) == 16000 and po.num_bins(g) == 121
    assert po.bin_freq(g, -7, 1) == 1e6 - 7 * 562500 - 6000 and po.bin_freq(g, 6, 121) == 1e6 + 6 * 562500 + 6000
    s = po.AcqSettings.gps(acqSearchBand=20.0, acqCohIntegration=1)
    assert po.num_bins(s) == 41 and po.bin_freq(s, 1, 1) == 2.42e6 - 10000 and po.bin_freq(s, 1, 41) == 2.42e6 + 10000
    w = po.AcqSettings.gps(acqSearchBand=20.0, acqCohIntegration=10, n_noncoh=20)
    assert po.num_bins(w) == 401 and po.samples_needed(w) == 200 * 16000
    c = po.sampled_code(s, 1)
    assert c.shape == (16000,) and set(np.unique(c)) == {-1.0, 1.0}
    # 16000/1023 = 15.64 samples per chip: first chip covers samples 0..14 (ceil rule), last sample = chip 1023
    from gnss_sdr_ru_b200.codes import ca_code

    assert (c[:15] == ca_code(1)[0]).all() and c[15] == ca_code(1)[1] and c[-1] == ca_code(1)[1022]


def test_exclusion_ranges(oracle_lib):
    from oracle import pcps_oracle as po

    s = po.AcqSettings.gps()
    n = 16000
    r = po.exclusion_range(s, 5000)
    assert r[0] == 1 and 4984 in r and 4985 not in r and 5015 not in r and 5016 in r and r[-1] == n
    r = po.exclusion_range(s, 3)  # e1 < 2: range wraps past the end
    assert r[0] == 19 and r[-1] == n + 3 - 16
    r = po.exclusion_range(s, 15990)  # e2 > n
    assert r[0] == 15990 + 16 - n and r[-1] == 15990 - 16
    g = po.AcqSettings.glonass()
    assert po.exclusion_range(g, 5000)[-1] == n and 5031 in po.exclusion_range(g, 5000) and 5030 not in po.exclusion_range(g, 5000)


def test_oracle_regression_fixture(oracle_lib):
    from oracle import pcps_oracle as po

    g = np.load(os.path.join(HERE, "golden", "acq_golden.npz"))
    rec = unpack2(g["packed"])
    s = po.AcqSettings.gps(acqSearchBand=float(g["band"]), acqCohIntegration=int(g["coh"]), svList=[int(x) for x in g["sv"]])
    res = po.acquisition(po.to_complex(rec), s)
    assert [r["bin"] for r in res] == [int(x) for x in g["bin"]]
    assert [r["codePhaseRaw"] for r in res] == [int(x) for x in g["codePhaseRaw"]]
    assert np.allclose([r["peakMetric"] for r in res], g["peakMetric"], rtol=1e-9)
    assert [r["codePhase"] for r in res] == [int(x) for x in g["codePhase"]]
    assert res[0]["carrFreq"] == float(g["carrFreq"][0]) and res[1]["codePhase"] == 0  # PRN 6 absent -> below threshold
