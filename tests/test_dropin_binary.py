"""The reference receiver's own host code (main, gpsisr, register accessors -- compiled unchanged
from /root/reference into oracle/_ref/osgnss_gpu) running on top of libgnssb200.so instead of
correlator.c, against the stock all-CPU binary oracle/_ref/osgnss_ref34 on the same record:
corr_out.csv (the reference's own export of the E/P/L dumps, osgpsisr.c:50-60) must be identical."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")


@pytest.mark.skipif(not os.path.exists(os.path.join(REFDIR, "osgnss_gpu")), reason="oracle/_ref/osgnss_gpu not built (needs /root/reference at build time)")
def test_reference_main_on_gpu_correlator_writes_identical_corr_out():
    import torch

    from gnss_sdr_ru_b200 import abi
    from gnss_sdr_ru_b200.lib import check, lib
    from gnss_sdr_ru_b200.scenarios import TrackScenario, synth_sat_array
    from gnss_sdr_ru_b200.synth import Sat

    # stock main allocates PRN 27 on channel 0 and PRN 9 on channel 8 (osgnss_next_step.c:78-79) and
    # searches bin 0 first (2.05 s), then +1 kHz.
    sats = [Sat(prn=27, doppler_hz=1150.0, cn0_dbhz=51, code_phase_chips=1008.3, data_seed=11),
            Sat(prn=9, doppler_hz=-880.0, cn0_dbhz=48, code_phase_chips=1000.0, data_seed=12)]
    n = 8192 * 9500  # 4.86 s: 2.05 s bin-0 sweep + hit in bin +1 + ~1.4 s pull-in
    L = lib()
    h = L.gnssb200_open(0, None)
    buf = torch.empty(2 * n, dtype=torch.int8, device="cuda")
    arr, nsat = synth_sat_array([TrackScenario(sats=sats, prns=[], n_freq=[])])
    check(L.gnssb200_synth(h, buf.data_ptr(), 2 * n, abi.FMT_INT8_IQ, 1, n, C.addressof(arr), nsat, 99, None), "synth")
    rec = buf.cpu().numpy()
    L.gnssb200_close(h)
    outs = {}
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "record.bin")
        rec.tofile(path)
        for name in ("osgnss_ref34", "osgnss_gpu"):
            wd = os.path.join(d, name)
            os.makedirs(wd)
            subprocess.run([os.path.join(REFDIR, name), "-f", path], cwd=wd, check=True, timeout=600, stdout=subprocess.DEVNULL)
            outs[name] = open(os.path.join(wd, "e:\\corr_out.csv"), "rb").read()
    assert len(outs["osgnss_ref34"]) > 1000, "record too short: the reference never finished a pull-in"
    assert outs["osgnss_gpu"] == outs["osgnss_ref34"]
