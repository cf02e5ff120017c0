# small end-to-end invocations of every kernel family (tracking both formats with forced slices, GPS-SDR acquisition, navbits): a quick manual smoke
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gnss_sdr_ru_b200 import abi
from gnss_sdr_ru_b200.receiver import TrackingEngine
from gnss_sdr_ru_b200.synth import Sat, make_record, pack2
NS, nblk = 8192, 40
rec = make_record([Sat(prn=27, doppler_hz=1200, cn0_dbhz=50, code_phase_chips=1015.3, data_seed=5)], NS * nblk, seed=7)
for fmt, buf in ((abi.FMT_PACKED2, pack2(rec)), (abi.FMT_INT8_IQ, rec.view(np.uint8))):
    eng = TrackingEngine(n_streams=2)
    for s in range(2):
        eng.simple_cold_allocate(s, [27, 0, 0, 0, 0, 0, 0, 0, 9, 0, 0, 3])
        eng.warm_start(s, 0, 1)
    eng.upload()
    eng.set_track_slice(7)
    d, c = eng.run_host(np.stack([buf, buf]), nblk, NS, fmt, dump_cap=64)
    print("track fmt", fmt, c.sum())
    eng.close()
from gnss_sdr_ru_b200.gpssdr_acq import Acquisition
rng = np.random.default_rng(1)
a = Acquisition()
print(a.doAcqStrong(np.round(8 * rng.standard_normal((2048, 2))).astype(np.int16), [1, 2], -2000, 2000)[0])
a.close()
from gnss_sdr_ru_b200.navbits import NavBitsEngine
nb = NavBitsEngine()
print(nb.findTimeMarks("TT", rng.standard_normal((2, 3000))))
nb.close()
