# ad-hoc probe (not the bench): GPS-SDR weak acquisition, 32 satellites, +-10 kHz
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gnss_sdr_ru_b200 import gpssdr_codes
from gnss_sdr_ru_b200.gpssdr_acq import Acquisition
rng = np.random.default_rng(1)
n = 310 * 2048
rec = np.round(8.0 * rng.standard_normal((n, 2))).astype(np.int16)
a = Acquisition()
for _ in range(3):
    t = time.time(); r = a.doAcqWeak(rec, list(range(32)), -10000, 10000); dt = time.time() - t
    print(f"weak 32 sv: {dt*1e3:.1f} ms wall, kernels {a.L.gnssb200_last_kernel_ms(a.h):.2f} ms")
rec10 = rec[: 10 * 2048]
for _ in range(3):
    t = time.time(); r = a.doAcqMedium(rec10, list(range(32)), -10000, 10000, prior=(2, rec)); dt = time.time() - t
    print(f"medium 32 sv (after a 310-ms preparation): {dt*1e3:.1f} ms wall, kernels {a.L.gnssb200_last_kernel_ms(a.h):.2f} ms")
for _ in range(3):
    t = time.time(); r = a.doAcqMedium(rec10, list(range(32)), -10000, 10000); dt = time.time() - t
    print(f"medium 32 sv (new object): {dt*1e3:.1f} ms wall, kernels {a.L.gnssb200_last_kernel_ms(a.h):.2f} ms")
a.close()
