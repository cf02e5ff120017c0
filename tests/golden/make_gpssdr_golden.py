"""Generates tests/golden/gpssdr_ref_golden.npz: inputs and outputs of the REFERENCE's own GPS-SDR primitives
(oracle/_ref/libgpssdr_ref.so = RT/objects/fft.cpp -DNO_SIMD, RT/simd/x86.cpp, RT/accessories/misc.cpp compiled
in place by oracle/build_ref_gpssdr.sh).  Run in the container that has /root/reference:
    python tests/golden/make_gpssdr_golden.py
The file lets tests/test_gpssdr_oracle.py pin the restatement where the reference tree is absent."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gpssdr_oracle_api as G  # noqa: E402

R = G.ref()
rng = np.random.default_rng(2024)
out = {}
pats = {"R1": np.zeros(16, np.int32), "R2": np.array([0, 0, 0, 0, 0, 0, 0, 1, 0, 1, 0, 1, 1, 1, 1, 1], np.int32), "Rall": np.ones(16, np.int32)}
for name, Rp in pats.items():
    for inv in (0, 1):
        for amp in (60, 20000):
            x = rng.integers(-amp, amp + 1, size=(2048, 2)).astype(np.int16)
            y = x.copy()
            R.gsr_fft(y.ctypes.data, 2048, Rp.ctypes.data, inv, 1)
            out[f"fft_{name}_{inv}_{amp}_in"] = x
            out[f"fft_{name}_{inv}_{amp}_out"] = y
for shift in (9, 10, 14):
    A = rng.integers(-23000, 23001, size=(3000, 2)).astype(np.int16)
    B = rng.integers(-23000, 23001, size=(3000, 2)).astype(np.int16)
    Cc = np.zeros_like(A)
    R.gsr_cmulsc(A.ctypes.data, B.ctypes.data, Cc.ctypes.data, 3000, shift)
    out[f"cmulsc_{shift}_a"], out[f"cmulsc_{shift}_b"], out[f"cmulsc_{shift}_c"] = A, B, Cc
for k, f in enumerate((-38400.0, -38650.0, -38900.0, -39150.0)):
    s = np.zeros((20480, 2), np.int16)
    R.gsr_sine_gen(s.ctypes.data, f, 2048000.0, 20480)
    out[f"sine_{k}"] = s
dft = np.zeros((10, 10, 4), np.int16)
for r in range(10):
    R.gsr_wipeoff_gen(dft[r].ctypes.data, float(np.float32(r) * 25.0 - 112.5), 1000.0, 10)
out["dft_rows"] = dft
d = rng.integers(-3000, 3000, size=(64, 10, 2)).astype(np.int16)
acc = np.zeros((64, 10, 2), np.int32)
for n in range(64):
    for r in range(10):
        i, q = C.c_int32(), C.c_int32()
        R.gsr_cacc(d[n].ctypes.data, dft[r].ctypes.data, 10, C.byref(i), C.byref(q))
        acc[n, r] = (i.value, q.value)
out["cacc_in"], out["cacc_out"] = d, acc
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gpssdr_ref_golden.npz"), **out)
print("written", len(out), "arrays")
