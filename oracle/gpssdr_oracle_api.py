"""TEST INFRASTRUCTURE ONLY -- ctypes front of oracle/gpssdr_oracle.c (the restatement) and of
oracle/_ref/libgpssdr_ref.so (the reference's own primitives compiled in place)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libgpssdr_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libgpssdr_ref.so")
RT = "/root/reference/trunk/GNSS_SOFTWARE_RECEIVERS/REALTIME_RECEIVERS/GPS/GPS_SDR_REAL_TIME_GPS_RECEIVER"

_lib = None
_ref = None


class Result(C.Structure):
    _fields_ = [("code_phase", C.c_int32), ("doppler", C.c_int32), ("magnitude", C.c_uint32), ("pad_", C.c_int32)]


def build(force: bool = False) -> None:
    src = os.path.join(HERE, "gpssdr_oracle.c")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-Wall", "-std=gnu11", "-shared", "-o", SO, src, "-lm"])
    if os.path.isdir(RT) and (force or not os.path.exists(REF_SO)):
        subprocess.check_call(["bash", os.path.join(HERE, "build_ref_gpssdr.sh")], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(SO)
        vp = C.c_void_p
        L.gso_fft.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int]
        L.gso_cmulsc.argtypes = [vp, vp, vp, C.c_int, C.c_int]
        L.gso_cacc.argtypes = [vp, vp, C.c_int, vp, vp]
        L.gso_cmag.argtypes = [vp, C.c_int]
        L.gso_max.argtypes = [vp, vp, vp, C.c_int]
        L.gso_sine_gen.argtypes = [vp, C.c_double, C.c_double, C.c_int]
        L.gso_wipeoff_gen.argtypes = [vp, C.c_double, C.c_double, C.c_int]
        L.gso_acq_new.argtypes = [C.c_double]
        L.gso_acq_new.restype = vp
        L.gso_acq_free.argtypes = [vp]
        L.gso_prep_if.argtypes = [vp, C.c_int, vp]
        L.gso_acq_strong.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(Result)]
        L.gso_acq_weak.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(Result)]
        L.gso_acq_medium.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(Result)]
        _lib = L
    return _lib


def have_ref() -> bool:
    build()
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        build()
        R = C.CDLL(REF_SO)
        vp = C.c_void_p
        R.gsr_fft.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int]
        R.gsr_cmulsc.argtypes = [vp, vp, vp, C.c_int, C.c_int]
        R.gsr_cmuls.argtypes = [vp, vp, C.c_int, C.c_int]
        R.gsr_cacc.argtypes = [vp, vp, C.c_int, vp, vp]
        R.gsr_cmag.argtypes = [vp, C.c_int]
        R.gsr_max.argtypes = [vp, vp, vp, C.c_int]
        R.gsr_sine_gen.argtypes = [vp, C.c_double, C.c_double, C.c_int]
        R.gsr_wipeoff_gen.argtypes = [vp, C.c_double, C.c_double, C.c_int]
        _ref = R
    return _ref


class GpsSdrAcquisition:
    """Acquisition(fsample, fif) + doPrepIF + doAcqStrong / doAcqMedium / doAcqWeak of the restatement"""

    def __init__(self, fif: float = 38400.0):
        self.L = lib()
        self.a = self.L.gso_acq_new(fif)

    def close(self):
        if self.a:
            self.L.gso_acq_free(self.a)
            self.a = None

    def doPrepIF(self, _type: int, buff: np.ndarray):
        b = np.ascontiguousarray(buff, dtype=np.int16)
        need = {0: 1, 1: 10, 2: 310}[_type] * 2048 * 2
        assert b.size >= need
        self.L.gso_prep_if(self.a, _type, b.ctypes.data)

    def _run(self, fn, code: np.ndarray, doppmin: int, doppmax: int):
        c = np.ascontiguousarray(code, dtype=np.int16)
        r = Result()
        fn(self.a, c.ctypes.data, doppmin, doppmax, C.byref(r))
        return dict(code_phase=int(r.code_phase), doppler=int(r.doppler), magnitude=int(r.magnitude))

    def doAcqStrong(self, code, doppmin, doppmax):
        return self._run(self.L.gso_acq_strong, code, doppmin, doppmax)

    def doAcqWeak(self, code, doppmin, doppmax):
        return self._run(self.L.gso_acq_weak, code, doppmin, doppmax)

    def doAcqMedium(self, code, doppmin, doppmax):
        return self._run(self.L.gso_acq_medium, code, doppmin, doppmax)
