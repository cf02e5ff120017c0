"""Host-side mirror of the GPS-SDR acquisition object (RT/objects/acquisition.h:48-94): doPrepIF followed by
doAcqStrong / doAcqMedium / doAcqWeak for a list of satellites, computed by csrc/gpssdr_acq.cu."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi, gpssdr_codes
from .lib import GnssB200Error, check, lib

ACQ_TYPE_STRONG, ACQ_TYPE_MEDIUM, ACQ_TYPE_WEAK = 0, 1, 2
_PREP_MS = {ACQ_TYPE_STRONG: 1, ACQ_TYPE_MEDIUM: 10, ACQ_TYPE_WEAK: 310}
IF_FREQUENCY = 38400.0  # RT/includes/signaldef.h:34


class Acquisition:
    def __init__(self, fif: float = IF_FREQUENCY, handle=None, device: int = 0):
        self.L = lib()
        self._own = handle is None
        self.h = handle if handle is not None else self.L.gnssb200_open(device, None)
        if not self.h:
            raise GnssB200Error("gnssb200_open failed (no CUDA device? there is no CPU path)")
        self.fif = fif
        self.codes = np.ascontiguousarray(gpssdr_codes.fft_codes())  # PRN_Codes

    def close(self):
        if self._own and self.h:
            self.L.gnssb200_close(self.h)
            self.h = None

    @staticmethod
    def _buffer(_type, buff):
        iq = np.ascontiguousarray(buff, dtype=np.int16)
        need = _PREP_MS[_type] * 2048 * 2
        if iq.size < need:
            raise ValueError(f"buffer holds {iq.size // 2} complex samples, {need // 2} needed")
        return iq

    def _acquire(self, _type, buff, svs, doppmin, doppmax, prior=None):
        iq = self._buffer(_type, buff)
        sv = np.ascontiguousarray(svs, dtype=np.int32)
        res = (abi.GpsSdrResult * len(sv))()
        if _type == ACQ_TYPE_MEDIUM:
            piq = self._buffer(prior[0], prior[1]) if prior is not None else None
            check(self.L.gnssb200_gpssdr_acquire_medium(self.h, iq.ctypes.data, piq.ctypes.data if piq is not None else None,
                                                        prior[0] if prior is not None else 0, self.fif, self.codes.ctypes.data,
                                                        self.codes.shape[0], sv.ctypes.data, len(sv), doppmin, doppmax,
                                                        C.addressof(res)), "gnssb200_gpssdr_acquire_medium")
        else:
            check(self.L.gnssb200_gpssdr_acquire(self.h, iq.ctypes.data, _type, self.fif, self.codes.ctypes.data, self.codes.shape[0],
                                                 sv.ctypes.data, len(sv), doppmin, doppmax, C.addressof(res)), "gnssb200_gpssdr_acquire")
        return [dict(sv=int(r.sv), type=int(r.type), code_phase=int(r.code_phase), doppler=int(r.doppler), magnitude=int(r.magnitude),
                     success=int(r.success)) for r in res]

    def doAcqStrong(self, buff, svs, doppmin, doppmax):
        """doPrepIF(0, buff) + doAcqStrong(sv, doppmin, doppmax) for every sv"""
        return self._acquire(ACQ_TYPE_STRONG, buff, svs, doppmin, doppmax)

    def doAcqWeak(self, buff, svs, doppmin, doppmax):
        """doPrepIF(2, buff) + doAcqWeak(sv, doppmin, doppmax) for every sv"""
        return self._acquire(ACQ_TYPE_WEAK, buff, svs, doppmin, doppmax)

    def doAcqMedium(self, buff, svs, doppmin, doppmax, prior=None):
        """doPrepIF(1, buff) + doAcqMedium(sv, doppmin, doppmax) for every sv.  prior = (type, buffer) of the doPrepIF
        that ran on the same object before this one (its rows 40-69 are read, acquisition.cpp:340); None = a new object"""
        return self._acquire(ACQ_TYPE_MEDIUM, buff, svs, doppmin, doppmax, prior)
