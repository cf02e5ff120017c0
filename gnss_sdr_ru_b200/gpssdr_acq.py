"""Host-side mirror of the GPS-SDR acquisition object (RT/objects/acquisition.h:48-94): doPrepIF followed by
doAcqStrong / doAcqWeak for a list of satellites, computed by csrc/gpssdr_acq.cu."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi, gpssdr_codes
from .lib import GnssB200Error, check, lib

ACQ_TYPE_STRONG, ACQ_TYPE_WEAK = 0, 2
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

    def _acquire(self, _type, buff, svs, doppmin, doppmax):
        iq = np.ascontiguousarray(buff, dtype=np.int16)
        need = (1 if _type == ACQ_TYPE_STRONG else 310) * 2048 * 2
        if iq.size < need:
            raise ValueError(f"buffer holds {iq.size // 2} complex samples, {need // 2} needed")
        sv = np.ascontiguousarray(svs, dtype=np.int32)
        res = (abi.GpsSdrResult * len(sv))()
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
