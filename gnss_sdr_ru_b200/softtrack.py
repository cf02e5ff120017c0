"""Host-side mirror of the Scilab receivers' ``[trackResults, channel] = tracking(fid, channel, settings)``
and ``channel = preRun(acqResults, settings)`` (SCI/GLONASS/L1/tracking.sci, include/preRun.sci:66-81).
Setting and result field names follow the reference; the computation is csrc/softtrack.cu."""
from __future__ import annotations

import ctypes as C
import sys
from dataclasses import dataclass

import numpy as np

from . import abi
from .lib import GnssB200Error, check, lib


@dataclass
class TrackSettings:
    system: str = "glonass"
    samplingFreq: float = 16e6
    IF: float = 1e6
    L1_IF_step: float = 0.5625e6
    GLONASS_zero_channel: float = 1602e6
    codeFreqBasis: float = 0.511e6
    codeLength: int = 511
    skipNumberOfSamples: int = 0
    msToProcess: int = 1000
    numberOfChannels: int = 8
    dllDampingRatio: float = 0.7
    dllNoiseBandwidth: float = 0.5
    dllCorrelatorSpacing: float = 0.05
    pllNoiseBandwidth: float = 25.0
    fllNoiseBandwidth: float = 250.0

    @staticmethod
    def gps(**kw):
        d = dict(system="gps", IF=2.42e6, L1_IF_step=0.0, codeFreqBasis=1.023e6, codeLength=1023,
                 dllNoiseBandwidth=0.1, dllCorrelatorSpacing=0.2)  # SCI/GPS/L1/initSettings.sci:91-98
        d.update(kw)
        return TrackSettings(**d)

    def to_c(self) -> abi.SoftTrackCfg:
        c = abi.SoftTrackCfg()
        c.system = abi.SYS_GPS if self.system == "gps" else abi.SYS_GLONASS
        c.code_length = self.codeLength
        c.samp_freq = self.samplingFreq
        c.IF = self.IF
        c.IF_step = self.L1_IF_step
        c.glonass_zero_channel = self.GLONASS_zero_channel
        c.code_freq = self.codeFreqBasis
        c.dll_damping_ratio = self.dllDampingRatio
        c.dll_noise_bandwidth = self.dllNoiseBandwidth
        c.dll_correlator_spacing = self.dllCorrelatorSpacing
        c.pll_noise_bandwidth = self.pllNoiseBandwidth
        c.fll_noise_bandwidth = self.fllNoiseBandwidth
        c.skip_samples = self.skipNumberOfSamples
        c.ms_to_process = self.msToProcess
        return c


def preRun(acqResults: dict, settings: TrackSettings) -> list:
    """Channel list from acquisition results, strongest peakMetric first (preRun.sci:66-81)."""
    order = np.argsort(-np.asarray(acqResults["peakMetric"]), kind="stable")
    n = min(settings.numberOfChannels, int(np.sum(np.asarray(acqResults["carrFreq"]) != 0)))
    return [dict(SVN=int(i) + 1, FCH=int(acqResults["freqChannel"][i]), acquiredFreq=float(acqResults["carrFreq"][i]),
                 codePhase=int(acqResults["codePhase"][i]), status="T") for i in order[:n]]


class SoftTrackingEngine:
    def __init__(self, device: int = 0, handle=None):
        self.L = lib()
        self._own = handle is None
        if handle is None:
            handle = self.L.gnssb200_open(device, None)
            if not handle:
                raise GnssB200Error("gnssb200_open failed: " + (self.L.gnssb200_last_error_string() or b"?").decode())
        self.h = handle

    def close(self):
        if self._own and self.h:
            self.L.gnssb200_close(self.h)
        self.h = None

    def __del__(self):
        if sys is None or sys.is_finalizing():
            return
        try:
            self.close()
        except Exception:
            pass

    def tracking_device(self, d_iq_ptr: int, n_samples: int, channel: list, settings: TrackSettings, d_out_ptr: int,
                        d_ms_done_ptr: int, stream: int = 0):
        ch = (abi.SoftTrackChan * len(channel))()
        for i, c in enumerate(channel):
            ch[i].sv = c["FCH"]
            ch[i].code_phase = c["codePhase"]
            ch[i].acquired_freq = c["acquiredFreq"]
        cfg = settings.to_c()
        check(self.L.gnssb200_softtrack(self.h, C.byref(cfg), d_iq_ptr, n_samples, ch, len(channel), d_out_ptr, d_ms_done_ptr,
                                        stream or None), "gnssb200_softtrack")

    def tracking(self, iq_int8: np.ndarray, channel: list, settings: TrackSettings) -> list:
        """Host record in, list of trackResults dicts (one per channel) out."""
        import torch

        buf = np.ascontiguousarray(iq_int8, dtype=np.int8)
        d_iq = torch.from_numpy(buf.view(np.uint8)).cuda()
        n_ch = len(channel)
        d_out = torch.zeros((n_ch, settings.msToProcess, 13), dtype=torch.float64, device="cuda")
        d_done = torch.zeros(n_ch, dtype=torch.int32, device="cuda")
        self.tracking_device(d_iq.data_ptr(), buf.size // 2, channel, settings, d_out.data_ptr(), d_done.data_ptr())
        out = d_out.cpu().numpy()
        done = d_done.cpu().numpy()
        res = []
        for i, c in enumerate(channel):
            r = {f: out[i, : done[i], k].copy() for k, f in enumerate(abi.SOFTTRACK_FIELDS)}
            r["SVN"], r["FCH"], r["status"] = c.get("SVN", 0), c["FCH"], c.get("status", "T")
            res.append(r)
        return res

    def last_kernel_ms(self) -> float:
        return float(self.L.gnssb200_last_kernel_ms(self.h))
