"""TEST INFRASTRUCTURE ONLY (oracle) -- NumPy float64 restatement of the Scilab receivers' tracking.

PARITY UNPINNED beyond the restatement itself (no Scilab/Octave here, no recording or expected output in
the reference).  Follows, line by line,

  SCI/GLONASS/L1/tracking.sci:100-425        (SCI = trunk/GNSS_SOFTWARE_RECEIVERS/POSTPROCESSING_SCILAB_RECEIVERS)
  SCI/GLONASS/L1/include/calcLoopCoef.sci:38-43, calcFLLPLLLoopCoef.sci:36-38
  SCI/GLONASS/L1/include/preRun.sci:66-81    (channel list from acqResults, strongest first)
  SCI/GLONASS/L1/initSettings.sci:41-107

for GLONASS L1OF, and the same algorithm with the C/A code and the GPS code-aiding term
(tracking.sci:366, the commented GPS formula) when settings.system == "gps".

One code period per iteration, variable block size ceil((L-rem)/step), float carrier/code NCOs,
FLL-assisted PLL + DLL (tracking.sci:226-400).  Elementwise expressions keep the reference's operation
order; the only intentional freedom is the summation order inside sum().
"""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from gnss_sdr_ru_b200.codes import ca_code, st_code  # noqa: E402


@dataclass
class TrackSettings:
    system: str = "glonass"
    samplingFreq: float = 16e6
    IF: float = 1e6
    L1_IF_step: float = 0.5625e6
    GLONASS_zero_channel: float = 1602e6
    codeFreqBasis: float = 0.511e6
    codeLength: int = 511
    skipNumberOfSamples: int = 0  # skipNumberOfBytes / dataTypeSizeInBytes (complex samples)
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
                 dllNoiseBandwidth=0.1, dllCorrelatorSpacing=0.2)
        d.update(kw)
        return TrackSettings(**d)


def calc_loop_coef(LBW, zeta, k):
    Wn = LBW * 8 * zeta / (4 * zeta ** 2 + 1)
    return k / (Wn * Wn), 2.0 * zeta / Wn


def calc_fll_pll_coef(pllbw, fllbw, T):
    k1 = T * ((pllbw / 0.53) ** 2) + 1.414 * (pllbw / 0.53)
    k2 = 1.414 * (pllbw / 0.53)
    k3 = T * (fllbw / 0.25)
    return k1, k2, k3


def pre_run(acq, settings: TrackSettings):
    """preRun.sci:66-81: channels = detected signals sorted by peakMetric (descending)."""
    order = np.argsort(-np.asarray(acq["peakMetric"]), kind="stable")
    n = min(settings.numberOfChannels, int(np.sum(np.asarray(acq["carrFreq"]) != 0)))
    chans = []
    for ii in range(n):
        i = int(order[ii])
        chans.append(dict(SVN=i + 1, FCH=int(acq["freqChannel"][i]), acquiredFreq=float(acq["carrFreq"][i]),
                          codePhase=int(acq["codePhase"][i])))
    return chans


FIELDS = ("I_E", "I_P", "I_L", "Q_E", "Q_P", "Q_L", "carrFreq", "codeFreq", "dllDiscr", "dllDiscrFilt", "pllDiscr",
          "pllDiscrFilt", "absoluteSample")


def tracking(iq_int8: np.ndarray, channel: dict, s: TrackSettings):
    """One channel of tracking.sci.  iq_int8: interleaved I,Q record (the whole file).  Returns a dict of
    arrays of length msToProcess (shorter if the record ends)."""
    x = np.asarray(iq_int8, dtype=np.int8)
    sig_all = x[0::2].astype(np.float64) + 1j * x[1::2].astype(np.float64)
    fs = s.samplingFreq
    L = s.codeLength
    code = (ca_code(channel["FCH"]) if s.system == "gps" else st_code()).astype(np.float64)
    caCode = np.concatenate([code[-1:], code, code[:1]])
    earlyLateSpc = s.dllCorrelatorSpacing
    PDIcode = 0.001
    tau1code, tau2code = calc_loop_coef(s.dllNoiseBandwidth, s.dllDampingRatio, 1.0)
    k1, k2, k3 = calc_fll_pll_coef(s.pllNoiseBandwidth, s.fllNoiseBandwidth, 0.001)
    pos = s.skipNumberOfSamples + (channel["codePhase"] - 1)
    currentSample = 2 * pos  # bytes (dataAdaptCoeff = 2, 1 byte per value)
    codeFreq = s.codeFreqBasis
    remCodePhase = 0.0
    carrFreq = channel["acquiredFreq"]
    carrFreqBasis = channel["acquiredFreq"]
    remCarrPhase = 0.0
    oldCodeNco = oldCodeError = oldCarrNco = oldCarrError = 0.0
    I1 = I2 = Q1 = Q2 = 0.001
    out = {f: [] for f in FIELDS}
    fch = channel["FCH"]
    for _ in range(s.msToProcess):
        codePhaseStep = codeFreq / fs
        blksize = int(np.ceil((L - remCodePhase) / codePhaseStep))
        if pos + blksize > sig_all.size:
            break
        rawSignal = sig_all[pos:pos + blksize]
        pos += blksize
        currentSample += 2 * blksize
        j = np.arange(blksize, dtype=np.float64)
        tE = (remCodePhase - earlyLateSpc) + j * codePhaseStep
        tL = (remCodePhase + earlyLateSpc) + j * codePhaseStep
        tP = remCodePhase + j * codePhaseStep
        earlyCode = caCode[np.ceil(tE).astype(np.int64)]   # Scilab: caCode(ceil(tcode)+1), 1-based
        lateCode = caCode[np.ceil(tL).astype(np.int64)]
        promptCode = caCode[np.ceil(tP).astype(np.int64)]
        remCodePhase = (tP[blksize - 1] + codePhaseStep) - L
        time = np.arange(blksize + 1, dtype=np.float64) / fs
        trigarg = ((carrFreq * 2.0 * np.pi) * time) + remCarrPhase
        last = trigarg[blksize]
        remCarrPhase = last - np.fix(last / (2 * np.pi)) * (2 * np.pi)
        carrsig = np.exp(1j * trigarg[:blksize])
        mixed = carrsig * rawSignal
        qBB = mixed.real
        iBB = mixed.imag
        I_E = float(np.sum(earlyCode * iBB))
        Q_E = float(np.sum(earlyCode * qBB))
        I_P = float(np.sum(promptCode * iBB))
        Q_P = float(np.sum(promptCode * qBB))
        I_L = float(np.sum(lateCode * iBB))
        Q_L = float(np.sum(lateCode * qBB))
        I2, Q2 = I1, Q1
        I1, Q1 = I_P, Q_P
        cross = I1 * Q2 - I2 * Q1
        dot = abs(I1 * I2 + Q1 * Q2)
        freqError = np.arctan2(cross, dot) / np.pi
        carrError = np.arctan(Q_P / I_P) / (2.0 * np.pi)
        carrNco = oldCarrNco + k1 * carrError - k2 * oldCarrError - k3 * freqError
        oldCarrNco = carrNco
        oldCarrError = carrError
        carrFreq = carrFreqBasis + carrNco
        sE = np.sqrt(I_E * I_E + Q_E * Q_E)
        sL = np.sqrt(I_L * I_L + Q_L * Q_L)
        codeError = (sE - sL) / (sE + sL)
        codeNco = oldCodeNco + (tau2code / tau1code) * (codeError - oldCodeError) + codeError * (PDIcode / tau1code)
        oldCodeNco = codeNco
        oldCodeError = codeError
        if s.system == "glonass":
            codeFreq = s.codeFreqBasis - codeNco + (carrFreq - (s.IF + s.L1_IF_step * fch)) / (
                (s.GLONASS_zero_channel + fch * s.L1_IF_step) / s.codeFreqBasis)
        else:
            codeFreq = s.codeFreqBasis - codeNco + ((carrFreq - s.IF) / 1540)
        absoluteSample = currentSample / 2 - remCodePhase * (fs / 1000) / L
        for f, v in zip(FIELDS, (I_E, I_P, I_L, Q_E, Q_P, Q_L, carrFreq, codeFreq, codeError, codeNco, carrError, carrNco,
                                 absoluteSample)):
            out[f].append(v)
    return {f: np.array(v) for f, v in out.items()}
