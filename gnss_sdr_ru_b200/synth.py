"""Synthetic IF record generator (host side, numpy) and the packed 2-bit sample format.

Modelled on the reference's only in-repo generator,
SIM/glonass_l3_generator.sce:60-186 (phase-continuous carrier + code NCOs, 4-level quantiser
``round(1.5*s+1.5)*2-3`` -> int8 {-3,-1,+1,+3}, interleaved I,Q written with mput(...,'c')), with the
additions SURVEY.md §8(d) asks for: complex AWGN of unit power, amplitude from C/N0, data bits.

Signal convention (verified against the compiled reference C receiver):
a satellite at IF frequency f appears in the record as  A*code(t)*data(t)*exp(-i(2*pi*f*t+phi0)),
because both receivers wipe the carrier off with exp(+i...) and read I_accum=Im, Q_accum=Re.

Record layout (what OSG/osgnss_next_step.c:172 and SCI/*/postProcessing.sce:91-102 consume):
int8, interleaved I,Q, fs complex samples per second.

Packed 2-bit layout (row U of SURVEY.md §8a; defined by this build because no decoder exists in the
reference): one byte holds two complex samples as four 2-bit fields, LSB first, in the order
I0,Q0,I1,Q1; field code -> value {0:+1, 1:-1, 2:+3, 3:-3} (FE/.../win32_sampler.h:45-55:
"(0 1 2 3) = {1,-1,3,-3}; lsb of each byte is the earliest sample").
"""
from __future__ import annotations

from dataclasses import dataclass, field
from fractions import Fraction

import numpy as np

from .codes import ca_code, st_code

FS = 16_000_000  # complex samples/s (OSG/include/globals.h:12, FE/.../win32_sampler.h:61)
GPS_IF = 2.42e6  # OSG/include/globals.h:16, SCI/GPS/L1/initSettings.sci
GPS_L1 = 1575.42e6
GLO_IF0 = 1.0e6  # SCI/GLONASS/L1/initSettings.sci (IF of frequency channel 0)
GLO_IF_STEP = 562_500.0
GLO_L1 = 1602.0e6


@dataclass
class Sat:
    """One emitter in a synthetic record."""

    system: str = "gps"  # "gps" | "glonass"
    prn: int = 1  # GPS PRN 1..32, or GLONASS frequency channel k=-7..+6
    cn0_dbhz: float = 48.0
    doppler_hz: float = 0.0
    code_phase_chips: float = 0.0  # code phase of the first sample, in chips
    carrier_phase_cycles: float = 0.0
    data_seed: int | None = None  # None -> no data modulation
    data_rate_hz: float = 50.0  # GPS 50 bps; GLONASS L1OF 100 sym/s (meander)
    data_bits: np.ndarray = field(default=None, repr=False, compare=False)  # explicit bits (0/1), repeated cyclically (device generator)
    _chips: np.ndarray = field(default=None, repr=False, compare=False)

    def carrier_if(self) -> float:
        if self.system == "gps":
            return GPS_IF
        return GLO_IF0 + self.prn * GLO_IF_STEP

    def code_rate(self) -> float:
        if self.system == "gps":
            return 1.023e6 * (1.0 + self.doppler_hz / GPS_L1)
        return 0.511e6 * (1.0 + self.doppler_hz / (GLO_L1 + self.prn * GLO_IF_STEP))

    def chips(self) -> np.ndarray:
        if self._chips is None:
            self._chips = ca_code(self.prn) if self.system == "gps" else st_code()
        return self._chips


def _mod_exact(freq_hz: float, n0: int, fs: int, mod: int = 1) -> float:
    """(freq*n0/fs) mod `mod`, computed exactly in rationals (freq quantised to 1 uHz)."""
    f = Fraction(int(round(freq_hz * 1_000_000)), 1_000_000)
    x = f * n0 / fs
    return float(x - mod * (x.numerator // (x.denominator * mod)))


def make_record(
    sats: list[Sat],
    n_samples: int,
    seed: int,
    fs: int = FS,
    noise: bool = True,
    chunk: int = 1 << 20,
) -> np.ndarray:
    """Return an int8 array of shape (2*n_samples,) with interleaved I,Q in {-3,-1,+1,+3}."""
    rng = np.random.default_rng(seed)
    out = np.empty(2 * n_samples, dtype=np.int8)
    data_bits = {}
    for k, s in enumerate(sats):
        if s.data_bits is not None:  # explicit bits, cyclic; bit 1 inverts the carrier (as the device generator)
            nbits = int(n_samples / fs * s.data_rate_hz) + 2
            b = np.asarray(s.data_bits, dtype=np.int64)
            data_bits[k] = (1 - 2 * b[np.arange(nbits) % b.size]).astype(np.float64)
        elif s.data_seed is not None:
            nbits = int(n_samples / fs * s.data_rate_hz) + 2
            data_bits[k] = (2 * np.random.default_rng(s.data_seed).integers(0, 2, nbits) - 1).astype(np.float64)
    sigma = np.sqrt(0.5)  # per component, complex noise power 1
    thr = sigma  # magnitude threshold of the 2-bit quantiser (~1 sigma)
    for n0 in range(0, n_samples, chunk):
        m = min(chunk, n_samples - n0)
        k = np.arange(m, dtype=np.float64)
        acc_i = np.zeros(m)
        acc_q = np.zeros(m)
        for si, s in enumerate(sats):
            amp = np.sqrt(10.0 ** (s.cn0_dbhz / 10.0) / fs)
            f = s.carrier_if() + s.doppler_hz
            ph0 = _mod_exact(f, n0, fs) + s.carrier_phase_cycles
            ph = 2.0 * np.pi * (ph0 + (f / fs) * k)
            chips = s.chips()
            L = len(chips)
            fc = s.code_rate()
            cp0 = (s.code_phase_chips + _mod_exact(fc, n0, fs, L)) % L
            cidx = np.floor(cp0 + (fc / fs) * k).astype(np.int64) % L
            a = amp * chips[cidx].astype(np.float64)
            if si in data_bits:
                bidx = ((n0 + k) * (s.data_rate_hz / fs)).astype(np.int64)
                a = a * data_bits[si][bidx]
            # exp(-i*ph): I = cos, Q = -sin
            acc_i += a * np.cos(ph)
            acc_q -= a * np.sin(ph)
        if noise:
            acc_i += sigma * rng.standard_normal(m)
            acc_q += sigma * rng.standard_normal(m)
            t = thr
        else:
            t = 0.5 * max(np.abs(acc_i).max(), np.abs(acc_q).max(), 1e-30)
        qi = np.where(np.abs(acc_i) > t, 3, 1) * np.where(acc_i >= 0, 1, -1)
        qq = np.where(np.abs(acc_q) > t, 3, 1) * np.where(acc_q >= 0, 1, -1)
        out[2 * n0 : 2 * (n0 + m) : 2] = qi.astype(np.int8)
        out[2 * n0 + 1 : 2 * (n0 + m) : 2] = qq.astype(np.int8)
    return out


# ----------------------------------------------------------------------------------------------
# packed 2-bit format
_CODE_TO_VAL = np.array([1, -1, 3, -3], dtype=np.int8)


def pack2(iq: np.ndarray) -> np.ndarray:
    """int8 interleaved I,Q in {-3,-1,1,3} (length multiple of 4) -> packed bytes (length/4)."""
    iq = np.asarray(iq, dtype=np.int8)
    if iq.size % 4:
        raise ValueError("need a multiple of 4 values (2 complex samples per byte)")
    if not np.isin(iq, _CODE_TO_VAL).all():
        raise ValueError("packed format holds only the values -3,-1,+1,+3")
    code = ((iq < 0).astype(np.uint8)) | ((np.abs(iq) == 3).astype(np.uint8) << 1)
    c = code.reshape(-1, 4)
    return (c[:, 0] | (c[:, 1] << 2) | (c[:, 2] << 4) | (c[:, 3] << 6)).astype(np.uint8)


def unpack2(packed: np.ndarray) -> np.ndarray:
    """Inverse of :func:`pack2` (CPU reference unpacker used to feed the oracle)."""
    p = np.asarray(packed, dtype=np.uint8)
    codes = np.stack([(p >> s) & 3 for s in (0, 2, 4, 6)], axis=1).reshape(-1)
    return _CODE_TO_VAL[codes]
