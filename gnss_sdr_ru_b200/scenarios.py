"""Synthetic workloads for the BASELINE.json configurations.

Everything here is deterministic in its seed so tests, bench.py and the CPU baseline see the same
records.  GPS tracking scenarios place 12 satellites (one per correlator channel) with Dopplers in
+-4.5 kHz that avoid abs(fd) < 600 Hz (reference quirk Q2: a detection in Doppler bin 0 on the
first pass leaves carrFreqBasis = 0) and warm-start every channel in the Doppler bin of its
satellite, i.e. in the state ch_acq (OSG/isr/osgpsisr.c:443-449) is in when its serial search enters
that bin, so that a 10 s record contains search, confirm, pull-in and tracking.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import abi
from .synth import FS, GPS_IF, GPS_L1, GLO_IF0, GLO_IF_STEP, GLO_L1, Sat


@dataclass
class TrackScenario:
    sats: list  # list[Sat], one per channel
    prns: list  # PRN per channel (12)
    n_freq: list  # warm-start Doppler bin per channel
    fch: list = None  # GLONASS frequency channel per correlator channel (None: GPS scenario)


def gps_tracking_scenario(seed: int, n_sats: int = 12, cn0=(46.0, 52.0)) -> TrackScenario:
    rng = np.random.default_rng(seed)
    prns = [int(p) for p in rng.choice(np.arange(1, 33), size=n_sats, replace=False)]
    sats, bins = [], []
    for i, prn in enumerate(prns):
        n = int(rng.integers(1, 5)) * (1 if rng.random() < 0.5 else -1)  # bins +-1..+-4
        fd = 1000.0 * n + float(rng.uniform(-300.0, 300.0))
        k = int(rng.integers(20, 400))  # half-chip search cell in which the serial search meets the code
        theta = (1023.0 - 0.5 * k + float(rng.uniform(-0.1, 0.1))) % 1023.0
        sats.append(
            Sat(system="gps", prn=prn, cn0_dbhz=float(rng.uniform(*cn0)), doppler_hz=fd, code_phase_chips=theta,
                carrier_phase_cycles=float(rng.random()), data_seed=int(seed * 100 + i + 1))
        )
        bins.append(n)
    while len(prns) < abi.N_CHANNELS:
        prns.append(0)
        bins.append(0)
    return TrackScenario(sats=sats, prns=prns, n_freq=bins)


def synth_sat_array(scenarios: list) -> tuple:
    """ctypes array of gnssb200_synth_sat [n_streams*12] for the device generator."""
    n_sats = max(len(sc.sats) for sc in scenarios)
    arr = (abi.SynthSat * (len(scenarios) * n_sats))()
    keep = []
    for s, sc in enumerate(scenarios):
        for k in range(n_sats):
            d = arr[s * n_sats + k]
            if k >= len(sc.sats):
                d.cn0_dbhz = 0.0
                d.samp_rate = FS
                continue
            sat = sc.sats[k]
            d.system = abi.SYS_GPS if sat.system == "gps" else abi.SYS_GLONASS
            d.prn = sat.prn if sat.system == "gps" else 0
            d.cn0_dbhz = sat.cn0_dbhz
            d.samp_rate = FS
            d.carrier_hz = sat.carrier_if() + sat.doppler_hz
            d.code_hz = sat.code_rate()
            d.code_phase_chips = sat.code_phase_chips
            d.carrier_phase_cycles = sat.carrier_phase_cycles
            d.data_seed = sat.data_seed or 0
            d.data_rate_hz = sat.data_rate_hz
            if sat.data_bits is not None:
                bits = np.ascontiguousarray(sat.data_bits, dtype=np.uint8)
                keep.append(bits)  # the array must outlive the ctypes pointer
                d.data_bits = bits.ctypes.data
                d.n_data_bits = bits.size
    arr._keepalive = keep
    return arr, n_sats


def glonass_tracking_scenario(seed: int, n_sats: int = 12, cn0=(46.0, 52.0)) -> TrackScenario:
    """Twelve GLONASS L1OF satellites on distinct frequency channels (k = -7..+6, IF 1 MHz + k * 562.5 kHz), one per
    correlator channel of the integer receiver (PRN register abi.PRN_GLONASS), each warm-started in its Doppler bin."""
    rng = np.random.default_rng(seed)
    ks = [int(k) for k in rng.choice(np.arange(-7, 7), size=n_sats, replace=False)]
    sats, bins = [], []
    for i, k in enumerate(ks):
        n = int(rng.integers(1, 5)) * (1 if rng.random() < 0.5 else -1)
        fd = 1000.0 * n + float(rng.uniform(-300.0, 300.0))
        cell = int(rng.integers(20, 300))  # half-chip search cell in which the serial search meets the code
        theta = (511.0 - 0.5 * cell + float(rng.uniform(-0.1, 0.1))) % 511.0
        sats.append(Sat(system="glonass", prn=k, cn0_dbhz=float(rng.uniform(*cn0)), doppler_hz=fd, code_phase_chips=theta,
                        carrier_phase_cycles=float(rng.random()), data_seed=int(seed * 100 + i + 1), data_rate_hz=100.0))
        bins.append(n)
    return TrackScenario(sats=sats, prns=[abi.PRN_GLONASS] * n_sats + [0] * (abi.N_CHANNELS - n_sats),
                         n_freq=bins + [0] * (abi.N_CHANNELS - n_sats), fch=ks + [0] * (abi.N_CHANNELS - n_sats))


def apply_tracking_scenario(engine, stream: int, sc: TrackScenario) -> None:
    """simple_cold_allocate + per-channel warm start on a TrackingEngine host state."""
    engine.simple_cold_allocate(stream, sc.prns)
    for ch, (prn, n) in enumerate(zip(sc.prns, sc.n_freq)):
        if prn > 0:
            if sc.fch is not None and prn == abi.PRN_GLONASS:
                engine.set_glonass_channel(stream, ch, sc.fch[ch])
            engine.warm_start(stream, ch, n)


# ---- acquisition scenarios (C1, C3, C4) -----------------------------------------------------------
def gps_acq_scenario(seed: int, prns=(3, 7, 9, 14, 19, 22, 27, 31), cn0=(44.0, 50.0), doppler_span=9000.0):
    rng = np.random.default_rng(seed)
    sats = []
    for i, prn in enumerate(prns):
        fd = float(rng.uniform(-doppler_span, doppler_span))
        sats.append(Sat(system="gps", prn=int(prn), cn0_dbhz=float(rng.uniform(*cn0)), doppler_hz=fd,
                        code_phase_chips=float(rng.uniform(0, 1023)), carrier_phase_cycles=float(rng.random()),
                        data_seed=None))
    return sats


def glonass_acq_scenario(seed: int, channels=(-7, -4, -1, 0, 2, 5, 6), cn0=(44.0, 50.0), doppler_span=5000.0):
    rng = np.random.default_rng(seed)
    sats = []
    for i, k in enumerate(channels):
        fd = float(rng.uniform(-doppler_span, doppler_span))
        sats.append(Sat(system="glonass", prn=int(k), cn0_dbhz=float(rng.uniform(*cn0)), doppler_hz=fd,
                        code_phase_chips=float(rng.uniform(0, 511)), carrier_phase_cycles=float(rng.random()),
                        data_seed=int(seed * 100 + i + 1), data_rate_hz=100.0))
    return sats


def gps_weak_acq_scenario(seed: int):
    rng = np.random.default_rng(seed)
    prns = [int(p) for p in rng.choice(np.arange(1, 33), size=6, replace=False)]
    sats = []
    for i, prn in enumerate(prns):
        c = float(rng.uniform(28.0, 33.0)) if i < 4 else 45.0
        sats.append(Sat(system="gps", prn=prn, cn0_dbhz=c, doppler_hz=float(rng.uniform(-9000, 9000)),
                        code_phase_chips=float(rng.uniform(0, 1023)), carrier_phase_cycles=float(rng.random()),
                        data_seed=None))
    return sats
