"""TEST INFRASTRUCTURE ONLY -- ctypes front ends for the oracle.

* :class:`Oracle`  -> oracle/liboracle.so  (gp2021_oracle.c, the CPU restatement)
* :class:`RefReceiver` -> oracle/_ref/libosgnss_ref34.so (the reference's own C sources compiled
  in place by oracle/build_ref.sh; global state, so one instance per process)

Nothing in the product package imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from gnss_sdr_ru_b200 import abi  # noqa: E402  (struct layouts only)

ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libosgnss_ref34.so")
REF_SO_PRISTINE = os.path.join(HERE, "_ref", "libosgnss_ref.so")


def build(force: bool = False) -> None:
    """Compile liboracle.so (always possible) and oracle/_ref (only where /root/reference exists)."""
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
        os.path.join(HERE, "gp2021_oracle.c")
    ):
        subprocess.check_call(["make", "-C", HERE, os.path.join(HERE, "liboracle.so")], stdout=subprocess.DEVNULL)
    gpu_bin = os.path.join(HERE, "_ref", "osgnss_gpu")
    gpu_lib = os.path.join(os.path.dirname(HERE), "gnss_sdr_ru_b200", "libgnssb200.so")
    stale = os.path.exists(gpu_lib) and (not os.path.exists(gpu_bin))
    if os.path.isdir("/root/reference") and (force or stale or not os.path.exists(REF_SO)):
        subprocess.check_call(["bash", os.path.join(HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


class Oracle:
    """One receiver instance of the CPU restatement."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            build()
            L = C.CDLL(ORACLE_SO)
            L.orc_cfg_default.argtypes = [C.POINTER(abi.Cfg)]
            L.orc_cfg_derive.argtypes = [C.POINTER(abi.Cfg)]
            L.orc_rx_init.argtypes = [C.POINTER(abi.Rx), C.POINTER(abi.Cfg)]
            L.orc_rx_cold_allocate.argtypes = [C.POINTER(abi.Rx), C.POINTER(abi.Cfg), C.POINTER(C.c_int32)]
            L.orc_sim_gp2021.argtypes = [C.POINTER(abi.Rx), C.POINTER(abi.Cfg), C.c_void_p, C.c_long, C.c_int]
            L.orc_gpsisr.argtypes = [C.POINTER(abi.Rx), C.POINTER(abi.Cfg)]
            L.orc_gpsisr.restype = C.c_int
            L.orc_run.argtypes = [C.POINTER(abi.Rx), C.POINTER(abi.Cfg), C.c_void_p, C.c_long, C.c_long,
                                  C.c_void_p, C.c_int, C.c_void_p]
            L.orc_run.restype = C.c_long
            L.orc_ch_cntl.argtypes = [C.POINTER(abi.Rx), C.c_int, C.c_int]
            L.orc_ch_code_slew.argtypes = [C.POINTER(abi.Rx), C.c_int, C.c_int]
            L.orc_ch_epoch_load.argtypes = [C.POINTER(abi.Rx), C.c_int, C.c_uint]
            L.orc_ch_carrier.argtypes = [C.POINTER(abi.Rx), C.POINTER(abi.Cfg), C.c_int, C.c_int64]
            L.orc_ch_code.argtypes = [C.POINTER(abi.Rx), C.POINTER(abi.Cfg), C.c_int, C.c_int64]
            L.orc_code_bits.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int * 3)]
            L.orc_lo_table.argtypes = [C.c_int, C.POINTER(C.c_int * 2)]
            L.orc_code_bit.argtypes = [C.c_int, C.c_int, C.c_int]
            cls._lib = L
        return cls._lib

    def __init__(self, cfg: abi.Cfg | None = None):
        L = self.lib()
        if cfg is None:
            cfg = abi.Cfg()
            L.orc_cfg_default(C.byref(cfg))
            L.orc_cfg_derive(C.byref(cfg))
        self.cfg = cfg
        self.rx = abi.Rx()
        L.orc_rx_init(C.byref(self.rx), C.byref(self.cfg))

    @staticmethod
    def default_cfg(**over) -> abi.Cfg:
        L = Oracle.lib()
        cfg = abi.Cfg()
        L.orc_cfg_default(C.byref(cfg))
        for k, v in over.items():
            setattr(cfg, k, v)
        L.orc_cfg_derive(C.byref(cfg))
        return cfg

    def cold_allocate(self, prns):
        arr = (C.c_int32 * abi.N_CHANNELS)(*prns)
        self.lib().orc_rx_cold_allocate(C.byref(self.rx), C.byref(self.cfg), arr)

    def sim(self, iq: np.ndarray, nsamp: int, iq_mode: int = 1):
        buf = np.ascontiguousarray(iq, dtype=np.int8)
        assert buf.size >= (2 if iq_mode else 1) * nsamp
        self.lib().orc_sim_gp2021(C.byref(self.rx), C.byref(self.cfg), buf.ctypes.data, nsamp, iq_mode)

    def gpsisr(self) -> int:
        return self.lib().orc_gpsisr(C.byref(self.rx), C.byref(self.cfg))

    def run(self, iq: np.ndarray, nsamp: int, nblocks: int, dump_cap: int = 0):
        buf = np.ascontiguousarray(iq, dtype=np.int8)
        assert buf.size >= 2 * nsamp * nblocks
        if dump_cap:
            dumps = np.zeros((abi.N_CHANNELS, dump_cap), dtype=abi.DUMP_DTYPE)
            cnt = np.zeros(abi.N_CHANNELS, dtype=np.int32)
            n = self.lib().orc_run(C.byref(self.rx), C.byref(self.cfg), buf.ctypes.data, nsamp, nblocks,
                                   dumps.ctypes.data, dump_cap, cnt.ctypes.data)
            return n, dumps, cnt
        n = self.lib().orc_run(C.byref(self.rx), C.byref(self.cfg), buf.ctypes.data, nsamp, nblocks, None, 0, None)
        return n, None, None

    def ch_cntl(self, ch, v):
        self.lib().orc_ch_cntl(C.byref(self.rx), ch, v)

    def ch_code_slew(self, ch, v):
        self.lib().orc_ch_code_slew(C.byref(self.rx), ch, v)

    def ch_carrier(self, ch, f):
        self.lib().orc_ch_carrier(C.byref(self.rx), C.byref(self.cfg), ch, f)

    def ch_code(self, ch, f):
        self.lib().orc_ch_code(C.byref(self.rx), C.byref(self.cfg), ch, f)

    def ch_epoch_load(self, ch, v):
        self.lib().orc_ch_epoch_load(C.byref(self.rx), ch, v)

    @staticmethod
    def lo_table():
        out = (C.c_int * 2)()
        tab = []
        for k in range(8):
            Oracle.lib().orc_lo_table(k, C.byref(out))
            tab.append((out[0], out[1]))
        return tab

    @staticmethod
    def code_bits(prn: int, h: int):
        out = (C.c_int * 3)()
        Oracle.lib().orc_code_bits(prn, h, C.byref(out))
        return tuple(out)


# ------------------------------------------------------------------------------------------------
# mirror of the reference's struct tracking_channel (OSG/include/structs.h:86-128) on LP64


class _RefAccum(C.Structure):
    _fields_ = [(n, C.c_short) for n in ("i_prompt", "q_prompt", "i_late", "q_late", "i_early", "q_early")]


class _RefAccumMag(C.Structure):
    _fields_ = [("early_mag", C.c_long), ("prompt_mag", C.c_long), ("late_mag", C.c_long)]


class RefChan(C.Structure):
    _fields_ = [
        ("system", C.c_int), ("state", C.c_int), ("accum", _RefAccum), ("prev_accum", _RefAccum),
        ("accum_mean", _RefAccumMag), ("cross", C.c_long), ("dot", C.c_long), ("carrError", C.c_long),
        ("oldCarrError", C.c_long), ("freqError", C.c_long), ("carrNco", C.c_long), ("oldCarrNco", C.c_long),
        ("carrFreq", C.c_long), ("carrFreqBasis", C.c_long), ("codeError", C.c_long), ("oldCodeError", C.c_long),
        ("codeFreq", C.c_long), ("codeFreqBasis", C.c_long), ("codeNco", C.c_long), ("oldCodeNco", C.c_long),
        ("ch_time", C.c_long), ("n_freq", C.c_int), ("i_confirm", C.c_int), ("n_thresh", C.c_int),
        ("codes", C.c_int), ("del_freq", C.c_int), ("CN0", C.c_char), ("carrier_freq", C.c_long),
        ("carrier_cold_corr", C.c_long), ("sign_pos", C.c_int), ("prev_sign_pos", C.c_int),
        ("sign_count", C.c_int), ("ms_sign", C.c_ulong), ("ms_count", C.c_int), ("ms_set", C.c_int),
        ("fifo0", C.c_ulong), ("fifo1", C.c_ulong), ("bit", C.c_char), ("search_max_PRN_delay", C.c_int),
        ("search_max_f", C.c_int), ("coherent_integration_time", C.c_int),
    ]


class RefGpChan(C.Structure):  # struct gp2021_channel, OSG/correlator/correlator.c:36-47
    _fields_ = [
        ("int_carrier_phase", C.c_uint32), ("int_carrier_cycle", C.c_uint32), ("int_code_phase", C.c_uint32),
        ("int_code_half_chip", C.c_uint16), ("i_prompt_accum", C.c_int32), ("q_prompt_accum", C.c_int32),
        ("i_late_accum", C.c_int32), ("q_late_accum", C.c_int32), ("i_early_accum", C.c_int32),
        ("q_early_accum", C.c_int32),
    ]


class RefReceiver:
    """The compiled reference (global state!).  Use one per process, re-initialised by reset()."""

    def __init__(self, pristine: bool = False, tic_period: float | None = None):
        path = REF_SO_PRISTINE if pristine else REF_SO
        if not os.path.exists(path):
            build()
        # DEEPBIND + -Bsymbolic: the reference's Sim_GP2021_int/REG_* must bind to the reference, even when
        # libgnssb200.so (which exports the same drop-in symbols) lives in the same process
        self.L = C.CDLL(path, mode=os.RTLD_LOCAL | getattr(os, "RTLD_DEEPBIND", 0))
        L = self.L
        L.correlator_init.argtypes = [C.c_double]
        L.Sim_GP2021_int.argtypes = [C.c_void_p, C.c_long]
        L.ch_carrier.argtypes = [C.c_int, C.c_long]
        L.ch_code.argtypes = [C.c_int, C.c_long]
        L.ch_cntl.argtypes = [C.c_int, C.c_int]
        L.ch_code_slew.argtypes = [C.c_int, C.c_int]
        L.ch_epoch_load.argtypes = [C.c_int, C.c_uint]
        self.REG_read = (C.c_int * 256).in_dll(L, "REG_read")
        self.REG_write = (C.c_int * 256).in_dll(L, "REG_write")
        self.chan = (RefChan * 12).in_dll(L, "chan")
        self.gpchan = (RefGpChan * 12).in_dll(L, "gpchan")
        self.acq_thresh = C.c_int.in_dll(L, "acq_thresh")
        self.freq_bin_width = C.c_double.in_dll(L, "freq_bin_width")
        self.gps_carrier_ref = C.c_long.in_dll(L, "gps_carrier_ref")
        self.gps_code_ref = C.c_long.in_dll(L, "gps_code_ref")
        self.d_freq = C.c_long.in_dll(L, "d_freq")
        self.use_iq = C.c_int.in_dll(L, "use_iq_processing")
        self.corr_out = C.c_void_p.in_dll(L, "corr_out")
        self.reset(tic_period)

    def reset(self, tic_period: float | None = None):
        L = self.L
        C.memset(self.REG_read, 0, 1024)
        C.memset(self.REG_write, 0, 1024)
        C.memset(self.chan, 0, C.sizeof(self.chan))
        L.init_tracking_loops_parameter()
        # stock main passes the int global tic_period (0.1 truncated to 0), osgnss_next_step.c:147
        tp = float(C.c_int.in_dll(L, "tic_period").value) if tic_period is None else tic_period
        L.correlator_init(C.c_double(tp))
        # corr_out is only written by output_test_data(); give it a sink
        libc = C.CDLL(None)
        libc.fopen.restype = C.c_void_p
        self.corr_out.value = libc.fopen(b"/dev/null", b"w")

    def cold_allocate(self, prns):
        self.L.reset_all_correlator_channles()
        for ch, p in enumerate(prns):
            if p > 0:
                self.L.ch_cntl(ch, p)

    def sim(self, iq: np.ndarray, nsamp: int):
        buf = np.ascontiguousarray(iq, dtype=np.int8)
        self.L.Sim_GP2021_int(buf.ctypes.data, nsamp)

    def gpsisr(self):
        self.L.gpsisr()

    def regs(self):
        return np.array(self.REG_read[:], dtype=np.int32), np.array(self.REG_write[:], dtype=np.int32)

    def run(self, iq: np.ndarray, nsamp: int, nblocks: int, dump_cap: int = 0, block0: int = 0):
        """Native loop (oracle/ref_driver.c) around the reference's Sim_GP2021_int + gpsisr."""
        buf = np.ascontiguousarray(iq, dtype=np.int8)
        assert buf.size >= 2 * nsamp * nblocks
        self.L.ref_run.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_long, C.c_void_p, C.c_int, C.c_void_p]
        self.L.ref_run.restype = C.c_long
        if dump_cap:
            dumps = np.zeros((abi.N_CHANNELS, dump_cap), dtype=abi.DUMP_DTYPE)
            cnt = np.zeros(abi.N_CHANNELS, dtype=np.int32)
            n = self.L.ref_run(buf.ctypes.data, nsamp, nblocks, block0, dumps.ctypes.data, dump_cap, cnt.ctypes.data)
            return n, dumps, cnt
        n = self.L.ref_run(buf.ctypes.data, nsamp, nblocks, block0, None, 0, None)
        return n, None, None

    def warm_start(self, ch: int, n_freq: int):
        k = self.chan[ch]
        k.n_freq = n_freq
        k.del_freq = -2 * n_freq if n_freq > 0 else 1 - 2 * n_freq
        k.carrier_freq = self.gps_carrier_ref.value + k.carrier_cold_corr + self.d_freq.value * n_freq
        k.codes = 0
        self.L.ch_carrier(ch, k.carrier_freq)
