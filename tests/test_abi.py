"""The C-ABI library loads without a GPU and exports every symbol include/gnssb200.h declares."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gnssb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    funcs = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text))
    funcs -= {"defined", "sizeof", "__attribute__"}
    funcs = {f for f in funcs if f.startswith("gnssb200_") or f in ("correlator_init", "Sim_GP2021_int")}
    return sorted(funcs), ["REG_read", "REG_write"]


def test_library_exports_every_declared_symbol():
    from gnss_sdr_ru_b200 import lib

    if not os.path.exists(lib.SO_PATH):
        lib.build()
    L = C.CDLL(lib.SO_PATH)
    funcs, data = _declared_symbols()
    assert "gnssb200_track_run" in funcs and "gnssb200_acq_search" in funcs and "Sim_GP2021_int" in funcs and "gnssb200_softtrack" in funcs
    for f in funcs:
        assert hasattr(L, f), f"{f} declared in include/gnssb200.h but not exported"
    for d in data:
        (C.c_int * 256).in_dll(L, d)


def test_struct_layouts_match_header():
    """ctypes mirrors (gnss_sdr_ru_b200/abi.py) vs sizeof() from the header compiled with gcc"""
    import subprocess
    import tempfile

    from gnss_sdr_ru_b200 import abi

    src = r'''
#include <stdio.h>
#include "gnssb200.h"
int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n",sizeof(gnssb200_cfg),sizeof(gnssb200_chan),sizeof(gnssb200_corr),
 sizeof(gnssb200_rx),sizeof(gnssb200_dump),sizeof(gnssb200_acq_cfg),sizeof(gnssb200_acq_row),sizeof(gnssb200_acq_result),sizeof(gnssb200_synth_sat),sizeof(gnssb200_softtrack_cfg),sizeof(gnssb200_softtrack_chan),sizeof(gnssb200_ingest_stat),sizeof(gnssb200_gpssdr_result),sizeof(gnssb200_serial_cell));return 0;}
'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")])
        out = subprocess.check_output([os.path.join(d, "s")]).split()
    want = [C.sizeof(x) for x in (abi.Cfg, abi.Chan, abi.Corr, abi.Rx, abi.Dump, abi.AcqCfg, abi.AcqRow, abi.AcqResult, abi.SynthSat, abi.SoftTrackCfg, abi.SoftTrackChan, abi.IngestStat, abi.GpsSdrResult, abi.SerialCell)]
    assert [int(x) for x in out] == want


def test_host_helpers_match_oracle(oracle_lib):
    """cfg derivation and register accessors of the library (host code, no device) vs the oracle"""
    from gnss_sdr_ru_b200 import abi, lib

    L = lib.lib()
    cfg = lib.default_cfg()
    ocfg = oracle_lib.Oracle.default_cfg()
    assert bytes(memoryview(cfg).cast("B")) == bytes(memoryview(ocfg).cast("B"))
    rx, o = abi.Rx(), oracle_lib.Oracle()
    L.gnssb200_rx_init(C.byref(rx), C.byref(cfg))
    prns = (C.c_int32 * 12)(27, 0, 3, 0, 0, 0, 0, 0, 9, 0, 32, 5)
    L.gnssb200_rx_cold_allocate(C.byref(rx), C.byref(cfg), prns)
    o.cold_allocate(list(prns))
    for ch, f in ((0, 32480690 + 13421), (5, -12345), (11, 2**31 + 17)):
        L.gnssb200_ch_carrier(C.byref(rx), C.byref(cfg), ch, f)
        o.ch_carrier(ch, f)
        L.gnssb200_ch_code(C.byref(rx), C.byref(cfg), ch, f // 3)
        o.ch_code(ch, f // 3)
    L.gnssb200_ch_code_slew(C.byref(rx), 2, 70000)
    o.ch_code_slew(2, 70000)
    L.gnssb200_ch_epoch_load(C.byref(rx), 4, 0x1234)
    o.ch_epoch_load(4, 0x1234)
    assert bytes(memoryview(rx).cast("B")) == bytes(memoryview(o.rx).cast("B"))


def test_no_cpu_fallback():
    """without a CUDA device the library must fail loudly, not compute on the host"""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gnss_sdr_ru_b200 import lib
    from gnss_sdr_ru_b200.receiver import TrackingEngine

    with pytest.raises(lib.GnssB200Error):
        TrackingEngine(n_streams=1)
    assert lib.lib().gnssb200_last_error() != 0
