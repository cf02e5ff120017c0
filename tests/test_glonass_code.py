"""GLONASS channels of the integer correlator (an extension: the C reference has the hooks only, include/gnssb200.h
GNSSB200_PRN_GLONASS).  CPU part: the ST-code row of the E/P/L table against a restatement of the Namuru RTL generator
(NAM/rtl/code_gen.v:121-133) and against the Scilab generator (generateSTcode.sci), the GLONASS reference words
against the values the compiled reference derives (correlator.c:116-118), and the library's cold allocation of such a
channel against the oracle's."""
import ctypes as C

import numpy as np

from gnss_sdr_ru_b200 import abi


def _rtl_g3_chips():
    """code_gen.v:121-133: g3 <= 9'b111111111 at the PRN-key write; per full-chip enable g3_q <= g3[2],
    g3 <= {g3[4]^g3[0], g3[8:1]}."""
    g3 = [1] * 9  # g3[0] .. g3[8]
    out = []
    for _ in range(511):
        out.append(g3[2])
        fb = g3[4] ^ g3[0]
        g3 = g3[1:] + [fb]
    return np.array(out)


def test_st_code_row_matches_rtl_and_scilab(oracle_lib):
    from gnss_sdr_ru_b200.codes import st_code

    L = oracle_lib.Oracle.lib()
    chips = _rtl_g3_chips()
    assert chips.sum() == 256  # m-sequence of length 511: 256 ones
    pm = 2 * chips - 1
    sc = st_code()  # +-1, generateSTcode.sci:35-42
    assert np.array_equal(pm, sc) or np.array_equal(pm, -sc)
    for which in range(3):
        got = np.array([L.orc_code_bit(which, abi.PRN_GLONASS, h) for h in range(1030)])
        want = np.array([pm[((h + which) % 1022) >> 1] for h in range(1022)])  # correlator.c:85-89 with 1022 for 2046
        assert np.array_equal(got[:1022], want)
        assert not got[1022:].any()  # past the row: zeros
    # the C/A rows did not move: PRN 1 starts 1100100000 (IS-GPS-200), chip 0 forced to 1 (correlator.c:75)
    e = [L.orc_code_bit(0, 1, 2 * k) for k in range(10)]
    assert e == [1, 1, -1, -1, 1, -1, -1, -1, -1, -1]
    assert L.orc_code_bit(0, 34, 5) == 0 and L.orc_code_bit(0, 35, 5) == 0  # not addressable as PRN 34 / 35


def test_glonass_reference_words(oracle_lib):
    cfg = oracle_lib.Oracle.default_cfg()
    assert int(cfg.glonass_code_ref) == 3429262 and int(cfg.glonass_carrier_ref) == 0  # SURVEY 8a row A4 (probe of the compiled reference)
    if oracle_lib.have_ref():
        ref = oracle_lib.RefReceiver()
        assert C.c_long.in_dll(ref.L, "glonass_code_ref").value == int(cfg.glonass_code_ref)
        assert C.c_long.in_dll(ref.L, "glonass_carrier_ref").value == int(cfg.glonass_carrier_ref)
    cfg2 = oracle_lib.Oracle.default_cfg(glonass_carrier_if=1.0e6)
    assert int(cfg2.glonass_carrier_ref) == int(1.0e6 / (5 * 16e6 / 2**30))


def test_cold_allocation_of_a_glonass_channel_equals_oracle(oracle_lib):
    from gnss_sdr_ru_b200.lib import default_cfg, lib

    L = lib()
    cfg = default_cfg(glonass_carrier_if=1.0e6)
    prns = [27, abi.PRN_GLONASS, 0, 9, abi.PRN_GLONASS, 0, 0, 0, 0, 0, 0, 32]
    rx = abi.Rx()
    L.gnssb200_rx_init(C.byref(rx), C.byref(cfg))
    L.gnssb200_rx_cold_allocate(C.byref(rx), C.byref(cfg), (C.c_int32 * 12)(*prns))
    o = oracle_lib.Oracle(oracle_lib.Oracle.default_cfg(glonass_carrier_if=1.0e6))
    o.cold_allocate(prns)
    assert bytes(memoryview(rx).cast("B")) == bytes(memoryview(o.rx).cast("B"))
    assert rx.chan[1].system == 1 and rx.chan[1].search_max_PRN_delay == 1021 and rx.chan[0].system == 0
    assert rx.reg_write[1 << 3] == abi.PRN_GLONASS
