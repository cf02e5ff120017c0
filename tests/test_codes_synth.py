"""Code generators, sampled code tables and the packed 2-bit format (host side, no GPU)."""
import numpy as np
import pytest

from gnss_sdr_ru_b200.codes import ca_code, st_code
from gnss_sdr_ru_b200.synth import Sat, make_record, pack2, unpack2


def test_ca_code_known_answers():
    # IS-GPS-200: first 10 chips of PRN 1 are 1100100000 (octal 1440); balanced Gold code
    assert "".join(str((c + 1) // 2) for c in ca_code(1)[:10]) == "1100100000"
    assert "".join(str((c + 1) // 2) for c in ca_code(2)[:10]) == "1110010000"  # octal 1620
    for prn in range(1, 33):
        assert int(ca_code(prn).sum()) in (-1, 1, 63, -65, 65, -63)


def test_ca_code_equals_reference_c_generator(oracle_lib):
    """oracle tables follow OSG/correlator/correlator.c:63-91; early[h] = chip[h>>1] etc."""
    for prn in (1, 7, 19, 32):
        c = ca_code(prn)
        for h in (0, 1, 2, 3, 100, 1023, 2043, 2044, 2045):
            e, p, l = oracle_lib.Oracle.code_bits(prn, h)
            assert e == c[(h % 2046) >> 1] and p == c[((h + 1) % 2046) >> 1] and l == c[((h + 2) % 2046) >> 1]
    # spill-over rule: index 2046 of PRN p is entry 0 of PRN p+1; past row 33 -> 0
    assert oracle_lib.Oracle.code_bits(5, 2046) == oracle_lib.Oracle.code_bits(6, 0)
    assert oracle_lib.Oracle.code_bits(32, 2046) == (0, 0, 0)


def test_st_code_is_m_sequence():
    s = st_code()
    assert len(s) == 511 and int(s.sum()) == 1
    # two-valued autocorrelation of a maximal-length sequence
    x = s.astype(int)
    for lag in (1, 7, 100, 255):
        assert int(np.dot(x, np.roll(x, lag))) == -1


def test_pack_roundtrip_and_layout():
    rng = np.random.default_rng(1)
    iq = rng.choice(np.array([-3, -1, 1, 3], dtype=np.int8), size=4096)
    p = pack2(iq)
    assert p.dtype == np.uint8 and p.size == 1024
    assert np.array_equal(unpack2(p), iq)
    # LSB-first, I then Q; codes {0:+1,1:-1,2:+3,3:-3} (FE/.../win32_sampler.h:45-55)
    assert pack2(np.array([1, -1, 3, -3], dtype=np.int8))[0] == (0 | (1 << 2) | (2 << 4) | (3 << 6))
    with pytest.raises(ValueError):
        pack2(np.array([1, 2, 1, 1], dtype=np.int8))
    with pytest.raises(ValueError):
        pack2(np.array([1, 1, 1], dtype=np.int8))
    assert unpack2(np.zeros(0, dtype=np.uint8)).size == 0


def test_record_is_deterministic_and_four_level():
    sats = [Sat(prn=3, doppler_hz=1500.0, code_phase_chips=10.0, data_seed=3)]
    a = make_record(sats, 20000, seed=9)
    b = make_record(sats, 20000, seed=9, chunk=7777)  # chunking must not change the signal part
    assert set(np.unique(a)) <= {-3, -1, 1, 3}
    assert a.shape == (40000,)
    # different chunking draws noise in a different order, but the noiseless records must agree exactly
    a0 = make_record(sats, 20000, seed=9, noise=False)
    b0 = make_record(sats, 20000, seed=9, noise=False, chunk=7777)
    assert np.array_equal(a0, b0)
    assert not np.array_equal(a, make_record(sats, 20000, seed=10))
    del b
