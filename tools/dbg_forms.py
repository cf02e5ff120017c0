"""debug helper: one stream, packed input, selected kernel forms vs the oracle; prints the records around the first mismatch"""
import sys
import numpy as np
sys.path.insert(0, ".")
from gnss_sdr_ru_b200 import abi
from gnss_sdr_ru_b200.receiver import TrackingEngine
from gnss_sdr_ru_b200.synth import Sat, make_record, pack2
from oracle import oracle_api

NS = 8192
nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 300
forms = [tuple(int(x) for x in f.split(",")) for f in sys.argv[2:]] or [(2, 0), (5, 0), (4, 0), (3, 0)]
sats = [Sat(prn=27, doppler_hz=1200, cn0_dbhz=50, code_phase_chips=1000.3, data_seed=5),
        Sat(prn=9, doppler_hz=-900, cn0_dbhz=47, code_phase_chips=980.0, data_seed=6),
        Sat(prn=32, doppler_hz=1000, cn0_dbhz=49, code_phase_chips=1010.0, data_seed=7)]
rec = make_record(sats, NS * nblk, seed=2002)
PRNS = [27, 0, 0, 31, 0, 0, 0, 0, 9, 0, 32, 5]
WARM = ((0, 1), (8, -1), (10, 1))
oracle_api.build()
o = oracle_api.Oracle()
o.cold_allocate(PRNS)
for ch, nf in WARM:
    k = o.rx.chan[ch]
    k.n_freq, k.del_freq, k.codes = nf, (-2 * nf if nf > 0 else 1 - 2 * nf), 0
    k.carrier_freq = o.cfg.gps_carrier_ref + o.cfg.d_freq * nf
    o.ch_carrier(ch, k.carrier_freq)
_, od, oc = o.run(rec, NS, nblk, dump_cap=2000)
pk = pack2(rec)[None, :]
for form, occ in forms:
    eng = TrackingEngine(n_streams=1)
    eng.simple_cold_allocate(0, PRNS)
    for ch, nf in WARM:
        eng.warm_start(0, ch, nf)
    eng.upload()
    eng.set_track_variant(form, occ)
    d, c = eng.run_host(pk, nblk, NS, abi.FMT_PACKED2, dump_cap=2000)
    bad = None
    for ch in range(12):
        n = min(c[0, ch], oc[ch])
        a, b = d[0, ch, :n], od[ch, :n]
        ne = np.nonzero(a != b)[0]
        if len(ne) or c[0, ch] != oc[ch]:
            bad = (ch, int(ne[0]) if len(ne) else n)
            break
    print(f"form {form} occ {occ}: counts equal {np.array_equal(c[0], oc)}; first mismatch {bad}")
    if bad:
        ch, i = bad
        for j in range(max(0, i - 3), min(i + 4, c[0, ch])):
            print("   gpu", j, d[0, ch, j])
            print("   ora", j, od[ch, j])
