"""Regenerates the committed golden fixtures.  Run HERE (needs /root/reference); the fixtures travel.

  lo_golden.npz         the reference's own golden LO sequences NAM/sci/i_carr.dat, q_carr.dat
                        (Verilator dump of the Namuru carrier NCO, tb_carrier_nco.cpp:43,52-55),
                        decoded with the mapping of NAM/sci/sci_carrier_nco.sce:8-19.
  ref_track_golden.npz  outputs of the REFERENCE C receiver itself (oracle/_ref, compiled from the
                        reference's sources) on a seeded synthetic record: every dump record, final
                        REG_read / REG_write.  The packed record is stored so nothing depends on libm.
  acq_golden.npz        outputs of the NumPy restatement of acquisition.sci on a seeded record
                        (regression fixture; parity for this path is unpinned, see oracle/pcps_oracle.py).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from gnss_sdr_ru_b200.synth import Sat, make_record, pack2  # noqa: E402
from oracle import oracle_api, pcps_oracle  # noqa: E402

NAM = "/root/reference/trunk/Verilog_VERIFICATION_PROJECTS/namuru/sci"


def lo():
    dec = np.array([-1, -2, 1, 2], dtype=np.int8)
    i = dec[np.loadtxt(os.path.join(NAM, "i_carr.dat"), dtype=np.int64)]
    q = dec[np.loadtxt(os.path.join(NAM, "q_carr.dat"), dtype=np.int64)]
    np.savez_compressed(os.path.join(HERE, "lo_golden.npz"), i=i, q=q, f_control=np.uint32(0x0318FC50))


def track():
    NS, nblk = 8192, 120
    sats = [Sat(prn=27, doppler_hz=1200, cn0_dbhz=50, code_phase_chips=1015.3, data_seed=5),
            Sat(prn=9, doppler_hz=-900, cn0_dbhz=47, code_phase_chips=1018.0, data_seed=6),
            Sat(prn=32, doppler_hz=1000, cn0_dbhz=49, code_phase_chips=1019.0, data_seed=7)]
    rec = make_record(sats, NS * nblk, seed=424242)
    prns = [27, 0, 0, 31, 0, 0, 0, 0, 9, 0, 32, 5]
    warm = [(0, 1), (8, -1), (10, 1)]
    ref = oracle_api.RefReceiver()
    ref.cold_allocate(prns)
    for ch, n in warm:
        ref.warm_start(ch, n)
    n, dumps, cnt = ref.run(rec, NS, nblk, dump_cap=80)
    rr, rw = ref.regs()
    np.savez_compressed(os.path.join(HERE, "ref_track_golden.npz"), packed=pack2(rec), nblk=nblk, nsamp=NS,
                        prns=np.array(prns), warm=np.array(warm), dumps=dumps, cnt=cnt, reg_read=rr, reg_write=rw)


def acq():
    sats = [Sat(prn=5, doppler_hz=2300.0, cn0_dbhz=50, code_phase_chips=321.5),
            Sat(prn=17, doppler_hz=-3100.0, cn0_dbhz=46, code_phase_chips=77.25)]
    rec = make_record(sats, 16000 * 3, seed=77)
    s = pcps_oracle.AcqSettings.gps(acqSearchBand=8.0, acqCohIntegration=1, svList=[5, 6, 17])
    res = pcps_oracle.acquisition(pcps_oracle.to_complex(rec), s)
    nb = pcps_oracle.num_bins(s)
    rows = np.array([[r["rows"][b + 1] for b in range(nb)] for r in res])  # (sv, bin, [peak, arg, blk])
    np.savez_compressed(os.path.join(HERE, "acq_golden.npz"), packed=pack2(rec), sv=np.array(s.svList), band=8.0, coh=1,
                        peakMetric=np.array([r["peakMetric"] for r in res]), bin=np.array([r["bin"] for r in res]),
                        codePhaseRaw=np.array([r["codePhaseRaw"] for r in res]), codePhase=np.array([r["codePhase"] for r in res]),
                        carrFreq=np.array([r["carrFreq"] for r in res]), rows=rows)


if __name__ == "__main__":
    lo()
    track()
    acq()
    print({f: os.path.getsize(os.path.join(HERE, f)) for f in sorted(os.listdir(HERE)) if f.endswith(".npz")})
