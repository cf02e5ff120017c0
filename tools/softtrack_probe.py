# ad-hoc probe (not the bench): Scilab float tracking, 8 GLONASS channels x 1 s on a device-generated record
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnss_sdr_ru_b200 import abi
from gnss_sdr_ru_b200.lib import check, lib
from gnss_sdr_ru_b200.receiver import TrackingEngine
from gnss_sdr_ru_b200.scenarios import TrackScenario, synth_sat_array
from gnss_sdr_ru_b200.softtrack import SoftTrackingEngine, TrackSettings
from gnss_sdr_ru_b200.synth import Sat
L = lib()
eng = TrackingEngine(n_streams=1)
fsats = [Sat(system="glonass", prn=k, cn0_dbhz=48.0, doppler_hz=400.0 * k, code_phase_chips=37.0 * (k + 8), data_seed=50 + k, data_rate_hz=100.0)
         for k in (-7, -5, -3, -1, 0, 2, 4, 6)]
fms = 1000
fn = 16000 * (fms + 8)
frec = torch.empty(2 * fn, dtype=torch.uint8, device="cuda")
farr, fns = synth_sat_array([TrackScenario(sats=fsats, prns=[], n_freq=[])])
check(L.gnssb200_synth(eng.h, frec.data_ptr(), 2 * fn, abi.FMT_INT8_IQ, 1, fn, C.addressof(farr), fns, 77, None), "synth")
fchan = [dict(FCH=s.prn, acquiredFreq=1e6 + 562500.0 * s.prn + s.doppler_hz + 30.0,
              codePhase=int(round((511.0 - s.code_phase_chips % 511.0) * 16000.0 / 511.0)) % 16000 + 1) for s in fsats]
fout = torch.zeros((len(fchan), fms, 13), dtype=torch.float64, device="cuda")
fdone = torch.zeros(len(fchan), dtype=torch.int32, device="cuda")
ste = SoftTrackingEngine(handle=eng.h)
for it in range(3):
    ste.tracking_device(frec.data_ptr(), fn, fchan, TrackSettings(msToProcess=fms), fout.data_ptr(), fdone.data_ptr())
    print("softtrack 8 ch x %d ms: kernel %.2f ms" % (fms, ste.last_kernel_ms()))
