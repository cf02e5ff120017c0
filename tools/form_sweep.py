# ad-hoc: tracking throughput (packed input, 2 s) vs number of streams for each kernel form / occupancy variant / slice length
import sys, ctypes as C, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from gnss_sdr_ru_b200 import abi
from gnss_sdr_ru_b200.receiver import TrackingEngine
from gnss_sdr_ru_b200.scenarios import gps_tracking_scenario, synth_sat_array, apply_tracking_scenario
from gnss_sdr_ru_b200.lib import lib, check
NS, nblk = 8192, 3906
streams = [int(x) for x in sys.argv[1].split(",")]
variants = [tuple(int(y) for y in x.split(":")) for x in sys.argv[2].split(",")]  # form:occ:slice
for S in streams:
    eng = TrackingEngine(n_streams=S)
    scs = [gps_tracking_scenario(5000 + s) for s in range(S)]
    buf = torch.empty((S, NS * nblk // 2), dtype=torch.uint8, device='cuda')
    arr, nsat = synth_sat_array(scs)
    check(lib().gnssb200_synth(eng.h, buf.data_ptr(), buf.stride(0), abi.FMT_PACKED2, S, NS * nblk, C.addressof(arr), nsat, 1234, None), 'synth')
    out = []
    for form, occ, sl in variants:
        eng.set_track_variant(form, occ)
        eng.set_track_slice(sl)
        best = None
        for rep in range(3):
            for s in range(S):
                eng.L.gnssb200_rx_init(C.byref(eng.rx[s]), C.byref(eng.cfg)); apply_tracking_scenario(eng, s, scs[s])
            eng.upload()
            eng.run_device(buf.data_ptr(), buf.stride(0), nblk, NS, abi.FMT_PACKED2)
            torch.cuda.synchronize()
            ms = eng.last_kernel_ms()
            best = ms if best is None or ms < best else best
        out.append(f"{form}:{occ}:{sl}={S*12*NS*nblk/(best*1e-3)/1e9:.0f}k")
    print(f"S={S:4d}", "  ".join(out), flush=True)
    eng.close(); del buf
