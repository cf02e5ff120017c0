# prints a SHA-256 over every dump record and the final receiver states of a synthetic multi-stream run;
# used by tests/test_track_gpu.py::test_config5_width_kernels_agree to compare independent kernel variants
# (environment: GNSSB200_TRACK_WS=0 selects the barrier-synchronised predecessor kernel)
import ctypes as C, hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gnss_sdr_ru_b200 import abi
from gnss_sdr_ru_b200.lib import check, lib
from gnss_sdr_ru_b200.receiver import TrackingEngine
from gnss_sdr_ru_b200.scenarios import apply_tracking_scenario, gps_tracking_scenario, synth_sat_array

S, nblk, slice_blocks = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
NS = 8192
eng = TrackingEngine(n_streams=S)
scs = [gps_tracking_scenario(9000 + s) for s in range(S)]
for s in range(S):
    apply_tracking_scenario(eng, s, scs[s])
eng.upload()
eng.set_track_slice(slice_blocks)
d_if = torch.empty((S, NS * nblk // 2), dtype=torch.uint8, device="cuda")
arr, nsat = synth_sat_array(scs)
check(lib().gnssb200_synth(eng.h, d_if.data_ptr(), d_if.stride(0), abi.FMT_PACKED2, S, NS * nblk, C.addressof(arr), nsat, 31337, None), "synth")
cap = nblk // 2 + 64
d_d = torch.zeros((S, 12, cap, 48), dtype=torch.uint8, device="cuda")
d_c = torch.zeros((S, 12), dtype=torch.int32, device="cuda")
eng.run_device(d_if.data_ptr(), d_if.stride(0), nblk, NS, abi.FMT_PACKED2, d_dumps_ptr=d_d.data_ptr(), dump_cap=cap, d_count_ptr=d_c.data_ptr())
torch.cuda.synchronize()
eng.download()
cnt = d_c.cpu().numpy()
dumps = d_d.cpu().numpy()
h = hashlib.sha256()
h.update(cnt.tobytes())
for s in range(S):
    for ch in range(12):
        h.update(dumps[s, ch, : cnt[s, ch]].tobytes())
    h.update(bytes(memoryview(eng.rx[s]).cast("B")))
print("DIGEST", h.hexdigest(), int(cnt.sum()))
