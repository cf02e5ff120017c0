# ad-hoc timing probe (not the bench): tracking throughput vs streams / SPT
import sys, time, ctypes as C, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from gnss_sdr_ru_b200 import abi
from gnss_sdr_ru_b200.receiver import TrackingEngine
from gnss_sdr_ru_b200.scenarios import gps_tracking_scenario, synth_sat_array, apply_tracking_scenario
from gnss_sdr_ru_b200.lib import lib, check
S=int(sys.argv[1]); nblk=int(sys.argv[2]); fmt=int(sys.argv[3]) if len(sys.argv)>3 else 0
NS=8192
eng=TrackingEngine(n_streams=S)
scs=[gps_tracking_scenario(5000+s) for s in range(S)]
bytes_per=NS*nblk*2 if fmt==0 else NS*nblk//2
buf=torch.empty((S,bytes_per),dtype=torch.uint8,device='cuda')
arr,nsat=synth_sat_array(scs)
t=time.time()
check(lib().gnssb200_synth(eng.h, buf.data_ptr(), buf.stride(0), fmt, S, NS*nblk, C.addressof(arr), nsat, 1234, None),'synth')
torch.cuda.synchronize(); print('synth s',time.time()-t)
for rep in range(3):
    for s in range(S): 
        eng.L.gnssb200_rx_init(C.byref(eng.rx[s]), C.byref(eng.cfg)); apply_tracking_scenario(eng,s,scs[s])
    eng.upload()
    torch.cuda.synchronize(); t=time.time()
    eng.run_device(buf.data_ptr(), buf.stride(0), nblk, NS, fmt)
    torch.cuda.synchronize(); dt=time.time()-t
    print(f'S={S} nblk={nblk} fmt={fmt} wall {dt*1e3:.2f} ms kernel {eng.last_kernel_ms():.2f} ms  -> {S*12*NS*nblk/dt/1e6:.0f} ch*Msamples/s, per-block {eng.last_kernel_ms()*1e3/nblk:.2f} us')
eng.download()
print('states', [[eng.rx[s].chan[ch].state for ch in range(12)] for s in range(min(S,3))])
