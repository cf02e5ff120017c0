# ad-hoc probe: time the acquisition kernels on a device-resident record
import sys, os, time, ctypes as C, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnss_sdr_ru_b200 import abi
from gnss_sdr_ru_b200.acquisition import AcquisitionEngine, Settings
from gnss_sdr_ru_b200.lib import lib, check
from gnss_sdr_ru_b200.scenarios import gps_weak_acq_scenario, TrackScenario, synth_sat_array
coh=int(sys.argv[1]) if len(sys.argv)>1 else 10
K=int(sys.argv[2]) if len(sys.argv)>2 else 20
nsv=int(sys.argv[3]) if len(sys.argv)>3 else 32
st=Settings.gps(acqSearchBand=20.0, acqCohIntegration=coh, n_noncoh=K, acqSatelliteList=list(range(1,nsv+1)))
ae=AcquisitionEngine()
n=ae.samples_needed(st); n4=(n+3)//4*4
rec=torch.empty(2*n4,dtype=torch.uint8,device='cuda')
arr,nsat=synth_sat_array([TrackScenario(sats=gps_weak_acq_scenario(4004),prns=[],n_freq=[])])
check(lib().gnssb200_synth(ae.h, rec.data_ptr(), 2*n4, abi.FMT_INT8_IQ, 1, n4, C.addressof(arr), nsat, 4004, None),'synth')
nb=ae.num_bins(st)
rows=torch.zeros(nsv*nb*16,dtype=torch.uint8,device='cuda')
for it in range(3):
    torch.cuda.synchronize(); t=time.time()
    ae.search_device(rec.data_ptr(), n4, st, rows.data_ptr())
    torch.cuda.synchronize(); dt=time.time()-t
    nfft=nsv*nb*max(K,2)
    print(f'coh={coh} K={K} nsv={nsv} bins={nb}: wall {dt*1e3:.3f} ms kernel {ae.last_kernel_ms():.3f} ms -> {nsv*nb*16000/dt/1e9:.2f} Gcells/s, {ae.last_kernel_ms()*1e6/nfft*148/1e3:.2f} us per IFFT per SM')
