import os, sys, time
sys.path.insert(0,'/root/repo')
import numpy as np, torch
import bench
from t41_sdr_b200 import rx
for flags in (0, 4):
    S,T=16384,32
    params,sigs=bench.workload(T,"c4")
    D=len(params)
    iq=torch.from_numpy(np.stack(sigs)).cuda().index_select(0, torch.arange(S,device="cuda")%D).contiguous()
    audio=torch.empty((S,T,2048),dtype=torch.float32,device="cuda")
    spec=torch.empty((S,T,512),dtype=torch.int16,device="cuda"); wf=torch.empty((S,T,512),dtype=torch.int16,device="cuda")
    with rx.Receiver(S) as eng:
        eng.set_params_each([params[s%D] for s in range(S)])
        for _ in range(2): eng.process_device(iq.data_ptr(),audio.data_ptr(),T,1,spec.data_ptr(),wf.data_ptr(),None,None,flags)
        eng.synchronize(); t0=time.perf_counter()
        for _ in range(3): eng.process_device(iq.data_ptr(),audio.data_ptr(),T,1,spec.data_ptr(),wf.data_ptr(),None,None,flags)
        eng.synchronize(); dt=(time.perf_counter()-t0)/3
    print("flags",flags,"ms",dt*1e3,"Mrows/s",S*T/dt/1e6)
    del iq,audio,spec,wf
