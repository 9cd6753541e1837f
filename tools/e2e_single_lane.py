import sys, time, os
sys.path[:0]=['coeb-slam_b200/python']
import numpy as np, coeb_b200 as cb
from coeb_b200 import synth
B=256
b=synth.make_batch(B, base_seed=0, unique=16)
pin={}
owners=[]
for k in b:
    a,o=cb.pinned_array(b[k].shape,b[k].dtype); a[...]=b[k]; pin[k]=a; owners.append(o)
ex=cb.Extractor(1000,1.2,8,20,7,device=0)
cap=ex.default_cap()
outs=[cb.pinned_array((B,cap),cb.KP_DTYPE)[0], cb.pinned_array((B,cap,32),np.uint8)[0], cb.pinned_array((B,),np.int32)[0], cb.pinned_array((B,),np.int32)[0]]
for _ in range(4): ex.extract_batch_host(pin['gray'],pin['boxes'],pin['nbox'],pin['tm'],pin['ntm'],pin['blur'],cap=cap,out=outs)
t=time.perf_counter(); n=int(os.environ.get("N_CALLS","20"))
for _ in range(n): ex.extract_batch_host(pin['gray'],pin['boxes'],pin['nbox'],pin['tm'],pin['ntm'],pin['blur'],cap=cap,out=outs)
dt=time.perf_counter()-t
print(os.environ.get('COEB_PIPE_CHUNK'), 'single lane frames/s', B*n/dt)
