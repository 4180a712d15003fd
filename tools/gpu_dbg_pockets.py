import os, sys
ROOT='/root/repo'
for p in (ROOT, os.path.join(ROOT, "netl-ap-map-flow_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
from helpers import synth_aperture, anchored_open_mask
from apmapflow_b200 import lcl
from apmapflow_b200._lib import Params, TOP
n=2048
a = synth_aperture(n, n, seed=7, sigma_c=8, contact=0.2, sigma_k=16)
p = Params.defaults(inlet_pressure=1000.0, outlet_pressure=0.0, density=1000.0, viscosity=1.0e-3, voxel_size=10.0e-6, avg_factor=1.0, roughness=0.0, high_mask=100.0, low_mask=1e-4, outlet_side=TOP)
f,s=lcl.lcl_solve(p,a,fields=("pressure",))
P=f["pressure"]
print("iters",s["iterations"],"res",s["rel_residual"],s["rel_residual_scaled"],"min",P.min(),"max",P.max())
bad=np.argwhere((P>1000.0+1e-6)|(P<-1e-6))
print("cells outside [0,1000]:",len(bad))
main=anchored_open_mask(a)
print("of them anchored-open:", int(main[bad[:,0],bad[:,1]].sum()), " contact cells:", int((a[bad[:,0],bad[:,1]]==0).sum()))
if len(bad):
    z,x=bad[np.argmax(P[bad[:,0],bad[:,1]])]
    print("worst at",z,x,"a=",a[z,x]); 
    sub=a[z-6:z+7,x-10:x+11]
    for r_,pr in zip(sub,P[z-6:z+7,x-10:x+11]):
        print("".join('#' if v==0 else '.' for v in r_), " ".join("%6.0f"%v for v in pr[::2]))
