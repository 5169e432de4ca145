import time, numpy as np, sys
sys.path.insert(0,'tests')
from helpers import make_case, disp_args, beta_args
from deseq2_b200 import wrappers as W
c=make_case(50000,100,seed=5)
a=disp_args(c,c["mu"],np.log(c["alpha0"]))
for i in range(3):
    t=time.time(); g=W.fitDisp(**a); dt=time.time()-t
    print("fitDisp 50k x100 host-API s",dt, "genes/s", len(c["counts"])/dt, "mean iter", g["iter"].mean())
alpha=np.clip(0.1+4/c["baseMean"],1e-8,10)
b=beta_args(c,alpha)
for i in range(3):
    t=time.time(); r=W.fitBeta(**b); dt=time.time()-t
    print("fitBeta host-API s",dt,"genes/s",len(c["counts"])/dt,"mean iter",r["iter"].mean())
