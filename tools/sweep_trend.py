import sys; sys.path.insert(0,'.')
import numpy as np
from insider_b200 import synth
from oracle import oracle
P=int(sys.argv[1]) if len(sys.argv)>1 else 1000
pb=synth.ageing_like(N=377,P=P,K=23)
tr,te=synth.random_masks(377,P)
F0,V0=synth.init_factors(pb.levels,23,P)
prev=0; previt=0
for iters in (30,60,100,150,200,300,400):
    r=oracle.optimize(pb.Y,F0,V0,pb.confounder,None,tr,te,0,23,10.0,10.0,0.4,0,1e-9,1e-5,iters,perm_mode=1,seed=1)
    print(f"iters={r.iters_run}: sweeps/gene-iter since last {(r.cd_sweeps-prev)/P/max(1,(r.iters_run-previt)):.2f}; loss {r.loss:.8g} delta {r.checks[-1]['delta_loss']:.3g} decay {r.checks[-1]['decay']} t={r.seconds_in_loop:.1f}s", flush=True)
    prev=r.cd_sweeps; previt=r.iters_run
    if r.iters_run < iters: print("converged"); break
