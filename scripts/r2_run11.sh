#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
python - <<'PY' 2>&1 | tail -12
import sys, time, numpy as np
sys.path.insert(0,'.')
import gpcc_b200, os
ctx=gpcc_b200.Context(1,profiling=True)
t,y,s,_=gpcc_b200.synthetic_bands([2048]*3,seed=4)
p=gpcc_b200.Problem(t,y,s,"matern52",ctx)
rg=np.random.default_rng(2); M=128
d=np.zeros((M,3)); d[:,1:]=rg.uniform(0,19.8,(M,2)); a=np.tile([1.0,2.2,4.0],(M,1)); r=np.full(M,3.5)
p.loglik_batch(d[:32],a[:32],r[:32])
for grad in (False,True):
    t0=time.perf_counter(); out=p.loglik_batch(d,a,r,want_grad=grad); dt=time.perf_counter()-t0; st=ctx.stats()
    fl=M*6144.0**3*(1.0 if grad else 1/3)
    print("N=6144 wave 32, grad",grad,"ms/eval %.3f"%(dt*1e3/M),"factor TF %.2f"%(fl/(st["ms_factor"]*1e-3)/1e12),"frac %.3f"%(fl/(st["ms_factor"]*1e-3)/1e12/37.0), "info",out[-1].max())
PY
timeout 900 python scripts/cfg5_sweep.py > gpurun_out/cfg5_sweep_r2.log 2>&1; tail -3 gpurun_out/cfg5_sweep_r2.log | cut -c1-250
