#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 -k "large or cfg4 or structure or exact or fit_state or prior_with" 2>&1 | tail -15
python - <<'PY' 2>&1 | tail -8
import sys, time, numpy as np
sys.path.insert(0,'.')
import gpcc_b200
ctx=gpcc_b200.Context(1,profiling=True)
t,y,s,_=gpcc_b200.synthetic_bands([2048]*3,seed=4)
p=gpcc_b200.Problem(t,y,s,"matern52",ctx)
c4=np.arange(0.0,19.8001,0.2); grid=np.array([[0.0,a,b] for b in c4 for a in c4])
th=np.concatenate([np.log(np.expm1(np.array([1.0,2.2,4.0]))),[np.log((3.5-0.1)/(300.0-3.5))]])[None]
p.grid_posterior(grid[:32],th,iterations=0,rhomin=0.1,rhomax=300.0)
for M in (400, 1200):
    sub=grid[:M] if M<1000 else np.array([[0.0,a,b] for b in c4[:12] for a in c4])
    t0=time.perf_counter(); r=p.grid_posterior(sub,th,iterations=0,rhomin=0.1,rhomax=300.0); dt=time.perf_counter()-t0; st=ctx.stats()
    print("M=%d: %.3f ms per candidate, shared %d, factor ms %.1f, assembly ms %.1f, posterior sum %.6f"%(len(sub),dt*1e3/len(sub),st["n_shared_prefix"],st["ms_factor"],st["ms_assembly"],r["posterior"].sum()))
PY
(cd scripts/microbench && timeout 200 ./potrf_yardstick) | tee gpurun_out/potrf_yardstick_r2.log
