import sys, numpy as np, time
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..')); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from lb import *
import oracle.simulate as S
from concurrent.futures import ProcessPoolExecutor
t,y,s,_=S.simulatethreelightcurves()
p=Problem(t,y,s,'matern32')
rhomin,rhomax=0.1,300.0
theta0,_=initial_solutions(p,1,1,5,rhomin,rhomax)
theta0=np.asarray(theta0).reshape(-1,p.L+1)
g=np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'tests', 'golden', 'fit_cfg3_full.npz'))
print(list(g.keys()))
rng=np.random.default_rng(0)
grid=np.arange(0,20.0001,0.2)
idx=np.concatenate([rng.choice(101*101,size=int(sys.argv[1]) if len(sys.argv)>1 else 120,replace=False),[10039,9629,10038,10143,9836,9837,10041,9939, 2041, 2040, 2142, 1940]])
VARS={'base':None,'relcap.5':{'relcap':0.5},'relcap.25':{'relcap':0.25}}
def work(m):
    d2=grid[m%101]; d3=grid[m//101]
    delays=np.array([0.0,d2,d3])
    out={}
    for k,v in VARS.items():
        out[k]=fit(p,delays,theta0,rhomin,rhomax,v)
    return m,out
if __name__=='__main__':
    t0=time.time()
    with ProcessPoolExecutor(16) as ex: res=list(ex.map(work,idx))
    print('time',time.time()-t0)
    for k in VARS:
        nf=np.array([r[1][k][2] for r in res]); it=np.array([r[1][k][1] for r in res]); f=np.array([r[1][k][0] for r in res]); fb=np.array([r[1]['base'][0] for r in res])
        d=f-fb; print('max gap',d.max(),'evals saved on the longest 10%%:', 1-nf[np.argsort(-np.array([q[1]['base'][2] for q in res]))[:max(1,len(res)//10)]].sum()/np.sort(np.array([q[1]['base'][2] for q in res]))[::-1][:max(1,len(res)//10)].sum()); print(f"{k:12s} mean nfev {nf.mean():6.2f} max {nf.max():4d}  iters {it.mean():6.2f}  nfev/iter {nf.sum()/it.sum():.2f}  worse>1e-6 than base: {(f>fb+1e-6).sum()}  better>1e-6: {(f<fb-1e-6).sum()}  max|df| among same-basin {np.max(np.abs(f-fb)[np.abs(f-fb)<1e-3]):.1e}")
