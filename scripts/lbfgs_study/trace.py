import sys, numpy as np, time
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..')); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from lb import *
import oracle.simulate as S
from concurrent.futures import ProcessPoolExecutor
t,y,s,_=S.simulatethreelightcurves()
p=Problem(t,y,s,'matern32')
rhomin,rhomax=0.1,300.0
theta0,_=initial_solutions(p,1,1,5,rhomin,rhomax); theta0=np.asarray(theta0).reshape(-1,p.L+1)
grid=np.arange(0,20.0001,0.2)
rng=np.random.default_rng(1); idx=rng.choice(101*101,size=48,replace=False)
def work(m):
    delays=np.array([0.0,grid[m%101],grid[m//101]])
    def fg(th):
        try:
            ll,g=p.objective_grad_theta(th,delays,rhomin,rhomax)
            if not np.isfinite(ll): return False,np.inf,np.zeros_like(th)
            return True,-ll,-g
        except Exception: return False,np.inf,np.zeros_like(th)
    vals=[]
    for th in theta0:
        try: v=-p.objective_theta(th,delays,rhomin,rhomax)
        except Exception: v=np.inf
        vals.append(v)
    th0=theta0[int(np.argmin(vals))]
    ok,f0,g0=fg(th0); L=LB(len(th0)); L.start(th0,f0,g0); tr=[(0,f0,np.max(np.abs(g0)))]
    while L.status=='RUN':
        ok,ft,gt=fg(L.xt); it=L.iters; L.feed(ok,ft,gt)
        if L.iters>it or L.status!='RUN': tr.append((L.nfev,L.f,np.max(np.abs(L.g))))
    return m,tr,L.status
if __name__=='__main__':
    with ProcessPoolExecutor(8) as ex: res=list(ex.map(work,idx))
    n6=[];n8=[];nt=[]
    for m,tr,st in res:
        fin=tr[-1][1]; nf=[a for a,f,g in tr]; 
        k6=next(a for a,f,g in tr if f-fin<1e-7); k8=next(a for a,f,g in tr if f-fin<1e-9)
        n6.append(k6); n8.append(k8); nt.append(tr[-1][0])
    print('mean nfev total',np.mean(nt),'to within 1e-7',np.mean(n6),'to within 1e-9',np.mean(n8))
    for m,tr,st in res[:3]:
        print(m,st,[(a,f'{f-tr[-1][1]:.1e}',f'{g:.1e}') for a,f,g in tr])
