import sys, numpy as np
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..')); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from lb import *
import oracle.simulate as S
from oracle.model import makepositive, transformbetween
from concurrent.futures import ProcessPoolExecutor
z=np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'profiles', 'cfg3_nfev_r2.npz')); print(z.files)
nf=z[z.files[0]] if 'nfev' not in z.files else z['nfev']
nf=np.asarray(nf).ravel()
top=np.argsort(-nf)[:8]; print('longest', top, nf[top])
t,y,s,_=S.simulatethreelightcurves(); p=Problem(t,y,s,'matern32'); rhomin,rhomax=0.1,300.0
theta0,_=initial_solutions(p,1,1,5,rhomin,rhomax); theta0=np.asarray(theta0).reshape(-1,p.L+1)
grid=np.arange(0,20.0001,0.2)
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
    ok,f0,g0=fg(th0); L=LB(len(th0)); L.start(th0,f0,g0)
    while L.status=='RUN':
        ok,ft,gt=fg(L.xt); L.feed(ok,ft,gt)
    return m, delays, L.nfev, L.x, L.g
if __name__=='__main__':
    with ProcessPoolExecutor(8) as ex: res=list(ex.map(work,top))
    for m,dl,n,x,g in res:
        alpha=np.log1p(np.exp(x[:3]))+1e-8; rho=rhomin+(rhomax-rhomin)/(1+np.exp(-x[3]))
        print(m, dl, 'nfev',n,'theta',np.round(x,2),'alpha',np.round(alpha,4),'rho',round(float(rho),4),'|g|',np.abs(g).max())
