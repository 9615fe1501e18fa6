# Python port of gpcc_b200/csrc/lbfgs.h for evaluation-count studies on the oracle objective (CPU).  variant=None is the state machine as it
# was before the relative step cap; variant={"relcap": 0.25} is the rule lbfgs.h ships now (n_scale = L, rel_cap = 0.25).
import numpy as np, sys, time
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from oracle.model import Problem
from oracle.simulate import simulatethreelightcurves
from oracle.fit import initial_solutions

STEP_CAP = 2.0
class LB:
    def __init__(s, n, hist=8, gtol=1e-7, ftol=1e-13, max_iter=1000, max_ls=30, variant=None):
        s.n=n; s.hist=hist; s.gtol=gtol; s.ftol=ftol; s.max_iter=max_iter; s.max_ls=max_ls; s.v=variant or {}
    def start(s, x0, f0, g0):
        s.dfh=[]; s.H=None; s.x=x0.copy(); s.g=g0.copy(); s.f=f0; s.status='RUN'; s.iters=0; s.S=[]; s.Y=[]; s.R=[]; s.small_df=0; s.nfev=0
        if not (np.max(np.abs(s.g)) > s.gtol): s.status='CONV'; return
        s.new_direction(True)
    def new_direction(s, first):
        if s.v.get('bfgs') and getattr(s,'H',None) is not None:
            s.d=-s.H.dot(s.g); s.gd0=s.g.dot(s.d)
            if not (s.gd0<0) or not np.isfinite(s.gd0):
                s.H=None; s.S=[];s.Y=[];s.R=[]; s.d=-s.g.copy(); s.gd0=s.g.dot(s.d); first=True
            s.t=1.0
            if first:
                gn=np.sqrt(s.g.dot(s.g)); s.t=min(1.0,1.0/gn)
            dmax=np.max(np.abs(s.d)); s.tcap=STEP_CAP/dmax if dmax>0 else np.inf
            s.t=min(s.t,s.tcap); s.tlo=0.0; s.thi=np.inf; s.ls=0; s.fb=None
            s.xt=s.x+s.t*s.d
            return
        q=s.g.copy(); a=[]
        for S,Y,R in zip(reversed(s.S),reversed(s.Y),reversed(s.R)):
            ak=R*S.dot(q); a.append(ak); q-=ak*Y
        if s.S:
            q*= 1.0/(s.R[-1]*s.Y[-1].dot(s.Y[-1]))
        for (S,Y,R),ak in zip(zip(s.S,s.Y,s.R), reversed(a)):
            bb=R*Y.dot(q); q+=(ak-bb)*S
        s.d=-q; s.gd0=s.g.dot(s.d)
        if not (s.gd0<0) or not np.isfinite(s.gd0):
            s.S=[];s.Y=[];s.R=[]; s.d=-s.g.copy(); s.gd0=s.g.dot(s.d); first=True
        s.t=1.0
        if first or not s.S:
            gn=np.sqrt(s.g.dot(s.g)); s.t=min(1.0,1.0/gn)
            if 'first_t' in s.v: s.t=min(1.0, s.v['first_t']/gn)
        dmax=np.max(np.abs(s.d)); s.tcap=STEP_CAP/dmax if dmax>0 else np.inf
        if 'relcap' in s.v:
            cap=np.full(s.n,STEP_CAP); cap[:s.n-1]=np.maximum(STEP_CAP, s.v['relcap']*np.maximum(s.x[:s.n-1],0.0))
            with np.errstate(divide='ignore'): s.tcap=np.min(np.where(np.abs(s.d)>0, cap/np.abs(s.d), np.inf))
        s.t=min(s.t,s.tcap); s.tlo=0.0; s.thi=np.inf; s.ls=0; s.fb=None
        s.xt=s.x+s.t*s.d
    def accept(s, xn, fn, gn):
        sv=xn-s.x; y=gn-s.g; sy=sv.dot(y)
        if sy>1e-10*np.sqrt(sv.dot(sv)*y.dot(y)) and sy>0:
            if s.v.get('bfgs'):
                n=s.n; I=np.eye(n)
                if getattr(s,'H',None) is None: s.H=(sy/y.dot(y))*I
                rho=1.0/sy
                s.H=(I-rho*np.outer(sv,y)).dot(s.H).dot(I-rho*np.outer(y,sv))+rho*np.outer(sv,sv)
            s.S.append(sv); s.Y.append(y); s.R.append(1.0/sy)
            if len(s.S)>s.hist: s.S.pop(0); s.Y.pop(0); s.R.pop(0)
        df=s.f-fn; s.x=xn.copy(); s.g=gn.copy(); s.f=fn; s.iters+=1
        if 'geo' in s.v:
            q,tol=s.v['geo']
            h=getattr(s,'dfh',[]); h.append(df); s.dfh=h[-3:]
            if len(s.dfh)==3 and s.dfh[2]>=0 and s.dfh[2]<=q*s.dfh[1] and s.dfh[1]<=q*s.dfh[0] and q/(1-q)*s.dfh[2]<=tol and np.max(np.abs(s.g))<=s.v.get('geo_g',1e9):
                s.status='CONV'; return
        if np.max(np.abs(s.g))<=s.gtol: s.status='CONV'; return
        if df<=s.ftol*max(1.0,abs(s.f)):
            s.small_df+=1
            if s.small_df>=2: s.status='CONV'; return
        else: s.small_df=0
        if s.iters>=s.max_iter: s.status='CAP'; return
        s.new_direction(False)
    def feed(s, ok, ft, gt):
        s.nfev+=1; s.ls+=1
        c1=1e-4; c2=s.v.get('c2',0.9)
        armijo = ok and np.isfinite(ft) and ft<=s.f+c1*s.t*s.gd0
        if armijo:
            gtd=gt.dot(s.d)
            if gtd>=c2*s.gd0: s.accept(s.xt,ft,gt); return
            if s.thi==np.inf and s.t>=s.tcap: s.accept(s.xt,ft,gt); return
            if s.v.get('armijo_only') and s.ls==1 and len(s.S)>0 and s.t==1.0: s.accept(s.xt,ft,gt); return
            if s.fb is None or ft<s.fb[1]: s.fb=(s.xt.copy(),ft,gt.copy())
            s.tlo=s.t
            if s.thi==np.inf:
                tn=min(2.0*s.t,s.tcap)
                if s.v.get('expand'):
                    # extrapolate with the secant of the directional derivative: g'd(t) linear between 0 and t
                    den=gtd-s.gd0
                    if den>0:
                        tq=-s.gd0*s.t/den
                        tn=min(max(tq,1.5*s.t),s.v['expand']*s.t,s.tcap)
                    else: tn=min(s.v['expand']*s.t,s.tcap)
                s.t=tn
            else: s.t=0.5*(s.tlo+s.thi)
        else:
            noise=4e-13*max(1.0,abs(s.f))
            if ok and np.isfinite(ft) and abs(ft-s.f)<=noise and -s.t*s.gd0<=noise:
                if s.fb is not None and s.fb[1]<s.f:
                    s.accept(*s.fb)
                    if s.status=='RUN': s.status='CONV'
                    return
                s.status='CONV'; return
            s.thi=s.t; tn=0.5*(s.tlo+s.thi)
            if s.tlo==0.0 and ok and np.isfinite(ft):
                den=2.0*(ft-s.f-s.gd0*s.t)
                if den>0:
                    tq=-s.gd0*s.t*s.t/den; tn=min(max(tq,0.1*s.t),0.5*s.t)
            s.t=tn
        if s.ls>=s.max_ls or not (s.thi-s.tlo>1e-16*max(1.0,s.thi)):
            if s.fb is not None and s.fb[1]<s.f: s.accept(*s.fb); return
            s.status='STALL'; return
        s.xt=s.x+s.t*s.d

def fit(p, delays, theta0, rhomin, rhomax, variant=None):
    def fg(th):
        try:
            ll,g=p.objective_grad_theta(th,delays,rhomin,rhomax)
            if not np.isfinite(ll): return False, np.inf, np.zeros_like(th)
            return True,-ll,-g
        except Exception:
            return False, np.inf, np.zeros_like(th)
    vals=[]
    for th in theta0:
        try: v=-p.objective_theta(th,delays,rhomin,rhomax)
        except Exception: v=np.inf
        vals.append(v if np.isfinite(v) else np.inf)
    th0=theta0[int(np.argmin(vals))]
    ok,f0,g0=fg(th0)
    L=LB(len(th0),hist=(variant or {}).get("hist",8),variant=variant); L.start(th0,f0,g0)
    trace=[]
    while L.status=='RUN':
        ok,ft,gt=fg(L.xt); L.feed(ok,ft,gt)
    return L.f, L.iters, L.nfev, L.status

if __name__=='__main__':
    t,y,s,_=simulatethreelightcurves()[:4] if len(simulatethreelightcurves())>3 else (*simulatethreelightcurves(),None)
