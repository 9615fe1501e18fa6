"""50-digit mpmath evaluation of postb and predictTest (oracle; test infrastructure only).

Follows /root/reference/src/gpccfixdelay_marginaliseb.jl literally, in exact-enough arithmetic:
  :248-250   Sigma_postb = (Sigma_b \\ I + Q'((Sobs+K) \\ Q)) \\ I ;  mu_postb = Sigma_postb ((Q'/(Sobs+K)) Y + Sigma_b \\ mu_b)
  :262-285   kB* = delayedCovariance(t, t*) + Q Sigma_b Q*' ; cB = delayedCovariance(t*) + Q* Sigma_b Q*' ;
             Sigma_pred = cB - kB*'(KSobsB \\ kB*) (+ symmetrise, + 1e-8 I) ; mu_pred = kB*'(KSobsB \\ (Y - bbar)) + Q* mu_b
  :303       sigma = sqrt(max(diag(Sigma_pred), 1e-6))
The float64 oracle (oracle/model.py, LU solves like the reference's `\\`) and the device library are BOTH compared with these
values: sigma_pred cancels ~4 digits ((alpha^2 + Sigma_b) - k*'K^-1 k*, Sigma_b = 100 var(y)), so a float64 evaluation of
the reference's formula is only good to ~1e-8..1e-6 relative and cannot arbitrate between two float64 implementations.
The covariance matrices are built in float64 (the inputs every implementation shares) and converted exactly.
"""
import numpy as np

from .covariance import delayed_covariance
from .model import JITTER, SIGMA_FLOOR, Problem


def _mp():
    import mpmath as mp
    mp.mp.dps = 50
    return mp


def postb_exact(p: Problem, delays, alpha, rho):
    """(mu_postb, Sigma_postb) as float64 arrays rounded from the 50-digit result."""
    mp = _mp()
    K = delayed_covariance(p.kernel, alpha, delays, rho, p.t) + np.diag(p.sobs)          # Sobs + K, without B (:248)
    Q = (p.band[:, None] == np.arange(p.L)[None, :]).astype(np.float64)
    KSi = mp.inverse(mp.matrix(K.tolist()))
    Qm, Y = mp.matrix(Q.tolist()), mp.matrix(p.Y.tolist())
    Sbi = mp.diag([1 / mp.mpf(float(v)) for v in p.Sigmab])
    Spost = mp.inverse(Sbi + Qm.T * KSi * Qm)
    mupost = Spost * (Qm.T * (KSi * Y) + Sbi * mp.matrix(p.mub.tolist()))
    S = np.array([[float(Spost[i, j]) for j in range(p.L)] for i in range(p.L)])
    return np.array([float(mupost[i]) for i in range(p.L)]), 0.5 * (S + S.T)


def predict_exact(p: Problem, delays, alpha, rho, ttest):
    """pred(ttest::Vector{Vector}) -> (mu, Sigma incl. jitter, sd) as float64 arrays rounded from the 50-digit result."""
    mp = _mp()
    ttest = [np.asarray(a, dtype=np.float64) for a in ttest]
    nt = np.array([len(a) for a in ttest])
    bt = np.repeat(np.arange(p.L), nt)
    NT = int(nt.sum())
    KSB = mp.matrix(p.Ktilde(delays, alpha, rho).tolist())
    Bs = np.where(p.band[:, None] == bt[None, :], p.Sigmab[p.band][:, None], 0.0)
    Bss = np.where(bt[:, None] == bt[None, :], p.Sigmab[bt][:, None], 0.0)
    kBs = mp.matrix((delayed_covariance(p.kernel, alpha, delays, rho, p.t, ttest) + Bs).tolist())
    cB = mp.matrix((delayed_covariance(p.kernel, alpha, delays, rho, ttest) + Bss).tolist())
    Ki = mp.inverse(KSB)
    W = Ki * kBs
    Sp = cB - kBs.T * W
    r = mp.matrix((p.Y - p.bbar).tolist())
    mu = kBs.T * (Ki * r)
    mu_f = np.array([float(mu[i]) for i in range(NT)]) + p.mub[bt]
    S = np.array([[float(Sp[i, j]) for j in range(NT)] for i in range(NT)])
    S = 0.5 * (S + S.T) + JITTER * np.eye(NT)
    sd = np.sqrt(np.maximum(np.diag(S), SIGMA_FLOOR))
    return mu_f, S, sd
