"""Fixed-delay GP model with the shift b marginalised (oracle; test infrastructure only).

Follows /root/reference/src/gpccfixdelay_marginaliseb.jl:
  data prep        :85-98     Y, Sobs, mu_b = mean(y_l), Sigma_b = 100*var(y_l) (unbiased), B = Q Sigma_b Q', bbar
  transforms       :112-126   alpha = makepositive(x)+1e-8, rho = transformbetween(x, rhomin, rhomax)
  objective        :133-141   logpdf(MvNormal(bbar, K+Sobs+B), Y)
  postb            :248-252
  predictTest x3   :259-343
``makepositive`` / ``transformbetween`` live in the un-vendored MiscUtil package; softplus and the
scaled logistic are the presumed definitions.
The analytic gradient is an addition of north_star (the reference is derivative-free).
"""
import numpy as np
import scipy.linalg as sla

from .covariance import delayed_covariance

LOG2PI = float(np.log(2.0 * np.pi))
ALPHA_FLOOR = 1e-8      # gpccfixdelay_marginaliseb.jl:112
JITTER = 1e-8           # :69, used at :279
SIGMA_FLOOR = 1e-6      # :303


# ---- MiscUtil transforms (presumed) -------------------------------------------------------------
def makepositive(x):
    x = np.asarray(x, dtype=np.float64)
    return np.logaddexp(0.0, x)                    # log(1+exp(x)), overflow-safe


def invmakepositive(y):
    y = np.asarray(y, dtype=np.float64)
    return y + np.log(-np.expm1(-y))               # log(exp(y)-1)


def logistic(x):
    x = np.asarray(x, dtype=np.float64)
    return 0.5 * (1.0 + np.tanh(0.5 * x))


def transformbetween(x, lo, hi):
    return lo + (hi - lo) * logistic(x)


def invtransformbetween(y, lo, hi):
    u = (np.asarray(y, dtype=np.float64) - lo) / (hi - lo)
    return np.log(u) - np.log1p(-u)


class Problem:
    """Data of one ``gpcc`` call: L bands of (t, y, sigma)."""

    def __init__(self, tarray, yarray, stdarray, kernel):
        self.kernel = kernel
        self.t = [np.asarray(a, dtype=np.float64) for a in tarray]
        self.y = [np.asarray(a, dtype=np.float64) for a in yarray]
        self.s = [np.asarray(a, dtype=np.float64) for a in stdarray]
        self.L = len(self.t)
        assert self.L == len(self.y) == len(self.s)                      # :78
        self.n = np.array([len(a) for a in self.t])
        self.N = int(self.n.sum())
        self.band = np.repeat(np.arange(self.L), self.n)
        self.Y = np.concatenate(self.y)                                   # :85
        self.sobs = np.concatenate(self.s) ** 2                           # :89
        self.mub = np.array([a.mean() for a in self.y])                   # :92
        self.Sigmab = 100.0 * np.array([a.var(ddof=1) for a in self.y])   # :94 (diagonal)
        self.bbar = self.mub[self.band]                                   # :98
        self.B = np.where(self.band[:, None] == self.band[None, :], self.Sigmab[self.band][:, None], 0.0)  # :96

    # ---- parameters ---------------------------------------------------------------------------
    def unpack(self, theta, rhomin, rhomax):                              # :116-126
        theta = np.asarray(theta, dtype=np.float64)
        assert len(theta) == self.L + 1
        return makepositive(theta[: self.L]) + ALPHA_FLOOR, float(transformbetween(theta[self.L], rhomin, rhomax))

    # ---- covariance ---------------------------------------------------------------------------
    def Ktilde(self, delays, alpha, rho):                                 # :135
        K = delayed_covariance(self.kernel, alpha, delays, rho, self.t)
        K = K + np.diag(self.sobs) + self.B
        return K

    # ---- objective ----------------------------------------------------------------------------
    def loglik(self, delays, alpha, rho):                                 # :133-141
        K = self.Ktilde(delays, alpha, rho)
        c, low = sla.cho_factor(K, lower=True, check_finite=False)        # PDMat -> dpotrf (:139)
        z = sla.solve_triangular(c, self.Y - self.bbar, lower=True, check_finite=False)
        logdet = 2.0 * np.sum(np.log(np.diag(c)))
        return -0.5 * (self.N * LOG2PI + logdet + z @ z)

    def loglik_grad(self, delays, alpha, rho):
        """logL and d logL / d(alpha_1..alpha_L, rho)   (SURVEY.md section 8 row a5)."""
        alpha = np.asarray(alpha, dtype=np.float64)
        K0 = delayed_covariance(self.kernel, alpha, delays, rho, self.t)
        dK = delayed_covariance(self.kernel, alpha, delays, rho, self.t, drho=True)
        K = K0 + np.diag(self.sobs) + self.B
        c, low = sla.cho_factor(K, lower=True, check_finite=False)
        r = self.Y - self.bbar
        z = sla.solve_triangular(c, r, lower=True, check_finite=False)
        logdet = 2.0 * np.sum(np.log(np.diag(c)))
        ll = -0.5 * (self.N * LOG2PI + logdet + z @ z)
        Kinv = sla.cho_solve((c, True), np.eye(self.N), check_finite=False)
        a = Kinv @ r
        W = np.outer(a, a) - Kinv
        rows = (W * K0).sum(axis=1)                     # s_i = sum_j W_ij K0_ij
        g = np.empty(self.L + 1)
        for p in range(self.L):
            g[p] = rows[self.band == p].sum() / alpha[p]
        g[self.L] = 0.5 * np.sum(W * dK)
        return ll, g

    def objective_theta(self, theta, delays, rhomin, rhomax):
        alpha, rho = self.unpack(theta, rhomin, rhomax)
        return self.loglik(delays, alpha, rho)

    def objective_grad_theta(self, theta, delays, rhomin, rhomax):
        """logL and gradient w.r.t. the unconstrained parameters (chain rule through the transforms)."""
        theta = np.asarray(theta, dtype=np.float64)
        alpha, rho = self.unpack(theta, rhomin, rhomax)
        ll, g = self.loglik_grad(delays, alpha, rho)
        s = logistic(theta)
        jac = np.empty(self.L + 1)
        jac[: self.L] = s[: self.L]                               # d softplus = logistic
        jac[self.L] = (rhomax - rhomin) * s[self.L] * (1.0 - s[self.L])
        return ll, g * jac

    # ---- posterior of the shifts (:248-252) ------------------------------------------------------
    def postb(self, delays, alpha, rho):
        K = delayed_covariance(self.kernel, alpha, delays, rho, self.t)
        KS = K + np.diag(self.sobs)                                       # Sobs + K, *without* B
        Q = (self.band[:, None] == np.arange(self.L)[None, :]).astype(np.float64)   # util.jl:56-70
        Sb_inv = np.diag(1.0 / self.Sigmab)
        Spost = np.linalg.solve(Sb_inv + Q.T @ np.linalg.solve(KS, Q), np.eye(self.L))        # :248
        mupost = Spost @ (np.linalg.solve(KS, Q).T @ self.Y + self.mub / self.Sigmab)        # :250
        Spost = 0.5 * (Spost + Spost.T)                                                       # :252
        return mupost, Spost

    # ---- predictions (:259-343) --------------------------------------------------------------------
    def predict_full(self, delays, alpha, rho, ttest):
        """pred(ttest::Vector{Vector}) -> (mu, Sigma) (:259-289)."""
        ttest = [np.asarray(a, dtype=np.float64) for a in ttest]
        nt = np.array([len(a) for a in ttest])
        bt = np.repeat(np.arange(self.L), nt)
        KSB = self.Ktilde(delays, alpha, rho)
        Bs = np.where(self.band[:, None] == bt[None, :], self.Sigmab[self.band][:, None], 0.0)   # :264
        Bss = np.where(bt[:, None] == bt[None, :], self.Sigmab[bt][:, None], 0.0)                # :266
        kBs = delayed_covariance(self.kernel, alpha, delays, rho, self.t, ttest) + Bs            # :269
        cB = delayed_covariance(self.kernel, alpha, delays, rho, ttest) + Bss                    # :272
        Spred = cB - kBs.T @ np.linalg.solve(KSB, kBs)                                            # :275
        Spred = 0.5 * (Spred + Spred.T)                                                           # :277
        Spred = Spred + JITTER * np.eye(len(bt))                                                  # :279
        mupred = kBs.T @ np.linalg.solve(KSB, self.Y - self.bbar) + self.mub[bt]                  # :283-285
        return mupred, Spred

    def predict(self, delays, alpha, rho, ttest):
        """pred(ttest::Vector) -> per-band means and standard deviations (:293-307)."""
        ttest = np.asarray(ttest, dtype=np.float64)
        nt = len(ttest)
        mu, S = self.predict_full(delays, alpha, rho, [ttest] * self.L)
        sd = np.sqrt(np.maximum(np.diag(S), SIGMA_FLOOR))
        return [mu[i * nt:(i + 1) * nt] for i in range(self.L)], [sd[i * nt:(i + 1) * nt] for i in range(self.L)]

    def predict_loglik(self, delays, alpha, rho, ttest, ytest, stest):
        """pred(ttest, ytest, sigmatest) -> test log-likelihood (:311-343)."""
        mu, S = self.predict_full(delays, alpha, rho, ttest)
        S = S + np.diag(np.concatenate([np.asarray(a, dtype=np.float64) for a in stest]) ** 2)
        S = 0.5 * (S + S.T)
        yt = np.concatenate([np.asarray(a, dtype=np.float64) for a in ytest])
        try:
            c = sla.cholesky(S, lower=True, check_finite=False)
        except np.linalg.LinAlgError:
            # nearestposdef(S; minimumeigenvalue=1e-6): eigen-clamp (UNUSED/gpcc.jl:294-300)
            w, V = np.linalg.eigh(S)
            S = (V * np.maximum(w, 1e-6)) @ V.T
            S = 0.5 * (S + S.T)
            c = sla.cholesky(S, lower=True, check_finite=False)
        z = sla.solve_triangular(c, yt - mu, lower=True, check_finite=False)
        return -0.5 * (len(yt) * LOG2PI + 2.0 * np.sum(np.log(np.diag(c))) + z @ z)
