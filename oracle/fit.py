"""Hyper-parameter fit at fixed delays (oracle; test infrastructure only).

Follows /root/reference/src/gpccfixdelay_marginaliseb.jl:
  :62        rg = MersenneTwister(seed)                (here: numpy default_rng(seed); the Julia stream is not reproducible)
  :160-176   initial rho values (<=2 restarts: Uniform(rhomin+1e-3, rhomax-1e-3); >=3: log-spaced grid)
  :188       sample_alpha() = var.(y) .* U(0.8,1.2)
  :195-196   unconstrained start = [invmakepositive.(alpha); invtransformbetween(rho_i)]
  :203-215   screening of `initialrandom` starts, then optimize(..., NelderMead(), Options(iterations, g_tol=1e-6))
  :222-226   best over restarts
  :351       return (-minimum, predictTest, (alpha, postb, rho))
Optim.jl is not vendored; `nelder_mead` restates its published algorithm (Gao & Han adaptive parameters,
AffineSimplexer(a=0.025, b=0.5), stop when sqrt(var(f)*n/(n+1)) < g_tol).  `optimizer="lbfgs"` is the
variant north_star asks the GPU path to use (scipy L-BFGS-B on the analytic gradient).
"""
import numpy as np
from scipy.optimize import minimize

from .model import Problem, invmakepositive, invtransformbetween

PENALTY = 1e300    # safewrapper (:153): exceptions (non-PD matrix, rho<=0) -> a huge objective value (presumed)


def initial_solutions(problem, seed=1, numberofrestarts=1, initialrandom=5, rhomin=0.1, rhomax=20.0):
    """theta0[restart, draw, L+1] in the order the reference consumes its RNG (:166, :188, :207)."""
    rg = np.random.default_rng(seed)
    if numberofrestarts in (1, 2):
        rho0 = rg.uniform(rhomin + 1e-3, rhomax - 1e-3, numberofrestarts)            # :166
    else:
        rho0 = np.exp(np.linspace(np.log(rhomin + 1e-3), np.log(rhomax - 1e-3), numberofrestarts))  # :172
    var_y = np.array([a.var(ddof=1) for a in problem.y])
    out = np.empty((numberofrestarts, initialrandom, problem.L + 1))
    for i in range(numberofrestarts):
        for j in range(initialrandom):
            alpha0 = var_y * (rg.random(problem.L) * (1.2 - 0.8) + 0.8)              # :188
            out[i, j, : problem.L] = invmakepositive(alpha0)                           # :195
            out[i, j, problem.L] = invtransformbetween(rho0[i], rhomin, rhomax)       # :196
    return out, rho0


def nelder_mead(f, x0, iterations, g_tol=1e-6):
    """Optim.jl-style adaptive Nelder-Mead (minimisation).  Returns (x, f(x), iterations, f_calls)."""
    n = len(x0)
    alpha, beta, gamma, delta = 1.0, 1.0 + 2.0 / n, 0.75 - 1.0 / (2 * n), 1.0 - 1.0 / n
    simplex = [np.array(x0, dtype=np.float64)]
    for i in range(n):                                    # AffineSimplexer(a=0.025, b=0.5)
        v = np.array(x0, dtype=np.float64)
        v[i] = (1.0 + 0.5) * v[i] + 0.025
        simplex.append(v)
    fs = np.array([f(v) for v in simplex])
    calls = n + 1
    it = 0
    while it < iterations:
        order = np.argsort(fs, kind="stable")
        simplex = [simplex[i] for i in order]
        fs = fs[order]
        if np.sqrt(np.var(fs) * n / (n + 1)) < g_tol:
            break
        it += 1
        centroid = np.mean(simplex[:-1], axis=0)
        xr = centroid + alpha * (centroid - simplex[-1])
        fr = f(xr); calls += 1
        if fr < fs[0]:
            xe = centroid + beta * (xr - centroid)
            fe = f(xe); calls += 1
            if fe < fr:
                simplex[-1], fs[-1] = xe, fe
            else:
                simplex[-1], fs[-1] = xr, fr
        elif fr < fs[-2]:
            simplex[-1], fs[-1] = xr, fr
        else:
            if fr < fs[-1]:
                xc = centroid + gamma * (xr - centroid)          # outside contraction
                fc = f(xc); calls += 1
                ok = fc <= fr
            else:
                xc = centroid - gamma * (centroid - simplex[-1])  # inside contraction
                fc = f(xc); calls += 1
                ok = fc < fs[-1]
            if ok:
                simplex[-1], fs[-1] = xc, fc
            else:
                for i in range(1, n + 1):                         # shrink
                    simplex[i] = simplex[0] + delta * (simplex[i] - simplex[0])
                    fs[i] = f(simplex[i])
                calls += n
    j = int(np.argmin(fs))
    return simplex[j], float(fs[j]), it, calls


def gpcc(tarray, yarray, stdarray, *, kernel, delays, iterations, seed=1, numberofrestarts=1,
         initialrandom=5, rhomin=0.1, rhomax, optimizer="neldermead", theta0=None, return_info=False):
    """Restatement of `gpcc` (:46-53, :56-352).  `theta0` (restarts x draws x (L+1)) overrides the RNG."""
    p = Problem(tarray, yarray, stdarray, kernel)
    delays = np.asarray(delays, dtype=np.float64)
    assert len(delays) == p.L
    if theta0 is None:
        theta0, _ = initial_solutions(p, seed, numberofrestarts, initialrandom, rhomin, rhomax)
    theta0 = np.asarray(theta0, dtype=np.float64).reshape(-1, theta0.shape[-2], p.L + 1) \
        if np.ndim(theta0) == 3 else np.asarray(theta0, dtype=np.float64)[None]
    nfev = [0]

    def safenegativeobj(theta):                               # :149-153
        nfev[0] += 1
        try:
            v = -p.objective_theta(theta, delays, rhomin, rhomax)
            return v if np.isfinite(v) else PENALTY
        except Exception:
            return PENALTY

    def negobj_grad(theta):
        nfev[0] += 1
        try:
            ll, g = p.objective_grad_theta(theta, delays, rhomin, rhomax)
            if not np.isfinite(ll):
                return PENALTY, np.zeros_like(theta)
            return -ll, -g
        except Exception:
            return PENALTY, np.zeros_like(theta)

    best = None
    for i in range(theta0.shape[0]):                           # :222
        starts = theta0[i]
        vals = [safenegativeobj(th) for th in starts]          # :209
        th0 = starts[int(np.argmin(vals))]
        if optimizer == "neldermead":
            x, fx, _, _ = nelder_mead(safenegativeobj, th0, iterations, g_tol=1e-6)   # :205-211
        elif optimizer == "lbfgs":
            r = minimize(negobj_grad, th0, jac=True, method="L-BFGS-B",
                         options=dict(maxiter=iterations, ftol=1e-15, gtol=1e-8, maxcor=8))
            x, fx = r.x, float(r.fun)
        else:
            raise ValueError(optimizer)
        if best is None or fx < best[1]:                       # :224
            best = (x, fx)
    theta_opt, fmin = best
    alpha, rho = p.unpack(theta_opt, rhomin, rhomax)           # :235
    mupost, Spost = p.postb(delays, alpha, rho)                # :248-252

    def pred(ttest, ytest=None, stest=None):                   # :259-343
        if ytest is not None:
            return p.predict_loglik(delays, alpha, rho, ttest, ytest, stest)
        if len(ttest) > 0 and np.ndim(ttest[0]) > 0:
            return p.predict_full(delays, alpha, rho, ttest)
        return p.predict(delays, alpha, rho, ttest)

    out = (-fmin, pred, (alpha, (mupost, Spost), rho))         # :351
    if return_info:
        return out + (dict(nfev=nfev[0], theta=theta_opt),)
    return out


def cv_folds(nper, numberoffolds=5, seedcv=1):
    """Fold assignment per band, each band partitioned on its own with seed seedcv + b (UNUSED/performcv.jl:66).  The
    reference takes its partitions from MiscUtil.CVindices (Julia RNG, not reproducible here): a numpy permutation dealt
    round robin stands in for it, and callers can pass their own folds instead."""
    folds = []
    for b, n in enumerate(nper):
        perm = np.random.default_rng(seedcv + b + 1).permutation(n)
        f = np.empty(n, dtype=np.int64)
        f[perm] = np.arange(n) % numberoffolds
        folds.append(f)
    return folds


def performcv(tobs, yobs, sobs, *, delays, kernel, iterations=1, seedcv=1, numberofrestarts=1, initialrandom=1,
              numberoffolds=5, rhomin=0.1, rhomax=20.0, folds=None, optimizer="lbfgs", theta0=None):
    """K-fold cross-validation of the GPCC model at fixed delays (UNUSED/performcv.jl:41-139): fit on the training part,
    test log-likelihood pred(ttest, ytest, stest) on the held-out part; returns the vector of fold scores."""
    if folds is None:
        folds = cv_folds([len(t) for t in tobs], numberoffolds, seedcv)
    fitness = np.zeros(numberoffolds)
    for k in range(numberoffolds):
        tr = [np.asarray(f) != k for f in folds]
        sel = lambda arrs, mask_list: [np.asarray(a, dtype=np.float64)[m] for a, m in zip(arrs, mask_list)]
        te = [~m for m in tr]
        th = None if theta0 is None else theta0[k]
        r = gpcc(sel(tobs, tr), sel(yobs, tr), sel(sobs, tr), kernel=kernel, delays=delays, iterations=iterations, seed=seedcv,
                 numberofrestarts=numberofrestarts, initialrandom=initialrandom, rhomin=rhomin, rhomax=rhomax,
                 optimizer=optimizer, theta0=th)
        fitness[k] = r[1](sel(tobs, te), sel(yobs, te), sel(sobs, te))
    return fitness
