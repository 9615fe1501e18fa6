"""simulatetwolightcurves / simulatethreelightcurves and larger synthetic problems of the same family.

Host-side data generators (the reference's are host-side too: /root/reference/src/simulatedata.jl:40-162;
exported at src/GPCC.jl:30).  Structure mirrored: N=[60,50,40], times U(0,20) with a gap 8..12 in band 2,
latent OU process rho=3.5 under delays [0,2,4], alpha=[1,1.5,2], b=[6,15,25], noise sigma=0.75, eigenvalue
clamp 1e-6.  Julia's MersenneTwister stream is not reproducible outside Julia: numpy default_rng(seed) is
used, so the data follow the reference's distribution, not its bits.  The plotting side effects
(simulatedata.jl:45-58) are not reproduced.
"""
import numpy as np

RHO_TRUE = 3.5
TRUEDELAYS = np.array([0.0, 2.0, 4.0])
ALPHA_TRUE = np.array([1.0, 1.5, 2.0])
B_TRUE = np.array([6.0, 15.0, 25.0])


def _ou_delayed_cov(alpha, delays, rho, t):
    ts = np.concatenate([np.asarray(a) - d for a, d in zip(t, delays)])
    sc = np.concatenate([np.full(len(a), s) for a, s in zip(t, alpha)])
    return (sc[:, None] * sc[None, :]) * np.exp(-np.abs(ts[:, None] - ts[None, :]) / rho)   # util.jl:15-23, delayedCovariance.jl:27


def simulatedata(sigma=0.75, seed=1):
    rg = np.random.default_rng(seed)
    N = [60, 50, 40]                                                           # simulatedata.jl:119
    t = [rg.random(N[0]) * 20.0,
         np.concatenate([rg.random(25) * 8.0, 12.0 + rg.random(25) * 8.0]),
         rg.random(N[2]) * 20.0]                                                # :121
    C = _ou_delayed_cov(ALPHA_TRUE, TRUEDELAYS, RHO_TRUE, t)                    # :128
    U, S, _ = np.linalg.svd(C)                                                  # :132-136
    C = (U * np.maximum(1e-6, np.abs(S))) @ U.T
    C = 0.5 * (C + C.T)
    w, V = np.linalg.eigh(C)
    Y = V @ (np.sqrt(np.maximum(w, 0.0)) * rg.standard_normal(sum(N)))          # :145
    y, mark = [], 0
    for i in range(3):                                                          # :151-157
        y.append(Y[mark:mark + N[i]] * ALPHA_TRUE[i] + B_TRUE[i] + sigma * rg.standard_normal(N[i]))
        mark += N[i]
    return t, y, [sigma * np.ones(n) for n in N], TRUEDELAYS.copy(), ALPHA_TRUE.copy(), B_TRUE.copy()


def simulatetwolightcurves(sigma=0.75, seed=1):
    t, y, s, d, _, _ = simulatedata(sigma, seed)
    return t[:2], y[:2], s[:2], d[:2]                                           # :61


def simulatethreelightcurves(sigma=0.75, seed=1):
    t, y, s, d, _, _ = simulatedata(sigma, seed)
    return t, y, s, d                                                           # :91


def synthetic_bands(n_per_band, seed=1, sigma=0.75, span=None):
    """Large problems for BASELINE configs 4/5: times U(0, span) (default 0.4*n per band, "constant density"); the
    latent OU process is drawn exactly in O(N log N) through its Markov property on the sorted shifted times."""
    rg = np.random.default_rng(seed)
    L = len(n_per_band)
    delays = TRUEDELAYS[:L] if L <= 3 else 2.0 * np.arange(L)
    alpha = ALPHA_TRUE[:L] if L <= 3 else 1.0 + 0.5 * np.arange(L)
    b = B_TRUE[:L] if L <= 3 else 6.0 + 9.0 * np.arange(L)
    t = [rg.random(n) * (span if span is not None else 0.4 * n) for n in n_per_band]
    ts = np.concatenate([tl - d for tl, d in zip(t, delays)])
    order = np.argsort(ts)
    f = np.empty(len(ts))
    eps = rg.standard_normal(len(ts))
    prev_t, prev_f = None, 0.0
    for k, idx in enumerate(order):
        if prev_t is None:
            f[idx] = eps[k]
        else:
            c = np.exp(-(ts[idx] - prev_t) / RHO_TRUE)
            f[idx] = c * prev_f + np.sqrt(max(1.0 - c * c, 0.0)) * eps[k]
        prev_t, prev_f = ts[idx], f[idx]
    y, mark = [], 0
    for l, n in enumerate(n_per_band):
        y.append(alpha[l] * alpha[l] * f[mark:mark + n] + b[l] + sigma * rg.standard_normal(n))
        mark += n
    return t, y, [sigma * np.ones(n) for n in n_per_band], np.array(delays, dtype=np.float64)
