"""Randomised A/B of the last-band cache of the tiled path: for a few three- and four-band problems (band boundaries on and off the tile
boundaries) and candidate sets (full grids, random subsets, repeated candidates) the log-likelihoods with the cache must be
BITWISE those of a process started with GPCC_LARGE_NO_TAUCACHE=1.  Usage: check_last_band_cache.py   (parent)"""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [([128, 128, 128], "matern32"), ([130, 200, 140], "OU"), ([400, 100, 300], "rbf"), ([1000, 700, 900], "matern52"), ([256, 512, 384], "matern32"), ([300, 128, 200, 260], "matern52")]

def candidates(rg, case, L):
    c2, c3 = np.sort(rg.uniform(0, 8, rg.integers(2, 7))), np.sort(rg.uniform(0, 8, rg.integers(2, 9)))
    grid = np.array([[0.0] + [2.1] * (L - 3) + [a, b] for b in c3 for a in c2])
    kind = case % 3
    if kind == 1: grid = grid[rg.permutation(len(grid))[: max(8, 2 * len(grid) // 3)]]
    if kind == 2: grid = np.concatenate([grid, grid[rg.integers(0, len(grid), 5)]])
    return grid

def run(tag):
    import gpcc_b200
    out = {}
    rg = np.random.default_rng(7)
    ctx = gpcc_b200.default_context()
    for ci, (nper, kernel) in enumerate(CASES):
        t, y, s, _ = gpcc_b200.synthetic_bands(nper, seed=30 + ci)
        p = gpcc_b200.Problem(t, y, s, kernel, ctx)
        for rep in range(3):
            d = candidates(rg, ci + rep, len(nper))
            M = len(d)
            alpha, rho = np.tile(rg.uniform(0.5, 2.5, len(nper)), (M, 1)), np.full(M, rg.uniform(1.0, 6.0))
            ll, info = p.loglik_batch(d, alpha, rho)
            st = ctx.stats()
            out["ll_%d_%d" % (ci, rep)] = ll
            out["st_%d_%d" % (ci, rep)] = np.array([st["n_tau_cache"], st["n_shared_prefix"], M, int(np.sum(info != 0))])
        p.close()
    np.savez(tag, **out)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1])
        sys.exit(0)
    with tempfile.TemporaryDirectory() as tmp:
        res = {}
        for name, env in (("cache", {}), ("plain", {"GPCC_LARGE_NO_TAUCACHE": "1"})):
            r = subprocess.run([sys.executable, os.path.abspath(__file__), os.path.join(tmp, name + ".npz")], env=dict(os.environ, **env), capture_output=True, text=True)
            if r.returncode: print(r.stderr[-2000:]); sys.exit(1)
            res[name] = np.load(os.path.join(tmp, name + ".npz"))
        bad = 0
        for k in res["cache"].files:
            if not k.startswith("ll_"): continue
            a, b, st, stp = res["cache"][k], res["plain"][k], res["cache"]["st_" + k[3:]], res["plain"]["st_" + k[3:]]
            same = np.array_equal(a, b)
            bad += (not same) or stp[0] != 0
            print("%s: %3d candidates, cache served %3d, shared leading block %3d, info!=0 %d | bitwise equal to the run without cache: %s (max rel diff %.1e)" % (
                k, st[2], st[0], st[1], st[3], same, float(np.max(np.abs(a - b) / np.abs(b)))))
        print("ALL EQUAL" if not bad else "MISMATCH in %d sets" % bad)
        sys.exit(1 if bad else 0)
