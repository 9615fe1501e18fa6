import numpy as np, sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import gpcc_b200
from conftest import load_golden
g = load_golden("fit_cfg1_cfg2")
p = gpcc_b200.Problem(g["tb"], g["yb"], g["sb"], "matern32")
print("variant", os.environ.get("GPCC_SMALL_VARIANT"))
ll, grad, info = p.loglik_batch([g["truedelays"]], [g["alpha"]], [float(g["rho"])], want_grad=True)
print(" loglik", ll, info)
try:
    mu, S = p.postb(g["truedelays"], g["alpha"], float(g["rho"]))
    print(" postb ok", np.allclose(mu, g["postb_mu"], rtol=1e-8))
except Exception as e:
    print(" postb failed:", e)
