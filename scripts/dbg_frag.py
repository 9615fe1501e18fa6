import numpy as np, sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import gpcc_b200, oracle
from conftest import load_golden
for name in ('loglik_ragged_OU', 'loglik_3x64_matern52', 'loglik_2band_OU'):
    g = load_golden(name)
    p = gpcc_b200.Problem(g['tb'], g['yb'], g['sb'], g['kernel'])
    ll, grad, info = p.loglik_batch(g['delays'], g['alpha'], g['rho'], want_grad=True)
    print(os.environ.get("GPCC_FRAG_NMAT"), name, "info", info.max(), "ll err", np.max(np.abs(ll - g['loglik']) / np.abs(g['loglik'])),
          "grad err per eval", np.max(np.abs(grad - g['grad']) / np.max(np.abs(g['grad']), axis=1, keepdims=True), axis=1))
    ll1, grad1, info1 = p.loglik_batch(g['delays'][:1], g['alpha'][:1], g['rho'][:1], want_grad=True)
    print("   single eval grad err", np.max(np.abs(grad1 - g['grad'][:1]) / np.max(np.abs(g['grad'][:1]))))
