"""CPU oracle for the GPCC.jl hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy/scipy restatement of the reference algorithm (GPCC.jl v0.1.35), used only by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs as the checker and the reported CPU baseline.  Nothing under ``gpcc_b200/`` imports it.

PARITY UNPINNED: the reference ships an empty test suite (``test/runtests.jl:4-6``), no golden
vectors and no fixtures, Julia is not installed here, and the dependencies ``MiscUtil``, ``Optim``,
``Distributions``, ``BlockArrays`` are not vendored under /root/reference.  The oracle is therefore
validated against first principles instead (see tests/test_oracle.py): 50-digit mpmath evaluation of
the log-likelihood, finite-difference gradients, the Woodbury route to the marginalised likelihood
and to ``postb``, and the element-wise statement of K+Sobs+B in
``src/gpccfixdelay_verifications.jl:130-150``.

Modules
  kernels      src/util.jl:15-52               scalar kernels OU / rbf / matern32 / matern52 (+ d/drho)
  covariance   src/delayedCovariance.jl:1-38   delayed block covariance
  model        src/gpccfixdelay_marginaliseb.jl:85-141, 235-343   data prep, objective, gradient, postb, pred
  fit          src/gpccfixdelay_marginaliseb.jl:160-226           starts, screening, Nelder-Mead / L-BFGS
  posterior    src/getprobabilities.jl:1-20, src/uniformpriordelay.jl:10-16
  simulate     src/simulatedata.jl:96-162      synthetic light curves of the reference's shape
"""
from .kernels import KERNELS, KERNEL_IDS, kernel_value, kernel_drho          # noqa: F401
from .covariance import delayed_covariance                                    # noqa: F401
from .model import (Problem, makepositive, invmakepositive, transformbetween,  # noqa: F401
                    invtransformbetween)
from .posterior import getprobabilities, uniformpriordelay, Uniform           # noqa: F401
from .simulate import simulatedata, simulatetwolightcurves, simulatethreelightcurves, synthetic_bands  # noqa: F401
from .fit import gpcc, initial_solutions, performcv, cv_folds                  # noqa: F401
