"""small tensor-core passes (generation 6 at N = 1024, generation 5 at N = 600, generation 4 below 512; the int8 trailing
update) for compute-sanitizer --tool memcheck / racecheck / synccheck"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bayesian_optimization_b200 as b2
from bayesian_optimization_b200 import _lib

rng = np.random.default_rng(0)
for N, D, M in [(200, 5, 300), (600, 4, 600), (1024, 4, 700)]:
    X = rng.uniform(0, 1, (N, D)); y = np.sin(3 * X).sum(axis=1) + 0.3 * rng.standard_normal(N)
    gp = b2.GaussianProcess(mean=b2.constant_trend(D), corr="matern52", thetaL=[1e-5] * D, thetaU=[1e2] * D, nugget=1e-4)
    gp._check_data(X, y)
    gp.engine.set_chol_tc(8, 128)
    print(N, D, gp.fit_fixed(X, y, np.full(D, 2.0), 1.0))
    Xc = rng.uniform(0, 1, (M, D))
    gp.engine.set_precision(_lib.PREC_FAST)
    for prod in (1, 3):
        gp.engine.set_fast_products(prod)
        print("fast", prod, b2.MGFI(model=gp, t=2.0).argmax(Xc, [1.0, 2.0])[1])
    yh2, ms2 = gp.predict(Xc, eval_MSE=True)
print("sanitize_fast done")
