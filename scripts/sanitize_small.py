"""tiny end-to-end pass for compute-sanitizer (memcheck): every kernel of the path at ragged sizes"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bayesian_optimization_b200 as b2

rng = np.random.default_rng(0)
for N, D, M in [(70, 3, 37), (200, 5, 300)]:
    X = rng.uniform(0, 1, (N, D)); y = np.sin(3 * X).sum(axis=1) + 0.3 * rng.standard_normal(N)
    for corr in ("squared_exponential", "matern52"):
        gp = b2.GaussianProcess(mean=b2.constant_trend(D), corr=corr, thetaL=[1e-5] * D, thetaU=[1e2] * D, nugget=1e-4)
        print(N, D, corr, gp.fit_fixed(X, y, np.full(D, 2.0), 1.0))
        assert gp.is_fitted
        Xc = rng.uniform(0, 1, (M, D))
        yh, ms = gp.predict(Xc, eval_MSE=True)
        print(b2.MGFI(model=gp, t=2.0).argmax(Xc, [1.0, 2.0, 3.0]), b2.EI(model=gp)(Xc)[:3], gp.C.shape)
print("sanitize_small done")
