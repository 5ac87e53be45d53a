"""tiny end-to-end pass for compute-sanitizer (memcheck): every kernel of the path at ragged sizes"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bayesian_optimization_b200 as b2
from bayesian_optimization_b200 import _lib

rng = np.random.default_rng(0)
for N, D, M in [(70, 3, 37), (200, 5, 300), (600, 4, 300)]:
    X = rng.uniform(0, 1, (N, D)); y = np.sin(3 * X).sum(axis=1) + 0.3 * rng.standard_normal(N)
    for corr in ("squared_exponential", "matern52"):
        gp = b2.GaussianProcess(mean=b2.constant_trend(D), corr=corr, thetaL=[1e-5] * D, thetaU=[1e2] * D, nugget=1e-4)
        print(N, D, corr, gp.fit_fixed(X, y, np.full(D, 2.0), 1.0))
        assert gp.is_fitted
        Xc = rng.uniform(0, 1, (M, D))
        yh, ms = gp.predict(Xc, eval_MSE=True)
        print(b2.MGFI(model=gp, t=2.0).argmax(Xc, [1.0, 2.0, 3.0]), b2.EI(model=gp)(Xc)[:3], gp.C.shape)
        # likelihood gradient, posterior / acquisition gradients
        llf, g = gp.log_likelihood_concentrated(np.r_[np.full(D, 2.0), 1.0], eval_grad=True)
        _, _, ydx, mdx = gp.engine.gradient(Xc)
        v, dx = b2.EI(model=gp).value_and_gradient(Xc)
        print("grad", float(np.abs(g).max()), float(np.abs(ydx).max()), float(np.abs(dx).max()))
        # tensor-core flavour: generations 4 (partial replay), 5 (N >= 512) and 6 (N % 256 == 0, N >= 1024) of the fused kernel
        gp.engine.set_precision(_lib.PREC_FAST)
        for gen in (4, 5, 6):
            gp.engine.set_fast_kernel(gen)
            if gen == 4:
                gp.engine.set_replay(64, 2)
            for prod in (1, 3):
                gp.engine.set_fast_products(prod)
                print("fast", gen, prod, b2.MGFI(model=gp, t=2.0).argmax(Xc, [1.0, 2.0])[1])
            yh2, ms2 = gp.predict(Xc, eval_MSE=True)
            print("fast predict diff", float(np.abs(yh2 - yh).max()), "of", float(np.abs(yh).max()))
    # restricted likelihood, linear / quadratic trends, generalized_exponential, two targets
    gr = b2.GaussianProcess(mean=b2.constant_trend(D), corr="matern", thetaL=[1e-5] * D, thetaU=[1e2] * D, nugget=1e-3, likelihood="restricted")
    gr._check_data(X, y)
    print("restricted", gr.log_likelihood_restricted(np.r_[np.full(D, 2.0), 0.8], eval_grad=True)[0])
    for mean in (b2.linear_trend(D), b2.quadratic_trend(D)):
        gt = b2.GaussianProcess(mean=mean, corr="squared_exponential", thetaL=[1e-5] * D, thetaU=[1e2] * D, nugget=1e-3)
        print("trend", type(mean).__name__, gt.fit_fixed(X, y, np.full(D, 2.0), 1.0), gt.predict(Xc, eval_MSE=True)[1][:2].ravel())
    ge = b2.GaussianProcess(mean=b2.constant_trend(D), corr="generalized_exponential", thetaL=[1e-5] * (D + 1), thetaU=[1e2] * (D + 1), nugget=1e-3)
    print("genexp", ge.fit_fixed(X, y, np.r_[np.full(D, 2.0), 1.5], 1.0), ge.predict(Xc[:5]).ravel()[:2])
    gm = b2.GaussianProcess(mean=b2.constant_trend(D), corr="matern", thetaL=[1e-5] * D, thetaU=[1e2] * D, nugget=1e-3)
    print("multi", gm.fit_fixed(X, np.c_[y, y[::-1]], np.full(D, 2.0), 1.0), gm.predict(Xc[:5]).shape)
print("sanitize_small done")
