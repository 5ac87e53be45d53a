"""GPU parity of likelihood="restricted" (gpr.py:813-918): value, gradient, fitted state and predict against golden
vectors produced by the reference, then a full restricted fit() checked for self-consistency against the oracle."""
import numpy as np
import pytest

import bayesian_optimization_b200 as b2
from oracle import gp_oracle as go

from conftest import load_golden
from gpu_common import CORR_ARG

pytestmark = pytest.mark.gpu

RESTRICTED = load_golden("restricted")
MODE_KW = {go.MODE_NOISELESS: dict(nugget=None), go.MODE_NOISY: dict(nugget=1e-2), go.MODE_NOISE_ESTIM: dict(nugget=1e-2, noise_estim=True)}


def device_gp(c):
    D = c["X"].shape[1]
    ok = bool(c["ok"])
    mean = b2.constant_trend(D) if ok else b2.constant_trend(D, beta=float(np.ravel(c["beta_in"])[0]))
    nt = c["theta"].size
    return b2.GaussianProcess(mean=mean, corr=CORR_ARG[int(c["corr"])], thetaL=[1e-5] * nt, thetaU=[1e2] * nt,
                              likelihood="restricted", **MODE_KW[int(c["mode"])])


@pytest.mark.parametrize("name", sorted(RESTRICTED))
def test_restricted_value_gradient_predict(name):
    c = RESTRICTED[name]
    gp = device_gp(c)
    gp._check_data(c["X"], c["y"])
    mode = int(c["mode"])
    par = np.r_[c["theta"], float(c["sigma2"])] if mode != go.MODE_NOISE_ESTIM else np.r_[c["theta"], float(c["sigma2"]), float(c["noise_var"])]
    # the noiseless RBF matrix (no nugget) has condition ~1e16: rho^T rho carries only a few digits on any machine
    loose = "_nl_" in name and "rbf" in name
    llf, grad = gp.log_likelihood_restricted(par, eval_grad=True)
    assert llf == pytest.approx(float(c["llf"]), rel=2e-2 if loose else 1e-9)
    if not loose:
        np.testing.assert_allclose(grad.ravel(), c["llf_grad"], rtol=1e-6, atol=1e-7 * np.abs(c["llf_grad"]).max())
        llf2 = gp.fit_fixed_restricted(c["X"], c["y"], c["theta"], float(c["sigma2"]), float(c["noise_var"]))
        assert llf2 == llf and gp.is_fitted
        yh, ms = gp.predict(c["Xc"], eval_MSE=True)
        np.testing.assert_allclose(yh.ravel(), c["yhat"], rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(ms.ravel(), c["mse"], rtol=1e-7, atol=1e-9 * float(c["sigma2"]))
        np.testing.assert_allclose(np.ravel(gp.mean.beta), c["beta"], rtol=1e-8, atol=1e-11)


@pytest.mark.parametrize("mode", ["noisy", "noise_estim"])
def test_restricted_full_fit(mode):
    """fit() with the restricted likelihood: the host L-BFGS-B loop drives the device value + gradient; the result
    must be a valid (finite, improved) likelihood whose state the oracle reproduces at the returned parameters"""
    rng = np.random.default_rng(5)
    N, D = 120, 3
    X = rng.uniform(0, 1, (N, D))
    y = np.sin(5 * X).sum(axis=1) + 0.2 * rng.standard_normal(N)
    y = (y - y.mean()) / y.std()
    kw = dict(nugget=1e-2) if mode == "noisy" else dict(nugget=1e-2, noise_estim=True)
    gp = b2.GaussianProcess(mean=b2.constant_trend(D), corr="squared_exponential", thetaL=[1e-2] * D, thetaU=[1e2] * D,
                            theta0=[1.0] * D, likelihood="restricted", random_start=2, **kw)
    np.random.seed(3)
    gp.fit(X, y)
    assert gp.is_fitted and np.isfinite(gp.log_likelihood_)
    s2, nv = float(gp.sigma2[0]), float(np.atleast_1d(gp.noise_var)[0])
    ora = go.fit_fixed_restricted(X, y, go.CORR_RBF, gp.theta_, s2, nv)
    assert gp.log_likelihood_ == pytest.approx(ora.llf, rel=1e-9)
    start = go.fit_fixed_restricted(X, y, go.CORR_RBF, [1.0] * D, s2, nv)
    assert gp.log_likelihood_ >= start.llf - 1e-9          # the optimiser did not make things worse than theta0
    Xc = rng.uniform(0, 1, (40, D))
    yh, ms = gp.predict(Xc, eval_MSE=True)
    yo, mo = go.predict(ora, Xc)
    np.testing.assert_allclose(yh, yo, rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(ms, mo, rtol=1e-7, atol=1e-9 * s2)
