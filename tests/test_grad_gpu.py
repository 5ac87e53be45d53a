"""GPU parity of the posterior gradient (gpr.py:537-576) and of the acquisition gradients (acquisition_fun.py
return_dx=True) against golden vectors produced by the reference one point at a time, and against the oracle.
Tolerance: 1e-7 relative (+1e-9 absolute) -- the device contracts against L^-1 where the reference solves twice."""
import numpy as np
import pytest

import bayesian_optimization_b200 as b2
from bayesian_optimization_b200 import _lib
from oracle import gp_oracle as go

from conftest import load_golden
from gpu_common import fit_case, oracle_case

pytestmark = pytest.mark.gpu

ACQ_GRAD = load_golden("acq_grad")
MEDIUM = load_golden("medium")
ACQS = (("ei", _lib.ACQ_EI, lambda c: 0.0), ("ucb", _lib.ACQ_UCB, lambda c: float(c["alpha_ucb"])),
        ("mgfi", _lib.ACQ_MGFI, lambda c: float(c["t"])), ("mgfi_big", _lib.ACQ_MGFI, lambda c: 30.0),
        ("epi", _lib.ACQ_PI, lambda c: float(c["eps"])))


def close(a, ref, rtol=1e-7, atol=1e-9):
    fin = np.isfinite(ref)
    np.testing.assert_allclose(a[fin], ref[fin], rtol=rtol, atol=atol * max(1.0, np.abs(ref[fin]).max() if fin.any() else 1.0))


@pytest.mark.parametrize("name", sorted(ACQ_GRAD))
def test_gradients_vs_reference(name):
    c = ACQ_GRAD[name]
    gp, llf = fit_case(c)
    assert llf == pytest.approx(float(c["llf"]), rel=1e-10)
    yh, ms, ydx, mdx = gp.engine.gradient(c["Xc"])
    np.testing.assert_allclose(yh, c["yhat"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(ms, c["mse"], rtol=1e-7, atol=1e-9 * float(c["sigma2"]))
    close(ydx, c["y_dx"])
    close(mdx, c["mse_dx"], rtol=1e-6)
    mn = bool(c["minimize"])
    for key, acq, par in ACQS:
        val, dx = gp.engine.acq_grad(c["Xc"], acq, mn, float(c["plugin"]), par(c))
        np.testing.assert_allclose(val, c[key], rtol=1e-6, atol=1e-300, err_msg=key)
        ref = c[key + "_dx"]
        for i in range(ref.shape[0]):
            if np.all(np.isfinite(ref[i])):
                # rows 0..2 are training points: sd ~ 1e-8 there, and dx divides by it
                np.testing.assert_allclose(dx[i], ref[i], rtol=1e-5, atol=1e-7 * max(1.0, np.abs(ref[i]).max()), err_msg=f"{key} row {i}")
            # else: sd = 0 at a training point and the reference divides by it (UCB, PI): nan / inf upstream, while the
            # device's MSE there may be a few ulp above the clip -- nothing to compare


@pytest.mark.parametrize("name", sorted(k for k, c in MEDIUM.items() if "y_dx" in c and int(c["trend"]) == go.TREND_CONSTANT))
def test_posterior_gradient_medium(name):
    """every estimation mode x OK / SK for the three kernels the reference differentiates"""
    c = MEDIUM[name]
    gp, _ = fit_case(c)
    n = c["y_dx"].shape[0]
    _, _, ydx, mdx = gp.engine.gradient(c["Xc"][:n])
    close(ydx, c["y_dx"], rtol=1e-6 if "_nl_" in name else 1e-7)
    close(mdx, c["mse_dx"], rtol=1e-5 if "_nl_" in name else 1e-6)


def test_python_surface_and_finite_differences():
    """gradient(x) / return_dx shapes as upstream; Matern-5/2 (no reference gradient) against central differences of
    the oracle's own predict"""
    c = ACQ_GRAD["rbf_ok_min"]
    gp, _ = fit_case(c)
    x = c["Xc"][5]
    a, b = gp.gradient(x)
    assert a.shape == (x.size, 1) and b.shape == (x.size, 1)
    np.testing.assert_allclose(a.ravel(), c["y_dx"][5], rtol=1e-7, atol=1e-9)
    with pytest.raises(Exception):
        gp.gradient(c["Xc"][:2])
    f = b2.EI(model=gp, minimize=True)
    v, dx = f(x, return_dx=True)
    assert dx.shape == (1, x.size)
    assert float(v) == pytest.approx(float(c["ei"][5]), rel=1e-6)
    np.testing.assert_allclose(dx.ravel(), c["ei_dx"][5], rtol=1e-5, atol=1e-9)
    vals, dxs = b2.MGFI(model=gp, t=float(c["t"])).value_and_gradient(c["Xc"])
    np.testing.assert_allclose(vals, c["mgfi"], rtol=1e-6, atol=1e-300)
    # Matern-5/2 and -1/2: analytic derivative vs finite differences of the oracle
    for corr_id, corr in ((go.CORR_MATERN52, "matern52"), (go.CORR_MATERN12, "matern12")):
        X, y, theta = c["X"], c["y"], c["theta"]
        D = X.shape[1]
        g2 = b2.GaussianProcess(mean=b2.constant_trend(D), corr=corr, thetaL=[1e-5] * D, thetaU=[1e2] * D, nugget=1e-3)
        g2.fit_fixed(X, y, theta, 0.8)
        ora = go.fit_fixed(X, y, corr_id, theta, go.MODE_NOISY, sigma2=0.8, noise_var=1e-3)
        Xq = c["Xc"][6:12]
        _, _, ydx, mdx = g2.engine.gradient(Xq)
        eps = 1e-6
        for i, xq in enumerate(Xq):
            for d in range(D):
                e = np.zeros(D)
                e[d] = eps
                yp, mp = go.predict(ora, (xq + e)[None, :])
                ym, mm = go.predict(ora, (xq - e)[None, :])
                assert ydx[i, d] == pytest.approx((yp[0, 0] - ym[0, 0]) / (2 * eps), rel=2e-4, abs=1e-6)
                assert mdx[i, d] == pytest.approx((mp[0, 0] - mm[0, 0]) / (2 * eps), rel=2e-4, abs=1e-6)
