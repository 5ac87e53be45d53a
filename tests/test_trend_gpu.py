"""GPU parity of the linear / quadratic regression trends (trend.py:94-142; _compute_aux_var gpr.py:800-808; predict
gpr.py:486-510 with the p-vector u) against golden vectors produced by the reference."""
import numpy as np
import pytest

import bayesian_optimization_b200 as b2
from bayesian_optimization_b200 import _lib
from oracle import gp_oracle as go

from conftest import load_golden
from gpu_common import CORR_ARG

pytestmark = pytest.mark.gpu

TRENDS = load_golden("trends")
MEDIUM = load_golden("medium")
TCLS = {go.TREND_LINEAR: b2.linear_trend, go.TREND_QUADRATIC: b2.quadratic_trend}


def device_fit(c):
    D = c["X"].shape[1]
    mode, ok = int(c["mode"]), bool(c["ok"])
    tcls = TCLS[int(c["trend"])]
    mean = tcls(D) if ok else tcls(D, beta=np.asarray(c["beta_in"], float).ravel())
    kw = dict(mean=mean, corr=CORR_ARG[int(c["corr"])], thetaL=[1e-5] * D, thetaU=[1e2] * D)
    if mode == go.MODE_NOISELESS:
        kw.update(nugget=None)
    elif mode == go.MODE_NOISY:
        kw.update(nugget=float(c["nugget"]))
    else:
        kw.update(nugget=float(c["nugget"]), noise_estim=True)
    gp = b2.GaussianProcess(**kw)
    last = None if mode == go.MODE_NOISELESS else float(c["par_last"])
    return gp, gp.fit_fixed(c["X"], c["y"], c["theta"], last)


CASES = sorted(TRENDS) + ["rbf_ny_ok_lin", "rbf_ny_sk_lin"]


@pytest.mark.parametrize("name", CASES)
def test_trend_fit_and_predict(name):
    c = TRENDS[name] if name in TRENDS else MEDIUM[name]
    gp, llf = device_fit(c)
    rt = 1e-7 if "_nl_" in name else 1e-9
    assert llf == pytest.approx(float(c["llf"]), rel=1e-9 if "_nl_" not in name else 1e-8)
    assert float(gp.sigma2[0]) == pytest.approx(float(c["sigma2"]), rel=rt)
    np.testing.assert_allclose(np.ravel(gp.mean.beta), np.ravel(c["beta"]), rtol=10 * rt, atol=1e-10)
    np.testing.assert_allclose(gp.gamma.ravel(), c["gamma"], rtol=1e-6, atol=1e-8 * np.abs(c["gamma"]).max())
    if bool(c["ok"]):
        np.testing.assert_allclose(gp.G, c["G"], rtol=1e-8, atol=1e-10 * np.abs(c["G"]).max())     # LAPACK's sign convention
        np.testing.assert_allclose(gp.Ft, c["Ft"], rtol=1e-7, atol=1e-9 * np.abs(c["Ft"]).max())
    yh, ms = gp.predict(c["Xc"], eval_MSE=True)
    np.testing.assert_allclose(yh.ravel(), c["yhat"], rtol=rt, atol=1e-9)
    np.testing.assert_allclose(ms.ravel(), c["mse"], rtol=10 * rt, atol=1e-9 * float(c["sigma2"]))
    np.testing.assert_allclose(gp.predict(c["Xc"]).ravel(), c["yhat"], rtol=rt, atol=1e-9)     # mean only
    # acquisition on top of the trend path (values from one device pass; arg-max index as numpy's)
    pl = float(c["plugin"])
    ei = b2.EI(model=gp, minimize=bool(c["minimize"]))
    assert float(ei.plugin) == pytest.approx(pl, rel=1e-14)
    np.testing.assert_allclose(ei(c["Xc"]), c["ei"], rtol=1e-6, atol=1e-300)
    bv, bi = b2.MGFI(model=gp, minimize=bool(c["minimize"]), t=float(c["t"])).argmax(c["Xc"])
    assert int(bi[0]) == int(np.argmax(c["mgfi"]))
    # the tensor-core precision falls back to the float64 path for p > 1
    gp.engine.set_precision(_lib.PREC_FAST)
    bv2, bi2 = b2.MGFI(model=gp, minimize=bool(c["minimize"]), t=float(c["t"])).argmax(c["Xc"])
    assert int(bi2[0]) == int(bi[0]) and bv2[0] == bv[0]


def test_trend_full_fit_and_limits():
    rng = np.random.default_rng(9)
    N, D = 100, 2
    X = rng.uniform(0, 1, (N, D))
    y = 2 * X[:, 0] - X[:, 1] + 0.3 * np.sin(6 * X).sum(axis=1)
    y = (y - y.mean()) / y.std()
    gp = b2.GaussianProcess(mean=b2.linear_trend(D), corr="matern", thetaL=[1e-2] * D, thetaU=[1e2] * D, theta0=[1.0] * D, nugget=1e-4)
    np.random.seed(1)
    gp.fit(X, y)
    assert gp.is_fitted and gp.mean.beta.shape == (D + 1, 1)
    ora = go.fit_fixed(X, y, go.CORR_MATERN32, gp.theta_, go.MODE_NOISY, sigma2=float(gp.sigma2[0]), noise_var=1e-4, trend=go.TREND_LINEAR)
    assert gp.log_likelihood_ == pytest.approx(ora.llf, rel=1e-9)
    Xc = rng.uniform(0, 1, (50, D))
    yh, ms = gp.predict(Xc, eval_MSE=True)
    yo, mo = go.predict(ora, Xc)
    np.testing.assert_allclose(yh, yo, rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(ms, mo, rtol=1e-7, atol=1e-10)
    with pytest.raises(NotImplementedError):                      # p = 78 > 64
        b2.GaussianProcess(mean=b2.quadratic_trend(11), thetaL=[1e-2] * 11, thetaU=[1e2] * 11)._check_data(rng.uniform(0, 1, (30, 11)), rng.uniform(0, 1, 30))
    with pytest.raises(b2.B200BOError):                           # gradient: constant trend only
        gp.engine.gradient(Xc[:2])
