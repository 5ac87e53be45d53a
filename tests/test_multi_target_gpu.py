"""GPU parity for multi-target y (N, k): per-target state over one factorisation (gpr.py:799-808, :934-979), the summed
likelihood (:1040), (M, k) predictions (:490, :502-505) -- against golden vectors of the reference (simple kriging: the
only multi-target form that runs upstream; with beta estimated its trend setter raises, trend.py:25-28) and, for
ordinary kriging, against the oracle column by column."""
import numpy as np
import pytest

import bayesian_optimization_b200 as b2
from oracle import gp_oracle as go

from conftest import load_golden
from gpu_common import CORR_ARG

pytestmark = pytest.mark.gpu

MT = load_golden("multi_target")


def build(c, ok):
    D = c["X"].shape[1]
    mode = int(c["mode"])
    mean = b2.constant_trend(D) if ok else b2.constant_trend(D, beta=0.1)
    kw = dict(mean=mean, corr=CORR_ARG[int(c["corr"])], thetaL=[1e-5] * D, thetaU=[1e2] * D)
    if mode == go.MODE_NOISELESS:
        kw.update(nugget=None)
    elif mode == go.MODE_NOISY:
        kw.update(nugget=float(c["nugget"]))
    else:
        kw.update(nugget=float(c["nugget"]), noise_estim=True)
    return b2.GaussianProcess(**kw), (None if mode == go.MODE_NOISELESS else float(c["par_last"]))


@pytest.mark.parametrize("name", sorted(MT))
def test_multi_target_vs_reference(name):
    c = MT[name]
    gp, last = build(c, False)
    llf = gp.fit_fixed(c["X"], c["y"], c["theta"], last)
    rt = 1e-7 if "_nl_" in name else 1e-9
    assert llf == pytest.approx(float(c["llf"]), rel=rt)
    np.testing.assert_allclose(np.ravel(gp.sigma2), c["sigma2"], rtol=rt)
    np.testing.assert_allclose(np.ravel(gp.noise_var), np.ravel(c["noise_var"]), rtol=rt, atol=1e-300)
    np.testing.assert_allclose(gp.gamma, c["gamma"], rtol=1e-6, atol=1e-8 * np.abs(c["gamma"]).max())
    yh, ms = gp.predict(c["Xc"], eval_MSE=True)
    assert yh.shape == c["yhat"].shape == (c["Xc"].shape[0], 2)
    np.testing.assert_allclose(yh, c["yhat"], rtol=rt, atol=1e-9)
    np.testing.assert_allclose(ms, c["mse"], rtol=10 * rt, atol=1e-9)
    np.testing.assert_allclose(gp.predict(c["Xc"]), c["yhat"], rtol=rt, atol=1e-9)


def test_multi_target_ordinary_kriging_and_fit():
    """beta estimated per target (upstream raises here); each column must equal the single-target oracle"""
    c = MT["mt_m32_ny_sk"]
    gp, last = build(c, True)
    llf = gp.fit_fixed(c["X"], c["y"], c["theta"], last)
    oras = [go.fit_fixed(c["X"], c["y"][:, t], go.CORR_MATERN32, c["theta"], go.MODE_NOISY, sigma2=last, noise_var=float(c["nugget"]))
            for t in range(2)]
    assert llf == pytest.approx(sum(o.llf for o in oras), rel=1e-10)
    assert gp.mean.beta.shape == (1, 2)
    np.testing.assert_allclose(np.ravel(gp.mean.beta), [o.beta[0, 0] for o in oras], rtol=1e-8)
    yh, ms = gp.predict(c["Xc"], eval_MSE=True)
    for t, o in enumerate(oras):
        yo, mo = go.predict(o, c["Xc"])
        np.testing.assert_allclose(yh[:, t], yo.ravel(), rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(ms[:, t], mo.ravel(), rtol=1e-8, atol=1e-10)
    # full fit(): the host L-BFGS-B loop on the summed device likelihood
    g2 = b2.GaussianProcess(mean=b2.constant_trend(3), corr="squared_exponential", thetaL=[1e-2] * 3, thetaU=[1e2] * 3,
                            theta0=[1.0] * 3, nugget=1e-2, random_start=1)
    np.random.seed(2)
    g2.fit(c["X"], c["y"])
    assert g2.is_fitted and np.isfinite(g2.log_likelihood_) and g2.predict(c["Xc"]).shape == (c["Xc"].shape[0], 2)
    s2 = float(np.ravel(g2.sigma2)[0])
    start = sum(go.fit_fixed(c["X"], c["y"][:, t], go.CORR_RBF, [1.0] * 3, go.MODE_NOISY, sigma2=s2, noise_var=1e-2).llf for t in range(2))
    assert g2.log_likelihood_ >= start - 1e-9
    with pytest.raises(NotImplementedError):
        g2.gradient(c["Xc"][0])
