"""GPU: likelihood gradient (gpr.py:994-1038, quirks included) and the full fit() loop (gpr.py:355-417,
:1058-1197) against reference-generated goldens."""
import numpy as np
import pytest

import bayesian_optimization_b200 as b2
from bayesian_optimization_b200 import _lib
from oracle import gp_oracle as go

from conftest import load_golden
from gpu_common import CORR_ARG, device_gp, fit_case

pytestmark = pytest.mark.gpu

MEDIUM = load_golden("medium")
FITS = load_golden("fit_full")
GRAD_CASES = sorted(k for k, c in MEDIUM.items() if "llf_grad" in c and np.isfinite(c["llf"]))


@pytest.mark.parametrize("name", GRAD_CASES)
def test_llf_gradient_matches_reference(name):
    c = MEDIUM[name]
    gp, llf = fit_case(c)
    par = c["theta"] if int(c["mode"]) == go.MODE_NOISELESS else np.r_[c["theta"], float(c["par_last"])]
    l2, g = gp.log_likelihood_concentrated(par, eval_grad=True)
    assert l2 == llf
    ill = 100.0 if "_nl_" in name else 1.0
    np.testing.assert_allclose(g, c["llf_grad"], rtol=1e-7 * ill, atol=1e-8 * ill * np.abs(c["llf_grad"]).max())


def test_llf_gradient_rejected_point_is_zero():
    c = MEDIUM["rejected"]
    gp = device_gp(c, c["X"].shape[1])
    gp._check_data(c["X"], c["y"])
    llf, g = gp.log_likelihood_concentrated(np.r_[c["theta"], float(c["par_last"])], eval_grad=True)
    assert np.isneginf(llf) and g.shape == (len(c["theta"]) + 1, 1) and not g.any()   # gpr.py:981-982


def test_matern52_gradient_is_the_true_derivative():
    """extension: the reference has no Matern-5/2 theta-gradient (gpr.py:758-759); ours must match central
    finite differences of the device likelihood (noise_estim mode, where the analytic form is exact, App. A g1)"""
    c = MEDIUM["m52_ne_sk"]
    gp = device_gp(c, c["X"].shape[1])
    gp._check_data(c["X"], c["y"])
    par = np.r_[c["theta"], float(c["par_last"])]
    _, g = gp.log_likelihood_concentrated(par, eval_grad=True)
    for i in range(len(par)):
        h = 1e-6 * par[i]
        p1, p2 = par.copy(), par.copy()
        p1[i] += h
        p2[i] -= h
        fd = (gp.log_likelihood_concentrated(p1) - gp.log_likelihood_concentrated(p2)) / (2 * h)
        assert g[i] == pytest.approx(fd, rel=2e-4, abs=1e-6), i


def _fit_kwargs(c):
    D = c["X"].shape[1]
    mode = int(c["mode"])
    kw = dict(corr=CORR_ARG[int(c["corr"])], thetaL=[1e-2] * D, thetaU=[1e2] * D, theta0=[1.0] * D, random_start=2)
    if mode == go.MODE_NOISELESS:
        kw.update(nugget=None)
    elif mode == go.MODE_NOISY:
        kw.update(nugget=1e-2)
    else:
        kw.update(nugget=1e-2, noise_estim=True)
    return D, mode, kw


@pytest.mark.parametrize("name", sorted(FITS))
def test_full_fit_matches_reference_over_seeds(name):
    """fit() = host L-BFGS-B restarts (global numpy RNG) on the device likelihood + gradient (gpr.py:1058-1197).
    The reference feeds L-BFGS-B an inconsistent gradient (quirk g4) and draws sigma2 / restarts from the global
    RNG, so its OWN result varies wildly with the seed (golden ``llf_seeds``: e.g. -28 / -152 / -6863 for one
    data set) and last-bit differences of the objective change the path.  Parity is therefore on the outcome over
    the same 12 seeds: the best optimum is the reference's best, and as many runs reach it."""
    c = FITS[name]
    D, mode, kw = _fit_kwargs(c)
    ref = np.asarray(c["llf_seeds"], dtype=float)
    ours = []
    for seed in range(len(ref)):
        gp = b2.GaussianProcess(mean=b2.constant_trend(D), **kw)
        np.random.seed(seed)
        assert gp.fit(c["X"], c["y"]) is gp and gp.is_fitted
        assert np.isfinite(gp.log_likelihood_)
        ours.append(gp.log_likelihood_)
    ours = np.array(ours)
    best = ref.max()
    assert ours.max() >= best - 1e-5 * abs(best), (ours, ref)
    # as many runs end (within 0.5 %) at the best optimum as for the reference, give or take two of the twelve
    good = lambda v: int((v >= best - 5e-3 * abs(best)).sum())
    assert good(ours) >= good(ref) - 2, (ours, ref)


@pytest.mark.parametrize("name", sorted(FITS))
def test_full_fit_state_is_consistent(name):
    c = FITS[name]
    D, mode, kw = _fit_kwargs(c)
    gp = b2.GaussianProcess(mean=b2.constant_trend(D), **kw)
    np.random.seed(5)
    gp.fit(c["X"], c["y"])
    yh, ms = gp.predict(c["Xc"], eval_MSE=True)
    if abs(gp.log_likelihood_ - float(c["llf"])) <= 1e-6 * abs(float(c["llf"])):   # same optimum as the reference's seed-5 run
        np.testing.assert_allclose(yh.ravel(), c["yhat"], rtol=0, atol=2e-2 * np.abs(c["yhat"]).max())
    # the fitted state is the fixed-theta fit at the optimum found (a fresh trend object: fit() stored beta in gp's)
    last = None if mode == go.MODE_NOISELESS else gp._par_last
    gp2 = b2.GaussianProcess(mean=b2.constant_trend(D), **kw)
    assert gp2.fit_fixed(c["X"], c["y"], gp.theta_, last) == gp.log_likelihood_
    y2, m2 = gp2.predict(c["Xc"], eval_MSE=True)
    np.testing.assert_array_equal(y2, yh)
    np.testing.assert_array_equal(m2, ms)
    # warm start from theta_ (gpr.py:1095-1096); only theta is warm-started, sigma2 / alpha are redrawn (:1101-1107),
    # so "not worse" holds where theta is the whole parameter vector
    l1 = gp.log_likelihood_
    gp.fit(c["X"], c["y"])
    assert gp.is_fitted and np.isfinite(gp.log_likelihood_)
    if mode == go.MODE_NOISELESS:
        assert gp.log_likelihood_ >= l1 - 1e-6 * abs(l1)


def test_isotropic_theta_gradient_quirk():
    """one theta for D > 1: the reference indexes its gradient tensor by parameter index (gpr.py:1004-1005),
    i.e. the theta component only sees feature 0 -- reproduced, checked against the ARD gradient's slice"""
    c = MEDIUM["rbf_nl_sk"]
    X, y = c["X"], c["y"]
    D = X.shape[1]
    gp = b2.GaussianProcess(mean=b2.constant_trend(D, beta=0.1), thetaL=[1e-5], thetaU=[1e2], nugget=None)
    gp._check_data(X, y)
    _, g_iso = gp.log_likelihood_concentrated(np.array([0.4]), eval_grad=True)
    gp2 = b2.GaussianProcess(mean=b2.constant_trend(D, beta=0.1), thetaL=[1e-5] * D, thetaU=[1e2] * D, nugget=None)
    gp2._check_data(X, y)
    _, g_ard = gp2.log_likelihood_concentrated(np.full(D, 0.4), eval_grad=True)
    assert g_iso.shape == (1,) and g_iso[0] == pytest.approx(g_ard[0], rel=1e-12)
