"""GPU: b200bo_append / GaussianProcess.update(reoptimize=False) -- the bordering update of L and L^-1 that upstream left
as a TODO (gpr.py:419-422).  The state after appending must be the state of a fresh fixed-parameter fit of all the
data: likelihood 1e-10, L and gamma 1e-9 relative, posterior 1e-9, and the oracle (CPU restatement of gpr.py:920-991 on
the full data) agrees with both.  Cases cross the 64-row block size, the 128-row padding (pitch change), several
kernels, estimation modes, fixed and estimated beta, the restricted likelihood, and repeated appends."""
import numpy as np
import pytest

import bayesian_optimization_b200 as b2
from bayesian_optimization_b200 import _lib, workloads
from oracle import gp_oracle as go

pytestmark = pytest.mark.gpu

CASES = [  # (N0, m, D, corr, corr id, nugget / None, noise_estim, ordinary kriging)
    (100, 1, 3, "squared_exponential", go.CORR_RBF, 1e-6, False, True),
    (100, 28, 3, "squared_exponential", go.CORR_RBF, 1e-6, False, True),       # stays inside the 128 padding
    (120, 20, 5, "matern52", go.CORR_MATERN52, 1e-6, False, True),             # crosses it: new pitch
    (500, 64, 8, "matern32", go.CORR_MATERN32, 1e-4, False, False),            # one full block, fixed beta
    (300, 150, 4, "squared_exponential", go.CORR_RBF, 1e-2, True, True),        # three blocks, noise_estim
    (256, 5, 6, "absolute_exponential", go.CORR_ABSEXP, None, False, True),     # noiseless mode
    (1000, 32, 16, "matern52", go.CORR_MATERN52, 1e-6, False, True),
]


def make(D, corr, nugget, noise_estim, ok, likelihood="concentrated"):
    mean = b2.constant_trend(D) if ok else b2.constant_trend(D, beta=0.3)
    return b2.GaussianProcess(mean=mean, corr=corr, thetaL=[1e-5] * D, thetaU=[1e2] * D, nugget=nugget, noise_estim=noise_estim,
                              likelihood=likelihood)


def fit_args(gp, theta):
    if gp.estimation_mode == "noiseless":
        return theta, None
    return theta, (0.7 if gp.estimation_mode == "noise_estim" else 1.0)


@pytest.mark.parametrize("N0,m,D,corr,corr_id,nugget,noise_estim,ok", CASES)
def test_append_equals_refit(N0, m, D, corr, corr_id, nugget, noise_estim, ok):
    X, y, theta = workloads.canonical_problem(N0 + m, D)
    Xc = workloads.canonical_candidates(300, D)
    a = make(D, corr, nugget, noise_estim, ok)
    th, last = fit_args(a, theta)
    assert np.isfinite(a.fit_fixed(X[:N0], y[:N0], th, last))
    a.update(X, y, reoptimize=False)
    assert a.engine.N == N0 + m and a.X.shape == (N0 + m, D)
    b = make(D, corr, nugget, noise_estim, ok)
    llf_b = b.fit_fixed(X, y, th, last)
    assert abs(a.log_likelihood_ - llf_b) <= 1e-10 * abs(llf_b), (a.log_likelihood_, llf_b)
    La, Lb = a.C, b.C
    assert np.abs(La - Lb).max() <= 1e-9 * np.abs(Lb).max()
    assert np.abs(a.gamma - b.gamma).max() <= 1e-8 * np.abs(b.gamma).max()
    assert np.allclose(np.ravel(a.sigma2), np.ravel(b.sigma2), rtol=1e-10)
    ya, ma = a.predict(Xc, eval_MSE=True)
    yb, mb = b.predict(Xc, eval_MSE=True)
    assert np.abs(ya - yb).max() <= 1e-9 * max(1.0, np.abs(yb).max())
    assert np.abs(ma - mb).max() <= 1e-9 * float(np.ravel(b.sigma2)[0])
    # and both are the oracle's model of the full data
    mode = {"noiseless": go.MODE_NOISELESS, "noisy": go.MODE_NOISY, "noise_estim": go.MODE_NOISE_ESTIM}[a.estimation_mode]
    kw = {} if ok else dict(beta_fixed=[0.3])
    if mode == go.MODE_NOISY:
        kw.update(sigma2=last, noise_var=nugget)
    elif mode == go.MODE_NOISE_ESTIM:
        kw.update(alpha=last)
    ora = go.fit_fixed(X, y, corr_id, theta, mode, **kw)
    assert abs(a.log_likelihood_ - ora.llf) <= 1e-9 * abs(ora.llf)
    yo, mo = go.predict_chunked(ora, Xc, 128)
    assert np.abs(ya - yo).max() <= 1e-8 * max(1.0, np.abs(yo).max())
    assert np.abs(ma - mo).max() <= 1e-8 * ora.sigma2
    # the tensor-core state is rebuilt on the appended factor
    a.engine.set_precision(_lib.PREC_FAST)
    f = b2.EI(model=a, minimize=True)
    bv, bi = f.argmax(Xc)
    a.engine.set_precision(_lib.PREC_FP64)
    bv2, bi2 = f.argmax(Xc)
    assert int(bi[0]) == int(bi2[0]) and abs(bv[0] - bv2[0]) <= 1e-11 * abs(bv2[0]) + 1e-300


def test_repeated_appends_like_a_bo_loop():
    """q points per iteration appended 12 times (y re-standardised each time, base.py:437): no drift against a refit"""
    D, q = 6, 5
    X, yraw, theta = workloads.canonical_problem(200 + 12 * q, D)
    a = make(D, "matern52", 1e-6, False, True)
    n = 200
    std = lambda v: (v - v.mean()) / v.std()  # noqa: E731
    a.fit_fixed(X[:n], std(yraw[:n]), theta, 1.0)
    for _ in range(12):
        n += q
        a.update(X[:n], std(yraw[:n]), reoptimize=False)
    b = make(D, "matern52", 1e-6, False, True)
    llf = b.fit_fixed(X[:n], std(yraw[:n]), theta, 1.0)
    assert abs(a.log_likelihood_ - llf) <= 1e-9 * abs(llf)
    Xc = workloads.canonical_candidates(200, D)
    ya, ma = a.predict(Xc, eval_MSE=True)
    yb, mb = b.predict(Xc, eval_MSE=True)
    assert np.abs(ya - yb).max() <= 1e-8 and np.abs(ma - mb).max() <= 1e-8


def test_append_restricted_likelihood():
    D = 4
    X, y, theta = workloads.canonical_problem(180, D)
    a = make(D, "squared_exponential", 1e-4, False, True, likelihood="restricted")
    a.fit_fixed_restricted(X[:150], y[:150], theta, 0.9)
    a.update(X, y, reoptimize=False)
    b = make(D, "squared_exponential", 1e-4, False, True, likelihood="restricted")
    llf = b.fit_fixed_restricted(X, y, theta, 0.9)
    assert abs(a.log_likelihood_ - llf) <= 1e-10 * abs(llf)
    Xc = workloads.canonical_candidates(100, D)
    assert np.abs(a.predict(Xc) - b.predict(Xc)).max() <= 1e-9


def test_append_timing_vs_refit():
    """the point of the update: O(m N^2) against O(N^3) -- at N = 4096 it must be several times cheaper than a refit"""
    import time

    D, N0, m = 16, 4064, 32
    X, y, theta = workloads.canonical_problem(N0 + m, D)
    a = make(D, "matern52", 1e-6, False, True)
    a.fit_fixed(X[:N0], y[:N0], theta, 1.0)
    a.update(X, y, reoptimize=False)           # warm-up (allocations)
    a.fit_fixed(X[:N0], y[:N0], theta, 1.0)
    t0 = time.perf_counter()
    a.update(X, y, reoptimize=False)
    t_app = time.perf_counter() - t0
    t_dev_app = a.engine.fit_timings()[0]
    t0 = time.perf_counter()
    a.fit_fixed(X, y, theta, 1.0)
    t_fit = time.perf_counter() - t0
    t_dev_fit = a.engine.fit_timings()[0]
    print(f"append {m} rows at N={N0}: device {t_dev_app:.3f} ms (wall {1e3 * t_app:.2f} ms); refit: device {t_dev_fit:.3f} ms (wall {1e3 * t_fit:.2f} ms)")
    assert t_dev_app < 0.5 * t_dev_fit
