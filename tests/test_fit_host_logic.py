"""CPU: the host side of fit() -- parameter lists per estimation mode, the L-BFGS-B restart loop with the reference's
RNG consumption, noise escalation, the restricted likelihood, multi-target plumbing, pickling -- with the device
replaced by an oracle-backed stand-in (tests/fake_engine.py; test infrastructure, monkeypatched in, never selectable
from the product).  The same assertions run against the real device in tests/test_fit_gpu.py."""
import pickle

import numpy as np
import pytest

import bayesian_optimization_b200 as b2
from bayesian_optimization_b200 import gp as gp_module
from oracle import gp_oracle as go

from conftest import load_golden
from fake_engine import FakeEngine

FITS = load_golden("fit_full")


@pytest.fixture(autouse=True)
def oracle_backed_engine(monkeypatch):
    monkeypatch.setattr(gp_module, "Engine", FakeEngine)


def _kwargs(c):
    D, mode = c["X"].shape[1], int(c["mode"])
    kw = dict(corr={go.CORR_RBF: "squared_exponential", go.CORR_MATERN32: "matern"}[int(c["corr"])],
              thetaL=[1e-2] * D, thetaU=[1e2] * D, theta0=[1.0] * D, random_start=2)
    kw.update({go.MODE_NOISELESS: dict(nugget=None), go.MODE_NOISY: dict(nugget=1e-2),
               go.MODE_NOISE_ESTIM: dict(nugget=1e-2, noise_estim=True)}[mode])
    return D, mode, kw


@pytest.mark.parametrize("name", sorted(FITS))
def test_fit_loop_reaches_the_reference_optimum(name):
    """the reference's own optimum over seeds 0..5 (golden ``llf_seeds``, see tests/test_fit_gpu.py for why parity is on
    the outcome over seeds): our best run is its best, and the seed-5 run adopts a consistent state"""
    c = FITS[name]
    D, mode, kw = _kwargs(c)
    ref = np.asarray(c["llf_seeds"], dtype=float)[:6]
    ours = []
    for seed in range(6):
        gp = b2.GaussianProcess(mean=b2.constant_trend(D), **kw)
        np.random.seed(seed)
        assert gp.fit(c["X"], c["y"]) is gp and gp.is_fitted and np.isfinite(gp.log_likelihood_)
        ours.append(gp.log_likelihood_)
    best = ref.max()
    assert max(ours) >= best - 1e-5 * abs(best), (ours, ref)
    assert gp.eval_count > 0 and gp.theta_.shape == (D,)
    assert set(gp.par) == {"theta"} | ({"sigma2"} if mode == go.MODE_NOISY else set()) | ({"alpha"} if mode == go.MODE_NOISE_ESTIM else set())
    # the adopted state is the fixed-theta fit at the optimum
    last = None if mode == go.MODE_NOISELESS else gp._par_last
    ora = go.fit_fixed(c["X"], c["y"], int(c["corr"]), gp.theta_, mode, **(
        {} if mode == go.MODE_NOISELESS else dict(sigma2=last, noise_var=1e-2) if mode == go.MODE_NOISY else dict(alpha=last)))
    assert gp.log_likelihood_ == pytest.approx(ora.llf, rel=1e-12)
    yh, ms = gp.predict(c["Xc"], eval_MSE=True)
    assert yh.shape == ms.shape == (c["Xc"].shape[0], 1)
    np.testing.assert_allclose(np.ravel(gp.mean.beta), ora.beta.ravel(), rtol=1e-12)


def test_same_seed_same_path_as_reference_for_the_first_evaluations():
    """RNG consumption: with the global seed fixed, the first likelihood point the loop evaluates is the reference's
    (theta0 given -> only sigma2 is drawn, gpr.py:1101-1107); golden eval_count pins the budget arithmetic"""
    c = FITS["rbf_ny"]
    D, mode, kw = _kwargs(c)
    gp = b2.GaussianProcess(mean=b2.constant_trend(D), **kw)
    np.random.seed(5)
    gp.fit(c["X"], c["y"])
    assert gp.eval_count <= 200 * (D + 1) + 60                       # eval_budget = 200 n_par (+ the last restart's overshoot)
    np.random.seed(5)
    from bayesian_optimization_b200.hyperopt import hyperparameter_bounds
    b = np.log10(hyperparameter_bounds(gp, ["theta", "sigma2"]))
    first_sigma2 = 10 ** np.random.uniform(b[D:, 0], b[D:, 1])       # the draw fit() makes before its first evaluation
    assert 1e-5 <= first_sigma2[0] <= max(1e-3, c["y"].std() ** 2)


def test_noise_escalation_on_a_singular_matrix(capsys):
    """duplicate points without a nugget: the likelihood is -inf, fit() switches to "noisy" with noise_var = 1e-5 and
    multiplies by 10 until it works (gpr.py:384-399)"""
    rng = np.random.default_rng(0)
    X = rng.uniform(0, 1, (30, 2))
    X[1] = X[0]
    y = np.sin(4 * X).sum(axis=1)
    y[1] = y[0] + 0.3
    gp = b2.GaussianProcess(mean=b2.constant_trend(2), corr="squared_exponential", thetaL=[1e-1] * 2, thetaU=[1e1] * 2,
                            theta0=[1.0] * 2, nugget=None, random_start=1)
    np.random.seed(0)
    gp.fit(X, y)
    assert gp.is_fitted and gp.estimation_mode == "noisy" and float(np.atleast_1d(gp.noise_var)[0]) >= 1e-5
    assert "Increasing nugget" in capsys.readouterr().out


@pytest.mark.parametrize("mode", ["noiseless", "noisy", "noise_estim"])
def test_restricted_fit_loop(mode):
    c = FITS["rbf_ny"]
    D = c["X"].shape[1]
    kw = dict(nugget=None) if mode == "noiseless" else dict(nugget=1e-2) if mode == "noisy" else dict(nugget=1e-2, noise_estim=True)
    gp = b2.GaussianProcess(mean=b2.constant_trend(D), corr="squared_exponential", thetaL=[1e-2] * D, thetaU=[1e2] * D,
                            theta0=[1.0] * D, likelihood="restricted", random_start=1, **kw)
    if mode == "noiseless":
        gp.thetaL = np.full(D, 5.0)    # keep the nugget-free matrix well conditioned
        gp.theta0 = np.full(D, 10.0)
    np.random.seed(1)
    gp.fit(c["X"], c["y"])
    assert gp.is_fitted and np.isfinite(gp.log_likelihood_)
    assert set(gp.par) == {"theta", "sigma2"} | ({"noise_var"} if mode == "noise_estim" else set())   # gpr.py:1073-1084
    s2, nv = float(gp.sigma2[0]), float(np.atleast_1d(gp.noise_var)[0])
    ora = go.fit_fixed_restricted(c["X"], c["y"], go.CORR_RBF, gp.theta_, s2, nv)
    assert gp.log_likelihood_ == pytest.approx(ora.llf, rel=1e-12)
    assert gp._restricted_par == (s2, nv)


def test_multi_target_plumbing_and_pickle():
    c = FITS["m32_ny"]
    D = c["X"].shape[1]
    Y = np.c_[c["y"], c["y"][::-1]]
    gp = b2.GaussianProcess(mean=b2.constant_trend(D), corr="matern", thetaL=[1e-2] * D, thetaU=[1e2] * D, nugget=1e-2)
    llf = gp.fit_fixed(c["X"], Y, [1.0] * D, 0.8)
    oras = [go.fit_fixed(c["X"], Y[:, t], go.CORR_MATERN32, [1.0] * D, go.MODE_NOISY, sigma2=0.8, noise_var=1e-2) for t in range(2)]
    assert llf == pytest.approx(sum(o.llf for o in oras), rel=1e-12)
    assert gp.mean.beta.shape == (1, 2) and gp.gamma.shape == (c["X"].shape[0], 2) and gp.sigma2.shape == (2,)
    yh, ms = gp.predict(c["Xc"], eval_MSE=True)
    assert yh.shape == ms.shape == (c["Xc"].shape[0], 2)
    with pytest.raises(NotImplementedError):
        gp.gradient(c["Xc"][0])
    # pickling drops the device handles; the deterministic factorisation is redone on first use
    g2 = pickle.loads(pickle.dumps(gp))
    assert g2._engine is None and all(s._engine is None for s in g2._sub)
    y2, m2 = g2.predict(c["Xc"], eval_MSE=True)
    np.testing.assert_array_equal(y2, yh)
    np.testing.assert_array_equal(m2, ms)
    # single target as well
    g1 = b2.GaussianProcess(mean=b2.constant_trend(D), corr="matern", thetaL=[1e-2] * D, thetaU=[1e2] * D, nugget=1e-2)
    g1.fit_fixed(c["X"], c["y"], [1.0] * D, 0.8)
    g3 = pickle.loads(pickle.dumps(g1))
    np.testing.assert_array_equal(g3.predict(c["Xc"]), g1.predict(c["Xc"]))


def test_error_conventions_of_predict_and_acquisition():
    c = FITS["rbf_ny"]
    D = c["X"].shape[1]
    gp = b2.GaussianProcess(mean=b2.constant_trend(D), corr="squared_exponential", thetaL=[1e-2] * D, thetaU=[1e2] * D, nugget=1e-2)
    gp.fit_fixed(c["X"], c["y"], [1.0] * D, 0.8)
    with pytest.raises(ValueError, match="number of features"):
        gp.predict(np.zeros((3, D + 1)))                                   # gpr.py:467-475
    with pytest.raises(Exception, match="batch_size"):
        gp.predict(c["Xc"], batch_size=0)                                  # gpr.py:515-516
    ei = b2.EI(model=gp)
    v, dx = ei(c["Xc"][0], return_dx=True)                                 # one point: (value, dx (1, D)) as upstream
    assert np.ndim(v) == 0 and dx.shape == (1, D)
    assert ei(c["Xc"]).shape == (c["Xc"].shape[0],)
    # a rejected likelihood is -inf with a zero gradient, never an exception (gpr.py:946-982): tiny total variance -> llf > 0
    llf, g = gp.log_likelihood_concentrated([1e-2] * D + [1e-5], eval_grad=True)
    assert np.isfinite(llf) or (llf == -np.inf and not np.any(g))


MEDIUM = load_golden("medium")


@pytest.mark.parametrize("name", ["rbf_ny_ok", "m52_ny_ok_max", "rbf_ny_ok_max", "abs_ne_sk", "rbf_ny_ok_lin"])
def test_acquisition_classes_plumbing_vs_reference(name):
    """EI / MGFI / UCB / EpsilonPI objects of the mirror (plug-in sign, minimise / maximise, the t cap, argmax) against
    the reference's row-by-row values -- the arithmetic behind them is the oracle here, the device in test_parity_gpu.py"""
    import functools

    from gpu_common import CORR_ARG

    c = MEDIUM[name]
    D = c["X"].shape[1]
    trend = {go.TREND_CONSTANT: b2.constant_trend, go.TREND_LINEAR: b2.linear_trend}[int(c["trend"])]
    ok = bool(c["ok"])
    mean = trend(D) if ok else trend(D, beta=np.asarray(c["beta_in"], float).ravel() if np.size(c["beta_in"]) > 1 else float(np.ravel(c["beta_in"])[0]))
    mode = int(c["mode"])
    kw = {go.MODE_NOISELESS: dict(nugget=None), go.MODE_NOISY: dict(nugget=float(c["nugget"])),
          go.MODE_NOISE_ESTIM: dict(nugget=float(c["nugget"]), noise_estim=True)}[mode]
    gp = b2.GaussianProcess(mean=mean, corr=CORR_ARG[int(c["corr"])], thetaL=[1e-5] * D, thetaU=[1e2] * D, **kw)
    llf = gp.fit_fixed(c["X"], c["y"], c["theta"], None if mode == go.MODE_NOISELESS else float(c["par_last"]))
    assert llf == pytest.approx(float(c["llf"]), rel=1e-11)
    mn = bool(c["minimize"])
    Xc = c["Xc"]
    a = dict(rtol=1e-6, atol=1e-300)
    ei = b2.EI(model=gp, minimize=mn)
    assert float(ei.plugin) == pytest.approx(float(c["plugin"]), rel=1e-14)
    np.testing.assert_allclose(ei(Xc), c["ei"], **a)
    np.testing.assert_allclose(b2.MGFI(model=gp, minimize=mn, t=float(c["t"]))(Xc), c["mgfi"], **a)
    np.testing.assert_allclose(b2.MGFI(model=gp, minimize=mn, t=30.0)(Xc), c["mgfi_big_t"], **a)          # t capped at 22.36
    np.testing.assert_allclose(b2.UCB(model=gp, minimize=mn, alpha=float(c["alpha_ucb"]))(Xc), c["ucb"], rtol=1e-12)
    np.testing.assert_allclose(b2.EpsilonPI(model=gp, minimize=mn, epsilon=float(c["eps"]))(Xc), c["epi"], **a)
    bv, bi = b2.MGFI(model=gp, minimize=mn, t=float(c["t"])).argmax(Xc, [float(c["t"]), 30.0])
    assert int(bi[0]) == int(np.argmax(c["mgfi"])) and int(bi[1]) == int(np.argmax(c["mgfi_big_t"]))
    with pytest.raises(AssertionError):
        b2.UCB(model=gp, alpha=-1.0)                                                                       # acquisition_fun.py:124
    # the candidate-set maximiser on top (host glue end to end)
    x, v = b2.argmax_candidates(functools.partial(ei, return_dx=False), [[-1, 2]] * D, n_candidates=2000,
                                rng=np.random.default_rng(0), refine_steps=0 if int(c["corr"]) == go.CORR_MATERN52 else 3)
    assert len(x) == D and v >= ei(np.array([x]))[0] * (1 - 1e-9)


def _sequential_search(gp):
    """the search as a plain sequential loop (what gpr.py:1127-1162 does), written for this test: one restart after the
    other on the model's primary engine, each start point drawn when its turn comes"""
    from scipy.optimize import fmin_l_bfgs_b

    from bayesian_optimization_b200 import hyperopt as ho

    box = ho.parameter_box(gp)
    budget = 200 * box.n if gp.eval_budget is None else gp.eval_budget
    obj = ho.NegLikelihood(gp, gp.engine, gp.likelihood == "restricted")
    z0 = ho.first_start(gp, box)
    best, waited, used = None, 0, 0
    for i in range(gp.random_start):
        if i:
            z0 = np.random.uniform(box.lo, box.hi)
        z, f, info = fmin_l_bfgs_b(obj, z0, bounds=box.bounds, maxfun=budget)
        if best is None:
            best = (z, f)
        elif f <= best[1]:
            best, waited = (z, f), 0
        else:
            waited += 1
        used += info["funcalls"]
        budget -= info["funcalls"]
        if budget <= 0 or waited >= gp.wait_iter:
            break
    return 10.0 ** best[0], -best[1], used


@pytest.mark.parametrize("random_start,eval_budget,wait_iter", [(1, None, 5), (3, None, 5), (7, None, 5), (7, 60, 5),
                                                              (6, 25, 5), (7, None, 1), (5, 200, 2)])
@pytest.mark.parametrize("name", ["rbf_ny", "m32_ny", "rbf_ne"])
def test_concurrent_restarts_equal_the_sequential_loop(name, random_start, eval_budget, wait_iter):
    """the waves of concurrent restarts (hyperopt.py) give the sequential loop's parameters, likelihood, evaluation count
    AND leave numpy's global generator in the same state -- also when the budget binds inside a wave or the stagnation
    counter stops it early"""
    c = FITS[name]
    D, mode, kw = _kwargs(c)
    kw.update(random_start=random_start, eval_budget=eval_budget, wait_iter=wait_iter)
    kw.pop("theta0")  # the first start is drawn too
    a = b2.GaussianProcess(mean=b2.constant_trend(D), **kw)
    b = b2.GaussianProcess(mean=b2.constant_trend(D), **kw)
    np.random.seed(11)
    a.fit(c["X"], c["y"])
    state_a = np.random.get_state()
    np.random.seed(11)
    b._check_data(c["X"], c["y"])
    par, llf, used = _sequential_search(b)
    state_b = np.random.get_state()
    assert a.eval_count == used
    assert a.log_likelihood_ == pytest.approx(llf, rel=1e-12)
    np.testing.assert_allclose(np.r_[a.par["theta"], [a.par[k][0] for k in a.par if k != "theta"]], par, rtol=1e-12)
    assert state_a[0] == state_b[0] and np.array_equal(state_a[1], state_b[1]) and state_a[2:] == state_b[2:]


def test_update_appends_rows_or_refits():
    """update(X, y, reoptimize=False): new rows behind the current training set go through Engine.append (same state as
    a fixed-parameter refit); reordered data falls back to that refit; the default stays upstream's update = fit"""
    c = FITS["rbf_ny"]
    D, mode, kw = _kwargs(c)
    X, y = c["X"], c["y"]
    n0 = X.shape[0] - 7
    a = b2.GaussianProcess(mean=b2.constant_trend(D), **kw)
    np.random.seed(3)
    a.fit(X[:n0], y[:n0])
    theta, last = a.theta_.copy(), a._par_last
    calls = a.engine.n_factor
    assert a.update(X, y, reoptimize=False) is a
    assert a.engine.n_factor == calls + 1 and a.X.shape == X.shape and np.array_equal(a.theta_, theta)
    ref = go.fit_fixed(X, y, int(c["corr"]), theta, mode, sigma2=last, noise_var=1e-2)
    assert a.log_likelihood_ == pytest.approx(ref.llf, rel=1e-12)
    np.testing.assert_allclose(a.predict(c["Xc"]), go.predict_chunked(ref, c["Xc"], 64, eval_MSE=False), rtol=1e-10)
    # reordered rows: not an append, still the fixed-parameter model of the new data
    b = b2.GaussianProcess(mean=b2.constant_trend(D), **kw)
    np.random.seed(3)
    b.fit(X[:n0], y[:n0])
    b.update(X[::-1], y[::-1], reoptimize=False)
    assert b.log_likelihood_ == pytest.approx(ref.llf, rel=1e-10) and np.array_equal(b.theta_, theta)
    # default: upstream's update() re-estimates the hyper-parameters
    np.random.seed(4)
    b.update(X, y)
    assert b.eval_count > 0
