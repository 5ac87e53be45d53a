"""argmax_candidates: the candidate-set replacement of the reference's argmax_restart
(acquisition/optim/__init__.py:55-153).  CPU part: host logic on a stand-in criterion; GPU part: EI on a fitted GP."""
import functools

import numpy as np
import pytest

import bayesian_optimization_b200 as b2
from bayesian_optimization_b200 import candidates as cd


class Quadratic:
    """stand-in criterion with the two methods the maximiser uses: peak 3.0 at x* inside the box"""

    def __init__(self, xstar):
        self.xstar = np.asarray(xstar, float)
        self.calls = 0

    def _param(self):
        return 0.0

    def __call__(self, X, return_dx=False):
        v, g = self.value_and_gradient(np.atleast_2d(X))
        return (v, g) if return_dx else v

    def batch(self, X, params):
        self.calls += 1
        return (3.0 - ((X - self.xstar) ** 2).sum(axis=1))[None, :]

    def value_and_gradient(self, X):
        self.calls += 1
        return 3.0 - ((X - self.xstar) ** 2).sum(axis=1), -2.0 * (X - self.xstar)


class Space:
    def __init__(self, bounds):
        self.bounds = bounds


def test_returns_argmax_restart_shape_and_polishes():
    f = Quadratic([0.3, -0.2, 0.7])
    sp = Space([[-1, 1]] * 3)
    x, v = cd.argmax_candidates(f, sp, n_candidates=4096, rng=np.random.default_rng(0))
    assert isinstance(x, list) and len(x) == 3 and isinstance(v, float)
    raw = cd.argmax_candidates(f, sp, n_candidates=4096, rng=np.random.default_rng(0), refine_steps=0)
    assert v >= raw[1] and v > 3.0 - 1e-4            # the polish only ever improves on the raw candidate arg-max
    np.testing.assert_allclose(x, f.xstar, atol=1e-2)
    assert f.calls < 100                              # a handful of batched passes, not one call per candidate


def test_unwraps_partials_like_base_py():
    f = Quadratic([0.0, 0.0])
    wrapped = functools.partial(functools.partial(f, return_dx=True))
    assert cd.unwrap_criterion(wrapped) is f
    # upstream's partial_argument (utils/utils.py:163-203) decorates its closure with functools.wraps(func)
    inner = functools.partial(f, return_dx=True)

    @functools.wraps(inner)
    def wrapper(X):
        return inner(X)

    assert cd.unwrap_criterion(wrapper) is f
    with pytest.raises(TypeError):
        cd.unwrap_criterion(lambda x: 0.0)


def test_constraints_duplicates_and_empty_result():
    f = Quadratic([0.5, 0.5])
    sp = Space([[0, 1]] * 2)
    rng = np.random.default_rng(1)
    # inequality g(x) <= 0 cuts the optimum off: x0 <= 0.25
    x, v = cd.argmax_candidates(f, sp, g=lambda x: [x[0] - 0.25], n_candidates=2048, rng=rng)
    assert x[0] <= 0.25 and v < 3.0
    # everything infeasible -> ([], []) as acquisition/optim/__init__.py:146-147
    assert cd.argmax_candidates(f, sp, g=lambda x: [1.0], n_candidates=256, rng=rng) == ([], [])
    # the best point is already evaluated: pre_eval_check semantics drop it (bayes_opt.py:42-50)
    x1, _ = cd.argmax_candidates(f, sp, n_candidates=1024, rng=np.random.default_rng(2), refine_steps=0)
    x2, _ = cd.argmax_candidates(f, sp, n_candidates=1024, rng=np.random.default_rng(2), refine_steps=0, data=np.array([x1]))
    assert x2 != x1
    with pytest.raises(ValueError):
        cd.argmax_candidates(f, Space([[0, np.inf]] * 2))


def test_sampling_uses_the_space_when_it_can():
    class S(Space):
        def sample(self, N, method):
            assert method == "uniform"
            return [[0.25, 0.75]] * N

    X = cd.sample_candidates(S([[0, 1]] * 2), 5)
    assert X.shape == (5, 2) and X.dtype == np.float64 and X.flags.c_contiguous


@pytest.mark.gpu
def test_ei_on_device_matches_oracle():
    from bayesian_optimization_b200 import workloads
    from oracle import gp_oracle as go

    N, D = 256, 4
    X, y, theta = workloads.canonical_problem(N, D)
    gp = b2.GaussianProcess(mean=b2.constant_trend(D), corr="squared_exponential", thetaL=[1e-5] * D, thetaU=[1e2] * D, nugget=1e-6)
    gp.fit_fixed(X, y, theta, 1.0)
    ora = go.fit_fixed(X, y, go.CORR_RBF, theta, go.MODE_NOISY, sigma2=1.0, noise_var=1e-6)
    ei = b2.EI(model=gp, minimize=True)
    rng = np.random.default_rng(3)
    x, v = b2.argmax_candidates(functools.partial(ei, return_dx=True), Space([[0, 1]] * D), n_candidates=20000, rng=rng, data=X)
    assert len(x) == D and all(0.0 <= t <= 1.0 for t in x)
    yo, mo = go.predict(ora, np.array([x]))
    pl = go.plugin_value(ora.y, True)
    assert v == pytest.approx(float(go.ei(yo, mo, ora.sigma2, pl)[0]), rel=1e-6)
    # at least as good as the raw candidate set scored by the oracle
    Xc = cd.sample_candidates(Space([[0, 1]] * D), 20000, np.random.default_rng(3))
    yo, mo = go.predict_chunked(ora, Xc, 2048)
    assert v >= go.ei(yo, mo, ora.sigma2, pl).max() * (1 - 1e-9)


@pytest.mark.gpu
def test_fast_precision_ranking_matches_float64_choice():
    """precision="fast": candidates ranked by the tensor-core pass, picks re-scored exactly -- the raw (unpolished)
    answer must be the float64 path's arg-max"""
    from bayesian_optimization_b200 import workloads

    N, D = 600, 6
    X, y, theta = workloads.canonical_problem(N, D)
    res = {}
    for prec in ("fp64", "fast"):
        gp = b2.GaussianProcess(mean=b2.constant_trend(D), corr="matern52", thetaL=[1e-5] * D, thetaU=[1e2] * D, nugget=1e-6,
                                precision=prec)
        gp.fit_fixed(X, y, theta, 1.0)
        res[prec] = b2.argmax_candidates(b2.MGFI(model=gp, t=2.0), Space([[0, 1]] * D), n_candidates=50000,
                                         rng=np.random.default_rng(4), refine_steps=0)
    assert res["fast"][0] == res["fp64"][0]
    assert res["fast"][1] == pytest.approx(res["fp64"][1], rel=1e-9)


@pytest.mark.parametrize("method", ["uniform", "LHS", "sobol"])
def test_sampling_designs(method):
    """the three designs of SearchSpace._sample (search_space.py:742-754) on plain bounds"""
    from bayesian_optimization_b200.candidates import sample_candidates

    b = np.array([[-5.0, 5.0], [0.0, 1.0], [10.0, 30.0]])
    M = 1024
    X = sample_candidates(b, M, rng=np.random.default_rng(3), method=method)
    assert X.shape == (M, 3) and X.dtype == np.float64
    assert np.all(X >= b[:, 0]) and np.all(X <= b[:, 1])
    U = (X - b[:, 0]) / (b[:, 1] - b[:, 0])
    if method == "LHS":      # exactly one point per stratum and coordinate
        for d in range(3):
            assert np.array_equal(np.sort(np.floor(U[:, d] * M).astype(int)), np.arange(M))
    if method == "sobol":    # a (t, m, s)-net with the origin skipped: points 1 .. M-1 fill every dyadic slab but the first
        for d in range(3):
            assert np.array_equal(np.sort(np.floor(U[: M - 1, d] * M).astype(int)), np.arange(1, M))
        assert U[0].tolist() == [0.5, 0.5, 0.5]
    assert sample_candidates(b, 1, rng=np.random.default_rng(0), method="LHS").shape == (1, 3)
    with pytest.raises(ValueError):
        sample_candidates(b, 4, method="halton")


def test_sampling_goes_through_the_space_when_it_can():
    """a search space with its own ``sample`` (upstream's SearchSpace) is asked, with the method passed on"""
    from bayesian_optimization_b200.candidates import sample_candidates

    class Space:
        bounds = [[0.0, 1.0], [0.0, 2.0]]

        def sample(self, N, method):
            self.asked = (N, method)
            return [[0.5, 1.0]] * N

    sp = Space()
    X = sample_candidates(sp, 5, method="sobol")
    assert sp.asked == (5, "sobol") and X.shape == (5, 2)
