"""CPU: host-side mirror of the reference interface (constructor checks, kernel-name resolution, trend
classes, acquisition parameter handling, pickling) -- nothing here touches the GPU."""
import functools
import pickle

import numpy as np
import pytest

import bayesian_optimization_b200 as b2
from bayesian_optimization_b200 import _lib
from bayesian_optimization_b200.hyperopt import hyperparameter_bounds


def matern(theta, X, nu=1.5):  # stands in for the reference's kernel function object (recognised by name)
    raise AssertionError("never called: kernels run on the device")


def test_constructor_matches_reference_defaults():
    gp = b2.GaussianProcess(thetaL=[1e-3] * 3, thetaU=[1e2] * 3)
    assert gp.estimation_mode == "noisy" and gp.nugget == 1e-6          # gpr.py:219, :258-263
    assert gp.estimate_trend is False and float(gp.mean.beta[0, 0]) == 0  # default = simple kriging, gpr.py:269-270
    assert gp.is_fitted is False
    gp = b2.GaussianProcess(mean=b2.constant_trend(3), thetaL=[1e-3] * 3, thetaU=[1e2] * 3, nugget=None)
    assert gp.estimation_mode == "noiseless" and gp.estimate_trend is True
    gp = b2.GaussianProcess(thetaL=[1e-3], thetaU=[1e2], noise_estim=True)
    assert gp.estimation_mode == "noise_estim"
    with pytest.raises(ValueError, match="finite"):
        b2.GaussianProcess(thetaL=[1e-3], thetaU=[np.inf])                # gpr.py:241-242
    with pytest.raises(AssertionError):
        b2.GaussianProcess(thetaL=[1e-3], thetaU=[1.0], likelihood="nope")


def test_check_params_errors():
    gp = b2.GaussianProcess(thetaL=[1e-3, 1e-3], thetaU=[1.0])
    with pytest.raises(ValueError, match="same length"):
        gp._check_params()
    gp = b2.GaussianProcess(thetaL=[-1.0], thetaU=[1.0])
    with pytest.raises(ValueError, match="bounds"):
        gp._check_params()
    gp = b2.GaussianProcess(thetaL=[1e-3], thetaU=[1.0], optimizer="SGD")
    with pytest.raises(ValueError, match="optimizer"):
        gp._check_params()
    gp = b2.GaussianProcess(thetaL=[1e-3], thetaU=[1.0], corr="laplace")
    with pytest.raises(ValueError, match="corr"):
        gp._check_params()


def test_resolve_corr():
    assert b2.resolve_corr("squared_exponential") == _lib.CORR_RBF
    assert b2.resolve_corr("matern") == _lib.CORR_MATERN32           # string API is nu=1.5
    assert b2.resolve_corr(functools.partial(matern, nu=2.5)) == _lib.CORR_MATERN52
    assert b2.resolve_corr(functools.partial(matern, nu=0.5)) == _lib.CORR_MATERN12
    assert b2.resolve_corr("absolute_exponential") == _lib.CORR_ABSEXP
    assert b2.resolve_corr(functools.partial(matern, nu=3.5)) == _lib.CORR_MATERN_NU   # any other nu: K_nu on the device
    with pytest.raises(ValueError):
        b2.resolve_corr(functools.partial(matern, nu=-1.0))
    with pytest.raises(ValueError):
        b2.resolve_corr(lambda t, d: d)


def test_trend_mirror():
    t = b2.constant_trend(4, beta=0.3)
    assert t.beta.shape == (1, 1) and t.n_dim == 1
    X = np.zeros((5, 4))
    np.testing.assert_array_equal(t.F(X), np.ones((5, 1)))
    np.testing.assert_allclose(t(X), 0.3 * np.ones((5, 1)))
    np.testing.assert_array_equal(t.F(X.T), np.ones((5, 1)))       # trend.py:51-58 transposes silently
    with pytest.raises(Exception, match="beta is not set"):
        b2.constant_trend(4)(X)
    with pytest.raises(Exception, match="right size"):
        t.F(np.zeros((3, 5)))


def test_hyperparameter_bounds():
    gp = b2.GaussianProcess(thetaL=[1e-3, 1e-2], thetaU=[10.0, 20.0])
    gp.y = np.array([[0.0], [2.0], [4.0]])
    bnd = hyperparameter_bounds(gp, ["theta", "sigma2"])
    np.testing.assert_allclose(bnd, [[1e-3, 10.0], [1e-2, 20.0], [1e-5, max(1e-3, gp.y.std() ** 2)]])
    bnd = hyperparameter_bounds(gp, ["theta", "alpha"])
    np.testing.assert_allclose(bnd[-1], [1e-10, 1 - 1e-10])


class FakeModel:
    y = np.array([[3.0], [-2.0], [1.0]])
    is_fitted = True

    def predict(self, X, eval_MSE=False):
        raise AssertionError


def test_acquisition_parameters():
    m = FakeModel()
    assert b2.EI(model=m, minimize=True).plugin == -2.0                 # acquisition_fun.py:96-104
    assert b2.EI(model=m, minimize=False).plugin == -3.0
    assert b2.EI(model=m, minimize=False, plugin=5.0).plugin == -5.0
    assert b2.MGFI(model=m, t=30).t == 22.36                            # :262
    with pytest.raises(AssertionError):
        b2.MGFI(model=m, t=0)
    with pytest.raises(AssertionError):
        b2.UCB(model=m, alpha=0)                                        # :124
    with pytest.raises(AssertionError):
        b2.EpsilonPI(model=m, epsilon=0)                                # :203-206
    assert b2.PI(model=m)._param() == 0.0                               # constructible here
    with pytest.raises(ValueError):
        b2.EI(model=None)
    with pytest.raises(TypeError, match="engine"):
        b2.EI(model=m)(np.zeros((2, 3)))                                # no silent CPU fallback
    np.testing.assert_array_equal(b2.MGFI(model=m)._clean_params([1.0, 50.0]), [1.0, 22.36])


def test_pickle_drops_device_handle():
    gp = b2.GaussianProcess(mean=b2.constant_trend(2), thetaL=[1e-3] * 2, thetaU=[1e2] * 2)
    gp._engine = object()  # pretend
    gp._cache = {"C": 1}
    g2 = pickle.loads(pickle.dumps(gp))
    assert g2._engine is None and g2._cache == {}
    assert g2.estimation_mode == gp.estimation_mode and g2.mean.beta is None


def test_trend_classes_match_oracle_basis():
    """linear / quadratic bases (trend.py:94-142) against the oracle's restatement; (p, k) beta for multi-target fits"""
    from oracle import gp_oracle as go

    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (7, 3))
    np.testing.assert_array_equal(b2.linear_trend(3).F(X), go.trend_basis(go.TREND_LINEAR, X))
    np.testing.assert_array_equal(b2.quadratic_trend(3).F(X), go.trend_basis(go.TREND_QUADRATIC, X))
    assert b2.quadratic_trend(3).n_dim == 10 and b2.linear_trend(3).n_dim == 4
    assert b2.linear_trend(3).Jacobian(X[:1]).shape == (4, 3)
    with pytest.raises(NotImplementedError):
        b2.quadratic_trend(3).Jacobian(X[:1])
    t = b2.linear_trend(3, beta=0.5)
    assert t.beta.shape == (4, 1) and t(X).shape == (7, 1)
    with pytest.raises(Exception, match="Shapes"):
        b2.linear_trend(3, beta=[1.0, 2.0])
    c = b2.constant_trend(3)
    c.beta = np.array([[0.1, 0.2]])            # one column per target (upstream's setter raises here)
    assert c.beta.shape == (1, 2)


def test_kernel_ids_and_restricted_parameter_split():
    assert b2.resolve_corr("generalized_exponential") == _lib.CORR_GENEXP
    gp = b2.GaussianProcess(thetaL=[1e-3] * 2, thetaU=[1e2] * 2, likelihood="restricted")                 # noisy (nugget 1e-6)
    th, s2, nv = gp._split_par_restricted([0.5, 0.6, 0.9])
    assert list(th) == [0.5, 0.6] and s2 == 0.9 and nv == 1e-6                                            # gpr.py:829-830
    gp = b2.GaussianProcess(thetaL=[1e-3] * 2, thetaU=[1e2] * 2, likelihood="restricted", nugget=None)
    assert gp._split_par_restricted([0.5, 0.6, 0.9])[2] == 0.0                                            # gpr.py:826-827
    gp = b2.GaussianProcess(thetaL=[1e-3] * 2, thetaU=[1e2] * 2, likelihood="restricted", noise_estim=True)
    th, s2, nv = gp._split_par_restricted([0.5, 0.6, 0.9, 0.05])
    assert list(th) == [0.5, 0.6] and (s2, nv) == (0.9, 0.05)                                             # gpr.py:832-833
