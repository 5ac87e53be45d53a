"""Acquisition functions EI / EpsilonPI / PI / UCB / MGFI on the B200 engine.

Mirror of bayes_optim/acquisition/acquisition_fun.py: same constructor keywords (``model, minimize,
plugin | alpha | epsilon | t``) and ``__call__(X)``; but ``X`` may hold any number of rows.  The reference
evaluates one row at a time (its EI / MGFI / EpsilonPI raise on >1 row); here row i of the
result is what the reference returns for row i alone, computed for all rows -- and for a whole list of
parameter values (``ParallelBO``'s q sampled ``t`` / ``alpha``, bayes_opt.py:82-98) -- in one device pass.

``argmax(X)`` returns only the best value and its lowest index per criterion (numpy argmax rule): the
candidate-set stand-in for ``argmax_restart`` (acquisition/optim/__init__.py:55-153).
"""
from __future__ import annotations

from abc import ABC
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib


class AcquisitionFunction(ABC):
    """acquisition_fun.py:22-84."""

    _acq_id: int = -1

    def __init__(self, model=None, minimize: bool = True):
        self.model = model
        self.minimize = minimize

    @property
    def model(self):
        return self._model

    @model.setter
    def model(self, model):
        if model is None:
            raise ValueError("model cannot be None")
        self._model = model
        assert hasattr(self._model, "predict")

    def check_X(self, X) -> np.ndarray:
        # acquisition_fun.py:82-84; partial_argument hands in object arrays (utils/utils.py:186-195)
        return np.atleast_2d(np.asarray(X, dtype=np.float64))

    # parameters of the criterion; subclasses override
    def _param(self) -> float:
        return 0.0

    def _plugin_value(self) -> float:
        return 0.0

    def _engine(self):
        if getattr(self._model, "_sub", None):
            # a k > 1 model holds one device model per target; a scalar criterion over it would silently score target 0
            raise NotImplementedError("acquisition functions take a single-target model (y of shape (N,) or (N, 1))")
        eng = getattr(self._model, "engine", None)
        if eng is None:
            raise TypeError("the model has no B200 engine; use bayesian_optimization_b200.GaussianProcess")
        return eng

    def _run(self, X, params, return_values):
        X = self.check_X(X)
        if not getattr(self._model, "is_fitted", False):
            raise RuntimeError("the model is not fitted")
        return self._engine().acq(X, self._acq_id, self.minimize, self._plugin_value(), params,
                                  return_values=return_values)

    def __call__(self, X, return_dx: bool = False):
        if return_dx:
            # the reference's return_dx path takes one point (model.gradient raises on more, gpr.py:547-548) and
            # returns (value, dx (1, D)); more rows give (values (M,), dx (M, D)) from one device pass
            X = self.check_X(X)
            val, dx = self.value_and_gradient(X)
            return (val[0], dx) if X.shape[0] == 1 else (val, dx)
        _, _, vals = self._run(X, [self._param()], True)
        return vals[0]

    def value_and_gradient(self, X) -> Tuple[np.ndarray, np.ndarray]:
        """(values (M,), gradients (M, D)): acquisition_fun.py return_dx=True for every row of X."""
        X = self.check_X(X)
        if not getattr(self._model, "is_fitted", False):
            raise RuntimeError("the model is not fitted")
        return self._engine().acq_grad(X, self._acq_id, self.minimize, self._plugin_value(), self._param())

    def batch(self, X, params: Sequence[float]) -> np.ndarray:
        """Values for q parameter settings from ONE predict pass -> (q, M)."""
        _, _, vals = self._run(X, self._clean_params(params), True)
        return vals

    def argmax(self, X, params: Optional[Sequence[float]] = None) -> Tuple[np.ndarray, np.ndarray]:
        """(best values (q,), lowest arg-max indices (q,)) without materialising the (q, M) values."""
        p = [self._param()] if params is None else self._clean_params(params)
        bv, bi, _ = self._run(X, p, False)
        return bv, bi

    def _clean_params(self, params):
        return np.asarray(params, dtype=np.float64).ravel()


class ImprovementBased(AcquisitionFunction):
    """acquisition_fun.py:87-104."""

    def __init__(self, plugin: float = None, **kwargs):
        super().__init__(**kwargs)
        self.plugin = plugin

    @property
    def plugin(self):
        return self._plugin

    @plugin.setter
    def plugin(self, plugin):
        if plugin is None:
            if hasattr(self._model, "y"):
                self._plugin = np.min(self._model.y) if self.minimize else -1.0 * np.max(self._model.y)
            else:
                self._plugin = None
        else:
            self._plugin = plugin if self.minimize else -1.0 * plugin

    def _plugin_value(self) -> float:
        if self._plugin is None:
            raise ValueError("plugin is not set and the model has no training targets")
        return float(self._plugin)


class UCB(AcquisitionFunction):
    """Upper confidence bound ``m(x) + alpha * s(x)`` (acquisition_fun.py:107-147)."""

    _acq_id = _lib.ACQ_UCB

    def __init__(self, alpha: float = 0.5, **kwargs):
        super().__init__(**kwargs)
        self.alpha = alpha

    @property
    def alpha(self):
        return self._alpha

    @alpha.setter
    def alpha(self, alpha):
        assert alpha > 0
        self._alpha = alpha

    def _param(self):
        return float(self._alpha)

    def _clean_params(self, params):
        p = super()._clean_params(params)
        assert np.all(p > 0)
        return p


class EI(ImprovementBased):
    """Expected improvement (acquisition_fun.py:150-189)."""

    _acq_id = _lib.ACQ_EI


class EpsilonPI(ImprovementBased):
    """epsilon-probability of improvement (acquisition_fun.py:192-228)."""

    _acq_id = _lib.ACQ_PI

    def __init__(self, epsilon=1e-10, **kwargs):
        super().__init__(**kwargs)
        self.epsilon = epsilon

    @property
    def epsilon(self):
        return self._epsilon

    @epsilon.setter
    def epsilon(self, eps):
        assert eps > 0
        self._epsilon = eps

    def _param(self):
        return float(self._epsilon)


class PI(EpsilonPI):
    """Probability of improvement = EpsilonPI with epsilon = 0.  Upstream this class cannot be constructed
    (its __init__ forces epsilon = 0 and the setter asserts eps > 0, acquisition_fun.py:231-235, :203-206);
    here the intended function Phi((plugin - yhat)/s) is provided."""

    def __init__(self, **kwargs):
        kwargs.pop("epsilon", None)
        ImprovementBased.__init__(self, **kwargs)
        self._epsilon = 0.0


class MGFI(ImprovementBased):
    """Moment-generating function of the improvement (acquisition_fun.py:238-310)."""

    _acq_id = _lib.ACQ_MGFI

    def __init__(self, t: float = 1, **kwargs):
        super().__init__(**kwargs)
        self.t = t

    @property
    def t(self):
        return self._t

    @t.setter
    def t(self, t):
        assert t > 0
        self._t = min(t, 22.36)  # acquisition_fun.py:262

    def _param(self):
        return float(self._t)

    def _clean_params(self, params):
        p = super()._clean_params(params)
        assert np.all(p > 0)
        return np.minimum(p, 22.36)
