"""Host-side mirror of the reference's trend (prior mean) classes for the part of them the hot path uses.

Reference: bayes_optim/surrogate/gaussian_process/trend.py.  ``constant_trend`` (:69-91) is what ``fmin``
and every upstream GP test use (bayes_optim/__init__.py:148, unittest/test_BO.py:34); ``beta=None`` means
the coefficient is estimated (ordinary kriging), a number means simple kriging (gpr.py:269-275).
The device evaluates the bases itself (constant: the literal 1 inside the kernels; linear / quadratic:
csrc/trend_kernels.cuh); these classes carry the coefficients and answer the host-side questions (F, Jacobian).
"""
from __future__ import annotations

import numpy as np


class BasisExpansionTrend:
    """trend.py:10-64: ``m(X) = F(X) beta`` with ``beta`` stored as a (p, 1) column."""

    def __init__(self, n_feature: int, n_dim: int, beta=None):
        self.n_feature = int(n_feature)
        self.n_dim = int(n_dim)
        self.beta = beta

    @property
    def beta(self):
        return self._beta

    @beta.setter
    def beta(self, beta):
        # trend.py:21-29: scalars are broadcast to n_dim, everything is reshaped to a column
        if beta is not None:
            if not hasattr(beta, "__iter__"):
                beta = np.array([beta] * self.n_dim)
            b2d = np.atleast_2d(beta)
            if b2d.ndim == 2 and b2d.shape[0] == self.n_dim and b2d.shape[1] > 1 and self.n_dim > 1 or (
                    self.n_dim == 1 and np.ndim(beta) == 2 and np.shape(beta)[0] == 1 and np.shape(beta)[1] > 1):
                # (p, k): one coefficient column per target of a multi-target fit with beta estimated -- upstream
                # flattens this and raises (trend.py:25-28), which is why its ordinary kriging cannot take y (N, k > 1)
                self._beta = np.asarray(b2d, dtype=np.float64)
                return
            beta = b2d.reshape(-1, 1)
            if len(beta) != self.n_dim:
                raise Exception("Shapes of beta and F do not match.")
        self._beta = beta

    def check_input(self, X):
        # trend.py:51-58: a (D, M) input is silently transposed
        X = np.atleast_2d(X)
        if X.shape[1] != self.n_feature:
            X = X.T
        if X.shape[1] != self.n_feature:
            raise Exception("X does not have the right size!")
        return X

    def __call__(self, X):
        if self._beta is None:
            raise Exception("beta is not set!")
        return self.F(X).dot(self._beta)

    def F(self, X):  # pragma: no cover - abstract
        raise NotImplementedError


class constant_trend(BasisExpansionTrend):
    """trend.py:69-91: zero-order polynomial, p = 1, F(x) = 1."""

    def __init__(self, n_feature: int, beta=None):
        super().__init__(n_feature, 1, beta)

    def F(self, X):
        X = self.check_input(X)
        return np.ones((X.shape[0], 1))

    def Jacobian(self, x):
        self.check_input(x)
        return np.zeros((1, self.n_feature))

    def Hessian(self, x):
        self.check_input(x)
        return np.zeros((self.n_feature, self.n_feature, self.n_dim))


class linear_trend(BasisExpansionTrend):
    """trend.py:94-116: first-order polynomial, p = n + 1, f(x) = [1, x_1, ..., x_n]."""

    def __init__(self, n_feature: int, beta=None):
        super().__init__(n_feature, n_feature + 1, beta)

    def F(self, X):
        X = self.check_input(X)
        return np.c_[np.ones(X.shape[0]), X]

    def Jacobian(self, x):
        x = self.check_input(x)
        assert x.shape[0] == 1
        return np.r_[np.zeros((1, self.n_feature)), np.eye(self.n_feature)]

    def Hessian(self, x):
        self.check_input(x)
        return np.zeros((self.n_feature, self.n_feature, self.n_dim))


class quadratic_trend(BasisExpansionTrend):
    """trend.py:119-142: second-order polynomial, f(x) = [1, {x_i}, {x_k x_j, j >= k}], p = (n + 1)(n + 2) / 2."""

    def __init__(self, n_feature: int, beta=None):
        super().__init__(n_feature, (n_feature + 1) * (n_feature + 2) // 2, beta)

    def F(self, X):
        X = self.check_input(X)
        f = np.c_[np.ones(X.shape[0]), X]
        for k in range(self.n_feature):
            f = np.c_[f, X[:, k, np.newaxis] * X[:, k:]]
        return f

    def Jacobian(self, X):
        raise NotImplementedError  # trend.py:138-139

    def Hessian(self, X):
        raise NotImplementedError
