"""Candidate-set maximiser of an acquisition function with the calling convention of the reference's
``argmax_restart`` (bayes_optim/acquisition/optim/__init__.py:55-153): same positional arguments, same return
value ``(xopt: list, fopt: float)`` (``([], [])`` when nothing feasible was found, :146-147).

The reference restarts a single-point L-BFGS-B / CMA-ES / MIES search and pays one ``model.predict`` per function
evaluation.  Here the M candidates are scored in ONE device pass (``AcquisitionFunction.batch``), the best K are
polished together by a few steps of projected gradient ascent on the device gradients
(``AcquisitionFunction.value_and_gradient`` = the return_dx path, acquisition_fun.py:139-309), candidates that
duplicate evaluated points are dropped the way ``BO.pre_eval_check`` does (bayes_opt.py:27-55, ``np.isclose`` on
every coordinate), and the constraints are applied as upstream (:127-128).

Use it wherever ``argmax_restart`` is bound, e.g. ``bo._argmax_restart = functools.partial(argmax_candidates,
search_space=bo.search_space, ...)``.
"""
from __future__ import annotations

import functools
import logging
from typing import Callable, Optional, Sequence, Tuple

import numpy as np

__all__ = ["argmax_candidates", "sample_candidates", "unwrap_criterion"]


def unwrap_criterion(obj_func):
    """The AcquisitionFunction behind ``obj_func``: the object itself, or the ``func`` of (nested) functools.partial
    wrappers (base.py:489-494 wraps ``functools.partial(criterion, return_dx=...)``)."""
    f = obj_func
    for _ in range(8):
        if hasattr(f, "value_and_gradient") and hasattr(f, "batch"):
            return f
        if isinstance(f, functools.partial):
            f = f.func
            continue
        inner = getattr(f, "criterion", None) or getattr(f, "__wrapped__", None)
        if inner is None:
            break
        f = inner
    raise TypeError("obj_func is not (a wrapper of) a bayesian_optimization_b200 acquisition function")


def _bounds_of(search_space) -> np.ndarray:
    b = getattr(search_space, "bounds", search_space)
    b = np.asarray(b, dtype=np.float64)
    if b.ndim != 2 or b.shape[1] != 2:
        raise ValueError("search_space must expose bounds of shape (D, 2)")
    if not np.all(np.isfinite(b)) or np.any(b[:, 0] > b[:, 1]):
        raise ValueError("bounds must be finite with lower <= upper")
    return b


SAMPLING_METHODS = ("uniform", "LHS", "sobol")


def sample_candidates(search_space, M: int, rng: Optional[np.random.Generator] = None, method: str = "uniform") -> np.ndarray:
    """(M, D) float64 candidates by one of the designs of ``SearchSpace._sample`` (search_space.py:742-754):
    "uniform", "LHS" (Latin hypercube) or "sobol".  When the space has a ``sample`` method and no generator is
    forced, that method is called (``search_space.sample(N, method=...)``: upstream's own code, pyDOE / sobol_seq
    included); on plain bounds the designs are built here:
      * LHS: one point per stratum and coordinate, strata permuted independently per coordinate.  (Upstream asks
        pyDOE for its "maximin" variant -- the best of five such designs by smallest pairwise distance, an O(M^2)
        criterion that is out of reach for 1e6 candidates; the stratification is what matters for a scoring pass.)
      * sobol: the unscrambled Sobol sequence after the origin, as ``i4_sobol_generate(dim, N)`` (skip = 1) returns it,
        from scipy's generator (Joe-Kuo direction numbers, up to 21201 dimensions)."""
    if method not in SAMPLING_METHODS:
        raise ValueError("method should be one of %s, %s was given." % (list(SAMPLING_METHODS), method))
    M = int(M)
    if rng is None and hasattr(search_space, "sample"):
        return np.ascontiguousarray(np.asarray(search_space.sample(N=M, method=method), dtype=np.float64))
    b = _bounds_of(search_space)
    D = b.shape[0]
    rng = np.random.default_rng() if rng is None else rng
    if method == "uniform" or (method == "LHS" and M == 1):  # search_space.py:748-749: one LHS point is a uniform draw
        U = rng.random((M, D))
    elif method == "LHS":
        U = np.empty((M, D))
        for d in range(D):
            U[:, d] = (rng.permutation(M) + rng.random(M)) / M
    else:
        from scipy.stats import qmc

        s = qmc.Sobol(d=D, scramble=False)
        s.fast_forward(1)
        import warnings

        with warnings.catch_warnings():  # scipy warns when M is not a power of two; any prefix of the sequence is wanted here
            warnings.simplefilter("ignore")
            U = s.random(M)
    return b[:, 0] + (b[:, 1] - b[:, 0]) * U


def _is_duplicate(x: np.ndarray, data: Optional[np.ndarray]) -> bool:
    # bayes_opt.py:42-50: a candidate equal (np.isclose on every coordinate) to an evaluated point is dropped
    return data is not None and data.size > 0 and bool(np.any(np.all(np.isclose(data, x), axis=1)))


def _feasible(x: Sequence[float], h: Optional[Callable], g: Optional[Callable]) -> bool:
    # acquisition/optim/__init__.py:127-128
    cond_h = all(np.isclose(np.abs(np.atleast_1d(h(x))), 0, atol=1e-1)) if h else True
    cond_g = all(np.atleast_1d(g(x)) <= 0) if g else True
    return bool(cond_h and cond_g)


def rank_values(crit, Xc: np.ndarray) -> Tuple[np.ndarray, bool]:
    """(values (M,), exact?) used to RANK the candidates.  A model built with precision="fast" is scored by the fused
    tensor-core pass (posterior moments to ~1e-3, criterion evaluated on the device from them); whatever is picked
    from that ranking is re-scored exactly afterwards, and the exact arg-max of the pass joins the picks.  Any other
    model takes the float64 path, whose values are the reference's to 1e-7."""
    model = getattr(crit, "_model", None)
    eng = getattr(model, "engine", None) if getattr(model, "precision", "fp64") == "fast" else None
    if eng is not None and hasattr(eng, "acq_from_moments") and getattr(model, "_sub", None) is None:
        yh, ms = eng.predict(Xc, True)
        _, _, vals = eng.acq_from_moments(yh, ms, crit._acq_id, crit.minimize, crit._plugin_value(), [crit._param()], True)
        return np.asarray(vals)[0], False
    return np.asarray(crit.batch(Xc, [crit._param()]))[0], True


def refine(criterion, X0: np.ndarray, bounds: np.ndarray, steps: int = 20, step0: float = 0.05) -> Tuple[np.ndarray, np.ndarray]:
    """Projected gradient ascent on K points at once; every iteration is one device call.  Per-point step sizes
    (relative to the box), doubled after an accepted step and halved after a rejected one.  Returns the best
    evaluated (X (K, D), values (K,)); never worse than the start."""
    X = np.array(X0, dtype=np.float64, copy=True)
    width = bounds[:, 1] - bounds[:, 0]
    val, dx = criterion.value_and_gradient(X)
    val = np.where(np.isfinite(val), val, -np.inf)
    step = np.full(X.shape[0], step0)
    for _ in range(int(steps)):
        g = np.where(np.isfinite(dx), dx, 0.0) * width          # gradient in box-relative coordinates
        nrm = np.linalg.norm(g, axis=1)
        live = nrm > 0
        if not live.any():
            break
        d = np.zeros_like(g)
        d[live] = g[live] / nrm[live, None]
        Xn = np.clip(X + (step[:, None] * d) * width, bounds[:, 0], bounds[:, 1])
        vn, dxn = criterion.value_and_gradient(Xn)
        better = np.isfinite(vn) & (vn > val)
        X[better], val[better], dx[better] = Xn[better], vn[better], dxn[better]
        step = np.where(better, step * 2.0, step * 0.5)
        if step.max() < 1e-9:
            break
    return X, val


def argmax_candidates(
    obj_func: Callable,
    search_space,
    h: Callable = None,
    g: Callable = None,
    eval_budget: int = 100,
    n_restart: int = 10,
    wait_iter: int = 3,
    optimizer: str = "B200_candidates",
    logger: logging.Logger = None,
    n_candidates: Optional[int] = None,
    refine_top: int = 64,
    refine_steps: int = 20,
    data: Optional[np.ndarray] = None,
    rng: Optional[np.random.Generator] = None,
    max_constraint_checks: int = 32768,
    sampling: str = "uniform",
):
    """Drop-in for ``argmax_restart``.  ``eval_budget`` x ``n_restart`` (the reference's total number of single-point
    evaluations) scales the default candidate count: max(2^16, 1024 x eval_budget x n_restart), capped at 2^22.
    ``data``: evaluated points (N, D) to de-duplicate against; ``wait_iter`` is accepted for signature
    compatibility (there are no sequential restarts to stop early); ``sampling``: "uniform" | "LHS" | "sobol", the
    designs of ``SearchSpace._sample`` (search_space.py:742-754)."""
    crit = unwrap_criterion(obj_func)
    bounds = _bounds_of(search_space)
    M = int(n_candidates) if n_candidates else int(min(2**22, max(2**16, 1024 * int(eval_budget) * int(n_restart))))
    Xc = sample_candidates(search_space, M, rng, sampling)
    if Xc.ndim != 2 or Xc.shape[1] != bounds.shape[0]:
        raise ValueError("sampled candidates do not match the bounds")
    vals, exact = rank_values(crit, Xc)
    vals = np.where(np.isfinite(vals), vals, -np.inf)
    K = int(min(max(refine_top, 1), M))
    if h is None and g is None:
        top = np.argpartition(-vals, K - 1)[:K]
        top = top[np.argsort(-vals[top], kind="stable")]
    else:
        # constrained: the K best FEASIBLE candidates (the user's h / g are host callables: bounded number of checks)
        top = []
        for n_checked, i in enumerate(np.argsort(-vals, kind="stable")):
            if len(top) >= K or n_checked >= max_constraint_checks or not np.isfinite(vals[i]):
                break
            if _feasible(Xc[i].tolist(), h, g):
                top.append(i)
        if not top:
            return [], []
        top = np.asarray(top)
    if not exact:
        # tensor-core ranking: add the exact arg-max of the pass (band re-scored in float64) and re-score the picks
        _, bi = crit.argmax(Xc)
        if h is None and g is None and int(bi[0]) not in set(int(t) for t in top):
            top = np.concatenate([[int(bi[0])], top])
        Xk = Xc[top]
        vk = np.asarray(crit.batch(Xk, [crit._param()]))[0]
        vk = np.where(np.isfinite(vk), vk, -np.inf)
    else:
        Xk, vk = Xc[top], vals[top]
    if refine_steps > 0:
        Xr, vr = refine(crit, Xk, bounds, refine_steps)
        Xk, vk = np.vstack([Xr, Xk]), np.concatenate([vr, vk])
    order = np.argsort(-vk, kind="stable")
    for i in order:
        x = Xk[i]
        if not np.isfinite(vk[i]) or _is_duplicate(x, data):
            continue
        xl = x.tolist()
        if _feasible(xl, h, g):
            if logger is not None:
                logger.debug("B200 candidates : %d - refined : %d - Fopt : %f" % (M, K, vk[i]))
            return xl, float(vk[i])
    return [], []
