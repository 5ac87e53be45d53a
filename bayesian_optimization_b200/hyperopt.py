"""Maximum-likelihood search for the GP hyper-parameters, organised around the device.

What it must reproduce (bayes_optim/surrogate/gaussian_process/gpr.py:1042-1197): the parameter
vector per estimation mode, the search in log10 space inside the reference's bounds, the warm start from a previous
``theta_``, random restarts drawn from numpy's GLOBAL generator in the reference's order, the acceptance rule
(``<=`` on the negated likelihood), the stagnation counter ``wait_iter``, the evaluation budget that shrinks by every
restart's ``funcalls``, and quirk g4: L-BFGS-B is handed the gradient w.r.t. the RAW parameters although it moves in
their log10 (gpr.py:1113-1121) -- the optima the reference finds depend on it, so it is kept.

How it is organised here.  A restart is a pure function of (start point, evaluation cap): L-BFGS-B is deterministic and
nothing else consumes random numbers.  So the restarts of a wave are drawn up front and run CONCURRENTLY, each on its
own engine handle (its own CUDA stream and factorisation buffers; the likelihood at N <= 2048 is a latency-bound chain
of small kernels, several of them overlap on one GPU), under the budget that was open when the wave started.  The
reference's sequential bookkeeping is then replayed over the wave in order: a restart that would have met a smaller
cap than it actually used is re-run under that cap, results after the stopping point are discarded, and the global
generator is rewound and advanced by exactly the draws the sequential loop would have made -- the outcome (parameters,
likelihood, ``eval_count``, generator state) is the sequential one.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import numpy as np
from scipy.optimize import fmin_l_bfgs_b

WAVE = 4  # restarts in flight (engine handles alive at once)


@dataclass(frozen=True)
class ParameterBox:
    """the optimisation variables: blocks of the parameter vector and their log10 box"""

    names: Tuple[str, ...]
    sizes: Tuple[int, ...]
    lo: np.ndarray  # log10 lower bounds, concatenated
    hi: np.ndarray

    @property
    def n(self) -> int:
        return int(sum(self.sizes))

    @property
    def bounds(self) -> np.ndarray:
        return np.c_[self.lo, self.hi]

    def unpack(self, raw: np.ndarray) -> Dict[str, np.ndarray]:
        out, at = {}, 0
        for name, k in zip(self.names, self.sizes):
            out[name] = raw[at:at + k]
            at += k
        return out


def hyperparameter_bounds(gp, names: Sequence[str]) -> np.ndarray:
    """(n_par, 2) raw bounds: theta from the model; sigma2 in [1e-5, max(1e-3, var y)]; alpha / noise_var in
    [1e-10, 1 - 1e-10]                                                                      (gpr.py:1042-1056)"""
    rows = []
    for name in names:
        if name == "theta":
            rows.append(np.c_[gp.thetaL, gp.thetaU])
        elif name == "sigma2":
            rows.append([[1e-5, max(1e-3, float(gp.y.std()) ** 2)]])
        elif name in ("alpha", "noise_var"):
            rows.append([[1e-10, 1.0 - 1e-10]])
        else:
            raise KeyError(name)
    return np.vstack(rows).astype(np.float64)


def parameter_box(gp) -> ParameterBox:
    """which parameters the likelihood takes: theta; + sigma2 for the restricted likelihood and for a fixed nugget;
    + alpha (concentrated) or noise_var (restricted) when the noise is estimated             (gpr.py:1066-1086)"""
    restricted = gp.likelihood == "restricted"
    names, sizes = ["theta"], [len(gp.thetaL)]
    if restricted or gp.estimation_mode == "noisy":
        names.append("sigma2")
        sizes.append(1)
    if gp.estimation_mode == "noise_estim":
        names.append("noise_var" if restricted else "alpha")
        sizes.append(1)
    b = np.log10(hyperparameter_bounds(gp, names))
    return ParameterBox(tuple(names), tuple(sizes), b[:, 0].copy(), b[:, 1].copy())


def first_start(gp, box: ParameterBox) -> np.ndarray:
    """start of restart 0 in log10 space: the previous optimum, else theta0, else a uniform draw; the remaining
    parameters are always drawn (gpr.py:1093-1107).  Draws come from numpy's global generator, theta first."""
    k = box.sizes[0]
    if hasattr(gp, "theta_"):
        z = np.log10(np.asarray(gp.theta_, dtype=np.float64))
    elif gp.theta0 is not None:
        z = np.log10(np.asarray(gp.theta0, dtype=np.float64))
    else:
        z = np.random.uniform(box.lo[:k], box.hi[:k])
    if box.n > k:
        z = np.r_[z, np.random.uniform(box.lo[k:], box.hi[k:])]
    return z


class NegLikelihood:
    """objective of one restart on one engine handle: z = log10(parameters) -> (-llf, -d llf / d parameters)"""

    def __init__(self, gp, engine, restricted: bool):
        self.gp, self.engine, self.restricted = gp, engine, restricted
        self.calls = 0

    def __call__(self, z):
        self.calls += 1
        raw = np.power(10.0, np.asarray(z, dtype=np.float64))
        llf, grad = self.gp._likelihood_on(self.engine, raw, restricted=self.restricted, eval_grad=True)
        return -llf, -np.asarray(grad, dtype=np.float64).ravel()  # quirk g4: no chain rule for the log10 map


@dataclass
class Restart:
    z0: np.ndarray
    cap: int                 # maxfun this result was obtained under
    z: np.ndarray = None
    neg_llf: float = np.inf
    funcalls: int = 0


def _run_restart(objective: NegLikelihood, r: Restart, box: ParameterBox) -> Restart:
    z, f, info = fmin_l_bfgs_b(objective, r.z0, bounds=box.bounds, maxfun=r.cap)
    r.z, r.neg_llf, r.funcalls = z, float(f), int(info["funcalls"])
    return r


def optimize_hyperparameter(gp):
    """-> ({name: values}, log-likelihood at the optimum, env of that evaluation); leaves the model's device state AT
    the optimum and ``gp.eval_count`` at the number of likelihood evaluations the sequential loop would have made."""
    if gp.optimizer != "BFGS":
        raise NotImplementedError('optimizer="CMA" hangs in the reference on Python 3 and is not provided')
    restricted = gp.likelihood == "restricted"
    box = parameter_box(gp)
    budget = 200 * box.n if gp.eval_budget is None else int(gp.eval_budget)  # gpr.py:1110
    n_restarts = max(1, int(gp.random_start))
    z_first = first_start(gp, box)

    engines = gp._engine_pool(min(WAVE, n_restarts))
    objectives = [NegLikelihood(gp, e, restricted) for e in engines]

    best: Restart = None
    waited, used, done = 0, 0, False
    gp.eval_count = 0
    index = 0
    while not done and index < n_restarts:
        # ---- draw the wave's start points in the sequential order, remembering where the generator stood ---------
        wave_n = min(len(engines), n_restarts - index)
        rng_state = np.random.get_state()
        wave: List[Restart] = []
        for k in range(wave_n):
            z0 = z_first if index + k == 0 else np.random.uniform(box.lo, box.hi)
            wave.append(Restart(z0=z0, cap=budget))
        # ---- run it: one thread per engine handle (ctypes releases the GIL inside the library) --------------------
        if wave_n == 1:
            _run_restart(objectives[0], wave[0], box)
        else:
            with ThreadPoolExecutor(max_workers=wave_n) as ex:
                list(ex.map(lambda a: _run_restart(*a), [(objectives[k], wave[k], box) for k in range(wave_n)]))
        # ---- replay the sequential bookkeeping over the wave ---------------------------------------------------------
        consumed = 0
        for k, r in enumerate(wave):
            if r.funcalls >= budget and r.cap != budget:
                # sequentially this restart would have run under a smaller cap than it was given: redo it under that one
                r = _run_restart(objectives[0], Restart(z0=r.z0, cap=budget), box)
            consumed += 1
            if best is None:
                best = r
            elif r.neg_llf <= best.neg_llf:   # ties move to the later restart, as upstream
                best, waited = r, 0
            else:
                waited += 1
            used += r.funcalls
            budget -= r.funcalls
            if gp.verbose:
                print(f"restart {index + k + 1}: {r.funcalls} likelihood evaluations, best so far {-best.neg_llf:.10g}")
            if budget <= 0 or waited >= gp.wait_iter:
                done = True
                break
        # ---- leave the generator where the sequential loop would have left it ---------------------------------------
        if consumed < wave_n:
            np.random.set_state(rng_state)
            for k in range(consumed):
                if index + k > 0:
                    np.random.uniform(box.lo, box.hi)
        index += consumed
    gp.eval_count = used

    raw = np.power(10.0, best.z)
    env: dict = {}
    llf = gp._likelihood_on(gp.engine, raw, restricted=restricted, env=env)  # the primary handle ends up at the optimum
    return box.unpack(raw), llf, env
