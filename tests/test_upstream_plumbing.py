"""CPU: the UNMODIFIED upstream drivers (``bayes_optim.BO`` / ``ParallelBO`` from /root/reference) running on this repo's
``GaussianProcess`` and acquisition classes -- BASELINE.json config 1 ("fmin sphere 2D ... plumbing only").

What is exercised:
  * ``BaseBO.update_model`` hands ``Solution`` arrays and a (N, 1) standardised fitness to ``model.fit`` and calls
    ``model.predict(data)`` (base.py:441-442);
  * upstream's own EI / MGFI drive ``model.predict(x, eval_MSE=True)`` / ``model.gradient(x)`` one point at a time
    through ``argmax_restart`` (acquisition/optim/__init__.py:55-153);
  * ``_create_acquisition`` looks the class up by name in ``bayes_optim.acquisition.acquisition_fun`` (base.py:482-494):
    swapping this repo's classes in is a ``setattr``;
  * ``argmax_candidates`` bound where ``argmax_restart`` is bound (base.py:231-243) returns the ``(xopt, fopt)`` shape
    ``arg_max_acquisition`` expects;
  * ``BaseBO.save`` / ``load`` pickle the model with dill (base.py:499-540).

The device is replaced by the oracle-backed stand-in of tests/fake_engine.py (test infrastructure); upstream needs three
small pure-Python packages that are not installed here: tests/shims/ provides the one function of each it calls.
Skipped when /root/reference is absent (the GPU box)."""
import functools
import os
import sys

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "bayes_optim")), reason="upstream sources not present")

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def upstream():
    added = [p for p in (os.path.join(HERE, "shims"), REF) if p not in sys.path]
    for p in added:
        sys.path.insert(0, p)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import bayes_optim
    yield bayes_optim
    for p in added:
        sys.path.remove(p)


@pytest.fixture(autouse=True)
def oracle_backed_engine(monkeypatch):
    from bayesian_optimization_b200 import gp as gp_module
    from fake_engine import FakeEngine

    monkeypatch.setattr(gp_module, "Engine", FakeEngine)


def sphere(x):
    return float(np.sum(np.asarray(x, dtype=float) ** 2))


def make_model(dim):
    import bayesian_optimization_b200 as b2

    # the keywords fmin() builds for its GaussianProcess (bayes_optim/__init__.py:147-160)
    return b2.GaussianProcess(mean=b2.constant_trend(dim), corr="matern", thetaL=1e-10 * np.ones(dim),
                              thetaU=10 * np.ones(dim), nugget=1e-6, noise_estim=False, optimizer="BFGS", wait_iter=3,
                              random_start=dim, likelihood="concentrated", eval_budget=100 * dim)


def test_bo_runs_on_the_b200_model(upstream):
    dim = 2
    space = upstream.RealSpace([-5, 5]) * dim
    model = make_model(dim)
    np.random.seed(42)
    opt = upstream.BO(search_space=space, obj_fun=sphere, model=model, DoE_size=5, max_FEs=9, verbose=False, n_point=1,
                      acquisition_optimization={"max_FEs": 40, "n_restart": 2})
    xopt, fopt, _ = opt.run()
    assert model.is_fitted and model.X.shape == (9, dim) and model.X.dtype == np.float64   # Solution (object) rows were accepted
    assert model.y.shape == (9, 1) and abs(model.y.mean()) < 1e-9                           # standardised fitness (base.py:437)
    assert np.isfinite(fopt) and len(xopt) == dim
    assert opt._optimizer == "BFGS"       # hasattr(model, "gradient") selects upstream's L-BFGS-B maximiser (base.py:201)
    # the evaluated points are where upstream's EI (not ours) had its maximum: its row-by-row calls went through model.predict
    yh = model.predict(opt.data)
    assert yh.shape == (9, 1)


def test_acquisition_classes_swap_in_by_name_and_candidates_bind(upstream, monkeypatch):
    import bayesian_optimization_b200 as b2
    from bayes_optim.acquisition import acquisition_fun as AF
    from bayes_optim.base import BaseBO

    dim = 2
    for name in ("EI", "PI", "EpsilonPI", "UCB", "MGFI"):
        monkeypatch.setattr(AF, name, getattr(b2, name))   # what the stub of INTEGRATION.md 2b does

    calls = {"n": 0}

    def set_argmax(self, fixed=None):                         # base.py:231-243 with the candidate maximiser bound instead
        fixed = {} if fixed is None else fixed
        data = np.asarray(self.data, dtype=float) if hasattr(self, "data") else None

        def maximiser(criteria, logger=None):
            calls["n"] += 1
            return b2.argmax_candidates(criteria, search_space=self.search_space.filter(fixed.keys(), invert=True),
                                        h=None, g=None, eval_budget=self.AQ_max_FEs, n_restart=self.AQ_n_restart,
                                        wait_iter=self.AQ_wait_iter, optimizer=self._optimizer, logger=logger,
                                        n_candidates=2048, refine_top=8, refine_steps=3, data=data)

        self._argmax_restart = maximiser

    monkeypatch.setattr(BaseBO, "_BaseBO__set_argmax", set_argmax)
    space = upstream.RealSpace([-5, 5]) * dim
    np.random.seed(1)
    opt = upstream.BO(search_space=space, obj_fun=sphere, model=make_model(dim), DoE_size=6, max_FEs=9, verbose=False,
                      n_point=1, acquisition_fun="EI", acquisition_optimization={"optimizer": "B200_candidates", "max_FEs": 50})
    # (an optimiser name upstream does not know needs an explicit "max_FEs": base.py:212-215 looks the default up by name)
    xopt, fopt, _ = opt.run()
    assert calls["n"] == 3 and np.isfinite(fopt)
    crit = opt._create_acquisition(par={}, return_dx=False)
    assert isinstance(b2.candidates.unwrap_criterion(crit), b2.EI)       # partial_argument(partial(criterion)) unwraps
    xo, fo = opt._argmax_restart(crit)
    assert isinstance(xo, list) and len(xo) == dim and isinstance(fo, float)   # argmax_restart's return shape (:149-153)

    # ParallelBO: q sampled t values, one maximiser call per criterion (bayes_opt.py:100-113)
    calls["n"] = 0
    np.random.seed(2)
    popt = upstream.ParallelBO(search_space=space, obj_fun=sphere, model=make_model(dim), DoE_size=6, max_FEs=12, verbose=False,
                               n_point=3, acquisition_fun="MGFI", acquisition_par={"t": 2}, n_job=1,
                               acquisition_optimization={"optimizer": "B200_candidates", "max_FEs": 50})
    popt.run()
    assert calls["n"] == 6 and popt.data.shape[0] >= 12


def test_save_and_load_with_dill(upstream, tmp_path):
    dim = 2
    space = upstream.RealSpace([-5, 5]) * dim
    np.random.seed(3)
    opt = upstream.BO(search_space=space, obj_fun=sphere, model=make_model(dim), DoE_size=5, max_FEs=7, verbose=False,
                      n_point=1, acquisition_optimization={"max_FEs": 30, "n_restart": 1})
    opt.run()
    Xq = np.array([[0.3, -1.2], [2.0, 2.0]])
    before = opt.model.predict(Xq, eval_MSE=True)
    path = str(tmp_path / "bo.dump")
    opt.save(path)
    back = upstream.BO.load(path)
    assert back.model._engine is None and back.model.is_fitted     # the device handle is dropped, the state comes back lazily
    after = back.model.predict(Xq, eval_MSE=True)
    np.testing.assert_allclose(after[0], before[0], rtol=1e-12)
    np.testing.assert_allclose(after[1], before[1], rtol=1e-10, atol=1e-14)
