"""GPU parity of the general-nu Matern kernel (kernel.py:201-207: 2^(1-nu)/Gamma(nu) t^nu K_nu(t), upstream through
scipy.special.kv, here through Temme's method on the device) against golden vectors produced by the reference with
``corr=functools.partial(matern, nu=...)``."""
import functools

import numpy as np
import pytest

import bayesian_optimization_b200 as b2
from bayesian_optimization_b200 import _lib
from oracle import gp_oracle as go

from conftest import load_golden
from gpu_common import matern

pytestmark = pytest.mark.gpu

MNU = load_golden("matern_nu")


def build(c):
    D = c["X"].shape[1]
    mode, ok = int(c["mode"]), bool(c["ok"])
    mean = b2.constant_trend(D) if ok else b2.constant_trend(D, beta=float(np.ravel(c["beta_in"])[0]))
    kw = dict(mean=mean, corr=functools.partial(matern, nu=float(c["nu"])), thetaL=[1e-5] * D, thetaU=[1e2] * D)
    if mode == go.MODE_NOISELESS:
        kw.update(nugget=None)
    elif mode == go.MODE_NOISY:
        kw.update(nugget=float(c["nugget"]))
    else:
        kw.update(nugget=float(c["nugget"]), noise_estim=True)
    return b2.GaussianProcess(**kw), (None if mode == go.MODE_NOISELESS else float(c["par_last"]))


@pytest.mark.parametrize("name", sorted(MNU))
def test_matern_nu_fit_predict_acq(name):
    c = MNU[name]
    gp, last = build(c)
    llf = gp.fit_fixed(c["X"], c["y"], c["theta"], last)
    assert gp._corr_id == _lib.CORR_MATERN_NU and gp._corr_extra == float(c["nu"])
    if not np.isfinite(c["llf"]):
        assert np.isneginf(llf) and not gp.is_fitted          # llf > 0 is rejected, gpr.py:981-982
        return
    rt = 1e-7 if "_nl_" in name else 1e-9
    assert llf == pytest.approx(float(c["llf"]), rel=1e-8, abs=1e-9)
    assert float(gp.sigma2[0]) == pytest.approx(float(c["sigma2"]), rel=rt)
    yh, ms = gp.predict(c["Xc"], eval_MSE=True)
    np.testing.assert_allclose(yh.ravel(), c["yhat"], rtol=rt, atol=1e-9)
    np.testing.assert_allclose(ms.ravel(), c["mse"], rtol=10 * rt, atol=1e-9 * float(c["sigma2"]))
    np.testing.assert_allclose(b2.EI(model=gp)(c["Xc"]), c["ei"], rtol=1e-6, atol=1e-300)
    gp.engine.set_precision(_lib.PREC_FAST)                    # no tensor-core form: the float64 path answers
    bv, bi = b2.MGFI(model=gp, t=float(c["t"])).argmax(c["Xc"])
    assert int(bi[0]) == int(np.argmax(c["mgfi"]))
