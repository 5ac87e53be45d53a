"""GPU parity of the generalized_exponential kernel (kernel.py:332-374; theta = [theta_1..n, p]) against golden
vectors produced by the reference; float64 path (the tensor-core flavour falls back to it)."""
import numpy as np
import pytest

import bayesian_optimization_b200 as b2
from bayesian_optimization_b200 import _lib
from oracle import gp_oracle as go

from conftest import load_golden
from gpu_common import fit_case

pytestmark = pytest.mark.gpu

GENEXP = load_golden("genexp")


@pytest.mark.parametrize("name", sorted(GENEXP))
def test_genexp_fit_predict_acq(name):
    c = GENEXP[name]
    gp, llf = fit_case(c)
    rt = 1e-7 if "_nl_" in name else 1e-9
    assert llf == pytest.approx(float(c["llf"]), rel=1e-9 if "_nl_" not in name else 1e-7)
    assert float(gp.sigma2[0]) == pytest.approx(float(c["sigma2"]), rel=rt)
    yh, ms = gp.predict(c["Xc"], eval_MSE=True)
    np.testing.assert_allclose(yh.ravel(), c["yhat"], rtol=rt, atol=1e-9)
    np.testing.assert_allclose(ms.ravel(), c["mse"], rtol=10 * rt, atol=1e-9 * float(c["sigma2"]))
    ei = b2.EI(model=gp)
    np.testing.assert_allclose(ei(c["Xc"]), c["ei"], rtol=1e-6, atol=1e-300)
    gp.engine.set_precision(_lib.PREC_FAST)          # no tensor-core form for |d|^p: runs the float64 path
    bv, bi = b2.MGFI(model=gp, t=float(c["t"])).argmax(c["Xc"])
    assert int(bi[0]) == int(np.argmax(c["mgfi"]))
    with pytest.raises(b2.B200BOError):              # corr_dx leaves this kernel unimplemented (gpr.py:652-653)
        gp.engine.gradient(c["Xc"][:2])


def test_genexp_isotropic_and_argument_checks():
    """theta = [theta, p]: upstream raises IndexError on this form (kernel.py:365-376); the device takes it as the
    docstring describes it and is checked against the oracle"""
    c = GENEXP["gexp_ard_ny_ok"]
    X, y, Xc = c["X"], c["y"], c["Xc"]
    D = X.shape[1]
    gp = b2.GaussianProcess(mean=b2.constant_trend(D), corr="generalized_exponential", thetaL=[1e-5] * 2, thetaU=[1e2] * 2, nugget=1e-2)
    llf = gp.fit_fixed(X, y, [0.7, 1.3], 0.8)
    ora = go.fit_fixed(X, y, go.CORR_GENEXP, [0.7, 1.3], go.MODE_NOISY, sigma2=0.8, noise_var=1e-2)
    assert llf == pytest.approx(ora.llf, rel=1e-10)
    yh, ms = gp.predict(Xc, eval_MSE=True)
    yo, mo = go.predict(ora, Xc)
    np.testing.assert_allclose(yh, yo, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(ms, mo, rtol=1e-8, atol=1e-10)
    with pytest.raises(b2.B200BOError):
        gp.engine.factor(_lib.CORR_GENEXP, [0.5, 0.5, 1.5], _lib.MODE_NOISY, 0.8, 1e-2)   # neither 2 nor D + 1 entries
