#!/usr/bin/env python
"""Developer measurements that back numbers quoted in DESIGN.md / include/b200bo.h:
  (1) errors of the three-product tensor-core predict() against the float64 device path (and what the a-priori model
      states) on the shapes of tests/test_fast_gpu.py and on the bench workloads;
  (2) throughput of b200bo_gradient / b200bo_acq_grad at C3 (wall clock around the C ABI call, host buffers).
  python scripts/measure_misc.py [--out file.json]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bayesian_optimization_b200 as b2  # noqa: E402
from bayesian_optimization_b200 import _lib, workloads as wl  # noqa: E402


def model(N, D, corr, nugget=1e-6):
    X, y, theta = wl.canonical_problem(N, D)
    gp = b2.GaussianProcess(mean=b2.constant_trend(D), corr=corr, thetaL=[1e-5] * D, thetaU=[1e2] * D, nugget=nugget)
    gp.fit_fixed(X, y, theta, 1.0)
    return gp


def main():
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    res = {"fast_predict": {}, "gradient": {}}
    shapes = [(512, 8, "squared_exponential"), (640, 5, "matern52"), (1100, 16, "matern32"), (300, 3, "matern12"),
              (2048, 40, "squared_exponential"), (1024, 8, "squared_exponential"), (4096, 16, "matern52"), (8192, 32, "squared_exponential"),
              (2048, 64, "squared_exponential")]
    for N, D, corr in shapes:
        nug = 1e-2 if N == 8192 else 1e-6
        gp = model(N, D, corr, nug)
        M = 20000
        Xc = wl.canonical_candidates(M, D)
        y64, m64 = gp.engine.predict(Xc, True)
        gp.engine.set_precision(_lib.PREC_FAST)
        yf, mf = gp.engine.predict(Xc, True)
        info = gp.engine.band_info()
        s2 = float(np.ravel(gp.sigma2)[0])
        ss = np.maximum(1.0 - m64 / s2, 0.0)
        allowed_s = info["ds_abs_3"] + info["ds_rel_3"] * np.sqrt(ss + 1e-3)
        r = {"max_err_yhat": float(np.abs(yf - y64).max()), "max_err_mse_over_sigma2": float(np.abs(mf - m64).max() / s2),
             "dy_model": info["dy_model"], "ds_model_max_over_sigma2": float(allowed_s.max() / s2),
             "max_err_mse_over_model": float((np.abs(mf - m64) / allowed_s).max()), "max_err_y_over_model": float(np.abs(yf - y64).max() / info["dy_model"])}
        res["fast_predict"][f"N{N}_D{D}_{corr}"] = r
        print("fast_predict", N, D, corr, json.dumps(r), flush=True)
        gp.engine.set_precision(_lib.PREC_FP64)
        if (N, D) == (4096, 16):
            for M in (4096, 65536):
                Xg = wl.canonical_candidates(M, D)
                gp.engine.gradient(Xg[:1024])
                t0 = time.perf_counter(); gp.engine.gradient(Xg); t1 = time.perf_counter()
                gp.engine.acq_grad(Xg, _lib.ACQ_EI, True, float(gp.y.min()), 0.0)
                t2 = time.perf_counter(); gp.engine.acq_grad(Xg, _lib.ACQ_EI, True, float(gp.y.min()), 0.0); t3 = time.perf_counter()
                res["gradient"][f"C3_M{M}"] = {"gradient_points_per_s": M / (t1 - t0), "acq_grad_points_per_s": M / (t3 - t2),
                                               "gradient_ms": 1e3 * (t1 - t0), "acq_grad_ms": 1e3 * (t3 - t2)}
                print("gradient", M, json.dumps(res["gradient"][f"C3_M{M}"]), flush=True)
        del gp
    if out:
        os.makedirs(os.path.dirname(out), exist_ok=True)
        json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
