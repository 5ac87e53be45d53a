#!/usr/bin/env python
"""Fixed-theta fit timing on a bench workload: device time (CUDA events inside the library) and wall time per factor()
(developer tool; B200BO_CHOL_LOOKAHEAD / B200BO_GRAPHS apply).  usage: python scripts/fit_time.py [workload] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bayesian_optimization_b200 as b2
from bayesian_optimization_b200 import workloads

wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cfg = workloads.WORKLOADS[wl]
N, D = cfg.N, cfg.D
X, y, theta = workloads.canonical_problem(N, D)
gp = b2.GaussianProcess(mean=b2.constant_trend(D), corr=cfg.corr, thetaL=[1e-5] * D, thetaU=[1e2] * D, nugget=cfg.nugget)
gp._check_data(X, y)
par = np.r_[theta, 1.0]
dev, wall, llf = [], [], None
for i in range(reps):
    t0 = time.perf_counter()
    llf = gp.log_likelihood_concentrated(par * (1.0 + 1e-3 * i))   # a slightly different theta every call, as L-BFGS does
    wall.append((time.perf_counter() - t0) * 1e3)
    dev.append(gp.engine.fit_timings()[:5].copy())
d = np.array(dev)
print("workload=%s N=%d lookahead=%s graphs=%s: device ms first %.3f | median of the rest %.3f (assemble %.3f cholesky %.3f trtri %.3f solves %.3f) | wall ms median %.3f  llf=%r" % (
    wl, N, os.environ.get("B200BO_CHOL_LOOKAHEAD", "1"), os.environ.get("B200BO_GRAPHS", "1"), d[0, 0],
    np.median(d[2:, 0]), *np.median(d[2:, 1:5], axis=0), np.median(wall[2:]), llf))
