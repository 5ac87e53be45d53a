import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bayesian_optimization_b200 as b2
from bayesian_optimization_b200 import workloads
D, N0, m = 16, 4064, 32
X, y, theta = workloads.canonical_problem(N0 + m, D)
a = b2.GaussianProcess(mean=b2.constant_trend(D), corr="matern52", thetaL=[1e-5] * D, thetaU=[1e2] * D, nugget=1e-6)
for it in range(4):
    print("fit", a.fit_fixed(X[:N0], y[:N0], theta, 1.0), a.engine.fit_timings()[:6])
    print("  engine N", a.engine.N, "X", a.X.shape)
    r = a._engine.append(X[N0:], y)
    print("  append ->", r, a.engine.fit_timings()[:6])
