#!/usr/bin/env python
"""Kernel-only timing of the fused tensor-core kernel on a bench workload (developer tool; B200BO_FAST_KERNEL / B200BO_REPLAY_MB apply).
usage: python scripts/fused_time.py [workload] [M] [products] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bayesian_optimization_b200 as b2
from bayesian_optimization_b200 import workloads

wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 303104
prod = int(sys.argv[3]) if len(sys.argv) > 3 else 1
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
cfg = workloads.WORKLOADS[wl]
N, D = cfg.N, cfg.D
X, y, theta = workloads.canonical_problem(N, D)
gp = b2.GaussianProcess(mean=b2.constant_trend(D), corr=cfg.corr, thetaL=[1e-5] * D, thetaU=[1e2] * D, nugget=cfg.nugget)
gp.fit_fixed(X, y, theta, 1.0)
Xc = workloads.canonical_candidates(M, D)
ms = gp.engine.debug_fused_time(Xc, prod, reps)
print("workload=%s N=%d D=%d M=%d products=%d gen=%s: %.3f ms  %.2f Mcand/s  %.0f algorithmic TFLOP/s" % (
    wl, N, D, M, prod, os.environ.get("B200BO_FAST_KERNEL", "default"),
    ms, M / ms / 1e3, M * float(N) * N / ms / 1e9))
