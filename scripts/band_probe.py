#!/usr/bin/env python
"""Developer probe of the arg-max band of the tensor-core path: for each workload, the a-priori half-widths, the
calibrated ones, the errors observed on a float64-checked sample, the band size and the step / band-stage times.
  python scripts/band_probe.py [C3 C4 ...] [--m 1250000] [--out gpurun_out/band_probe.json]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bayesian_optimization_b200 as b2  # noqa: E402
from bayesian_optimization_b200 import _lib, workloads as wl  # noqa: E402


def main():
    names = [a for a in sys.argv[1:] if a in wl.WORKLOADS] or ["C3"]
    M = int(sys.argv[sys.argv.index("--m") + 1]) if "--m" in sys.argv else 1_250_000
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    res = {}
    for name in names:
        w = wl.WORKLOADS[name]
        X, y, theta = wl.canonical_problem(w.N, w.D)
        gp = b2.GaussianProcess(mean=b2.constant_trend(w.D), corr=w.corr, thetaL=[1e-5] * w.D, thetaU=[1e2] * w.D, nugget=w.nugget)
        gp.fit_fixed(X, y, theta, 1.0)
        eng = gp.engine
        eng.set_precision(_lib.PREC_FAST)
        params = wl.acquisition_params(w)
        acq = wl.ACQ_IDS[w.acq]
        plugin = float(np.min(gp.y))
        Xc = wl.canonical_candidates(M, w.D)
        import torch

        xd = torch.from_numpy(Xc).cuda()
        eng.acq(xd, acq, True, plugin, params)  # calibration + warm-up
        steps = []
        for _ in range(5):
            t0 = time.perf_counter()
            bv, bi, _ = eng.acq(xd, acq, True, plugin, params)
            wall = 1e3 * (time.perf_counter() - t0)
            t = eng.timings()
            steps.append({"wall_ms": wall, "device_ms": t[0], "fused_ms": t[2], "band_ms": t[3], "launches": t[5],
                          "rescored": t[6], "passes": t[7], "products": t[8]})
        info = eng.band_info()
        chk = eng.fast_check(stride=50)
        eng.set_precision(_lib.PREC_FP64)
        bv_e, bi_e, _ = eng.acq(xd, acq, True, plugin, params)
        res[name] = {"M": M, "steps": steps, "band_info": info, "fast_check": chk,
                     "argmax_equal_fp64": bool(np.array_equal(bi, bi_e)),
                     "value_rel_diff": float(np.abs(bv - bv_e).max() / max(np.abs(bv_e).max(), 1e-300))}
        print(name, json.dumps(res[name]["steps"][-1]))
        print("   info", json.dumps(info))
        print("   check", json.dumps(chk), "argmax equal:", res[name]["argmax_equal_fp64"], flush=True)
        del gp, eng
    if out:
        os.makedirs(os.path.dirname(out), exist_ok=True)
        json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
