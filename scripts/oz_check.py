#!/usr/bin/env python
"""Developer check + timing of the tensor-core trailing update (oz_kernels.cuh) through b200bo_debug_oz_syrk:
C - P P^T on the lower tiles against an extended-precision host product, for 7 / 8 digit planes and the fp64 DMMA kernel.
  python scripts/oz_check.py [rows ...] [--out file.json]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bayesian_optimization_b200 import _lib  # noqa: E402


def reference(P, Cm):
    Pl = P.astype(np.longdouble)
    out = Cm.astype(np.longdouble)
    bs = 512
    for i in range(0, P.shape[0], bs):
        out[i:i + bs] -= Pl[i:i + bs] @ Pl.T
    return out


def main():
    rows_list = [int(a) for a in sys.argv[1:] if a.isdigit()] or [64, 192, 1024, 3968]
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    eng = _lib.Engine(0)
    rng = np.random.default_rng(7)
    res = {}
    for rows in rows_list:
        P = rng.standard_normal((rows, 64)) * 10.0 ** rng.uniform(-3, 0, size=(rows, 1))
        P[:, ::7] *= 1e-4   # wide dynamic range inside a row
        Cm = rng.standard_normal((rows, rows))
        Cm = Cm + Cm.T
        exact = reference(P, Cm) if rows <= 4096 else None
        tril = np.tril_indices(rows)
        scale = np.abs(P).max(1)
        denom = (scale[:, None] * scale[None, :] * 64 + np.abs(Cm) * 2.0 ** -53)[tril]
        rec = {}
        for digits in (8, 7, 0):
            got, ms = eng.debug_oz_syrk(P, Cm, digits=digits, reps=5 if rows >= 1024 else 1)
            r = {"ms": ms}
            if exact is not None:
                err = np.abs((got.astype(np.longdouble) - exact))[tril].astype(np.float64)
                r["max_abs_err"] = float(err.max())
                r["max_err_over_rowscale_product_x64"] = float((err / denom).max())
            else:
                r["max_abs_diff_vs_digits8"] = float(np.abs(got - first)[tril].max()) if digits != 8 else 0.0
            if digits == 8:
                first = got
            upd = rows * (rows + 1) / 2
            r["G_elem_per_s"] = upd / ms / 1e6
            r["eff_TFLOPs"] = upd * 128 / ms / 1e9
            r["GBps_C_traffic"] = upd * 16 / ms / 1e6
            rec[str(digits)] = r
            print(rows, digits, json.dumps(r), flush=True)
        res[str(rows)] = rec
    if out:
        os.makedirs(os.path.dirname(out), exist_ok=True)
        json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
