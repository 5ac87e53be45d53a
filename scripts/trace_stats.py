#!/usr/bin/env python
"""Per-super-tile summary of a B200BO_TRACE timeline of the generation-5 kernel (MMA issuer of CTA 0)."""
import sys
import numpy as np
rows = [l.split() for l in open(sys.argv[1]) if not l.startswith('#')]
a = np.array([[int(x) for x in r] for r in rows])
ld = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
nch = ld // 64
per_super = [min(nch, 6 * (s + 1)) for s in range((ld + 383) // 384)]
w, r, i = a[:, 1], a[:, 2], a[:, 3]
n = int((w >= 0).sum())
print("chunks traced", n, " mean clk/chunk %.0f" % np.diff(w[:n]).mean())
idx = 0
t = 0
while idx < n:
    for s, k in enumerate(per_super):
        if idx + k >= n:
            idx = n
            break
        seg = slice(idx, idx + k)
        print("tile %d s=%2d chunks=%2d  clk/chunk %6.0f  pre-issue(max) %5d  issue(med) %5.0f" % (
            t, s, k, (w[idx + k] - w[idx]) / k, (r - w)[seg].max(), np.median((i - r)[seg])))
        idx += k
    t += 1
