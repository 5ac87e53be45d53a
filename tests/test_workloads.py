"""CPU: workload recipes (package copy == oracle copy), shard arithmetic, arg-max merge, and the world_size-2
gloo run of the one multi-GPU exchange step."""
import os
import subprocess
import sys

import numpy as np

from bayesian_optimization_b200 import sharded, workloads
from oracle import gp_oracle as go

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_recipes_agree_with_oracle():
    for N, D in [(64, 3), (1024, 8)]:
        a, b = workloads.canonical_problem(N, D), go.canonical_problem(N, D)
        for u, v in zip(a, b):
            np.testing.assert_array_equal(u, v)
    np.testing.assert_array_equal(workloads.canonical_candidates(100, 5, 3), go.canonical_candidates(100, 5, 3))
    out = np.empty((100, 5))
    np.testing.assert_array_equal(workloads.canonical_candidates(100, 5, 3, out=out), go.canonical_candidates(100, 5, 3))
    np.testing.assert_array_equal(workloads.acquisition_params(workloads.WORKLOADS["C3"]), go.mgfi_t_samples(2.0, 32))
    np.testing.assert_array_equal(workloads.acquisition_params(workloads.WORKLOADS["C4"]), go.ucb_alpha_samples(0.5, 32))
    np.testing.assert_allclose(workloads.acquisition_params(workloads.WORKLOADS["C5"]), go.annealing_t_schedule(2.0, 0.1, 32))


def test_shard_bounds_cover_exactly():
    for M, W in [(10, 3), (10_000_000, 8), (5, 8), (0, 2)]:
        b = [sharded.shard_bounds(M, W, r) for r in range(W)]
        assert b[0][0] == 0 and b[-1][1] == M
        assert all(b[i][1] == b[i + 1][0] for i in range(W - 1))
        assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_merge_argmax_rule():
    v = np.array([[1.0, 5.0, np.nan], [3.0, 5.0, 7.0], [3.0, 2.0, np.nan]])
    i = np.array([[4, 10, 30], [9, 3, 1], [2, 8, 12]])
    bv, bi = sharded.merge_argmax(v, i)
    assert list(bi) == [2, 3, 12] and bv[0] == 3.0 and bv[1] == 5.0 and np.isnan(bv[2])
    # empty shards are skipped
    bv, bi = sharded.merge_argmax(np.array([[0.0], [-5.0]]), np.array([[-1], [6]]))
    assert bi[0] == 6 and bv[0] == -5.0
    # agrees with numpy on random data split into shards
    rng = np.random.default_rng(1)
    x = rng.integers(0, 50, (4, 1000)).astype(float)  # many ties
    W = 4
    vals, idxs = np.empty((W, 4)), np.empty((W, 4), dtype=np.int64)
    for r in range(W):
        lo, hi = sharded.shard_bounds(1000, W, r)
        for c in range(4):
            k = int(np.argmax(x[c, lo:hi]))
            vals[r, c], idxs[r, c] = x[c, lo + k], lo + k
    bv, bi = sharded.merge_argmax(vals, idxs)
    assert list(bi) == [int(np.argmax(x[c])) for c in range(4)]


WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch.distributed as dist
from bayesian_optimization_b200 import sharded
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(7)
x = rng.integers(0, 40, (3, 999)).astype(float)
lo, hi = sharded.shard_bounds(999, world, rank)
lv = np.array([x[c, lo:hi].max() for c in range(3)])
li = np.array([int(np.argmax(x[c, lo:hi])) for c in range(3)], dtype=np.int64)
for coll in ("allreduce", "allgather"):
    bv, bi = sharded.global_argmax(lv + (0.1 if coll == "allreduce" else 0.0) * 0, li, lo, collective=coll)
    assert list(bi) == [int(np.argmax(x[c])) for c in range(3)], (rank, coll, bi)
    assert list(bv) == [x[c].max() for c in range(3)]
# negative values, NaN and an empty shard survive the integer-sum transport bit for bit
lv2 = np.array([-1.5 - rank, np.nan if rank == 1 else 2.0, 0.0])
li2 = np.array([3, 4, -1 if rank == 0 else 5], dtype=np.int64)
bv, bi = sharded.global_argmax(lv2, li2, 100 * rank)
assert bv[0] == -1.5 and bi[0] == 3 and np.isnan(bv[1]) and bi[1] == 104 and bi[2] == 105, (bv, bi)
# the pipelined exchange (two tickets in flight) gives the same merges
ex = sharded.ArgmaxExchange(3)
t1 = ex.submit(local_val=lv, local_idx=li, offset=lo)
t2 = ex.submit(local_val=lv2, local_idx=li2, offset=100 * rank)
bv, bi = t1.result()
assert list(bi) == [int(np.argmax(x[c])) for c in range(3)] and list(bv) == [x[c].max() for c in range(3)]
bv, bi = t2.result()
assert bv[0] == -1.5 and bi[0] == 3 and np.isnan(bv[1]) and bi[1] == 104 and bi[2] == 105, (bv, bi)
dist.destroy_process_group()
print("OK", rank)
"""


def test_global_argmax_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
         "127.0.0.1", "--master-port", "29731", str(script), ROOT],
        capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("OK") == 2
