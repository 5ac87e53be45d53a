"""2+ ranks on NCCL: the single-all-reduce global arg-max against the all-gather form and numpy (developer check)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from bayesian_optimization_b200 import sharded

rank = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl")
world = dist.get_world_size()
rng = np.random.default_rng(7)
x = rng.integers(0, 40, (5, 1001)).astype(float) - 20.5
lo, hi = sharded.shard_bounds(1001, world, rank)
lv = np.array([x[c, lo:hi].max() for c in range(5)])
li = np.array([int(np.argmax(x[c, lo:hi])) for c in range(5)], dtype=np.int64)
for coll in ("allreduce", "allgather"):
    bv, bi = sharded.global_argmax(lv, li, lo, device=torch.device("cuda", rank), collective=coll)
    assert list(bi) == [int(np.argmax(x[c])) for c in range(5)], (coll, bi)
    assert list(bv) == [x[c].max() for c in range(5)], (coll, bv)
dist.destroy_process_group()
print("NCCL argmax OK", rank, flush=True)
