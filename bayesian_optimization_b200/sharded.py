"""Candidate data-parallelism over the GPUs of one box.

The path shards embarrassingly: every candidate's (yhat, s^2, acquisition) depends only on the fitted
state, which each rank recomputes deterministically from the same (X, y, theta) -- bit-identical, no
broadcast.  The one exchange is the global arg-max: a single all-reduce (or all-gather) of q (value, index) pairs per
rank (16 q bytes each) over NCCL/NVLink, merged with numpy's arg-max rule (largest value, ties -> lowest GLOBAL
index).  The reference's analogue is the joblib fan-out of q argmax_restart calls (bayes_opt.py:108-111).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_bounds(M_total: int, world: int, rank: int) -> Tuple[int, int]:
    """contiguous block [lo, hi) of rank; the first (M_total % world) ranks take one extra candidate"""
    base, extra = divmod(int(M_total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def merge_argmax(vals: np.ndarray, idxs: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """vals, idxs: (world, q) per-rank bests with GLOBAL indices (idx < 0: empty shard).
    numpy arg-max rule: NaN wins, else larger value, ties -> lowest index."""
    world, q = vals.shape
    bv = np.empty(q)
    bi = np.empty(q, dtype=np.int64)
    for c in range(q):
        best, arg = 0.0, -1
        for r in range(world):
            v, i = vals[r, c], int(idxs[r, c])
            if i < 0:
                continue
            if arg < 0:
                best, arg = v, i
                continue
            vn, bn = v != v, best != best
            if vn or bn:
                better = vn and (not bn or i < arg)
            else:
                better = v > best or (v == best and i < arg)
            if better:
                best, arg = v, i
        bv[c], bi[c] = best, arg
    return bv, bi


def global_argmax(local_val: np.ndarray, local_idx: np.ndarray, offset: int, group=None, device=None,
                  collective: str = "allreduce"):
    """Exchange the per-rank (value, global index) pairs with ONE collective and merge with numpy's arg-max rule.
    Works on NCCL (CUDA tensors) and gloo (CPU tensors).  Returns (best_val (q,), best_idx (q,)).

    collective="allreduce" (default): every rank writes its 2 q int64 words (value bits, index) into its own row of a
    zeroed (world, 2 q) buffer and the buffer is summed -- one ``ncclAllReduce`` of 16 q world bytes over NVLink; each
    element has exactly one non-zero contributor, so the integer sum is exact and the float64 values arrive bit for
    bit (a max-reduction on a packed key, SURVEY 8e, would have to truncate them).  collective="allgather": the same
    payload through ``all_gather_into_tensor``."""
    import torch
    import torch.distributed as dist

    q = local_val.shape[0]
    gidx = np.where(local_idx >= 0, local_idx + offset, -1).astype(np.int64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return merge_argmax(local_val[None], gidx[None])
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    payload = torch.from_numpy(np.concatenate([np.ascontiguousarray(local_val, dtype=np.float64).view(np.int64), gidx]))
    if device is not None:
        payload = payload.to(device)
    if collective == "allreduce":
        out = torch.zeros(world, 2 * q, dtype=torch.int64, device=payload.device)
        out[rank] = payload
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    elif collective == "allgather":
        out = torch.empty(world * 2 * q, dtype=torch.int64, device=payload.device)
        dist.all_gather_into_tensor(out, payload, group=group)
    else:
        raise ValueError("collective should be 'allreduce' or 'allgather'")
    host = out.cpu().numpy().reshape(world, 2, q)
    return merge_argmax(host[:, 0].copy().view(np.float64), host[:, 1])
