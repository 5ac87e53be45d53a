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


class ArgmaxExchange:
    """The global arg-max exchange with the host out of the data path, pipelined one step deep.

    ``submit(engine, offset)`` (right after ``engine.acq``) has the engine write this rank's pairs into its row of a
    (world, 2q) int64 block ON THE DEVICE (``b200bo_best_pairs_device``, ordered on the engine's stream), enqueues ONE
    ``all_reduce(SUM)`` of the block (NCCL over NVLink; gloo on CPU tensors in the tests) and an asynchronous copy of the
    reduced block into pinned host memory, and returns a ticket at once -- the next step's kernels can be launched
    before the exchange of this one has run.  ``ticket.result()`` waits for that copy and merges the (world, q) pairs
    with numpy's rule.  Two blocks alternate, so one exchange may be in flight while the next is submitted."""

    def __init__(self, q: int, device=None, group=None):
        import torch
        import torch.distributed as dist

        self.q, self.group = int(q), group
        self.dist_on = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.dist_on else 1
        self.rank = dist.get_rank(group) if self.dist_on else 0
        self.device = torch.device("cpu") if device is None else torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.blocks = [torch.zeros(self.world, 2 * self.q, dtype=torch.int64, device=self.device) for _ in range(2)]
        self.host = [torch.zeros(self.world, 2 * self.q, dtype=torch.int64, pin_memory=self.cuda) for _ in range(2)]
        self.turn = 0
        self.done = [None, None]
        # the collective and the copy back run on a high-priority side stream: the compute stream goes straight on to
        # the next step's kernel instead of waiting for the slowest rank's contribution
        self.side = torch.cuda.Stream(self.device, priority=-1) if self.cuda else None

    class Ticket:
        def __init__(self, ex, slot, event):
            self.ex, self.slot, self.event = ex, slot, event

        def result(self):
            if self.event is not None:
                self.event.synchronize()
            h = self.ex.host[self.slot].numpy().reshape(self.ex.world, 2, self.ex.q)
            return merge_argmax(h[:, 0].copy().view(np.float64), h[:, 1].copy())

    def submit(self, engine=None, offset: int = 0, local_val=None, local_idx=None):
        """pairs from the engine's device buffers (``engine`` given), or from host arrays (CPU tests, gloo)"""
        import torch
        import torch.distributed as dist

        slot, self.turn = self.turn, self.turn ^ 1
        blk = self.blocks[slot]
        if self.done[slot] is not None:
            self.done[slot].synchronize()   # the exchange that used this block two submits ago has been read
        if engine is not None and self.cuda:
            engine.best_pairs_device(blk, offset, self.rank, self.world)
        else:
            gidx = np.where(local_idx >= 0, local_idx + offset, -1).astype(np.int64)
            row = np.concatenate([np.ascontiguousarray(local_val, dtype=np.float64).view(np.int64), gidx])
            blk.zero_()
            blk[self.rank] = torch.from_numpy(row).to(self.device)
        if not self.cuda:
            if self.dist_on:
                dist.all_reduce(blk, op=dist.ReduceOp.SUM, group=self.group)   # gloo: blocking
            self.host[slot].copy_(blk)
            return ArgmaxExchange.Ticket(self, slot, None)
        cur = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        self.side.wait_event(ready)
        with torch.cuda.stream(self.side):
            if self.dist_on:
                dist.all_reduce(blk, op=dist.ReduceOp.SUM, group=self.group)   # NCCL: stream-ordered on the side stream
            self.host[slot].copy_(blk, non_blocking=True)
            event = torch.cuda.Event()
            event.record(self.side)
        self.done[slot] = event
        return ArgmaxExchange.Ticket(self, slot, event)
