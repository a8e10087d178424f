"""Multi-GPU sharding of the trajectory batch (one process per GPU, torch.distributed).

Trajectories are independent (the reference has no cross-trajectory state), so the batch is
split contiguously over the ranks and nothing is exchanged inside the forward, linearize
and backward kernels.  The only collective is one all-gather per iLQR iteration of the
per-trajectory costs and status flags (8-16 KB at B=1024), which every rank uses for the
global termination test and for reporting.  Backend: NCCL over NVLink on GPUs, gloo in the
CPU tests of this host logic.
"""
from __future__ import annotations

import numpy as np


def shard_range(B: int, rank: int, world: int):
    """Contiguous [lo, hi) slice of the batch owned by ``rank``; sizes differ by at most 1."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_sizes(B: int, world: int):
    return [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]


def all_gather_ragged(local, B: int, group=None):
    """All-gather per-trajectory values (1-D tensor, this rank's shard) into the full batch
    order.  Shards may differ by one element, so pad to the largest shard."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = shard_sizes(B, world)
    pad = max(sizes)
    buf = torch.zeros(pad, dtype=local.dtype, device=local.device)
    buf[: local.numel()] = local
    out = torch.empty(world * pad, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    return torch.cat([out[r * pad: r * pad + sizes[r]] for r in range(world)])


class CostGather:
    """The path's one collective, without stalling the solver: all-gather of the per-trajectory
    cost vectors of all ranks (8 KB per rank at B = 1024) on a side stream, from a snapshot taken
    on the solver's stream, into pre-allocated buffers.  ``issue()`` after an iteration returns at
    once; the gather overlaps the next iteration's line search; ``result()`` waits for it.
    Equal shards (the usual case) gather straight into the output, ragged shards are padded."""

    def __init__(self, solver, B_global: int, group=None):
        import torch
        import torch.distributed as dist
        from . import _lib

        self.torch, self.dist, self.group = torch, dist, group
        self.solver = solver
        self.world = dist.get_world_size(group)
        self.sizes = shard_sizes(B_global, self.world)
        self.pad = max(self.sizes)
        self.equal = min(self.sizes) == self.pad
        self.cost = solver.device_tensor(_lib.COST)
        dev = self.cost.device
        # two snapshot / output pairs used in turn: the solver's stream only ever waits for the gather
        # issued TWO iterations ago, so a rank that runs ahead by a step is not held back by its peer
        self.snap = [torch.zeros(self.pad, dtype=torch.float64, device=dev) for _ in range(2)]
        self.out = [torch.empty(self.world * self.pad, dtype=torch.float64, device=dev) for _ in range(2)]
        self.side = torch.cuda.Stream(device=dev)
        self.ev_snap = [torch.cuda.Event() for _ in range(2)]
        self.ev_done = [torch.cuda.Event() for _ in range(2)]
        self.issued = 0

    def issue(self):
        torch = self.torch
        st = self.solver._stream
        k = self.issued & 1
        with torch.cuda.stream(st):
            if self.issued >= 2:
                st.wait_event(self.ev_done[k])       # the gather that last used this pair has read its snapshot
            self.snap[k][: self.cost.numel()].copy_(self.cost, non_blocking=True)
            self.ev_snap[k].record(st)
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.ev_snap[k])
            self.dist.all_gather_into_tensor(self.out[k], self.snap[k], group=self.group)
            self.ev_done[k].record(self.side)
        self.issued += 1

    def result(self):
        """Global cost vector in batch order (waits for the gather issued last)."""
        assert self.issued > 0, "issue() first"
        k = (self.issued - 1) & 1
        self.ev_done[k].synchronize()
        out = self.out[k]
        if self.equal:
            return out
        return self.torch.cat([out[r * self.pad: r * self.pad + self.sizes[r]] for r in range(self.world)])


class ShardedILQR:
    """BatchedILQR over a global batch, sharded across the ranks of a process group.

    ``make_local(B_local)`` builds this rank's solver (a BatchedILQR on its GPU); inputs are
    given for the *global* batch and sliced here, results are gathered on request.
    """

    def __init__(self, make_local, B: int, group=None):
        import torch.distributed as dist

        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.B = B
        self.lo, self.hi = shard_range(B, self.rank, self.world)
        self.local = make_local(self.hi - self.lo)

    def set_initial_state(self, x0_global):
        self.local.set_initial_state(np.asarray(x0_global)[self.lo:self.hi])

    def set_initial_guess(self, u_global):
        u = np.asarray(u_global)
        self.local.set_initial_guess(u if u.ndim == 2 else u[self.lo:self.hi])

    def iterate(self):
        """One iLQR iteration on every rank's shard, then the all-gather of per-trajectory
        costs.  Returns (global cost vector tensor, number of trajectories still active)."""
        import torch
        import torch.distributed as dist
        from . import _lib

        n_active = self.local.iterate()
        cost = self.local.device_tensor(_lib.COST)
        if self.world > 1:
            cost_all = all_gather_ragged(cost, self.B, self.group)
            act = torch.tensor([n_active], dtype=torch.int64, device=cost.device)
            dist.all_reduce(act, group=self.group)
            n_active = int(act.item())
        else:
            cost_all = cost.clone()
        return cost_all, n_active

    def solve(self, max_iters=0):
        self.local.begin_solve()
        it, n_active, cost = 0, 1, None
        while n_active > 0 and (max_iters <= 0 or it < max_iters):
            cost, n_active = self.iterate()
            it += 1
        return cost, it
