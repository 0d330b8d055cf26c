"""Data-parallel host logic of the hot path: trajectories shard across ranks,
every rank holds a full parameter replica, scales its loss by ``1 / B_global`` (``nj_ode_loss(...,
traj_scale=...)``) and ONE all-reduce of the flat gradient vector with the loss appended precedes the
optimiser step.  There is no data-path collective.  Works with any ``torch.distributed`` backend
(NCCL over NVLink on the B200 box, gloo in the CPU tests).

Two ways to get that all-reduce -- use ONE of them, never both (the gradients would be summed twice):
  * ``model.enable_data_parallel()``: the reverse sweep all-reduces its flat gradient buffer in place (what
    ``bench.py`` and the packed training loop use);
  * ``allreduce_gradients(params, loss)`` below, for gradients that did not come out of the sweep with data
    parallelism enabled (it also carries the loss in the same collective).
Every rank must take part in the collective, so every rank needs at least one trajectory: ``shard_bounds`` refuses
a batch with fewer trajectories than ranks on ALL ranks alike (a rank with an empty shard would raise "empty
batch" on its own while its peers block in the all-reduce).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch


def shard_bounds(n_traj: int, world_size: int, steps_per_traj: Optional[Sequence[int]] = None) -> List[int]:
    """``world_size + 1`` trajectory indices cutting ``range(n_traj)`` into contiguous slices.

    Without ``steps_per_traj`` the slices have (almost) equal trajectory counts; with it they are balanced
    by the number of Euler steps (the work), still contiguous so a packed batch is sliced, not gathered."""
    if world_size < 1 or n_traj < 0:
        raise ValueError("shard_bounds: world_size must be >= 1 and n_traj >= 0")
    if 0 < n_traj < world_size:
        raise ValueError(f"shard_bounds: {n_traj} trajectories cannot be split over {world_size} ranks without an empty "
                         "shard (every rank takes part in the gradient all-reduce)")
    if steps_per_traj is None:
        return [(n_traj * r) // world_size for r in range(world_size + 1)]
    if len(steps_per_traj) != n_traj:
        raise ValueError("shard_bounds: steps_per_traj must have one entry per trajectory")
    total = float(sum(steps_per_traj))
    bounds, acc, b = [0], 0.0, 0
    for r in range(1, world_size):
        target = total * r / world_size
        while b < n_traj and acc + steps_per_traj[b] * 0.5 <= target:
            acc += steps_per_traj[b]
            b += 1
        bounds.append(b)
    bounds.append(n_traj)
    for r in range(1, world_size):                       # work balancing must not starve a rank either
        bounds[r] = min(max(bounds[r], bounds[r - 1] + 1), n_traj - (world_size - r))
    return bounds


def shard_lists(batch_times: List[torch.Tensor], batch_values: List[torch.Tensor], rank: int, world_size: int,
                steps_per_traj: Optional[Sequence[int]] = None) -> Tuple[List[torch.Tensor], List[torch.Tensor], float]:
    """This rank's slice of a list batch and the ``traj_scale`` (= 1 / global batch) to hand to ``nj_ode_loss``."""
    bounds = shard_bounds(len(batch_times), world_size, steps_per_traj)
    lo, hi = bounds[rank], bounds[rank + 1]
    return batch_times[lo:hi], batch_values[lo:hi], 1.0 / max(len(batch_times), 1)


def allreduce_gradients(params: Sequence[torch.Tensor], loss: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the parameter gradients and the (already ``1/B_global``-scaled) loss over the ranks with ONE
    all-reduce of the flat vector ``[grad_0, grad_1, ..., loss]``; gradients are written back in place
    (a missing ``.grad`` counts as zero and is materialised).  Returns the global loss (0-dim tensor).
    Do NOT combine with ``model.enable_data_parallel()`` (see the module docstring)."""
    import torch.distributed as dist
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    grads = [p.grad for p in params]
    flat = torch.cat([g.reshape(-1) for g in grads] + [loss.detach().reshape(1).to(grads[0].dtype if grads else loss.dtype)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    views, o = [], 0
    for g in grads:
        views.append(flat[o:o + g.numel()].view_as(g))
        o += g.numel()
    if grads:
        torch._foreach_copy_(grads, views)       # one multi-tensor kernel instead of one copy per parameter
    return flat[o]
