"""Packed training step around the hot path.

The reference's ``Trainer.train_epoch`` (utils/training.py:33-103) shuffles the trajectory indices with Python's
``random``, then for every mini-batch copies ``batch_size`` list elements to the device one by one, zeroes the
gradients, runs ``model`` / ``nj_ode_loss`` / ``backward`` / ``optimizer.step()`` and synchronises on ``loss.item()``.
With the sweeps on the GPU that host work is the step time.  ``train_epoch_packed`` keeps the SAME sequence of
mini-batches and optimiser steps (same shuffle when given the same ``random`` state, same tail batch, same
``ignore_first_continuity`` / ``moment_weights`` / ``variance_method`` semantics) on a dataset that lives on the
device as ONE ``PackedBatch``: a mini-batch is a device-side gather, the losses stay on the device until the epoch
ends (one D2H read), and with ``FlatAdam`` the optimiser step is one launch on the sweep's own flat gradient.
"""
from __future__ import annotations

import random
from typing import List, Optional

import torch

from .models.jump_ode import NeuralJumpODE, nj_ode_loss
from .packed import PackedBatch


def epoch_order(n: int, shuffle: bool = True, rng: Optional[random.Random] = None) -> List[int]:
    """The reference's mini-batch order: ``indices = list(range(n)); random.shuffle(indices)`` (training.py:55-56)."""
    idx = list(range(n))
    if shuffle:
        (rng or random).shuffle(idx)
    return idx


def train_epoch_packed(model: NeuralJumpODE, optimizer: torch.optim.Optimizer, data: PackedBatch, batch_size: Optional[int] = None,
                       ignore_first_continuity: bool = False, moment_weights=None, variance_method: Optional[str] = None,
                       shuffle: bool = True, rng: Optional[random.Random] = None, order: Optional[List[int]] = None) -> float:
    """One epoch over a device-resident dataset; returns the mean mini-batch loss like ``Trainer.train_epoch``
    (training.py:101-103).  ``batch_size=None`` = the whole dataset in one step (training.py:58-72)."""
    model.train()
    n = data.B
    if batch_size is None or batch_size >= n:
        batch_size = n
    idx = epoch_order(n, shuffle, rng) if order is None else list(order)
    dev = data.device
    idx_dev = torch.tensor(idx, dtype=torch.int64).to(dev, non_blocking=True)
    vm = variance_method if variance_method is not None else getattr(model, "variance_method", "direct")
    losses = []
    for lo in range(0, n, batch_size):
        hi = min(lo + batch_size, n)
        mb = data if (lo == 0 and hi == n and not shuffle) else data.gather(idx_dev[lo:hi], idx[lo:hi])
        optimizer.zero_grad(set_to_none=True)
        preds, before = model.forward_packed(mb)
        loss = nj_ode_loss(mb, None, preds, before, ignore_first_continuity=ignore_first_continuity,
                           moment_weights=moment_weights, variance_method=vm)
        loss.backward()
        optimizer.step()
        losses.append(loss.detach())
    return float(torch.stack(losses).mean().item())          # the epoch's only device -> host read


@torch.no_grad()
def validate_packed(model: NeuralJumpODE, data: PackedBatch, ignore_first_continuity: bool = False, moment_weights=None,
                    variance_method: Optional[str] = None) -> float:
    """``Trainer.validate`` (training.py:105-124): the whole set in one forward-only call."""
    was = model.training
    model.eval()
    vm = variance_method if variance_method is not None else getattr(model, "variance_method", "direct")
    preds, before = model.forward_packed(data)
    loss = nj_ode_loss(data, None, preds, before, ignore_first_continuity=ignore_first_continuity,
                       moment_weights=moment_weights, variance_method=vm)
    model.train(was)
    return float(loss.item())


@torch.no_grad()
def relative_loss_packed(model: NeuralJumpODE, data: PackedBatch, process_type: str, process_params: dict,
                         moment_weights=None, variance_method: Optional[str] = None) -> float:
    """The reference's relative-loss metric (utils/training.py:219-261) on a packed batch:
    ``(L_model - L_true) / max(L_true, 1e-8)`` where ``L_true`` is ``nj_ode_loss`` evaluated on the closed-form
    conditional moments at the observations (``simulation.conditional_moments_packed``).  As in the reference neither
    loss call passes ``ignore_first_continuity`` (training.py:225-227, :250-252)."""
    from .simulation import conditional_moments_packed
    was = model.training
    model.eval()
    vm = variance_method if variance_method is not None else getattr(model, "variance_method", "direct")
    preds, before = model.forward_packed(data)
    l_model = nj_ode_loss(data, None, preds, before, moment_weights=moment_weights, variance_method=vm)
    true, true_before = conditional_moments_packed(data, process_type, num_moments=model.num_moments, variance_method=vm,
                                                   **process_params)
    l_true = nj_ode_loss(data, None, true, true_before, moment_weights=moment_weights, variance_method=vm)
    model.train(was)
    lm, lt = float(l_model.item()), float(l_true.item())
    return (lm - lt) / max(lt, 1e-8)
