"""Closed-form conditional moments at the observation times, vectorised over a packed batch.

The reference evaluates them trajectory by trajectory and observation by observation in Python
(simulation/data_generation.py:543-816, dispatcher :819-921) to feed ``nj_ode_loss`` with the "true" predictions of
the relative-loss metric (utils/training.py:219-261).  Same formulas and conventions here, as a handful of tensor
operations on the device the batch lives on:

  at an observation           mean = the observed value, variance = 0
  just before observation i   conditioned on observation i-1 over dt = t_i - t_{i-1}:
      black_scholes           mean x e^{mu dt}              var x^2 (e^{sigma^2 dt} - 1) e^{2 mu dt}
      ornstein_uhlenbeck      mean x e^{-theta dt} + mu (1 - e^{-theta dt})     var sigma^2 (1 - e^{-2 theta dt}) / (2 theta)
      heston                  the Black-Scholes formulas with sigma := xi       (data_generation.py:621-636, :706-719)
      hybrid_ou_bs            OU formulas among the observations before ``switch_time``, Black-Scholes among those from
                              it on, each regime treated as its own sequence (:722-816); zeros when ``switch_time`` is
                              None (random switch: the reference disables the metric that way, :857-860)
  first observation of a trajectory (or of a hybrid regime): before-mean = the value itself, before-variance = 0
  second column: the variance (``variance_method='direct'``) or variance + mean^2 (``'second_moment'``); columns
  beyond the second stay 0, as in the reference.
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from ..packed import PackedBatch

_PROCESSES = ("black_scholes", "ornstein_uhlenbeck", "heston", "hybrid_ou_bs")


def _bs(x_prev, dt, mu, sigma):
    mean = x_prev * torch.exp(mu * dt)
    var = x_prev ** 2 * (torch.exp(sigma ** 2 * dt) - 1) * torch.exp(2 * mu * dt)
    return mean, var


def _ou(x_prev, dt, theta, mu, sigma):
    decay = torch.exp(-theta * dt)
    mean = x_prev * decay + mu * (1 - decay)
    var = (sigma ** 2 / (2 * theta) * (1 - torch.exp(-2 * theta * dt))).expand_as(x_prev)
    return mean, var


def conditional_moments_packed(batch: PackedBatch, process_type: str, num_moments: int = 1, variance_method: str = "direct",
                               **process_params) -> Tuple[torch.Tensor, torch.Tensor]:
    """(moments, moments_before), each (N, d_x, num_moments) on the batch's device -- the packed form of what
    ``get_conditional_moments_at_obs`` returns (reference data_generation.py:819-921, same keyword defaults)."""
    if process_type not in _PROCESSES:
        raise ValueError(f"Unknown process type for conditional moments: {process_type}")
    if num_moments > 1 and variance_method not in ("direct", "second_moment"):
        raise ValueError(f"Unknown variance_method: {variance_method}")
    t, x = batch.times, batch.values                               # (N,), (N, d_x)
    N, dev = batch.N, batch.device
    pp = process_params
    first = torch.zeros(N, dtype=torch.bool, device=dev)
    first[batch.offsets[:-1][batch.offsets[:-1] < N]] = True       # first observation of every (non-empty) trajectory
    prev = torch.clamp(torch.arange(N, device=dev) - 1, min=0)     # previous observation (unused where `first`)
    mean = x.clone()
    var = torch.zeros_like(x)

    if process_type == "hybrid_ou_bs":
        sw = pp.get("switch_time")
        if sw is None:
            mean, mean_b, var_b = torch.zeros_like(x), torch.zeros_like(x), torch.zeros_like(x)
        else:
            # each regime is its own sequence: the previous observation OF THE SAME REGIME, regime starts count as first
            ou_side = t < sw
            same = ou_side == ou_side[prev]
            start = first | ~same                                   # (the OU prefix is contiguous, so "previous in regime" = prev)
            dt = (t - t[prev]).unsqueeze(-1)
            m_ou, v_ou = _ou(x[prev], dt, pp.get("theta_ou", 1.0), pp.get("mu_ou", 0.0), pp.get("sigma_ou", 0.3))
            m_bs, v_bs = _bs(x[prev], dt, pp.get("mu_bs", 0.0), pp.get("sigma_bs", 0.2))
            side = ou_side.unsqueeze(-1)
            mean_b = torch.where(start.unsqueeze(-1), x, torch.where(side, m_ou, m_bs))
            var_b = torch.where(start.unsqueeze(-1), torch.zeros_like(x), torch.where(side, v_ou, v_bs))
    else:
        dt = (t - t[prev]).unsqueeze(-1)
        if process_type == "ornstein_uhlenbeck":
            m, v = _ou(x[prev], dt, pp.get("theta", 1.0), pp.get("mu", 0.0), pp.get("sigma", 0.3))
        else:   # black_scholes, and heston through the Black-Scholes formulas with sigma := xi
            sigma = pp.get("xi", 0.5) if process_type == "heston" else pp.get("sigma", 0.2)
            m, v = _bs(x[prev], dt, pp.get("mu", 0.0), sigma)
        f = first.unsqueeze(-1)
        mean_b = torch.where(f, x, m)
        var_b = torch.where(f, torch.zeros_like(x), v)

    out = torch.zeros(N, x.shape[1], num_moments, dtype=x.dtype, device=dev)
    out_b = torch.zeros_like(out)
    out[..., 0], out_b[..., 0] = mean, mean_b
    if num_moments > 1:
        if variance_method == "direct":
            out[..., 1], out_b[..., 1] = var, var_b
        else:
            out[..., 1], out_b[..., 1] = var + mean ** 2, var_b + mean_b ** 2
    return out, out_b


def get_conditional_moments_at_obs(batch_times: List[torch.Tensor], batch_values: List[torch.Tensor], process_type: str,
                                   num_moments: int = 1, variance_method: str = "direct", **process_params):
    """The reference's list API (data_generation.py:819-921): two lists of (n_i, d_x, num_moments) tensors."""
    batch = PackedBatch.from_lists(batch_times, batch_values)
    m, mb = conditional_moments_packed(batch, process_type, num_moments, variance_method, **process_params)
    return batch.split(m), batch.split(mb)
