"""Batched SDE path simulation and observation sub-sampling on the device.

The reference simulates one trajectory at a time in Python (0.4-15 ms per trajectory,
data_generation.py:11-218), which cannot feed the 262 144 / 1 M trajectory configurations.  These
generators follow the same recurrences, vectorised over the batch, and emit a ``PackedBatch``:

* black_scholes   log-Euler, data_generation.py:26-43
* ornstein_uhlenbeck   exact one-step transition, data_generation.py:76-91
* heston          full-truncation Euler with correlated increments, data_generation.py:186-216
* hybrid_ou_bs    OU up to a per-path switch time (U[0.2T, 0.8T] unless given), then Black-Scholes in log space from
                  the value reached, data_generation.py:130-160 (a path whose OU leg is <= 0 at the switch is NaN in
                  the reference, ~1 in 2000; here such paths are re-drawn from the valid ones so batches stay finite)
* observation rule: first and last grid point always observed plus a uniform random interior subset,
  n_obs = max(2, int(obs_fraction * n_grid)) (data_generation.py:235-249)

They use their own RNG stream (a ``torch.Generator``), so paths are statistically, not bitwise, the
reference's; parity tests use the reference generator's fixtures instead.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from ..packed import PackedBatch

_DEFAULTS = {
    "black_scholes": dict(mu=0.0, sigma=0.2, x0=1.0),
    "ornstein_uhlenbeck": dict(theta=1.0, mu=0.0, sigma=0.3, x0=0.0),
    "heston": dict(mu=0.0, kappa=2.0, theta=0.04, xi=0.5, rho=-0.5, x0=1.0, v0=0.04),
    "hybrid_ou_bs": dict(theta_ou=1.0, mu_ou=0.0, sigma_ou=0.3, mu_bs=0.0, sigma_bs=0.2, x0=1.0, switch_time=None),
}


def simulate_paths(process_type: str, n_paths: int, n_steps: int = 100, T: float = 1.0, device="cuda",
                   generator: Optional[torch.Generator] = None, **kw) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (grid_times (n_steps+1,), X (n_paths, n_steps+1)) float32 on ``device``."""
    if process_type not in _DEFAULTS:
        raise ValueError(f"Unknown process type: {process_type}. Supported: {', '.join(_DEFAULTS)}")
    p = dict(_DEFAULTS[process_type])
    unknown = set(kw) - set(p)
    if unknown:
        raise TypeError(f"unexpected parameters for {process_type}: {sorted(unknown)}")
    p.update(kw)
    dt = T / n_steps
    times = torch.linspace(0.0, T, n_steps + 1, device=device)

    def randn():
        return torch.randn(n_paths, n_steps, device=device, generator=generator)

    if process_type == "black_scholes":
        incr = (p["mu"] - 0.5 * p["sigma"] ** 2) * dt + p["sigma"] * math.sqrt(dt) * randn()
        logx = math.log(p["x0"]) + torch.cumsum(incr, dim=1)
        X = torch.cat([torch.full((n_paths, 1), float(p["x0"]), device=device), torch.exp(logx)], dim=1)
    elif process_type == "ornstein_uhlenbeck":
        th = p["theta"]
        decay = math.exp(-th * dt)
        shift = p["mu"] * (1.0 - decay)
        amp = p["sigma"] * (math.sqrt((1.0 - math.exp(-2.0 * th * dt)) / (2.0 * th)) if th > 0 else math.sqrt(dt))
        noise = amp * randn()
        cols = [torch.full((n_paths,), float(p["x0"]), device=device)]
        for i in range(n_steps):
            cols.append(cols[-1] * decay + shift + noise[:, i])
        X = torch.stack(cols, dim=1)
    elif process_type == "hybrid_ou_bs":
        th = p["theta_ou"]
        decay = math.exp(-th * dt)
        shift = p["mu_ou"] * (1.0 - decay)
        amp = p["sigma_ou"] * (math.sqrt((1.0 - math.exp(-2.0 * th * dt)) / (2.0 * th)) if th > 0 else math.sqrt(dt))
        noise = amp * randn()
        cols = [torch.full((n_paths,), float(p["x0"]), device=device)]
        for i in range(n_steps):
            cols.append(cols[-1] * decay + shift + noise[:, i])
        ou = torch.stack(cols, dim=1)                                         # OU leg on the whole grid
        if p["switch_time"] is None:
            sw = 0.2 * T + 0.6 * T * torch.rand(n_paths, device=device, generator=generator)
        else:
            sw = torch.full((n_paths,), float(p["switch_time"]), device=device)
        sidx = torch.clamp((sw / dt).long(), max=n_steps)                     # int(switch_time / dt)
        incr = (p["mu_bs"] - 0.5 * p["sigma_bs"] ** 2) * dt + p["sigma_bs"] * math.sqrt(dt) * randn()
        csum = torch.cat([torch.zeros(n_paths, 1, device=device), torch.cumsum(incr, dim=1)], dim=1)   # (n, n_steps+1)
        x_sw = torch.gather(ou, 1, sidx[:, None])                             # value at the switch
        c_sw = torch.gather(csum, 1, sidx[:, None])
        bs = torch.exp(torch.log(x_sw) + csum - c_sw)                         # log-space BS leg (NaN if x_sw <= 0)
        grid = torch.arange(n_steps + 1, device=device)[None, :]
        X = torch.where(grid <= sidx[:, None], ou, bs)
        bad = ~torch.isfinite(X).all(dim=1)
        if bool(bad.any()) and not bool(bad.all()):
            good = torch.nonzero(~bad).flatten()
            X[bad] = X[good[torch.arange(int(bad.sum()), device=device) % good.numel()]]
    else:  # heston
        z1, z2 = randn(), randn()
        sq = math.sqrt(dt)
        dw1 = sq * z1
        dw2 = sq * (p["rho"] * z1 + math.sqrt(1.0 - p["rho"] ** 2) * z2)
        x = torch.full((n_paths,), float(p["x0"]), device=device)
        v = torch.full((n_paths,), float(p["v0"]), device=device)
        cols = [x]
        for i in range(n_steps):
            sv = torch.sqrt(torch.clamp(v, min=1e-6))
            x = x + p["mu"] * x * dt + sv * x * dw1[:, i]
            v = torch.clamp(v + p["kappa"] * (p["theta"] - v) * dt + p["xi"] * sv * dw2[:, i], min=1e-6)
            cols.append(x)
        X = torch.stack(cols, dim=1)
    return times, X.float()


def sample_observations(times: torch.Tensor, X: torch.Tensor, obs_fraction: float = 0.1,
                        generator: Optional[torch.Generator] = None) -> PackedBatch:
    """Equal observation count per path: first + last grid point and a uniform interior subset."""
    B, n_grid = X.shape
    n_obs = max(2, int(obs_fraction * n_grid))
    n_int = min(n_obs - 2, n_grid - 2)
    dev = X.device
    if n_int > 0:
        scores = torch.rand(B, n_grid - 2, device=dev, generator=generator)
        interior = torch.topk(scores, n_int, dim=1).indices + 1
        idx = torch.cat([torch.zeros(B, 1, dtype=torch.long, device=dev), interior,
                         torch.full((B, 1), n_grid - 1, dtype=torch.long, device=dev)], dim=1)
        idx = torch.sort(idx, dim=1).values
    else:
        idx = torch.tensor([[0, n_grid - 1]], device=dev).expand(B, 2)
    n = idx.shape[1]
    t = times[idx].reshape(-1)
    v = torch.gather(X, 1, idx).reshape(-1, 1)
    off = torch.arange(0, (B + 1) * n, n, dtype=torch.int64, device=dev)
    return PackedBatch(t, v, off, sizes=[n] * B)


def sample_observations_ragged(times: torch.Tensor, X: torch.Tensor, frac_lo: float, frac_hi: float,
                               generator: Optional[torch.Generator] = None) -> PackedBatch:
    """Per-path observation fraction U[frac_lo, frac_hi] (BASELINE config 5): ragged n_obs."""
    B, n_grid = X.shape
    dev = X.device
    frac = frac_lo + (frac_hi - frac_lo) * torch.rand(B, device=dev, generator=generator)
    n_obs = torch.clamp((frac * n_grid).long(), min=2, max=n_grid)
    scores = torch.rand(B, n_grid, device=dev, generator=generator)
    scores[:, 0] = 2.0           # first and last grid point are always observed
    scores[:, -1] = 2.0
    rank = torch.argsort(torch.argsort(scores, dim=1, descending=True), dim=1)
    keep = rank < n_obs[:, None]
    off = torch.zeros(B + 1, dtype=torch.int64, device=dev)
    off[1:] = torch.cumsum(keep.sum(1), 0)
    t = times.expand(B, n_grid)[keep]
    v = X[keep].reshape(-1, 1)
    return PackedBatch(t, v, off)


def make_packed_batch(process_type: str, n_paths: int, obs_fraction: float = 0.1, n_steps: int = 100,
                      T: float = 1.0, device="cuda", seed: int = 0, **process_kwargs) -> PackedBatch:
    """simulate + sub-sample in one call, deterministic in ``seed``."""
    g = torch.Generator(device=device).manual_seed(seed)
    times, X = simulate_paths(process_type, n_paths, n_steps=n_steps, T=T, device=device, generator=g,
                              **process_kwargs)
    return sample_observations(times, X, obs_fraction, generator=g)


def concat_batches(batches) -> PackedBatch:
    """Concatenate packed batches (same device, same d_x) into one."""
    batches = list(batches)
    times = torch.cat([b.times for b in batches])
    values = torch.cat([b.values for b in batches])
    offs, base = [batches[0].offsets[:1]], 0
    for b in batches:
        offs.append(b.offsets[1:] + base)
        base += b.N
    return PackedBatch(times, values, torch.cat(offs))


def make_mixed_ragged_batch(n_paths: int, frac_lo: float = 0.02, frac_hi: float = 0.2, n_steps: int = 100,
                            T: float = 1.0, device="cuda", seed: int = 0) -> PackedBatch:
    """BASELINE config 5: equal parts Black-Scholes / OU / Heston / hybrid OU->BS paths (experiment_hybrid.py-style
    mixed batch, process parameters of the experiment scripts) with a per-path observation fraction U[frac_lo, frac_hi]."""
    g = torch.Generator(device=device).manual_seed(seed)
    procs = [("black_scholes", dict(mu=0.1, sigma=0.5, x0=1.0)),
             ("ornstein_uhlenbeck", dict(theta=1.0, mu=0.5, sigma=0.3, x0=0.0)),
             ("heston", dict(mu=0.5, kappa=2.0, theta=0.04, xi=0.5, rho=-0.5, x0=1.0, v0=0.04)),
             ("hybrid_ou_bs", dict(theta_ou=1.0, mu_ou=0.5, sigma_ou=0.3, mu_bs=0.1, sigma_bs=0.3, x0=1.0))]
    parts, left = [], n_paths
    for i, (name, kw) in enumerate(procs):
        n = left if i == len(procs) - 1 else n_paths // len(procs)
        left -= n
        if n <= 0:
            continue
        times, X = simulate_paths(name, n, n_steps=n_steps, T=T, device=device, generator=g, **kw)
        parts.append(sample_observations_ragged(times, X, frac_lo, frac_hi, generator=g))
    return concat_batches(parts)
