"""TEST INFRASTRUCTURE -- CPU restatement of the reference's per-trajectory path generators and observation sampler
(SURVEY.md section 8f, row N2).  Only tests/ and bench.py's CPU-baseline legs (bench.rows_cpu_baselines) may import this.

Follows neural_jump_ode/simulation/data_generation.py of the reference, one trajectory at a time and with the
reference's consumption of the global torch / numpy RNG streams, so that the same seed gives the same path BIT FOR BIT:

* ``black_scholes``       data_generation.py:23-43   one randn(n) draw, cumulative sum in log space
* ``ornstein_uhlenbeck``  data_generation.py:60-91   TWO randn(n) draws (the first is scaled by sqrt(dt) and never used),
                                                    exact one-step transition, sequential recursion
* ``heston``              data_generation.py:182-216 z1, z2 = randn(n), randn(n); full-truncation Euler, variance floored
                                                    at 1e-6 both when read and when written
* observation rule        data_generation.py:226-249 first and last grid point + np.random.choice of the interior
                                                    without replacement, n_obs = max(2, int(fraction * n_grid))
* batch builder           data_generation.py:268-289 trajectory i uses seed i for the path AND for the sampler

PINNED: tests/golden/aux/paths_ref.npz holds paths and observation sets produced by the unmodified reference in the build
container (tests/golden/aux/make_paths_golden.py); tests/test_oracle.py::test_paths_oracle_matches_reference compares
bit for bit.  The hybrid OU -> BS generator is not restated here (the device generator's hybrid paths are checked
against the reference's sample statistics, tests/golden/generator_stats.json).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

PROCESS_DEFAULTS = {
    "black_scholes": dict(mu=0.0, sigma=0.2, x0=1.0),
    "ornstein_uhlenbeck": dict(theta=1.0, mu=0.0, sigma=0.3, x0=0.0),
    "heston": dict(mu=0.0, kappa=2.0, theta=0.04, xi=0.5, rho=-0.5, x0=1.0, v0=0.04),
}


def _grid(T: float, n_steps: int) -> Tuple[float, torch.Tensor]:
    return T / n_steps, torch.linspace(0.0, T, n_steps + 1)


def path_black_scholes(seed: int, T: float, n_steps: int, mu: float, sigma: float, x0: float):
    torch.manual_seed(seed)
    dt, times = _grid(T, n_steps)
    dW = torch.randn(n_steps) * torch.sqrt(torch.tensor(dt))
    steps = (mu - 0.5 * sigma ** 2) * dt + sigma * dW
    logx = torch.zeros(n_steps + 1)
    logx[0] = torch.log(torch.tensor(x0))
    logx[1:] = logx[0] + torch.cumsum(steps, dim=0)
    return times, torch.exp(logx)


def path_ou(seed: int, T: float, n_steps: int, theta: float, mu: float, sigma: float, x0: float):
    torch.manual_seed(seed)
    dt, times = _grid(T, n_steps)
    torch.randn(n_steps)                                  # the reference's unused dW draw: keeps the stream aligned
    th, m, s = torch.tensor(theta), torch.tensor(mu), torch.tensor(sigma)
    decay = torch.exp(-th * dt)
    pull = m * (1 - decay)
    amp = s * torch.sqrt((1 - torch.exp(-2 * th * dt)) / (2 * th)) if theta > 0 else s * torch.sqrt(dt)
    noise = amp * torch.randn(n_steps)
    X = torch.zeros(n_steps + 1)
    X[0] = x0
    for k in range(n_steps):
        X[k + 1] = X[k] * decay + pull + noise[k]
    return times, X


def path_heston(seed: int, T: float, n_steps: int, mu: float, kappa: float, theta: float, xi: float, rho: float,
                x0: float, v0: float):
    torch.manual_seed(seed)
    dt, times = _grid(T, n_steps)
    z1, z2 = torch.randn(n_steps), torch.randn(n_steps)
    root_dt = torch.sqrt(torch.tensor(dt))
    dW1 = root_dt * z1
    dW2 = root_dt * (rho * z1 + torch.sqrt(torch.tensor(1 - rho ** 2)) * z2)
    X, V = torch.zeros(n_steps + 1), torch.zeros(n_steps + 1)
    X[0], V[0] = x0, v0
    for k in range(n_steps):
        vol = torch.sqrt(torch.clamp(V[k], min=1e-6))
        X[k + 1] = X[k] + mu * X[k] * dt + vol * X[k] * dW1[k]
        V[k + 1] = torch.clamp(V[k] + kappa * (theta - V[k]) * dt + xi * vol * dW2[k], min=1e-6)
    return times, X


_PATHS = {"black_scholes": path_black_scholes, "ornstein_uhlenbeck": path_ou, "heston": path_heston}


def observe(times: torch.Tensor, values: torch.Tensor, obs_fraction: float, seed: int):
    torch.manual_seed(seed)
    np.random.seed(seed)
    n_grid = len(times)
    n_obs = max(2, int(obs_fraction * n_grid))
    picked = [0, n_grid - 1]
    if n_obs > 2:
        interior = list(range(1, n_grid - 1))
        picked.extend(np.random.choice(interior, min(n_obs - 2, len(interior)), replace=False))
    idx = torch.tensor(sorted(set(picked)), dtype=torch.long)
    return times[idx], values[idx]


def trajectory_batch(n_trajectories: int, process_type: str, obs_fraction: float = 0.1, T: float = 1.0, n_steps: int = 100,
                     **params) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
    """``create_trajectory_batch`` of the reference: lists of (n_i,) observation times and (n_i, 1) values."""
    if process_type not in _PATHS:
        raise ValueError(f"Unknown process type: {process_type}")
    kw = dict(PROCESS_DEFAULTS[process_type])
    kw.update(params)
    bt, bv = [], []
    for i in range(n_trajectories):
        times, X = _PATHS[process_type](i, T, n_steps, **kw)
        t_obs, x_obs = observe(times, X, obs_fraction, i)
        bt.append(t_obs)
        bv.append(x_obs.unsqueeze(-1))
    return bt, bv
