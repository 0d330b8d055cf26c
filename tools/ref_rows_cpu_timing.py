"""Times the UNMODIFIED reference on the rows AROUND the hot path in the BUILD CONTAINER (the
reference does not exist on the GPU box) and writes profiles/r2_reference_rows_build_container.json.  tools/bench_rows.py
prints these beside its GPU numbers as labelled context -- they are this container's CPU, not the GPU box's.

    PYTHONDONTWRITEBYTECODE=1 python tools/ref_rows_cpu_timing.py
"""
import json
import os
import sys
import time
from unittest import mock

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.dont_write_bytecode = True
plt = mock.MagicMock(name="matplotlib.pyplot")
sys.modules["matplotlib"] = mock.MagicMock(name="matplotlib")
sys.modules["matplotlib.pyplot"] = plt
sys.modules["matplotlib"].pyplot = plt
sys.path.insert(0, "/root/reference")
from neural_jump_ode.models.jump_ode import NeuralJumpODE, nj_ode_loss   # noqa: E402
from neural_jump_ode.simulation import data_generation as dg             # noqa: E402
from neural_jump_ode.utils import plotting                               # noqa: E402

HESTON = dict(mu=0.5, kappa=2.0, theta=0.04, xi=0.5, rho=-0.5, x0=1.0, v0=0.04)
BS = dict(mu=0.1, sigma=0.5, x0=1.0)


def wall(fn, repeats=1):
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        r = fn()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, r


def main():
    out = {"where": "build container CPU (NOT the GPU box)", "threads": torch.get_num_threads(), "cpu_count": os.cpu_count(),
           "torch": torch.__version__}

    # N1: one training step of config 1 as Trainer.train_epoch runs it (utils/training.py:78-101): 128 trajectories,
    # zero_grad, forward, loss, backward, Adam(weight_decay=5e-4) step, loss.item()
    bt, bv = dg.create_trajectory_batch(128, "black_scholes", obs_fraction=0.1, T=1.0, n_steps=100, **BS)
    torch.manual_seed(0)
    model = NeuralJumpODE(1, 32, 1, dt_ode_step=0.01, num_moments=2)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=5e-4)
    mw = torch.tensor([1.0, 10.0])

    def step():
        opt.zero_grad()
        p, pb = model(bt, bv)
        loss = nj_ode_loss(bt, bv, p, pb, ignore_first_continuity=True, moment_weights=mw)
        loss.backward()
        opt.step()
        return loss.item()
    step()
    sec, _ = wall(step, 2)
    out["N1_training_step_config1_b128"] = dict(seconds=sec, trajectories=128, trajectories_per_s=128 / sec)

    # N2: path generation + observation sampling, per trajectory (create_trajectory_batch)
    gen = {}
    for name, proc, n, kw in (("black_scholes_n100", "black_scholes", 64, dict(T=1.0, n_steps=100, **BS)),
                              ("ornstein_uhlenbeck_n100", "ornstein_uhlenbeck", 64, dict(T=1.0, n_steps=100, theta=1.0, mu=0.5, sigma=0.3, x0=0.0)),
                              ("heston_n200", "heston", 32, dict(T=1.0, n_steps=200, **HESTON))):
        sec, _ = wall(lambda: dg.create_trajectory_batch(n, proc, obs_fraction=0.1, **kw))
        gen[name] = dict(trajectories=n, seconds=sec, trajectories_per_s=n / sec, ms_per_trajectory=1e3 * sec / n)
    out["N2_generators"] = gen

    # N3: closed-form conditional moments at the observations (data_generation.py:819-921), Heston, 2 moments
    hbt, hbv = dg.create_trajectory_batch(64, "heston", obs_fraction=0.1, T=1.0, n_steps=200, **HESTON)
    n_obs = sum(len(t) for t in hbt)
    sec, _ = wall(lambda: dg.get_conditional_moments_at_obs(hbt, hbv, "heston", num_moments=2, variance_method="direct",
                                                            mu=0.5, xi=0.5, kappa=2.0), 3)
    out["N3_conditional_moments_heston"] = dict(observations=n_obs, seconds=sec, observations_per_s=n_obs / sec)

    # N4: the dense-grid model simulation of plot_single_trajectory_with_condexp (utils/plotting.py:133-256), one
    # trajectory on its 101-point grid (matplotlib stubbed; the time includes the path generation and the closed-form
    # conditional expectation the function also computes)
    params = dict(T=1.0, n_steps=100, **BS)
    sec, _ = wall(lambda: plotting.plot_single_trajectory_with_condexp(model, "black_scholes", params, obs_fraction=0.1, seed=123), 3)
    out["N4_dense_grid_one_trajectory_bs_n100"] = dict(grid_points=101, seconds=sec, grid_points_per_s=101 / sec)

    path = os.path.join(ROOT, "profiles", "r2_reference_rows_build_container.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
