"""Golden vectors for the rows around the hot path, from the UNMODIFIED reference in the build container:

  cond_moments.npz   get_conditional_moments_at_obs (simulation/data_generation.py:819-921) for every process,
                     both variance methods, 1..3 moments, on one ragged batch (incl. a fixed-switch hybrid batch and
                     the random-switch case, which the reference answers with zeros)
  dense_grid_*.npz   the dense-grid model simulation inside plot_single_trajectory_with_condexp
                     (utils/plotting.py:133-256): the arrays it hands to matplotlib (model mean and, with two moments,
                     the +-2 sigma band) captured through a stub of matplotlib.pyplot (matplotlib is not installed)

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/aux/make_aux_golden.py
"""
import json
import os
import sys
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))          # tests/golden/aux
sys.dont_write_bytecode = True
plt = mock.MagicMock(name="matplotlib.pyplot")
sys.modules["matplotlib"] = mock.MagicMock(name="matplotlib")
sys.modules["matplotlib.pyplot"] = plt
sys.modules["matplotlib"].pyplot = plt
sys.path.insert(0, "/root/reference")
from neural_jump_ode.models.jump_ode import NeuralJumpODE          # noqa: E402
from neural_jump_ode.simulation import data_generation as dg       # noqa: E402
from neural_jump_ode.utils import plotting                         # noqa: E402


def cond_moments():
    g = torch.Generator().manual_seed(3)
    sizes = [1, 5, 9, 3, 12, 2]
    bt, bv = [], []
    for n in sizes:
        t = torch.sort(torch.rand(n, generator=g))[0]
        t[0] = 0.0
        bt.append(t)
        bv.append(0.6 + torch.rand(n, 1, generator=g))
    cases = {
        "black_scholes": dict(mu=0.1, sigma=0.5),
        "ornstein_uhlenbeck": dict(theta=1.5, mu=0.5, sigma=0.3),
        "heston": dict(mu=0.5, xi=0.5, kappa=2.0),
        "hybrid_ou_bs": dict(switch_time=0.45, theta_ou=1.0, mu_ou=0.5, sigma_ou=0.3, mu_bs=0.1, sigma_bs=0.3),
        "hybrid_ou_bs/random_switch": dict(switch_time=None, theta_ou=1.0, mu_ou=0.5),
    }
    out = {"times": torch.cat(bt).numpy(), "values": torch.cat(bv).numpy(), "sizes": np.array(sizes)}
    meta = {}
    for name, params in cases.items():
        proc = name.split("/")[0]
        for vm in ("direct", "second_moment"):
            for M in (1, 2, 3):
                m, mb = dg.get_conditional_moments_at_obs(bt, bv, proc, num_moments=M, variance_method=vm, **params)
                key = f"{name}|{vm}|{M}"
                out["m/" + key] = torch.cat(m).numpy()
                out["mb/" + key] = torch.cat(mb).numpy()
                meta[key] = dict(process=proc, params=params, variance_method=vm, num_moments=M)
    out["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "cond_moments.npz"), **out)
    print("cond_moments:", len(meta), "cases")


def dense_grid():
    cases = {
        "bs_sep_dt01": dict(model=dict(input_dim=1, hidden_dim=32, output_dim=1, dt_ode_step=0.01, num_moments=2),
                            process="black_scholes", params=dict(mu=0.1, sigma=0.5, T=1.0, n_steps=100, x0=1.0), obs=0.1, seed=123),
        "ou_shared_second_moment_dtnone": dict(
            model=dict(input_dim=1, hidden_dim=32, output_dim=1, dt_ode_step=None, num_moments=2, shared_network=True,
                       variance_method="second_moment"),
            process="ornstein_uhlenbeck", params=dict(theta=1.0, mu=0.5, sigma=0.3, T=1.0, n_steps=100, x0=0.0), obs=0.1, seed=7),
        "heston_h64_l2_tanh_dt0035": dict(
            model=dict(input_dim=1, hidden_dim=64, output_dim=1, dt_ode_step=0.0035, num_moments=1, n_hidden_layers=2,
                       activation="tanh", input_scaling="tanh"),
            process="heston", params=dict(mu=0.5, kappa=2.0, theta=0.04, xi=0.5, rho=-0.5, x0=1.0, v0=0.04, T=1.0, n_steps=200),
            obs=0.05, seed=11),
    }
    for name, c in cases.items():
        torch.manual_seed(0)
        model = NeuralJumpODE(**c["model"])
        plt.reset_mock()
        plotting.plot_single_trajectory_with_condexp(model, c["process"], c["params"], obs_fraction=c["obs"], seed=c["seed"])
        plots = [call.args for call in plt.plot.call_args_list]
        times_full, model_mean = np.asarray(plots[1][0]), np.asarray(plots[1][1])       # plt.plot(times, model_mean, 'r-')
        assert plots[1][2] == "r-"
        obs_t, obs_v = (np.asarray(a) for a in plt.scatter.call_args_list[0].args[:2])
        out = {"config_json": np.frombuffer(json.dumps(dict(model=c["model"], process=c["process"], params=c["params"],
                                                            obs_fraction=c["obs"], seed=c["seed"])).encode(), dtype=np.uint8),
               "grid": times_full.astype(np.float32), "obs_times": obs_t.astype(np.float32), "obs_values": obs_v.astype(np.float32),
               "mean": model_mean.astype(np.float32)}
        if c["model"].get("num_moments", 1) > 1:
            band = plt.fill_between.call_args_list[0].args            # (times, lower, upper): mean -+ 2 sqrt(max(var, 0))
            out["std"] = ((np.asarray(band[2]) - np.asarray(band[1])) / 4.0).astype(np.float32)
        for k, p in model.state_dict().items():
            out["param/" + k] = p.numpy()
        np.savez_compressed(os.path.join(HERE, f"dense_grid_{name}.npz"), **out)
        print(f"dense_grid_{name}: G={len(times_full)} n_obs={len(obs_t)} mean range [{model_mean.min():.4f}, {model_mean.max():.4f}]")


if __name__ == "__main__":
    cond_moments()
    dense_grid()
