"""Rows N3 / N4 of SURVEY.md section 8f against vectors produced by the unmodified reference
(tests/golden/aux/make_aux_golden.py): closed-form conditional moments at the observations (CPU and device), the
relative-loss metric, and dense-grid inference."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, rel_err

AUX = os.path.join(GOLDEN_DIR, "aux")
DEV = "cuda:0"


def _cond_cases():
    z = np.load(os.path.join(AUX, "cond_moments.npz"))
    return z, json.loads(bytes(z["meta_json"]).decode())


def _check_cond_moments(device):
    from neural_jump_ode import PackedBatch
    from neural_jump_ode.simulation import conditional_moments_packed, get_conditional_moments_at_obs
    z, meta = _cond_cases()
    sizes = z["sizes"].tolist()
    off = np.concatenate([[0], np.cumsum(sizes)])
    bt = [torch.from_numpy(z["times"][off[i]:off[i + 1]].copy()) for i in range(len(sizes))]
    bv = [torch.from_numpy(z["values"][off[i]:off[i + 1]].copy()) for i in range(len(sizes))]
    batch = PackedBatch.from_lists(bt, bv, device=device)
    for key, m in meta.items():
        got, got_b = conditional_moments_packed(batch, m["process"], num_moments=m["num_moments"],
                                                variance_method=m["variance_method"], **m["params"])
        assert got.shape == (sum(sizes), 1, m["num_moments"])
        for a, ref in ((got, z["m/" + key]), (got_b, z["mb/" + key])):
            ref = torch.from_numpy(ref)
            assert float((a.cpu() - ref).abs().max()) <= 2e-6 * max(1.0, float(ref.abs().max())), key
    m_list, mb_list = get_conditional_moments_at_obs(bt, bv, "black_scholes", num_moments=2, mu=0.1, sigma=0.5)
    assert [tuple(t.shape) for t in m_list] == [(n, 1, 2) for n in sizes] and len(mb_list) == len(sizes)
    with pytest.raises(ValueError):
        conditional_moments_packed(batch, "brownian")
    with pytest.raises(ValueError):
        conditional_moments_packed(batch, "black_scholes", num_moments=2, variance_method="nope")


def test_conditional_moments_match_reference_cpu():
    _check_cond_moments("cpu")


@pytest.mark.gpu
def test_conditional_moments_match_reference_device():
    _check_cond_moments(DEV)


@pytest.mark.gpu
def test_relative_loss_metric():
    """training.py:219-261 on the device: loss of the model, loss of the closed-form moments (tensors that did not come
    from the model, through the same nj_ode_loss), their relative difference -- against the CPU port of the loss."""
    from neural_jump_ode import NeuralJumpODE, PackedBatch, relative_loss_packed, nj_ode_loss
    from neural_jump_ode.simulation import make_packed_batch, conditional_moments_packed
    from oracle import njode_oracle as orc
    params = dict(mu=0.1, sigma=0.5, x0=1.0)
    data = make_packed_batch("black_scholes", 10, 0.1, device=DEV, seed=3, **params)
    torch.manual_seed(0)
    model = NeuralJumpODE(1, 32, 1, dt_ode_step=0.01, num_moments=2).to(DEV)
    mw = torch.tensor([1.0, 10.0], device=DEV)
    rel = relative_loss_packed(model, data, "black_scholes", params, moment_weights=mw)
    with torch.no_grad():
        p, b = model.forward_packed(data)
        true, true_b = conditional_moments_packed(data, "black_scholes", num_moments=2, **params)
    bv = [v.cpu() for v in data.split(data.values)]
    l_model = float(orc.loss_port(bv, [t.cpu() for t in data.split(p)], [t.cpu() for t in data.split(b)], moment_weights=[1.0, 10.0]))
    l_true = float(orc.loss_port(bv, [t.cpu() for t in data.split(true)], [t.cpu() for t in data.split(true_b)], moment_weights=[1.0, 10.0]))
    want = (l_model - l_true) / max(l_true, 1e-8)
    assert abs(rel - want) <= 1e-5 * max(1.0, abs(want))
    assert l_true > 0.0 and np.isfinite(rel)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["bs_sep_dt01", "ou_shared_second_moment_dtnone", "heston_h64_l2_tanh_dt0035"])
def test_dense_grid_inference_matches_reference_plots(name):
    """predict_on_grid against what the unmodified plot_single_trajectory_with_condexp (utils/plotting.py:133-256)
    handed to matplotlib: model mean on the full grid (1e-5 max-norm relative) and the standard deviation band."""
    from neural_jump_ode import NeuralJumpODE
    z = np.load(os.path.join(AUX, f"dense_grid_{name}.npz"))
    conf = json.loads(bytes(z["config_json"]).decode())
    model = NeuralJumpODE(**conf["model"])
    model.load_state_dict({k[len("param/"):]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith("param/")})
    model = model.to(DEV)
    t = torch.from_numpy(z["obs_times"].copy())
    v = torch.from_numpy(z["obs_values"].copy()).view(-1, 1)
    grid = torch.from_numpy(z["grid"].copy())
    dense = model.predict_on_grid([t, t[:4]], grid, [v, v[:4]])               # a second, shorter trajectory on the same grid
    assert dense.shape == (2, grid.shape[0], 1, conf["model"].get("num_moments", 1))
    mean, var = model.grid_moments(dense)
    assert rel_err(mean[0, :, 0].cpu(), z["mean"]) <= 1e-5
    if "std" in z.files:
        assert rel_err(torch.sqrt(torch.clamp(var[0, :, 0], min=0.0)).cpu(), z["std"]) <= 2e-5
    # the short trajectory agrees with the long one up to its last observation (same observations, same grid) ...
    g_last = int((grid < t[3]).sum())
    assert torch.equal(dense[1, :g_last], dense[0, :g_last])
    # ... and grid times before the first observation would be 0
    late = model.predict_on_grid([t[2:]], grid, [v[2:]])
    assert float(late[0, : int((grid < t[2]).sum())].abs().max()) == 0.0
