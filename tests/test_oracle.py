"""Pin the oracle against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import rel_err, loss_close
from oracle import njode_oracle as orc


def _cfg(g):
    m = g["model"]
    return orc.make_cfg(m["input_dim"], m["hidden_dim"], m["output_dim"], m.get("dt_ode_step"),
                        m.get("num_moments", 1), m.get("n_hidden_layers", 1), m.get("activation", "relu"),
                        m.get("shared_network", False), m.get("input_scaling", "identity"))


def test_schedule_bit_exact(golden):
    """euler_schedule reproduces every (t_last, t_next) pair of every euler_step call, bit for bit."""
    cfg = _cfg(golden)
    got = []
    for t in golden["batch_times"]:
        t = t.numpy()
        for i in range(len(t) - 1):
            got += orc.euler_schedule(t[i], t[i + 1], cfg["dt_ode_step"])
    got = np.array(got, dtype=np.float32).reshape(-1, 2)
    ref = golden["step_log"]
    assert got.shape == ref.shape
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_port_matches_reference(golden):
    """The eager float32 port follows the reference op for op: tight agreement."""
    cfg = _cfg(golden)
    r = orc.run_port(golden["params"], cfg, golden["batch_times"], golden["batch_values"], golden["loss"])
    assert rel_err(torch.cat(r["preds"]), golden["preds"]) <= 2e-6
    assert rel_err(torch.cat(r["preds_before"]), golden["preds_before"]) <= 2e-6
    assert loss_close(r["loss"], golden["ref_loss"], 1e-6)
    for k, gref in golden["grads"].items():
        assert rel_err(r["grads"][k], gref) <= 1e-5, k
    log = np.array(r["step_log"], dtype=np.float32).reshape(-1, 2)
    assert np.array_equal(log.view(np.uint32), golden["step_log"].view(np.uint32))


def test_flat_f64_matches_reference(golden):
    """The interval-flattened float64 oracle agrees with the float32 reference to 1e-5
    (max-norm relative per tensor), loss to 1e-6; step counts are exact."""
    cfg = _cfg(golden)
    r = orc.run_flat(golden["params"], cfg, golden["batch_times"], golden["batch_values"], golden["loss"],
                     dtype=torch.float64)
    assert rel_err(r["preds"], golden["preds"]) <= 1e-5
    assert rel_err(r["preds_before"], golden["preds_before"]) <= 1e-5
    assert loss_close(r["loss"], golden["ref_loss"], 2e-6)
    for k, gref in golden["grads"].items():
        if not golden["has_grad"][k]:
            assert float(torch.nan_to_num(r["grads"][k]).abs().max()) == 0.0, k
        assert rel_err(r["grads"][k], gref) <= 1e-5, k
    assert int(r["K"].sum()) == len(golden["step_log"])
    first = golden["offsets"][:-1]
    assert float(r["preds_before"][first].abs().max()) == 0.0          # (a NaN observation never reaches a first row)


def test_flat_f32_close(golden):
    cfg = _cfg(golden)
    r = orc.run_flat(golden["params"], cfg, golden["batch_times"], golden["batch_values"], golden["loss"],
                     dtype=torch.float32)
    assert rel_err(r["preds"], golden["preds"]) <= 1e-5
    for k, gref in golden["grads"].items():
        assert rel_err(r["grads"][k], gref) <= 2e-5, k


def test_loss_rejects_unknown_variance_method():
    x = [torch.ones(2, 1)]
    y = [torch.ones(2, 1, 2)]
    with pytest.raises(ValueError):
        orc.loss_port(x, y, y, variance_method="nope")


def test_unknown_activation_is_relu():
    cfg = orc.make_cfg(1, 8, 1, 0.1, activation="identity")
    cfg2 = orc.make_cfg(1, 8, 1, 0.1, activation="relu")
    P = orc.init_params(cfg, seed=1)
    t = [torch.tensor([0.0, 0.35, 1.0])]
    v = [torch.tensor([[0.3], [0.1], [0.2]])]
    a, _ = orc.forward_port(P, cfg, t, v)
    b, _ = orc.forward_port(P, cfg2, t, v)
    assert torch.equal(a[0], b[0])


def test_paths_oracle_matches_reference():
    """oracle/paths_oracle.py (CPU baseline of the generator row) reproduces the reference's create_trajectory_batch
    bit for bit: same seeds, same RNG consumption (incl. the OU generator's unused draw), same float32 ops
    (fixture: tests/golden/aux/make_paths_golden.py)."""
    import json
    import os
    from conftest import GOLDEN_DIR
    from oracle import paths_oracle as po
    z = np.load(os.path.join(GOLDEN_DIR, "aux", "paths_ref.npz"))
    meta = json.loads(bytes(z["meta_json"]).decode())
    assert set(m["process"] for m in meta.values()) == {"black_scholes", "ornstein_uhlenbeck", "heston"}
    for name, m in meta.items():
        bt, bv = po.trajectory_batch(m["n_traj"], m["process"], **m["kwargs"])
        assert [len(t) for t in bt] == list(z[f"{name}|sizes"]), name
        assert all(v.shape == (len(t), 1) for t, v in zip(bt, bv)), name
        assert np.array_equal(torch.cat(bt).numpy().view(np.uint32), z[f"{name}|times"].view(np.uint32)), name
        assert np.array_equal(torch.cat(bv).numpy().view(np.uint32), z[f"{name}|values"].view(np.uint32)), name
    with pytest.raises(ValueError):
        po.trajectory_batch(1, "no_such_process")
