"""What sits around the hot path on a B200: the reference's training step replayed against
the drop-in, the packed epoch loop, FlatAdam checkpoints, batches in waves, the on-device generators, and the NCCL
gradient all-reduce.  `pytest -m gpu`."""
import os
import random

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import njode_oracle as orc

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
TOL = 1e-5
MK = dict(input_dim=1, hidden_dim=32, output_dim=1, dt_ode_step=0.01, num_moments=2)
LK = dict(ignore_first_continuity=True, moment_weights=[1.0, 10.0], variance_method="direct")


def _cfg(mk):
    return orc.make_cfg(mk["input_dim"], mk["hidden_dim"], mk["output_dim"], mk.get("dt_ode_step"),
                        mk.get("num_moments", 1), mk.get("n_hidden_layers", 1), mk.get("activation", "relu"),
                        mk.get("shared_network", False), mk.get("input_scaling", "identity"))


def _bs_lists(n, seed=0, n_obs=10, n_steps=100):
    """Black-Scholes-like paths observed at n_obs grid points (first and last always): the shape of BASELINE config 1."""
    rng = np.random.RandomState(seed)
    bt, bv = [], []
    for _ in range(n):
        idx = np.sort(np.concatenate([[0, n_steps], rng.choice(np.arange(1, n_steps), n_obs - 2, replace=False)]))
        logx = np.concatenate([[0.0], np.cumsum((0.1 - 0.125) / n_steps + 0.5 / np.sqrt(n_steps) * rng.randn(n_steps))])
        bt.append(torch.linspace(0.0, 1.0, n_steps + 1)[torch.from_numpy(idx)])
        bv.append(torch.from_numpy(np.exp(logx[idx]).astype(np.float32)).view(-1, 1))
    return bt, bv


def _oracle_training(P0, cfg, batches, lr, wd, lk):
    """The reference's optimisation (Adam(lr, weight_decay), training.py:396) in float64 on the CPU: loss per step."""
    P = {k: v.double().clone().requires_grad_(True) for k, v in P0.items()}
    opt = torch.optim.Adam(list(P.values()), lr=lr, weight_decay=wd)
    losses = []
    for bt, bv in batches:
        r = orc.run_flat({k: v.detach() for k, v in P.items()}, cfg, bt, bv, lk, dtype=torch.float64)
        opt.zero_grad()
        for k, v in P.items():
            v.grad = r["grads"][k].double().clone()
        opt.step()
        losses.append(float(r["loss"]))
    return losses, {k: v.detach() for k, v in P.items()}


def test_reference_training_step_replay_and_packed_epoch():
    """The call pattern of the reference Trainer (utils/training.py:78-101, :116-122, :396 and the reload at
    experiments/experiment_heston.py:152-168), written against the public API: mini-batches of 128 with a 104 tail,
    per-element .to(device), zero_grad, Adam(weight_decay=5e-4), a device-tensor moment_weights, loss.item() per step,
    validation under no_grad on the whole set, state_dict reload into a fresh model.  The loss trajectory over 6
    optimiser steps equals the float64 oracle's; the packed epoch loop (device-resident dataset, device-side gathers,
    FlatAdam, one loss read per epoch) walks the same trajectory."""
    from neural_jump_ode import NeuralJumpODE, nj_ode_loss, FlatAdam, PackedBatch, train_epoch_packed, validate_packed
    n_train, batch_size, epochs, lr, wd = 232, 128, 3, 1e-3, 5e-4
    bt, bv = _bs_lists(n_train, seed=5)
    order_rng = random.Random(123)
    orders = []
    for _ in range(epochs):
        idx = list(range(n_train))
        order_rng.shuffle(idx)                                                  # training.py:55-56
        orders.append(idx)

    torch.manual_seed(0)
    model = NeuralJumpODE(**MK)
    P0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to(DEV)
    opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=wd)          # training.py:396
    mw = torch.tensor(LK["moment_weights"], device=DEV)                         # training.py:24
    losses, batches = [], []
    for idx in orders:
        model.train()
        for lo in range(0, n_train, batch_size):                               # training.py:78-101
            sel = idx[lo:lo + batch_size]
            mb_t = [bt[i].to(DEV) for i in sel]
            mb_v = [bv[i].to(DEV) for i in sel]
            opt.zero_grad()
            preds, before = model(mb_t, mb_v)
            loss = nj_ode_loss(mb_t, mb_v, preds, before, ignore_first_continuity=True, moment_weights=mw,
                               variance_method=model.variance_method)
            loss.backward()
            opt.step()
            losses.append(loss.item())
            batches.append(([bt[i] for i in sel], [bv[i] for i in sel]))
    assert [len(b[0]) for b in batches[:2]] == [128, 104]
    ref_losses, P_ref = _oracle_training(P0, _cfg(MK), batches, lr, wd, LK)
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= 2e-5 * abs(b), (losses, ref_losses)
    for k, p in model.named_parameters():
        assert rel_err(p.detach().cpu(), P_ref[k]) <= 2e-4, k                   # (Adam normalises: tiny gradient entries amplify)

    # validation (training.py:116-122) and the reload of experiment_heston.py:152-168
    model.eval()
    with torch.no_grad():
        vt, vv = [t.to(DEV) for t in bt[:50]], [v.to(DEV) for v in bv[:50]]
        vp, vb = model(vt, vv)
        val = nj_ode_loss(vt, vv, vp, vb, ignore_first_continuity=True, moment_weights=mw).item()
    fresh = NeuralJumpODE(**MK).to(DEV)
    fresh.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
    fresh.eval()
    with torch.no_grad():
        fp, fb = fresh(vt, vv)
    assert torch.equal(fp.packed, vp.packed) and torch.equal(fb.packed, vb.packed)

    # the packed epoch loop on a device-resident dataset: same mini-batches, same optimiser semantics
    torch.manual_seed(0)
    m2 = NeuralJumpODE(**MK).to(DEV)
    opt2 = FlatAdam(m2.parameters(), lr=lr, weight_decay=wd)
    data = PackedBatch.from_lists(bt, bv, device=DEV)
    means = [train_epoch_packed(m2, opt2, data, batch_size, ignore_first_continuity=True, moment_weights=mw, order=o)
             for o in orders]
    for e, m in enumerate(means):
        want = float(np.mean(ref_losses[2 * e:2 * e + 2]))
        assert abs(m - want) <= 2e-5 * abs(want)
    for (k, p), (_, q) in zip(model.named_parameters(), m2.named_parameters()):
        assert rel_err(q.detach().cpu(), p.detach().cpu()) <= 2e-4, k
    v2 = validate_packed(m2, PackedBatch.from_lists(bt[:50], bv[:50], device=DEV), ignore_first_continuity=True, moment_weights=mw)
    assert abs(v2 - val) <= 1e-4 * abs(val)


def test_flat_adam_checkpoint_round_trips_with_torch_adam():
    """optimizer_state_dict of the reference Trainer (training.py:152-154, :291-304): a run resumed from a checkpoint
    continues exactly -- FlatAdam -> FlatAdam, torch.optim.Adam -> FlatAdam and FlatAdam -> torch.optim.Adam."""
    from neural_jump_ode import NeuralJumpODE, nj_ode_loss, FlatAdam, PackedBatch
    bt, bv = _bs_lists(64, seed=9)
    data = PackedBatch.from_lists(bt, bv, device=DEV)

    def make(kind):
        torch.manual_seed(3)
        m = NeuralJumpODE(**MK).to(DEV)
        o = (FlatAdam if kind == "flat" else torch.optim.Adam)(m.parameters(), lr=1e-2, weight_decay=5e-4)
        return m, o

    def steps(m, o, n):
        out = []
        for _ in range(n):
            o.zero_grad()
            p, b = m.forward_packed(data)
            loss = nj_ode_loss(data, None, p, b, **LK)
            loss.backward()
            o.step()
            out.append(loss.item())
        return out

    straight_m, straight_o = make("torch")
    straight = steps(straight_m, straight_o, 6)
    for first, second in (("flat", "flat"), ("torch", "flat"), ("flat", "torch")):
        m1, o1 = make(first)
        assert o1.state_dict()["state"] == {}                                  # nothing before the first step, like Adam
        got = steps(m1, o1, 3)
        ckpt = {"model": {k: v.clone() for k, v in m1.state_dict().items()}, "opt": o1.state_dict()}
        assert set(ckpt["opt"]["state"][0]) >= {"step", "exp_avg", "exp_avg_sq"}
        assert float(ckpt["opt"]["state"][0]["step"]) == 3.0
        m2, o2 = make(second)
        m2.load_state_dict(ckpt["model"])
        o2.load_state_dict(ckpt["opt"])
        got += steps(m2, o2, 3)
        for a, b in zip(got, straight):
            assert abs(a - b) <= 1e-4 * abs(b), (first, second, got, straight)
        for (k, p), (_, q) in zip(straight_m.named_parameters(), m2.named_parameters()):
            assert rel_err(q.detach().cpu(), p.detach().cpu()) <= 2e-4, (first, second, k)


def test_backward_guards():
    """What stock autograd would catch, the zero-copy parameter view and the early reverse sweep must catch too
    (ADVICE round 1): a FlatAdam step between forward and backward, a second backward through released checkpoints,
    and a tensor hook that edits the gradient in place."""
    from neural_jump_ode import NeuralJumpODE, nj_ode_loss, FlatAdam, PackedBatch
    torch.manual_seed(2)
    model = NeuralJumpODE(**MK).to(DEV)
    model.flatten_parameters()                                                 # the sweeps read the parameters in place ...
    opt = FlatAdam(model.parameters(), lr=1e-3)                                # ... and FlatAdam steps on that very buffer
    assert opt._flat[0]["flat"].data_ptr() == model.flat_parameters()[0].data_ptr()
    data = PackedBatch.from_lists(*_bs_lists(40, seed=2), device=DEV)

    def fwd():
        model.zero_grad(set_to_none=True)
        p, b = model.forward_packed(data)
        return p, b, nj_ode_loss(data, None, p, b, **LK)

    p, b, loss = fwd()
    loss.backward()
    opt.step()                                                                 # njode_adam_step writes through a raw pointer ...
    p, b, loss = fwd()
    opt.step()
    with pytest.raises(RuntimeError, match="modified in place"):              # ... and the guard still sees it
        loss.backward()

    p, b, loss = fwd()
    loss.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="already consumed"):
        loss.backward()

    p, b, loss = fwd()
    loss.backward()
    plain = [q.grad.clone() for q in model.flat_parameters()]
    p, b, _ = fwd()
    p.register_hook(lambda g: g.mul_(2.0))                                     # in place: same address as the loss's own buffer
    b.register_hook(lambda g: g.mul_(2.0))
    nj_ode_loss(data, None, p, b, **LK).backward()
    for g2, g1 in zip((q.grad for q in model.flat_parameters()), plain):
        assert rel_err(g2, 2.0 * g1) <= 1e-6


@pytest.mark.parametrize("hidden,layers,act", [(32, 1, "relu"), (64, 1, "relu"), (128, 3, "tanh")])
def test_waves_equal_single_batch(hidden, layers, act):
    """forward_backward_waves (BASELINE config 4 is run in waves: 10 MB of checkpoints per trajectory): predictions of a
    wave are bit-identical to the same trajectories inside the whole batch, loss and gradients equal the single-batch
    ones to tolerance (only the summation order differs)."""
    from neural_jump_ode import NeuralJumpODE, nj_ode_loss, PackedBatch
    mk = dict(MK, hidden_dim=hidden, n_hidden_layers=layers, activation=act)
    torch.manual_seed(4)
    model = NeuralJumpODE(**mk).to(DEV)
    from test_gpu_parity import _random_batch
    data = PackedBatch.from_lists(*_random_batch(300, seed=8), device=DEV)
    model.zero_grad(set_to_none=True)
    p, b = model.forward_packed(data)
    loss = nj_ode_loss(data, None, p, b, **LK)
    loss.backward()
    whole = [q.grad.clone() for q in model.flat_parameters()]
    model.zero_grad(set_to_none=True)
    total = model.forward_backward_waves(data, 77, **LK)
    assert abs(total.item() - loss.item()) <= 5e-6 * abs(loss.item())
    for gw, g1 in zip((q.grad for q in model.flat_parameters()), whole):
        assert rel_err(gw, g1) <= TOL
    with torch.no_grad():
        sub = data.slice(77, 154)
        ps, bs_ = model.forward_packed(sub)
        a, z = int(data.offsets[77]), int(data.offsets[154])
        assert torch.equal(ps, p[a:z].detach()) and torch.equal(bs_, b[a:z].detach())


def test_device_generators_match_reference_moments():
    """Row N2 on the device: sample moments of the vectorised BS / OU / Heston / hybrid generators against the
    reference generators' fixture and the closed forms (same check as the CPU test, 20 000 paths, CUDA RNG)."""
    from test_host import _check_against_reference_moments
    _check_against_reference_moments(DEV, n=20000)
    from neural_jump_ode.simulation import make_packed_batch, make_mixed_ragged_batch
    b = make_packed_batch("heston", 1000, 0.1, n_steps=200, device=DEV, seed=1)
    assert b.B == 1000 and b.sizes == [20] * 1000 and b.times.is_cuda
    m = make_mixed_ragged_batch(1000, 0.02, 0.2, device=DEV, seed=1)
    sizes = torch.tensor(m.sizes)
    assert m.B == 1000 and int(sizes.min()) >= 2 and int(sizes.max()) <= 20 and bool(torch.isfinite(m.values).all())


def _dp_worker(rank, world, port, out):
    import torch.distributed as dist
    from neural_jump_ode import NeuralJumpODE, nj_ode_loss
    from neural_jump_ode.sharding import shard_lists
    from test_gpu_parity import _random_batch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    worst = 0.0
    for mk in (MK, dict(MK, hidden_dim=64, n_hidden_layers=2, activation="tanh")):
        bt, bv = _random_batch(257, seed=7)
        bt, bv = [t.to(dev) for t in bt], [v.to(dev) for v in bv]
        torch.manual_seed(0)
        model = NeuralJumpODE(**mk).to(dev)
        p, b = model(bt, bv)
        full = nj_ode_loss(bt, bv, p, b, **LK)
        full.backward()
        ref = {k: v.grad.clone() for k, v in model.named_parameters()}
        for waves in (False, True):
            model.zero_grad(set_to_none=True)
            model.enable_data_parallel()
            t, v, scale = shard_lists(bt, bv, rank, world)
            if waves:
                tot = model.forward_backward_waves(model.pack(t, v), 50, traj_scale=scale, **LK).clone()
            else:
                p, b = model(t, v)
                loss = nj_ode_loss(t, v, p, b, traj_scale=scale, **LK)
                loss.backward()
                tot = loss.detach().clone()
            dist.all_reduce(tot)
            model.enable_data_parallel(None)
            err = max(float((q.grad - ref[k]).abs().max() / ref[k].abs().max().clamp_min(1e-30)) for k, q in model.named_parameters())
            worst = max(worst, err, abs(float(tot) - float(full)) / abs(float(full)))
    if rank == 0:
        with open(out, "w") as f:
            f.write(repr(worst))
    dist.destroy_process_group()


def test_nccl_gradient_allreduce_two_gpus(tmp_path):
    """Trajectory sharding over 2 GPUs with the in-place NCCL all-reduce inside the reverse sweep (and once at the end
    of forward_backward_waves): loss and every gradient equal the single-GPU ones."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    out = str(tmp_path / "worst.txt")
    mp.spawn(_dp_worker, args=(2, 29533, out), nprocs=2, join=True)
    assert float(open(out).read()) <= TOL
