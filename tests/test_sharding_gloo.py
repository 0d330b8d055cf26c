"""The N > 1 path on CPU: two gloo ranks shard a ragged batch, each computes its slice's loss and
gradients with the float64 oracle (test infrastructure; on the B200 box the CUDA sweep kernels do this
part), scales by 1/B_global, and ONE all-reduce of the flat [gradients, loss] vector reproduces the
single-process result.  Also checks the shard bounds (equal counts / balanced by Euler steps)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT, PKG
from oracle import njode_oracle as orc

MODEL = dict(input_dim=1, hidden_dim=16, output_dim=1, dt_ode_step=0.05, num_moments=2)
LOSS = dict(ignore_first_continuity=True, moment_weights=[1.0, 10.0], variance_method="direct")


def _batch(B=11, seed=3):
    rng = np.random.RandomState(seed)
    bt, bv = [], []
    for b in range(B):
        n = rng.randint(1, 7)
        idx = np.sort(rng.choice(np.arange(1, 40), max(n - 1, 0), replace=False))
        t = torch.from_numpy(np.concatenate([[0.0], idx / 40.0]).astype(np.float32))
        bt.append(t)
        bv.append(torch.from_numpy((1.0 + 0.3 * rng.randn(len(t), 1)).astype(np.float32)))
    return bt, bv


def _cfg():
    return orc.make_cfg(MODEL["input_dim"], MODEL["hidden_dim"], MODEL["output_dim"], MODEL["dt_ode_step"], MODEL["num_moments"])


def _worker(rank, world, port, out):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from neural_jump_ode import NeuralJumpODE
    from neural_jump_ode.sharding import shard_lists, allreduce_gradients
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = NeuralJumpODE(**MODEL).double()          # identical replicas on every rank
        P = {k: v.detach().clone() for k, v in model.state_dict().items()}
        bt, bv = _batch()
        steps = [int(orc.step_counts(t.numpy(), MODEL["dt_ode_step"]).sum()) for t in bt]
        my_t, my_v, traj_scale = shard_lists(bt, bv, rank, world, steps)
        params = [p for _, p in model.named_parameters()]
        names = [k for k, _ in model.named_parameters()]
        if len(my_t) > 0:
            ref = orc.run_flat(P, _cfg(), my_t, my_v, LOSS, dtype=torch.float64)
            w = len(my_t) * traj_scale                   # oracle loss is a mean over the local slice
            for k, p in zip(names, params):
                p.grad = (ref["grads"][k].double() * w).reshape(p.shape).clone()
            loss = ref["loss"].double() * w
        else:
            loss = torch.zeros((), dtype=torch.float64)
        total = allreduce_gradients(params, loss)
        if rank == 0:
            torch.save({"loss": total, "grads": {k: p.grad.clone() for k, p in zip(names, params)}}, out)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
def test_two_rank_allreduce_matches_single_process(tmp_path):
    out = str(tmp_path / "rank0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    from neural_jump_ode import NeuralJumpODE
    model = NeuralJumpODE(**MODEL).double()
    P = {k: v.detach().clone() for k, v in model.state_dict().items()}
    bt, bv = _batch()
    ref = orc.run_flat(P, _cfg(), bt, bv, LOSS, dtype=torch.float64)
    assert abs(float(got["loss"]) - float(ref["loss"])) <= 1e-12 * abs(float(ref["loss"]))
    for k, g in got["grads"].items():
        want = ref["grads"][k].double().reshape(g.shape)
        assert float((g - want).abs().max()) <= 1e-12 * max(float(want.abs().max()), 1e-30), k


def test_shard_bounds():
    from neural_jump_ode.sharding import shard_bounds, shard_lists
    assert shard_bounds(10, 1) == [0, 10]
    assert shard_bounds(10, 4) == [0, 2, 5, 7, 10]
    with pytest.raises(ValueError, match="empty"):                              # more ranks than trajectories: every rank
        shard_bounds(3, 8)                                                      # refuses alike (no rank may sit out the all-reduce)
    assert shard_bounds(4, 4, [100, 1, 1, 1]) == [0, 1, 2, 3, 4]                # balancing by steps never starves a rank
    b = shard_bounds(6, 2, [100, 1, 1, 1, 1, 96])
    assert b == [0, 1, 6]                                                       # balanced by steps, contiguous
    b = shard_bounds(8, 3, [5] * 8)
    assert b[0] == 0 and b[-1] == 8 and all(x <= y for x, y in zip(b, b[1:]))
    bt, bv = _batch(5)
    seen = []
    for r in range(3):
        t, v, scale = shard_lists(bt, bv, r, 3)
        assert scale == 1.0 / 5 and len(t) == len(v)
        seen += [id(x) for x in t]
    assert seen == [id(x) for x in bt]                                          # a partition, in order
    with pytest.raises(ValueError):
        shard_bounds(4, 0)
