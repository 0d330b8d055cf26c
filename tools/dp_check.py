"""2-rank (or N-rank) check on GPUs, run under torch.distributed.run: every rank integrates its shard with the
data-parallel gradient all-reduce inside the reverse sweep; the result must equal the whole batch on one GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-jump-ode_b200"))
import numpy as np, torch, torch.distributed as dist
from neural_jump_ode import NeuralJumpODE, nj_ode_loss
from neural_jump_ode.sharding import shard_lists

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rng = np.random.RandomState(7)
bt, bv = [], []
for b in range(257):
    n = rng.randint(2, 14)
    idx = np.sort(np.concatenate([[0, 100], rng.choice(np.arange(1, 100), n - 2, replace=False)]))
    bt.append(torch.linspace(0.0, 1.0, 101)[torch.from_numpy(idx)].to(dev))
    bv.append(torch.from_numpy((1.0 + 0.4 * rng.randn(n, 1)).astype(np.float32)).to(dev))
lk = dict(ignore_first_continuity=True, moment_weights=[1.0, 10.0])
worst = 0.0
for mk in (dict(input_dim=1, hidden_dim=32, output_dim=1, dt_ode_step=0.01, num_moments=2),
           dict(input_dim=1, hidden_dim=64, output_dim=1, dt_ode_step=0.01, num_moments=2, n_hidden_layers=2, activation="tanh")):
    torch.manual_seed(0)
    model = NeuralJumpODE(**mk).to(dev)
    ref = {}
    p, b = model(bt, bv)
    full = nj_ode_loss(bt, bv, p, b, **lk)
    full.backward()
    ref = {k: v.grad.clone() for k, v in model.named_parameters()}
    model.zero_grad(set_to_none=True)
    model.enable_data_parallel()
    t, v, scale = shard_lists(bt, bv, rank, world)
    p, b = model(t, v)
    loss = nj_ode_loss(t, v, p, b, traj_scale=scale, **lk)
    loss.backward()
    tot = loss.detach().clone()
    dist.all_reduce(tot)
    err = max(float((q.grad - ref[k]).abs().max() / ref[k].abs().max().clamp_min(1e-30)) for k, q in model.named_parameters())
    lerr = abs(float(tot) - float(full)) / abs(float(full))
    worst = max(worst, err, lerr)
    if rank == 0:
        print(f"H={mk['hidden_dim']}: world {world}: max grad rel err {err:.2e}, loss rel err {lerr:.2e}")
assert worst < 1e-5, worst
if rank == 0:
    print("dp_check OK")
dist.destroy_process_group()
