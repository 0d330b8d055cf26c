"""Host/device time of each stage of the end-to-end step (pinned host inputs -> loss on the host), default workload."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-jump-ode_b200"))
import torch
from bench import WORKLOADS, make_batch
from neural_jump_ode import NeuralJumpODE, nj_ode_loss, PackedBatch
wl = WORKLOADS["ou_shared_b4096"]; dev = torch.device("cuda")
torch.manual_seed(0); model = NeuralJumpODE(**wl["model"]).to(dev)
batch = make_batch(wl, wl["B"], dev, 1000)
h_t, h_v, h_o = batch.times.cpu().pin_memory(), batch.values.cpu().pin_memory(), batch.offsets.cpu().pin_memory()
sizes = batch.sizes; desc = model.descriptor(); params = model.flat_parameters()
def sync(): torch.cuda.synchronize(); return time.perf_counter()
acc = {}
def add(k, dt): acc[k] = acc.get(k, 0.0) + dt
N = 50
for it in range(N + 5):
    t0 = sync()
    b = PackedBatch(h_t.to(dev, non_blocking=True), h_v.to(dev, non_blocking=True), h_o.to(dev, non_blocking=True), sizes)
    t1 = sync()
    sched = b.schedule(desc)
    t2 = sync()
    for p in params: p.grad = None
    preds, before = model.forward_packed(b)
    t3 = sync()
    loss = nj_ode_loss(b, None, preds, before, **wl["loss"])
    t4 = sync()
    loss.backward()
    t5 = sync()
    v = loss.item()
    t6 = sync()
    if it >= 5:
        for k, d in (("h2d", t1 - t0), ("schedule", t2 - t1), ("forward", t3 - t2), ("loss", t4 - t3), ("backward", t5 - t4), ("item", t6 - t5)): add(k, d)
print({k: round(v / N * 1e6, 1) for k, v in acc.items()}, "us per stage (each stage synchronised);  sum", round(sum(acc.values()) / N * 1e6, 1))
