"""Minimal driver for ncu: a few fwd+loss+bwd steps of a workload (no timing, no CPU baseline)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-jump-ode_b200"))
import torch
from bench import WORKLOADS
from neural_jump_ode import NeuralJumpODE, nj_ode_loss
from bench import make_batch

name = sys.argv[1] if len(sys.argv) > 1 else "ou_shared_b4096"
B = int(sys.argv[2]) if len(sys.argv) > 2 else WORKLOADS[name]["B"]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
wl = WORKLOADS[name]
torch.manual_seed(0)
model = NeuralJumpODE(**wl["model"]).to("cuda")
batch = make_batch(wl, B, "cuda", 1000)
for _ in range(n):
    model.zero_grad()
    p, b = model.forward_packed(batch)
    loss = nj_ode_loss(batch, None, p, b, **wl["loss"])
    loss.backward()
torch.cuda.synchronize()
print("loss", loss.item(), "steps", batch.schedule(model.descriptor()).total_steps)
