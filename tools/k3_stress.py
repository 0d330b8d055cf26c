"""Repeat one golden case on the wide flavour and count runs whose gradients miss the golden tolerance
(flushes out timing-dependent bugs in the weight-gradient GEMM pipeline).
usage: python tools/k3_stress.py [golden name] [repeats]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-jump-ode_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from neural_jump_ode import NeuralJumpODE, nj_ode_loss, PackedBatch
from conftest import load_golden

name = sys.argv[1] if len(sys.argv) > 1 else "heston_h128_l3_tanh"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
g = load_golden(name)
model = NeuralJumpODE(**g["model"])
model.load_state_dict(g["params"])
model = model.to("cuda:0")
model.kernel_impl = "wide"
batch = PackedBatch.from_lists(g["batch_times"], g["batch_values"], device="cuda:0")
junk = torch.empty(64 << 20, device="cuda:0")
bad = {}
for it in range(n):
    if it % 3 == 1:
        junk.normal_()          # disturb L2 / timing
    model.zero_grad(set_to_none=True)
    p, b = model.forward_packed(batch)
    loss = nj_ode_loss(batch, None, p, b, **g["loss"])
    loss.backward()
    torch.cuda.synchronize()
    for k, q in model.named_parameters():
        ref = g["grads"][k].numpy().astype(np.float64)
        e = np.abs(q.grad.cpu().numpy().astype(np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30)
        if not e <= 1e-5:
            bad.setdefault(k, []).append((it, float(e)))
print(f"{name}: {n} runs, parameters out of tolerance:", {k: (len(v), v[:3]) for k, v in bad.items()} or "none")
