"""Where the HOST time of the end-to-end step goes (default workload): un-synchronised stage clocks + cProfile."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-jump-ode_b200"))
import torch
from bench import WORKLOADS, make_batch
from neural_jump_ode import NeuralJumpODE, nj_ode_loss, PackedBatch
name = sys.argv[1] if len(sys.argv) > 1 else "ou_shared_b4096"
wl = WORKLOADS[name]; dev = torch.device("cuda")
torch.manual_seed(0); model = NeuralJumpODE(**wl["model"]).to(dev)
batch = make_batch(wl, wl["B"], dev, 1000)
h_t, h_v, h_o = batch.times.cpu().pin_memory(), batch.values.cpu().pin_memory(), batch.offsets.cpu().pin_memory()
sizes = batch.sizes; params = model.flat_parameters()
acc = {}
def step(clock):
    t0 = time.perf_counter()
    b = PackedBatch(h_t.to(dev, non_blocking=True), h_v.to(dev, non_blocking=True), h_o.to(dev, non_blocking=True), sizes)
    for p in params: p.grad = None
    t1 = time.perf_counter()
    preds, before = model.forward_packed(b)
    t2 = time.perf_counter()
    loss = nj_ode_loss(b, None, preds, before, **wl["loss"])
    t3 = time.perf_counter()
    loss.backward()
    t4 = time.perf_counter()
    v = loss.item()
    t5 = time.perf_counter()
    if clock:
        for k, d in (("h2d+pack", t1 - t0), ("forward_packed(host)", t2 - t1), ("loss(host)", t3 - t2), ("backward(host)", t4 - t3), ("item(wait)", t5 - t4), ("step", t5 - t0)):
            acc[k] = acc.get(k, 0.0) + d
    return v
for _ in range(10): step(False)
torch.cuda.synchronize()
N = 200
for _ in range(N): step(True)
print({k: round(v / N * 1e6, 1) for k, v in acc.items()}, "us of host wall per stage, no added syncs")
pr = cProfile.Profile(); pr.enable()
for _ in range(N): step(False)
pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(28)
