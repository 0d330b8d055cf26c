"""Where a wave-mode step spends its GPU time: torch.profiler kernel totals over two steps of heston_h128_l3_1m.
usage: python tools/wave_overhead.py [batch] [wave]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-jump-ode_b200"))
import bench
from neural_jump_ode import NeuralJumpODE
wl = dict(bench.WORKLOADS["heston_h128_l3_1m"])
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
wave = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = NeuralJumpODE(**wl["model"]).to(dev)
batch = bench.make_batch(wl, B, dev, 1000)
def step():
    model.zero_grad(set_to_none=True)
    return model.forward_backward_waves(batch, wave, **wl["loss"])
for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); step(); e1.record(); torch.cuda.synchronize()
print("2 steps:", e0.elapsed_time(e1), "ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
