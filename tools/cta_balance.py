"""Per-CTA cycles and work of the wide sweeps on a bench workload (njode_debug_cta_cycles): is the kernel time the
slowest CTA's, and is a CTA's time proportional to the chain GEMMs it runs?   python tools/cta_balance.py [workload] [batch]"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-jump-ode_b200"))
import bench
from neural_jump_ode import NeuralJumpODE, nj_ode_loss, _native as nat

name = sys.argv[1] if len(sys.argv) > 1 else "heston_h128_l3"
wl = dict(bench.WORKLOADS[name])
if len(sys.argv) > 2:
    wl["B"] = int(sys.argv[2])
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = NeuralJumpODE(**wl["model"]).to(dev)
batch = bench.make_batch(wl, wl["B"], dev, 1000)
lib = nat.load()
for it in range(3):
    model.zero_grad(set_to_none=True)
    p, b = model.forward_packed(batch)
    torch.cuda.synchronize()
    buf = (ctypes.c_uint64 * (2 * 148))()
    nat.check(lib.njode_debug_cta_cycles(buf, 148), "dbg")
    fwd = np.array(buf[:], dtype=np.float64).reshape(148, 2)
    loss = nj_ode_loss(batch, None, p, b, **wl["loss"])
    loss.backward()
    torch.cuda.synchronize()
for tag, a in (("forward", fwd),):
    cyc, work = a[:, 0], a[:, 1]
    order = np.argsort(cyc)
    print(f"{name} {tag}: cycles min {cyc.min():.3g} mean {cyc.mean():.3g} max {cyc.max():.3g} (max/mean {cyc.max() / cyc.mean():.2f}); "
          f"GEMMs min {work.min():.0f} mean {work.mean():.0f} max {work.max():.0f}; cycles per GEMM min {(cyc / work).min():.0f} "
          f"mean {(cyc / work).mean():.0f} max {(cyc / work).max():.0f}")
    print("  slowest CTAs (block, cycles, GEMMs):", [(int(i), int(cyc[i]), int(work[i])) for i in order[-6:]])
    print("  fastest CTAs (block, cycles, GEMMs):", [(int(i), int(cyc[i]), int(work[i])) for i in order[:6]])
