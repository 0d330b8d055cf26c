"""Where the wide kernels spend their time, per role phase (needs `make -C neural-jump-ode_b200/csrc phase`):
    NJODE_B200_LIB=neural-jump-ode_b200/lib/libnjode_b200_phase.so python tools/phase_wide.py [workload] [batch]"""
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
n = 148


def fetch(which):
    buf = (ctypes.c_uint64 * (8 * n))()
    nat.check(lib.njode_debug_phase(which, buf, n), "njode_debug_phase")
    return np.array(buf[:], dtype=np.float64).reshape(n, 8)


def cycles():
    buf = (ctypes.c_uint64 * (2 * n))()
    nat.check(lib.njode_debug_cta_cycles(buf, n), "njode_debug_cta_cycles")
    return np.array(buf[:], dtype=np.float64).reshape(n, 2)


for it in range(3):
    model.zero_grad(set_to_none=True)
    model.eager_backward = False
    p, b = model.forward_packed(batch)
    fwd, fwd_c, fwd_i = fetch(1), cycles(), fetch(2)
    loss = nj_ode_loss(batch, None, p, b, **wl["loss"])
    loss.backward()
    bwd, bwd_c, wg, bwd_i = fetch(1), cycles(), fetch(3), fetch(2)
for tag, ph, cyc, iss in (("forward sweep", fwd, fwd_c, fwd_i), ("reverse sweep", bwd, bwd_c, bwd_i)):
    g = cyc[:, 1].mean()
    print(f"{name} {tag}: {cyc[:, 0].mean():.4g} cycles per CTA, {g:.0f} chain GEMMs -> {cyc[:, 0].mean() / g:.0f} cycles per GEMM; "
          f"worker thread 0: epilogue + emission {ph[:, 0].mean() / g:.0f}, accumulator wait {ph[:, 1].mean() / g:.0f} per GEMM; "
          f"issuer: operand wait {iss[:, 0].mean() / g:.0f}, weight wait {iss[:, 1].mean() / g:.0f}, issue {iss[:, 2].mean() / g:.0f} per GEMM")
names = ["bookkeeping", "raw-stage wait", "smem loads + operand-stage wait", "split+stores", "fence+hand-over", "merge", "flush"]
tot = wg[:, :7].sum(1).mean()
print(f"{name} weight-gradient GEMM converter thread 0: {tot:.4g} cycles per CTA: " + ", ".join(f"{nm} {100 * wg[:, i].mean() / tot:.1f}%" for i, nm in enumerate(names)))
