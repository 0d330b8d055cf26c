"""Bring-up probe: one fwd+bwd of a bench workload at a chosen batch with NJODE_NO_TRAP=1, then the device status words
(which barrier of which CTA gave up).  usage: NJODE_NO_TRAP=1 python tools/k3_debug_big.py [workload] [batch]"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-jump-ode_b200"))
import bench
from neural_jump_ode import NeuralJumpODE, nj_ode_loss, _native as nat
name = sys.argv[1] if len(sys.argv) > 1 else "heston_h128_l3"
wl = dict(bench.WORKLOADS[name])
B = int(sys.argv[2]) if len(sys.argv) > 2 else wl["B"]
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = NeuralJumpODE(**wl["model"]).to(dev)
batch = bench.make_batch(wl, B, dev, 1000)
p, b = model.forward_packed(batch)
loss = nj_ode_loss(batch, None, p, b, **wl["loss"])
try:
    loss.backward()
    torch.cuda.synchronize()
    print("loss", float(loss), "grad norm", float(torch.cat([q.grad.flatten() for q in model.parameters()]).norm()))
except Exception as e:
    print("FAILED:", str(e).splitlines()[0])
lib = nat.load()
st = ctypes.c_uint32(99)
lib.njode_device_status(ctypes.byref(st))
det = (ctypes.c_uint32 * 4)()
lib.njode_device_status_detail(det)
print(f"B={B} tiles={p._njode_state.sched.n_tiles} status {st.value:#x} detail " + " ".join(hex(v) for v in det))
