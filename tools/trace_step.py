"""Phase trace of CTA 0 of the tiled sweep kernels (needs `make -C neural-jump-ode_b200/csrc trace`).
Prints, per trace id, the median cycle distance to the previous record."""
import ctypes, os, sys, collections, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["NJODE_B200_LIB"] = os.path.join(ROOT, "neural-jump-ode_b200", "lib", "libnjode_b200_trace.so")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-jump-ode_b200"))
import torch
from bench import WORKLOADS
from neural_jump_ode import NeuralJumpODE, nj_ode_loss, _native as nat
from neural_jump_ode.simulation import make_packed_batch

name = sys.argv[1] if len(sys.argv) > 1 else "ou_shared_b4096"
B = int(sys.argv[2]) if len(sys.argv) > 2 else WORKLOADS[name]["B"]
wl = WORKLOADS[name]
lib = nat.load()
lib.njode_tiled_trace_fetch.restype = ctypes.c_int
lib.njode_tiled_trace_fetch.argtypes = [ctypes.c_void_p, ctypes.c_int]
torch.manual_seed(0)
model = NeuralJumpODE(**wl["model"]).to("cuda")
batch = make_packed_batch(wl["process"], B, wl["obs_fraction"], n_steps=wl["n_steps"], T=wl["T"], device="cuda", seed=1000, **wl["pkw"])
buf = (ctypes.c_longlong * 12288)()
for it in range(3):
    model.zero_grad()
    p, b = model.forward_packed(batch)
    loss = nj_ode_loss(batch, None, p, b, **wl["loss"])
    loss.backward()
    torch.cuda.synchronize()
    n = lib.njode_tiled_trace_fetch(buf, 12288)
rec = [(buf[i] & 255, buf[i] >> 8) for i in range(n)]
print("records", n)
d = collections.defaultdict(list)
for (i0, t0), (i1, t1) in zip(rec, rec[1:]):
    d[(i0, i1)].append(t1 - t0)
for k in sorted(d, key=lambda k: (k[1] >= 32, k[1], k[0])):
    v = d[k]
    if len(v) >= 3:
        print(f"{k[0]:3d} -> {k[1]:3d}: n={len(v):4d} median={statistics.median(v):8.0f} min={min(v):7d} max={max(v):8d}")

# merged absolute timeline (reverse-sweep worker thread 0 and the MMA issuer share the SM clock): a few steps mid-run
bw = sorted((t, i) for i, t in rec if (1 <= i <= 13) or i >= 64)
if bw:
    mid = len(bw) // 2
    t0 = bw[mid][0]
    print("timeline (cycles relative, id): worker ids 1-12, issuer 64=loop top 65=ops#1 seen 66=chain1 issued 67=ops#2 seen 68=chain2+wgrad issued")
    print("  " + "  ".join(f"{t - t0}:{i}" for t, i in bw[mid:mid + 60]))

# reverse-sweep worker, whole tiles (ids 17-31 mark the per-tile phases outside the Euler loop: 17 tile start, 18/19 first
# hand-over of a readout, 20 chain 1 done, 21 second hand-over, 22 chain 2 done, 23 weight gradients done, 24 merged,
# 25/26 around the wait for the last step's weight gradients, 27-31 jump net)
wk = [(t, i) for i, t in rec if 1 <= i <= 31]
if wk:
    tail = wk[-150:]
    t0 = tail[0][0]
    print("worker timeline, last records (cycles relative : id):")
    print("  " + "  ".join(f"{t - t0}:{i}" for t, i in tail))
    starts = [t for t, i in wk if i == 17]
    if len(starts) > 2:
        gaps = [b - a for a, b in zip(starts, starts[1:])]
        print("cycles per tile (17 -> 17):", gaps[-12:])
