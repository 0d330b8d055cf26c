"""Bring-up / regression probe for the wide tcgen05 flavour (njode_wide.cu, njode_wgrad.cu), run on a B200:

  1. forward sweep:  predictions and every checkpoint plane (h, hidden-layer outputs) against the row-tiled FP32
     flavour on the same inputs (matched unit by unit through the two schedules);
  2. weight-gradient GEMM in isolation: the gradient it produced against numpy contractions of the very
     (d plane, activation plane, aux rows) it read -- separates its bugs from the reverse chain sweep's;
  3. end to end: every parameter gradient against the row-tiled flavour.

usage: python tests/wide_debug_probe.py [H L act B]      (prints one line per check; exit code 1 on a mismatch)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "neural-jump-ode_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from neural_jump_ode import NeuralJumpODE, nj_ode_loss, PackedBatch  # noqa: E402
from neural_jump_ode import _native as nat  # noqa: E402

DEV = "cuda:0"
TOL = 1e-5


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def random_batch(B, seed, n_steps=100, d_x=1):
    rng = np.random.RandomState(seed)
    bt, bv = [], []
    for _ in range(B):
        n = rng.randint(1, 14)
        if n == 1:
            idx = np.array([0])
        else:
            idx = np.sort(np.concatenate([[0, n_steps], rng.choice(np.arange(1, n_steps), n - 2, replace=False)]))
        bt.append(torch.linspace(0.0, 1.0, n_steps + 1)[torch.from_numpy(idx)])
        bv.append(torch.from_numpy((1.0 + 0.4 * rng.randn(n, d_x)).astype(np.float32)))
    return bt, bv


def run(model, impl, batch, lk):
    model.kernel_impl = impl
    model.zero_grad(set_to_none=True)
    batch._schedules.clear()
    p, b = model.forward_packed(batch)
    st = p._njode_state
    ckpt, sched = st.ckpt, st.sched
    loss = nj_ode_loss(batch, None, p, b, **lk)
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: q.grad.detach().cpu().numpy().copy() for k, q in model.named_parameters()}
    return dict(p=p.detach().cpu().numpy(), b=b.detach().cpu().numpy(), loss=loss.item(), grads=grads,
                ckpt=ckpt.detach().cpu().numpy(), sched=sched)


def unit_map(sched):
    """unit -> (tile, row)"""
    perm = sched.perm.cpu().numpy().reshape(-1, sched.tile_rows)
    out = {}
    for t in range(perm.shape[0]):
        for r in range(perm.shape[1]):
            if perm[t, r] >= 0:
                out[int(perm[t, r])] = (t, r)
    return out


def golden_main(name):
    """wide flavour on a golden case: per-parameter error against the reference's own gradients, worst element,
    and the weight-gradient GEMM against numpy on its own inputs."""
    from conftest import load_golden
    g = load_golden(name)
    mk = g["model"]
    model = NeuralJumpODE(**mk)
    model.load_state_dict(g["params"])
    model = model.to(DEV)
    H, L = mk["hidden_dim"], mk.get("n_hidden_layers", 1)
    S = 1 if mk.get("shared_network", False) else mk.get("num_moments", 1)
    batch = PackedBatch.from_lists(g["batch_times"], g["batch_values"], device=DEV)
    for impl in ("rowtile", "wide"):
        r = run(model, impl, batch, g["loss"])
        print(f"== {name} impl={impl} loss {r['loss']:.9g} (ref {g['ref_loss']:.9g})")
        for k_, ref in g["grads"].items():
            ref = ref.numpy().astype(np.float64)
            got = r["grads"][k_].astype(np.float64)
            d = np.abs(got - ref)
            e = d.max() / max(np.abs(ref).max(), 1e-30)
            idx = np.unravel_index(d.argmax(), d.shape)
            print(f"  {k_:<30} rel {e:.2e}  worst at {idx}: got {got[idx]:+.6e} ref {ref[idx]:+.6e}  tensor max {np.abs(ref).max():.3e}")
    sw = r["sched"]
    slots_w = sw.total_slots
    PL = 128 * H
    halfA = S * slots_w * (L + 1) * PL
    A = r["ckpt"][:halfA].reshape(S, slots_w, L + 1, 4, H // 8, 32, 8)
    D = r["ckpt"][halfA:2 * halfA].reshape(S, slots_w, L + 1, 16, H, 8)                     # half D: [row octet][feature][8 rows]
    X = r["ckpt"][2 * halfA:2 * halfA + S * slots_w * 128 * 8].reshape(S, slots_w, 128, 8)
    kmax_w = sw.tile_kmax.cpu().numpy()
    so_w = sw.tile_slot_off.cpu().numpy()
    print("  tiles", sw.n_tiles, "kmax", kmax_w.tolist())

    def plane(buf, s, slot, pl):
        if buf is D:
            return buf[s, slot, pl].transpose(0, 2, 1).reshape(128, H).astype(np.float64)
        return buf[s, slot, pl].transpose(0, 2, 1, 3).reshape(128, H).astype(np.float64)

    dx = mk["input_dim"]
    for s in range(S):
        for l in range(L + 1):
            dW = np.zeros((H, H + (dx + 2 if l == 0 else 0)))
            absW = np.zeros_like(dW)
            for t in range(sw.n_tiles):
                for k in range(int(kmax_w[t])):
                    slot = so_w[t] + k
                    d = plane(D, s, slot, l)
                    a_ = plane(A, s, slot, l)
                    if l == 0:
                        a_ = np.concatenate([a_, X[s, slot].astype(np.float64)[:, 1:dx + 3]], axis=1)
                    dW += d.T @ a_
                    absW += np.abs(d).T @ np.abs(a_)
            key = (f"ode_func.net.{3 * l}.weight" if S == 1 and mk.get("shared_network", False) else f"ode_funcs.{s}.net.{3 * l}.weight")
            got = r["grads"][key].astype(np.float64)
            ref = g["grads"][key].numpy().astype(np.float64)
            dd = np.abs(got - dW)
            idx = np.unravel_index(dd.argmax(), dd.shape)
            print(f"  stack {s} ode layer {l}: GEMM vs numpy-of-its-inputs {dd.max() / np.abs(dW).max():.2e} at {idx} "
                  f"(cancellation there {absW[idx] / max(abs(dW[idx]), 1e-30):.1f}x);  numpy-of-its-inputs vs golden {rel(dW, ref):.2e}")
    return 0


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--golden":
        return golden_main(sys.argv[2])
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    act = sys.argv[3] if len(sys.argv) > 3 else "tanh"
    B = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    scaling = sys.argv[5] if len(sys.argv) > 5 else "identity"
    M = 2
    torch.manual_seed(0)
    model = NeuralJumpODE(1, H, 1, dt_ode_step=0.01, num_moments=M, n_hidden_layers=L, activation=act,
                          input_scaling=scaling).to(DEV)
    bt, bv = random_batch(B, seed=3)
    batch = PackedBatch.from_lists(bt, bv, device=DEV)
    lk = dict(ignore_first_continuity=True, moment_weights=[1.0, 10.0])
    ok = True

    ref = run(model, "rowtile", batch, lk)
    got = run(model, "wide", batch, lk)
    S = M
    K = (batch.step_counts(model.descriptor()).cpu().numpy())
    print(f"H={H} L={L} act={act} scaling={scaling} B={B} units={batch.N} steps={int(K.sum())} "
          f"tiles wide={got['sched'].n_tiles} rowtile={ref['sched'].n_tiles}")

    for name in ("p", "b"):
        e = rel(got[name], ref[name])
        print(f"  forward {name:>2}: rel err {e:.3e}")
        ok &= e <= TOL
    print(f"  loss wide {got['loss']:.9g} rowtile {ref['loss']:.9g}")

    # ---- checkpoint planes: wide [half A][stack][slot][plane][chunk][row 128][8] vs rowtile [stack][slot][plane][row 32][H]
    sw, sr = got["sched"], ref["sched"]
    slots_w, slots_r = sw.total_slots, sr.total_slots
    PL = 128 * H
    halfA = S * slots_w * (L + 1) * PL
    A = got["ckpt"][:halfA].reshape(S, slots_w, L + 1, 4, H // 8, 32, 8)                    # half A: [row group][chunk][32 rows][8]
    Rk = ref["ckpt"].reshape(S, slots_r, L + 1, 32, H)
    mw, mr = unit_map(sw), unit_map(sr)
    so_w, so_r = sw.tile_slot_off.cpu().numpy(), sr.tile_slot_off.cpu().numpy()
    worst = np.zeros(L + 1)
    scale = np.zeros(L + 1) + 1e-30
    for u, (tw, rw) in mw.items():
        tr, rr = mr[u]
        for k in range(int(K[u]) + 1):
            for pl in range(L + 1):
                if pl > 0 and k == int(K[u]):
                    continue
                a = A[:, so_w[tw] + k, pl, rw // 32, :, rw % 32, :].reshape(S, H)
                r_ = Rk[:, so_r[tr] + k, pl, rr, :]
                worst[pl] = max(worst[pl], np.abs(a - r_).max())
                scale[pl] = max(scale[pl], np.abs(r_).max())
    for pl in range(L + 1):
        e = worst[pl] / scale[pl]
        print(f"  ckpt plane {pl}: rel err {e:.3e}")
        ok &= e <= TOL

    # ---- the weight-gradient GEMM against numpy contractions of its own inputs (ODE net only: the bulk) ----
    D = got["ckpt"][halfA:2 * halfA].reshape(S, slots_w, L + 1, 16, H, 8)                   # half D: [row octet][feature][8 rows]
    X = got["ckpt"][2 * halfA:2 * halfA + S * slots_w * 128 * 8].reshape(S, slots_w, 128, 8)
    kmax_w = sw.tile_kmax.cpu().numpy()

    def plane(buf, s, slot, pl):                                                            # -> [row][feature]
        if buf is D:
            return buf[s, slot, pl].transpose(0, 2, 1).reshape(128, H).astype(np.float64)
        return buf[s, slot, pl].transpose(0, 2, 1, 3).reshape(128, H).astype(np.float64)

    keys = list(got["grads"].keys())
    for s in range(S):
        for l in range(L + 1):
            dW = np.zeros((H, H + (3 if l == 0 else 0)))
            db = np.zeros(H)
            for t in range(sw.n_tiles):
                for k in range(int(kmax_w[t])):
                    slot = so_w[t] + k
                    d = plane(D, s, slot, l)
                    a_ = plane(A, s, slot, l)
                    if l == 0 and scaling == "tanh":
                        a_ = np.tanh(a_)
                    x = X[s, slot].astype(np.float64)
                    if l == 0:
                        a_ = np.concatenate([a_, x[:, 1:4]], axis=1)
                    dW += d.T @ a_
                    db += d.T @ x[:, 0]
            kw = [k_ for k_ in keys if k_.startswith(f"ode_funcs.{s}.net.{3 * l}.")]
            ew = rel(got["grads"][kw[0]], dW)
            eb = rel(got["grads"][kw[1]], db)
            print(f"  wgrad GEMM vs numpy  stack {s} ode layer {l}: W {ew:.3e}  b {eb:.3e}")
            ok &= ew <= TOL and eb <= TOL

    # ---- end to end: against the float64 oracle (ground truth) and against the row-tiled flavour ----
    from oracle import njode_oracle as orc
    cfg = orc.make_cfg(1, H, 1, 0.01, M, L, act, False, scaling)
    P = {k_: v.detach().cpu() for k_, v in model.state_dict().items()}
    tru = orc.run_flat(P, cfg, bt, bv, lk, dtype=torch.float64)
    for name, key in (("p", "preds"), ("b", "preds_before")):
        print(f"  {name} vs f64 oracle: wide {rel(got[name], tru[key].numpy()):.3e}   rowtile {rel(ref[name], tru[key].numpy()):.3e}")
    worst_w = worst_r = 0.0
    signed = []
    for k_ in keys:
        t_ = tru["grads"][k_].numpy()
        ew, er = rel(got["grads"][k_], t_), rel(ref["grads"][k_], t_)
        big = np.abs(t_) > 0.1 * np.abs(t_).max()
        signed.append(float(np.mean((got["grads"][k_][big] - t_[big]) / t_[big])))
        worst_w, worst_r = max(worst_w, ew), max(worst_r, er)
        flag = "" if ew <= TOL else "   <-- MISMATCH"
        if ew > 3e-6 or flag:
            print(f"  grad {k_:<32} vs f64: wide {ew:.3e} rowtile {er:.3e}{flag}")
        ok &= ew <= TOL
    print(f"  worst gradient error vs f64 oracle: wide {worst_w:.3e}   rowtile {worst_r:.3e};  "
          f"mean signed relative error of the large entries (wide): {np.mean(signed):+.3e}")
    import ctypes
    stt = ctypes.c_uint32(99)
    nat.check(nat.load().njode_device_status(ctypes.byref(stt)), "njode_device_status")
    det = (ctypes.c_uint32 * 4)()
    nat.check(nat.load().njode_device_status_detail(det), "njode_device_status_detail")
    print(f"  device status word: {stt.value}  detail: " + " ".join(hex(v) for v in det))
    print("WIDE DEBUG", "OK" if ok and stt.value == 0 else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
