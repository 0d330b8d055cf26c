"""pytest configuration: the `gpu` marker and import paths.

`-m "not gpu"`: oracle vs golden vectors, host logic, C-ABI symbol check (no GPU needed).
`-m gpu`      : parity tests proper, through the C-ABI, on a real B200.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neural-jump-ode_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))


def load_golden(name):
    """Load one golden case -> dict with cfg, loss kwargs, lists of tensors and reference outputs."""
    import torch
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    conf = json.loads(bytes(z["config_json"]).decode())
    off = z["offsets"]
    times, values = z["times"], z["values"]
    bt = [torch.from_numpy(times[off[b]:off[b + 1]].copy()) for b in range(len(off) - 1)]
    bv = [torch.from_numpy(values[off[b]:off[b + 1]].copy()) for b in range(len(off) - 1)]
    params = {k[len("param/"):]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith("param/")}
    grads = {k[len("grad/"):]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith("grad/")}
    has_grad = {k[len("has_grad/"):]: bool(z[k]) for k in z.files if k.startswith("has_grad/")}
    return dict(name=name, model=conf["model"], loss=conf["loss"], seed=int(conf.get("seed", 0)),
                kink_margin=conf.get("kink_margin"), batch_times=bt, batch_values=bv,
                offsets=off, times=times, values=values, params=params, grads=grads, has_grad=has_grad,
                preds=torch.from_numpy(z["preds"].copy()), preds_before=torch.from_numpy(z["preds_before"].copy()),
                ref_loss=float(z["loss"]), step_log=z["step_log"])


def rel_err(a, b):
    """max-norm relative error ||a-b||inf / max(||b||inf, tiny) over the finite entries of the reference b.
    Where b is NaN (the reference propagates a NaN observation into predictions, loss and gradients) a must be
    NaN too and vice versa: a mismatch of the NaN pattern is an infinite error."""
    import torch
    a = torch.as_tensor(a, dtype=torch.float64).detach()
    b = torch.as_tensor(b, dtype=torch.float64).detach()
    if a.numel() == 0:
        return 0.0
    nan_a, nan_b = torch.isnan(a), torch.isnan(b)
    if not torch.equal(nan_a, nan_b):
        return float("inf")
    ok = ~nan_b
    if not bool(ok.any()):
        return 0.0
    return float((a[ok] - b[ok]).abs().max() / max(float(b[ok].abs().max()), 1e-30))


def loss_close(got, ref, tol):
    """|got - ref| <= tol * |ref|, or both NaN."""
    import math
    got, ref = float(got), float(ref)
    if math.isnan(ref) or math.isnan(got):
        return math.isnan(ref) and math.isnan(got)
    return abs(got - ref) <= tol * abs(ref)


@pytest.fixture(params=golden_names())
def golden(request):
    return load_golden(request.param)
