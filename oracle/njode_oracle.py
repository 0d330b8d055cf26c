"""CPU oracle for the Neural Jump ODE hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``neural-jump-ode_b200/``) never imports it and has no CPU fallback.

Parity status: PINNED.  ``tests/golden/make_golden.py`` ran the *unmodified*
reference (``/root/reference/neural_jump_ode/models/jump_ode.py``) in the build
container and stored inputs / state_dict / preds / preds_before / loss / every
parameter gradient / the exact float32 Euler step log as ``tests/golden/*.npz``;
``tests/test_oracle.py`` checks every function below against those vectors.

Three restatements live here:

* ``euler_schedule``    exact float32 restatement of the step rule, jump_ode.py:188-203
* ``forward_port`` / ``loss_port``   a per-trajectory, per-step eager restatement with
  the same op granularity as the reference (one tiny Linear per layer per step);
  it is the CPU baseline that ``bench.py`` times (``cpu_baseline.kind == "port"``)
* ``forward_flat`` / ``loss_flat``  an interval-flattened, vectorised restatement
  (float32 or float64) used as the checker at sizes the eager port cannot reach.

Parameters are passed as a plain ``dict`` keyed by the reference's state_dict
names (``jump_nns.0.net.0.weight`` ...), see jump_ode.py:19-22, :36-40, :70-74.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------- #
# configuration helpers
# --------------------------------------------------------------------------- #

_ACTS = {
    # jump_ode.py:6-13; unknown names fall back to ReLU (jump_ode.py:18)
    "relu": torch.relu,
    "tanh": torch.tanh,
    "sigmoid": torch.sigmoid,
    "elu": F.elu,
    "leaky_relu": lambda v: F.leaky_relu(v, 0.01),
    "selu": F.selu,
}
_SCALINGS = {
    # jump_ode.py:43-50
    "identity": lambda v: v,
    "none": lambda v: v,
    "tanh": torch.tanh,
    "sigmoid": torch.sigmoid,
}


def make_cfg(input_dim, hidden_dim, output_dim, dt_ode_step=None, num_moments=1,
             n_hidden_layers=1, activation="relu", shared_network=False,
             input_scaling="identity"):
    return dict(input_dim=int(input_dim), hidden_dim=int(hidden_dim), output_dim=int(output_dim),
                dt_ode_step=None if dt_ode_step is None else float(dt_ode_step),
                num_moments=int(num_moments), n_hidden_layers=int(n_hidden_layers),
                activation=str(activation), shared_network=bool(shared_network),
                input_scaling=str(input_scaling))


def n_stacks(cfg):
    return 1 if cfg["shared_network"] else cfg["num_moments"]


def stack_prefixes(cfg, s):
    """state_dict prefixes of stack ``s`` (jump_ode.py:100-116)."""
    if cfg["shared_network"]:
        return "jump_nn.", "ode_func.", "output_nn."
    return f"jump_nns.{s}.", f"ode_funcs.{s}.", f"output_nns.{s}."


def layer_keys(prefix, n_hidden_layers):
    """Linear layers sit at Sequential indices 0,3,...,3L in all three nets."""
    return [(f"{prefix}net.{3 * i}.weight", f"{prefix}net.{3 * i}.bias")
            for i in range(n_hidden_layers + 1)]


def init_params(cfg, seed=0, dtype=torch.float32):
    """Random parameters with nn.Linear-like uniform bounds (NOT the reference RNG stream;
    golden tests load the reference's own state_dict instead)."""
    g = torch.Generator().manual_seed(seed)
    H, dx, L = cfg["hidden_dim"], cfg["input_dim"], cfg["n_hidden_layers"]
    O = cfg["output_dim"] * (cfg["num_moments"] if cfg["shared_network"] else 1)
    P = {}

    def lin(wk, bk, out_f, in_f):
        bound = 1.0 / np.sqrt(in_f)
        P[wk] = ((torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound).to(dtype)
        P[bk] = ((torch.rand(out_f, generator=g) * 2 - 1) * bound).to(dtype)

    for s in range(n_stacks(cfg)):
        pj, po, pq = stack_prefixes(cfg, s)
        for i, (wk, bk) in enumerate(layer_keys(pj, L)):
            lin(wk, bk, H, dx if i == 0 else H)
        for i, (wk, bk) in enumerate(layer_keys(po, L)):
            lin(wk, bk, H, H + dx + 2 if i == 0 else H)
        for i, (wk, bk) in enumerate(layer_keys(pq, L)):
            lin(wk, bk, O if i == L else H, H)
    return P


# --------------------------------------------------------------------------- #
# (1) the float32 Euler step schedule  -- jump_ode.py:188-203
# --------------------------------------------------------------------------- #

def euler_schedule(t_i, t_next, dt):
    """Return the list of (t_cur, t_new) float32 pairs the reference integrates over
    between two observations.

    dt is None  -> one step (jump_ode.py:188-190).
    else        -> ``while t_cur + dt < t_next`` full steps in float32 accumulation,
                   then one closing step ``if t_cur < t_next`` (jump_ode.py:193-203).
    The reference evaluates ``t_cur + dt`` as float32 tensor + python float, i.e.
    fl32(t_cur + fl32(dt)); both comparisons are done on float32 values.
    """
    t_i = np.float32(t_i)
    t_next = np.float32(t_next)
    if dt is None:
        return [(t_i, t_next)]
    dt32 = np.float32(dt)
    out = []
    t_cur = t_i
    while np.float32(t_cur + dt32) < t_next:
        t_new = np.float32(t_cur + dt32)
        out.append((t_cur, t_new))
        t_cur = t_new
    if t_cur < t_next:
        out.append((t_cur, t_next))
    return out


def step_counts(times, dt):
    """Per-observation Euler step counts of one trajectory (last observation: 0)."""
    t = np.asarray(times, dtype=np.float32)
    k = np.zeros(len(t), dtype=np.int32)
    for i in range(len(t) - 1):
        k[i] = len(euler_schedule(t[i], t[i + 1], dt))
    return k


# --------------------------------------------------------------------------- #
# (2) eager per-trajectory port (CPU baseline; same op granularity as the reference)
# --------------------------------------------------------------------------- #

def _act(cfg):
    return _ACTS.get(cfg["activation"].lower(), torch.relu)


def _scale(cfg):
    name = cfg["input_scaling"]
    if name not in _SCALINGS:
        raise ValueError(f"Unknown input_scaling: {name}. Use 'identity', 'tanh', or 'sigmoid'.")
    return _SCALINGS[name]


def _jump(P, keys, act, x):
    # JumpNN: Linear, act, [Linear, act] x L  (jump_ode.py:19-22)
    v = x
    for wk, bk in keys:
        v = act(F.linear(v, P[wk], P[bk]))
    return v


def _ode(P, keys, act, scale, t_new, h, x_last, t_cur):
    # ODEFunc.forward (jump_ode.py:52-63): inp = [s(h), s(x), t_cur, t_new - t_cur]
    t_el = (t_new - t_cur).expand_as(h[..., :1])
    t_rel = t_cur.expand_as(h[..., :1])
    v = torch.cat([scale(h), scale(x_last), t_rel, t_el], dim=-1)
    for i, (wk, bk) in enumerate(keys):
        v = F.linear(v, P[wk], P[bk])
        if i < len(keys) - 1:
            v = act(v)
    return v


def _out(P, keys, act, h):
    # OutputNN: [Linear, act] x L, Linear  (jump_ode.py:70-74)
    v = h
    for i, (wk, bk) in enumerate(keys):
        v = F.linear(v, P[wk], P[bk])
        if i < len(keys) - 1:
            v = act(v)
    return v


def forward_port(P, cfg, batch_times, batch_values, step_log=None):
    """Eager restatement of NeuralJumpODE.forward (jump_ode.py:142-233).

    Returns (preds, preds_before): lists of (n_i, d_y, M) tensors.
    ``step_log`` (optional list) receives one (t_cur, t_new) float pair per Euler step.
    """
    act, scale = _act(cfg), _scale(cfg)
    L, M, dy = cfg["n_hidden_layers"], cfg["num_moments"], cfg["output_dim"]
    S = n_stacks(cfg)
    keys = []
    for s in range(S):
        pj, po, pq = stack_prefixes(cfg, s)
        keys.append((layer_keys(pj, L), layer_keys(po, L), layer_keys(pq, L)))
    dt = cfg["dt_ode_step"]

    def readout(hs):
        if cfg["shared_network"]:
            return _out(P, keys[0][2], act, hs[0]).view(1, dy, M)       # jump_ode.py:170-172
        return torch.stack([_out(P, keys[s][2], act, hs[s]) for s in range(S)], dim=-1)  # :177-179

    def euler(hs, x, t_cur, t_new):
        # jump_ode.py:122-140, dt as a float32 tensor difference
        if step_log is not None:
            step_log.append((float(t_cur), float(t_new)))
        return [hs[s] + (t_new - t_cur) * _ode(P, keys[s][1], act, scale, t_new, hs[s], x, t_cur)
                for s in range(S)]

    preds, preds_before = [], []
    for times, values in zip(batch_times, batch_values):
        n = values.shape[0]
        ys, ybs = [], []
        y_before = torch.zeros(1, dy, M, dtype=values.dtype)            # jump_ode.py:161
        for i in range(n):
            x = values[i].unsqueeze(0)
            hs = [_jump(P, keys[s][0], act, x) for s in range(S)]
            ys.append(readout(hs).squeeze(0))
            ybs.append(y_before.squeeze(0))
            if i < n - 1:
                t_cur, t_next = times[i], times[i + 1]
                if dt is None:
                    hs = euler(hs, x, t_cur, t_next)
                else:
                    while t_cur + dt < t_next:                          # jump_ode.py:196
                        t_new = t_cur + dt
                        hs = euler(hs, x, t_cur, t_new)
                        t_cur = t_new
                    if t_cur < t_next:                                  # jump_ode.py:201
                        hs = euler(hs, x, t_cur, t_next)
                y_before = readout(hs)
        preds.append(torch.stack(ys, dim=0))
        preds_before.append(torch.stack(ybs, dim=0))
    return preds, preds_before


def loss_port(batch_values, preds, preds_before, ignore_first_continuity=False,
              moment_weights=None, eps=1e-10, variance_method="direct"):
    """Restatement of nj_ode_loss (jump_ode.py:295-383), one trajectory at a time."""
    if variance_method not in ("direct", "second_moment"):
        raise ValueError(f"Unknown variance_method: {variance_method}")
    per_traj = []
    for X, Y, Yb in zip(batch_values, preds, preds_before):
        n = Y.shape[0]
        keep = torch.ones(n, dtype=Y.dtype)
        if ignore_first_continuity and n > 0:
            keep[0] = 0.0                                               # jump_ode.py:315-317

        def pair(target, got, target_b, got_b):
            a = ((target - got) ** 2).sum(dim=1)
            c = ((target_b - got_b) ** 2).sum(dim=1) * keep
            return ((torch.sqrt(a + eps) + torch.sqrt(c + eps)) ** 2).mean()

        w0 = 1.0 if moment_weights is None else moment_weights[0]
        total = w0 * pair(X, Y[:, :, 0], X, Yb[:, :, 0])                # jump_ode.py:304-325
        if Y.shape[2] > 1:
            W, Wb = Y[:, :, 1], Yb[:, :, 1]
            if variance_method == "direct":                             # jump_ode.py:333-344
                tgt = (X - Y[:, :, 0].detach()) ** 2
                tgt_b = (X - Yb[:, :, 0].detach()) ** 2
                got, got_b = W ** 2, Wb ** 2
            else:                                                       # jump_ode.py:346-353
                tgt = tgt_b = X ** 2
                got, got_b = W, Wb
            w1 = 1.0 if moment_weights is None else moment_weights[1]
            total = total + w1 * pair(tgt, got, tgt_b, got_b)           # jump_ode.py:362-378
        per_traj.append(total)
    return torch.stack(per_traj).mean()                                 # jump_ode.py:383


# --------------------------------------------------------------------------- #
# (3) interval-flattened vectorised restatement (float32 or float64)
# --------------------------------------------------------------------------- #

def pack(batch_times, batch_values):
    """lists -> (times (N,), values (N,d_x), offsets (B+1,)) numpy arrays."""
    n = [int(t.shape[0]) for t in batch_times]
    off = np.zeros(len(n) + 1, dtype=np.int64)
    off[1:] = np.cumsum(n)
    times = np.concatenate([np.asarray(t, dtype=np.float32).reshape(-1) for t in batch_times]) \
        if n else np.zeros(0, np.float32)
    vals = np.concatenate([np.asarray(v, dtype=np.float32).reshape(len(v), -1) for v in batch_values]) \
        if n else np.zeros((0, 1), np.float32)
    return times, vals, off


def flat_schedule(times, offsets, dt):
    """For every observation o: K[o] Euler steps and the (K[o]+1) float32 time knots
    t_0=t_o ... t_K=t_{o+1} (K=0 and has_next=False for the last observation of a trajectory)."""
    N = len(times)
    K = np.zeros(N, dtype=np.int32)
    has_next = np.zeros(N, dtype=bool)
    knots = [None] * N
    for b in range(len(offsets) - 1):
        lo, hi = int(offsets[b]), int(offsets[b + 1])
        for o in range(lo, hi):
            if o < hi - 1:
                sched = euler_schedule(times[o], times[o + 1], dt)
                K[o] = len(sched)
                has_next[o] = True
                knots[o] = np.array([times[o]] + [p[1] for p in sched], dtype=np.float32)
            else:
                knots[o] = np.array([times[o]], dtype=np.float32)
    return K, has_next, knots


def forward_flat(P, cfg, times, values, offsets, dtype=torch.float64):
    """All (trajectory, observation) units integrated side by side: the jump resets the
    hidden state from x_i only (jump_ode.py:169/:176), so units are independent IVPs.

    Returns packed preds, preds_before of shape (N, d_y, M) with autograd history to P.
    The step schedule and the dt/t features are computed in float32 exactly as the
    reference does and only then promoted to ``dtype``.
    """
    act, scale = _act(cfg), _scale(cfg)
    L, M, dy, H = cfg["n_hidden_layers"], cfg["num_moments"], cfg["output_dim"], cfg["hidden_dim"]
    S = n_stacks(cfg)
    N = len(times)
    K, has_next, knots = flat_schedule(times, offsets, cfg["dt_ode_step"])
    Kmax = int(K.max()) if N else 0
    # (N, Kmax+1) knot table, padded by repeating the last knot (zero-length steps are masked)
    T = np.zeros((N, Kmax + 1), dtype=np.float32)
    for o in range(N):
        T[o, :len(knots[o])] = knots[o]
        T[o, len(knots[o]):] = knots[o][-1]
    t_cur = torch.from_numpy(T[:, :-1].copy())
    delta = torch.from_numpy((T[:, 1:] - T[:, :-1]).astype(np.float32))   # fl32(t_new - t_cur)
    active = torch.from_numpy(np.arange(Kmax)[None, :] < K[:, None])
    x = torch.from_numpy(np.asarray(values, dtype=np.float32)).to(dtype)
    Pd = {k: v.to(dtype) for k, v in P.items()}

    outs, outs_before = [], []
    for s in range(S):
        pj, po, pq = stack_prefixes(cfg, s)
        kj, ko, kq = layer_keys(pj, L), layer_keys(po, L), layer_keys(pq, L)
        h = _jump(Pd, kj, act, x)
        outs.append(_out(Pd, kq, act, h))
        for k in range(Kmax):
            tc = t_cur[:, k:k + 1].to(dtype)
            de = delta[:, k:k + 1].to(dtype)
            v = torch.cat([scale(h), scale(x), tc, de], dim=-1)
            for i, (wk, bk) in enumerate(ko):
                v = F.linear(v, Pd[wk], Pd[bk])
                if i < len(ko) - 1:
                    v = act(v)
            h = torch.where(active[:, k:k + 1], h + de * v, h)
        outs_before.append(_out(Pd, kq, act, h))
    if cfg["shared_network"]:
        y = outs[0].view(N, dy, M)
        yb_src = outs_before[0].view(N, dy, M)
    else:
        y = torch.stack(outs, dim=-1)
        yb_src = torch.stack(outs_before, dim=-1)
    # preds_before[o+1] = readout(h_end of unit o); first observation of each trajectory is 0
    hn = torch.from_numpy(has_next)
    yb = torch.zeros_like(y)
    if N > 1:
        shifted = torch.where(hn[:-1, None, None], yb_src[:-1], torch.zeros_like(yb_src[:-1]))
        yb = torch.cat([torch.zeros_like(y[:1]), shifted], dim=0)
    return y, yb, K


def loss_flat(values, offsets, preds, preds_before, ignore_first_continuity=False,
              moment_weights=None, eps=1e-10, variance_method="direct"):
    """Vectorised nj_ode_loss on packed tensors (same formulae as ``loss_port``)."""
    if variance_method not in ("direct", "second_moment"):
        raise ValueError(f"Unknown variance_method: {variance_method}")
    dtype = preds.dtype
    X = torch.as_tensor(np.asarray(values, dtype=np.float32)).to(dtype)
    N = X.shape[0]
    B = len(offsets) - 1
    n = np.diff(np.asarray(offsets)).astype(np.int64)
    traj = torch.from_numpy(np.repeat(np.arange(B), n))
    first = torch.zeros(N, dtype=torch.bool)
    first[torch.from_numpy(np.asarray(offsets[:-1], dtype=np.int64)[n > 0])] = True
    keep = torch.ones(N, dtype=dtype)
    if ignore_first_continuity:
        keep = keep.masked_fill(first, 0.0)
    inv_n = torch.from_numpy(1.0 / np.maximum(n, 1)).to(dtype)[traj]

    def pair(target, got, target_b, got_b):
        a = ((target - got) ** 2).sum(dim=1)
        c = ((target_b - got_b) ** 2).sum(dim=1) * keep
        per_obs = (torch.sqrt(a + eps) + torch.sqrt(c + eps)) ** 2
        return torch.zeros(B, dtype=dtype).index_add(0, traj, per_obs * inv_n)

    w0 = 1.0 if moment_weights is None else float(moment_weights[0])
    total = w0 * pair(X, preds[:, :, 0], X, preds_before[:, :, 0])
    if preds.shape[2] > 1:
        W, Wb = preds[:, :, 1], preds_before[:, :, 1]
        if variance_method == "direct":
            tgt = (X - preds[:, :, 0].detach()) ** 2
            tgt_b = (X - preds_before[:, :, 0].detach()) ** 2
            got, got_b = W ** 2, Wb ** 2
        else:
            tgt = tgt_b = X ** 2
            got, got_b = W, Wb
        w1 = 1.0 if moment_weights is None else float(moment_weights[1])
        total = total + w1 * pair(tgt, got, tgt_b, got_b)
    return total.mean()


def run_flat(P, cfg, batch_times, batch_values, loss_kwargs=None, dtype=torch.float64):
    """Convenience: forward + loss + all parameter gradients with the flattened oracle.
    Returns dict(preds, preds_before, loss, grads{name: tensor}, K)."""
    loss_kwargs = dict(loss_kwargs or {})
    times, values, offsets = pack(batch_times, batch_values)
    Pg = {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in P.items()}
    y, yb, K = forward_flat(Pg, cfg, times, values, offsets, dtype=dtype)
    loss = loss_flat(values, offsets, y, yb, **loss_kwargs)
    names = list(Pg)
    gr = torch.autograd.grad(loss, [Pg[k] for k in names], allow_unused=True)
    grads = {k: (torch.zeros_like(Pg[k]) if g is None else g) for k, g in zip(names, gr)}
    return dict(preds=y.detach(), preds_before=yb.detach(), loss=loss.detach(), grads=grads, K=K,
                times=times, values=values, offsets=offsets)


def run_port(P, cfg, batch_times, batch_values, loss_kwargs=None):
    """Forward + loss + backward with the eager port (float32); the timed CPU baseline."""
    loss_kwargs = dict(loss_kwargs or {})
    Pg = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    log = []
    preds, preds_before = forward_port(Pg, cfg, batch_times, batch_values, step_log=log)
    loss = loss_port(batch_values, preds, preds_before, **loss_kwargs)
    loss.backward()
    grads = {k: (torch.zeros_like(v) if v.grad is None else v.grad) for k, v in Pg.items()}
    return dict(preds=[p.detach() for p in preds], preds_before=[p.detach() for p in preds_before],
                loss=loss.detach(), grads=grads, step_log=log)
