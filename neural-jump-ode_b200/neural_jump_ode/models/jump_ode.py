"""Neural Jump ODE model and loss -- B200-native drop-in for the reference module of the same
name (reference: neural_jump_ode/models/jump_ode.py).

Same import surface, constructor, ``forward(batch_times, batch_values)`` and ``nj_ode_loss``
signature, sub-module names and ``state_dict`` keys as the reference; the arithmetic of the hot
path (forward sweep, loss, reverse sweep) runs in hand-written sm_100a CUDA kernels behind the
C-ABI in ``include/njode.h`` through ``torch.autograd.Function``s.  There is no CPU fallback:
calling the hot path with CPU parameters raises ``RuntimeError``.

The small sub-modules (``JumpNN``, ``ODEFunc``, ``OutputNN``) and ``euler_step`` stay callable on
``(1, d)`` tensors because the reference's plotting code drives them directly
(utils/plotting.py:153-256); they hold the parameters, the kernels read them.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .. import _native as nat
from ..packed import PackedBatch, PredList

# name -> module class; anything else silently means ReLU, as in the reference (jump_ode.py:6-13, :18)
ACTIVATION_FUNCTIONS = {
    "relu": nn.ReLU, "tanh": nn.Tanh, "sigmoid": nn.Sigmoid,
    "elu": nn.ELU, "leaky_relu": nn.LeakyReLU, "selu": nn.SELU,
}


def _activation(name: str):
    return ACTIVATION_FUNCTIONS.get(str(name).lower(), nn.ReLU)


def _activation_code(name: str) -> int:
    return nat.ACT.get(str(name).lower(), nat.ACT["relu"])


def _linear_stack(widths: Sequence[int], act, dropout_rate: float, pattern: str) -> nn.Sequential:
    """Sequential whose Linear layers sit at indices 0,3,6,... so that state_dict keys equal the
    reference's (``net.{3i}.weight``).  ``pattern`` places activation/dropout around the Linears:
      'jump'  Lin act | drop Lin act | ...            (jump_ode.py:19-21)
      'ode'   Lin act | drop Lin act | ... | drop Lin (jump_ode.py:36-39)
      'out'   Lin act drop | ... | Lin                (jump_ode.py:72-73)
    """
    mods: List[nn.Module] = []
    n = len(widths) - 1
    for i in range(n):
        lin = nn.Linear(widths[i], widths[i + 1])
        last = i == n - 1
        if pattern == "out":
            mods += [lin] if last else [lin, act(), nn.Dropout(p=dropout_rate)]
        else:
            if i > 0:
                mods.append(nn.Dropout(p=dropout_rate))
            mods.append(lin)
            if pattern == "jump" or not last:
                mods.append(act())
    return nn.Sequential(*mods)


class JumpNN(nn.Module):
    """x_i -> hidden state right after an observation (reference jump_ode.py:15-26)."""

    def __init__(self, input_dim, hidden_dim, n_hidden_layers=1, activation="relu", dropout_rate=0.0):
        super().__init__()
        widths = [input_dim] + [hidden_dim] * (n_hidden_layers + 1)
        self.net = _linear_stack(widths, _activation(activation), dropout_rate, "jump")

    def forward(self, x):
        return self.net(x)


class ODEFunc(nn.Module):
    """dh/dt = f(s(h), s(x_last), t_last, t - t_last) (reference jump_ode.py:29-63)."""

    def __init__(self, hidden_dim, input_dim, n_hidden_layers=1, activation="relu", dropout_rate=0.0,
                 input_scaling="identity"):
        super().__init__()
        widths = [hidden_dim + input_dim + 2] + [hidden_dim] * (n_hidden_layers + 1)
        self.net = _linear_stack(widths, _activation(activation), dropout_rate, "ode")
        if input_scaling in ("identity", "none"):
            self.scaling_fn = nn.Identity()
        elif input_scaling == "tanh":
            self.scaling_fn = nn.Tanh()
        elif input_scaling == "sigmoid":
            self.scaling_fn = nn.Sigmoid()
        else:
            raise ValueError(f"Unknown input_scaling: {input_scaling}. Use 'identity', 'tanh', or 'sigmoid'.")

    def forward(self, t, h, x_last, t_last):
        ones = torch.ones_like(h[..., :1])
        feats = [self.scaling_fn(h), self.scaling_fn(x_last), t_last * ones, (t - t_last) * ones]
        return self.net(torch.cat(feats, dim=-1))


class OutputNN(nn.Module):
    """hidden state -> prediction(s) (reference jump_ode.py:66-77)."""

    def __init__(self, hidden_dim, output_dim, n_hidden_layers=1, activation="relu", dropout_rate=0.0):
        super().__init__()
        widths = [hidden_dim] * (n_hidden_layers + 1) + [output_dim]
        self.net = _linear_stack(widths, _activation(activation), dropout_rate, "out")

    def forward(self, h):
        return self.net(h)


# ------------------------------------------------------------------------------------------------
# autograd glue: forward sweep / reverse sweep
# ------------------------------------------------------------------------------------------------

class _SweepFunction(torch.autograd.Function):
    """preds, preds_before = sweep(batch; params).  Forward = ``njode_forward`` (writes per-step
    hidden-state checkpoints when a gradient will be needed), backward = ``njode_backward``."""

    @staticmethod
    def forward(ctx, desc, batch: PackedBatch, sched, want_grad: bool, dp_group, *params):
        lib = nat.load()
        dev = batch.device
        N, B = batch.N, batch.B
        d_y, M, H = desc.d_y, desc.num_moments, desc.hidden
        S = 1 if desc.shared_network else M
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            flat = torch.cat([p.detach().reshape(-1) for p in params]).float()
            preds = torch.empty((N, d_y, M), dtype=torch.float32, device=dev)
            before = torch.empty((N, d_y, M), dtype=torch.float32, device=dev)
            ckpt = None
            if want_grad:
                row_floats = lib.njode_ckpt_row_floats(desc)
                if row_floats < 0:
                    raise RuntimeError("njode_ckpt_row_floats: " + lib.njode_last_error().decode(errors="replace"))
                ckpt = torch.empty(S * sched.total_slots * sched.tile_rows * row_floats, dtype=torch.float32, device=dev)
            ws_bytes = lib.njode_forward_workspace_bytes(desc)
            ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
            nat.check(lib.njode_forward(desc, nat.ptr(flat), nat.ptr(batch.times), nat.ptr(batch.values),
                                        nat.ptr(batch.offsets), B, N, nat.ptr(sched.kenc), nat.ptr(sched.perm),
                                        nat.ptr(sched.tile_kmax), nat.ptr(sched.tile_slot_off), nat.ptr(sched.knots),
                                        sched.n_tiles, sched.total_slots, sched.tile_rows,
                                        nat.ptr(preds), nat.ptr(before), nat.ptr(ckpt), nat.ptr(ws), ws_bytes, stream),
                      "njode_forward")
        ctx.desc, ctx.batch, ctx.sched = desc, batch, sched
        ctx.dp_group = dp_group
        ctx.shapes = [p.shape for p in params]
        ctx.flat, ctx.ckpt = flat, ckpt
        return preds, before

    @staticmethod
    def backward(ctx, g_preds, g_before):
        if ctx.ckpt is None:
            raise RuntimeError("NeuralJumpODE: backward requested but the forward sweep ran without checkpoints")
        lib = nat.load()
        desc, batch, sched = ctx.desc, ctx.batch, ctx.sched
        dev = batch.device
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            g_preds = g_preds.contiguous().float()
            g_before = g_before.contiguous().float()
            grad_flat = torch.empty_like(ctx.flat)
            ws_bytes = lib.njode_backward_workspace_bytes(desc, sched.n_tiles)
            ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
            nat.check(lib.njode_backward(desc, nat.ptr(ctx.flat), nat.ptr(batch.times), nat.ptr(batch.values),
                                         nat.ptr(batch.offsets), batch.B, batch.N, nat.ptr(sched.kenc),
                                         nat.ptr(sched.perm), nat.ptr(sched.tile_kmax), nat.ptr(sched.tile_slot_off),
                                         nat.ptr(sched.knots), sched.n_tiles, sched.total_slots, sched.tile_rows,
                                         nat.ptr(g_preds), nat.ptr(g_before), nat.ptr(ctx.ckpt), nat.ptr(grad_flat),
                                         nat.ptr(ws), ws_bytes, stream), "njode_backward")
        ctx.ckpt = None     # checkpoints are the big buffer: release them as soon as they are consumed
        if ctx.dp_group is not None:
            # data parallel: the reverse sweep left this rank's share of the gradient (the loss is scaled by
            # 1/B_global) in ONE flat buffer -- sum it over the ranks in place, no gather / scatter copies
            import torch.distributed as dist
            dist.all_reduce(grad_flat, op=dist.ReduceOp.SUM, group=None if ctx.dp_group is True else ctx.dp_group)
        # Stacks of moments >= 2 get an all-zero gradient from nj_ode_loss (jump_ode.py:328-378); the
        # reference reports zero tensors for them too (torch.stack backward), so nothing is special-cased.
        grads, o = [], 0
        for shp in ctx.shapes:
            n = shp.numel()
            grads.append(grad_flat[o:o + n].view(shp))
            o += n
        return (None, None, None, None, None, *grads)


class _LossFunction(torch.autograd.Function):
    """nj_ode_loss value; d loss / d preds and d loss / d preds_before are closed-form and produced by
    the same kernel pass (``njode_loss``)."""

    @staticmethod
    def forward(ctx, ldesc, batch: PackedBatch, traj_scale: float, want_grad: bool, preds, before):
        lib = nat.load()
        dev = preds.device
        N, d, M = preds.shape
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            preds_c = preds.detach().contiguous().float()
            before_c = before.detach().contiguous().float()
            loss = torch.empty((), dtype=torch.float32, device=dev)
            gp = torch.empty_like(preds_c) if want_grad else None
            gb = torch.empty_like(before_c) if want_grad else None
            ws_bytes = lib.njode_loss_workspace_bytes(batch.B)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            nat.check(lib.njode_loss(ldesc, nat.ptr(batch.values), nat.ptr(preds_c), nat.ptr(before_c),
                                     nat.ptr(batch.offsets), batch.B, N, d, M, float(traj_scale),
                                     nat.ptr(loss), nat.ptr(gp), nat.ptr(gb), nat.ptr(ws), ws_bytes, stream),
                      "njode_loss")
        ctx.gp, ctx.gb = gp, gb
        return loss

    @staticmethod
    def backward(ctx, g):
        if ctx.gp is None:
            raise RuntimeError("nj_ode_loss: backward requested but gradients were not computed")
        return None, None, None, None, ctx.gp * g, ctx.gb * g


# ------------------------------------------------------------------------------------------------
# the model
# ------------------------------------------------------------------------------------------------

class NeuralJumpODE(nn.Module):
    """Neural Jump ODE (reference jump_ode.py:80-233) with the hot path on B200.

    Constructor arguments are the reference's, in the reference's positional order.
    ``n_steps_between`` is accepted and ignored: the reference's README / tests still pass it
    although its constructor dropped it.
    """

    def __init__(self, input_dim, hidden_dim, output_dim,
                 dt_between_obs=None, dt_ode_step=None, num_moments=1, n_hidden_layers=1, activation="relu",
                 shared_network=False, dropout_rate=0.0, input_scaling="identity", variance_method="direct",
                 n_steps_between=None):
        super().__init__()
        self.num_moments = num_moments
        self.shared_network = shared_network
        self.variance_method = variance_method
        mk = dict(n_hidden_layers=n_hidden_layers, activation=activation, dropout_rate=dropout_rate)
        if shared_network:
            self.jump_nn = JumpNN(input_dim, hidden_dim, **mk)
            self.ode_func = ODEFunc(hidden_dim, input_dim, input_scaling=input_scaling, **mk)
            self.output_nn = OutputNN(hidden_dim, output_dim * num_moments, **mk)
            self.jump_nns = self.ode_funcs = self.output_nns = None
        else:
            self.jump_nns = nn.ModuleList(JumpNN(input_dim, hidden_dim, **mk) for _ in range(num_moments))
            self.ode_funcs = nn.ModuleList(
                ODEFunc(hidden_dim, input_dim, input_scaling=input_scaling, **mk) for _ in range(num_moments))
            self.output_nns = nn.ModuleList(OutputNN(hidden_dim, output_dim, **mk) for _ in range(num_moments))
            self.jump_nn = self.ode_func = self.output_nn = None
        self.dt_ode_step = dt_ode_step
        self.dt_between_obs = dt_between_obs      # deprecated in the reference, unused
        self.output_dim = output_dim
        # host-side description of the kernels' view of this module
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim
        self.n_hidden_layers = n_hidden_layers
        self.activation = activation
        self.input_scaling = input_scaling
        self.dropout_rate = dropout_rate
        self.kernel_impl = "auto"                 # 'auto' | 'generic' | 'rowtile' | 'tiled' (testing / profiling knob)
        self._dp_group = None                     # see enable_data_parallel

    def enable_data_parallel(self, group=True):
        """Sum the parameter gradients over the ranks of ``group`` (``True`` = the default process group, ``None``
        = off) inside the reverse sweep: one in-place all-reduce of the flat gradient buffer per backward.  Every
        rank must hold the same parameters, integrate its own slice of the batch and scale its loss with
        ``nj_ode_loss(..., traj_scale=1 / B_global)`` (see ``neural_jump_ode.sharding``)."""
        self._dp_group = group
        return self

    # -- reference-compatible small-tensor API (plotting, tests) ---------------------------------
    def euler_step(self, h_list, x_last, t_last, t_next):
        """One explicit Euler step for every stack (reference jump_ode.py:122-140)."""
        funcs = [self.ode_func] if self.shared_network else list(self.ode_funcs)
        step = t_next - t_last
        return [h + step * f(t_next, h, x_last, t_last) for f, h in zip(funcs, h_list)]

    # -- kernel plumbing -----------------------------------------------------------------------------
    def _stacks(self):
        if self.shared_network:
            return [(self.jump_nn, self.ode_func, self.output_nn)]
        return list(zip(self.jump_nns, self.ode_funcs, self.output_nns))

    def flat_parameters(self) -> List[nn.Parameter]:
        """Parameters in the flat order of include/njode.h: stack-major; jump, ode, out; layer; W then b."""
        out = []
        for nets in self._stacks():
            for net in nets:
                for i in range(self.n_hidden_layers + 1):
                    lin = net.net[3 * i]
                    out += [lin.weight, lin.bias]
        return out

    def descriptor(self) -> "nat.NjodeDesc":
        d = nat.NjodeDesc()
        d.d_x, d.d_y, d.hidden = self.input_dim, self.output_dim, self.hidden_dim
        d.n_hidden_layers, d.num_moments = self.n_hidden_layers, self.num_moments
        d.shared_network = 1 if self.shared_network else 0
        d.activation = _activation_code(self.activation)
        d.input_scaling = nat.SCALE[self.input_scaling]
        d.has_dt = 0 if self.dt_ode_step is None else 1
        d.dt = 0.0 if self.dt_ode_step is None else float(self.dt_ode_step)
        d.impl = nat.IMPL[self.kernel_impl]
        return d

    def _check_runnable(self, params):
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("NeuralJumpODE (B200 build): the forward/backward hot path runs on CUDA only; "
                               "there is no CPU fallback. Call model.to('cuda').")
        if self.dropout_rate and self.training:
            raise NotImplementedError("dropout_rate > 0 in training mode is not supported by the fused kernels "
                                      "(no BASELINE configuration uses it); use model.eval() or dropout_rate=0")
        for p in params:
            if p.dtype != torch.float32 or p.device != dev:
                raise RuntimeError("NeuralJumpODE: all parameters must be float32 on one CUDA device")
        return dev

    def pack(self, batch_times, batch_values) -> PackedBatch:
        dev = next(self.parameters()).device
        return PackedBatch.from_lists(batch_times, batch_values, device=dev)

    def forward_packed(self, batch: PackedBatch) -> Tuple[torch.Tensor, torch.Tensor]:
        """Packed fast path: returns packed ``preds, preds_before`` of shape (N, d_y, M)."""
        params = self.flat_parameters()
        dev = self._check_runnable(params)
        if batch.device != dev:
            batch = batch.to(dev)
        if batch.values.shape[1] != self.input_dim:
            raise ValueError(f"values have d_x={batch.values.shape[1]}, model expects {self.input_dim}")
        desc = self.descriptor()
        sched = batch.schedule(desc)
        want_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _SweepFunction.apply(desc, batch, sched, want_grad, self._dp_group, *params)

    def forward(self, batch_times, batch_values=None):
        """batch_times / batch_values: lists of (n_i,) / (n_i, d_x) tensors (reference jump_ode.py:218-233),
        or a ``PackedBatch`` as the first argument.  Returns two lists of (n_i, d_y, M) tensors."""
        batch = batch_times if isinstance(batch_times, PackedBatch) else self.pack(batch_times, batch_values)
        preds, before = self.forward_packed(batch)
        return PredList(preds, batch), PredList(before, batch)

    def forward_single(self, times, values):
        """One trajectory (reference jump_ode.py:142-216)."""
        preds, before = self.forward([times], [values])
        return preds[0], before[0]


# ------------------------------------------------------------------------------------------------
# the loss
# ------------------------------------------------------------------------------------------------

_MW_CACHE = {}


def _moment_weights(moment_weights, M):
    """(w0, w1) as Python floats.  A device tensor (the reference Trainer passes one, training.py:24)
    is read back once and cached by identity/version."""
    if moment_weights is None:
        return 1.0, 1.0
    if torch.is_tensor(moment_weights):
        key = (id(moment_weights), moment_weights._version)
        w = _MW_CACHE.get(key)
        if w is None:
            if len(_MW_CACHE) > 64:
                _MW_CACHE.clear()
            w = _MW_CACHE[key] = [float(v) for v in moment_weights.detach().cpu().tolist()]
    else:
        w = [float(v) for v in moment_weights]
    if len(w) < min(M, 2):
        raise IndexError("moment_weights has fewer entries than moments used by the loss")
    return w[0], (w[1] if M > 1 else 1.0)


def _as_packed(x, batch: Optional[PackedBatch], what: str) -> torch.Tensor:
    if torch.is_tensor(x):
        return x
    if isinstance(x, PredList) and batch is not None and x.batch is batch and x.untouched():
        return x.packed
    if len(x) == 0:
        raise ValueError(f"nj_ode_loss: empty {what}")
    return torch.cat(list(x), dim=0)


def nj_ode_loss(batch_times, batch_values, preds, preds_before,
                ignore_first_continuity: bool = False, moment_weights=None, eps: float = 1e-10,
                variance_method: str = "direct", traj_scale: Optional[float] = None):
    """Neural Jump ODE loss (reference jump_ode.py:235-383), evaluated by ``njode_loss`` on the device.

    ``preds`` / ``preds_before`` may be the lists returned by the model, arbitrary lists of
    (n_i, d, M) CUDA tensors (e.g. closed-form moments, utils/training.py:250), or packed (N, d, M)
    tensors; ``batch_times`` may be a ``PackedBatch`` (then ``batch_values`` is ignored).
    ``traj_scale`` (default 1/B) lets data-parallel ranks weight by the global batch size.
    """
    if variance_method not in nat.VAR:
        raise ValueError(f"Unknown variance_method: {variance_method}")
    batch = None
    if isinstance(batch_times, PackedBatch):
        batch = batch_times
    elif isinstance(preds, PredList) and preds.batch.came_from(batch_times, batch_values):
        batch = preds.batch
    p = _as_packed(preds, batch, "preds")
    pb = _as_packed(preds_before, batch, "preds_before")
    if not p.is_cuda:
        raise RuntimeError("nj_ode_loss (B200 build) runs on CUDA only; there is no CPU fallback")
    if batch is None:
        batch = PackedBatch.from_lists(batch_times, batch_values, device=p.device)
    elif batch.device != p.device:
        batch = batch.to(p.device)
    if p.dim() != 3 or p.shape != pb.shape or p.shape[0] != batch.N or p.shape[1] != batch.values.shape[1]:
        raise ValueError(f"nj_ode_loss: preds {tuple(p.shape)} / preds_before {tuple(pb.shape)} do not match "
                         f"values {tuple(batch.values.shape)}")
    M = p.shape[2]
    ld = nat.NjodeLossDesc()
    ld.ignore_first_continuity = 1 if ignore_first_continuity else 0
    ld.variance_method = nat.VAR[variance_method]
    ld.eps = float(eps)
    ld.w0, ld.w1 = _moment_weights(moment_weights, M)
    scale = 1.0 / batch.B if traj_scale is None else float(traj_scale)
    want_grad = torch.is_grad_enabled() and (p.requires_grad or pb.requires_grad)
    return _LossFunction.apply(ld, batch, scale, want_grad, p, pb)
