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

import weakref
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .. import _native as nat
from ..packed import PackedBatch, PredList, Schedule, pinned_header

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

class _SweepState:
    """What a reverse sweep needs, shared by the sweep's autograd node and -- through ``preds._njode_state`` -- by
    ``nj_ode_loss``, which may run the reverse sweep EARLY (see ``_LossFunction``)."""
    __slots__ = ("desc", "batch", "sched", "dp_group", "shapes", "flat", "ckpt", "versions", "params",
                 "eager_ok", "early_grad", "early_key", "consumed", "outputs")

    def current_versions(self):
        """autograd's version counters of the parameters plus the count of raw-pointer writes to their flat buffer
        (FlatAdam.step goes through njode_adam_step, which autograd does not see)."""
        return sum(p._version for p in self.params) + nat.generation(self.flat.data_ptr())

    def check_unmodified(self):
        if self.current_versions() != self.versions:
            raise RuntimeError("NeuralJumpODE: a parameter was modified in place between the forward sweep and its "
                               "backward (the reverse sweep reads the parameters where they live)")

    def hooked(self):
        """True when a tensor hook sits on the sweep's outputs: a hook may edit the gradient in place (same address),
        so the early reverse sweep's result must not be matched by address then."""
        for ref in self.outputs or ():
            t = ref()
            if t is not None and getattr(t, "_backward_hooks", None):
                return True
        return False

    def reverse_sweep(self, g_preds, g_before):
        """njode_backward (+ the data-parallel all-reduce): the flat parameter gradient for these output gradients."""
        if self.ckpt is None:
            if self.consumed:
                raise RuntimeError("NeuralJumpODE: the checkpoints of this forward sweep were already consumed by a "
                                   "backward pass and released (they are the large buffer); a second backward through "
                                   "the same predictions -- retain_graph=True, or another loss -- needs a new forward")
            raise RuntimeError("NeuralJumpODE: backward requested but the forward sweep ran without checkpoints")
        self.check_unmodified()
        lib = nat.load()
        desc, batch, sched = self.desc, self.batch, self.sched
        dev = batch.device
        with nat.on_device(dev):
            stream = nat.current_stream(dev)
            g_preds = g_preds.contiguous().float()
            g_before = g_before.contiguous().float()
            grad_flat = torch.empty_like(self.flat)
            ws_bytes = lib.njode_backward_workspace_bytes(desc, sched.n_tiles)
            ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
            nat.check(lib.njode_backward(desc, nat.ptr(self.flat), nat.ptr(batch.times), nat.ptr(batch.values),
                                         nat.ptr(batch.offsets), batch.B, batch.N, *sched.ptrs,
                                         sched.n_tiles, sched.total_slots, sched.tile_rows,
                                         nat.ptr(g_preds), nat.ptr(g_before), nat.ptr(self.ckpt), nat.ptr(grad_flat),
                                         nat.ptr(ws), ws_bytes, stream), "njode_backward")
        if self.dp_group is not None:
            # data parallel: the reverse sweep left this rank's share of the gradient (the loss is scaled by
            # 1/B_global) in ONE flat buffer -- sum it over the ranks in place, no gather / scatter copies
            import torch.distributed as dist
            dist.all_reduce(grad_flat, op=dist.ReduceOp.SUM, group=None if self.dp_group is True else self.dp_group)
        return grad_flat


class _BatchPlan:
    """What the sweep needs to know about a batch before it runs (see NeuralJumpODE._begin_batch)."""
    __slots__ = ("tile_rows", "n_tiles", "row_floats", "key", "sched", "slots", "arena", "layout", "scratch", "scratch_bytes", "host")


def _alloc_ckpt(model, n_floats: int, dev):
    """Checkpoint buffer of a sweep.  Normally a fresh allocation (the caching allocator recycles it); inside
    ``forward_backward_waves`` one grow-only buffer is shared by all waves: consecutive waves need slightly different
    sizes (41 GB +- a few MB at hidden 128), which the allocator can only serve by freeing and re-allocating device
    memory -- a device synchronisation and ~4 ms per wave.  Re-use is safe because the next wave's forward sweep is
    enqueued behind the previous wave's reverse sweep on the same stream."""
    n_floats = max(int(n_floats), 1)
    if not getattr(model, "_ckpt_pool_on", False):
        return torch.empty(n_floats, dtype=torch.float32, device=dev)
    pool = getattr(model, "_ckpt_pool", None)
    if pool is None or pool.numel() < n_floats or pool.device != torch.device(dev):
        model._ckpt_pool = pool = None                       # release the old one before growing
        pool = torch.empty(n_floats + n_floats // 32 + 4096, dtype=torch.float32, device=dev)
        model._ckpt_pool = pool
    return pool[:n_floats]


class _SweepFunction(torch.autograd.Function):
    """preds, preds_before = sweep(batch; params).  Forward = ``njode_forward`` on a batch whose schedule is cached,
    ``njode_forward_batch_begin`` / ``_finish`` (schedule + knots + sweep, no Python between the schedule's host sync
    and the sweep) on a new one; per-step hidden-state checkpoints are written when a gradient will be needed.
    Backward = ``njode_backward``."""

    @staticmethod
    def forward(ctx, model, desc, batch: PackedBatch, want_grad: bool, plan, *params):
        lib = nat.load()
        dev = batch.device
        N, B = batch.N, batch.B
        d_y, M = desc.d_y, desc.num_moments
        S = 1 if desc.shared_network else M
        tile_rows, n_tiles, row_floats, key, sched = plan.tile_rows, plan.n_tiles, plan.row_floats, plan.key, plan.sched
        with nat.on_device(dev):
            stream = nat.current_stream(dev)
            flat = model._flat_view(params)
            out = torch.empty((2, N, d_y, M), dtype=torch.float32, device=dev)
            preds, before = out[0], out[1]
            ckpt = None
            if sched is not None:
                if want_grad:
                    ckpt = _alloc_ckpt(model, S * sched.total_slots * tile_rows * row_floats, dev)
                ws_bytes = lib.njode_forward_workspace_bytes(desc)
                ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
                nat.check(lib.njode_forward(desc, nat.ptr(flat), nat.ptr(batch.times), nat.ptr(batch.values),
                                            nat.ptr(batch.offsets), B, N, *sched.ptrs,
                                            n_tiles, sched.total_slots, tile_rows,
                                            nat.ptr(preds), nat.ptr(before), nat.ptr(ckpt), nat.ptr(ws), ws_bytes, stream),
                          "njode_forward")
            else:
                # new batch: its schedule is being built on the device since _begin_batch (njode_forward_batch_begin);
                # arena and checkpoints are sized from a guess of its slot count (exact when a batch of this shape
                # was seen before) so that no Python runs between the schedule's host sync and the sweep
                slots, arena, layout, scratch, scratch_bytes, host = plan.slots, plan.arena, plan.layout, plan.scratch, plan.scratch_bytes, plan.host
                ckpt_floats = S * slots * tile_rows * row_floats if want_grad else 0
                ckpt = _alloc_ckpt(model, ckpt_floats, dev) if want_grad else None
                rc = lib.njode_forward_batch_finish(desc, nat.ptr(flat), nat.ptr(batch.times), nat.ptr(batch.values),
                                                    nat.ptr(batch.offsets), B, N, nat.ptr(arena), arena.numel(),
                                                    1 if want_grad else 0, nat.ptr(ckpt), ckpt_floats,
                                                    nat.ptr(scratch), scratch_bytes, host.data_ptr(),
                                                    nat.ptr(preds), nat.ptr(before), stream)
                if rc == nat.ECAPACITY:
                    # the guess was too small; the schedule is known now: size exactly and run the one-call form
                    slots = int(host[nat.HDR_TOTAL_SLOTS])
                    arena = ckpt = None
                    arena = torch.empty(lib.njode_batch_arena_bytes(desc, B, N, slots, layout), dtype=torch.uint8, device=dev)
                    ckpt_floats = S * slots * tile_rows * row_floats if want_grad else 0
                    ckpt = _alloc_ckpt(model, ckpt_floats, dev) if want_grad else None
                    rc = lib.njode_forward_batch(desc, nat.ptr(flat), nat.ptr(batch.times), nat.ptr(batch.values),
                                                 nat.ptr(batch.offsets), B, N, nat.ptr(arena), arena.numel(),
                                                 1 if want_grad else 0, nat.ptr(ckpt), ckpt_floats,
                                                 nat.ptr(scratch), scratch_bytes, host.data_ptr(),
                                                 nat.ptr(preds), nat.ptr(before), stream)
                    nat.check(rc, "njode_forward_batch")
                else:
                    nat.check(rc, "njode_forward_batch_finish")
                sched = Schedule(arena, layout, N, tile_rows, n_tiles, host.tolist())
                model._note_slots(N, B, n_tiles, sched.total_slots)
                batch._schedules[key] = sched
        st = _SweepState()
        st.desc, st.batch, st.sched = desc, batch, sched
        st.dp_group = model._dp_group
        st.shapes = [p.shape for p in params]
        st.flat, st.ckpt = flat, ckpt
        st.params = params
        st.versions = st.current_versions()
        st.eager_ok = bool(model.eager_backward) and want_grad
        st.early_grad = st.early_key = None
        st.consumed, st.outputs = False, None
        ctx.state = st
        model._last_state = st
        return preds, before

    @staticmethod
    def backward(ctx, g_preds, g_before):
        st = ctx.state
        early = st.early_grad
        if (early is not None and st.early_key is not None
                and st.early_key[0] == (g_preds.data_ptr(), g_before.data_ptr()) and not st.hooked()):
            # nj_ode_loss already ran the reverse sweep on its un-scaled gradients, and what arrives here is exactly
            # those gradients times the loss's upstream gradient (same buffers, no hook in between): scale the result
            st.check_unmodified()
            grad_flat = early * st.early_key[1]
        else:
            grad_flat = st.reverse_sweep(g_preds, g_before)
        st.ckpt = st.early_grad = st.early_key = None     # checkpoints are the big buffer: release them once consumed
        st.consumed = True
        # Stacks of moments >= 2 get an all-zero gradient from nj_ode_loss (jump_ode.py:328-378); the
        # reference reports zero tensors for them too (torch.stack backward), so nothing is special-cased.
        grads = [g.view(shp) for g, shp in zip(grad_flat.split([shp.numel() for shp in st.shapes]), st.shapes)]
        return (None, None, None, None, None, *grads)


class _LossFunction(torch.autograd.Function):
    """nj_ode_loss value; d loss / d preds and d loss / d preds_before are closed-form and produced by
    the same kernel pass (``njode_loss``).

    When ``preds`` / ``preds_before`` come straight from the model, the reverse sweep is launched HERE, on those
    un-scaled gradients, instead of ~100 us later when the autograd engine reaches the sweep's node (thread hand-off,
    two Python nodes): the GPU no longer idles between the loss and the reverse sweep.  The sweep's backward node
    then only scales the finished parameter gradient by the loss's upstream gradient -- if what reaches it is not
    exactly this loss's gradient buffers (a second loss on the same predictions, a hook) it runs the reverse sweep
    itself as before.  ``model.eager_backward = False`` turns this off."""

    @staticmethod
    def forward(ctx, ldesc, batch: PackedBatch, traj_scale: float, want_grad: bool, state, preds, before):
        lib = nat.load()
        dev = preds.device
        N, d, M = preds.shape
        with nat.on_device(dev):
            stream = nat.current_stream(dev)
            preds_c = preds.detach().contiguous().float()
            before_c = before.detach().contiguous().float()
            loss = torch.empty((), dtype=torch.float32, device=dev)
            g = torch.empty((2, N, d, M), dtype=torch.float32, device=dev) if want_grad else None   # (d preds, d preds_before)
            ws_bytes = lib.njode_loss_workspace_bytes(batch.B)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            nat.check(lib.njode_loss(ldesc, nat.ptr(batch.values), nat.ptr(preds_c), nat.ptr(before_c),
                                     nat.ptr(batch.offsets), batch.B, N, d, M, float(traj_scale),
                                     nat.ptr(loss), nat.ptr(g[0]) if want_grad else None,
                                     nat.ptr(g[1]) if want_grad else None, nat.ptr(ws), ws_bytes, stream),
                      "njode_loss")
            if state is not None and want_grad and state.eager_ok and state.ckpt is not None and state.early_grad is None:
                state.early_grad = state.reverse_sweep(g[0], g[1])
            else:
                state = None
        ctx.g, ctx.state = g, state
        return loss

    @staticmethod
    def backward(ctx, g):
        if ctx.g is None:
            raise RuntimeError("nj_ode_loss: backward requested but gradients were not computed")
        scaled = ctx.g * g                      # one launch for both gradients
        gp, gb = scaled[0], scaled[1]
        if ctx.state is not None:
            ctx.state.early_key = ((gp.data_ptr(), gb.data_ptr()), g, scaled)    # (scaled kept alive: its address is the key)
        return None, None, None, None, None, gp, gb


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
        self.kernel_impl = "auto"                 # 'auto' | 'generic' | 'rowtile' | 'tiled' | 'wide' (testing / profiling knob)
        self._dp_group = None                     # see enable_data_parallel
        self.auto_flatten = True                  # see flatten_parameters
        self.eager_backward = True                # see _LossFunction
        self._last_state = None
        self._slots_memo, self._slots_per_tile = {}, 0.0

    def enable_data_parallel(self, group=True):
        """Sum the parameter gradients over the ranks of ``group`` (``True`` = the default process group, ``None``
        = off) inside the reverse sweep: one in-place all-reduce of the flat gradient buffer per backward.  Every
        rank must hold the same parameters, integrate its own slice of the batch and scale its loss with
        ``nj_ode_loss(..., traj_scale=1 / B_global)`` (see ``neural_jump_ode.sharding``)."""
        self._dp_group = group
        return self

    # -- parameters as one flat buffer ---------------------------------------------------------------
    def flatten_parameters(self):
        """Re-home every parameter as a view of ONE flat float32 buffer in the C-ABI's order (values, ``Parameter``
        objects and ``state_dict`` are unchanged -- the ``nn.RNN.flatten_parameters`` idea): the sweeps then read
        the parameters where they live instead of gathering 6 * (L + 1) tensors per call.  Done automatically by the
        first sweep while every parameter still owns its storage (``auto_flatten = False`` to opt out), redone after
        ``.to()``.  ``FlatAdam`` adopts this buffer when it is built afterwards."""
        params = self.flat_parameters()
        dev = params[0].device
        total = sum(p.numel() for p in params)
        flat = torch.empty(total, dtype=torch.float32, device=dev)
        o = 0
        with torch.no_grad():
            for p in params:
                k = p.numel()
                flat[o:o + k].copy_(p.detach().reshape(-1))
                p.data = flat[o:o + k].view(p.shape)
                o += k
        return self

    def _flat_view(self, params) -> torch.Tensor:
        """The parameters as one flat float32 tensor: a zero-copy view when they sit back to back in one storage
        (``flatten_parameters`` / ``FlatAdam``), else a gathered copy."""
        for attempt in (0, 1):
            p0 = params[0]
            addr, total, ok = p0.data_ptr(), 0, True
            for p in params:
                if p.data_ptr() != addr + 4 * total or not p.is_contiguous():
                    ok = False
                    break
                total += p.numel()
            if ok and p0.untyped_storage().nbytes() >= (p0.storage_offset() + total) * 4:
                return p0.detach().as_strided((total,), (1,), p0.storage_offset())
            # (never re-home parameters that are views of someone else's buffer, e.g. a FlatAdam built first)
            if attempt == 0 and self.auto_flatten and all(
                    p.is_leaf and p.storage_offset() == 0 and p.untyped_storage().nbytes() == 4 * p.numel() for p in params):
                self.flatten_parameters()
                continue
            break
        return torch.cat([p.detach().reshape(-1) for p in params]).float()

    def _begin_batch(self, desc, batch: PackedBatch):
        """Tiling of the batch for this model; if its schedule is not cached yet, launch the schedule build right away
        (``njode_forward_batch_begin``) so that it runs on the device while the host prepares the sweep."""
        lib = nat.load()
        plan = _BatchPlan()
        N, B = batch.N, batch.B
        plan.tile_rows = lib.njode_tile_rows(desc)
        plan.n_tiles = lib.njode_num_tiles(desc, N)
        plan.row_floats = lib.njode_ckpt_row_floats(desc)
        if plan.tile_rows < 1 or plan.n_tiles < 0 or plan.row_floats < 0:
            raise RuntimeError("NeuralJumpODE: " + lib.njode_last_error().decode(errors="replace"))
        plan.key = batch.schedule_key(desc, plan.tile_rows, plan.n_tiles)
        plan.sched = batch._schedules.get(plan.key)
        if plan.sched is None:
            dev = batch.device
            with nat.on_device(dev):
                plan.slots = self._guess_slots(N, B, plan.n_tiles)
                plan.layout = (nat.C.c_int64 * nat.ARENA_WORDS)()
                plan.arena = torch.empty(lib.njode_batch_arena_bytes(desc, B, N, plan.slots, plan.layout), dtype=torch.uint8, device=dev)
                plan.scratch_bytes = lib.njode_batch_scratch_bytes(desc, B, N)
                plan.scratch = torch.empty(plan.scratch_bytes, dtype=torch.uint8, device=dev)
                plan.host = pinned_header(dev)
                nat.check(lib.njode_forward_batch_begin(desc, nat.ptr(batch.times), nat.ptr(batch.offsets), B, N,
                                                        nat.ptr(plan.arena), plan.arena.numel(), nat.ptr(plan.scratch),
                                                        plan.scratch_bytes, plan.host.data_ptr(), nat.current_stream(dev)),
                          "njode_forward_batch_begin")
        return plan

    # -- checkpoint-slot guesses for batches whose schedule is not known yet (njode_forward_batch) -------
    def _guess_slots(self, N, B, n_tiles) -> int:
        exact = self._slots_memo.get((N, B))
        if exact is not None:
            return exact
        return int(self._slots_per_tile * n_tiles * 1.03) + 2 if self._slots_per_tile > 0 else 0

    def _note_slots(self, N, B, n_tiles, total_slots):
        if len(self._slots_memo) > 256:
            self._slots_memo.clear()
        self._slots_memo[(N, B)] = total_slots
        if n_tiles > 0:
            self._slots_per_tile = max(self._slots_per_tile, total_slots / n_tiles)

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
        plan = self._begin_batch(desc, batch)
        want_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        preds, before = _SweepFunction.apply(self, desc, batch, want_grad, plan, *params)
        if want_grad:
            preds._njode_state = before._njode_state = self._last_state     # lets nj_ode_loss start the reverse sweep early
            self._last_state.outputs = (weakref.ref(preds), weakref.ref(before))
        self._last_state = None
        return preds, before

    # -- dense-grid inference (plots) ----------------------------------------------------------------
    @torch.no_grad()
    def predict_on_grid(self, batch, grid_times: torch.Tensor, batch_values=None) -> torch.Tensor:
        """The model's raw readouts at EVERY time of ``grid_times`` (ascending, shared by all trajectories) for the
        observed trajectories in ``batch`` (a ``PackedBatch``, or lists of times / values): a (B, G, d_y, M) tensor.
        This is the simulation the reference's plotting code runs in Python, one ``euler_step`` at a time
        (utils/plotting.py:133-256), with its own step rule (``n_sub = max(1, int(gap / dt_ode_step))`` equal sub-steps
        to each grid time) and its conventions: a grid time that is an observation time shows the value after the
        jump, except at a trajectory's last observation; times before the first observation are 0.
        ``grid_moments`` turns the second readout into a variance the way the plots do."""
        params = self.flat_parameters()
        dev = self._check_runnable(params)
        if not isinstance(batch, PackedBatch):
            batch = PackedBatch.from_lists(batch, batch_values, device=dev)
        elif batch.device != dev:
            batch = batch.to(dev)
        grid = grid_times.to(dev, torch.float32).contiguous()
        lib = nat.load()
        desc = self.descriptor()
        G = int(grid.shape[0])
        with nat.on_device(dev):
            flat = self._flat_view(params)
            dense = torch.empty((batch.B, G, self.output_dim, self.num_moments), dtype=torch.float32, device=dev)
            ws_bytes = lib.njode_dense_workspace_bytes(desc)
            ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
            nat.check(lib.njode_dense_forward(desc, nat.ptr(flat), nat.ptr(batch.times), nat.ptr(batch.values),
                                              nat.ptr(batch.offsets), batch.B, batch.N, nat.ptr(grid), G, nat.ptr(dense),
                                              nat.ptr(ws), ws_bytes, nat.current_stream(dev)), "njode_dense_forward")
        return dense

    def grid_moments(self, dense: torch.Tensor):
        """(mean, variance) from ``predict_on_grid`` output as the reference's plots derive them (plotting.py:183-196):
        variance = W^2 (``variance_method='direct'``) or clamp(W - mean^2, 0) (``'second_moment'``); None for one moment."""
        mean = dense[..., 0]
        if self.num_moments < 2:
            return mean, None
        w = dense[..., 1]
        var = w * w if self.variance_method == "direct" else torch.clamp(w - mean * mean, min=0.0)
        return mean, var

    # -- large batches in waves ----------------------------------------------------------------------
    def _flat_grad(self, params):
        """The parameter gradients as ONE flat tensor when they tile one storage in parameter order (what the reverse
        sweep hands out and in-place accumulation preserves), else None."""
        first = params[0].grad
        if first is None:
            return None
        base, o = first.data_ptr(), 0
        for p in params:
            g = p.grad
            if g is None or not g.is_contiguous() or g.dtype != torch.float32 or g.data_ptr() != base + 4 * o:
                return None
            o += p.numel()
        if first.untyped_storage().nbytes() - first.storage_offset() * 4 < 4 * o:
            return None
        return torch.as_strided(first, (o,), (1,), first.storage_offset())

    def release_wave_buffers(self) -> None:
        """Free the checkpoint buffer that ``forward_backward_waves`` keeps between calls."""
        self._ckpt_pool = None

    def forward_backward_waves(self, batch: PackedBatch, wave: int, traj_scale: Optional[float] = None, **loss_kwargs):
        """Loss and parameter gradients of a batch that is too large for one sweep's checkpoints (BASELINE config 4:
        10 MB of checkpoints per trajectory at hidden 128 / 3 layers / ~1040 Euler steps): the batch is cut into
        waves of ``wave`` trajectories; every wave runs forward sweep, ``nj_ode_loss`` (scaled by ``traj_scale``,
        default 1 / batch.B -- pass 1 / B_global under data parallelism) and reverse sweep, and its checkpoints are
        released before the next one starts; gradients accumulate in ``.grad`` (zero them first, as with any
        backward).  With ``enable_data_parallel()`` the ranks' gradients are summed ONCE, after the last wave.
        All waves share one checkpoint buffer, which the model keeps for the next call (``release_wave_buffers()``
        frees it).  Returns the summed loss (0-dim tensor on the device; no host synchronisation).  The reference computes the
        same thing in one ``model(batch); nj_ode_loss(...); loss.backward()`` (utils/training.py:88-97): trajectories
        are independent and the loss is a mean over them (jump_ode.py:383), so waves only change the summation order."""
        if wave < 1:
            raise ValueError("forward_backward_waves: wave must be >= 1")
        B = batch.B
        scale = 1.0 / B if traj_scale is None else float(traj_scale)
        dp, self._dp_group = self._dp_group, None
        total = None
        self._ckpt_pool_on = True             # one checkpoint buffer for all waves (_alloc_ckpt); kept for the next call
        try:
            for lo in range(0, B, wave):
                sub = batch.slice(lo, min(lo + wave, B))
                preds, before = self.forward_packed(sub)
                loss = nj_ode_loss(sub, None, preds, before, traj_scale=scale, **loss_kwargs)
                loss.backward()
                total = loss.detach() if total is None else total + loss.detach()
                del preds, before, loss
        finally:
            self._dp_group = dp
            self._ckpt_pool_on = False
        if dp is not None:
            import torch.distributed as dist
            group = None if dp is True else dp
            params = self.flat_parameters()
            flat = self._flat_grad(params)
            if flat is not None:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            else:
                from ..sharding import allreduce_gradients
                allreduce_gradients(params, total, group=group)
        return total

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
        # (the entry holds a weak reference to the tensor: CPython may hand the id of a collected tensor to a new one)
        key = (id(moment_weights), moment_weights._version)
        hit = _MW_CACHE.get(key)
        if hit is not None and hit[0]() is moment_weights:
            w = hit[1]
        else:
            if len(_MW_CACHE) > 64:
                _MW_CACHE.clear()
            w = [float(v) for v in moment_weights.detach().cpu().tolist()]
            _MW_CACHE[key] = (weakref.ref(moment_weights), w)
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
    state = getattr(p, "_njode_state", None)
    if state is not None and (getattr(pb, "_njode_state", None) is not state or state.batch is not batch):
        state = None
    return _LossFunction.apply(ld, batch, scale, want_grad, state, p, pb)
