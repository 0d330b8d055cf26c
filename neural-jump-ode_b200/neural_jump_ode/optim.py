"""Flat-buffer Adam for the hot path's parameters.

The reference trains with ``optim.Adam(model.parameters(), lr, weight_decay)`` (utils/training.py:396): a
multi-tensor update over ~12-24 small tensors per step.  ``FlatAdam`` takes the same constructor arguments and
has the same semantics (weight decay as L2 added to the gradient, bias correction, eps outside the square root),
but keeps parameters, gradients and both moments in ONE flat device buffer each, so ``step()`` is a single launch
of ``njode_adam_step`` (include/njode.h) and the reverse sweep's flat gradient is consumed without a gather.

    opt = FlatAdam(model.parameters(), lr=1e-3, weight_decay=5e-4)     # after model.to("cuda")
    loss.backward(); opt.step(); opt.zero_grad()

On construction the parameters are re-pointed at slices of the flat buffer (``p.data`` becomes a view; values,
``state_dict`` keys and shapes are unchanged).  CUDA float32 parameters only -- like the rest of the package there
is no CPU fallback.

Checkpoints.  ``state_dict()`` / ``load_state_dict()`` speak ``torch.optim.Adam``'s layout (per parameter ``step``,
``exp_avg``, ``exp_avg_sq``; the per-parameter tensors are views of the flat moment buffers), so the reference
Trainer's ``optimizer_state_dict`` (utils/training.py:152-154, :291-304) round-trips and a checkpoint written by
``torch.optim.Adam`` resumes under ``FlatAdam`` and vice versa.
"""
from __future__ import annotations

from typing import Iterable

import torch

from . import _native as nat


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0):
        if lr < 0.0 or eps < 0.0 or weight_decay < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0):
            raise ValueError("FlatAdam: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._flat = []
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.requires_grad]
            if not ps:
                self._flat.append(None)
                continue
            dev = ps[0].device
            if dev.type != "cuda" or any(p.device != dev or p.dtype != torch.float32 for p in ps):
                raise RuntimeError("FlatAdam needs float32 parameters on one CUDA device (no CPU fallback); "
                                   "build the optimizer after model.to('cuda')")
            n = sum(p.numel() for p in ps)
            flat = self._adopt(ps, n)
            if flat is not None:
                # already tiling one storage (model.flatten_parameters()): step on that buffer, in its order -- which
                # is also the order of the reverse sweep's flat gradient, so nothing is ever gathered
                ps = sorted(ps, key=lambda p: p.data_ptr())
            else:
                flat = torch.empty(n, dtype=torch.float32, device=dev)
                o = 0
                with torch.no_grad():
                    for p in ps:
                        k = p.numel()
                        flat[o:o + k].copy_(p.detach().reshape(-1))
                        p.data = flat[o:o + k].view(p.shape)          # the parameter now lives in the flat buffer
                        o += k
            self._flat.append(dict(params=ps, flat=flat, grad=torch.zeros_like(flat), exp_avg=torch.zeros_like(flat),
                                   exp_avg_sq=torch.zeros_like(flat), step=0, step_t=torch.zeros((), dtype=torch.float32)))

    # -- torch.optim.Adam-compatible optimizer state (views of the flat moment buffers) ---------------------
    def _publish_state(self, st):
        """self.state[p] = {step, exp_avg, exp_avg_sq} as torch.optim.Adam keeps them; the moments are views of the
        flat buffers, ``step`` is one tensor shared by the group's parameters."""
        o = 0
        for p in st["params"]:
            k = p.numel()
            self.state[p] = dict(step=st["step_t"], exp_avg=st["exp_avg"][o:o + k].view(p.shape),
                                 exp_avg_sq=st["exp_avg_sq"][o:o + k].view(p.shape))
            o += k

    def state_dict(self):
        """``torch.optim.Adam``'s layout.  Every parameter gets its OWN ``step`` tensor in the returned dict: inside
        FlatAdam the group's parameters share one, but torch.optim.Adam increments the step tensors of all parameters
        with one ``_foreach_add_`` -- a shared tensor would advance once per parameter."""
        sd = super().state_dict()
        for ps in sd["state"].values():
            if torch.is_tensor(ps.get("step")):
                ps["step"] = ps["step"].clone()
        return sd

    def load_state_dict(self, state_dict):
        """Accepts what ``FlatAdam.state_dict()`` or ``torch.optim.Adam.state_dict()`` produced for the same
        parameters (in the same order); moments and step counts are copied into the flat buffers."""
        super().load_state_dict(state_dict)          # validates groups / sizes, casts tensors to the parameters' device
        for st in self._flat:
            if st is None:
                continue
            o, steps = 0, set()
            for p in st["params"]:
                k = p.numel()
                ps = self.state.get(p)
                if ps:
                    st["exp_avg"][o:o + k].copy_(ps["exp_avg"].reshape(-1))
                    st["exp_avg_sq"][o:o + k].copy_(ps["exp_avg_sq"].reshape(-1))
                    steps.add(int(float(ps["step"])))
                else:                                # a parameter that had not been stepped yet
                    st["exp_avg"][o:o + k].zero_()
                    st["exp_avg_sq"][o:o + k].zero_()
                    steps.add(0)
                o += k
            if len(steps) > 1:
                raise ValueError("FlatAdam.load_state_dict: the parameters of one group carry different step counts "
                                 f"({sorted(steps)}); a flat group has one")
            st["step"] = steps.pop() if steps else 0
            st["step_t"] = torch.tensor(float(st["step"]), dtype=torch.float32)
            for p in st["params"]:
                self.state.pop(p, None)
            if st["step"] > 0:
                self._publish_state(st)

    @staticmethod
    def _adopt(ps, n):
        """The flat tensor the parameters already tile (in some order, back to back in one storage), or None."""
        if any(not p.is_contiguous() for p in ps):
            return None
        order = sorted(ps, key=lambda p: p.data_ptr())
        first = order[0]
        addr = first.data_ptr()
        for p in order:
            if p.data_ptr() != addr:
                return None
            addr += 4 * p.numel()
        if first.untyped_storage().nbytes() < (first.storage_offset() + n) * 4:
            return None
        return first.detach().as_strided((n,), (1,), first.storage_offset())

    def _gather_grads(self, st):
        """The flat gradient.  The reverse sweep hands out views of one flat buffer in parameter order: if the
        ``.grad`` tensors still are such views nothing is copied; otherwise one multi-tensor copy packs them."""
        ps = st["params"]
        first = ps[0].grad
        if first is not None:
            base, o, ok = first.data_ptr(), 0, True
            for p in ps:
                g = p.grad
                if g is None or not g.is_contiguous() or g.dtype != torch.float32 or g.data_ptr() != base + 4 * o:
                    ok = False
                    break
                o += p.numel()
            if ok and first.untyped_storage().nbytes() - first.storage_offset() * 4 >= 4 * o:
                return torch.as_strided(first, (o,), (1,), first.storage_offset())
        flat, o, dst, src = st["grad"], 0, [], []
        for p in ps:
            k = p.numel()
            if p.grad is None:
                flat[o:o + k].zero_()
            else:
                dst.append(flat[o:o + k].view(p.shape))
                src.append(p.grad)
            o += k
        if dst:
            torch._foreach_copy_(dst, src)
        return flat

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = nat.load()
        for group, st in zip(self.param_groups, self._flat):
            if st is None:
                continue
            g = self._gather_grads(st)
            st["step"] += 1
            st["step_t"] += 1
            if st["step"] == 1 or st["params"][0] not in self.state:
                self._publish_state(st)
            flat = st["flat"]
            with torch.cuda.device(flat.device):
                stream = torch.cuda.current_stream(flat.device).cuda_stream
                nat.check(lib.njode_adam_step(nat.ptr(flat), nat.ptr(g), nat.ptr(st["exp_avg"]), nat.ptr(st["exp_avg_sq"]),
                                              flat.numel(), float(group["lr"]), float(group["betas"][0]),
                                              float(group["betas"][1]), float(group["eps"]), float(group["weight_decay"]),
                                              st["step"], 1.0, stream), "njode_adam_step")
            # the kernel wrote the parameters through a raw pointer: autograd's version counters did not move, so tell
            # the sweeps' in-place-modification guard (jump_ode._SweepState) that this buffer changed
            nat.bump_generation(flat.data_ptr())
        return loss
