"""Packed batch format and step schedule -- the host side of the hot path's data layout.

The reference API passes Python lists of per-trajectory tensors (jump_ode.py:218-225).  The kernels
work on one packed batch: ``times (N,)``, ``values (N,d_x)``, ``offsets (B+1,) int64``.  A
``PackedBatch`` can be built once (e.g. with ``--cache-data``) and passed wherever the lists go:
``model(batch, None)`` and ``nj_ode_loss(batch, None, preds, preds_before)``.
"""
from __future__ import annotations

import itertools
from typing import List, Optional, Sequence

import torch

from . import _native as nat


class Schedule:
    """Device-resident Euler step schedule of one packed batch for one (dt, tile_rows).

    Built by ``njode_schedule_build`` / ``njode_schedule_knots`` (jump_ode.py:188-203 reproduced in
    float32 on the device).  ``total_steps`` is the number of trajectory-ODE-steps of the batch.
    """

    __slots__ = ("kenc", "perm", "tile_kmax", "tile_slot_off", "knots", "tile_rows", "n_tiles",
                 "total_steps", "total_slots", "kmax")


class PackedBatch:
    """A batch of irregularly observed trajectories, packed for the device."""

    def __init__(self, times: torch.Tensor, values: torch.Tensor, offsets: torch.Tensor,
                 sizes: Optional[Sequence[int]] = None):
        if times.dim() != 1 or values.dim() != 2 or values.shape[0] != times.shape[0]:
            raise ValueError("PackedBatch: times must be (N,), values (N, d_x)")
        if offsets.dim() != 1 or offsets.dtype != torch.int64:
            raise ValueError("PackedBatch: offsets must be a 1-D int64 tensor of length B+1")
        self.times = times.contiguous().float()
        self.values = values.contiguous().float()
        self.offsets = offsets.contiguous()
        self.B = int(offsets.shape[0]) - 1
        self.N = int(times.shape[0])
        self._sizes = None if sizes is None else list(sizes)
        self._schedules = {}
        self._src = None          # (times_list, values_list) this batch was packed from, if any

    # -- construction -----------------------------------------------------------------------
    @classmethod
    def from_lists(cls, batch_times: List[torch.Tensor], batch_values: List[torch.Tensor],
                   device=None) -> "PackedBatch":
        if len(batch_times) != len(batch_values):
            raise ValueError("batch_times and batch_values must have the same length")
        if len(batch_times) == 0:
            raise ValueError("empty batch")
        sizes = [int(t.shape[0]) for t in batch_times]
        for t, v in zip(batch_times, batch_values):
            if v.dim() != 2 or t.dim() != 1 or v.shape[0] != t.shape[0]:
                raise ValueError("each trajectory needs times (n_i,) and values (n_i, d_x)")
        dev = torch.device(device) if device is not None else batch_values[0].device
        times = torch.cat([t.reshape(-1) for t in batch_times]).to(dev, torch.float32, non_blocking=True)
        values = torch.cat(list(batch_values)).to(dev, torch.float32, non_blocking=True)
        off = torch.tensor(list(itertools.accumulate(sizes, initial=0)), dtype=torch.int64)
        batch = cls(times, values, off.to(dev, non_blocking=True), sizes)
        batch._src = (batch_times, batch_values)
        return batch

    @property
    def device(self):
        return self.values.device

    @property
    def sizes(self) -> List[int]:
        if self._sizes is None:
            off = self.offsets.cpu()
            self._sizes = (off[1:] - off[:-1]).tolist()
        return self._sizes

    def to(self, device) -> "PackedBatch":
        b = PackedBatch(self.times.to(device), self.values.to(device), self.offsets.to(device), self._sizes)
        b._src = self._src
        return b

    def came_from(self, batch_times, batch_values) -> bool:
        return self._src is not None and self._src[0] is batch_times and self._src[1] is batch_values

    def split(self, packed: torch.Tensor):
        """Per-trajectory views of a packed (N, ...) tensor."""
        return list(torch.split(packed, self.sizes, dim=0))

    # -- schedule ---------------------------------------------------------------------------
    def schedule(self, desc: "nat.NjodeDesc") -> Schedule:
        lib = nat.load()
        tile_rows = lib.njode_tile_rows(desc)
        if tile_rows < 1:
            nat.check(-1, "njode_tile_rows")
        key = (bool(desc.has_dt), float(desc.dt), int(tile_rows))
        s = self._schedules.get(key)
        if s is not None:
            return s
        if not self.times.is_cuda:
            raise RuntimeError("the Neural Jump ODE hot path runs on CUDA only (no CPU fallback); "
                               "move the model and the batch to a CUDA device")
        dev = self.device
        N, B = self.N, self.B
        n_tiles = lib.njode_num_tiles(desc, N)        # tiles may be partially filled (small batches, tcgen05 flavour)
        if n_tiles < 0:
            nat.check(-1, "njode_num_tiles")
        s = Schedule()
        s.tile_rows, s.n_tiles = tile_rows, n_tiles
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            s.kenc = torch.empty(max(N, 1), dtype=torch.int32, device=dev)
            s.perm = torch.empty(max(n_tiles * tile_rows, 1), dtype=torch.int32, device=dev)
            s.tile_kmax = torch.empty(max(n_tiles, 1), dtype=torch.int32, device=dev)
            s.tile_slot_off = torch.empty(n_tiles + 1, dtype=torch.int64, device=dev)
            header = torch.empty(nat.HDR_WORDS, dtype=torch.int64, device=dev)
            ws_bytes = lib.njode_schedule_workspace_bytes(B, N, tile_rows)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            nat.check(lib.njode_schedule_build(desc, nat.ptr(self.times), nat.ptr(self.offsets), B, N, tile_rows,
                                               nat.ptr(s.kenc), nat.ptr(s.perm), nat.ptr(s.tile_kmax),
                                               nat.ptr(s.tile_slot_off), nat.ptr(header), nat.ptr(ws), ws_bytes,
                                               stream), "njode_schedule_build")
            hdr = header.cpu().tolist()     # the one host sync of the schedule: sizes the checkpoints
            s.total_steps, s.total_slots, s.kmax = hdr[nat.HDR_TOTAL_STEPS], hdr[nat.HDR_TOTAL_SLOTS], hdr[nat.HDR_KMAX]
            s.knots = torch.empty(max(s.total_slots * tile_rows, 1), dtype=torch.float32, device=dev)
            nat.check(lib.njode_schedule_knots(nat.ptr(self.times), nat.ptr(s.kenc), nat.ptr(s.perm),
                                               nat.ptr(s.tile_kmax), nat.ptr(s.tile_slot_off), N, n_tiles, tile_rows,
                                               desc, nat.ptr(s.knots), stream), "njode_schedule_knots")
        self._schedules[key] = s
        return s

    def step_counts(self, desc) -> torch.Tensor:
        """Per-observation Euler step counts (int32, device); the last observation of a trajectory has 0."""
        return self.schedule(desc).kenc[: self.N] >> 1


class PredList(list):
    """The list of per-trajectory prediction tensors the reference API returns
    (jump_ode.py:227-233), remembering the packed tensor the entries are views of."""

    def __init__(self, packed: torch.Tensor, batch: PackedBatch):
        super().__init__(batch.split(packed))
        self.packed = packed
        self.batch = batch
        self._ids = [id(t) for t in self]

    def untouched(self) -> bool:
        return len(self) == len(self._ids) and all(id(t) == i for t, i in zip(self, self._ids))
