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
    float32 on the device), or by ``njode_forward_batch`` together with the first forward sweep.
    ``total_steps`` is the number of trajectory-ODE-steps of the batch.  The arrays live in one arena
    (layout: ``njode_batch_arena_bytes``); ``kenc`` / ``perm`` / ``tile_kmax`` / ``tile_slot_off`` /
    ``knots`` are tensor views of it, made on first use -- the kernels only need ``ptrs``.
    """

    _FIELDS = {"kenc": (nat.ARENA_KENC, torch.int32), "perm": (nat.ARENA_PERM, torch.int32),
               "tile_kmax": (nat.ARENA_TILE_KMAX, torch.int32), "tile_slot_off": (nat.ARENA_TILE_SLOT_OFF, torch.int64),
               "knots": (nat.ARENA_KNOTS, torch.float32)}

    def __init__(self, arena, layout, N, tile_rows, n_tiles, header, knots_buf=None):
        self._arena, self._layout, self._N = arena, list(layout), N
        self._knots_buf = knots_buf            # separate tensor when the arena was sized without knots
        self.tile_rows, self.n_tiles = tile_rows, n_tiles
        self.total_steps, self.total_slots = header[nat.HDR_TOTAL_STEPS], header[nat.HDR_TOTAL_SLOTS]
        self.kmax = header[nat.HDR_KMAX]
        base = arena.data_ptr()
        knots_ptr = knots_buf.data_ptr() if knots_buf is not None else base + layout[nat.ARENA_KNOTS]
        # (kenc, perm, tile_kmax, tile_slot_off, knots) as the C-ABI takes them
        self.ptrs = tuple(nat.C.c_void_p(base + layout[w]) for w in (nat.ARENA_KENC, nat.ARENA_PERM, nat.ARENA_TILE_KMAX,
                                                                    nat.ARENA_TILE_SLOT_OFF)) + (nat.C.c_void_p(knots_ptr),)

    def __getattr__(self, name):               # only reached for attributes not set yet: the lazy views
        f = Schedule._FIELDS.get(name)
        if f is None:
            raise AttributeError(name)
        word, dtype = f
        count = {"kenc": max(self._N, 1), "perm": max(self.n_tiles * self.tile_rows, 1), "tile_kmax": max(self.n_tiles, 1),
                 "tile_slot_off": self.n_tiles + 1, "knots": max(self.total_slots * self.tile_rows, 1)}[name]
        item = 8 if dtype == torch.int64 else 4
        if name == "knots" and self._knots_buf is not None:
            t = self._knots_buf
        else:
            o = self._layout[word]
            t = self._arena[o:o + count * item].view(dtype)
        setattr(self, name, t)
        return t


_HDR_HOST = {}


def pinned_header(dev):
    """A pinned int64[HDR_WORDS] host buffer per device for the schedule header read-back (read synchronously
    right after the call that fills it, so one buffer is enough)."""
    h = _HDR_HOST.get(dev)
    if h is None:
        h = _HDR_HOST[dev] = torch.empty(nat.HDR_WORDS, dtype=torch.int64).pin_memory()
    return h


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

    # -- sub-batches (device-side, no host copies of the data) ----------------------------------
    def _host_offsets(self) -> List[int]:
        if getattr(self, "_off_host", None) is None:
            self._off_host = list(itertools.accumulate(self.sizes, initial=0))
        return self._off_host

    def slice(self, lo: int, hi: int) -> "PackedBatch":
        """Trajectories [lo, hi) as a PackedBatch of views (cached: a wave of a large batch keeps its schedule)."""
        if not (0 <= lo < hi <= self.B):
            raise ValueError(f"PackedBatch.slice: need 0 <= lo < hi <= {self.B}")
        cache = self.__dict__.setdefault("_slices", {})
        sub = cache.get((lo, hi))
        if sub is None:
            off = self._host_offsets()
            a, b = off[lo], off[hi]
            sub = PackedBatch(self.times[a:b], self.values[a:b], self.offsets[lo:hi + 1] - a, self.sizes[lo:hi])
            if len(cache) < 4096:
                cache[(lo, hi)] = sub
        return sub

    def gather(self, index: torch.Tensor, index_host: Optional[Sequence[int]] = None) -> "PackedBatch":
        """The trajectories ``index`` (1-D int64 tensor on this batch's device, any order) as a new PackedBatch, built
        on the device with no host synchronisation: the host only needs the trajectory lengths it already knows
        (``index_host`` = the same indices as a Python sequence; taken from ``index`` with one D2H copy if omitted)."""
        if index_host is None:
            index_host = index.tolist()
        sizes_all = self.sizes
        sizes = [sizes_all[i] for i in index_host]
        n = sum(sizes)
        dev = self.device
        lens = torch.tensor(sizes, dtype=torch.int64).to(dev, non_blocking=True)
        new_off = torch.zeros(len(sizes) + 1, dtype=torch.int64, device=dev)
        torch.cumsum(lens, 0, out=new_off[1:])
        src0 = self.offsets.index_select(0, index)                       # first observation of every picked trajectory
        obs = torch.repeat_interleave(src0 - new_off[:-1], lens, output_size=n) + torch.arange(n, device=dev)
        return PackedBatch(self.times.index_select(0, obs), self.values.index_select(0, obs), new_off, sizes)

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
        n_tiles = lib.njode_num_tiles(desc, self.N)   # tiles may be partially filled (small batches, tcgen05 flavour)
        if n_tiles < 0:
            nat.check(-1, "njode_num_tiles")
        key = self.schedule_key(desc, tile_rows, n_tiles)
        s = self._schedules.get(key)
        if s is not None:
            return s
        if not self.times.is_cuda:
            raise RuntimeError("the Neural Jump ODE hot path runs on CUDA only (no CPU fallback); "
                               "move the model and the batch to a CUDA device")
        dev = self.device
        N, B = self.N, self.B
        with nat.on_device(dev):
            stream = nat.current_stream(dev)
            layout = (nat.C.c_int64 * nat.ARENA_WORDS)()
            arena_bytes = lib.njode_batch_arena_bytes(desc, B, N, 0, layout)      # everything but the knots
            arena = torch.empty(arena_bytes, dtype=torch.uint8, device=dev)
            base = arena.data_ptr()
            at = lambda w: nat.C.c_void_p(base + layout[w])
            ws_bytes = lib.njode_schedule_workspace_bytes(B, N, tile_rows)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            nat.check(lib.njode_schedule_build(desc, nat.ptr(self.times), nat.ptr(self.offsets), B, N, tile_rows,
                                               at(nat.ARENA_KENC), at(nat.ARENA_PERM), at(nat.ARENA_TILE_KMAX),
                                               at(nat.ARENA_TILE_SLOT_OFF), at(nat.ARENA_HEADER), nat.ptr(ws), ws_bytes,
                                               stream), "njode_schedule_build")
            host = pinned_header(dev)
            host.copy_(arena[layout[nat.ARENA_HEADER]:layout[nat.ARENA_HEADER] + 8 * nat.HDR_WORDS].view(torch.int64),
                       non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()   # the one host sync of the schedule: sizes the checkpoints
            hdr = host.tolist()
            knots = torch.empty(max(hdr[nat.HDR_TOTAL_SLOTS] * tile_rows, 1), dtype=torch.float32, device=dev)
            s = Schedule(arena, layout, N, tile_rows, n_tiles, hdr, knots_buf=knots)
            nat.check(lib.njode_schedule_knots(nat.ptr(self.times), *s.ptrs[:4], N, n_tiles, tile_rows,
                                               desc, s.ptrs[4], stream), "njode_schedule_knots")
        self._schedules[key] = s
        return s

    @staticmethod
    def schedule_key(desc, tile_rows, n_tiles):
        # (the flavour is part of the key: tiled and wide share tile_rows but not the slots a tile owns)
        return (bool(desc.has_dt), float(desc.dt), int(tile_rows), int(n_tiles), int(nat.load().njode_selected_impl(desc)))

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
