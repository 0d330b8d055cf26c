"""Host-side cost (us) of each C-ABI call of the step, launches only (one sync per iteration, outside the clocks)."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "neural-jump-ode_b200"))
import torch
from bench import WORKLOADS, make_batch
from neural_jump_ode import NeuralJumpODE, nj_ode_loss, PackedBatch, _native as nat
wl = WORKLOADS["ou_shared_b4096"]; dev = torch.device("cuda")
torch.manual_seed(0); model = NeuralJumpODE(**wl["model"]).to(dev)
batch = make_batch(wl, wl["B"], dev, 1000)
lib = nat.load(); desc = model.descriptor()
B, N = batch.B, batch.N
tile_rows = lib.njode_tile_rows(desc); n_tiles = lib.njode_num_tiles(desc, N)
s = batch.schedule(desc)
layout = (C.c_int64 * 8)()
arena = torch.empty(lib.njode_batch_arena_bytes(desc, B, N, s.total_slots, layout), dtype=torch.uint8, device=dev)
at = lambda w: C.c_void_p(arena.data_ptr() + layout[w])
ws_b = lib.njode_schedule_workspace_bytes(B, N, tile_rows); ws = torch.empty(ws_b, dtype=torch.uint8, device=dev)
flat = model._flat_view(model.flat_parameters())
out = torch.empty((2, N, 1, 2), device=dev)
ck = torch.empty(s.total_slots * tile_rows * lib.njode_ckpt_row_floats(desc), device=dev)
fws_b = lib.njode_forward_workspace_bytes(desc); fws = torch.empty(max(fws_b, 1), dtype=torch.uint8, device=dev)
sc_b = lib.njode_batch_scratch_bytes(desc, B, N); sc = torch.empty(sc_b, dtype=torch.uint8, device=dev)
host = torch.empty(8, dtype=torch.int64).pin_memory()
st = nat.current_stream(dev)
acc = {}
def clock(name, fn):
    t0 = time.perf_counter(); rc = fn(); t1 = time.perf_counter()
    assert rc == 0 or name.endswith("bytes"), (name, rc, lib.njode_last_error())
    acc[name] = acc.get(name, 0.0) + (t1 - t0)
n = 200
for it in range(n + 10):
    if it == 10: acc.clear()
    torch.cuda.synchronize()
    clock("sched_ws_bytes", lambda: lib.njode_schedule_workspace_bytes(B, N, tile_rows))
    clock("schedule_build", lambda: lib.njode_schedule_build(desc, nat.ptr(batch.times), nat.ptr(batch.offsets), B, N, tile_rows, at(0), at(1), at(2), at(3), at(4), nat.ptr(ws), ws_b, st))
    t0 = time.perf_counter(); torch.cuda.synchronize(); acc["schedule_build drain (device time left after the launch returns)"] = acc.get("schedule_build drain (device time left after the launch returns)", 0.0) + time.perf_counter() - t0
    clock("schedule_knots", lambda: lib.njode_schedule_knots(nat.ptr(batch.times), at(0), at(1), at(2), at(3), N, n_tiles, tile_rows, desc, at(5), st))
    clock("forward", lambda: lib.njode_forward(desc, nat.ptr(flat), nat.ptr(batch.times), nat.ptr(batch.values), nat.ptr(batch.offsets), B, N, at(0), at(1), at(2), at(3), at(5), n_tiles, s.total_slots, tile_rows, nat.ptr(out[0]), nat.ptr(out[1]), nat.ptr(ck), nat.ptr(fws), fws_b, st))
    torch.cuda.synchronize()
    clock("forward_batch(incl. its sync)", lambda: lib.njode_forward_batch(desc, nat.ptr(flat), nat.ptr(batch.times), nat.ptr(batch.values), nat.ptr(batch.offsets), B, N, nat.ptr(arena), arena.numel(), 1, nat.ptr(ck), ck.numel(), nat.ptr(sc), sc_b, host.data_ptr(), nat.ptr(out[0]), nat.ptr(out[1]), st))
    t0 = time.perf_counter(); torch.cuda.synchronize(); acc["forward kernel drain"] = acc.get("forward kernel drain", 0.0) + time.perf_counter() - t0
print({k: round(v / n * 1e6, 1) for k, v in acc.items()})
