"""CPU-only checks of the host side: import surface, state_dict layout, init RNG order, the C-ABI
library loads and exports every symbol include/njode.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, load_golden, golden_names


def test_import_surface():
    import neural_jump_ode
    from neural_jump_ode import NeuralJumpODE, nj_ode_loss          # reference __init__.py:3
    from neural_jump_ode.models import JumpNN, ODEFunc, OutputNN    # reference models/__init__.py:3
    assert neural_jump_ode.__version__ == "0.1.0"
    assert callable(nj_ode_loss) and callable(NeuralJumpODE)
    assert all(callable(c) for c in (JumpNN, ODEFunc, OutputNN))


@pytest.mark.parametrize("name", golden_names())
def test_state_dict_and_init_match_reference(name):
    """Same keys, shapes AND values as the reference under the same torch.manual_seed: parameters are
    created in the reference's order, so old checkpoints load and seeds reproduce."""
    from neural_jump_ode import NeuralJumpODE
    g = load_golden(name)
    torch.manual_seed(g["seed"])
    m = NeuralJumpODE(**g["model"])
    sd = m.state_dict()
    assert list(sd.keys()) == list(g["params"].keys()) or set(sd) == set(g["params"])
    for k, v in sd.items():
        assert torch.equal(v, g["params"][k]), k
    m.load_state_dict(g["params"])
    flat = m.flat_parameters()
    assert sum(p.numel() for p in flat) == sum(p.numel() for p in m.parameters())
    assert len({id(p) for p in flat}) == len(list(m.parameters()))


def test_constructor_compat():
    from neural_jump_ode import NeuralJumpODE
    m = NeuralJumpODE(1, 16, 1)                                      # positional, as in the reference tests
    assert m.jump_nns is not None and m.jump_nn is None and m.num_moments == 1
    m = NeuralJumpODE(input_dim=1, hidden_dim=16, output_dim=1, n_steps_between=3)   # stale README kwarg
    assert m.dt_ode_step is None
    m = NeuralJumpODE(1, 16, 1, num_moments=2, shared_network=True, activation="identity")
    assert m.jump_nn is not None and m.jump_nns is None
    assert isinstance(m.jump_nn.net[1], torch.nn.ReLU)              # unknown activation -> ReLU
    assert m.output_nn.net[3].weight.shape == (2, 16)
    for attr in ("num_moments", "shared_network", "variance_method", "dt_ode_step", "output_dim"):
        assert hasattr(m, attr)
    with pytest.raises(ValueError):
        NeuralJumpODE(1, 16, 1, input_scaling="bogus")


def test_submodules_callable_on_cpu():
    """The small sub-modules are ordinary torch modules (plotting drives them); only the batched
    hot path is CUDA-only."""
    from neural_jump_ode import NeuralJumpODE
    m = NeuralJumpODE(1, 8, 1, num_moments=2)
    x = torch.tensor([[0.5]])
    h = [m.jump_nns[i](x) for i in range(2)]
    h2 = m.euler_step(h, x, torch.tensor(0.0), torch.tensor(0.1))
    assert h2[0].shape == (1, 8) and m.output_nns[1](h2[1]).shape == (1, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m([torch.tensor([0.0, 1.0])], [torch.tensor([[1.0], [2.0]])])


def test_abi_library_exports_every_declared_symbol():
    from neural_jump_ode import _native as nat
    hdr = open(os.path.join(ROOT, "include", "njode.h")).read()
    declared = set(re.findall(r"\b(njode_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(nat.SIGNATURES), declared ^ set(nat.SIGNATURES)
    lib = ctypes.CDLL(nat.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib2 = nat.load()
    assert lib2.njode_abi_version() == nat.ABI_VERSION


def test_abi_param_count_and_validation():
    from neural_jump_ode import NeuralJumpODE, _native as nat
    lib = nat.load()
    for kw in (dict(num_moments=2), dict(num_moments=2, shared_network=True), dict(n_hidden_layers=3, hidden_dim=128)):
        args = dict(input_dim=1, hidden_dim=32, output_dim=1)
        args.update(kw)
        m = NeuralJumpODE(**args)
        d = m.descriptor()
        assert lib.njode_param_count(d) == sum(p.numel() for p in m.parameters())
        assert lib.njode_num_stacks(d) == (1 if m.shared_network else m.num_moments)
    bad = NeuralJumpODE(1, 32, 1).descriptor()
    bad.n_hidden_layers = 0
    assert lib.njode_param_count(bad) == -1
    assert b"n_hidden_layers" in lib.njode_last_error()


def test_abi_tiling_and_arena_layout():
    """Host-only arithmetic of the ABI: tiles per batch (njode_num_tiles: quarter / half / full tiles for the tcgen05
    flavour, 32-row tiles otherwise) and the arena layout of njode_forward_batch (256-byte aligned, knots last, only
    the knots depend on the slot count).  No kernel is launched."""
    import ctypes as C
    from neural_jump_ode import NeuralJumpODE, _native as nat
    lib = nat.load()
    tiled = NeuralJumpODE(1, 32, 1, dt_ode_step=0.01, num_moments=2, shared_network=True).descriptor()
    rowt = NeuralJumpODE(1, 96, 1, dt_ode_step=0.01, num_moments=2).descriptor()
    wide = NeuralJumpODE(1, 128, 1, dt_ode_step=0.01, num_moments=2, n_hidden_layers=3).descriptor()
    assert lib.njode_tile_rows(tiled) == 128 and lib.njode_tile_rows(rowt) == 32 and lib.njode_tile_rows(wide) == 128
    assert [lib.njode_selected_impl(d) for d in (tiled, rowt, wide)] == [nat.IMPL["tiled"], nat.IMPL["rowtile"], nat.IMPL["wide"]]
    # wide: (L+1) activation planes + (L+1) data-gradient planes + 8 aux floats per row and slot
    assert lib.njode_ckpt_row_floats(tiled) == 64 and lib.njode_ckpt_row_floats(rowt) == 192
    assert lib.njode_ckpt_row_floats(wide) == 2 * 4 * 128 + 8
    assert lib.njode_forward_workspace_bytes(wide) >= 2 * 10 * 2 * 128 * 128 * 4     # S x (3L+1) split weight images
    sms = 148                                            # what the library assumes when it cannot ask a device
    for N in (0, 1, 31, 32, 33, 1000, 4 * 128 * sms // 4, 40960, 10 ** 6):
        full = (N + 127) // 128
        units = 32 if 4 * full <= sms else 64 if full <= sms else 128
        assert lib.njode_num_tiles(tiled, N) == (N + units - 1) // units, N
        assert lib.njode_num_tiles(rowt, N) == (N + 31) // 32, N
        assert lib.njode_num_tiles(wide, N) == full, N
    assert lib.njode_num_tiles(tiled, -1) == -1
    for desc in (tiled, rowt, wide):
        rows = lib.njode_tile_rows(desc)
        for N, B in ((1, 1), (40960, 4096), (10 ** 6, 10 ** 5)):
            lay0, lay1 = (C.c_int64 * nat.ARENA_WORDS)(), (C.c_int64 * nat.ARENA_WORDS)()
            fixed = lib.njode_batch_arena_bytes(desc, B, N, 0, lay0)
            slots = 12345
            total = lib.njode_batch_arena_bytes(desc, B, N, slots, lay1)
            assert list(lay0) == list(lay1)              # nothing but the size depends on the slot count
            offs = [lay1[w] for w in (nat.ARENA_KENC, nat.ARENA_PERM, nat.ARENA_TILE_KMAX, nat.ARENA_TILE_SLOT_OFF,
                                      nat.ARENA_HEADER, nat.ARENA_KNOTS)]
            assert offs[0] == 0 and offs == sorted(offs) and all(o % 256 == 0 for o in offs)
            n_tiles = lib.njode_num_tiles(desc, N)
            assert offs[1] - offs[0] >= 4 * N and offs[2] - offs[1] >= 4 * n_tiles * rows
            assert offs[4] - offs[3] >= 8 * (n_tiles + 1) and offs[5] - offs[4] >= 8 * nat.HDR_WORDS
            assert fixed <= total and total - offs[5] >= 4 * slots * rows and total % 256 == 0
            assert lib.njode_batch_scratch_bytes(desc, B, N) >= lib.njode_schedule_workspace_bytes(B, N, rows)


def test_slot_guess_bookkeeping():
    """The checkpoint-slot guess for a batch whose schedule is not known yet: exact for a batch shape seen before,
    else the largest slots-per-tile ratio seen (+3 %), and 0 (= size exactly after the schedule is built) at first."""
    from neural_jump_ode import NeuralJumpODE
    m = NeuralJumpODE(1, 32, 1, dt_ode_step=0.01)
    assert m._guess_slots(1000, 100, 8) == 0
    m._note_slots(1000, 100, 8, 90)
    assert m._guess_slots(1000, 100, 8) == 90                       # same shape: exact
    assert m._guess_slots(2000, 200, 16) == int(90 / 8 * 16 * 1.03) + 2
    m._note_slots(500, 50, 4, 80)                                   # a batch with longer gaps raises the ratio
    assert m._guess_slots(2000, 200, 16) == int(20.0 * 16 * 1.03) + 2
    for i in range(300):                                            # the exact-shape memo is bounded
        m._note_slots(10000 + i, 7, 3, 5)
    assert len(m._slots_memo) <= 257


def test_packed_batch_host_logic():
    from neural_jump_ode import PackedBatch
    bt = [torch.tensor([0.0, 0.5, 1.0]), torch.tensor([0.0, 0.3])]
    bv = [torch.tensor([[1.0], [1.2], [0.9]]), torch.tensor([[0.5], [0.7]])]
    b = PackedBatch.from_lists(bt, bv)
    assert b.B == 2 and b.N == 5 and b.offsets.tolist() == [0, 3, 5] and b.sizes == [3, 2]
    parts = b.split(torch.arange(5.0).view(5, 1))
    assert [p.shape[0] for p in parts] == [3, 2]
    assert b.came_from(bt, bv) and not b.came_from(list(bt), bv)
    with pytest.raises(ValueError):
        PackedBatch.from_lists(bt, bv[:1])
    with pytest.raises(RuntimeError, match="CUDA only"):
        from neural_jump_ode import NeuralJumpODE
        b.schedule(NeuralJumpODE(1, 8, 1, dt_ode_step=0.1).descriptor())


@pytest.mark.parametrize("process", ["black_scholes", "ornstein_uhlenbeck", "heston", "hybrid_ou_bs"])
def test_path_generators_and_observation_rule(process):
    """Vectorised generators (row N2): shapes, the reference's observation rule (first and last grid point always
    observed, n_obs = max(2, int(frac * n_grid)), data_generation.py:235-249), strictly increasing float32 grid
    times, finite values, determinism in the seed."""
    from neural_jump_ode.simulation import simulate_paths, sample_observations, make_packed_batch
    g = torch.Generator().manual_seed(3)
    times, X = simulate_paths(process, 257, n_steps=100, T=1.0, device="cpu", generator=g)
    assert times.shape == (101,) and X.shape == (257, 101) and X.dtype == torch.float32
    assert torch.isfinite(X).all() and torch.equal(times, torch.linspace(0.0, 1.0, 101))
    batch = sample_observations(times, X, 0.1, generator=g)
    assert batch.B == 257 and batch.sizes == [10] * 257 and batch.N == 2570
    t = batch.times.view(257, 10)
    assert float(t[:, 0].max()) == 0.0 and float(t[:, -1].min()) == 1.0
    assert bool((t[:, 1:] > t[:, :-1]).all())
    a = make_packed_batch(process, 33, 0.1, device="cpu", seed=5)
    b = make_packed_batch(process, 33, 0.1, device="cpu", seed=5)
    assert torch.equal(a.values, b.values) and torch.equal(a.times, b.times)
    with pytest.raises(ValueError):
        simulate_paths("no_such_process", 4, device="cpu")


def test_mixed_ragged_batch():
    from neural_jump_ode.simulation import make_mixed_ragged_batch
    b = make_mixed_ragged_batch(101, 0.02, 0.2, n_steps=100, device="cpu", seed=2)
    assert b.B == 101 and b.N == sum(b.sizes) and int(b.offsets[-1]) == b.N
    assert min(b.sizes) >= 2 and max(b.sizes) <= 20 and len(set(b.sizes)) > 3          # ragged
    off = b.offsets.tolist()
    for i in range(b.B):
        t = b.times[off[i]:off[i + 1]]
        assert float(t[0]) == 0.0 and float(t[-1]) == 1.0 and bool((t[1:] > t[:-1]).all())
    assert torch.isfinite(b.values).all()


def _check_against_reference_moments(device, n=20000):
    """Moments of the vectorised generators at T/2 and T against 2000 paths of the unmodified reference generators
    (tests/golden/generator_stats.json, made by tests/golden/make_generator_stats.py) and, where one exists, the
    closed form: |mean - mean_ref| within 4.5 standard errors, standard deviations within 8 %."""
    import json, math, os
    from conftest import GOLDEN_DIR
    from neural_jump_ode.simulation import simulate_paths
    ref = json.load(open(os.path.join(GOLDEN_DIR, "generator_stats.json")))
    n_ref, n_steps, T = ref["n_paths"], ref["n_steps"], ref["T"]
    for name, r in ref["processes"].items():
        g = torch.Generator(device=device).manual_seed(11)
        _, X = simulate_paths(name, n, n_steps=n_steps, T=T, device=device, generator=g, **r["params"])
        X = X.double().cpu()
        for tag, col in (("mid", n_steps // 2), ("end", n_steps)):
            x = X[:, col]
            m, s = float(x.mean()), float(x.std())
            se = math.sqrt(r[f"std_{tag}"] ** 2 / n_ref + s * s / n)
            assert abs(m - r[f"mean_{tag}"]) <= 4.5 * se, (name, tag, m, r[f"mean_{tag}"], se)
            assert abs(s / r[f"std_{tag}"] - 1.0) <= 0.08, (name, tag, s, r[f"std_{tag}"])
        p = r["params"]
        xT = X[:, -1]
        if name == "black_scholes":            # E X_T = x0 e^{mu T}; Var = x0^2 e^{2 mu T} (e^{sigma^2 T} - 1)
            mean = p["x0"] * math.exp(p["mu"] * T)
            std = mean * math.sqrt(math.exp(p["sigma"] ** 2 * T) - 1.0)
        elif name == "ornstein_uhlenbeck":     # exact transition: mean mu + (x0 - mu) e^{-theta T}, var sigma^2 (1 - e^{-2 theta T}) / (2 theta)
            mean = p["mu"] + (p["x0"] - p["mu"]) * math.exp(-p["theta"] * T)
            std = p["sigma"] * math.sqrt((1.0 - math.exp(-2.0 * p["theta"] * T)) / (2.0 * p["theta"]))
        elif name == "heston":                 # Euler scheme: E X_T = x0 (1 + mu dt)^n
            mean, std = p["x0"] * (1.0 + p["mu"] * T / n_steps) ** n_steps, None
        else:
            continue
        assert abs(float(xT.mean()) - mean) <= 4.5 * float(xT.std()) / math.sqrt(n), (name, float(xT.mean()), mean)
        if std is not None:
            assert abs(float(xT.std()) / std - 1.0) <= 0.05, (name, float(xT.std()), std)


def test_generators_match_reference_moments_cpu():
    _check_against_reference_moments("cpu", n=8000)


def test_moment_weight_cache_ignores_recycled_ids():
    """The read-back cache of a device moment_weights tensor is keyed by id(): a collected tensor's id can be handed to
    a new tensor, whose weights must then be read afresh."""
    import weakref
    from neural_jump_ode.models import jump_ode as jo

    class Dead:
        pass
    d = Dead()
    ref = weakref.ref(d)
    del d
    w = torch.tensor([2.0, 7.0])
    jo._MW_CACHE[(id(w), w._version)] = (ref, [9.0, 9.0])          # what a recycled id looks like
    assert jo._moment_weights(w, 2) == (2.0, 7.0)
    assert jo._moment_weights(w, 2) == (2.0, 7.0)                   # second call hits the (now valid) entry
    w.mul_(2.0)
    assert jo._moment_weights(w, 2) == (4.0, 14.0)                  # in-place edit = new version = re-read


def test_packed_slice_and_gather():
    from neural_jump_ode import PackedBatch
    bt = [torch.tensor([0.0, 0.5, 1.0]), torch.tensor([0.0, 1.0]), torch.tensor([0.0, 0.2, 0.4, 1.0]), torch.tensor([0.3])]
    bv = [torch.arange(3.0).view(3, 1), 10 + torch.arange(2.0).view(2, 1), 20 + torch.arange(4.0).view(4, 1), torch.tensor([[30.0]])]
    b = PackedBatch.from_lists(bt, bv)
    s = b.slice(1, 3)
    assert s.B == 2 and s.sizes == [2, 4] and s.offsets.tolist() == [0, 2, 6]
    assert torch.equal(s.times, torch.cat(bt[1:3])) and torch.equal(s.values, torch.cat(bv[1:3]))
    assert b.slice(1, 3) is s                                       # cached: a wave keeps its schedule
    g = b.gather(torch.tensor([3, 0, 2]))
    assert g.sizes == [1, 3, 4] and g.offsets.tolist() == [0, 1, 4, 8]
    assert torch.equal(g.times, torch.cat([bt[3], bt[0], bt[2]])) and torch.equal(g.values, torch.cat([bv[3], bv[0], bv[2]]))
    with pytest.raises(ValueError):
        b.slice(2, 2)


def test_header_is_plain_c(tmp_path):
    """include/njode.h is the drop-in boundary: it must compile as C99 (no C++-isms, no torch types) and a C translation unit
    that takes the address of every declared entry point must link against the built library."""
    import re
    import shutil
    import subprocess
    from conftest import ROOT
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    hdr = os.path.join(ROOT, "include", "njode.h")
    subprocess.run(["gcc", "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-Werror", hdr], check=True)
    names = sorted(set(re.findall(r"\b(njode_[a-z0-9_]+)\s*\(", open(hdr).read())))
    assert len(names) >= 20
    src = tmp_path / "bind.c"
    src.write_text('#include "njode.h"\n#include <stdio.h>\nint main(void) {\n  const void* p[] = {' +
                   ", ".join(f"(const void*){n}" for n in names) + "};\n  printf(\"%d\\n\", (int)(sizeof p / sizeof p[0]));\n  return 0;\n}\n")
    lib_dir = os.path.join(ROOT, "neural-jump-ode_b200", "lib")
    if not os.path.exists(os.path.join(lib_dir, "libnjode_b200.so")):
        pytest.skip("library not built")
    exe = tmp_path / "bind"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", lib_dir,
                    "-lnjode_b200", f"-Wl,-rpath,{lib_dir}", "-Wl,--unresolved-symbols=ignore-in-shared-libs"], check=True)


def test_bench_flop_model_matches_survey():
    """bench.py's algorithmic flop count is SURVEY.md 8d's formula: 6 * S * MAC_ode per trajectory-step (25 728 for configs 1 / 3,
    12 864 shared, 100 608 at hidden 64, 791 040 at hidden 128 / 3 layers) and F = 6 S [E MAC_ode + n MAC_jump + (2n - B) MAC_out]."""
    import bench
    sep32 = dict(input_dim=1, hidden_dim=32, output_dim=1, num_moments=2)
    assert bench.mac_counts(sep32) == dict(S=2, ode=2144, jump=1056, out=1056)
    assert bench.algorithmic_flops(sep32, 1, 0, 0)["per_step_ode"] == 25728
    assert bench.algorithmic_flops(dict(sep32, shared_network=True), 1, 0, 0)["per_step_ode"] == 12864
    assert bench.mac_counts(dict(sep32, shared_network=True))["out"] == 32 * 32 + 32 * 2
    assert bench.algorithmic_flops(dict(sep32, hidden_dim=64), 1, 0, 0)["per_step_ode"] == 100608
    h128 = dict(input_dim=1, hidden_dim=128, output_dim=1, num_moments=2, n_hidden_layers=3)
    assert bench.mac_counts(h128)["ode"] == 65920
    assert bench.algorithmic_flops(h128, 1, 0, 0)["per_step_ode"] == 791040
    f = bench.algorithmic_flops(sep32, total_steps=1000, n_obs_total=100, n_traj=10)
    macs = 2 * (1000 * 2144 + 100 * 1056 + (200 - 10) * 1056)
    assert f["total"] == 6.0 * macs and f["fwd"] == 2.0 * macs and f["bwd"] == 4.0 * macs
    assert bench.METRIC == "trajectory-ODE-steps/sec (fwd+bwd)" and bench.DEFAULT_WORKLOAD == "heston_sep_b262144"
    assert bench.WORKLOADS[bench.DEFAULT_WORKLOAD]["B"] == 262144 and bench.WORKLOADS[bench.DEFAULT_WORKLOAD]["scaling"] == "strong"


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the oracle port on the host cores, a bounded sample): one JSON line with the same metric /
    unit / config.workload as the product arm, impl = reference, its own cpu_baseline block and an e2e block without copies."""
    import json
    import subprocess
    import sys
    import bench
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample", "3"], check=True, capture_output=True, text=True, timeout=600).stdout
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "trajectory-ODE-steps/s"
    assert d["config"]["workload"] == bench.DEFAULT_WORKLOAD and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # under torchrun only rank 0 works and prints
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    quiet = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                            "--cpu-sample", "3"], check=True, capture_output=True, text=True, timeout=600, env=env).stdout
    assert not [l for l in quiet.splitlines() if l.startswith("{")]
