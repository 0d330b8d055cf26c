#!/usr/bin/env python
"""Measurement of the rows AROUND the hot path (N1 trainer step, N2 on-device generators, N3
closed-form conditional moments, N4 dense-grid inference) on one B200.  bench.py measures the hot path itself.

    python tools/bench_rows.py [--rows N1,N2,N3,N4] [--out gpurun_out/rows.jsonl]

One JSON line per measurement: the GPU number (CUDA events, or host wall clock around a synchronised region where the
row's point IS the host work), a CPU baseline timed on this box's host cores where an oracle port exists (N1: the eager
port of the reference's training step, N2: oracle/paths_oracle.py -- pinned bit for bit to the reference; both through
bench.rows_cpu_baselines, bench.py being the one measurement script that executes oracle/), and the
unmodified reference's own timing from the build container (profiles/r2_reference_rows_build_container.json,
tools/ref_rows_cpu_timing.py) as labelled context where the reference code cannot travel (N3, N4).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neural-jump-ode_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

HESTON = dict(mu=0.5, kappa=2.0, theta=0.04, xi=0.5, rho=-0.5, x0=1.0, v0=0.04)
BS = dict(mu=0.1, sigma=0.5, x0=1.0)
OU = dict(theta=1.0, mu=0.5, sigma=0.3, x0=0.0)
LOSS = dict(ignore_first_continuity=True, moment_weights=[1.0, 10.0], variance_method="direct")


def ref_context():
    try:
        with open(os.path.join(ROOT, "profiles", "r2_reference_rows_build_container.json")) as f:
            return json.load(f)
    except OSError:
        return {}


def events_ms(fn, warmup=2, iters=5):
    """mean device time of fn() in ms (CUDA events on the current stream, synchronised on both sides)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def wall_ms(fn, warmup=1, iters=3):
    """mean host wall time of fn() in ms with the device drained on both sides (rows whose cost is host work)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / iters


def emit(out, rec):
    line = json.dumps(rec)
    print(line, flush=True)
    if out:
        out.write(line + "\n")
        out.flush()


def total_steps(model, batch):
    """trajectory-ODE-steps of a batch from the kernels' own schedule (as bench.py counts them)."""
    with torch.no_grad():
        model.forward_packed(batch)
    sched = next(iter(batch._schedules.values()))
    return int(sched.total_steps)


# ------------------------------------------------------------------------------------------------------------
def row_n1(out, ctx):
    """Trainer step (utils/training.py:78-101, :396).  (a) config 1's epoch -- 1000 Black-Scholes trajectories in
    mini-batches of 128 with a 104 tail, FlatAdam(weight_decay=5e-4) -- through train_epoch_packed on a device-resident
    dataset; (b) the same epoch through the LIST API with torch.optim.Adam, i.e. the reference's own call pattern against
    the drop-in; (c) one full-batch training step of configs[2] (262 144 Heston trajectories) incl. the optimiser;
    (d) njode_adam_step alone."""
    from neural_jump_ode import NeuralJumpODE, nj_ode_loss
    from neural_jump_ode.optim import FlatAdam
    from neural_jump_ode.simulation import make_packed_batch
    from neural_jump_ode.training import train_epoch_packed
    from bench import rows_cpu_baselines              # (the CPU legs live in bench.py: it owns every use of oracle/)
    dev = "cuda:0"
    mk = dict(input_dim=1, hidden_dim=32, output_dim=1, dt_ode_step=0.01, num_moments=2)
    data = make_packed_batch("black_scholes", 1000, 0.1, n_steps=100, T=1.0, device=dev, seed=5, **BS)
    torch.manual_seed(0)
    model = NeuralJumpODE(**mk).to(dev)
    steps = total_steps(model, data)
    opt = FlatAdam(model.parameters(), lr=1e-3, weight_decay=5e-4)
    mw = torch.tensor([1.0, 10.0], device=dev)
    kw = dict(ignore_first_continuity=True, moment_weights=mw, variance_method="direct")
    ms = wall_ms(lambda: train_epoch_packed(model, opt, data, batch_size=128, **kw), warmup=2, iters=5)

    # (b) reference call pattern: per-element .to(device) of list elements, zero_grad, model(lists), loss, backward, Adam, loss.item()
    bt = [t.cpu() for t in torch.split(data.times, data.sizes)]
    bv = [v.cpu() for v in torch.split(data.values, data.sizes)]
    torch.manual_seed(0)
    model2 = NeuralJumpODE(**mk).to(dev)
    opt2 = torch.optim.Adam(model2.parameters(), lr=1e-3, weight_decay=5e-4)

    def list_epoch():
        tot = 0.0
        for lo in range(0, 1000, 128):
            t = [x.to(dev) for x in bt[lo:lo + 128]]
            v = [x.to(dev) for x in bv[lo:lo + 128]]
            opt2.zero_grad()
            p, pb = model2(t, v)
            loss = nj_ode_loss(t, v, p, pb, **kw)
            loss.backward()
            opt2.step()
            tot += loss.item()
        return tot
    ms_list = wall_ms(list_epoch, warmup=1, iters=3)

    # CPU baseline on this box: the eager port of the reference's fwd + loss + bwd on one mini-batch of 128 + Adam
    cpu_s = rows_cpu_baselines()["training_step"](bt[:128], bv[:128], LOSS)
    emit(out, dict(row="N1", what="config-1 epoch (1000 trajectories, mini-batches of 128 + 104 tail, Adam weight_decay 5e-4)",
                   metric="training epoch, host wall clock (device drained)", packed_flatadam_ms=ms, list_api_torch_adam_ms=ms_list,
                   trajectories_per_s_packed=1000 / (ms * 1e-3), trajectories_per_s_list_api=1000 / (ms_list * 1e-3),
                   trajectory_ode_steps_per_epoch=steps, steps_per_s_packed=steps / (ms * 1e-3),
                   cpu_baseline=dict(kind="port", cores=torch.get_num_threads(), sample="one mini-batch of 128 (fwd + loss + bwd + Adam)",
                                     seconds=cpu_s, trajectories_per_s=128 / cpu_s),
                   reference_build_container=ctx.get("N1_training_step_config1_b128")))

    # (c) configs[2] full-batch training step incl. FlatAdam; (d) the Adam kernel alone
    del model2, opt2
    big = make_packed_batch("heston", 262144, 0.1, n_steps=200, T=1.0, device=dev, seed=1, **HESTON)
    torch.manual_seed(0)
    m3 = NeuralJumpODE(input_dim=1, hidden_dim=32, output_dim=1, dt_ode_step=0.005, num_moments=2).to(dev)
    steps3 = total_steps(m3, big)
    o3 = FlatAdam(m3.parameters(), lr=1e-3, weight_decay=5e-4)

    def big_step(with_opt):
        o3.zero_grad(set_to_none=True)
        p, pb = m3.forward_packed(big)
        loss = nj_ode_loss(big, None, p, pb, **kw)
        loss.backward()
        if with_opt:
            o3.step()
    ms_no = events_ms(lambda: big_step(False), warmup=2, iters=5)
    ms_with = events_ms(lambda: big_step(True), warmup=2, iters=5)
    ms_adam = events_ms(o3.step, warmup=5, iters=200)
    n_par = sum(p.numel() for p in m3.parameters())
    emit(out, dict(row="N1", what="configs[2] full-batch training step (262 144 Heston trajectories), eager, CUDA events",
                   fwd_loss_bwd_ms=ms_no, fwd_loss_bwd_adam_ms=ms_with, trajectory_ode_steps=steps3,
                   steps_per_s_with_optimiser=steps3 / (ms_with * 1e-3),
                   adam_step_alone_us=1e3 * ms_adam, parameters=n_par,
                   adam_note="one njode_adam_step launch on the flat buffers; 28 B per parameter -> latency-bound at this size"))


def row_n2(out, ctx):
    """On-device generators + observation sampler against the per-trajectory reference loops."""
    from neural_jump_ode.simulation import make_packed_batch, make_mixed_ragged_batch
    from bench import rows_cpu_baselines
    cpu_generate = rows_cpu_baselines()["generate"]
    dev = "cuda:0"
    cases = (("black_scholes", 262144, 100, BS, 64), ("ornstein_uhlenbeck", 262144, 100, OU, 64), ("heston", 262144, 200, HESTON, 32))
    refc = ctx.get("N2_generators", {})
    for proc, n, n_steps, pkw, n_cpu in cases:
        ms = events_ms(lambda: make_packed_batch(proc, n, 0.1, n_steps=n_steps, T=1.0, device=dev, seed=3, **pkw), warmup=1, iters=3)
        cpu_s = cpu_generate(proc, n_cpu, n_steps, **pkw)
        emit(out, dict(row="N2", what=f"{proc}: {n} paths x {n_steps} grid steps + observation sampling (obs 0.1) -> PackedBatch",
                       metric="trajectories/s (CUDA events)", ms=ms, value=n / (ms * 1e-3), grid_points_per_s=n * (n_steps + 1) / (ms * 1e-3),
                       cpu_baseline=dict(kind="port", cores=1, sample=f"{n_cpu} trajectories, oracle/paths_oracle.py (bit-exact restatement)",
                                         seconds=cpu_s, value=n_cpu / cpu_s, unit="trajectories/s"),
                       reference_build_container=refc.get(f"{proc}_n{n_steps}")))
    ms = events_ms(lambda: make_mixed_ragged_batch(131072, 0.02, 0.2, n_steps=100, T=1.0, device=dev, seed=3), warmup=1, iters=3)
    emit(out, dict(row="N2", what="config-5 mixed ragged batch: 131 072 paths (BS / OU / Heston / hybrid), per-path obs fraction U[0.02, 0.2]",
                   metric="trajectories/s (CUDA events)", ms=ms, value=131072 / (ms * 1e-3)))


def row_n3(out, ctx):
    """Closed-form conditional moments at the observations + the relative-loss metric, whole set in one call."""
    from neural_jump_ode import NeuralJumpODE
    from neural_jump_ode.simulation import make_packed_batch, conditional_moments_packed
    from neural_jump_ode.training import relative_loss_packed, validate_packed
    dev = "cuda:0"
    data = make_packed_batch("heston", 262144, 0.1, n_steps=200, T=1.0, device=dev, seed=2, **HESTON)
    pp = dict(mu=0.5, xi=0.5, kappa=2.0)
    ms = events_ms(lambda: conditional_moments_packed(data, "heston", num_moments=2, variance_method="direct", **pp), warmup=2, iters=10)
    emit(out, dict(row="N3", what="conditional_moments_packed: Heston, 2 moments, 262 144 trajectories x 20 observations",
                   metric="observations/s (CUDA events)", ms=ms, value=data.N / (ms * 1e-3),
                   reference_build_container=ctx.get("N3_conditional_moments_heston")))
    torch.manual_seed(0)
    model = NeuralJumpODE(input_dim=1, hidden_dim=32, output_dim=1, dt_ode_step=0.005, num_moments=2).to(dev)
    steps = total_steps(model, data)
    mw = torch.tensor([1.0, 10.0], device=dev)
    ms_val = wall_ms(lambda: validate_packed(model, data, ignore_first_continuity=True, moment_weights=mw), warmup=1, iters=3)
    ms_rel = wall_ms(lambda: relative_loss_packed(model, data, "heston", pp, moment_weights=mw), warmup=1, iters=3)
    emit(out, dict(row="N3", what="validate_packed / relative_loss_packed on the same set (forward-only sweep + loss [+ closed-form moments + second loss]), host wall incl. loss.item()",
                   validate_ms=ms_val, relative_loss_ms=ms_rel, trajectory_ode_steps=steps, forward_only_steps_per_s=steps / (ms_val * 1e-3)))


def row_n4(out, ctx):
    """Dense-grid inference (plotting.py:133-256): readouts at every grid time for a batch of trajectories."""
    from neural_jump_ode import NeuralJumpODE
    from neural_jump_ode.simulation import make_packed_batch
    dev = "cuda:0"
    for name, B, n_steps, mk, proc, pkw in (
            ("hidden 32, dt 0.01, Black-Scholes", 4096, 100, dict(input_dim=1, hidden_dim=32, output_dim=1, dt_ode_step=0.01, num_moments=2), "black_scholes", BS),
            ("hidden 128 / 3 layers / tanh, dt 0.005, Heston", 1024, 200, dict(input_dim=1, hidden_dim=128, output_dim=1, dt_ode_step=0.005, num_moments=2, n_hidden_layers=3, activation="tanh"), "heston", HESTON)):
        data = make_packed_batch(proc, B, 0.1, n_steps=n_steps, T=1.0, device=dev, seed=4, **pkw)
        torch.manual_seed(0)
        model = NeuralJumpODE(**mk).to(dev)
        grid = torch.linspace(0.0, 1.0, n_steps + 1, device=dev)
        ms = events_ms(lambda: model.predict_on_grid(data, grid), warmup=2, iters=5)
        emit(out, dict(row="N4", what=f"predict_on_grid: {B} trajectories x {n_steps + 1} grid points, {name}",
                       metric="grid points/s (CUDA events)", ms=ms, value=B * (n_steps + 1) / (ms * 1e-3),
                       reference_build_container=ctx.get("N4_dense_grid_one_trajectory_bs_n100") if n_steps == 100 else None))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", default="N1,N2,N3,N4")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    assert torch.cuda.is_available(), "bench_rows.py needs a CUDA device"
    torch.cuda.set_device(0)
    ctx = ref_context()
    out = open(args.out, "w") if args.out else None
    emit(out, dict(row="env", gpu=torch.cuda.get_device_name(0), host_threads=torch.get_num_threads(), cpu_count=os.cpu_count(), torch=torch.__version__))
    table = dict(N1=row_n1, N2=row_n2, N3=row_n3, N4=row_n4)
    for r in args.rows.split(","):
        try:
            table[r.strip()](out, ctx)
        except Exception as e:   # one failing row must not lose the others' numbers (the GPU budget is per call)
            emit(out, dict(row=r.strip(), error=f"{type(e).__name__}: {e}"))
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
