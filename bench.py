#!/usr/bin/env python
"""bench.py -- trajectory-ODE-steps/s (fwd+bwd, device-timed) of the Neural Jump ODE hot path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --steps K --warmup W    (CPU port of the reference on a bounded sample)

One "step" = forward sweep + nj_ode_loss + reverse sweep over one batch of the workload (Adam and
host packing excluded).  Default workload = BASELINE.json configs[2], the largest
configuration that fits one GPU: experiment_heston.py defaults (hidden 32, 1 layer, relu, 2 moments,
separate networks) scaled to 262 144 trajectories in ONE batch, n_steps 200, dt_ode_step 0.005, obs 0.1.
With --gpus N the SAME global batch is split N ways (strong scaling, as BASELINE asks: "262144
trajectories ... 1/2/4/8 B200").  Other workloads (--workload): configs[1] (OU shared, 4096 per GPU),
configs[0] at batch 128, the config-4 shape (hidden 128 / 3 layers / tanh, dt 0.001) at a small batch and
at its named size (131 072 trajectories per GPU = 1 M on eight, run in waves), the config-5 mixed ragged
hidden-64 batch.  Data is synthetic (on-device paths of that shape), weights are random-init of that
architecture.  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "neural-jump-ode_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1]
    "ou_shared_b4096": dict(
        scaling="weak", process="ornstein_uhlenbeck", pkw=dict(theta=1.0, mu=0.5, sigma=0.3, x0=0.0), B=4096, n_steps=100, T=1.0,
        obs_fraction=0.1, model=dict(input_dim=1, hidden_dim=32, output_dim=1, dt_ode_step=0.01, num_moments=2,
                                     n_hidden_layers=1, activation="identity", shared_network=True),
        loss=dict(ignore_first_continuity=True, moment_weights=[1.0, 10.0], variance_method="direct")),
    # BASELINE.json configs[0] at batch 128 (the reference's CPU-runnable case)
    "bs_sep_b128": dict(
        scaling="weak", process="black_scholes", pkw=dict(mu=0.1, sigma=0.5, x0=1.0), B=128, n_steps=100, T=1.0,
        obs_fraction=0.1, model=dict(input_dim=1, hidden_dim=32, output_dim=1, dt_ode_step=0.01, num_moments=2),
        loss=dict(ignore_first_continuity=True, moment_weights=[1.0, 10.0], variance_method="direct")),
    # BASELINE.json configs[2]: B is the GLOBAL batch, split over the ranks (strong scaling)
    "heston_sep_b262144": dict(
        scaling="strong", process="heston", pkw=dict(mu=0.5, kappa=2.0, theta=0.04, xi=0.5, rho=-0.5, x0=1.0, v0=0.04), B=262144,
        n_steps=200, T=1.0, obs_fraction=0.1,
        model=dict(input_dim=1, hidden_dim=32, output_dim=1, dt_ode_step=0.005, num_moments=2),
        loss=dict(ignore_first_continuity=True, moment_weights=[1.0, 10.0], variance_method="direct")),
    # BASELINE.json configs[4] (per-GPU slice of the sweep): mixed-process ragged batch, hidden 64 -> wide tcgen05 kernels
    "mixed_h64_ragged": dict(
        scaling="weak", process="mixed", pkw=dict(), B=32768, n_steps=100, T=1.0, obs_fraction=(0.02, 0.2),
        model=dict(input_dim=1, hidden_dim=64, output_dim=1, dt_ode_step=0.01, num_moments=2),
        loss=dict(ignore_first_continuity=True, moment_weights=[1.0, 10.0], variance_method="direct")),
    # BASELINE.json configs[3] shape at a small batch (hidden 128, 3 layers, tanh): one sweep, no waves
    "heston_h128_l3": dict(
        scaling="weak", process="heston", pkw=dict(mu=0.5, kappa=2.0, theta=0.04, xi=0.5, rho=-0.5, x0=1.0, v0=0.04), B=2048,
        n_steps=1000, T=1.0, obs_fraction=0.05,
        model=dict(input_dim=1, hidden_dim=128, output_dim=1, dt_ode_step=0.001, num_moments=2, n_hidden_layers=3,
                   activation="tanh"),
        loss=dict(ignore_first_continuity=True, moment_weights=[1.0, 10.0], variance_method="direct")),
    # BASELINE.json configs[3] at its named size: 1 M trajectories on 8 GPUs = 131 072 per GPU, run in waves of 4096
    # trajectories (10 MB of checkpoints per trajectory: NeuralJumpODE.forward_backward_waves)
    "heston_h128_l3_1m": dict(
        scaling="weak", process="heston", pkw=dict(mu=0.5, kappa=2.0, theta=0.04, xi=0.5, rho=-0.5, x0=1.0, v0=0.04), B=131072,
        wave=4096, n_steps=1000, T=1.0, obs_fraction=0.05,
        model=dict(input_dim=1, hidden_dim=128, output_dim=1, dt_ode_step=0.001, num_moments=2, n_hidden_layers=3,
                   activation="tanh"),
        loss=dict(ignore_first_continuity=True, moment_weights=[1.0, 10.0], variance_method="direct")),
}
DEFAULT_WORKLOAD = "heston_sep_b262144"
METRIC = "trajectory-ODE-steps/sec (fwd+bwd)"     # the same string in both arms: the driver divides one by the other


def make_batch(wl, n_traj, device, seed):
    """Synthetic on-device batch of the workload's shape."""
    from neural_jump_ode.simulation import make_packed_batch, make_mixed_ragged_batch
    if wl["process"] == "mixed":
        lo, hi = wl["obs_fraction"]
        return make_mixed_ragged_batch(n_traj, lo, hi, n_steps=wl["n_steps"], T=wl["T"], device=device, seed=seed)
    return make_packed_batch(wl["process"], n_traj, wl["obs_fraction"], n_steps=wl["n_steps"], T=wl["T"],
                             device=device, seed=seed, **wl["pkw"])


def mac_counts(mk):
    H, dx, L = mk["hidden_dim"], mk["input_dim"], mk.get("n_hidden_layers", 1)
    M = mk.get("num_moments", 1)
    shared = mk.get("shared_network", False)
    S = 1 if shared else M
    O = mk["output_dim"] * (M if shared else 1)
    return dict(S=S, ode=H * (H + dx + 2) + L * H * H, jump=dx * H + L * H * H, out=L * H * H + H * O)


def algorithmic_flops(mk, total_steps, n_obs_total, n_traj):
    """F = 6*S*[E*MAC_ode + n*MAC_jump + (2n-1)*MAC_out] (2 fwd + 4 bwd flop per MAC)."""
    c = mac_counts(mk)
    macs = c["S"] * (total_steps * c["ode"] + n_obs_total * c["jump"] + (2 * n_obs_total - n_traj) * c["out"])
    return dict(fwd=2.0 * macs, bwd=4.0 * macs, total=6.0 * macs, per_step_ode=6.0 * c["S"] * c["ode"])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_baseline(wl, n_traj, repeats=1):
    """Time the CPU port of the reference (oracle.run_port: same per-step eager op granularity as
    jump_ode.py) on `n_traj` trajectories of the workload.  Returns (steps/s, steps, seconds)."""
    from oracle import njode_oracle as orc
    from neural_jump_ode.simulation import make_packed_batch
    mk = wl["model"]
    batch = make_batch(wl, n_traj, "cpu", 1234)
    bt = list(torch.split(batch.times, batch.sizes))
    bv = list(torch.split(batch.values, batch.sizes))
    cfg = orc.make_cfg(mk["input_dim"], mk["hidden_dim"], mk["output_dim"], mk.get("dt_ode_step"),
                       mk.get("num_moments", 1), mk.get("n_hidden_layers", 1), mk.get("activation", "relu"),
                       mk.get("shared_network", False), mk.get("input_scaling", "identity"))
    P = orc.init_params(cfg, seed=0)
    best, steps = None, 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        r = orc.run_port(P, cfg, bt, bv, wl["loss"])
        dt = time.perf_counter() - t0
        steps = len(r["step_log"])
        best = dt if best is None else min(best, dt)
    return steps / best, steps, best


def rows_cpu_baselines():
    """CPU legs of tools/bench_rows.py (rows N1 / N2 of SURVEY.md 8f), kept HERE because bench.py's CPU-baseline legs are the
    one measurement place that may execute oracle/: returns callables timing the eager port of the reference's training step
    and the bit-exact restatement of its path generators."""
    from oracle import njode_oracle as orc
    from oracle import paths_oracle as po

    def training_step(bt, bv, loss_kwargs, lr=1e-3, weight_decay=5e-4):
        cfg = orc.make_cfg(1, 32, 1, 0.01, 2)
        P = orc.init_params(cfg, seed=0)
        t0 = time.perf_counter()
        r = orc.run_port(P, cfg, bt, bv, loss_kwargs)
        opt = torch.optim.Adam([torch.nn.Parameter(v.clone()) for v in P.values()], lr=lr, weight_decay=weight_decay)
        for p, g in zip(opt.param_groups[0]["params"], r["grads"].values()):
            p.grad = g
        opt.step()
        return time.perf_counter() - t0

    def generate(process, n, n_steps, **pkw):
        t0 = time.perf_counter()
        po.trajectory_batch(n, process, obs_fraction=0.1, T=1.0, n_steps=n_steps, **pkw)
        return time.perf_counter() - t0

    return dict(training_step=training_step, generate=generate)


def run_reference(args, wl, name):
    """--impl reference: the reference's algorithm on the host cores.  /root/reference (pure Python) does
    not exist on the GPU box and cannot be compiled into oracle/_ref, so the oracle port is timed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_traj = min(args.cpu_sample, 96)           # bounded sample per step (~3 s of CPU work)
    for _ in range(args.warmup):
        cpu_port_baseline(wl, max(2, n_traj // 8))
    t_tot, steps_tot = 0.0, 0
    for _ in range(args.steps):
        sps, steps, sec = cpu_port_baseline(wl, n_traj)
        t_tot += sec
        steps_tot += steps
    value = steps_tot / t_tot
    sample = f"{n_traj} trajectories of {name} per step ({steps_tot // args.steps} trajectory-ODE-steps), fwd+loss+bwd"
    out = {"impl": "reference", "metric": METRIC, "value": value,
           "unit": "trajectory-ODE-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "timing": "host wall clock (CPU)",
           "config": {"workload": name, "sample": sample},
           "cpu_baseline": {"value": value, "unit": "trajectory-ODE-steps/s", "cores": torch.get_num_threads(),
                            "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "trajectory-ODE-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# the kernels of each flavour that bench.py can time alone (njode_set_kernel_timing which = 1 forward / 2 reverse
# sweep / 3 weight-gradient GEMM), the share of the algorithmic flops each of them carries, and the pipe that bounds it
FLAVOURS = {
    "tiled": dict(bound="tensor", pipe="tcgen05 kind::tf32, 3xTF32 split (FP32-accurate); chain operands from TMEM, weight-gradient tiles from shared memory",
                  kernels={1: ("k_tiled_forward", "fwd", 1.0), 2: ("k_tiled_backward (data + weight gradients)", "bwd", 1.0)}),
    "wide": dict(bound="tensor", pipe="tcgen05 kind::tf32, 3xTF32 split (FP32-accurate); activations in TMEM, weights streamed by cp.async.bulk, weight gradients as a split-K GEMM",
                 kernels={1: ("k_wide_sweep<forward>", "fwd", 1.0), 2: ("k_wide_sweep<reverse> (data gradients)", "bwd", 0.5),
                          3: ("k_wide_wgrad (weight gradients)", "bwd", 0.5)}),
    "rowtile": dict(bound="fp32", pipe="FP32 FMA (CUDA cores), operands from shared memory",
                    kernels={1: ("k_rowtile_forward", "fwd", 1.0), 2: ("k_rowtile_backward", "bwd", 1.0)}),
    "generic": dict(bound="fp32", pipe="FP32 FMA (CUDA cores), warp per unit",
                    kernels={1: ("k_generic_forward", "fwd", 1.0), 2: ("k_generic_backward", "bwd", 1.0)}),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="override the workload's batch (global batch for strong-scaling workloads, per GPU otherwise)")
    ap.add_argument("--wave", type=int, default=None, help="trajectories per wave (workloads that run in waves)")
    ap.add_argument("--obs-fraction", type=float, default=None, help="override the workload's observation fraction (analysis: separates per-step from per-tile cost)")
    ap.add_argument("--kernel-impl", default="auto", choices=["auto", "generic", "tiled", "rowtile", "wide"])
    ap.add_argument("--cpu-sample", type=int, default=384, help="trajectories in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--no-cuda-graph", action="store_true", help="launch the timed steps eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    name = args.workload
    wl = dict(WORKLOADS[name])
    if args.batch:
        wl["B"] = args.batch
    if args.obs_fraction is not None:
        wl["obs_fraction"] = args.obs_fraction
    if args.wave:
        wl["wave"] = args.wave

    if args.impl == "reference":
        run_reference(args, wl, name)
        return

    import torch.distributed as dist
    from neural_jump_ode import NeuralJumpODE, nj_ode_loss, PackedBatch, _native as nat
    from neural_jump_ode.sharding import shard_bounds
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = nat.load()

    mk = wl["model"]
    torch.manual_seed(0)                       # identical replicas on every rank
    model = NeuralJumpODE(**mk)
    model.kernel_impl = args.kernel_impl
    model = model.to(dev)
    if world > 1:
        model.enable_data_parallel()           # gradient all-reduce (NCCL over NVLink) inside the reverse sweep
    params = model.flat_parameters()
    strong = wl["scaling"] == "strong"
    if strong:                                 # one global batch, every rank integrates its contiguous slice
        B_global = wl["B"]
        whole = make_batch(wl, B_global, dev, 1000)
        bounds = shard_bounds(B_global, world)
        batch = whole.slice(bounds[rank], bounds[rank + 1]) if world > 1 else whole
        batch = PackedBatch(batch.times.clone(), batch.values.clone(), batch.offsets.clone(), batch.sizes)
        del whole
    else:                                      # fixed batch per GPU
        batch = make_batch(wl, wl["B"], dev, 1000 + rank)
        B_global = wl["B"] * world
    B = batch.B
    wave = wl.get("wave")
    desc = model.descriptor()
    impl = nat.IMPL_NAME[lib.njode_selected_impl(desc)]
    lk = wl["loss"]

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    def run_on(b):
        for p in params:
            p.grad = None
        if wave:                               # checkpoints of the whole batch do not fit: waves of `wave` trajectories
            return model.forward_backward_waves(b, wave, traj_scale=1.0 / B_global, **lk)
        preds, before = model.forward_packed(b)
        loss = nj_ode_loss(b, None, preds, before, traj_scale=1.0 / B_global, **lk)
        loss.backward()                        # world > 1: ONE in-place all-reduce of the flat gradient inside backward
        return loss

    def step():
        return run_on(batch)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):               # --cache-data: schedules are built here, outside the timed region
        step()
    sync_all()
    subs = [batch.slice(lo, min(lo + wave, B)) for lo in range(0, B, wave)] if wave else [batch]
    scheds = [next(iter(b_._schedules.values())) for b_ in subs]
    total_steps_rank = sum(s_.total_steps for s_ in scheds)
    sched = scheds[0]

    # The step is a fixed sequence of ~10 launches on a cached batch (--cache-data): capture it once in a CUDA graph
    # and replay it, so the host (8 ranks share the box's cores) is out of the timed region's critical path.
    # Same kernels, same work; falls back to eager launches if capture is not possible.  Steps that run in waves
    # (hundreds of milliseconds each, checkpoint buffers re-allocated per wave) are launched eagerly.
    run_step, graphed = step, False
    if not args.no_cuda_graph and not wave:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step()                                 # allocator warm-up on the capture stream
            torch.cuda.current_stream().wait_stream(side)
            sync_all()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step()
            sync_all()
            graph.replay()
            sync_all()
            run_step, graphed = graph.replay, True
        except Exception as exc:                       # noqa: BLE001 - report and measure eagerly
            if rank == 0:
                print(f"bench.py: CUDA graph capture unavailable ({type(exc).__name__}: {exc}); eager launches", file=sys.stderr)
            torch.cuda.synchronize()
            run_step, graphed = step, False

    # ---- timed region: K steps, device-timed with CUDA events, L2 flushed between steps ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    lib.njode_kernel_launches(1)              # the library counts its own kernel launches
    sync_all()
    torch.cuda.profiler.start()                # `ncu --profile-from-start off` captures exactly the timed region
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()
        ev[i][0].record()
        run_step()
        ev[i][1].record()
    sync_all()
    wall = time.perf_counter() - wall0
    torch.cuda.profiler.stop()
    launches = int(lib.njode_kernel_launches(0))
    if graphed:                                # replays do not pass through the library's launch counter: count one eager step
        lib.njode_kernel_launches(1)
        step()
        torch.cuda.synchronize()
        launches = int(lib.njode_kernel_launches(0)) * args.steps
    clocks = sampler.stop() if rank == 0 else None
    ms = [a.elapsed_time(b) for a, b in ev]
    t_dev = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
    steps_all = torch.tensor([float(total_steps_rank)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        dist.all_reduce(steps_all)
    t_total_ms = float(t_dev.item())
    value = float(steps_all.item()) * args.steps / (t_total_ms * 1e-3)

    # ---- every sweep kernel of the flavour timed alone (events on its stream around that launch); the slowest is the
    #      dominant kernel of the roofline block.  With waves the one-shot events catch the first wave's launches, so
    #      the algorithmic flops are those of the first wave. ----
    fl_spec = FLAVOURS[impl]
    first = subs[0]
    fl = algorithmic_flops(mk, scheds[0].total_steps, first.N, first.B)
    kernel_ms = {}
    for which in fl_spec["kernels"]:
        samples = []
        for i in range(min(args.steps, 7)):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            # torch.cuda.Event creates its cudaEvent lazily on first record: record once so the handle exists
            e0.record(); e1.record()
            torch.cuda.synchronize()
            nat.check(lib.njode_set_kernel_timing(which, ctypes.c_void_p(e0.cuda_event), ctypes.c_void_p(e1.cuda_event)),
                      "njode_set_kernel_timing")
            run_on(first)
            torch.cuda.synchronize()
            samples.append(e0.elapsed_time(e1))
        samples.sort()
        kernel_ms[which] = samples[len(samples) // 2]
    dom = max(kernel_ms, key=kernel_ms.get)
    kname, kphase, kshare = fl_spec["kernels"][dom]
    dom_flop = fl[kphase] * kshare
    achieved = dom_flop / (kernel_ms[dom] * 1e-3) * 1e-12
    peak = ctypes.c_float(0.0)
    nat.check(lib.njode_ffma_peak(ctypes.byref(peak)), "njode_ffma_peak")
    peak_fma = float(peak.value)
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    measured = json.load(open(peaks_file)) if os.path.exists(peaks_file) else {}
    bf16 = measured.get("bf16_tflops", 1590.0)          # fallback: B200_PROFILING.md
    bf16_src = "MEASURED_PEAKS.json bf16_tflops (burst: the kernel is timed alone)" if "bf16_tflops" in measured else "fallback 1590 TFLOP/s (B200_PROFILING.md)"
    # measured DRAM traffic of the same kernel on this workload (one `ncu --set full` capture per entry, profiles/)
    traffic = None
    for tfile in ("r2_dram_traffic.json", "r1_dram_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tfile)
        if os.path.exists(tpath) and traffic is None:
            entry = json.load(open(tpath)).get(f"{name}:{first.B}", {})
            traffic = entry.get(f"which{dom}_dram_bytes_per_launch", entry.get("reverse_sweep_dram_bytes_per_launch") if dom == 2 else None)
    # algorithmic checkpoint traffic of the dominant kernel (the only large buffer): bytes per unit-slot and
    # stack that this kernel must move once, x (Euler steps + units) x stacks; slot padding and re-reads not counted
    Hh, Ll = mk["hidden_dim"], mk.get("n_hidden_layers", 1)
    n_stacks = 1 if mk.get("shared_network", False) else mk.get("num_moments", 1)
    plane = 4 * Hh
    ckpt_bytes = {"tiled": {1: 2 * plane, 2: 2 * plane},
                  "wide": {1: (1 + Ll) * plane, 2: 2 * (1 + Ll) * plane + 32, 3: 2 * (1 + Ll) * plane + 32},
                  "rowtile": {1: (1 + Ll) * plane, 2: (1 + Ll) * plane},
                  "generic": {1: plane, 2: plane}}[impl][dom]
    hbm_bytes = float(ckpt_bytes) * (scheds[0].total_steps + first.N) * n_stacks
    hbm_peak = measured.get("hbm_gbs", 7700.0)
    hbm_gbs = hbm_bytes / (kernel_ms[dom] * 1e-3) * 1e-9
    hbm_view = {"algorithmic_bytes_per_launch": hbm_bytes, "achieved_gbs": hbm_gbs, "peak_gbs": hbm_peak,
                "frac": hbm_gbs / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in measured else "fallback 7700 GB/s (B200_PROFILING.md)"}
    if fl_spec["bound"] == "tensor":
        # 3xTF32: three tf32 MMAs per FP32-accurate product, tf32 runs at half the bf16 rate -> bf16 / 6
        peak_used = bf16 / 6.0
        peak_source = (f"tensor pipe: {bf16_src} / 2 (kind::tf32) / 3 (the 3xTF32 split that makes the product FP32-accurate); "
                       "the FP32-FMA peak measured in this process (njode_ffma_peak, the north star's FP32 denominator) is given beside it")
    else:
        peak_used = peak_fma
        peak_source = "FP32-FMA peak measured in this process (njode_ffma_peak); MEASURED_PEAKS.json has no FP32 figure"
    compute_view = {"achieved_tflops": achieved, "peak_tflops": peak_used, "frac": achieved / peak_used, "peak_source": peak_source}
    if hbm_view["frac"] > achieved / peak_used:
        # the dominant kernel is closer to the HBM roof than to its compute roof: that is the roof that binds it
        head = {"bound": "hbm", "achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_view["frac"],
                "peak_source": hbm_view["peak_source"] + "; algorithmic checkpoint bytes of the kernel (DESIGN.md 5.2) / its event time"}
    else:
        head = {"bound": "tensor" if fl_spec["bound"] == "tensor" else "fp32", "achieved": achieved, "peak": peak_used,
                "unit": "TFLOP/s", "frac": achieved / peak_used, "peak_source": peak_source}
    roofline = {**head, "pipe": fl_spec["pipe"], "kernel": kname, "flavour": impl, "compute": compute_view,
                "traffic": traffic, "kernel_ms": kernel_ms[dom],
                "algorithmic_flop_per_launch": dom_flop,
                "all_kernels_ms": {fl_spec["kernels"][w][0]: kernel_ms[w] for w in kernel_ms},
                "hbm": hbm_view,
                "fp32_fma_peak_tflops": peak_fma, "frac_of_fp32_fma_peak": achieved / peak_fma,
                "whole_step_tflops": fl["total"] * (total_steps_rank / max(scheds[0].total_steps, 1)) * args.steps / (t_total_ms * 1e-3) * 1e-12,
                "whole_step_frac_of_fp32_fma_peak": fl["total"] * (total_steps_rank / max(scheds[0].total_steps, 1)) * args.steps / (t_total_ms * 1e-3) * 1e-12 / peak_fma}

    if impl == "tiled" and dom == 2 and sched.tile_rows == 128 and clocks and clocks.get("sm_mhz"):
        # Third view, for the H=32 reverse sweep only: neither roof above is the one that binds it.  By design it moves
        # 342 KB through shared memory per tile-step (48 SS-mode M64xN72xK8 MMAs read 206 KB of operand tiles, the row
        # workers write 136 KB of tile stores) and the tensor core's operand fetch and the CUDA cores share
        # the 128 B/cycle/SM port.  Design-derived bytes (NOT algorithmic): they say how close the kernel is to what this
        # formulation allows, the tensor / HBM views say what the formulation costs.
        tile_steps = float(scheds[0].tile_kmax.sum().item()) * n_stacks   # Euler tile-steps of one launch (every stack sweeps every tile)
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        smem_peak = 128.0 * sms * clocks["sm_mhz"] * 1e6 * 1e-9           # GB/s at the SM clock sampled under load
        smem_bytes = tile_steps * 342.0 * 1024
        smem_gbs = smem_bytes / (kernel_ms[dom] * 1e-3) * 1e-9
        roofline["smem"] = {"design_bytes_per_launch": smem_bytes, "achieved_gbs": smem_gbs, "peak_gbs": smem_peak,
                            "frac": smem_gbs / smem_peak, "tile_steps_per_launch": tile_steps,
                            "cycles_per_tile_step": kernel_ms[dom] * 1e-3 * clocks["sm_mhz"] * 1e6 * sms / max(tile_steps, 1.0),
                            "peak_source": "128 B/cycle/SM x SMs x median SM clock sampled during the timed region (readouts and jump-net phases of a tile not counted)"}

    # ---- end to end through the public API: pinned host inputs -> H2D -> schedule -> fwd/loss/bwd -> loss D2H ----
    e2e = None
    if not args.no_e2e:
        h_times = batch.times.cpu().pin_memory()
        h_values = batch.values.cpu().pin_memory()
        h_off = batch.offsets.cpu().pin_memory()
        sizes = batch.sizes

        def e2e_step():
            b = PackedBatch(h_times.to(dev, non_blocking=True), h_values.to(dev, non_blocking=True),
                            h_off.to(dev, non_blocking=True), sizes)
            return run_on(b).item()            # device -> host read of the step's result (this rank's share of the loss)

        n_e2e = args.steps if not wave else min(args.steps, 3)
        for _ in range(3 if not wave else 1):
            e2e_step()
        sync_all()
        e2e_ev = []
        for _ in range(n_e2e):
            flush.zero_()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            e2e_step()
            b_.record()
            e2e_ev.append((a, b_))
        sync_all()
        t_e2e = torch.tensor([sum(a.elapsed_time(b_) for a, b_ in e2e_ev)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        e2e_value = float(steps_all.item()) * n_e2e / (float(t_e2e.item()) * 1e-3)
        h2d = h_times.numel() * 4 + h_values.numel() * 4 + h_off.numel() * 8
        d2h = 4 + 8 * nat.HDR_WORDS * len(subs)
        e2e = {"value": e2e_value, "unit": "trajectory-ODE-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": n_e2e, "includes": "H2D of packed inputs from pinned memory, schedule build, fwd, loss, bwd, loss.item()"}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            sps, steps, sec = cpu_port_baseline(wl, args.cpu_sample)
            cpu = {"value": sps, "unit": "trajectory-ODE-steps/s", "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"{args.cpu_sample} trajectories of {name} ({steps} trajectory-ODE-steps, {sec:.1f} s), "
                             f"fwd+loss+bwd, oracle.run_port (eager per-step port of jump_ode.py)",
                   "host_cpus": os.cpu_count()}
        out = {"metric": METRIC, "value": value,
               "unit": "trajectory-ODE-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": t_total_ms / args.steps, "higher_is_better": True, "scaling": wl["scaling"],
               "vs_baseline": None, "dtype": "f32", "data": "synthetic", "timing": "device (CUDA events, max over ranks)",
               "config": {"workload": name, "process": wl["process"], "batch_per_gpu": B, "global_batch": B_global,
                          "n_steps": wl["n_steps"], "obs_fraction": wl["obs_fraction"], "model": mk,
                          "trajectory_ode_steps_per_gpu": total_steps_rank, "observations_per_gpu": batch.N,
                          "parallelism": f"dp{world}", "kernel_impl": args.kernel_impl, "kernel_flavour": impl,
                          "wave_trajectories": wave, "waves_per_step": len(subs),
                          "tile_rows": sched.tile_rows, "l2": "flushed between steps (256 MiB write)",
                          "cuda_graph": graphed,
                          "flop_per_trajectory_step_ode": fl["per_step_ode"]},
               "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
               "wall_s_timed_region": wall}
        print(json.dumps(out), flush=True)
    sys.stdout.flush()
    if world > 1:
        # tear-down must never hang the launcher: release the captured graph (it pins NCCL work) first, and leave
        # through a watchdog if destroy_process_group still blocks (seen once with a captured all-reduce)
        watchdog = threading.Timer(15.0, lambda: os._exit(0))
        watchdog.daemon = True
        watchdog.start()
        graph = None
        run_step = None
        torch.cuda.synchronize()
        try:
            dist.barrier()
            dist.destroy_process_group()
        finally:
            watchdog.cancel()


if __name__ == "__main__":
    main()
