"""Per-kernel roofline rows of the wide flavour from bench lines (profiles/r2_bench_<tag>_{h128,h64}.json): algorithmic
TFLOP/s against the 3xTF32 tensor peak and algorithmic checkpoint GB/s against the measured HBM bandwidth.
usage: python tools/roofline_table.py <tag>"""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r2e"
peaks = json.load(open("MEASURED_PEAKS.json"))
tpeak, hpeak = peaks["bf16_tflops"] / 6.0, peaks["hbm_gbs"]
for w, H, L in (("h128", 128, 3), ("h64", 64, 1)):
    d = json.loads(open(f"profiles/r2_bench_{tag}_{w}.json").read().strip().splitlines()[-1])
    r, c = d["roofline"], d["config"]
    ms = r["all_kernels_ms"]
    steps, N, S = c["trajectory_ode_steps_per_gpu"], c["observations_per_gpu"], 2
    X = r["whole_step_tflops"] * 1e12 * d["ms_per_step"] * 1e-3 / 3          # flops of each of the three kernels
    plane = 4 * H
    by = {"k_wide_sweep<forward>": (1 + L) * plane, "k_wide_sweep<reverse> (data gradients)": 2 * (1 + L) * plane + 32,
          "k_wide_wgrad (weight gradients)": 2 * (1 + L) * plane + 32}
    print(f"{c['workload']}: {d['value']:.4g} steps/s, {d['ms_per_step']:.2f} ms per step, e2e {d['e2e']['value']:.4g}")
    tot = 0.0
    for k, v in ms.items():
        b = by[k] * (steps + N) * S
        tot += b
        print(f"  {k:45s} {v:6.2f} ms  tensor {X / (v * 1e-3) * 1e-12:4.0f} TFLOP/s = {X / (v * 1e-3) * 1e-12 / tpeak:.2f}   "
              f"hbm {b / 1e9:5.2f} GB {b / (v * 1e-3) / 1e9:5.0f} GB/s = {b / (v * 1e-3) / 1e9 / hpeak:.2f}")
    T = d["ms_per_step"]
    print(f"  whole step: tensor {3 * X / (T * 1e-3) * 1e-12 / tpeak:.2f}, hbm {tot / 1e9:.1f} GB = {tot / (T * 1e-3) / 1e9 / hpeak:.2f}, "
          f"x FP32-FMA peak {r['whole_step_frac_of_fp32_fma_peak']:.2f}")
