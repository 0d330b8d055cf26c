"""Sample statistics of the UNMODIFIED reference generators (simulation/data_generation.py:11-218), produced in the
build container and committed as tests/golden/generator_stats.json: mean and standard deviation of X at T/2 and T
over 2000 seeded paths per process, with the process parameters used.  The vectorised device generators
(neural_jump_ode/simulation/device_paths.py) use their own RNG stream, so they are checked against these moments
(tests/test_host.py on the CPU, tests/test_gpu_training.py on the device), not path by path.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_generator_stats.py
"""
import importlib.util
import json
import os
import sys

import numpy as np

REF = "/root/reference/neural_jump_ode"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location("ref_data_generation", f"{REF}/simulation/data_generation.py")
    dg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(dg)
    n, n_steps, T = 2000, 100, 1.0
    procs = {
        "black_scholes": (dg.generate_black_scholes, dict(mu=0.1, sigma=0.5, x0=1.0)),
        "ornstein_uhlenbeck": (dg.generate_ou, dict(theta=1.0, mu=0.5, sigma=0.3, x0=0.0)),
        "heston": (dg.generate_heston, dict(mu=0.5, kappa=2.0, theta=0.04, xi=0.5, rho=-0.5, x0=1.0, v0=0.04)),
        "hybrid_ou_bs": (dg.generate_hybrid_ou_bs, dict(theta_ou=1.0, mu_ou=0.5, sigma_ou=0.3, mu_bs=0.1, sigma_bs=0.3,
                                                         x0=1.0, switch_time=0.5)),
    }
    out = {"n_paths": n, "n_steps": n_steps, "T": T, "processes": {}}
    for name, (fn, kw) in procs.items():
        mid, end = [], []
        for seed in range(n):
            res = fn(T=T, n_steps=n_steps, seed=seed, **kw)          # (times, X) -- heston also returns the variance path
            x = np.asarray(res[1], dtype=np.float64).reshape(-1)
            if not np.isfinite(x).all():                 # hybrid: ~1 in 2000 paths is NaN in the reference (log of OU <= 0)
                continue
            mid.append(x[n_steps // 2])
            end.append(x[-1])
        mid, end = np.array(mid), np.array(end)
        out["processes"][name] = dict(params=kw, n_finite=len(end), mean_mid=float(mid.mean()), std_mid=float(mid.std(ddof=1)),
                                      mean_end=float(end.mean()), std_end=float(end.std(ddof=1)))
        print(name, out["processes"][name])
    with open(os.path.join(HERE, "generator_stats.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
