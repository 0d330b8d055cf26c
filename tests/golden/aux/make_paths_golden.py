"""Paths and observation sets from the UNMODIFIED reference generators (simulation/data_generation.py:11-291), produced
in the build container and committed as tests/golden/aux/paths_ref.npz.  They pin oracle/paths_oracle.py (the CPU
baseline of the generator row, tools/bench_rows.py) bit for bit: same seeds, same RNG consumption, same float32 ops.
Also records how long the reference generators take per trajectory in this container (context only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/aux/make_paths_golden.py
"""
import importlib.util
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {
    "black_scholes": dict(obs_fraction=0.1, T=1.0, n_steps=100, mu=0.1, sigma=0.5, x0=1.0),
    "ornstein_uhlenbeck": dict(obs_fraction=0.1, T=1.0, n_steps=100, theta=1.0, mu=0.5, sigma=0.3, x0=0.0),
    "heston": dict(obs_fraction=0.1, T=1.0, n_steps=200, mu=0.5, kappa=2.0, theta=0.04, xi=0.5, rho=-0.5, x0=1.0, v0=0.04),
    "heston/short_ragged": dict(obs_fraction=0.35, T=0.5, n_steps=17, mu=0.0, kappa=1.0, theta=0.09, xi=0.3, rho=0.2, x0=2.0, v0=0.01),
}
N_TRAJ = 5


def main():
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location("ref_data_generation", "/root/reference/neural_jump_ode/simulation/data_generation.py")
    dg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(dg)
    out, meta = {}, {}
    for name, kw in CASES.items():
        proc = name.split("/")[0]
        t0 = time.perf_counter()
        bt, bv = dg.create_trajectory_batch(N_TRAJ, proc, **kw)
        ms = (time.perf_counter() - t0) * 1e3 / N_TRAJ
        out[f"{name}|times"] = torch.cat(bt).numpy()
        out[f"{name}|values"] = torch.cat(bv).numpy()
        out[f"{name}|sizes"] = np.array([len(t) for t in bt])
        meta[name] = dict(process=proc, kwargs=kw, n_traj=N_TRAJ, reference_ms_per_trajectory_build_container=ms)
        print(name, [len(t) for t in bt], f"{ms:.2f} ms / trajectory")
    out["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "paths_ref.npz"), **out)


if __name__ == "__main__":
    main()
