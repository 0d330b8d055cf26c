#!/bin/bash
# per-step vs per-tile cost of the H=32 sweeps: the default shape at 65 536 trajectories with 20 and with 4 observations per path
mkdir -p gpurun_out
for f in 0.1 0.02; do
  timeout 300 python bench.py --batch 65536 --obs-fraction $f --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/tilecost_$f.json 2> gpurun_out/tilecost_$f.err; echo "rc=$?"
done
python - <<PY
import json
for f in ("0.1","0.02"):
    d=json.loads(open(f"gpurun_out/tilecost_{f}.json").read().strip().splitlines()[-1])
    r=d["roofline"]; c=d["config"]
    print(f, "steps", c["trajectory_ode_steps_per_gpu"], "obs", c["observations_per_gpu"], "ms", d["ms_per_step"], r["all_kernels_ms"], r.get("smem"))
PY
