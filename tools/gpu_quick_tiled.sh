#!/bin/bash
# quick loop for the H = 32 tensor-core kernels: parity tests, then the default bench workload (kernel times)
TAG=${1:-t}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiled or auto" > gpurun_out/tiled_${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/tiled_${TAG}_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${TAG}_default.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("default value=%.4g ms=%.4g frac=%.3f |" % (d["value"], d["ms_per_step"], r["frac"]), {k: round(v,3) for k,v in r["all_kernels_ms"].items()})
PY
