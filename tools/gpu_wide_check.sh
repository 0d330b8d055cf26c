#!/bin/bash
# Runs on the GPU box: full parity suite, then the wide-flavour bench workloads (new kernels and the row-tiled ones they replace).
TAG=${1:-run}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_${TAG}.log
for impl in auto rowtile; do
  timeout 300 python bench.py --workload heston_h128_l3 --kernel-impl $impl --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_h128_${impl}.json 2> gpurun_out/bench_${TAG}_h128_${impl}.err; echo "bench h128 $impl rc=$?"
  timeout 300 python bench.py --workload mixed_h64_ragged --kernel-impl $impl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_h64_${impl}.json 2> gpurun_out/bench_${TAG}_h64_${impl}.err; echo "bench h64 $impl rc=$?"
done
python - <<PY
import json
for w in ("h128_auto","h128_rowtile","h64_auto","h64_rowtile"):
    try:
        d=json.loads(open(f"gpurun_out/bench_${TAG}_{w}.json").read().strip().splitlines()[-1])
        print(w, "value=%.4g e2e=%.4g ms=%.4g bwd_ms=%.4g frac=%.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
    except Exception as e:
        print(w, "ERR", e)
PY
