#!/bin/bash
# Runs on the GPU box: parity tests, then the three bench workloads.  Outputs under gpurun_out/.
TAG=${1:-run}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_${TAG}.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_ou.json 2> gpurun_out/bench_${TAG}_ou.err; echo "bench ou rc=$?"
timeout 300 python bench.py --workload heston_sep_b262144 --batch 65536 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_heston64k.json 2> gpurun_out/bench_${TAG}_heston.err; echo "bench heston rc=$?"
timeout 300 python bench.py --workload bs_sep_b128 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_bs128.json 2> gpurun_out/bench_${TAG}_bs.err; echo "bench bs rc=$?"
python - <<PY
import json
for w in ("ou","heston64k","bs128"):
    try:
        d=json.loads(open(f"gpurun_out/bench_${TAG}_{w}.json").read().strip().splitlines()[-1])
        print(w, "value=%.4g e2e=%.4g ms=%.4g bwd_ms=%.4g frac=%.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
    except Exception as e:
        print(w, "ERR", e)
PY
