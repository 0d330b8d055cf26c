#!/bin/bash
# Runs on the GPU box (final round-2 check): parity suite, smoke(), default bench, the rows tool.  Outputs under gpurun_out/.
TAG=${1:-r2f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"
tail -14 gpurun_out/pytest_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_${TAG}.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err; echo "bench default rc=$?"
timeout 600 python tools/bench_rows.py --out gpurun_out/rows_${TAG}.jsonl > gpurun_out/rows_${TAG}.log 2>&1; echo "rows rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${TAG}_default.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("default value=%.4g e2e=%.4g ms=%.4g | %s %.3f ms frac=%.3f (%s) fp32frac=%.3f" % (d["value"], (d["e2e"] or {}).get("value", 0), d["ms_per_step"], r["kernel"], r["kernel_ms"], r["frac"], r["bound"], r["frac_of_fp32_fma_peak"]))
print("smem view:", r.get("smem"))
print("cpu:", d.get("cpu_baseline"), "clocks:", d.get("clocks"))
PY
cat gpurun_out/rows_${TAG}.log | cut -c1-700
