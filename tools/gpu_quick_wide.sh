#!/bin/bash
# quick loop for the wide kernels: parity probe on both widths, then kernel times on the two bench shapes
TAG=${1:-q}
mkdir -p gpurun_out
for cfg in "128 3 tanh 60" "64 1 relu 60"; do
  t=$(echo $cfg | tr " " "_")
  timeout 250 python tests/wide_debug_probe.py $cfg > gpurun_out/wd_${TAG}_$t.log 2>&1; echo "== $cfg rc=$?"
  grep -E "vs f64|worst|status|WIDE|Error" gpurun_out/wd_${TAG}_$t.log | head -8
done
for w in heston_h128_l3 mixed_h64_ragged; do
  timeout 300 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${TAG}_$w.json 2> gpurun_out/bench_${TAG}_$w.err; echo "bench $w rc=$?"
done
python - <<PY
import json
for w in ("heston_h128_l3","mixed_h64_ragged"):
    try:
        d=json.loads(open(f"gpurun_out/bench_${TAG}_{w}.json").read().strip().splitlines()[-1])
        r=d["roofline"]
        print(w, "value=%.4g ms=%.4g whole-step/fma-peak %.3f |" % (d["value"], d["ms_per_step"], r["whole_step_frac_of_fp32_fma_peak"]), {k: round(v,3) for k,v in r["all_kernels_ms"].items()})
    except Exception as e:
        print(w, "ERR", e)
PY
