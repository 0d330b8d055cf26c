#!/bin/bash
# Final round-2 pass on the GPU box: parity suite, smoke(), the rows tool, every bench workload (tools/gpu_check_r2.sh + rows)
TAG=${1:-r2g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_${TAG}.log
run() { local n=$1; shift
  timeout 600 python bench.py "$@" > gpurun_out/bench_${TAG}_${n}.json 2> gpurun_out/bench_${TAG}_${n}.err; echo "bench $n rc=$?"; }
run default --steps 10 --warmup 3
run ou --workload ou_shared_b4096 --steps 20 --warmup 5 --no-cpu-baseline
run h128 --workload heston_h128_l3 --steps 3 --warmup 3 --no-cpu-baseline
run h64 --workload mixed_h64_ragged --steps 5 --warmup 3 --no-cpu-baseline
run h128_1m --workload heston_h128_l3_1m --batch 16384 --steps 2 --warmup 3 --no-cpu-baseline
run bs128 --workload bs_sep_b128 --steps 20 --warmup 5 --no-cpu-baseline
timeout 600 python tools/bench_rows.py --out gpurun_out/rows_${TAG}.jsonl > gpurun_out/rows_${TAG}.log 2>&1; echo "rows rc=$?"
python - <<PY
import json
for w in ("default","ou","h128","h64","h128_1m","bs128"):
    try:
        d=json.loads(open(f"gpurun_out/bench_${TAG}_{w}.json").read().strip().splitlines()[-1])
        r=d["roofline"]
        print(w, "value=%.4g e2e=%.4g ms=%.4g | %s %.3f ms frac=%.3f (%s; fp32-fma frac %.3f; whole step %.3f) | all:" % (d["value"], (d["e2e"] or {}).get("value", 0), d["ms_per_step"], r["kernel"], r["kernel_ms"], r["frac"], r["bound"], r["frac_of_fp32_fma_peak"], r["whole_step_frac_of_fp32_fma_peak"]), {k: round(v,3) for k,v in r["all_kernels_ms"].items()}, "smem" , (r.get("smem") or {}).get("frac"))
    except Exception as e:
        print(w, "ERR", e)
PY
