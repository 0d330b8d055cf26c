#!/bin/bash
# one iteration on the GPU box: parity suite, per-step / per-tile cost of the H=32 sweeps, default bench line
TAG=${1:-it}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_${TAG}.log
for f in 0.1 0.02; do
  timeout 300 python bench.py --batch 65536 --obs-fraction $f --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/tilecost_${TAG}_$f.json 2> gpurun_out/tilecost_${TAG}_$f.err; echo "rc=$?"
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_default.json 2> gpurun_out/bench_${TAG}_default.err; echo "bench default rc=$?"
python - <<PY
import json
r={}
for f in ("0.1","0.02"):
    d=json.loads(open(f"gpurun_out/tilecost_${TAG}_{f}.json").read().strip().splitlines()[-1])
    c=d["config"]; k=d["roofline"]["all_kernels_ms"]
    tiles=c["observations_per_gpu"]/128*2; ts=d["roofline"]["smem"]["tile_steps_per_launch"]
    r[f]=(ts,tiles,k["k_tiled_forward"],k["k_tiled_backward (data + weight gradients)"])
    print(f, "ms", d["ms_per_step"], k)
def fit(i, ctas):
    (s1,t1,*m1),(s2,t2,*m2)=r["0.1"],r["0.02"]
    y1=m1[i]*1e-3*1.965e9*ctas; y2=m2[i]*1e-3*1.965e9*ctas
    b=(y1-y2*s1/s2)/(t1-t2*s1/s2); a=(y2-b*t2)/s2
    return a,b
print("forward  cycles per tile-step %.0f, per tile %.0f" % fit(0,296))
print("backward cycles per tile-step %.0f, per tile %.0f" % fit(1,148))
d=json.loads(open("gpurun_out/bench_${TAG}_default.json").read().strip().splitlines()[-1])
print("default value=%.4g e2e=%.4g ms=%.4g" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["roofline"]["all_kernels_ms"])
PY
