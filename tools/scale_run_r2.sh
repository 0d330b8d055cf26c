#!/bin/bash
# Strong / weak scaling on one box (gpurun --gpus 8): the default workload (configs[2], strong) at N = 1, 2, 4, 8,
# config 4 at its named size (1 M trajectories = 131 072 per GPU, waves) at N = 8 and N = 1 (same per-GPU load: weak),
# the config-5 mixed batch at N = 8.  One JSON line per run under gpurun_out/scale_<tag>_*.json
TAG=${1:-r2}
mkdir -p gpurun_out
run() { # name n args...
  local name=$1 n=$2; shift; shift
  if [ "$n" = 1 ]; then
    timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/scale_${TAG}_${name}_n1.json 2> gpurun_out/scale_${TAG}_${name}_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n "$@" > gpurun_out/scale_${TAG}_${name}_n${n}.json 2> gpurun_out/scale_${TAG}_${name}_n${n}.err
  fi
  echo "$name N=$n rc=$?"
}
for n in 1 2 4 8; do run default $n --steps 10 --warmup 3 --no-cpu-baseline; done
for n in 1 8; do run ou $n --workload ou_shared_b4096 --steps 20 --warmup 5 --no-cpu-baseline; done
for n in 8 1; do run h128_1m $n --workload heston_h128_l3_1m --steps 2 --warmup 3 --no-cpu-baseline; done
for n in 8 1; do run h64 $n --workload mixed_h64_ragged --steps 5 --warmup 3 --no-cpu-baseline --no-e2e; done
for n in 8 1; do run h128 $n --workload heston_h128_l3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e; done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/scale_${TAG}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("scale_${TAG}_")[1][:-5], "value=%.4g e2e=%.4g ms=%.4g" % (d["value"], (d["e2e"] or {}).get("value", 0), d["ms_per_step"]))
    except Exception as e:
        print(f, "ERR", e)
PY
