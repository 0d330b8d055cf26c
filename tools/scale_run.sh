#!/bin/bash
# usage: tools/scale_run.sh N ["ou heston64k ..."]  -- the bench workloads on N GPUs of this box (weak scaling), JSON lines under gpurun_out/
N=$1
run() { # name, args...
  local name=$1; shift
  if [ "$N" = "1" ]; then timeout 400 python bench.py --gpus 1 --no-cpu-baseline "$@" 2>/dev/null | tail -1 > gpurun_out/scale_${name}_n${N}.json
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N "$@" 2>/dev/null | tail -1 > gpurun_out/scale_${name}_n${N}.json; fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_${name}_n${N}.json").read().strip())
    print("${name} N=%d value=%.4g e2e=%.4g ms=%.4g" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"]))
except Exception as e:
    print("${name} N=${N} ERR", e)
PY
}
WHICH=${2:-"ou heston64k mixed_h64 h128"}
for w in $WHICH; do
  case $w in
    ou) run ou --steps 20 --warmup 5 ;;
    heston64k) run heston64k --workload heston_sep_b262144 --batch 65536 --steps 5 --warmup 3 ;;
    mixed_h64) run mixed_h64 --workload mixed_h64_ragged --batch 32768 --steps 3 --warmup 3 ;;
    h128) run h128 --workload heston_h128_l3 --batch 2048 --steps 3 --warmup 3 ;;
  esac
done
