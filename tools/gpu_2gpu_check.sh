#!/bin/bash
# two B200s: the NCCL gradient-equality test (skipped on one GPU) and the default / ou bench lines at N = 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -q -k "nccl" > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_default.json 2> gpurun_out/bench_2gpu_default.err; echo "default n2 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload ou_shared_b4096 --steps 20 --warmup 5 > gpurun_out/bench_2gpu_ou.json 2> gpurun_out/bench_2gpu_ou.err; echo "ou n2 rc=$?"
python - <<PY
import json
for w in ("default","ou"):
    try:
        d=json.loads(open(f"gpurun_out/bench_2gpu_{w}.json").read().strip().splitlines()[-1])
        print(w, "n_gpus", d["n_gpus"], "value=%.4g e2e=%.4g ms=%.4g" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
    except Exception as e:
        print(w, "ERR", e)
PY
