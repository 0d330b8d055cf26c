#!/bin/bash
# Runs on the GPU box: ncu captures of the final kernels (each only after the plain command exited 0).
TAG=${1:-v4}
mkdir -p gpurun_out
timeout 120 python tools/profile_step.py ou_shared_b4096 > gpurun_out/plain_ou.log 2>&1 || { echo "plain ou failed"; exit 1; }
timeout 120 python tools/profile_step.py heston_sep_b262144 16384 > gpurun_out/plain_heston.log 2>&1 || { echo "plain heston failed"; exit 1; }
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_tiled --launch-skip 2 -c 2 -f -o gpurun_out/prof_ou_${TAG} \
    python tools/profile_step.py ou_shared_b4096 > gpurun_out/ncu_ou.log 2>&1; echo "ncu ou rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_tiled_backward --launch-skip 1 -c 1 -f -o gpurun_out/prof_bwd_heston_${TAG} \
    python tools/profile_step.py heston_sep_b262144 16384 > gpurun_out/ncu_heston.log 2>&1; echo "ncu heston rc=$?"
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain_${TAG}.json 2>&1 || { echo "plain bench failed"; exit 1; }
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_${TAG}.csv
