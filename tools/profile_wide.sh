#!/bin/bash
# ncu --set full capture of the three wide kernels on the config-4 shape (one launch each, second step), after the
# same command exited 0 without ncu.  Usage on the GPU box: bash tools/profile_wide.sh <tag> [workload] [extra bench args]
TAG=${1:-r2}; WL=${2:-heston_h128_l3}; shift; shift
mkdir -p gpurun_out
CMD="python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-cuda-graph $*"
$CMD > gpurun_out/prof_${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_wide_sweep|k_wide_wgrad" -s 3 -c 3 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/prof_${TAG}_ncu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/prof_${TAG}_ncu.log
