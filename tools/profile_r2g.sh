#!/bin/bash
# Final profiling pass for the H = 32 (tiled) kernels after the per-tile work of round 2 (each ncu run only after the same
# command exited 0 without ncu): launch list of the default bench, DRAM bytes per launch at the bench sizes, ncu --set full.
TAG=${1:-r2g}
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e"
timeout 300 $B --steps 2 --warmup 3 > gpurun_out/${TAG}_plain_default.json 2> gpurun_out/${TAG}_plain_default.err || { echo "plain default failed"; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|^Device" -c 400 --csv --log-file gpurun_out/${TAG}_launches_default.csv \
    $B --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "launch list rc=$?"
for w in heston_sep_b262144 ou_shared_b4096; do
  timeout 300 $B --workload $w --steps 1 --warmup 3 --no-cuda-graph > gpurun_out/${TAG}_plain_$w.json 2>/dev/null || { echo "plain $w failed"; continue; }
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k regex:"k_tiled_forward|k_tiled_backward" -s 6 -c 6 --csv --log-file gpurun_out/${TAG}_dram_$w.csv \
      $B --workload $w --steps 1 --warmup 3 --no-cuda-graph > gpurun_out/${TAG}_ncu_dram_$w.log 2>&1; echo "dram $w rc=$?"
done
timeout 300 $B --steps 1 --warmup 3 --no-cuda-graph --batch 32768 > gpurun_out/${TAG}_plain_full_h32.json 2>/dev/null || { echo "plain full failed"; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_tiled_forward|k_tiled_backward" -s 2 -c 2 -f -o gpurun_out/${TAG}_full_h32 \
    $B --steps 1 --warmup 3 --no-cuda-graph --batch 32768 > gpurun_out/${TAG}_ncu_full_h32.log 2>&1; echo "full h32 rc=$?"
python tools/ncu_summary.py gpurun_out/${TAG}_full_h32.ncu-rep "k_tiled" > gpurun_out/${TAG}_ncu_full_tiled_h32.txt 2>&1
python tools/ncu_lines.py gpurun_out/${TAG}_full_h32.ncu-rep "k_tiled" 14 > gpurun_out/${TAG}_ncu_lines_tiled_h32.txt 2>&1
ls -la gpurun_out/${TAG}_* | awk '{print $5, $9}'
