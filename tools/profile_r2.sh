#!/bin/bash
# Round-2 profiling pass on the GPU box (each ncu run only after the same command exited 0 without ncu).
#   1. launch list of the default bench command (gpu__time_duration per launch)
#   2. DRAM bytes per launch of the sweep kernels at the bench sizes (two counters: one pass, no replay of the 28 GB step)
#   3. ncu --set full of the wide kernels (H = 128 and H = 64) and of the H = 32 kernels at reduced batch
TAG=${1:-r2}
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e"
# 1
timeout 300 $B --steps 2 --warmup 3 > gpurun_out/${TAG}_plain_default.json 2> gpurun_out/${TAG}_plain_default.err || { echo "plain default failed"; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|^Device" -c 400 --csv --log-file gpurun_out/${TAG}_launches_default.csv \
    $B --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "launch list rc=$?"
# 2
for spec in "heston_sep_b262144:" "heston_h128_l3:" "mixed_h64_ragged:" "ou_shared_b4096:"; do
  w=${spec%%:*}
  timeout 300 $B --workload $w --steps 1 --warmup 3 --no-cuda-graph > gpurun_out/${TAG}_plain_$w.json 2>/dev/null || { echo "plain $w failed"; continue; }
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k regex:"k_tiled_forward|k_tiled_backward|k_wide_sweep|k_wide_wgrad" -s 6 -c 6 --csv --log-file gpurun_out/${TAG}_dram_$w.csv \
      $B --workload $w --steps 1 --warmup 3 --no-cuda-graph > gpurun_out/${TAG}_ncu_dram_$w.log 2>&1; echo "dram $w rc=$?"
done
# 3
full() { # name, kernel regex, skip, count, bench args...
  local n=$1 k=$2 s=$3 c=$4; shift 4
  timeout 300 $B --steps 1 --warmup 3 --no-cuda-graph "$@" > gpurun_out/${TAG}_plain_full_$n.json 2>/dev/null || { echo "plain full $n failed"; return; }
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$k" -s $s -c $c -f -o gpurun_out/${TAG}_full_$n \
      $B --steps 1 --warmup 3 --no-cuda-graph "$@" > gpurun_out/${TAG}_ncu_full_$n.log 2>&1; echo "full $n rc=$?"
}
full h128 "k_wide_sweep|k_wide_wgrad" 3 3 --workload heston_h128_l3 --batch 1024
full h64 "k_wide_sweep|k_wide_wgrad" 3 3 --workload mixed_h64_ragged --batch 8192
full h32 "k_tiled_forward|k_tiled_backward" 2 2 --workload heston_sep_b262144 --batch 32768
# summaries on the box (gpurun brings back at most 64 MiB: only the H = 128 report travels)
for n in h128 h64; do
  python tools/ncu_summary.py gpurun_out/${TAG}_full_$n.ncu-rep "k_wide" > gpurun_out/${TAG}_ncu_full_wide_$n.txt 2>&1
  python tools/ncu_lines.py gpurun_out/${TAG}_full_$n.ncu-rep "k_wide" 14 > gpurun_out/${TAG}_ncu_lines_wide_$n.txt 2>&1
done
python tools/ncu_summary.py gpurun_out/${TAG}_full_h32.ncu-rep "k_tiled" > gpurun_out/${TAG}_ncu_full_tiled_h32.txt 2>&1
rm -f gpurun_out/${TAG}_full_h64.ncu-rep gpurun_out/${TAG}_full_h32.ncu-rep
ls -la gpurun_out/${TAG}_*
