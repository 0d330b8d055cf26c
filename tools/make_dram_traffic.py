"""profiles/r2_dram_traffic.json from the ncu metric runs of tools/profile_r2.sh (dram__bytes_read/write.sum per launch of
every sweep kernel at the bench size).  bench.py looks an entry up by "<workload>:<per-GPU batch>".
usage: python tools/make_dram_traffic.py <tag> > profiles/r2_dram_traffic.json"""
import csv, glob, json, os, re, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r2f"
out = {"_how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none on "
               "`bench.py --workload W --steps 1 --warmup 3 --no-cuda-graph --no-e2e --no-cpu-baseline` (tools/profile_r2.sh), "
               "mean over the captured launches of each kernel; bytes = read + write"}
which_of = [("k_tiled_forward", 1), ("k_tiled_backward", 2), (r"k_wide_sweep<\d+, \d+, 0>", 1), (r"k_wide_sweep<\d+, \d+, 1>", 2), ("k_wide_wgrad", 3)]
for path in sorted(glob.glob(f"gpurun_out/{tag}_dram_*.csv")):
    wl = os.path.basename(path)[len(tag) + 6:-4]
    plain = json.loads(open(f"gpurun_out/{tag}_plain_{wl}.json").read().strip().splitlines()[-1])
    B = plain["config"]["batch_per_gpu"]
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"')) if len(r) > 14 and r[0] != "ID"]
    per = {}
    for r in rows:
        name, metric, val = r[4], r[12], float(r[14])
        for rx, w in which_of:
            if re.search(rx, name):
                per.setdefault(w, {}).setdefault(r[0], {})[metric] = val
    entry = {}
    for w, launches in sorted(per.items()):
        n = len(launches)
        rd = sum(v["dram__bytes_read.sum"] for v in launches.values()) / n
        wr = sum(v["dram__bytes_write.sum"] for v in launches.values()) / n
        ns = sum(v["gpu__time_duration.sum"] for v in launches.values()) / n
        entry[f"which{w}_dram_bytes_per_launch"] = rd + wr
        entry[f"which{w}_read_write_ms"] = [rd, wr, ns * 1e-6]
    out[f"{wl}:{B}"] = entry
print(json.dumps(out, indent=1))
