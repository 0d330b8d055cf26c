"""Aggregate the pc samples of an `ncu --set full --import-source on` report by CUDA source line (needs -lineinfo):
which statements the warps of a kernel sit on.   python tools/ncu_lines.py report.ncu-rep kernel-regex [top]"""
import collections, csv, io, os, re, subprocess, sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
kn = rows[0].index("Kernel Name")
for idx, r in enumerate(rows[2:]):
    name = r[kn]
    if not re.search(rx, name):
        continue
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--launch-skip", str(idx),
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    agg, total = collections.Counter(), 0
    fname, cur, c_smp = "?", "?", None
    for row in csv.reader(io.StringIO(out)):
        if not row:
            continue
        if row[0] == "File Path":
            fname = os.path.basename(row[1])
        elif row[0] == "Line No":
            c_smp = row.index("# Samples")
        elif c_smp is not None and row[0].strip().isdigit():
            cur = f"{fname}:{row[0]}  {row[1].strip()[:120]}"
        elif c_smp is not None and row[0] == "" and len(row) > c_smp:
            try:
                n = int(float(row[c_smp] or 0))
            except ValueError:
                n = 0
            agg[cur] += n
            total += n
    print(f"## {name[:90]}  ({total} samples)")
    for line, n in agg.most_common(top):
        print(f"  {100.0 * n / max(total, 1):5.1f}%  {line}")
