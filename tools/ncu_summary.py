"""Turn an `ncu --set full --import-source on` report into the short text summary kept under profiles/.

usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-regex] > profiles/rNN_x.txt
Needs `ncu` on PATH (it is in the build container; no GPU required to read a report).
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rx = sys.argv[2] if len(sys.argv) > 2 else None
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, rows = raw[0], raw[1], raw[2:]
    kn = hdr.index("Kernel Name")
    print(f"# ncu summary of {rep}\n")
    for r in rows:
        if rx and not re.search(rx, r[kn]):
            continue
        print(f"## {r[kn]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:92s} {r[i]:>16s} {units[i]}")
        print()
    # per-kernel SASS statistics (instruction mix of the hottest loop, top stall sites)
    names = sorted({r[kn] for r in rows if not rx or re.search(rx, r[kn])})
    for name in names:
        short = re.sub(r"\(.*", "", name).split("::")[-1]
        short = re.sub(r"<.*", "", short)
        txt = ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + short])
        tab = list(csv.reader(io.StringIO(txt)))
        if len(tab) < 3:
            continue
        h = tab[1]
        try:
            iS, iE, iSrc = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
        except ValueError:
            continue
        data = [r for r in tab[2:] if len(r) == len(h) and r[iE].isdigit()]
        seen, uniq = set(), []
        for r in data:           # the page repeats once per launch of the same kernel
            if r[0] in seen:
                break
            seen.add(r[0])
            uniq.append(r)
        data = uniq
        tot_e = sum(int(r[iE]) for r in data) or 1
        tot_s = sum(int(r[iS] or 0) for r in data) or 1
        print(f"## SASS view: {name}\n  static SASS instructions {len(data)}, warp-instructions executed {tot_e}, pc samples {tot_s}")
        byop, bysmp = collections.Counter(), collections.Counter()
        for r in data:
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iSrc])
            op = m.group(2) if m else "?"
            byop[op] += int(r[iE])
            bysmp[op] += int(r[iS] or 0)
        print("  opcode mix (share of executed warp-instructions / share of pc samples):")
        for op, c in byop.most_common(18):
            print(f"    {op:10s} {100.0 * c / tot_e:5.1f}%  {100.0 * bysmp[op] / tot_s:5.1f}%")
        print("  top stall sites (samples, executions, SASS):")
        for r in sorted(data, key=lambda r: -int(r[iS] or 0))[:14]:
            print(f"    {r[iS]:>7s} {r[iE]:>9s}  {r[iSrc].strip()[:100]}")
        print()


if __name__ == "__main__":
    main()
