#!/usr/bin/env python
"""Compact per-kernel summary of an `ncu --set full` report (read here, on the build box):
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [substring-of-kernel-name] > profiles/x.summary.txt
duration, tensor-pipe activity, issue-slot use, DRAM bytes, occupancy and the stall-sample histogram of every launch."""
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "duration us"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots used %"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("dram__bytes_read.sum", "dram read"),
        ("dram__bytes_write.sum", "dram write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
        ("smsp__warps_active.avg.per_cycle_active", "warps active / scheduler"),
        ("launch__registers_per_thread", "registers"),
        ("launch__grid_size", "grid"),
        ("launch__block_size", "block")]


def main():
    rep = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    ki = head.index("Kernel Name")
    for r in rows[2:]:
        if pat not in r[ki]:
            continue
        print("== %s" % r[ki][:110])
        col = {h: (r[i], units[i]) for i, h in enumerate(head)}
        for k, label in KEYS:
            if k in col and col[k][0] != "":
                print("   %-28s %s %s" % (label, col[k][0], col[k][1]))
        st = []
        for h, (v, _) in col.items():
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and v not in ("", "0"):
                st.append((int(float(v)), h[len("smsp__pcsamp_warps_issue_stalled_"):]))
        tot = sum(v for v, _ in st) or 1
        st.sort(reverse=True)
        print("   stall samples: " + ", ".join("%s %.0f%%" % (n, 100.0 * v / tot) for v, n in st[:7]))


if __name__ == "__main__":
    main()
