#!/usr/bin/env python3
"""Print the handful of ncu metrics this project tracks from a .ncu-rep (or its --page raw --csv dump)."""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"


def main():
    path = sys.argv[1]
    if path.endswith(".ncu-rep"):
        text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(text.splitlines()))
    else:
        rows = list(csv.reader(open(path)))
    hdr = rows[0]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:90] if "Kernel Name" in hdr else "")
        for w in WANT:
            if w in hdr:
                print("  %-80s %s" % (w, r[hdr.index(w)]))
        st = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        print("  stalls (warps per issue-active):", ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:7]))


if __name__ == "__main__":
    main()
