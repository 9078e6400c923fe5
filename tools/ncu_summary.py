"""Condense an .ncu-rep (ncu --set full) into a markdown table for profiles/.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_xxx.md"""

import csv
import io
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "ms"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "DRAM rd GB"),
    ("dram__bytes_write.sum", "DRAM wr GB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "LSU wavefronts %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wf %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank conflicts"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue %"),
    ("smsp__sass_inst_executed_op_local_ld.sum", "local ld"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    print("| kernel | " + " | ".join(label for _, label in COLS) + " |")
    print("|---|" + "---|" * len(COLS))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        cells = []
        for key, _ in COLS:
            v = r[idx[key]] if key in idx else ""
            try:
                f = float(v.replace(",", ""))
                v = ("%.3f" % f) if abs(f) < 1000 else ("%.0f" % f)
            except ValueError:
                pass
            cells.append(v)
        print("| `%s` | " % name + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
