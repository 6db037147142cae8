"""Summarise `ncu --page raw --csv` exports (one row per captured launch) into the handful of
counters DESIGN.md / profiles/ quote: duration, DRAM bytes, pipe utilisation, shared-memory
wavefronts and bank conflicts, issue-stall reasons."""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration_ns"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pipe_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_throttle"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    units = rows[1]
    name_i = hdr.index("Kernel Name")
    out = []
    for r in rows[2:]:
        rec = {"kernel": r[name_i].split("(")[0].replace("void ", "")}
        for k, short in KEYS:
            if k in hdr:
                i = hdr.index(k)
                rec[short] = r[i] + (" " + units[i] if short.startswith("dram_r") or short.startswith("dram_w") else "")
        out.append(rec)
    for rec in out:
        print(rec["kernel"])
        for k, v in rec.items():
            if k != "kernel":
                print(f"    {k:22s} {v}")


if __name__ == "__main__":
    main(sys.argv[1])
