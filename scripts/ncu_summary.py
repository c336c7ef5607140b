"""Condense an .ncu-rep (one kernel launch, --set full) into the handful of numbers DESIGN.md / profiles/README.md quote.
usage: python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep > profiles/r01_ncu_x.txt"""
import csv, subprocess, sys

WANT = [
    ("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"), ("launch__occupancy_limit_registers", "CTAs/SM limit (registers)"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM limit (shared memory)"),
    ("gpu__time_duration.sum", "duration (ncu, cold, us or ms as printed)"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_sectors_srcunit_tex.sum", "L2 sectors from SMs (x32 B)"), ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 throughput % of peak"),
    ("smsp__inst_executed.sum", "warp instructions"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe busy %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle / issue"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving / issue"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                print(f"{label:45s} {vals[i]} {units[i]}")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
