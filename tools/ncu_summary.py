"""Summarise an .ncu-rep (ncu --set full) into the handful of counters DESIGN.md argues from.
usage: python tools/ncu_summary.py <report.ncu-rep> [more reports...]   (prints one block per kernel launch)"""
import csv, io, subprocess, sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("lts__t_sectors.sum", "L2 sectors"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("lts__t_requests_srcunit_tex_op_red.sum", "L2 red requests"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX % of peak"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), CTAs"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), CTAs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long scoreboard"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected / issue"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle / issue"),
    ("sm__cycles_active.avg", "SM active cycles (avg)"),
    ("sm__cycles_active.max", "SM active cycles (max)"),
    ("sm__cycles_elapsed.max", "SM elapsed cycles"),
]


def main():
    for path in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        print("# %s" % path)
        for r in rows[2:]:
            print("## %s  grid %s block %s" % (r[col["Kernel Name"]][:90], r[col["Grid Size"]], r[col["Block Size"]]))
            for key, label in WANT:
                if key in col and r[col[key]] != "":
                    print("  %-34s %16s %s" % (label, r[col[key]], units[col[key]]))


if __name__ == "__main__":
    main()
