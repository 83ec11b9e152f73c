"""Key metrics of .ncu-rep files as a compact table (reads with `ncu -i ... --page raw --csv`)."""
import csv, io, subprocess, sys
WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("lts__t_sector_hit_rate.pct", "L2hit%"), ("l1tex__t_sector_hit_rate.pct", "L1hit%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_lsb"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_ssb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_bar")]
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(f"== {r[idx['Kernel Name']].split('(')[0]}  [{path.split('/')[-1]}]")
        print("   " + "  ".join(f"{short}={r[idx[m]]}{units[idx[m]] if units[idx[m]] in ('ms','us','Gbyte','Mbyte','Kbyte','byte') else ''}"
                                for m, short in WANT if m in idx))
