"""Summarise an .ncu-rep (raw page) into a compact per-kernel table: python tools/ncu_summary.py rep [out.txt]"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "inst"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel")]
out = []
for r in rows[2:]:
    name = r[idx["Kernel Name"]][:60]
    parts = [name]
    for key, lab in want:
        if key in idx:
            v = r[idx[key]]
            try:
                v = "%.4g" % float(v.replace(",", ""))
            except ValueError:
                pass
            parts.append("%s=%s%s" % (lab, v, units[idx[key]] if lab in ("time", "dram_rd", "dram_wr") else ""))
    out.append("  ".join(parts))
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
