"""Per-source-line summary of an ncu report captured with --import-source on (binary built with -lineinfo):
share of warp-stall samples and of executed instructions per line, top N.  usage: ncu_lines.py report.ncu-rep [N]
Also prints the handful of raw metrics the roofline discussion needs."""
import csv, io, subprocess, sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[-1]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:90s} {rows[1][i]:>10s} {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, h, agg = None, None, []
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) >= 2 and r[0] == "Line No":
        h = r
    elif h and r and r[0].isdigit():
        try:
            agg.append((int(r[4]), int(r[h.index("Instructions Executed")]), cur, r[0], r[1][:120]))
        except (ValueError, IndexError):
            pass
ts, ti = sum(a[0] for a in agg) or 1, sum(a[1] for a in agg) or 1
print(f"-- {ts} stall samples, {ti} instructions; share of samples / share of instructions per line")
for s, i, f, l, t in sorted(agg, reverse=True)[:top]:
    print(f"{100 * s / ts:5.1f}% {100 * i / ti:5.1f}%i {f}:{l:>4s} {t}")
