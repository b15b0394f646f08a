"""Per-launch table of an ncu report: kernel, duration, DRAM bytes read / written, tensor-pipe and issue activity, registers.
usage: ncu_table.py report.ncu-rep"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
cols = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__grid_size"]
idx = [hdr.index(c) for c in cols if c in hdr]
print(" | ".join(f"{hdr[i]} [{units[i]}]" for i in idx))
for r in rows[2:]:
    print(" | ".join((r[i][:90] if hdr[i] == "Kernel Name" else r[i]) for i in idx))
