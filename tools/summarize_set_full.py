"""Selected metrics of an `ncu -i <rep> --page raw --csv` dump (one row per profiled launch) as a small CSV for profiles/."""
import csv
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = [hdr.index(k) for k in KEEP if k in hdr]
w = csv.writer(sys.stdout)
w.writerow([hdr[i] for i in idx])
w.writerow([units[i] for i in idx])
for r in data:
    w.writerow([r[i] for i in idx])
