"""Key raw metrics per launch of an ncu report: python profiles/raw_metrics.py rep.ncu-rep"""
import csv, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]
units = rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")][:70]
    print(name)
    print("   " + "  ".join(f"{w.split('.')[0].replace('launch__','').replace('pct_of_peak_sustained','')}={r[i]}{units[i] if units[i] not in ('','%') else ''}" for w, i in idx))
