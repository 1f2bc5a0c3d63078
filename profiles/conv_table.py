"""Per-launch table of an ncu raw CSV (sections pass): python profiles/conv_table.py raw.csv [kernel substring]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
pat = sys.argv[2] if len(sys.argv) > 2 else ""
hdr, units = rows[0], rows[1]
c = {h: i for i, h in enumerate(hdr)}
def g(r, k):
    return r[c[k]].replace(",", "") if k in c else "-"
tot = 0.0
print(f"{'kernel':28s} {'us':>8s} {'dramGB/s':>9s} {'dram%':>6s} {'lts%':>6s} {'l2hit':>6s} {'sm%':>6s} {'tens%':>6s} {'grid':>7s} {'blk':>5s} {'warps%':>6s} {'regs':>4s}")
for r in rows[2:]:
    name = r[c["Kernel Name"]].replace("<unnamed>::", "").replace("(anonymous namespace)::", "").replace("void ", "")
    if pat not in name:
        continue
    short = name.split("(")[0][:28]
    d = float(g(r, "gpu__time_duration.sum")); u = units[c["gpu__time_duration.sum"]]
    d_us = d / 1e3 if u == "ns" else d if u == "us" else d * 1e3
    tot += d_us
    bps = float(g(r, "dram__bytes.sum.per_second")); bu = units[c["dram__bytes.sum.per_second"]]
    gbs = bps * {"Tbyte/s": 1e3, "Gbyte/s": 1.0, "Mbyte/s": 1e-3, "Kbyte/s": 1e-6, "byte/s": 1e-9}.get(bu, 1.0)
    print(f"{short:28s} {d_us:8.1f} {gbs:9.0f} {float(g(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):6.1f} "
          f"{float(g(r,'lts__throughput.avg.pct_of_peak_sustained_elapsed')):6.1f} {float(g(r,'lts__t_sector_hit_rate.pct')):6.1f} "
          f"{float(g(r,'sm__throughput.avg.pct_of_peak_sustained_elapsed')):6.1f} {float(g(r,'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')):6.1f} "
          f"{g(r,'launch__grid_size'):>7s} {g(r,'launch__block_size'):>5s} {float(g(r,'sm__warps_active.avg.pct_of_peak_sustained_active')):6.1f} {g(r,'launch__registers_per_thread'):>4s}")
print(f"total {tot:.1f} us")
