"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel."""
import collections, csv, re, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    n = row["Kernel Name"].replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
    n = re.sub(r"\(.*", "", re.sub(r"<.*", "", n))
    n = n.split("::")[-1] if "eitb" in n or "GLOBAL__N" in n else n
    agg[n][0] += 1
    agg[n][1] += float(row["Metric Value"].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot/1e3:.1f} us of kernel time (cold-cache, serialised)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{v[1]/1e3:10.1f} us {v[0]:5d} launches {100*v[1]/tot:5.1f}%  {k[:100]}")
