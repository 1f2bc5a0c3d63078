"""Hot SASS instructions of an `ncu --page source --csv` export (SASS view): python profiles/sass_hot.py file.csv [n] [launch index]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
want = int(sys.argv[3]) if len(sys.argv) > 3 else 0
launches = []; cur = None; hdr = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; launches.append(cur); continue
    if r and r[0] == "Address":
        hdr = r; continue
    if cur is not None and hdr and len(r) >= len(hdr) - 2 and r[0].startswith("0x"):
        cur["rows"].append(r)
L = launches[want]
h = {k: i for i, k in enumerate(hdr)}
tot_s = sum(int(r[h["# Samples"]] or 0) for r in L["rows"]) or 1
tot_i = sum(int(r[h["Instructions Executed"]] or 0) for r in L["rows"]) or 1
print(L["name"][:100], "| launches in file:", len(launches), "| samples", tot_s, "warp instr", tot_i)
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
agg = collections.Counter()
for r in L["rows"]:
    for k in stalls:
        agg[k] += int(r[h[k]] or 0)
print("stall mix:", ", ".join(f"{k[6:]} {100*v/tot_s:.0f}%" for k, v in agg.most_common(6)))
ops = collections.Counter()
for r in L["rows"]:
    ops[r[h["Source"]].split()[0]] += int(r[h["Instructions Executed"]] or 0)
print("instr mix:", ", ".join(f"{k} {100*v/tot_i:.0f}%" for k, v in ops.most_common(10)))
for r in sorted(L["rows"], key=lambda r: -int(r[h["# Samples"]] or 0))[:topn]:
    st = max(stalls, key=lambda k: int(r[h[k]] or 0))
    print(f"{100*int(r[h['# Samples']] or 0)/tot_s:5.1f}% smp {100*int(r[h['Instructions Executed']] or 0)/tot_i:5.1f}% inst  {r[h['Source']].strip()[:70]:70s} {st[6:]}")
