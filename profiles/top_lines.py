"""Top source lines of an ncu report by stall samples: python profiles/top_lines.py rep.ncu-rep [launch_skip] [n]"""
import csv, subprocess, sys
rep = sys.argv[1]; skip = sys.argv[2] if len(sys.argv) > 2 else "0"; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=cuda,sass", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None; out = []; fn = ""
for r in rows:
    if r and r[0] == "Function Name": fn = r[1]
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) > 8 and r[0].strip().isdigit():
        try:
            inst = int(r[hdr.index("Instructions Executed")] or 0); smp = int(r[hdr.index("# Samples")] or 0)
        except ValueError:
            continue
        out.append((inst, smp, r[0], r[1][:110]))
tot = sum(o[0] for o in out) or 1; ts = sum(o[1] for o in out) or 1
print(fn[:120]); print("warp instructions", tot, "samples", ts)
key = 0 if len(sys.argv) > 4 and sys.argv[4] == "inst" else 1
for o in sorted(out, key=lambda o: -o[key])[:topn]:
    print(f"{100*o[0]/tot:5.1f}% inst {100*o[1]/ts:5.1f}% samples  L{o[2]}: {o[3]}")
