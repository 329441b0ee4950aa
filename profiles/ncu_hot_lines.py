"""Hottest SASS instructions of an ncu report by warp-stall samples: python profiles/ncu_hot_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for k, r in enumerate(rows[2:]):
    if len(r) != len(hdr): continue
    data.append((int(r[ci["# Samples"]] or 0), k, r))
tot = sum(d[0] for d in data)
print("total samples", tot)
for s, k, r in sorted(data, key=lambda x: -x[0])[:top]:
    why = sorted(((int(r[ci[h]] or 0), h) for h in stalls), reverse=True)[:2]
    print("%6d %5.1f%%  #%d  %-70s %s" % (s, 100.0 * s / tot, k, r[ci["Source"]].strip()[:70], why))
