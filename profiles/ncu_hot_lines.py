"""Hottest SASS instructions of an ncu report by warp-stall samples, plus samples between synchronisation markers:
python profiles/ncu_hot_lines.py report.ncu-rep [top] [kernel-substring]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40; want = sys.argv[3] if len(sys.argv) > 3 else ""
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# split into kernels
kernels = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; kernels.append(cur)
    elif cur is not None: cur["rows"].append(r)
for k in kernels:
    if want not in k["name"]: continue
    hdr = k["rows"][0]; ci = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = [(int(r[ci["# Samples"]] or 0), i, r) for i, r in enumerate(k["rows"][1:]) if len(r) == len(hdr)]
    tot = sum(d[0] for d in data) or 1
    print("==", k["name"][:80], "total samples", tot)
    for s, i, r in sorted(data, key=lambda x: -x[0])[:top]:
        why = sorted(((int(r[ci[h]] or 0), h) for h in stalls), reverse=True)[:2]
        print("%6d %5.1f%%  #%d  %-64s %s" % (s, 100.0 * s / tot, i, r[ci["Source"]].strip()[:64], why))
    marks = [i for s, i, r in data if any(m in r[ci["Source"]] for m in ("UTCHMMA", "UTCBAR", "SYNCS", "BAR.SYNC", "LDTM", "UTCATOM", "DEPBAR"))]
    prev = 0
    print("-- samples between markers")
    for m in marks + [len(data)]:
        s = sum(d[0] for d in data[prev:m]); n = m - prev
        if s > tot * 0.004 or n > 20:
            print("insts %4d..%4d (%4d)  %5d %5.1f%%   next: %s" % (prev, m, n, s, 100 * s / tot, data[m][2][ci["Source"]].strip()[:50] if m < len(data) else "END"))
        if m < len(data) and data[m][0] > tot * 0.004:
            print("   marker #%4d  %5d %5.1f%%  %s" % (m, data[m][0], 100 * data[m][0] / tot, data[m][2][ci["Source"]].strip()[:60]))
        prev = m + 1
