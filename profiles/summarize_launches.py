"""ncu launch list (gpu__time_duration.sum per launch, CSV) -> markdown table: launches, total ms, average, share, for every
kernel (custom and library).  Usage: python profiles/summarize_launches.py launches.csv [first_id [last_id]] > summary.md"""
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
last = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 62
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
per, total, n = {}, 0.0, 0
for r in rows[1:]:
    if len(r) != len(hdr) or r[col["Metric Name"]] != "gpu__time_duration.sum" or not r[col["ID"]].isdigit():
        continue
    if int(r[col["ID"]]) < skip or int(r[col["ID"]]) > last:
        continue
    unit = r[col["Metric Unit"]]
    ms = float(r[col["Metric Value"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
    name = re.sub(r"^void ", "", name)
    if len(name) > 90:
        name = name[:87] + "..."
    e = per.setdefault(name, [0, 0.0])
    e[0] += 1
    e[1] += ms
    total += ms
    n += 1
print("%d launches, %.1f ms of kernel time under ncu (cold-cache, serialised: compare shares, not absolutes)\n" % (n, total))
print("| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|")
custom = 0.0
for name, (cnt, ms) in sorted(per.items(), key=lambda kv: -kv[1][1])[:45]:
    print("| `%s` | %d | %.2f | %.1f | %.1f %% |" % (name, cnt, ms, ms / cnt * 1e3, 100 * ms / total))
for name, (cnt, ms) in per.items():
    if name.startswith(("pnp::", "tc5::")):     # ncu prints the tcgen05 kernel without its outer namespace
        custom += ms
print("\ncustom (`pnp::`) kernels: %.1f ms = %.1f %% of the kernel time" % (custom, 100 * custom / total))
