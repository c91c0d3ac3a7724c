"""ncu launch list (csv of gpu__time_duration.sum) -> markdown table of the LAST `n` launches (= the timed step).
usage: python tools/launch_table.py gpurun_out/launches.csv 458 > profiles/rNN_launches_....md"""
import csv
import re
import sys

path, n = sys.argv[1], int(sys.argv[2])
rows = []
with open(path) as fh:
    rd = csv.reader(l for l in fh if l.startswith('"'))
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        v = float(r[iv].replace(",", ""))
        rows.append((r[ik], v / 1e3 if r[iu] == "ns" else v))
rows = rows[-n:]
agg = {}
for k, us in rows:
    k = re.sub(r"\(.*$", "", k).strip()
    k = re.sub(r"^void ", "", k)
    a = agg.setdefault(k[:70], [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print("last %d launches = the timed step; cold-cache, serialised, burst clocks: compare SHARES. total %.1f us\n" % (n, tot))
print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.1f | %.1f%% | %.1f |" % (k, c, us, 100 * us / tot, us / c))
