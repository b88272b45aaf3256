"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (no GPU needed).
    python scripts/launch_summary.py <launches.csv> "<header line>" > profiles/<name>.summary.txt"""
import csv, re, sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    if not r[0].isdigit():
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
    ns = float(r[vi].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[ui], 1.0)
    tot[name] += ns
    cnt[name] += 1
total = sum(tot.values())
print(sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
print("%-60s %8s %10s %9s  %s" % ("kernel", "launches", "total ms", "avg us", "share"))
for name in sorted(tot, key=lambda k: -tot[k]):
    print("%-60s %8d %10.2f %9.1f  %.3f" % (name[:60], cnt[name], tot[name] / 1e6, tot[name] / cnt[name] / 1e3, tot[name] / total))
print("total %.1f ms of kernel time in the window (%d launches)" % (total / 1e6, sum(cnt.values())))
