"""per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`):
    python scripts/launch_list_summary.py <launches.csv> [chains]
With `chains`, only launches whose grid has that many CTAs in y (or x) are counted in a second table: the steady-state batch, without
the calibration / parity / normalisation launches of smaller batches."""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    chains = sys.argv[2] if len(sys.argv) > 2 else None
    lines = [l for l in open(path) if not l.startswith("==")]
    tot = collections.defaultdict(lambda: [0, 0.0])
    big = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "nsecond": 1, "ms": 1e6, "msecond": 1e6}.get(row["Metric Unit"], 1)
        tot[name][0] += 1
        tot[name][1] += ns
        if chains and re.search(r"[ (]%s[,)]" % chains, row.get("Grid Size", "")):
            big[name][0] += 1
            big[name][1] += ns
    for title, table in (("all launches", tot), (f"launches with {chains} CTAs in one grid dimension (the lock-step batch itself)", big)):
        if not table:
            continue
        allns = sum(v[1] for v in table.values())
        print(f"== {title}: {sum(v[0] for v in table.values())} launches, {allns / 1e6:.1f} ms")
        for k, v in sorted(table.items(), key=lambda kv: -kv[1][1])[:40]:
            print(f"{k[:72]:72s} launches {v[0]:7d}  total {v[1] / 1e6:10.2f} ms  share {v[1] / allns:6.3f}  avg {v[1] / v[0] / 1e3:9.1f} us")


if __name__ == "__main__":
    main()
