# ncu launch list of the bench command (per-launch gpu__time_duration; cold-cache, serialised: shares matter, not absolutes)
tag=$1; nb=${2:-2368}
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_cfg2_nb${nb}_${tag}.csv \
  python bench.py --workload cfg2 --chains $nb --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_bench_${tag}.json 2> gpurun_out/r02_ncu_bench_${tag}.err
tail -c 300 gpurun_out/r02_ncu_bench_${tag}.err
python - <<PY
import csv, collections, re
rows = []
with open("gpurun_out/r02_launches_cfg2_nb${nb}_${tag}.csv") as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
tot = collections.defaultdict(lambda: [0, 0.0])
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "nsecond": 1, "ms": 1e6, "msecond": 1e6}.get(unit, 1)
    tot[name][0] += 1
    tot[name][1] += ns
allns = sum(v[1] for v in tot.values())
with open("gpurun_out/r02_launches_cfg2_nb${nb}_${tag}.summary.txt", "w") as out:
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        line = f"{k:70s} launches {v[0]:7d}  total {v[1]/1e6:10.2f} ms  share {v[1]/allns:6.3f}  avg {v[1]/v[0]/1e3:9.1f} us"
        print(line); out.write(line + "\n")
PY
gzip -f gpurun_out/r02_launches_cfg2_nb${nb}_${tag}.csv
