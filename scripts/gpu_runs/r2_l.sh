# call L: cfg3 at full size with the final code: 74 chains (dense-bound buffers) and 296 chains (learnt capacities)
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 400 python bench.py "$@" --no-cpu-baseline > gpurun_out/r2l_$tag.json 2> gpurun_out/r2l_$tag.err; tail -c 300 gpurun_out/r2l_$tag.err; }
run cfg3_nb74 --workload cfg3 --steps 2 --warmup 2
run cfg3_nb296 --workload cfg3 --chains 296 --steps 2 --warmup 2
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2l_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); kb = d["kernel_breakdown"]
        print(f, round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["ms_per_step"]), d["gpu_launches"], round(d["hbm_peak_allocated_gb"], 1), d["energy_per_site"],
              {k: round(v["ms"]) for k, v in kb.items() if v["ms"] > 50})
    except Exception as e:
        print(f, "failed", e)
PY
