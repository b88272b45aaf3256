# call I: fermionic configurations after the host-path work: cfg3 at 74 and 296 chains, cfg4 at 37; fermionic parity suite
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_sector_fermi_gpu.py tests/test_sector_kernels_gpu.py -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/r2i_tests.txt; cat gpurun_out/r2i_tests.txt
run() { tag=$1; shift; timeout 1500 python bench.py "$@" --no-cpu-baseline > gpurun_out/r2i_$tag.json 2> gpurun_out/r2i_$tag.err; tail -c 400 gpurun_out/r2i_$tag.err; }
run cfg3_nb74 --workload cfg3 --steps 2 --warmup 3
run cfg3_nb296 --workload cfg3 --chains 296 --steps 2 --warmup 3
run cfg4_nb37 --workload cfg4 --steps 1 --warmup 3
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2i_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        kb = d["kernel_breakdown"]
        print(f, round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["ms_per_step"]), d["gpu_launches"], round(d["hbm_peak_allocated_gb"], 1), d["energy_per_site"],
              {k: round(v["ms"]) for k, v in kb.items() if v["ms"] > 50})
    except Exception as e:
        print(f, "failed", e)
PY
