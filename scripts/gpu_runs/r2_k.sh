# call K (last single-GPU call of the round): verification of the final state -- whole -m gpu suite, smoke, cfg2, cfg4
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/r2k_gputests.txt; cat gpurun_out/r2k_gputests.txt
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) > gpurun_out/r2k_smoke.txt; cat gpurun_out/r2k_smoke.txt
timeout 400 python bench.py --workload cfg2 --steps 4 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2k_cfg2.json 2> gpurun_out/r2k_cfg2.err
tail -c 300 gpurun_out/r2k_cfg2.err
timeout 500 python bench.py --workload cfg4 --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/r2k_cfg4_nb37.json 2> gpurun_out/r2k_cfg4_nb37.err
tail -c 300 gpurun_out/r2k_cfg4_nb37.err
python - <<'PY'
import json
for f in ("gpurun_out/r2k_cfg2.json", "gpurun_out/r2k_cfg4_nb37.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); kb = d["kernel_breakdown"]
        print(f, round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["ms_per_step"]), d["gpu_launches"], round(d["hbm_peak_allocated_gb"], 1), d["energy_per_site"], d.get("parity_check") and d["parity_check"]["ok"],
              {k: round(v["ms"]) for k, v in kb.items() if v["ms"] > 50})
    except Exception as e:
        print(f, "failed", e)
PY
