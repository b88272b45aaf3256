# call F: big-class QR thread counts, the whole -m gpu suite + smoke, cfg2 with the secondary workloads, cfg3 / cfg4 at full size
mkdir -p gpurun_out
mb() { echo "== $*"; env "$@" timeout 300 python scripts/mb_rt_factor.py 2368 5 2>&1 | tail -3; }
( mb X=0
  mb TNSP_RT_QR_T0=512
  mb TNSP_RT_QR_T0=256 ) > gpurun_out/r2f_mb_factor.txt 2>&1
cat gpurun_out/r2f_mb_factor.txt
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/r2f_gputests.txt; cat gpurun_out/r2f_gputests.txt
( timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/r2f_smoke.txt; cat gpurun_out/r2f_smoke.txt
( time timeout 1500 python bench.py --steps 6 --warmup 3 > gpurun_out/r2f_cfg2_default.json 2> gpurun_out/r2f_cfg2_default.err ) 2>&1 | tail -3
tail -c 300 gpurun_out/r2f_cfg2_default.err
timeout 1500 python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_cfg3_nb74.json 2> gpurun_out/r2f_cfg3_nb74.err
tail -c 300 gpurun_out/r2f_cfg3_nb74.err
timeout 1500 python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_cfg4_nb37.json 2> gpurun_out/r2f_cfg4_nb37.err
tail -c 300 gpurun_out/r2f_cfg4_nb37.err
python - <<'PY'
import json, glob
for f in ("gpurun_out/r2f_cfg2_default.json", "gpurun_out/r2f_cfg3_nb74.json", "gpurun_out/r2f_cfg4_nb37.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        kb = d["kernel_breakdown"]
        print(f, round(d["value"], 2), round(d["e2e"]["value"], 2), round(d["ms_per_step"]), d["gpu_launches"], d["hbm_peak_allocated_gb"], d.get("parity_check") and d["parity_check"].get("ok"),
              {k: round(v["ms"]) for k, v in kb.items() if v["ms"] > 20})
        print("   cpu_baseline", d.get("cpu_baseline") and d["cpu_baseline"]["value"], "secondary", json.dumps(d.get("secondary"))[:900])
    except Exception as e:
        print(f, "failed", e)
PY
