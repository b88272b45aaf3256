# round 2, session 2, call A: full -m gpu suite, cfg2 bench, ncu --set full of the sector-engine kernels, cfg3 at 74 chains
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/r2a_gputests.txt
cat gpurun_out/r2a_gputests.txt
timeout 600 python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_cfg2.json 2> gpurun_out/r2a_cfg2.err
tail -c 300 gpurun_out/r2a_cfg2.err
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:rt_(gemm_warp|repack_tile|sort|qr_work|svd_work)' -s 20000 -c 120 \
   -o gpurun_out/r2a_sector_full python bench.py --workload cfg2 --chains 592 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu_bench.json 2> gpurun_out/r2a_ncu_bench.err
tail -c 300 gpurun_out/r2a_ncu_bench.err
ls -la gpurun_out/r2a_sector_full.ncu-rep
timeout 1500 python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_cfg3_nb74.json 2> gpurun_out/r2a_cfg3_nb74.err
tail -c 300 gpurun_out/r2a_cfg3_nb74.err
python - <<'PY'
import json
for f in ("gpurun_out/r2a_cfg2.json", "gpurun_out/r2a_cfg3_nb74.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["gpu_launches"], d["ms_per_step"], d["hbm_peak_allocated_gb"], d.get("parity_check", {}) and d["parity_check"].get("ok"))
        print({k: (v["launches"], round(v["ms"], 1)) for k, v in d["kernel_breakdown"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
