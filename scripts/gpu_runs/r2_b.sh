# call B: warp-level rt_sort parity, host-side profile (cProfile) of a host-bound cfg2 run, ncu --set full of the sector kernels (reports
# turned into csv on the box: gpurun_out/ carries at most 64 MiB back), cfg2 at full size
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_sector_kernels_gpu.py tests/test_cfg2_at_size_gpu.py -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/r2b_tests.txt; cat gpurun_out/r2b_tests.txt
timeout 900 python -m cProfile -o gpurun_out/r2b_cfg2_nb148.pstats bench.py --workload cfg2 --chains 148 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_cfg2_nb148.json 2> gpurun_out/r2b_cfg2_nb148.err
tail -c 300 gpurun_out/r2b_cfg2_nb148.err
timeout 600 python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_cfg2.json 2> gpurun_out/r2b_cfg2.err
tail -c 300 gpurun_out/r2b_cfg2.err
ncu_one() {  # tag, kernel regex, skip, count
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -o /tmp/$1 \
     python bench.py --workload cfg2 --chains 1184 --steps 1 --warmup 3 --no-cpu-baseline > /tmp/$1.json 2> /tmp/$1.err
  tail -c 200 /tmp/$1.err
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/$1.source.csv 2>/dev/null
  gzip -f gpurun_out/$1.source.csv
  ls -la /tmp/$1.ncu-rep gpurun_out/$1.*
}
ncu_one r2b_ncu_stream 'rt_(gemm_warp|repack_tile|sort)' 24000 36
ncu_one r2b_ncu_factor 'rt_(qr_work|svd_work)' 700 8
du -sh gpurun_out
python - <<'PY'
import json
for f in ("gpurun_out/r2b_cfg2_nb148.json", "gpurun_out/r2b_cfg2.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["gpu_launches"], d["ms_per_step"], d["hbm_peak_allocated_gb"], d.get("parity_check", {}) and d["parity_check"].get("ok"))
        print({k: (v["launches"], round(v["ms"], 1)) for k, v in d["kernel_breakdown"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
