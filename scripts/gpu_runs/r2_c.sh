# call C: host-path optimisations (backend fast paths, cached ctypes forms, contract plans) + launch-shape experiments of the sector
# factorisations at cfg2's hot shapes (scripts/mb_rt_factor.py)
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_sector_kernels_gpu.py tests/test_cfg2_at_size_gpu.py tests/test_sector_fermi_gpu.py -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/r2c_tests.txt; cat gpurun_out/r2c_tests.txt
mb() { echo "== $*"; env "$@" timeout 300 python scripts/mb_rt_factor.py 2368 5 2>&1 | tail -3; }
( mb X=0
  mb TNSP_RT_SVD_T2=128
  mb TNSP_RT_SVD_T2=128 TNSP_RT_MID_DOUBLES=5400
  mb TNSP_RT_SVD_T2=256 TNSP_RT_MID_DOUBLES=5400
  mb TNSP_RT_SVD_T2=192 TNSP_RT_MID_DOUBLES=5400
  mb TNSP_RT_SVD_T3=64 TNSP_RT_QR_T3=64
  mb TNSP_RT_SVD_T3=256 TNSP_RT_QR_T3=256
  mb TNSP_RT_QR_T1=128 TNSP_RT_QR_T2=128
  mb TNSP_RT_QR_T1=512 TNSP_RT_QR_T2=512 ) > gpurun_out/r2c_mb_factor.txt 2>&1
cat gpurun_out/r2c_mb_factor.txt
timeout 900 python bench.py --workload cfg2 --chains 148 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_cfg2_nb148.json 2> gpurun_out/r2c_cfg2_nb148.err
tail -c 300 gpurun_out/r2c_cfg2_nb148.err
timeout 600 python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_cfg2.json 2> gpurun_out/r2c_cfg2.err
tail -c 300 gpurun_out/r2c_cfg2.err
python - <<'PY'
import json
for f in ("gpurun_out/r2c_cfg2_nb148.json", "gpurun_out/r2c_cfg2.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["gpu_launches"], d["ms_per_step"], d["hbm_peak_allocated_gb"], d.get("parity_check", {}) and d["parity_check"].get("ok"))
        print({k: (v["launches"], round(v["ms"], 1)) for k, v in d["kernel_breakdown"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
