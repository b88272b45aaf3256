# call E: CTA-size sweeps of the sector GEMM (TNSP_RT_GEMM) and the regrouping kernels (TNSP_RT_REPACK_THREADS) on cfg2 at 2368 chains
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_sector_kernels_gpu.py tests/test_sector_fermi_gpu.py -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/r2e_tests.txt; cat gpurun_out/r2e_tests.txt
( TNSP_RT_GEMM=24 TNSP_RT_REPACK_THREADS=128 timeout 900 python -m pytest tests/test_sector_kernels_gpu.py tests/test_sector_fermi_gpu.py -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/r2e_tests_24.txt; cat gpurun_out/r2e_tests_24.txt
( TNSP_RT_GEMM=12 TNSP_RT_REPACK_THREADS=64 timeout 900 python -m pytest tests/test_sector_kernels_gpu.py -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/r2e_tests_12.txt; cat gpurun_out/r2e_tests_12.txt
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_cfg2_$tag.json 2> gpurun_out/r2e_cfg2_$tag.err; tail -c 200 gpurun_out/r2e_cfg2_$tag.err; }
run g18 TNSP_RT_GEMM=18
run g14 TNSP_RT_GEMM=14
run g12 TNSP_RT_GEMM=12
run g24 TNSP_RT_GEMM=24
run r128 TNSP_RT_REPACK_THREADS=128
run r64 TNSP_RT_REPACK_THREADS=64
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2e_cfg2_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        kb = d["kernel_breakdown"]
        print(f.split("_")[-1], round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"]), d.get("parity_check", {}) and d["parity_check"].get("ok"),
              {k: round(kb[k]["ms"]) for k in ("rt_gemm", "rt_repack", "rt_repack_pair", "rt_sort", "rt_qr_work", "rt_svd_work")})
    except Exception as e:
        print(f, "failed", e)
PY
