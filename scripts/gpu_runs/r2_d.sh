# call D: TMA-fed software-pipelined sector GEMM (rt_gemm_warp2_kernel) and the four-in-flight regrouping loop: parity, then A/B on cfg2
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_sector_kernels_gpu.py tests/test_cfg2_at_size_gpu.py tests/test_sector_fermi_gpu.py -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/r2d_tests.txt; cat gpurun_out/r2d_tests.txt
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_cfg2_$tag.json 2> gpurun_out/r2d_cfg2_$tag.err; tail -c 300 gpurun_out/r2d_cfg2_$tag.err; }
run gen2 X=0
run gen1 TNSP_RT_GEMM_GEN=1
timeout 900 python bench.py --workload cfg2 --chains 592 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_cfg2_nb592.json 2> gpurun_out/r2d_cfg2_nb592.err
python - <<'PY'
import json
for f in ("gpurun_out/r2d_cfg2_gen2.json", "gpurun_out/r2d_cfg2_gen1.json", "gpurun_out/r2d_cfg2_nb592.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["gpu_launches"], d["ms_per_step"], d["hbm_peak_allocated_gb"], d.get("parity_check", {}) and d["parity_check"].get("ok"))
        print({k: (v["launches"], round(v["ms"], 1)) for k, v in d["kernel_breakdown"].items()})
        print({k: round(v.get("frac_of_hbm_peak", 0), 4) for k, v in d["roofline"]["classes"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
