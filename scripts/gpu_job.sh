# usage: bash scripts/gpu_job.sh <tag> [chains]
tag=$1; nb=${2:-2368}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sector_kernels_gpu.py tests/test_cfg2_at_size_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 900 python bench.py --workload cfg2 --chains $nb --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_cfg2_sector_nb${nb}_${tag}.json 2> gpurun_out/r02_cfg2_sector_nb${nb}_${tag}.err
tail -c 600 gpurun_out/r02_cfg2_sector_nb${nb}_${tag}.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_cfg2_sector_nb${nb}_${tag}.json"))
print(d["value"], d["e2e"], d["gpu_launches"], d["energy_per_site"], d["hbm_peak_allocated_gb"], d.get("parity_check"))
print({k:(v["launches"], round(v["ms"],1)) for k,v in d["kernel_breakdown"].items()})
r=d["roofline"]["classes"]
print({k:{a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ("seconds","gbs","frac_of_hbm_peak","executed_over_algorithmic","algorithmic_tflops")} for k,v in r.items()})
for r in d["top_shapes"][:40]: print(r["kernel"], r["mnk"], r["chains"], r["launches"], round(r["ms"],1))
PY
