mkdir -p gpurun_out
nb=${1:-16}
timeout 2000 python bench.py --workload cfg3 --chains $nb --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_cfg3_nb$nb.json 2> gpurun_out/r02_cfg3_nb$nb.err
tail -c 800 gpurun_out/r02_cfg3_nb$nb.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_cfg3_nb$nb.json"))
print("cfg3", d["value"], d["e2e"]["value"], d["gpu_launches"], d["energy_per_site"], d["hbm_peak_allocated_gb"], d["ms_per_step"])
print({k:(v["launches"], round(v["ms"],1)) for k,v in d["kernel_breakdown"].items()})
for r in d["top_shapes"][:14]: print(r["kernel"], r["mnk"], r["chains"], r["launches"], round(r["ms"],1))
print(json.dumps(d["roofline"]["classes"], indent=0)[:1500])
PY
