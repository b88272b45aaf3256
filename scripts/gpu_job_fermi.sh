mkdir -p gpurun_out
for w in cfg3s cfg4s; do
timeout 600 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_${w}.json 2> gpurun_out/r02_${w}.err
tail -c 500 gpurun_out/r02_${w}.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_${w}.json"))
print("$w", d["value"], d["e2e"]["value"], d["gpu_launches"], d["energy_per_site"], d["hbm_peak_allocated_gb"])
print({k:(v["launches"], round(v["ms"],1)) for k,v in d["kernel_breakdown"].items()})
PY
done
timeout 1200 python bench.py --workload cfg3 --chains 16 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_cfg3_nb16.json 2> gpurun_out/r02_cfg3_nb16.err
tail -c 800 gpurun_out/r02_cfg3_nb16.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_cfg3_nb16.json"))
print("cfg3", d["value"], d["e2e"]["value"], d["gpu_launches"], d["energy_per_site"], d["hbm_peak_allocated_gb"], d["ms_per_step"])
print({k:(v["launches"], round(v["ms"],1)) for k,v in d["kernel_breakdown"].items()})
for r in d["top_shapes"][:12]: print(r["kernel"], r["mnk"], r["chains"], r["launches"], round(r["ms"],1))
PY
