# call J: ncu evidence of one STEADY-STATE step (launch list restricted by --launch-skip / --launch-count: ncu costs ~18 ms per profiled
# launch, the whole run has ~170 k), ncu --set full of the factorisation work kernels, then the fermionic configurations (call I)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 112000 --launch-count 14500 --csv --log-file gpurun_out/r2j_launches_cfg2_nb2368_step.csv \
  python bench.py --workload cfg2 --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2j_ncu_bench.json 2> gpurun_out/r2j_ncu_bench.err
tail -c 200 gpurun_out/r2j_ncu_bench.err
python scripts/launch_list_summary.py gpurun_out/r2j_launches_cfg2_nb2368_step.csv > gpurun_out/r2j_launches_cfg2_nb2368_step.summary.txt
head -24 gpurun_out/r2j_launches_cfg2_nb2368_step.summary.txt
gzip -f gpurun_out/r2j_launches_cfg2_nb2368_step.csv
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:rt_(qr_work|svd_work)' -s 7000 -c 24 -o /tmp/r2j_factor \
   python bench.py --workload cfg2 --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > /tmp/r2j_factor.json 2> /tmp/r2j_factor.err
tail -c 200 /tmp/r2j_factor.err
ncu -i /tmp/r2j_factor.ncu-rep --page raw --csv > gpurun_out/r2j_ncu_factor.raw.csv 2>/dev/null
python scripts/ncu_raw_summary.py gpurun_out/r2j_ncu_factor.raw.csv > gpurun_out/r2j_ncu_factor.summary.txt
grep -E "^==|time_duration|fp64_cycles|issue_active|warps_active" gpurun_out/r2j_ncu_factor.summary.txt | head -60
bash scripts/gpu_runs/r2_i.sh
