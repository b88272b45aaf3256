# call P: cfg2 at 2960 chains (20 x 148): does a larger lock-step batch still pay now that the step is GPU-bound?
mkdir -p gpurun_out
timeout 100 python bench.py --workload cfg2 --chains 2960 --steps 2 --warmup 2 --no-cpu-baseline --no-secondary > gpurun_out/r2p_cfg2_nb2960.json 2> gpurun_out/r2p_cfg2_nb2960.err
tail -c 300 gpurun_out/r2p_cfg2_nb2960.err
python -c "
import json; d=json.loads(open('gpurun_out/r2p_cfg2_nb2960.json').read().strip().splitlines()[-1]); print(d['config']['chains_per_gpu'], d['value'], d['e2e']['value'], d['ms_per_step'], d['hbm_peak_allocated_gb'], d['parity_check']['ok'])"
