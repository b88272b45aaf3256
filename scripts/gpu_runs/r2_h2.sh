# call H2 (two GPUs): the bench under torchrun after the parity check was moved to every rank (H: the rank-0-only check unbalanced the collectives)
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu-baseline --no-secondary > gpurun_out/r2h_cfg2_2gpu.json 2> gpurun_out/r2h_cfg2_2gpu.err
tail -c 400 gpurun_out/r2h_cfg2_2gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2h_cfg2_2gpu.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['e2e'], d['ms_per_step'], d['parity_check']['ok'], d['energy_per_site'])"
