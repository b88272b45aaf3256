# call H (two GPUs): NCCL parity on hardware + a short cfg2 run on 2 GPUs
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_nccl_gpu.py -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/r2h_nccl_test.txt; cat gpurun_out/r2h_nccl_test.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu-baseline --no-secondary > gpurun_out/r2h_cfg2_2gpu.json 2> gpurun_out/r2h_cfg2_2gpu.err
tail -c 300 gpurun_out/r2h_cfg2_2gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/r2h_cfg2_2gpu.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['e2e'], d['ms_per_step'])"
