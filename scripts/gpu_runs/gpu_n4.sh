set -x
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 2 --warmup 3 --chains 592 > gpurun_out/n4_bench_cfg2.json 2> gpurun_out/n4_bench_cfg2.err
