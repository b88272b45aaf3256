set -x
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --chains 592 > gpurun_out/s2l_bench_cfg2_2gpu.json 2> gpurun_out/s2l_bench_cfg2_2gpu.err
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 3 > gpurun_out/s2l_ref_2gpu.json 2> gpurun_out/s2l_ref_2gpu.err
