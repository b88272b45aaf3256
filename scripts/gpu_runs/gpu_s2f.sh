set -x
(cd scripts && timeout 600 python mb_gemm.py 592 > ../gpurun_out/s2f_mb_gemm.txt 2>&1)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/s2f_bench_cfg2_2gpu.json 2> gpurun_out/s2f_bench_cfg2_2gpu.err
timeout 600 python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s2f_bench_cfg2.json 2> gpurun_out/s2f_bench_cfg2.err
