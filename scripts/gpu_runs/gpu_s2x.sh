set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_dense_gpu.py tests/test_cfg5.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/s2x_ktests.txt
timeout 400 python bench.py --workload heis6 --chains 592 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s2x_bench_heis6.json 2> gpurun_out/s2x_bench_heis6.err
(cd scripts && timeout 300 python mb_cfg5.py 64 2048 128 512 > ../gpurun_out/s2x_mb_cfg5.txt 2>&1)
