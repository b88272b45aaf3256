set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_dense_gpu.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/s2w_ktests.txt
timeout 400 python bench.py --workload heis6 --chains 592 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s2w_bench_heis6.json 2> gpurun_out/s2w_bench_heis6.err
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/s2w_bench_cfg2.json 2> gpurun_out/s2w_bench_cfg2.err
