set -x
python -m pytest tests/test_kernels_gpu.py tests/test_dense_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/quick_tests.txt
(cd scripts && python mb_factor.py 148 > ../gpurun_out/mb_148.txt 2>&1)
timeout 900 python bench.py --workload cfg2 --chains 592 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/quick_bench_cfg2_592.json 2> gpurun_out/quick_bench_cfg2_592.err
