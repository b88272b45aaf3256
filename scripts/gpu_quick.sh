set -x
python -m pytest tests/test_kernels_gpu.py tests/test_dense_gpu.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/quick_tests.txt
python scripts/dbg_struct_gpu.py 6 36 > gpurun_out/dbg_struct.txt 2>&1
timeout 1200 python bench.py --workload cfg2 --chains 148 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/quick_bench_cfg2.json 2> gpurun_out/quick_bench_cfg2.err
