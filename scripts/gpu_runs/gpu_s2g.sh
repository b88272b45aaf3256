set -x
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/s2g_ktests.txt
(cd scripts && timeout 300 python mb_sector.py 296 > ../gpurun_out/s2g_mb_sector.txt 2>&1)
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/s2g_tests.txt
timeout 600 python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s2g_bench_cfg2.json 2> gpurun_out/s2g_bench_cfg2.err
