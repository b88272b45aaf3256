set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_dense_gpu.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/s2u_ktests.txt
(cd scripts && timeout 300 python mb_sector.py 296 > ../gpurun_out/s2u_mb_sector.txt 2>&1)
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/s2u_bench_cfg2.json 2> gpurun_out/s2u_bench_cfg2.err
