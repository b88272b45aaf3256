set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_dense_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s2s_ktests.txt
(cd scripts && timeout 300 python mb_gemm.py 592 > ../gpurun_out/s2s_mb_gemm.txt 2>&1)
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/s2s_bench_cfg2.json 2> gpurun_out/s2s_bench_cfg2.err
