set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s2q_ktests.txt
(cd scripts && timeout 300 python mb_gemm.py 592 > ../gpurun_out/s2q_mb_gemm.txt 2>&1)
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/s2q_bench_cfg2.json 2> gpurun_out/s2q_bench_cfg2.err
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/s2q_tests.txt
