set -x
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm" 2>&1 | tail -5 > gpurun_out/s2t_ktests.txt
(cd scripts && timeout 300 python mb_gemm_sparse.py 592 > ../gpurun_out/s2t_mb_gemm_sparse.txt 2>&1)
(cd scripts && timeout 300 python mb_gemm_sparse.py 2368 > ../gpurun_out/s2t_mb_gemm_sparse_2368.txt 2>&1)
