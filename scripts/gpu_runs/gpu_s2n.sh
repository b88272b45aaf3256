set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_cfg5.py tests/test_tat_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s2n_ktests.txt
(cd scripts && timeout 300 python mb_cfg5.py 64 2048 128 512 > ../gpurun_out/s2n_mb_cfg5.txt 2>&1)
timeout 60 scripts/bin/mb_fp64_pipes > gpurun_out/s2n_fp64_pipes.txt 2>&1
