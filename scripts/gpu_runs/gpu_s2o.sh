set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_cfg5.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s2o_ktests.txt
(cd scripts && timeout 300 python mb_cfg5.py 64 2048 128 512 > ../gpurun_out/s2o_mb_cfg5.txt 2>&1)
(cd scripts && timeout 300 python mb_sector.py 296 > ../gpurun_out/s2o_mb_sector.txt 2>&1)
