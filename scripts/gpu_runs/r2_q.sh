# call Q: the table-cache purge on the B200 (multi-sweep sector-engine tests)
mkdir -p gpurun_out
( timeout 55 python -m pytest tests/test_sector_kernels_gpu.py -m gpu -q -k "lockstep or trajectory or fixture" 2>&1 | tail -2 ) > gpurun_out/r2q_tests.txt; cat gpurun_out/r2q_tests.txt
