# call O: gauge fixing (unconditional noise draw) and the drivers on the final commit
mkdir -p gpurun_out
( timeout 95 python -m pytest tests/test_next_rows_gpu.py tests/test_vmc_gpu.py -m gpu -q -k "gauge or expansion or lowers or direct_sampling_driver" 2>&1 | tail -3 ) > gpurun_out/r2o_tests.txt; cat gpurun_out/r2o_tests.txt
