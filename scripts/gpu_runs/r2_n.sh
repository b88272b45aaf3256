# call N: the SURVEY 8f rows of tests/test_next_rows_gpu.py on the B200
mkdir -p gpurun_out
( timeout 140 python -m pytest tests/test_next_rows_gpu.py -m gpu -q 2>&1 | tail -6 ) > gpurun_out/r2n_next_rows.txt; cat gpurun_out/r2n_next_rows.txt
