set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/s2m_tests.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s2m_bench_cfg2.json 2> gpurun_out/s2m_bench_cfg2.err
timeout 600 python bench.py --chains 1776 --steps 4 --no-cpu-baseline > gpurun_out/s2m_bench_cfg2_1776.json 2> gpurun_out/s2m_bench_cfg2_1776.err
(cd scripts && timeout 400 python mb_cfg5.py 64 2048 128 512 > ../gpurun_out/s2m_mb_cfg5.txt 2>&1)
