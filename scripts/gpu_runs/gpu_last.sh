set -x
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/last_tests.txt
timeout 300 python bench.py --steps 3 --no-cpu-baseline > gpurun_out/last_bench_cfg2.json 2> gpurun_out/last_bench_cfg2.err
