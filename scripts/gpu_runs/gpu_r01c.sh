set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r01c_tests.txt
python bench.py --workload cfg2s --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r01c_bench_cfg2s.json 2> gpurun_out/r01c_bench_cfg2s.err
timeout 1500 python bench.py --workload cfg2 --chains 148 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01c_bench_cfg2.json 2> gpurun_out/r01c_bench_cfg2.err
