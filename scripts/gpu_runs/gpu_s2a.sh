set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/s2a_smi.txt
python -m pytest tests/test_kernels_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/s2a_ktests.txt
(cd scripts && timeout 600 python mb_sector.py 296 > ../gpurun_out/s2a_mb_sector.txt 2>&1)
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/s2a_tests.txt
timeout 900 python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s2a_bench_cfg2.json 2> gpurun_out/s2a_bench_cfg2.err
