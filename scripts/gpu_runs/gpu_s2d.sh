set -x
python -m pytest tests/test_kernels_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/s2d_ktests.txt
(cd scripts && timeout 600 python mb_gemm.py 592 > ../gpurun_out/s2d_mb_gemm.txt 2>&1)
(cd scripts && timeout 600 python mb_sector.py 296 > ../gpurun_out/s2d_mb_sector.txt 2>&1)
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/s2d_tests.txt
timeout 900 python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s2d_bench_cfg2.json 2> gpurun_out/s2d_bench_cfg2.err
cd scripts
ncu --set full --clock-control none --import-source on -k regex:gemm_rowstream -s 4 -c 1 -o ../gpurun_out/s2d_prof_rowstream python mb_gemm.py 592 > ../gpurun_out/s2d_ncu_rowstream.log 2>&1
