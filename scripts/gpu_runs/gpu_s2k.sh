set -x
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s2k_ktests.txt
(cd scripts && timeout 300 python mb_sector.py 296 > ../gpurun_out/s2k_mb_sector.txt 2>&1)
(cd scripts && timeout 300 python mb_gemm.py 592 > ../gpurun_out/s2k_mb_gemm.txt 2>&1)
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/s2k_tests.txt
timeout 120 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' > gpurun_out/s2k_smoke.txt 2>&1
timeout 900 python bench.py > gpurun_out/s2k_bench_cfg2.json 2> gpurun_out/s2k_bench_cfg2.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/s2k_ref_cfg2.json 2> gpurun_out/s2k_ref_cfg2.err
# launch list of one cfg2 step (592 chains; skip set-up + warm-up launches; shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 22000 -c 7000 --csv --log-file gpurun_out/s2k_launches_cfg2.csv python bench.py --workload cfg2 --chains 592 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/s2k_ncu_launch.log 2>&1
cd scripts
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_rowstream -s 70 -c 1 -o ../gpurun_out/s2k_prof_rowstream_1296x216x216 python mb_gemm.py 592 > ../gpurun_out/s2k_ncu_rowstream.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:svd_work -s 3 -c 1 -o ../gpurun_out/s2k_prof_svd_work_216x216 python mb_sector_one.py svd 216 216 296 > ../gpurun_out/s2k_ncu_svd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sector_discover -s 2 -c 1 -o ../gpurun_out/s2k_prof_discover python mb_sector_one.py lq 216 1296 296 > ../gpurun_out/s2k_ncu_disc.log 2>&1
