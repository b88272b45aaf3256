set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r01d_tests.txt
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r01d_smoke.txt 2>&1
python bench.py > gpurun_out/r01d_bench_cfg2.json 2> gpurun_out/r01d_bench_cfg2.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r01d_ref_cfg2.json 2> gpurun_out/r01d_ref_cfg2.err
python bench.py --workload cfg1 --steps 4 --warmup 3 > gpurun_out/r01d_bench_cfg1.json 2> gpurun_out/r01d_bench_cfg1.err
# launch list of one cfg2 step (skip set-up + warm-up launches; shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 9000 --csv --log-file gpurun_out/r01d_launches_cfg2.csv python bench.py --workload cfg2 --chains 592 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r01d_ncu_launch.log 2>&1
cd scripts
ncu --set full --clock-control none --import-source on -k regex:gemm_stream -s 2 -c 1 -o ../gpurun_out/r01d_prof_gemm_1296x36x36 python mb_one.py gemm 1296 36 36 592 > ../gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:svd_sector -s 2 -c 1 -o ../gpurun_out/r01d_prof_svd_216x216 python mb_one.py svd 216 216 6 592 > ../gpurun_out/ncu_svd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:qr_sector -s 2 -c 1 -o ../gpurun_out/r01d_prof_lq_216x1296 python mb_one.py lq 216 1296 7 592 > ../gpurun_out/ncu_lq.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pack_generic -s 40 -c 1 -o ../gpurun_out/r01d_prof_pack python ../bench.py --workload cfg2 --chains 148 --steps 1 --warmup 0 --no-cpu-baseline > ../gpurun_out/ncu_pack.log 2>&1
