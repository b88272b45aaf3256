set -x
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/fin_tests.txt
timeout 120 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' > gpurun_out/fin_smoke.txt 2>&1
timeout 900 python bench.py > gpurun_out/fin_bench_cfg2.json 2> gpurun_out/fin_bench_cfg2.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/fin_ref_cfg2.json 2> gpurun_out/fin_ref_cfg2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 22000 -c 7000 --csv --log-file gpurun_out/fin_launches_cfg2.csv python bench.py --workload cfg2 --chains 592 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/fin_ncu_launch.log 2>&1
cd scripts
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_rowstream -s 26 -c 1 -o ../gpurun_out/fin_prof_rowstream_sparse python mb_gemm_sparse.py 592 > ../gpurun_out/fin_ncu_rowstream.log 2>&1
