set -x
timeout 300 ncu --set full --clock-control none --import-source on -k regex:svd_work_kernel -s 60 -c 1 -o gpurun_out/real_prof_svd_work python bench.py --workload cfg2 --chains 592 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/real_ncu_svd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qr_work_kernel -s 60 -c 1 -o gpurun_out/real_prof_qr_work python bench.py --workload cfg2 --chains 592 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/real_ncu_qr.log 2>&1
