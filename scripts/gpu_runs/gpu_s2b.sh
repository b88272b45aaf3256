set -x
cd scripts
ncu --set full --clock-control none --import-source on -k regex:svd_work -s 2 -c 1 -o ../gpurun_out/s2b_prof_svd_work_216x216 python mb_sector_one.py svd 216 216 296 > ../gpurun_out/s2b_ncu_svd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:qr_work -s 4 -c 2 -o ../gpurun_out/s2b_prof_qr_work_216x1296 python mb_sector_one.py lq 216 1296 296 > ../gpurun_out/s2b_ncu_lq.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ../gpurun_out/s2b_launches_svd.csv python mb_sector_one.py svd 216 216 296 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ../gpurun_out/s2b_launches_lq.csv python mb_sector_one.py lq 216 1296 296 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_stream -s 2 -c 1 -o ../gpurun_out/s2b_prof_gemm_7776x36x36 python mb_one.py gemm 7776 36 36 592 > ../gpurun_out/s2b_ncu_gemm.log 2>&1
