set -x
cd scripts
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ../gpurun_out/s2j_launches_lq.csv python mb_sector_one.py lq 216 1296 296 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ../gpurun_out/s2j_launches_svd.csv python mb_sector_one.py svd 216 216 296 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qr_work -s 4 -c 1 -o ../gpurun_out/s2j_prof_qr_work_big python mb_sector_one.py lq 216 1296 296 > ../gpurun_out/s2j_ncu_lq.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sector_discover -s 2 -c 1 -o ../gpurun_out/s2j_prof_discover python mb_sector_one.py lq 216 1296 296 > ../gpurun_out/s2j_ncu_disc.log 2>&1
