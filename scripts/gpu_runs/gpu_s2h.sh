set -x
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "sector or factor or qr or svd" 2>&1 | tail -5 > gpurun_out/s2h_ktests.txt
(cd scripts && timeout 300 python mb_sector.py 296 > ../gpurun_out/s2h_mb_sector.txt 2>&1)
cd scripts
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ../gpurun_out/s2h_launches_lq.csv python mb_sector_one.py lq 216 1296 296 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ../gpurun_out/s2h_launches_qr.csv python mb_sector_one.py qr 1296 216 296 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_rowstream -s 70 -c 1 -o ../gpurun_out/s2h_prof_rowstream_1296x216x216 python mb_gemm.py 592 > ../gpurun_out/s2h_ncu_rowstream.log 2>&1
