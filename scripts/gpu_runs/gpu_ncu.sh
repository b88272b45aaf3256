set -x
cd scripts
ncu --set full --clock-control none --import-source on -k regex:sector -s 2 -c 1 -o ../gpurun_out/prof_lq_216x1296 python mb_one.py lq 216 1296 7 148 > ../gpurun_out/ncu_lq.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sector -s 2 -c 1 -o ../gpurun_out/prof_svd_216x216 python mb_one.py svd 216 216 6 148 > ../gpurun_out/ncu_svd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_grouped -s 2 -c 1 -o ../gpurun_out/prof_gemm_1296x36x36 python mb_one.py gemm 1296 36 36 592 > ../gpurun_out/ncu_gemm.log 2>&1
python mb_factor.py 148 > ../gpurun_out/mb_148.txt 2>&1
