# call G: ncu launch list of the default bench command at 2368 chains + ncu --set full of steady-state sector kernels (reports exported
# to csv on the box: gpurun_out/ carries at most 64 MiB back)
# (first: the pair-parallel rotation variant of the sector SVD, TNSP_RT_JACOBI=3)
mkdir -p gpurun_out
( for v in 2 3; do echo "== TNSP_RT_JACOBI=$v"; TNSP_RT_JACOBI=$v timeout 300 python scripts/mb_rt_factor.py 2368 5 2>&1 | tail -3; done ) > gpurun_out/r2g_mb_jacobi.txt 2>&1
cat gpurun_out/r2g_mb_jacobi.txt
( TNSP_RT_JACOBI=3 timeout 900 python -m pytest tests/test_sector_kernels_gpu.py tests/test_cfg2_at_size_gpu.py tests/test_sector_fermi_gpu.py -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/r2g_tests_jacobi3.txt; cat gpurun_out/r2g_tests_jacobi3.txt
TNSP_RT_JACOBI=3 timeout 600 python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2g_cfg2_jacobi3.json 2> gpurun_out/r2g_cfg2_jacobi3.err
python -c "
import json; d=json.loads(open('gpurun_out/r2g_cfg2_jacobi3.json').read().strip().splitlines()[-1]); kb=d['kernel_breakdown']; print('jacobi3', d['value'], d['e2e']['value'], d['ms_per_step'], d['parity_check']['ok'], {k: round(kb[k]['ms']) for k in ('rt_svd_work','rt_qr_work','rt_gemm')})"
TNSP_RT_GEMM=38 timeout 600 python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2g_cfg2_gemm38.json 2> gpurun_out/r2g_cfg2_gemm38.err
python -c "
import json; d=json.loads(open('gpurun_out/r2g_cfg2_gemm38.json').read().strip().splitlines()[-1]); kb=d['kernel_breakdown']; print('gemm38', d['value'], d['e2e']['value'], d['ms_per_step'], d['parity_check']['ok'], {k: round(kb[k]['ms']) for k in ('rt_svd_work','rt_qr_work','rt_gemm')})"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launches_cfg2_nb2368.csv \
  python bench.py --workload cfg2 --steps 1 --warmup 2 --no-cpu-baseline --no-secondary > gpurun_out/r2g_ncu_bench.json 2> gpurun_out/r2g_ncu_bench.err
tail -c 300 gpurun_out/r2g_ncu_bench.err
python scripts/launch_list_summary.py gpurun_out/r2g_launches_cfg2_nb2368.csv 2368 > gpurun_out/r2g_launches_cfg2_nb2368.summary.txt
head -30 gpurun_out/r2g_launches_cfg2_nb2368.summary.txt
gzip -f gpurun_out/r2g_launches_cfg2_nb2368.csv
ncu_one() {  # tag, kernel regex, skip, count
  timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -o /tmp/$1 \
     python bench.py --workload cfg2 --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > /tmp/$1.json 2> /tmp/$1.err
  tail -c 200 /tmp/$1.err
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  ls -la /tmp/$1.ncu-rep gpurun_out/$1.*
}
ncu_one r2g_ncu_stream 'rt_(gemm_warp|repack_tile|sort)' 64000 40
ncu_one r2g_ncu_factor 'rt_(qr_work|svd_work)' 12000 32
python scripts/ncu_raw_summary.py gpurun_out/r2g_ncu_stream.raw.csv gpurun_out/r2g_ncu_factor.raw.csv > gpurun_out/r2g_ncu_full.summary.txt
head -60 gpurun_out/r2g_ncu_full.summary.txt
du -sh gpurun_out
