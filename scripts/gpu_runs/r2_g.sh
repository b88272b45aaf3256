# call G: ncu launch list of the default bench command at 2368 chains + ncu --set full of steady-state sector kernels (reports exported
# to csv on the box: gpurun_out/ carries at most 64 MiB back)
mkdir -p gpurun_out
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
