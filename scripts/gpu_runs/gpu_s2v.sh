set -x
timeout 400 python bench.py --workload cfg1 --steps 4 --warmup 3 > gpurun_out/s2v_bench_cfg1.json 2> gpurun_out/s2v_bench_cfg1.err
timeout 400 python bench.py --workload heis6 --chains 592 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s2v_bench_heis6.json 2> gpurun_out/s2v_bench_heis6.err
