set -x
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/fin2_tests.txt
timeout 120 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' > gpurun_out/fin2_smoke.txt 2>&1
timeout 900 python bench.py > gpurun_out/fin2_bench_cfg2.json 2> gpurun_out/fin2_bench_cfg2.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/fin2_ref_cfg2.json 2> gpurun_out/fin2_ref_cfg2.err
