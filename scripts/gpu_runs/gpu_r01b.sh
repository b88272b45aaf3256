set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r01b_gpu.txt 2>&1
nproc >> gpurun_out/r01b_gpu.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r01b_tests.txt
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r01b_smoke.txt 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/r01b_bench_cfg1.json 2> gpurun_out/r01b_bench_cfg1.err
timeout 900 python bench.py --workload heis6 --chains 64 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01b_bench_heis6.json 2> gpurun_out/r01b_bench_heis6.err
python - > gpurun_out/r01b_peaks.txt 2>&1 <<'PY'
import torch
from tnsp_b200 import profiling
print("dgemm_tflops_4096", profiling.measure_fp64_gemm_tflops(4096))
print("dgemm_tflops_8192", profiling.measure_fp64_gemm_tflops(8192))
a = torch.empty(1<<28, dtype=torch.float64, device="cuda"); b = torch.empty_like(a)
b.copy_(a); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): b.copy_(a)
e1.record(); torch.cuda.synchronize()
print("copy_gbs", 5*2*a.numel()*8/ (e0.elapsed_time(e1)*1e-3)/1e9)
PY
