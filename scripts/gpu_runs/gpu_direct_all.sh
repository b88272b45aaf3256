# every direct-sampling fixture on the GPU backend, one line per case (a failing case does not hide the others)
cd tests
timeout 30 python - > ../gpurun_out/direct_gpu_all.txt 2>&1 <<'EOF'
import sys, traceback
sys.path.insert(0, "..")
from tnsp_b200 import backend
backend.set_backend(None)
backend.get()
from test_direct_sampling import test_direct_sampling_matches_the_reference as case
for c in ["tJ_4x4_D1_Dc8", "hubbardFF_4x4_D1_Dc8", "heis_4x4_D3_Dc5_truncating", "heis_3x3_D2_Dc4", "heisU1_4x4_d1_Dc6"]:
    try:
        case(c)
        print(c, "direct sampling on the GPU backend matches the reference: True", flush=True)
    except Exception as e:
        print(c, "FAILED:", repr(e)[:300], flush=True)
EOF
