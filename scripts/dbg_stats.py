import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from tnsp_b200 import backend
from tnsp_b200.TAT import ragged
import test_sector_kernels_gpu as t
B = backend.get()
case = t._case(1, 6, (40, 3), (90,), 70, 0)
print("on", B.rt_stats(enable=1, reset=True))
out = t._run(B, case, 20)
print("stats", B.rt_stats(enable=0, read=True))
print("overflow", B.rt_overflow())
