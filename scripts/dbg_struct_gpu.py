import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, time
from tnsp_b200 import backend
import tnsp_b200.TAT as TAT
from tnsp_b200.tetragono import models, dense_embedding as de
from tnsp_b200.tetragono.state import SamplingLattice
from tnsp_b200.tetragono.configuration import Configuration
from oracle.numpy_backend import _zero_pattern_blocks
B = backend.get()
L, D, Dc = 6, int(sys.argv[1]), int(sys.argv[2])
TAT.random.seed(2333)
lat = SamplingLattice(models.j1j2_abstract_lattice(TAT.BoseU1.D.Tensor, L, L, D, 1.0, 0.5))
pts = models.neel_points(lat)
dl = de.embed_lattice(lat)
stats = []
def analyse(kind, plan, a, dt):
    for (m, n, k, aoff, *_) in plan.sectors:
        M = a.cpu().numpy()[0, aoff:aoff + m * n].reshape(m, n)
        blocks = _zero_pattern_blocks(M)
        big = max([len(r) * len(c) for r, c in blocks] + [0])
        stats.append((kind, int(m), int(n), len(blocks), round(float((M != 0).mean()), 3), big, round(dt * 1e3, 2), bool(np.isfinite(M).all())))
oq, osv = B.qr, B.svd
def qr(plan, a, o1, o2):
    a0 = a.clone(); torch.cuda.synchronize(); t = time.perf_counter(); r = oq(plan, a, o1, o2); torch.cuda.synchronize(); analyse("qr", plan, a0, time.perf_counter() - t); return r
def svd(plan, a, o1, s, o2):
    torch.cuda.synchronize(); t = time.perf_counter(); r = osv(plan, a, o1, s, o2); torch.cuda.synchronize(); analyse("svd", plan, a, time.perf_counter() - t); return r
B.qr, B.svd = qr, svd
conf = Configuration(dl, Dc, 1)
conf.import_configuration(de.embed_configuration(lat, pts))
print(float(conf.hole(())))
for st in stats: print(st)
