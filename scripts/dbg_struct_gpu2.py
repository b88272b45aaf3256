import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, time
from tnsp_b200 import backend
import tnsp_b200.TAT as TAT
from tnsp_b200.tetragono import models, dense_embedding as de
from tnsp_b200.tetragono.state import SamplingLattice
from tnsp_b200.tetragono.configuration import Configuration
B = backend.get()
L, D, Dc = 6, 6, 36
TAT.random.seed(2333)
lat = SamplingLattice(models.j1j2_abstract_lattice(TAT.BoseU1.D.Tensor, L, L, D, 1.0, 0.5))
pts = models.neel_points(lat)
dl = de.embed_lattice(lat)
count = [0]
def pat(M):
    return "\n".join("".join("x" if v else "." for v in row) for row in (M != 0))
def show(name, t, shape):
    print(name, shape); print(pat(t.cpu().numpy()[0][:shape[0]*shape[1]].reshape(shape)))
oq, osv = B.qr, B.svd
def qr(plan, a, o1, o2):
    a0 = a.clone(); r = oq(plan, a, o1, o2); torch.cuda.synchronize()
    count[0] += 1
    if count[0] <= 14:
        m, n, k = [int(x) for x in plan.sectors[0][:3]]
        print("=== QR", m, n, k, "flag", plan.flag); show("A", a0, (m, n)); show("out1", o1, (m, k)); show("out2", o2, (k, n))
    return r
def svd(plan, a, o1, s, o2):
    r = osv(plan, a, o1, s, o2); torch.cuda.synchronize()
    count[0] += 1
    if count[0] <= 14:
        m, n, k = [int(x) for x in plan.sectors[0][:3]]
        print("=== SVD", m, n, k); show("A", a, (m, n)); show("U", o1, (m, k)); print("S", s.cpu().numpy()[0]); show("Vt", o2, (k, n))
    return r
B.qr, B.svd = qr, svd
conf = Configuration(dl, Dc, 1)
conf.import_configuration(de.embed_configuration(lat, pts))
print(float(conf.hole(())))
