"""child process of test_reference_seam.py: the UNMODIFIED reference tetragono / tetraku / lazy (imported from /root/reference)
running on this repository's TAT module installed as `TAT`: the heis_3x3_D2_Dc4 case of tests/golden/make_golden.py, printed as
JSON for comparison with the fixture the reference produced on its own PyTAT."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import numpy_backend  # noqa: E402

numpy_backend.install()
import tnsp_b200.TAT as T  # noqa: E402

T.install_as_TAT()
import numpy as np  # noqa: E402
import TAT  # noqa: E402
import tetragono as tet  # noqa: E402
from tetraku.models.heisenberg import abstract_lattice  # noqa: E402

TAT.random.seed(2333)
lattice = tet.SamplingLattice(abstract_lattice(3, 3, 2, 1.0))
Dc = 4
S = lattice.Symmetry
points = [[{0: (S(), (l1 + l2) % 2)} for l2 in range(3)] for l1 in range(3)]
conf = tet.sampling_lattice.Configuration(lattice, Dc)
for l1 in range(3):
    for l2 in range(3):
        conf[l1, l2, 0] = points[l1][l2][0]
ws = float(conf.hole(()))
obs = tet.sampling_lattice.Observer(lattice, enable_energy=True, enable_gradient=True)
with obs:
    obs(ws**2, conf)
energy_s = obs._whole_result_reweight["energy"] / obs._total_weight
TAT.random.seed(11)
sampling = tet.sampling_lattice.SweepSampling(lattice, Dc, None, None)
for l1 in range(3):
    for l2 in range(3):
        sampling.configuration[l1, l2, 0] = points[l1][l2][0]
obs = tet.sampling_lattice.Observer(lattice, enable_energy=True, enable_gradient=True, enable_natural_gradient=True)
traj, poss = [], []
with obs:
    for _ in range(12):
        p, c = sampling()
        traj.append(np.asarray(c.export_configuration()).tolist())
        poss.append(float(p))
        obs(p, c)
grad = obs.gradient
out = {"ws": ws, "energy_s": float(energy_s), "traj_config": traj, "traj_possibility": poss, "traj_energy": [float(x) for x in obs.total_energy],
       "gradient": [[np.asarray(grad[l1][l2].storage).tolist() for l2 in range(3)] for l1 in range(3)],
       "gradient_names": [[list(grad[l1][l2].names) for l2 in range(3)] for l1 in range(3)]}
print("RESULT " + json.dumps(out))
