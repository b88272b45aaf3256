"""Child process of tests/test_checkpoint.py: runs in the UNMODIFIED reference's environment (PyTAT from oracle/_ref,
tetragono from /root/reference, single-rank mpi4py stub) and loads a checkpoint written by this repository.
    python tests/ref_load_checkpoint.py <checkpoint.pkl> <fixture.npz>   -> prints "ws <value> energy <value>" """
import json
import pickle
import sys

import numpy as np
import TAT
import tetragono as tet

with open(sys.argv[1], "rb") as f:
    lattice = pickle.load(f)
assert type(lattice) is tet.SamplingLattice, type(lattice)
z = np.load(sys.argv[2])
meta = json.loads(bytes(z["meta"]).decode())
S = lattice.Symmetry
conf = tet.sampling_lattice.Configuration(lattice, meta["Dc"])
for l1 in range(meta["L1"]):
    for l2 in range(meta["L2"]):
        for o, p in meta["config"][l1][l2].items():
            conf[l1, l2, int(o)] = (S(*p[0]), p[1])
ws = conf.hole(())
obs = tet.sampling_lattice.Observer(lattice, enable_energy=True)
with obs:
    obs(float(ws)**2, conf)
print("ws %.17g energy %.17g" % (float(ws), obs._whole_result_reweight["energy"] / obs._total_weight))
