"""Lock-step batching: nb Markov chains advanced together must reproduce nb independent single-chain
runs (each of which is pinned to the reference by test_vmc_golden.py): identical configurations,
amplitudes / energies / gradients within 1e-10.  Uses truncation-free lattices (Dc = D^2 rows of
length 3/4) so that the different cache-warmth of the lock-step run cannot change values."""
import numpy as np
import pytest

import tnsp_b200.TAT as TAT
from golden_loader import build_lattice, load
from tnsp_b200.tetragono import models
from tnsp_b200.tetragono.observer import Observer
from tnsp_b200.tetragono.sampling import ChainRng, SweepSampling


def _run(lat, Dc, seeds, conf0, n_steps, batched):
    nb = len(seeds)
    out = {"conf": [], "poss": []}
    if batched:
        rng = ChainRng(nb)
        rng.seed(seeds)
        s = SweepSampling(lat, Dc, nb=nb, rng=rng)
        s.configuration.import_configuration(np.broadcast_to(conf0, (nb,) + conf0.shape))
        obs = Observer(lat, enable_energy=True, enable_gradient=True)
        with obs:
            for _ in range(n_steps):
                p, c = s()
                out["conf"].append(c.export_configuration())
                out["poss"].append(np.array(p))
                obs(p, c)
        out["energy"] = obs.total_energy
        out["grad"] = [np.asarray(t.storage) for row in obs.gradient for t in row]
        out["weight"] = obs._total_weight
        return out
    obs = Observer(lat, enable_energy=True, enable_gradient=True)
    confs, poss = [], []
    with obs:
        for c_i, seed in enumerate(seeds):
            rng = ChainRng(1)
            rng.seed([seed])
            s = SweepSampling(lat, Dc, nb=1, rng=rng)
            s.configuration.import_configuration(conf0)
            cc, pp = [], []
            for _ in range(n_steps):
                p, c = s()
                cc.append(c.export_configuration())
                pp.append(p)
                obs(p, c)
            confs.append(cc)
            poss.append(pp)
    out["conf"] = [np.stack([confs[c][t] for c in range(nb)]) for t in range(n_steps)]
    out["poss"] = [np.array([poss[c][t] for c in range(nb)]) for t in range(n_steps)]
    out["energy"] = obs.total_energy
    out["grad"] = [np.asarray(t.storage) for row in obs.gradient for t in row]
    out["weight"] = obs._total_weight
    return out


@pytest.mark.parametrize("shape", [(3, 3, 2, 4), (2, 4, 3, 9)])
def test_lockstep_equals_independent_chains(shape):
    L1, L2, D, Dc = shape
    lat = models.random_sampling_lattice(models.heisenberg_lattice(L1, L2, D), 99)
    conf0 = models.neel_configuration(L1, L2)
    seeds = [101, 202, 303, 404, 505]
    a = _run(lat, Dc, seeds, conf0, 3, batched=True)
    b = _run(lat, Dc, seeds, conf0, 3, batched=False)
    for t in range(3):
        assert np.array_equal(a["conf"][t], b["conf"][t])
        assert np.allclose(a["poss"][t], b["poss"][t], rtol=1e-10, atol=0)
    assert np.allclose(a["energy"], b["energy"], rtol=1e-9)
    assert abs(a["weight"] - b["weight"]) <= 1e-12 * abs(b["weight"])
    scale = max(np.abs(g).max() for g in b["grad"])
    for ga, gb in zip(a["grad"], b["grad"]):
        assert np.abs(ga - gb).max() <= 1e-9 * scale


def test_chain_rng_matches_reference_seed_recipe():
    """chain 0 of a ChainRng seeded like the reference's seed_differ draws the same numbers as the global
    engine after the same recipe (utility.py:146-150)."""
    TAT.random.seed(5)
    rng = ChainRng(3)
    base = rng.seed_like_reference()
    TAT.random.seed(base % 2**31)
    TAT.random.uniform_real(0, 1)()
    want = [TAT.random.uniform_int(0, 9)() for _ in range(5)]
    got = [int(rng.uniform_int(np.full(3, 9), None)[0]) for _ in range(5)]
    assert got == want


@pytest.mark.parametrize("size,nb", [(2, 3), (4, 5), (3, 7)])
def test_batched_ergodic_enumeration_covers_every_configuration_on_every_rank(size, nb):
    """several ranks x lock-step batches that do not divide the number of configurations: the union of what the ranks are handed
    (driver rule: rank r takes the calls with step % size == r) is every configuration exactly once"""
    from tnsp_b200.tetragono.sampling import ErgodicSampling
    lat = models.random_sampling_lattice(models.heisenberg_lattice(3, 3, 2), 5)
    seen = []
    calls = None
    for rank in range(size):
        s = ErgodicSampling(lat, 4, rank=rank, size=size, nb=nb)
        calls = s.calls
        for step in range(calls):
            if step % size == rank:
                p, c = s()
                conf = c.export_configuration().reshape(nb, -1)
                for k in range(nb):
                    if np.isfinite(np.atleast_1d(p)[k]):
                        seen.append(int("".join(str(int(x)) for x in conf[k]), 2))
    assert len(seen) == 512 and sorted(seen) == list(range(512))
