"""-m gpu: BASELINE cfg2 at its real size (6x6 J1-J2 U(1), D = 6, Dc = 36) through the sm_100a kernels against the fixture of the
unmodified reference: a lock-step batch of 64 chains on the sector-compact engine (work queues, footprint classes, grouped sector
GEMM are the code under test), the same batch on the charge-dense embedding of round 1, and the reference's sweep trajectory with
gradient and SR-CG."""
import numpy as np
import pytest

from test_cfg2_at_size import NAME, at_size_cold_check, at_size_trajectory_check

pytestmark = pytest.mark.gpu


def test_cfg2_at_size_cold_64_chains_sector_engine():
    at_size_cold_check(64, holes=True)


def test_cfg2_at_size_trajectory_gradient_sr():
    at_size_trajectory_check()


def test_cfg2_at_size_cold_64_chains_dense_embedding():
    """the round-1 engine at the benched size: dense tensors with exact zeros, sectors discovered on the device"""
    import tnsp_b200.TAT as TAT
    from golden_loader import build_lattice, load
    from tnsp_b200.tetragono import dense_embedding as de
    from tnsp_b200.tetragono.configuration import Configuration
    from tnsp_b200.tetragono.observer import Observer
    from golden_loader import config_points
    meta, z = load(NAME)
    lat = build_lattice(meta, z)
    B = TAT.tensor._bk.get()
    saved = (B.sector_discovery, B.lib.tnsp_gemm_skip_zero_fragments(-1))
    try:
        dl = de.embed_lattice(lat)
        nb = 64
        conf = Configuration(dl, meta["Dc"], nb)
        c0 = de.embed_configuration(lat, config_points(meta))
        conf.import_configuration(np.broadcast_to(c0, (nb,) + c0.shape))
        ws = np.asarray(conf.hole(()).storage).reshape(-1)
        assert np.abs(ws - z["ws"][0]).max() <= 1e-10 * abs(z["ws"][0])
        obs = Observer(dl, enable_energy=True)
        with obs:
            obs(ws**2, conf)
        e = obs._whole_result_reweight["energy"] / obs._total_weight
        assert abs(e - z["energy_s"][0]) <= 1e-10 * abs(z["energy_s"][0])
    finally:
        B.sector_discovery = saved[0]
        B.lib.tnsp_gemm_skip_zero_fragments(saved[1])
