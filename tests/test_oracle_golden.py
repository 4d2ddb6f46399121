"""The oracle port against the vectors produced by the unmodified reference (tests/golden)."""
import numpy as np
import pytest
import torch

from oracle import mcem_port
from tests.golden_io import Golden, case_names


@pytest.mark.parametrize("name", case_names())
def test_port_reproduces_reference_bit_for_bit(name):
    torch.set_num_threads(1)
    g = Golden(name)
    o = mcem_port.MCEMOracle(g.variant, g.niter, *g.sched, g.var_RW, draws=mcem_port.ReplayDraws(g.draws))
    o.init_parameters(g.X, g.S, g.sd, g.K, g.eps, y=g.y)
    assert np.array_equal(o.Z.numpy(), g.ref["Z0"])
    cost = o.run()
    assert o.draws.pos == len(g.draws)
    assert tuple(o.Vs.shape) == tuple(g.ref["Vs_shape"])
    # bit-exact on the machine that generated the fixtures; 1e-6 guards a different BLAS build
    for k, v in dict(cost=cost, S_hat=o.S_hat, N_hat=o.N_hat, W=o.W.numpy(), H=o.H.numpy(), g=o.g.numpy(),
                     Z=o.Z.numpy()).items():
        np.testing.assert_allclose(v, g.ref[k], rtol=1e-5, atol=1e-6, err_msg=k)


@pytest.mark.parametrize("name", case_names())
def test_container_forward(name):
    g = Golden(name)
    rows = torch.tensor(np.abs(g.S.T[:5]) ** 2)
    yrows = None if g.y is None else torch.tensor(g.y.T[:5])
    v = "M2v3" if g.variant == "M2v2" else g.variant
    x_mu, z, mu, lv = mcem_port.vae_forward(g.sd, v, rows, yrows, mcem_port.ReplayDraws([("randn", g.ref["fwd_eps"])]))
    np.testing.assert_allclose(x_mu.numpy(), g.ref["fwd_xmu"], rtol=1e-5)
    np.testing.assert_allclose(mu.numpy(), g.ref["fwd_mu"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(lv.numpy(), g.ref["fwd_lv"], rtol=1e-5, atol=1e-6)


def test_m1_schedule_shift():
    # MCEM_M1(niter, 10, 30, 25, 75) really runs 60 iterations keeping 30, then 105 keeping 75
    o = mcem_port.MCEMOracle("M1", 1, 10, 30, 25, 75)
    assert o.schedule() == ((30, 30), (75, 30))
    o = mcem_port.MCEMOracle("M2", 1, 10, 30, 25, 75)
    assert o.schedule() == ((10, 30), (25, 75))


def test_draw_count_matches_survey():
    # M1 consumes 4 + niter*2*60 + 2*105 draws; checked on the tiny fixture's own schedule
    g = Golden("tiny_M1")
    (kE, bE), (kW, bW) = mcem_port.MCEMOracle("M1", g.niter, *g.sched).schedule()
    assert len(g.draws) == 4 + g.niter * 2 * (kE + bE) + 2 * (kW + bW)
    assert [k for k, _ in g.draws[:4]] == ["rand", "rand", "randn", "randn"]


def test_energy_ratios_port_matches_reference_golden():
    """packages/metrics.py:39-82 on the seeded signals of tests/golden/metrics.npz (made by oracle/make_golden.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "metrics.npz"))
    for c in range(g["s"].shape[0]):
        r = mcem_port.energy_ratios(g["s_hat"][c], g["s"][c], g["n"][c])
        assert np.allclose(r, g["ref"][c, :3], rtol=0, atol=1e-9)
        assert abs(mcem_port.si_sdr(g["s_hat"][c], g["s"][c]) - g["ref"][c, 3]) < 1e-9
