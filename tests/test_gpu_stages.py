"""Stage-wise parity of the CUDA kernels (called through the C ABI) against the oracle with injected inputs.

Tolerances follow BASELINE.json: NMF / Wiener outputs rel <= 1e-4 (FP32); MLP outputs in FP32 mode to FP32 rounding.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from dvae_b200 import _lib
from dvae_b200.engine import McemConfig, McemEngine, RaggedBatch, VaeWeights, _ld_for, _p, _stream, mlp_forward
from oracle import mcem_port
from tests.golden_io import Golden
from tests.gpu_util import DEV, engine_for, fm, golden_draws, relerr, unfm

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["tiny_M1", "tiny_M2", "full_M1", "full_M2", "full_M2v3", "cfg1_M1"])
def test_encoder_decoder_match_oracle(name):
    g = Golden(name)
    w = VaeWeights(g.sd, g.variant, torch.device(DEV))
    P = torch.tensor(np.abs(g.X) ** 2)
    y = None if g.y is None else torch.tensor(g.y)
    enc_in = torch.t(torch.cat([P, y], 0)) if g.variant == "M2" else torch.t(P)
    _, mu_ref, lv_ref = mcem_port.encoder_forward(g.sd, enc_in, mcem_port.TorchDraws())
    x = enc_in.contiguous().to(DEV)
    mu = mlp_forward(w.enc_mu, x, _lib.ACT_NONE).cpu().numpy()
    lv = mlp_forward(w.enc_lv, x, _lib.ACT_NONE).cpu().numpy()
    assert relerr(mu, mu_ref.numpy()) <= 2e-5 and relerr(lv, lv_ref.numpy()) <= 2e-5
    # x2 path (labels as a second input) must agree with the concatenated input
    if g.variant == "M2":
        mu2 = mlp_forward(w.enc_mu, torch.t(P).contiguous().to(DEV), _lib.ACT_NONE, x2=torch.t(y).contiguous().to(DEV)).cpu().numpy()
        assert np.array_equal(mu2, mu)
    z = torch.t(mu_ref)
    dec_in = torch.t(z if y is None else torch.cat([z, y], 0)).contiguous()
    ref = mcem_port.decoder_forward(g.sd, dec_in).numpy()
    got = mlp_forward(w.dec, dec_in.to(DEV), _lib.ACT_EXP).cpu().numpy()
    assert np.max(np.abs(got - ref) / ref) <= 2e-5


def test_container_forward_matches_reference_golden():
    from dvae_b200.packages.models.models import DeepGenerativeModel, DeepGenerativeModel_v5, VariationalAutoencoder
    for name in ("full_M1", "full_M2", "full_M2v3"):
        g = Golden(name)
        if g.variant == "M1":
            model = VariationalAutoencoder([g.F, g.L, g.h])
        elif g.variant == "M2":
            model = DeepGenerativeModel([g.F, g.y_dim, g.L, g.h], None)
        else:
            model = DeepGenerativeModel_v5([g.F, g.y_dim, g.L, g.h]).enc_dec_clf
        missing, unexpected = model.load_state_dict({k: torch.tensor(v) for k, v in g.sd.items()}, strict=False)
        assert not unexpected and all(k.startswith("classifier.") for k in missing)
        model.to(DEV).eval()
        rows = torch.tensor(np.abs(g.S.T[:5]) ** 2).to(DEV)
        # replay the reference's reparametrisation draw through the CPU generator patch
        eps = torch.tensor(g.ref["fwd_eps"])
        real = torch.randn
        torch.randn = lambda *a, **k: eps.clone()
        try:
            out = model(rows) if g.variant == "M1" else model(rows, torch.tensor(g.y.T[:5]).to(DEV))
        finally:
            torch.randn = real
        assert np.max(np.abs(out[0].cpu().numpy() - g.ref["fwd_xmu"]) / g.ref["fwd_xmu"]) <= 5e-5
        assert relerr(out[-2].cpu().numpy(), g.ref["fwd_mu"]) <= 2e-5
        assert relerr(out[-1].cpu().numpy(), g.ref["fwd_lv"]) <= 2e-5


def _random_mstep_inputs(F, N, K, R, seed):
    rng = np.random.default_rng(seed)
    P = rng.gamma(1.0, 1.0, size=(F, N)).astype(np.float32) * 0.1
    Vs = rng.gamma(2.0, 0.05, size=(R, F, N)).astype(np.float32)
    W = np.maximum(rng.uniform(size=(F, K)), 1e-8).astype(np.float32)
    H = np.maximum(rng.uniform(size=(K, N)), 1e-8).astype(np.float32)
    g = rng.uniform(0.5, 1.5, size=N).astype(np.float32)
    return P, Vs, W, H, g


@pytest.mark.parametrize("F,N,K,R", [(513, 37, 10, 30), (513, 185, 10, 10), (33, 9, 3, 3), (513, 5, 10, 1),
                                     (513, 21, 10, 60), (513, 13, 10, 40), (513, 9, 7, 160)])      # windowed kernel (multi-chain sample counts)
def test_m_step_matches_oracle(F, N, K, R):
    P, Vs, W, H, g = _random_mstep_inputs(F, N, K, R, seed=F + N)
    ref = mcem_port.m_step_reference(torch.tensor(P), torch.tensor(Vs), torch.tensor(W), torch.tensor(H), torch.tensor(g))
    ld = _ld_for(F)
    # two utterances in one ragged batch: the same data twice but split differently must give the same answer
    for split in ([N], [N // 2, N - N // 2] if N > 1 else [N]):
        batch = RaggedBatch(split, DEV)
        B = batch.B
        Pd = fm(P)
        Vsd = torch.zeros((N, R, ld), device=DEV)
        Vsd[:, :, :F] = torch.from_numpy(np.ascontiguousarray(Vs.transpose(2, 0, 1))).to(DEV)
        Wd = torch.zeros((B, K, ld), device=DEV)
        Wd[:, :, :F] = torch.from_numpy(np.ascontiguousarray(W.T)).to(DEV)[None]
        Hd = torch.from_numpy(np.ascontiguousarray(H.T)).to(DEV).contiguous()
        gd = torch.from_numpy(g).to(DEV)
        Vbd = torch.zeros((N, ld), device=DEV)
        _lib.call("dvae_nmf_vb", _p(Wd), _p(Hd), _p(batch.frame_utt), N, F, K, ld, _p(Vbd), _stream())
        assert relerr(unfm(Vbd, F), W @ H) <= 1e-5
        if B > 1:
            continue   # a split batch updates each half's own W: only the Vb kernel is comparable
        cost = torch.zeros(B, dtype=torch.float64, device=DEV)
        ws = torch.empty(int(_lib.load().dvae_nmf_workspace_floats(B, K, ld, batch.max_frames)), device=DEV)
        _lib.call("dvae_nmf_mstep", _p(Pd), _p(Vsd), R, _p(Wd), _p(Hd), _p(gd), _p(Vbd), _p(cost), _p(batch.fr_off),
                  _p(batch.frame_utt), B, N, F, K, ld, batch.max_frames, _p(ws), None, None, _stream())
        assert relerr(Wd[0, :, :F].t().cpu().numpy(), ref["W"].numpy()) <= 1e-4
        assert relerr(Hd.t().cpu().numpy(), ref["H"].numpy()) <= 1e-4
        assert relerr(gd.cpu().numpy(), ref["g"].numpy()) <= 1e-4
        assert relerr(unfm(Vbd, F), ref["Vb"].numpy()) <= 1e-4
        assert abs(cost.item() - ref["cost"].item()) <= 1e-4 * abs(ref["cost"].item())


def test_m_step_batch_equals_single():
    """Ragged batching: utterances processed together give the results of processing them alone (bit-exact)."""
    F, K, R = 513, 10, 4
    ld = _ld_for(F)
    sizes = [7, 1, 12]
    data = [_random_mstep_inputs(F, n, K, R, seed=100 + i) for i, n in enumerate(sizes)]

    def run(idx):
        batch = RaggedBatch([sizes[i] for i in idx], DEV)
        N = batch.NT
        Pd = torch.cat([fm(data[i][0]) for i in idx])
        Vsd = torch.zeros((N, R, ld), device=DEV)
        Vsd[:, :, :F] = torch.cat([torch.from_numpy(np.ascontiguousarray(data[i][1].transpose(2, 0, 1))) for i in idx]).to(DEV)
        Wd = torch.zeros((len(idx), K, ld), device=DEV)
        for j, i in enumerate(idx):
            Wd[j, :, :F] = torch.from_numpy(np.ascontiguousarray(data[i][2].T)).to(DEV)
        Hd = torch.cat([torch.from_numpy(np.ascontiguousarray(data[i][3].T)) for i in idx]).to(DEV).contiguous()
        gd = torch.cat([torch.from_numpy(data[i][4]) for i in idx]).to(DEV)
        Vbd = torch.zeros((N, ld), device=DEV)
        _lib.call("dvae_nmf_vb", _p(Wd), _p(Hd), _p(batch.frame_utt), N, F, K, ld, _p(Vbd), _stream())
        cost = torch.zeros(len(idx), dtype=torch.float64, device=DEV)
        ws = torch.empty(int(_lib.load().dvae_nmf_workspace_floats(len(idx), K, ld, batch.max_frames)), device=DEV)
        _lib.call("dvae_nmf_mstep", _p(Pd), _p(Vsd), R, _p(Wd), _p(Hd), _p(gd), _p(Vbd), _p(cost), _p(batch.fr_off),
                  _p(batch.frame_utt), len(idx), N, F, K, ld, batch.max_frames, _p(ws), None, None, _stream())
        return Wd.cpu(), Hd.cpu(), gd.cpu(), cost.cpu(), batch

    Wb, Hb, gb, cb, batch = run([0, 1, 2])
    for j in range(3):
        W1, H1, g1, c1, _ = run([j])
        a, b = batch.fr_off_host[j], batch.fr_off_host[j + 1]
        assert torch.equal(Wb[j], W1[0]) and torch.equal(Hb[a:b], H1) and torch.equal(gb[a:b], g1)
        assert abs(cb[j].item() - c1[0].item()) <= 1e-12 * abs(c1[0].item())


def test_wiener_matches_oracle():
    F, N, R = 513, 23, 7
    rng = np.random.default_rng(5)
    Vs = rng.gamma(2.0, 0.05, size=(R, F, N)).astype(np.float32)
    Vb = rng.gamma(2.0, 0.02, size=(F, N)).astype(np.float32)
    g = rng.uniform(0.5, 1.5, size=N).astype(np.float32)
    X = (rng.standard_normal((F, N)) + 1j * rng.standard_normal((F, N))).astype(np.complex64)
    Vs_scaled = torch.tensor(g) * torch.tensor(Vs)
    Vx = Vs_scaled + torch.tensor(Vb)
    WFs, WFn = torch.mean(Vs_scaled / Vx, axis=0).numpy(), torch.mean(torch.tensor(Vb) / Vx, axis=0).numpy()
    ld = _ld_for(F)
    Vsd = torch.zeros((N, R, ld), device=DEV)
    Vsd[:, :, :F] = torch.from_numpy(np.ascontiguousarray(Vs.transpose(2, 0, 1))).to(DEV)
    a = torch.empty((N, ld), device=DEV)
    b = torch.empty((N, ld), device=DEV)
    Vbd, gd = fm(Vb), torch.from_numpy(g).to(DEV)          # keep the tensors alive while the kernel runs
    _lib.call("dvae_wiener_accum", _p(Vsd), R, _p(Vbd), _p(gd), N, F, ld, _p(a), _p(b), 1, _stream())
    Xd = fm(X, dtype=torch.complex64)
    S = torch.empty_like(Xd)
    Nn = torch.empty_like(Xd)
    _lib.call("dvae_wiener_apply", _p(Xd), _p(a), _p(b), R, N, F, ld, _p(S), _p(Nn), _stream())
    assert relerr(unfm(S, F), WFs * X) <= 1e-4 and relerr(unfm(Nn, F), WFn * X) <= 1e-4
    assert np.max(np.abs(unfm(a, F) / R - WFs) / WFs) <= 1e-5


@pytest.mark.parametrize("name", ["tiny_M1", "tiny_M2", "tiny_M2v3", "full_M1", "full_M2", "cfg1_M1"])
def test_mh_log_acceptance_and_decisions(name):
    """One sample_posterior call with the reference's own draws: log-acceptance values and accept decisions."""
    g = Golden(name)
    o = mcem_port.MCEMOracle(g.variant, g.niter, *g.sched, g.var_RW, draws=mcem_port.ReplayDraws(g.draws))
    o.init_parameters(g.X, g.S, g.sd, g.K, g.eps, y=g.y)
    (kE, bE), _ = o.schedule()
    o.taps = []
    Zs_ref = o.sample_posterior(o.Z, kE, bE)

    eng, X, P, y, batch = engine_for(g, o.schedule())
    draws = golden_draws(g, o.schedule())
    eng.init_parameters(X, P, batch, y, draws)
    assert relerr(eng.Z.t().cpu().numpy(), g.ref["Z0"]) <= 2e-5
    N, L = g.X.shape[1], g.L
    a_trace = torch.zeros((kE + bE, N), device=DEV)
    Zs = eng.sample_posterior(kE, bE, draws, a_trace=a_trace).cpu().numpy()       # [N][keep][L]
    a_trace = a_trace.cpu().numpy()

    same = np.ones(N, bool)         # chains whose accept history still equals the oracle's
    n_dec = n_flip = 0
    for it, tap in enumerate(o.taps):
        a_ref = tap["a"].numpy()
        # |a| sums 2F terms of size ~|log Vx| + P/Vx: FP32 summation noise scales with that magnitude
        tol = 2e-4 * (1.0 + np.abs(a_ref)) + 5e-3
        bad = same & (np.abs(a_trace[it] - a_ref) > tol)
        assert not bad.any(), "iteration %d: log-acceptance off by %g" % (it, np.max(np.abs(a_trace[it] - a_ref)[same]))
        acc_gpu = np.log(tap["u"].numpy()) < a_trace[it]
        acc_ref = tap["acc"].numpy()
        n_dec += int(same.sum())
        n_flip += int((same & (acc_gpu != acc_ref)).sum())
        same &= acc_gpu == acc_ref
    assert n_flip <= max(1, int(0.01 * n_dec)), "%d of %d decisions flipped" % (n_flip, n_dec)
    ref = Zs_ref.numpy()
    ok = same
    assert ok.sum() >= 0.9 * N
    assert np.max(np.abs(Zs[ok] - ref[ok])) <= 1e-5


def test_philox_draws_are_sharding_independent_and_sane():
    L, C_, n_iter = 16, 2, 3
    rng = _lib.DvaeRng()
    rng.seed, rng.iter0 = 1234, 5

    def dump(utt_ids, nfr):
        b = RaggedBatch(nfr, DEV, utt_ids)
        eps = torch.empty((n_iter, b.NT * C_, L), device=DEV)
        u = torch.empty((n_iter, b.NT * C_), device=DEV)
        _lib.call("dvae_rng_dump", C.byref(rng), _p(b.frame_gid), _p(b.frame_idx), b.NT, C_, L, n_iter, _p(eps), _p(u), _stream())
        return eps.cpu(), u.cpu(), b

    e_all, u_all, b = dump([7, 8, 9], [50, 20, 30])
    e_one, u_one, _ = dump([8], [20])
    a = int(b.fr_off_host[1]) * C_
    assert torch.equal(e_all[:, a:a + 20 * C_], e_one) and torch.equal(u_all[:, a:a + 20 * C_], u_one)
    big_e, big_u, _ = dump(list(range(64)), [185] * 64)
    assert abs(big_e.mean().item()) < 5e-3 and abs(big_e.var().item() - 1.0) < 1e-2
    assert 0.0 < big_u.min().item() and big_u.max().item() < 1.0 and abs(big_u.mean().item() - 0.5) < 5e-3
    assert abs(torch.corrcoef(torch.stack([big_e[0, :, 0], big_e[0, :, 1]]))[0, 1].item()) < 2e-2
    # chains of one frame and successive iterations get different numbers
    assert not torch.equal(big_e[0, 0], big_e[0, 1]) and not torch.equal(big_e[0, 0], big_e[1, 0])


def test_argument_errors_surface_as_exceptions():
    g = Golden("tiny_M1")
    w = VaeWeights(g.sd, "M1", torch.device(DEV))
    with pytest.raises(ValueError):
        McemEngine(w, McemConfig(nmf_rank=17), DEV)
    with pytest.raises(ValueError):
        VaeWeights(g.sd, "M2", torch.device(DEV))
    x = torch.zeros((4, 7), device=DEV)
    with pytest.raises(ValueError):
        mlp_forward(w.dec, x, _lib.ACT_EXP)           # wrong input width -> DVAE_ERR_ARG -> ValueError
