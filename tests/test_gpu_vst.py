"""The sampler's own BF16 emission of the kept samples' variances ("VsT", dvae_b200/csrc/vst.cu) and its consumers.

* the M-step on the emission against the oracle's M-step (mcem.py:91-153) GIVEN THE SAME (BF16-rounded) variances: 1e-4;
* pack / unpack round trip and bounds;
* what the sampler emits equals the decoder output of the kept latent samples (mcem.py:280-290) to the stated BF16 tolerance;
* Philox inside the sampler == the draws of dvae_rng_dump injected into the same sampler, bit for bit;
* NaN / Inf guard of the status word; device guard (engine on cuda:1 while cuda:0 is current).
"""
import ctypes as C

import numpy as np
import pytest
import torch

from dvae_b200 import _lib, synth, tc
from dvae_b200.engine import McemConfig, McemEngine, RaggedBatch, VaeWeights, _ld_for, _p, _stream, mlp_forward
from oracle import mcem_port
from tests.gpu_util import DEV, fm, relerr, unfm

pytestmark = pytest.mark.gpu
SENT = 0xA5


def _weights(variant="M1", seed=5, bias=float(np.log(0.05))):
    y_dim = 0 if variant == "M1" else 1
    sd = synth.xavier_state_dict(variant, 513, 16, [128, 128], y_dim, seed=seed, out_bias=bias)
    # a non-constant output bias, so that a wrong per-bin scale E[f] cannot hide
    sd["decoder.reconstruction.bias"] = (sd["decoder.reconstruction.bias"] + np.linspace(-1.0, 1.0, 513)).astype(np.float32)
    return VaeWeights(sd, variant, torch.device(DEV))


def _pack(w, Vsd, R, N):
    img = tc.decoder_image(w)
    nbytes = int(_lib.load().dvae_vst_bytes(N, R))
    vst = torch.full((nbytes + 4096,), SENT, dtype=torch.uint8, device=DEV)
    idx = torch.full(((N + 2) * 32,), SENT, dtype=torch.uint8, device=DEV)
    _lib.call("dvae_vst_pack", w.dec.ref, _p(img), w.z_dim, w.y_dim, _p(Vsd), R, N, Vsd.shape[2], _p(vst), _p(idx), _stream())
    return vst, idx, nbytes


def _unpack(w, vst, idx, R, N, ld=520):
    out = torch.full((N + 1, R, ld), -7.0, device=DEV)
    _lib.call("dvae_vst_unpack", w.dec.ref, _p(tc.decoder_image(w)), w.z_dim, w.y_dim, _p(vst), _p(idx), R, N, ld, _p(out), _stream())
    return out


@pytest.mark.parametrize("N,R", [(37, 30), (300, 10), (1, 10), (129, 7)])
def test_pack_unpack_round_trip_and_bounds(N, R):
    w = _weights()
    rng = np.random.default_rng(N)
    Vs = torch.tensor(rng.gamma(2.0, 0.05, size=(N, R, 520)).astype(np.float32)).to(DEV)
    vst, idx, nbytes = _pack(w, Vs, R, N)
    assert bool((vst[nbytes:] == SENT).all()) and bool((idx[N * 32:] == SENT).all()), "pack wrote past its buffers"
    out = _unpack(w, vst, idx, R, N)
    rel = ((out[:N, :, :513] - Vs[:, :, :513]) / Vs[:, :, :513]).abs().max().item()
    assert rel <= 2.0 ** -8 + 1e-6, rel                      # one BF16 rounding
    assert bool((out[N:] == -7.0).all()) and bool((out[:N, :, 513:] == -7.0).all()), "unpack wrote outside [NT][R][0..F)"
    # a slot index table that points every sample at slot 1 yields R copies of sample 0
    idx2 = idx.clone()
    idx2.view(-1, 32)[:N, :R] = 1
    out2 = _unpack(w, vst, idx2, R, N)
    assert torch.equal(out2[:N, :, :513], out[:N, :1, :513].expand(-1, R, -1))


@pytest.mark.parametrize("N,K,R,split,C,wpart",
                         [(n, k, r, sp, 1, wp) for wp in (False, True) for n, k, r, sp in
                          [(37, 10, 30, None), (185, 10, 10, None), (300, 7, 30, None), (130, 10, 10, [1, 100, 29]),
                           (400, 10, 10, [3, 120, 5, 1, 128, 143])]] +
                         # several chains per frame (R kept samples per chain): the windowed M-step kernel; 4 x 10 = 40 samples = one
                         # full window + 10, 16 x 10 = 160 = five windows + 10 (BASELINE configs[3]), 2 x 30 = two full windows
                         [(37, 10, 10, None, 4, True), (130, 10, 10, [1, 100, 29], 16, True), (70, 7, 30, [33, 37], 2, True),
                          (9, 10, 10, [4, 5], 128, True)])
def test_m_step_on_emission_matches_oracle_given_its_inputs(N, K, R, split, C, wpart):
    """dvae_vst_frame_stats / dvae_vst_w_partials + dvae_nmf_mstep_vst against the oracle's M-step on the variances the emission
    actually holds; with C chains per frame the frame's C x R samples are the oracle's sample axis."""
    F, ld = 513, 520
    w = _weights()
    rng = np.random.default_rng(F + N + R + C)
    P = rng.gamma(1.0, 1.0, size=(F, N)).astype(np.float32) * 0.1
    Rc, NC = R, N * C                                                            # kept per chain, chain rows
    R = C * Rc                                                                   # samples per frame
    Vs0 = rng.gamma(2.0, 0.05, size=(R, F, N)).astype(np.float32)
    W = np.maximum(rng.uniform(size=(F, K)), 1e-8).astype(np.float32)
    H = np.maximum(rng.uniform(size=(K, N)), 1e-8).astype(np.float32)
    g = rng.uniform(0.5, 1.5, size=N).astype(np.float32)
    Vsd = torch.zeros((N, R, ld), device=DEV)
    Vsd[:, :, :F] = torch.from_numpy(np.ascontiguousarray(Vs0.transpose(2, 0, 1))).to(DEV)
    vst, idx, _ = _pack(w, Vsd.view(NC, Rc, ld), Rc, NC)                          # chain row m = frame * C + chain
    # make the index table non-trivial: some samples repeat their predecessor's slot (a rejected proposal)
    it = idx.view(-1, 32)
    rep = torch.tensor(rng.uniform(size=(NC, Rc)) < 0.2).to(DEV)
    rep[:, 0] = False
    for r in range(1, Rc):
        it[:NC, r] = torch.where(rep[:, r], it[:NC, r - 1], it[:NC, r])
    Vq = _unpack(w, vst, idx, Rc, NC)[:NC, :, :F].reshape(N, R, F)                # what the kernels will see
    Vs = np.ascontiguousarray(Vq.permute(1, 2, 0).cpu().numpy())
    lens = split or [N]
    batch = RaggedBatch(lens, DEV)
    B = batch.B
    Pd = fm(P)
    Wd = torch.zeros((B, K, ld), device=DEV)
    Wd[:, :, :F] = torch.from_numpy(np.ascontiguousarray(W.T)).to(DEV)[None]
    Hd = torch.from_numpy(np.ascontiguousarray(H.T)).to(DEV).contiguous()
    gd = torch.from_numpy(g).to(DEV)
    Vbd = torch.zeros((N, ld), device=DEV)
    _lib.call("dvae_nmf_vb", _p(Wd), _p(Hd), _p(batch.frame_utt), N, F, K, ld, _p(Vbd), _stream())
    fstat = torch.zeros(2 * N * ld, device=DEV)
    img = tc.decoder_image(w)
    if C == 1:
        _lib.call("dvae_vst_frame_stats", w.dec.ref, _p(img), 16, 0, _p(vst), _p(idx), R, _p(Vbd), _p(gd), N, ld, _p(fstat),
                  _p(fstat[N * ld:]), _stream())
        Vx = torch.tensor(g)[None, None, :] * torch.tensor(Vs) + torch.tensor(W @ H)[None]
        a1 = (1.0 / Vx.double()).sum(0).numpy()
        a2 = (1.0 / Vx.double() ** 2).sum(0).numpy()
        assert relerr(unfm(fstat[: N * ld].view(N, ld), F), a1) <= 2e-5
        assert relerr(unfm(fstat[N * ld:].view(N, ld), F), a2) <= 2e-5
    cost = torch.zeros(B, dtype=torch.float64, device=DEV)
    st = torch.zeros(1, dtype=torch.int32, device=DEV)
    ws = torch.empty(int(_lib.load().dvae_nmf_workspace_floats(B, K, ld, batch.max_frames)), device=DEV)
    wp = utt_seg = None
    if wpart:                                    # the W sums reduced per (tile, utterance) segment inside the statistics pass
        seg_start, tile_seg, utt_seg, S = batch.segments(n_chains=C)
        assert S >= B and int(utt_seg[-1]) == S
        wp = torch.full((int(_lib.load().dvae_vst_w_partial_floats(S, K, ld)) + 64,), float("nan"), device=DEV)
        _lib.call("dvae_vst_w_partials", w.dec.ref, _p(img), 16, 0, _p(vst), _p(idx), Rc, _p(Pd), _p(Vbd), _p(gd), _p(Hd), K, N, C, ld,
                  _p(seg_start), _p(tile_seg), _p(wp), _stream())
        body = wp[:-64].view(S, K, 2, ld)
        assert bool(torch.isnan(wp[-64:]).all()) and bool(torch.isnan(body[..., F:]).all()), "partials written outside [S][K][2][0..F)"
        assert bool(torch.isfinite(body[..., :F]).all())
    _lib.call("dvae_nmf_mstep_vst", w.dec.ref, _p(img), 16, 0, _p(Pd), _p(vst), _p(idx), Rc, _p(Wd), _p(Hd), _p(gd), _p(Vbd), _p(cost),
              _p(batch.fr_off), B, N, K, ld, batch.max_frames, _p(ws), None if wpart else _p(fstat), _p(wp), _p(utt_seg), C, _p(st),
              _stream())
    assert int(st.item()) == 0
    off = batch.fr_off_host
    for u in range(B):
        a, b = int(off[u]), int(off[u + 1])
        ref = mcem_port.m_step_reference(torch.tensor(P[:, a:b]), torch.tensor(Vs[:, :, a:b]), torch.tensor(W),
                                         torch.tensor(H[:, a:b]), torch.tensor(g[a:b]))
        assert relerr(Wd[u, :, :F].t().cpu().numpy(), ref["W"].numpy()) <= 1e-4
        assert relerr(Hd[a:b].t().cpu().numpy(), ref["H"].numpy()) <= 1e-4
        assert relerr(gd[a:b].cpu().numpy(), ref["g"].numpy()) <= 1e-4
        assert relerr(unfm(Vbd[a:b], F), ref["Vb"].numpy()) <= 1e-4
        assert abs(cost[u].item() - ref["cost"].item()) <= 1e-4 * abs(ref["cost"].item())


def _engine(variant, keep, lens, seed=3, **kw):
    w = _weights(variant)
    NT = sum(lens)
    rng = np.random.default_rng(11)
    P = torch.tensor(rng.gamma(1.0, 0.05, size=(NT, 520)).astype(np.float32)).to(DEV)
    X = torch.zeros((NT, 520), dtype=torch.complex64, device=DEV)
    y = (torch.tensor(rng.uniform(size=(NT, 1))) > 0.5).float().to(DEV) if w.y_dim else None
    cfg = McemConfig(niter=1, keep_E=keep, burn_E=6, keep_WF=5, burn_WF=3, sampler="tc", seed=seed, **kw)
    eng = McemEngine(w, cfg, DEV)
    eng.init_parameters(X, P, RaggedBatch(lens, DEV, utt_ids=list(range(10, 10 + len(lens)))), y)
    return eng


@pytest.mark.parametrize("variant,keep", [("M1", 30), ("M2", 10), ("M2v3", 10)])
def test_emission_is_the_decoder_output_of_the_kept_samples(variant, keep):
    lens = [185, 64, 1, 37]
    eng = _engine(variant, keep, lens)
    NT = sum(lens)
    Zs = eng.sample_posterior(keep, 6, emit=True)
    tc.check_status(eng)
    idx = eng.vs_idx.view(-1, 32)[:NT, :keep].cpu().numpy().astype(int)
    # slot r+1 was the proposal of kept iteration r: a sample points at its own slot (accepted) or repeats its predecessor
    r = np.arange(keep)[None, :]
    prev = np.concatenate([np.zeros((NT, 1), int), idx[:, :-1]], axis=1)
    assert np.all((idx == r + 1) | (idx == prev))
    acc = (idx == r + 1).mean()
    assert 0.3 < acc < 0.999, acc
    # consecutive equal latent samples <=> same slot
    same_z = (Zs[:, 1:] == Zs[:, :-1]).all(dim=2).cpu().numpy()
    assert np.array_equal(same_z, idx[:, 1:] == idx[:, :-1])
    got = tc.vst_unpack(eng, keep)[:, :, :513]
    x2 = None if eng.y is None else eng.y
    ref = mlp_forward(eng.w.dec, Zs.reshape(NT * keep, 16), _lib.ACT_EXP, x2=x2, x2_row_div=keep).view(NT, keep, 513)
    rel = ((got - ref) / ref).abs()
    assert rel.max().item() <= 1.5e-2 and rel.mean().item() <= 3e-3, (rel.max().item(), rel.mean().item())


def test_emission_buffers_are_not_overrun():
    keep, lens = 10, [100, 29]                                  # 129 chains: two tiles, the second almost empty
    eng = _engine("M1", keep, lens)
    NT = sum(lens)
    nbytes = int(_lib.load().dvae_vst_bytes(NT, keep))
    big = torch.full((nbytes + 8192,), SENT, dtype=torch.uint8, device=DEV)
    big_idx = torch.full((NT * 32 + 4096,), SENT, dtype=torch.uint8, device=DEV)
    eng._buf[("VsT", (nbytes,), torch.uint8)] = big[:nbytes]                    # what McemEngine._get will hand out
    eng._buf[("vs_idx", (NT * 32,), torch.uint8)] = big_idx[: NT * 32]
    eng.sample_posterior(keep, 6, emit=True)
    tc.check_status(eng)
    assert eng.VsT.data_ptr() == big.data_ptr() and eng.vs_idx.data_ptr() == big_idx.data_ptr()
    assert bool((big[nbytes:] == SENT).all()), "the sampler wrote past the emission buffer"
    assert bool((big_idx[NT * 32:] == SENT).all()), "the sampler wrote past the slot index table"
    assert bool((big_idx[: NT * 32].view(-1, 32)[:, keep:] == SENT).all()), "only bytes 0..keep-1 of an index row are written"
    assert bool((big_idx[: NT * 32].view(-1, 32)[:, :keep] <= keep).all())


@pytest.mark.parametrize("variant,L", [("M1", 16), ("M2", 16), ("M1", 32)])
def test_philox_inside_the_sampler_equals_dumped_draws(variant, L):
    """Same seed and counters: the sampler drawing inside the kernel and the sampler fed with dvae_rng_dump's numbers walk
    bit-identical chains (kept samples, last states, accept counts, emitted variances and slot table)."""
    y_dim = 0 if variant == "M1" else 1
    sd = synth.xavier_state_dict(variant, 513, L, [128, 128] if L == 16 else [128], y_dim, seed=2, out_bias=float(np.log(0.05)))
    w = VaeWeights(sd, variant, torch.device(DEV))
    lens, keep, burn = [150, 40, 3], 10, 7
    NT = sum(lens)
    rng_np = np.random.default_rng(3)
    P = torch.tensor(rng_np.gamma(1.0, 0.05, size=(NT, 520)).astype(np.float32)).to(DEV)
    X = torch.zeros((NT, 520), dtype=torch.complex64, device=DEV)
    y = (torch.tensor(rng_np.uniform(size=(NT, 1))) > 0.5).float().to(DEV) if y_dim else None
    res = []
    for inject in (False, True):
        eng = McemEngine(w, McemConfig(niter=1, keep_E=keep, burn_E=burn, sampler="tc", seed=99), DEV)
        b = RaggedBatch(lens, DEV, utt_ids=[5, 1000, 77])
        eng.init_parameters(X, P, b, y)
        eng.mh_iter0 = 13
        draws = None
        if inject:
            rng = _lib.DvaeRng()
            rng.seed, rng.iter0 = 99, 13
            eps = torch.empty((keep + burn, NT, L), device=DEV)
            u = torch.empty((keep + burn, NT), device=DEV)
            _lib.call("dvae_rng_dump", C.byref(rng), _p(b.frame_gid), _p(b.frame_idx), NT, 1, L, keep + burn, _p(eps), _p(u), _stream())

            class D:
                W0 = H0 = None

                def mh_draws(self, call, n_iter, chains, L_):
                    return eps, u
            draws = D()
        Zs = eng.sample_posterior(keep, burn, draws, emit=True).clone()
        tc.check_status(eng)
        res.append((Zs, eng.Z.clone(), eng.n_accept.clone(), tc.vst_unpack(eng, keep), eng.vs_idx.clone().view(-1, 32)[:, :keep]))
    for a, b_ in zip(*res):
        assert torch.equal(a, b_)
    assert 0 < int(res[0][2].sum()) < NT * (keep + burn)


def test_status_word_reports_non_finite_values_and_is_cleared():
    eng = _engine("M1", 10, [40])
    eng.P[3, 17] = float("nan")
    eng.e_step()
    eng.m_step(0)
    with pytest.raises(_lib.DvaeError, match="NaN"):
        tc.check_status(eng)
    tc.check_status(eng)                                         # cleared by the raise: the engine is usable again


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_engine_and_shims_run_on_a_non_current_device():
    """The reference hands each worker a device index and never calls set_device (scripts/evaluate_ntcd_M1.py:69,252):
    everything must land on that device while cuda:0 stays current."""
    from dvae_b200.engine import Enhancer
    from dvae_b200.packages.models import mcem as shim_mcem
    from dvae_b200.packages.models import models as shim_models
    from dvae_b200.packages.processing.stft import istft, stft
    torch.cuda.set_device(0)
    x, s, _ = synth.synth_utterance(5, 1.0)
    sd = synth.xavier_state_dict("M1", 513, 16, [128, 128], 0, seed=3, out_bias=synth.speech_prior_bias(s))
    cfg = McemConfig(niter=3, keep_E=10, burn_E=5, keep_WF=5, burn_WF=5, seed=4, sampler="tc")
    out = []
    for dev in (0, 1):
        enh = Enhancer(sd, "M1", cfg, device=dev)
        s_hat, n_hat, cost = enh.enhance([x], utt_ids=[9])
        assert enh.engine.W.device.index == dev and torch.cuda.current_device() == 0
        out.append((s_hat[0].copy(), cost.copy()))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    model = shim_models.VariationalAutoencoder([513, 16, [128, 128]])
    model.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})
    model.to("cuda:1").eval()
    kw = dict(fs=16000, wlen_sec=64e-3, win="hann", hop_percent=0.25, center=False)
    algo = shim_mcem.MCEM_M1(3, 10, 5, 5, 5, 0.01, seed=4)
    Xtf = stft(x, pad_at_end=True, **kw)
    algo.init_parameters(X=Xtf, S=Xtf, vae=model, nmf_rank=10, eps=1e-8, device=1)
    algo.run()
    assert algo.W.device.index == 1 and torch.cuda.current_device() == 0
    assert algo.Vs.shape == (5, 513, Xtf.shape[1]) and algo.Vx.shape == algo.Vs.shape
    y = istft(algo.S_hat, max_len=len(x), **kw)
    assert np.isfinite(y).all()
