"""CUDA STFT / ISTFT (through the C ABI and the reference-signature wrappers) against the numpy oracle."""
import numpy as np
import pytest
import torch

from dvae_b200 import synth
from oracle import stft_np
from tests.gpu_util import DEV, relerr

pytestmark = pytest.mark.gpu
KW = dict(fs=16000, wlen_sec=64e-3, win="hann", hop_percent=0.25, center=False, pad_at_end=True)
IKW = dict(fs=16000, wlen_sec=64e-3, win="hann", hop_percent=0.25, center=False)


def test_stft_wrapper_matches_oracle():
    from dvae_b200.packages.processing.stft import stft
    for u, secs in ((0, 3.0), (1, 1.0), (2, 2.048), (3, 0.5)):
        x, _, _ = synth.synth_utterance(u, secs)
        ref = stft_np.stft(x, **KW)
        got = stft(x, **KW)
        assert got.dtype == np.complex64 and got.shape == ref.shape
        assert relerr(got, ref) <= 1e-4          # north-star tolerance for STFT (FP32 FFT vs FP64 reference)
        assert relerr(got, ref) <= 2e-6          # what an FP32 radix-8 FFT actually delivers


def test_stft_batch_ragged_and_power():
    from dvae_b200.engine import RaggedBatch, stft_batch
    lens = [48000, 16000, 20000, 1024, 33333]
    xs = [synth.synth_utterance(10 + i, l / 16000.0)[0][:l] for i, l in enumerate(lens)]
    off = np.zeros(len(lens) + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    flat = torch.from_numpy(np.concatenate(xs)).to(DEV)
    nfr = [synth.num_frames(l) for l in lens]
    batch = RaggedBatch(nfr, DEV)
    X, P = stft_batch(flat, torch.from_numpy(off[:-1].copy()).to(DEV), torch.tensor(lens, dtype=torch.int32, device=DEV), batch)
    X, P = X.cpu().numpy(), P.cpu().numpy()
    for u, x in enumerate(xs):
        ref = stft_np.stft(x, **KW)
        a, b = batch.fr_off_host[u], batch.fr_off_host[u + 1]
        assert b - a == ref.shape[1]
        assert relerr(X[a:b, :513].T, ref) <= 2e-6
        assert relerr(P[a:b, :513].T, np.abs(ref) ** 2) <= 1e-5


def test_istft_wrapper_matches_oracle():
    from dvae_b200.packages.processing.stft import istft
    rng = np.random.default_rng(0)
    for u, secs in ((4, 3.0), (5, 1.0)):
        x, _, _ = synth.synth_utterance(u, secs)
        X = stft_np.stft(x, **KW)
        X = (X * rng.uniform(0.0, 1.0, size=X.shape)).astype(np.complex64)      # a Wiener-like real mask
        ref = stft_np.istft(X, max_len=len(x), **IKW)
        got = istft(X, max_len=len(x), **IKW)
        assert got.dtype == np.float32 and got.shape == ref.shape
        # interior: the reference's own metric drops 50 ms at both ends (scripts/run_metrics.py:117-121)
        assert relerr(got[800:-800], ref[800:-800]) <= 1e-4
        assert relerr(got[800:-800], ref[800:-800]) <= 5e-6
        # edges: sum of squared windows is tiny there, errors are amplified by up to 1/w^2 -> absolute bound
        # (sample 1: w^2 = 9e-11, the FP32 irfft rounding is amplified ~1e5 times in BOTH implementations)
        scale = np.max(np.abs(ref[800:-800]))
        tol = 2e-3 * np.abs(ref[:800]) + 1e-4 * scale
        assert np.all(np.abs(got[:800] - ref[:800]) <= tol)
        assert np.all(np.abs(got[-800:] - ref[-800:]) <= 2e-3 * np.abs(ref[-800:]) + 1e-4 * scale)
        assert got[0] == 0.0


def test_istft_truncated_frames_zero_pad():
    from dvae_b200.packages.processing.stft import istft
    x, _, _ = synth.synth_utterance(6)
    X = stft_np.stft(x, **KW)[:, :100]
    ref = stft_np.istft(X, max_len=len(x), **IKW)
    got = istft(X, max_len=len(x), **IKW)
    assert len(got) == len(x)
    assert np.all(got[1024 + 256 * 99:] == 0)
    assert relerr(got[800:20000], ref[800:20000]) <= 5e-6
    got2 = istft(X, **IKW)                                  # max_len=None: natural length
    assert len(got2) == 1024 + 256 * 99


def test_round_trip_full_batch_property():
    """Size-independent property at the benchmark's full batch: STFT -> ISTFT reproduces the interior of every utterance."""
    from dvae_b200.engine import RaggedBatch, istft_batch, stft_batch
    B, T = 512, 48000
    g = torch.Generator(device=DEV).manual_seed(1)
    x = (torch.rand((B, T), device=DEV, generator=g) - 0.5)
    off = torch.arange(B, device=DEV, dtype=torch.int64) * T
    lens = torch.full((B,), T, dtype=torch.int32, device=DEV)
    batch = RaggedBatch([synth.num_frames(T)] * B, DEV)
    X, _ = stft_batch(x.view(-1), off, lens, batch)
    y = istft_batch(X, batch, off, lens, B * T, T).view(B, T)
    err = (y[:, 800:-800] - x[:, 800:-800]).abs().max().item()
    assert err <= 5e-6


def test_errors():
    from dvae_b200.packages.processing.stft import istft, stft
    with pytest.raises(ValueError):
        stft(np.zeros(4000, np.float32), fs=16000, wlen_sec=50.01e-3)
    with pytest.raises(NotImplementedError):
        stft(np.zeros(4000, np.float32), fs=16000, wlen_sec=50e-3, center=False)
    with pytest.raises(NotImplementedError):
        stft(np.zeros(4000, np.float32), fs=16000, wlen_sec=64e-3, center=True)
    with pytest.raises(ValueError):
        istft(np.zeros((513, 4), np.complex64), fs=16000, wlen_sec=50.01e-3)


def test_torch_front_end_matches_numpy_wrappers():
    """stft_pytorch / istft_pytorch (stft.py:102-193): tensors in and out, (F, N, 2) layout, max_len in seconds."""
    import torch
    from dvae_b200.packages.processing.stft import istft_pytorch, stft, stft_pytorch
    kw = dict(fs=16000, wlen_sec=64e-3, win="hann", hop_percent=0.25, center=False, pad_at_end=True)
    x = synth.synth_utterance(3, 1.0)[0].astype(np.float32)
    ref = stft(x, **kw)
    for dev in ("cuda:0", "cpu"):
        S = stft_pytorch(torch.from_numpy(x).to(dev), **kw)
        assert S.shape == (513, ref.shape[1], 2) and S.dtype == torch.float32 and S.device.type == dev[:4].rstrip(":")
        assert np.array_equal(torch.view_as_complex(S.contiguous()).cpu().numpy(), ref)
        y = istft_pytorch(S, fs=16000, wlen_sec=64e-3, win="hann", hop_percent=0.25, center=False, max_len=0.5)
        assert y.shape == (8000,) and y.dtype == torch.float32
        assert np.max(np.abs(y.cpu().numpy()[800:] - x[800:8000])) < 1e-5
    power = (stft_pytorch(torch.from_numpy(x).cuda(), **kw) ** 2).sum(-1)          # packages/data_handling.py:133-136
    assert np.allclose(power.cpu().numpy(), np.abs(ref) ** 2, rtol=1e-6)
    with pytest.raises(NotImplementedError):
        stft_pytorch(torch.zeros(4096), fs=16000, wlen_sec=64e-3, center=True)


def test_istft_masked_equals_wiener_apply_then_istft():
    """Fused Wiener mask + ISTFT (dvae_istft_masked_f32) against the two-kernel path and against the oracle."""
    from dvae_b200 import _lib
    from dvae_b200.engine import RaggedBatch, _p, _stream, istft_batch, istft_masked_batch, stft_batch
    lens = [48000, 16000, 20000, 1024, 33333]
    xs = [synth.synth_utterance(30 + i, l / 16000.0)[0][:l] for i, l in enumerate(lens)]
    off = np.zeros(len(lens) + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    flat = torch.from_numpy(np.concatenate(xs)).to(DEV)
    nfr = [synth.num_frames(l) for l in lens]
    batch = RaggedBatch(nfr, DEV)
    x_off = torch.from_numpy(off[:-1].copy()).to(DEV)
    x_len = torch.tensor(lens, dtype=torch.int32, device=DEV)
    X, _ = stft_batch(flat, x_off, x_len, batch)
    g = torch.Generator(device=DEV).manual_seed(3)
    R = 25
    ws = torch.rand(X.shape, device=DEV, generator=g) * R          # un-normalised sums over R samples, like WFs / WFn
    wn = R - ws
    S, Nn = torch.empty_like(X), torch.empty_like(X)
    _lib.call("dvae_wiener_apply", _p(X), _p(ws), _p(wn), R, batch.NT, 513, X.shape[1], _p(S), _p(Nn), _stream())
    total = int(off[-1])
    for mask, spec in ((ws, S), (wn, Nn)):
        two = istft_batch(spec, batch, x_off, x_len, total, max(lens))
        one = istft_masked_batch(X, mask, batch, x_off, x_len, total, max(lens), mask_scale=1.0 / R)
        assert torch.equal(one, two)                               # same products, same transform: bit-identical
    # oracle: mask the spectrum on the host, invert with the numpy restatement
    Xh, wh = X.cpu().numpy(), (ws / R).cpu().numpy()
    one = istft_masked_batch(X, ws, batch, x_off, x_len, total, max(lens), mask_scale=1.0 / R).cpu().numpy()
    for u, l in enumerate(lens):
        a, b = batch.fr_off_host[u], batch.fr_off_host[u + 1]
        ref = stft_np.istft((Xh[a:b, :513] * wh[a:b, :513]).T.astype(np.complex64), max_len=l, **IKW)
        got = one[off[u]:off[u] + l]
        if l > 4000:
            assert relerr(got[800:-800], ref[800:-800]) <= 1e-4     # north-star tolerance, interior samples
    with pytest.raises(ValueError):
        istft_masked_batch(X, ws[:, :512], batch, x_off, x_len, total, max(lens))


def test_cuda_stft_istft_against_scipy_signal_directly():
    """The CUDA kernels against scipy.signal's STFT / ISTFT (no restatement in between): STFT to FP32-FFT accuracy, ISTFT of a
    Wiener-masked spectrogram in the region where scipy and librosa share semantics (away from the first / last window)."""
    import warnings

    import scipy.signal
    from dvae_b200.packages.processing.stft import istft, stft
    x, _, _ = synth.synth_utterance(8, 3.0)
    xp = np.pad(x.astype(np.float64), (0, 256))
    w = scipy.signal.get_window("hann", 1024)
    _, _, Z = scipy.signal.stft(xp, fs=16000, window="hann", nperseg=1024, noverlap=768, nfft=1024, boundary=None, padded=False,
                                return_onesided=True, scaling="spectrum")
    got = stft(x, **KW)
    assert got.shape == Z.shape and relerr(got, Z * w.sum()) <= 2e-6
    M = np.random.default_rng(1).uniform(0.0, 1.0, size=Z.shape)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, ym = scipy.signal.istft(M * Z, fs=16000, window="hann", nperseg=1024, noverlap=768, nfft=1024, boundary=False, scaling="spectrum")
    y = istft((M * Z * w.sum()).astype(np.complex64), max_len=len(xp), **IKW)
    n = min(len(y), len(ym))
    assert np.max(np.abs(y[1024:n - 1024] - ym[1024:n - 1024])) <= 5e-6 * max(1.0, np.max(np.abs(ym)))
