"""``stft`` / ``istft`` with the reference's signatures, computed by libdvae_b200's CUDA kernels.

Drop-in for ``packages/processing/stft.py:13-99`` (wrappers over ``librosa.core.stft`` / ``istft``).  numpy in, numpy
out, same shapes and dtypes: ``stft -> (F, N) complex64``, ``istft -> (T,) float32``.  The kernels implement the one
configuration every caller of the reference uses (``wlen_sec=64e-3`` at 16 kHz -> n_fft = 1024, ``hop_percent=0.25``,
Hann window, ``center=False``, ``pad_at_end=True``); anything else raises instead of silently falling back.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from ...engine import RaggedBatch, _require_cuda, istft_batch, stft_batch

_DEVICE = 0


def set_device(device):
    """CUDA device used by the numpy-facing wrappers (default 0)."""
    global _DEVICE
    _DEVICE = device


def _sizes(fs, wlen_sec, hop_percent, what):
    if wlen_sec * fs != int(wlen_sec * fs):
        raise ValueError("wlen_sample of %s is not an integer." % what)
    nfft = int(wlen_sec * fs)
    hopsamp = int(hop_percent * nfft)
    return nfft, hopsamp


def _check(nfft, win, center):
    if nfft != 1024:
        raise NotImplementedError("dvae_b200 implements n_fft=1024 only (got %d); the reference scripts use wlen_sec=64e-3 at 16 kHz" % nfft)
    if win != 'hann':
        raise NotImplementedError("dvae_b200 implements the Hann window only")
    if center:
        raise NotImplementedError("dvae_b200 implements center=False only (as every reference script sets)")


def stft(x, fs=16e3, wlen_sec=50e-3, win='hann', hop_percent=0.25, center=True, pad_mode='reflect', pad_at_end=True,
         dtype='complex64'):
    nfft, hopsamp = _sizes(fs, wlen_sec, hop_percent, "STFT")
    if not pad_at_end:
        # the reference leaves x_ unbound on this branch (stft.py:45-52)
        raise UnboundLocalError("local variable 'x_' referenced before assignment")
    _check(nfft, win, center)
    x = np.asarray(x)
    T = len(x)
    utt_len = T / fs
    q = utt_len / wlen_sec / hop_percent
    Tp = T + hopsamp if math.ceil(q) != int(q) else T
    if Tp < nfft:
        raise ValueError("input shorter than one frame")
    n_frames = 1 + (Tp - nfft) // hopsamp
    dev = _require_cuda(_DEVICE)
    xd = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(dev)
    batch = RaggedBatch([n_frames], dev)
    X, _ = stft_batch(xd, torch.zeros(1, dtype=torch.int64, device=dev), torch.tensor([T], dtype=torch.int32, device=dev),
                      batch, nfft, hopsamp, want_power=False)
    F = nfft // 2 + 1
    return np.ascontiguousarray(X[:, :F].t().cpu().numpy()).astype(dtype, copy=False)


def istft(Sxx, fs=16000, wlen_sec=50e-3, win='hann', hop_percent=0.25, center=True, dtype='float32', max_len=None):
    nfft, hopsamp = _sizes(fs, wlen_sec, hop_percent, "iSTFT")
    _check(nfft, win, center)
    Sxx = np.asarray(Sxx)
    F = nfft // 2 + 1
    if Sxx.ndim != 2 or Sxx.shape[0] != F:
        raise ValueError("spectrogram must be (%d, N)" % F)
    N = Sxx.shape[1]
    T = nfft + hopsamp * (N - 1) if max_len is None else int(max_len)      # max_len is in SAMPLES
    dev = _require_cuda(_DEVICE)
    ld = (F + 7) // 8 * 8
    X = torch.zeros((N, ld), dtype=torch.complex64, device=dev)
    X[:, :F] = torch.from_numpy(np.ascontiguousarray(Sxx.T.astype(np.complex64))).to(dev)
    batch = RaggedBatch([N], dev)
    y = istft_batch(X, batch, torch.zeros(1, dtype=torch.int64, device=dev), torch.tensor([T], dtype=torch.int32, device=dev),
                    T, T, nfft, hopsamp)
    out = y.cpu().numpy().astype(dtype, copy=False)
    if max_len:
        out = out[:int(max_len * fs)]
    return out
