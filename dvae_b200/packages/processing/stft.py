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

_DEVICE = None


def set_device(device):
    """CUDA device used by the numpy-facing wrappers.  Default: the calling thread's current device; the ``MCEM_*`` shims
    set it to their own device in ``init_parameters``, so a worker that was handed a device index (the reference's
    ``process_sublist``) keeps its whole utterance loop on that GPU."""
    global _DEVICE
    _DEVICE = device


def _device():
    return torch.cuda.current_device() if (_DEVICE is None and torch.cuda.is_available()) else (0 if _DEVICE is None else _DEVICE)


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
    dev = _require_cuda(_device())
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
    dev = _require_cuda(_device())
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


# ---------------------------------------------------------------------------------------------- torch front end (N2)
def stft_pytorch(x, fs=16e3, wlen_sec=50e-3, win='hann', hop_percent=0.25, center=True, pad_mode='reflect', pad_at_end=True):
    """``stft_pytorch`` of the reference (stft.py:102-151) on the CUDA kernel: a 1-D float tensor in, ``(F, N, 2)`` float32
    (real, imaginary) out, on the tensor's device - the layout ``torch.stft`` returned when the reference was written
    (``packages/data_handling.py:126-139`` squares and adds the last axis).  The reference's own version no longer runs on
    torch >= 2 (``return_complex`` is mandatory); CPU tensors are moved to the wrapper device and the result moved back.
    """
    nfft, hopsamp = _sizes(fs, wlen_sec, hop_percent, "STFT")
    if not pad_at_end:
        raise UnboundLocalError("local variable 'x_' referenced before assignment")       # as the reference (stft.py:133-139)
    _check(nfft, win, center)
    if x.dim() != 1:
        raise ValueError("stft_pytorch takes a 1-D signal")
    T = int(x.numel())
    q = (T / fs) / wlen_sec / hop_percent
    Tp = T + hopsamp if math.ceil(q) != int(q) else T
    if Tp < nfft:
        raise ValueError("input shorter than one frame")
    n_frames = 1 + (Tp - nfft) // hopsamp
    dev = x.device if x.is_cuda else _require_cuda(_device())
    xd = x.detach().to(device=dev, dtype=torch.float32).contiguous()
    with torch.cuda.device(dev):
        X, _ = stft_batch(xd, torch.zeros(1, dtype=torch.int64, device=dev), torch.tensor([T], dtype=torch.int32, device=dev),
                          RaggedBatch([n_frames], dev), nfft, hopsamp, want_power=False)
    out = torch.view_as_real(X[:, :nfft // 2 + 1].t().contiguous())
    return out if x.is_cuda else out.to(x.device)


def istft_pytorch(Sxx, fs=16000, wlen_sec=50e-3, win='hann', hop_percent=0.25, center=True, max_len=None):
    """``istft_pytorch`` of the reference (stft.py:153-193): ``(F, N, 2)`` real view or ``(F, N)`` complex tensor in, 1-D
    float32 tensor out; ``max_len`` is in SECONDS here (``x[:int(max_len*fs)]``, stft.py:191-192)."""
    nfft, hopsamp = _sizes(fs, wlen_sec, hop_percent, "iSTFT")
    _check(nfft, win, center)
    if not torch.is_complex(Sxx):
        if Sxx.dim() != 3 or Sxx.shape[-1] != 2:
            raise ValueError("spectrogram must be (F, N, 2) or complex (F, N)")
        Sxx = torch.view_as_complex(Sxx.contiguous())
    F = nfft // 2 + 1
    if Sxx.dim() != 2 or Sxx.shape[0] != F:
        raise ValueError("spectrogram must be (%d, N)" % F)
    N = int(Sxx.shape[1])
    T = nfft + hopsamp * (N - 1)
    dev = Sxx.device if Sxx.is_cuda else _require_cuda(_device())
    ld = (F + 7) // 8 * 8
    X = torch.zeros((N, ld), dtype=torch.complex64, device=dev)
    X[:, :F] = Sxx.detach().to(device=dev, dtype=torch.complex64).t()
    with torch.cuda.device(dev):
        y = istft_batch(X, RaggedBatch([N], dev), torch.zeros(1, dtype=torch.int64, device=dev),
                        torch.tensor([T], dtype=torch.int32, device=dev), T, T, nfft, hopsamp)
    if max_len:
        y = y[:int(max_len * fs)]
    return y if Sxx.is_cuda else y.to(Sxx.device)
