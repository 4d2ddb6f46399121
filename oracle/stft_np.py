"""numpy restatement of the reference's STFT / ISTFT wrappers.  Oracle: test infrastructure only.

Follows ``packages/processing/stft.py:13-60`` (``stft``) and ``:63-99`` (``istft``), which are
thin wrappers over ``librosa.core.stft`` / ``librosa.core.istft`` (librosa 0.7-0.9 semantics,
restated here because librosa is not installed and the reference holds no vector at this boundary:
**not pinned by a reference-held vector**, see ``oracle/__init__.py``; pinned against three independent
implementations instead: ``torch.stft`` in float64 and a direct DFT (forward), ``scipy.signal.stft`` /
``scipy.signal.istft`` with the same framing (forward bit for bit after the complex64 cast, inverse to
float32 rounding on plain and on masked spectrograms), ``tests/test_oracle_stft.py``).

librosa semantics restated (for ``center=False``; ``center=True`` adds the reflect padding):
  stft : periodic Hann ``get_window('hann', n_fft, fftbins=True)`` in float64, frames
         ``y[i*hop : i*hop + n_fft]``, ``N = 1 + (len - n_fft)//hop``, ``rfft`` over each windowed
         frame, result stored as complex64 with shape ``(1 + n_fft/2, N)``.
  istft: per frame ``window * irfft(column)``, overlap-added into a float32 buffer of
         ``n_fft + hop*(N-1)`` samples, then divided by the float32 sum of squared windows
         wherever that sum exceeds ``np.finfo(float32).tiny``; no centre trimming; finally
         ``fix_length`` (crop or zero-pad) to ``length``.
"""
from __future__ import annotations

import math

import numpy as np


def hann_periodic(n_fft: int) -> np.ndarray:
    """``scipy.signal.get_window('hann', n_fft, fftbins=True)`` in float64."""
    k = np.arange(n_fft, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / n_fft)


def _sizes(fs, wlen_sec, hop_percent, what):
    # stft.py:34-37 / 82-85
    if wlen_sec * fs != int(wlen_sec * fs):
        raise ValueError("wlen_sample of %s is not an integer." % what)
    n_fft = int(wlen_sec * fs)
    hop = int(hop_percent * n_fft)
    return n_fft, hop


def end_pad(x: np.ndarray, fs, wlen_sec, hop_percent, hop: int) -> np.ndarray:
    """End-padding rule of ``stft.py:45-50``: pad ``hop`` zeros unless len/fs/wlen/hop% is an integer."""
    utt_len = len(x) / fs
    q = utt_len / wlen_sec / hop_percent
    if math.ceil(q) != int(q):
        return np.pad(x, (0, hop), mode="constant")
    return x


def stft(x, fs=16e3, wlen_sec=50e-3, win="hann", hop_percent=0.25, center=True, pad_mode="reflect",
         pad_at_end=True, dtype="complex64"):
    """Restates ``packages/processing/stft.py:13-60`` -> ``(F, N)`` complex64."""
    n_fft, hop = _sizes(fs, wlen_sec, hop_percent, "STFT")
    if win != "hann":
        raise ValueError("oracle restates the Hann window only (every reference caller uses it)")
    x = np.asarray(x)
    if not pad_at_end:
        # stft.py:45-52 leaves x_ undefined in this case: the reference raises.
        raise UnboundLocalError("local variable 'x_' referenced before assignment")
    x_ = end_pad(x, fs, wlen_sec, hop_percent, hop)
    y = np.asarray(x_, dtype=np.float64)
    if center:
        y = np.pad(y, n_fft // 2, mode=pad_mode)
    if len(y) < n_fft:
        raise ValueError("input shorter than one frame")
    n_frames = 1 + (len(y) - n_fft) // hop
    idx = np.arange(n_fft)[:, None] + hop * np.arange(n_frames)[None, :]
    frames = y[idx]                                   # (n_fft, N), librosa.util.frame layout
    spec = np.fft.rfft(hann_periodic(n_fft)[:, None] * frames, axis=0)
    return spec.astype(dtype)


def window_sumsquare(n_frames: int, n_fft: int, hop: int) -> np.ndarray:
    """librosa.filters.window_sumsquare(norm=None, dtype=float32): float32 accumulation frame by frame."""
    n = n_fft + hop * (n_frames - 1)
    out = np.zeros(n, dtype=np.float32)
    win_sq = (hann_periodic(n_fft) ** 2).astype(np.float64)
    for i in range(n_frames):
        s = i * hop
        out[s:s + n_fft] += win_sq[: max(0, min(n_fft, n - s))]
    return out


def istft(Sxx, fs=16000, wlen_sec=50e-3, win="hann", hop_percent=0.25, center=True, dtype="float32",
          max_len=None):
    """Restates ``packages/processing/stft.py:63-99`` -> ``(T,)`` float32.

    ``max_len`` is forwarded as librosa's ``length`` in SAMPLES (stft.py:95); the later
    ``x[:int(max_len*fs)]`` (stft.py:97-98) is then a no-op slice.
    """
    n_fft, hop = _sizes(fs, wlen_sec, hop_percent, "iSTFT")
    if win != "hann":
        raise ValueError("oracle restates the Hann window only")
    Sxx = np.asarray(Sxx)
    if 2 * (Sxx.shape[0] - 1) != n_fft:
        raise ValueError("spectrogram has %d bins, expected %d" % (Sxx.shape[0], n_fft // 2 + 1))
    n_frames = Sxx.shape[1]
    w = hann_periodic(n_fft)
    y = np.zeros(n_fft + hop * (n_frames - 1), dtype=dtype)
    ytmp = w[:, None] * np.fft.irfft(Sxx.astype(np.complex128), n=n_fft, axis=0)
    for i in range(n_frames):
        y[i * hop:i * hop + n_fft] += ytmp[:, i]      # float32 accumulation, frame order
    wss = window_sumsquare(n_frames, n_fft, hop)
    nz = wss > np.finfo(np.float32).tiny
    y[nz] /= wss[nz]
    if max_len is None:
        if center:
            y = y[n_fft // 2: -(n_fft // 2)]
    else:
        start = n_fft // 2 if center else 0
        y = y[start:]
        if len(y) > max_len:
            y = y[:max_len]
        elif len(y) < max_len:
            y = np.pad(y, (0, max_len - len(y)), mode="constant")
    if max_len:
        y = y[:int(max_len * fs)]
    return y


def dft_direct(frame: np.ndarray) -> np.ndarray:
    """O(n^2) one-sided DFT of one real frame in float64; independent check for the rfft path."""
    n = len(frame)
    k = np.arange(n // 2 + 1)[:, None]
    t = np.arange(n)[None, :]
    return (frame[None, :] * np.exp(-2j * np.pi * k * t / n)).sum(axis=1)


# ---------------------------------------------------------------------------------------------- labels (target.py)
def clean_speech_vad(speech_t, fs=16000, wlen_sec=64e-3, hop_percent=0.25, pad_at_end=True, vad_threshold=1.70):
    """``clean_speech_VAD`` (packages/processing/target.py:5-56) with ``center=False``; ``librosa.util.frame`` restated as
    ``y[j*hop : j*hop + n_fft]`` for ``j < 1 + (len - n_fft) // hop``."""
    n_fft = int(wlen_sec * fs)
    hop = int(hop_percent * n_fft)
    y = np.asarray(speech_t, np.float64)
    if pad_at_end:
        utt_len = len(y) / fs
        if math.ceil(utt_len / wlen_sec / hop_percent) != int(utt_len / wlen_sec / hop_percent):
            y = np.pad(y, (0, hop), mode="constant")
    n = 1 + (len(y) - n_fft) // hop
    frames = y[np.arange(n_fft)[:, None] + hop * np.arange(n)[None, :]]
    power = np.power(frames, 2).sum(axis=0)
    return np.float32(power > np.power(10, vad_threshold) * np.min(power))[None]


def clean_speech_ibm(speech_tf, eps=1e-8, ibm_threshold=50):
    """``clean_speech_IBM`` (packages/processing/target.py:58-70)."""
    power_db = 20 * np.log10(abs(speech_tf) + eps)
    return np.float32(power_db > np.max(power_db) - ibm_threshold)

